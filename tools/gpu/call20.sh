#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out
NCU="ncu --clock-control none"
python tools/prof_case.py --cams 1000 --points 100000 --vis 0.1 --solves 1 > $O/r2t_prof_pairs_plain.log 2>&1 &&
$NCU --set full --import-source on -k regex:schur_pairs -c 1 -f -o $O/r2t_pairs_reg python tools/prof_case.py --cams 1000 --points 100000 --vis 0.1 --solves 1 > $O/r2t_prof_pairs_ncu.log 2>&1
echo "ncu rc=$?" >> $O/r2t_prof_pairs_ncu.log
