#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out
NCU="ncu --clock-control none"
export BA_PAIRS_REG=1
python tools/prof_case.py --cams 1000 --points 100000 --vis 0.1 --solves 1 > $O/r2y_prof_pairs_plain.log 2>&1 &&
$NCU --set full --import-source on -k regex:schur_pairs -c 1 -f -o $O/r2y_pairs_reg2 python tools/prof_case.py --cams 1000 --points 100000 --vis 0.1 --solves 1 > $O/r2y_prof_pairs_ncu.log 2>&1
echo "ncu rc=$?" >> $O/r2y_prof_pairs_ncu.log
