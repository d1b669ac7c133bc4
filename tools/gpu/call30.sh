#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out
NCU="ncu --clock-control none"
python tools/prof_case.py --cams 200 --points 20000 --solves 1 > $O/r2ac_prof_plain.log 2>&1 &&
$NCU --set full --import-source on -k regex:chol_fused -c 1 -f -o $O/r2ac_chol_fused python tools/prof_case.py --cams 200 --points 20000 --solves 1 > $O/r2ac_prof_ncu.log 2>&1
echo "ncu rc=$?" >> $O/r2ac_prof_ncu.log
