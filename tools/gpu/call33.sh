#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out
python tools/prof_case.py --cams 200 --points 100000 --lm 2 > $O/r2af_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/r2af_launches_c3.csv python tools/prof_case.py --cams 200 --points 100000 --lm 2 > $O/r2af_ncu.log 2>&1
echo "rc=$?" >> $O/r2af_ncu.log
