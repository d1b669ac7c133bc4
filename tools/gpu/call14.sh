#!/bin/bash
# ncu evidence of the round-2 kernels (one tool per call: ncu only)
set -u
mkdir -p gpurun_out
O=gpurun_out
NCU="ncu --clock-control none"
python tools/prof_case.py --cams 200 --points 100000 --solves 2 > $O/r2_prof_c3_plain.log 2>&1 &&
$NCU --set full --import-source on -k regex:syrk_tma -s 1 -c 1 -f -o $O/r2_syrk_tma_c3 python tools/prof_case.py --cams 200 --points 100000 --solves 2 > $O/r2_prof_c3_ncu.log 2>&1
echo "c3 ncu rc=$?" >> $O/r2_prof_c3_ncu.log
python tools/prof_case.py --cams 50 --points 10000 --solves 2 > $O/r2_prof_c2_plain.log 2>&1 &&
$NCU --set full --import-source on -k "regex:syrk_tma|chol_step|syrk_reduce" -s 9 -c 10 -f -o $O/r2_kernels_c2 python tools/prof_case.py --cams 50 --points 10000 --solves 2 > $O/r2_prof_c2_ncu.log 2>&1
echo "c2 ncu rc=$?" >> $O/r2_prof_c2_ncu.log
python bench.py --steps 2 --warmup 1 --no-cpu-baseline --extras none > $O/r2_bench_short_plain.json 2> $O/r2_bench_short_plain.err &&
$NCU --metrics gpu__time_duration.sum -c 600 --csv --log-file $O/r2_launches_c3_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --extras none > $O/r2_bench_short_ncu.json 2> $O/r2_bench_short_ncu.err
echo "launch list rc=$?" >> $O/r2_bench_short_ncu.err
ls -la $O | grep r2_ 
