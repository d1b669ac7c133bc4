#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python bench.py --steps 20 --warmup 5 > $O/r2f_bench_c3.json 2> $O/r2f_bench_c3.err
echo "bench rc=$?" >> $O/r2f_bench_c3.err
python bench.py --steps 2 --warmup 1 --no-cpu-baseline --extras none > $O/r2f_bench_short_plain.json 2> $O/r2f_bench_short_plain.err &&
ncu --clock-control none --metrics gpu__time_duration.sum -c 600 --csv --log-file $O/r2f_launches_c3_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --extras none > $O/r2f_bench_short_ncu.json 2> $O/r2f_bench_short_ncu.err
