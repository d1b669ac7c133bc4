#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --extras none > $O/r2g_bench_c3.json 2> $O/r2g_bench_c3.err
echo "bench rc=$?" >> $O/r2g_bench_c3.err
