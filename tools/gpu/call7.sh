#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out
( time timeout 900 python -m pytest tests/test_projective_depth.py tests/test_projection.py -m gpu -q --timeout 600 ) > $O/r2g_pytest_depth.log 2>&1
echo "pytest rc=$?" >> $O/r2g_pytest_depth.log
timeout 600 python bench.py --steps 20 --warmup 5 > $O/r2g_bench_c3.json 2> $O/r2g_bench_c3.err
echo "bench rc=$?" >> $O/r2g_bench_c3.err
