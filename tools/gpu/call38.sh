#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 ) > $O/r2g_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> $O/r2g_pytest_gpu.log
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/r2g_bench_c3.json 2> $O/r2g_bench_c3.err
echo "bench rc=$?" >> $O/r2g_bench_c3.err
BA_NO_STAGED_UPLOAD=1 timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --extras none > $O/r2g_bench_c3_plainup.json 2> $O/r2g_bench_c3_plainup.err
