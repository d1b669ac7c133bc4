#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 600 -k "matrix_free" ) > $O/r2ae_pytest_mf.log 2>&1
echo "pytest rc=$?" >> $O/r2ae_pytest_mf.log
timeout 900 python bench.py --steps 20 --warmup 5 > $O/r2ae_bench_c3.json 2> $O/r2ae_bench_c3.err
echo "bench rc=$?" >> $O/r2ae_bench_c3.err
