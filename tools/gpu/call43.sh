#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out
nproc > $O/r2j_nproc.txt
( time timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 600 -k "c3_properties or camera_major or matrix_free" ) > $O/r2j_pytest.log 2>&1
echo "pytest rc=$?" >> $O/r2j_pytest.log
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --extras none > $O/r2j_bench_c3.json 2> $O/r2j_bench_c3.err
echo "bench rc=$?" >> $O/r2j_bench_c3.err
