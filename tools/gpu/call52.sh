#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 ) > $O/r2p_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> $O/r2p_pytest_gpu.log
timeout 900 python bench.py --workload c4 --extras none --steps 10 --warmup 3 > $O/r2p_bench_c4.json 2> $O/r2p_bench_c4.err
echo "bench rc=$?" >> $O/r2p_bench_c4.err
