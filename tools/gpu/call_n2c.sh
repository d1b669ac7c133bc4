#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_sharded.py -m gpu -q --timeout 600 ) > $O/r2n2_pytest_sharded.log 2>&1
echo "pytest rc=$?" >> $O/r2n2_pytest_sharded.log
timeout 900 python bench.py --gpus 2 --workload c4 --extras none --steps 10 --warmup 3 > $O/r2n2_bench_c4.json 2> $O/r2n2_bench_c4.err
echo "bench rc=$?" >> $O/r2n2_bench_c4.err
