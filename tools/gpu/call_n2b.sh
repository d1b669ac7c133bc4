#!/bin/bash
# two GPUs: the peer-memory exchange tests over NVLink, strong-scaled c3 / c4 with the in-run parity gate
set -u
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi -L > $O/r2n2_smi.txt 2>&1
( time timeout 900 python -m pytest tests/test_gpu_sharded.py -m gpu -q --timeout 600 ) > $O/r2n2_pytest_sharded.log 2>&1
echo "pytest rc=$?" >> $O/r2n2_pytest_sharded.log
timeout 900 python bench.py --gpus 2 --steps 20 --warmup 5 > $O/r2n2_bench_c3.json 2> $O/r2n2_bench_c3.err
echo "bench rc=$?" >> $O/r2n2_bench_c3.err
timeout 900 python bench.py --gpus 2 --steps 20 --warmup 5 --profile-ranks --no-parity --extras none > $O/r2n2_bench_c3_prof.json 2> $O/r2n2_bench_c3_prof.err
timeout 900 python bench.py --gpus 2 --workload c4 --extras none --steps 10 --warmup 3 > $O/r2n2_bench_c4.json 2> $O/r2n2_bench_c4.err
echo "bench rc=$?" >> $O/r2n2_bench_c4.err
