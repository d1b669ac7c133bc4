#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi -L > $O/r2n8_smi.txt 2>&1
timeout 1500 python bench.py --gpus 8 --steps 20 --warmup 5 > $O/r2n8_bench.json 2> $O/r2n8_bench.err
echo "bench rc=$?" >> $O/r2n8_bench.err
timeout 600 python bench.py --gpus 4 --steps 20 --warmup 5 --extras none > $O/r2n4_bench.json 2> $O/r2n4_bench.err
echo "bench rc=$?" >> $O/r2n4_bench.err
timeout 600 python bench.py --gpus 8 --steps 20 --warmup 5 --extras none --no-parity --profile-ranks > $O/r2n8_bench_prof.json 2> $O/r2n8_bench_prof.err
echo "bench rc=$?" >> $O/r2n8_bench_prof.err
