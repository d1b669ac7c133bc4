#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python bench.py --gpus 8 --steps 20 --warmup 5 > $O/r2n8_bench.json 2> $O/r2n8_bench.err
echo "bench rc=$?" >> $O/r2n8_bench.err
