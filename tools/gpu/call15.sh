#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 ) > $O/r2o_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> $O/r2o_pytest_gpu.log
: > $O/r2o_syrk_sweep.jsonl
for m in 199 200 50; do
  pts=100000; if [ $m = 50 ]; then pts=10000; fi
  timeout 300 python tools/syrk_sweep.py --cams $m --points $pts --tag box132 >> $O/r2o_syrk_sweep.jsonl 2>> $O/r2o_syrk_sweep.err
done
timeout 300 python tools/syrk_sweep.py --cams 200 --points 12500 --tag box132_n8shard >> $O/r2o_syrk_sweep.jsonl 2>> $O/r2o_syrk_sweep.err
timeout 600 python bench.py --steps 20 --warmup 5 > $O/r2o_bench_c3.json 2> $O/r2o_bench_c3.err
echo "bench rc=$?" >> $O/r2o_bench_c3.err
