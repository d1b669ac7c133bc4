#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 ) > $O/r2p_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> $O/r2p_pytest_gpu.log
: > $O/r2p_syrk_sweep.jsonl
for m in 199 200 50; do
  pts=100000; if [ $m = 50 ]; then pts=10000; fi
  timeout 300 python tools/syrk_sweep.py --cams $m --points $pts --tag tall_streamk >> $O/r2p_syrk_sweep.jsonl 2>> $O/r2p_syrk_sweep.err
done
BA_SYRK_NO_TALL=1 timeout 300 python tools/syrk_sweep.py --cams 200 --points 100000 --tag no_tall >> $O/r2p_syrk_sweep.jsonl 2>> $O/r2p_syrk_sweep.err
BA_SYRK_NO_STREAMK=1 timeout 300 python tools/syrk_sweep.py --cams 50 --points 10000 --tag no_streamk >> $O/r2p_syrk_sweep.jsonl 2>> $O/r2p_syrk_sweep.err
timeout 300 python tools/syrk_sweep.py --cams 200 --points 12500 --tag tall_n8shard >> $O/r2p_syrk_sweep.jsonl 2>> $O/r2p_syrk_sweep.err
timeout 300 python tools/syrk_sweep.py --cams 30 --points 20000 --tag streamk_30x20k >> $O/r2p_syrk_sweep.jsonl 2>> $O/r2p_syrk_sweep.err
timeout 600 python bench.py --steps 20 --warmup 5 > $O/r2p_bench_c3.json 2> $O/r2p_bench_c3.err
echo "bench rc=$?" >> $O/r2p_bench_c3.err
