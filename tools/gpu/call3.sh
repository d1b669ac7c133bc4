#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out/r2c_exit_probe.txt
: > $O
echo "LD_LIBRARY_PATH=$LD_LIBRARY_PATH" >> $O
for m in ref_only project_only ba_small ba_small_debug ba_small_torch_first ba_small_no_torch; do
  timeout 300 python -X faulthandler tools/gpu/exit_probe.py $m >> $O 2>&1
  echo "== $m rc=$?" >> $O
done
O2=gpurun_out
# TMA-fed SYRK: parity first, then timing against the cp.async kernel and planner floors
( time timeout 1200 python -m pytest tests -m gpu -q -x --timeout 600 ) > $O2/r2c_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> $O2/r2c_pytest_gpu.log
: > $O2/r2c_syrk_sweep.jsonl
for m in 199 200 50; do
  pts=100000; if [ $m = 50 ]; then pts=10000; fi
  timeout 300 python tools/syrk_sweep.py --cams $m --points $pts --tag tma >> $O2/r2c_syrk_sweep.jsonl 2>> $O2/r2c_syrk_sweep.err
  BA_SYRK_NO_TMA=1 timeout 300 python tools/syrk_sweep.py --cams $m --points $pts --tag cpasync >> $O2/r2c_syrk_sweep.jsonl 2>> $O2/r2c_syrk_sweep.err
done
for fl in 0.75 0.5 0.15; do BA_SYRK_FLOOR=$fl timeout 300 python tools/syrk_sweep.py --cams 200 --points 100000 --tag tma_floor$fl >> $O2/r2c_syrk_sweep.jsonl 2>> $O2/r2c_syrk_sweep.err; done
for fl in 1.0 0.5; do BA_SYRK_FLOOR=$fl timeout 300 python tools/syrk_sweep.py --cams 50 --points 10000 --tag tma_floor$fl >> $O2/r2c_syrk_sweep.jsonl 2>> $O2/r2c_syrk_sweep.err; done
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O2/r2c_bench_c3.json 2> $O2/r2c_bench_c3.err
echo "bench rc=$?" >> $O2/r2c_bench_c3.err
