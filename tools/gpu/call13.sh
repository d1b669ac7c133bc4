#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out
( time timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 900 ) > $O/r2m_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> $O/r2m_pytest_gpu.log
: > $O/r2m_syrk_sweep.jsonl
for m in 200 50; do
  pts=100000; if [ $m = 50 ]; then pts=10000; fi
  timeout 300 python tools/syrk_sweep.py --cams $m --points $pts --tag thin >> $O/r2m_syrk_sweep.jsonl 2>> $O/r2m_syrk_sweep.err
done
for fl in 0.25 0.15; do BA_SYRK_FLOOR=$fl timeout 300 python tools/syrk_sweep.py --cams 200 --points 100000 --tag thin_floor$fl >> $O/r2m_syrk_sweep.jsonl 2>> $O/r2m_syrk_sweep.err; done
for fl in 0.75 0.6; do BA_SYRK_FLOOR=$fl timeout 300 python tools/syrk_sweep.py --cams 50 --points 10000 --tag thin_floor$fl >> $O/r2m_syrk_sweep.jsonl 2>> $O/r2m_syrk_sweep.err; done
timeout 300 python tools/syrk_sweep.py --cams 200 --points 12500 --tag thin_n8shard >> $O/r2m_syrk_sweep.jsonl 2>> $O/r2m_syrk_sweep.err
