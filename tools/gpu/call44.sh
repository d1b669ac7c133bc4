#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out
: > $O/r2k_syrk_lpt.jsonl
for l in 1 0; do
  BA_SYRK_LPT=$l timeout 300 python tools/syrk_sweep.py --cams 50 --points 10000 --tag lpt$l >> $O/r2k_syrk_lpt.jsonl 2>> $O/r2k_syrk_lpt.err
  BA_SYRK_LPT=$l timeout 300 python tools/syrk_sweep.py --cams 30 --points 20000 --tag lpt$l >> $O/r2k_syrk_lpt.jsonl 2>> $O/r2k_syrk_lpt.err
  BA_SYRK_LPT=$l timeout 300 python tools/syrk_sweep.py --cams 100 --points 5000 --tag lpt$l >> $O/r2k_syrk_lpt.jsonl 2>> $O/r2k_syrk_lpt.err
done
( time timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 600 -k "kernels_match or full_run or c2_properties" ) > $O/r2k_pytest.log 2>&1
echo "pytest rc=$?" >> $O/r2k_pytest.log
