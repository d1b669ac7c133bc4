#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out
: > $O/r2k_syrk_lpt3.jsonl
for f in 0.6 0.7 0.85; do
  BA_SYRK_RUN_FLOOR=$f timeout 300 python tools/syrk_sweep.py --cams 50 --points 10000 --tag run_floor$f >> $O/r2k_syrk_lpt3.jsonl 2>> $O/r2k_syrk_lpt3.err
  BA_SYRK_RUN_FLOOR=$f timeout 300 python tools/syrk_sweep.py --cams 30 --points 20000 --tag run_floor$f >> $O/r2k_syrk_lpt3.jsonl 2>> $O/r2k_syrk_lpt3.err
  BA_SYRK_RUN_FLOOR=$f timeout 300 python tools/syrk_sweep.py --cams 100 --points 5000 --tag run_floor$f >> $O/r2k_syrk_lpt3.jsonl 2>> $O/r2k_syrk_lpt3.err
done
