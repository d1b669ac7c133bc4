#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out
: > $O/r2k_syrk_lpt2.jsonl
for f in 0.5 0.75; do
  BA_SYRK_FLOOR=$f timeout 300 python tools/syrk_sweep.py --cams 50 --points 10000 --tag lpt_floor$f >> $O/r2k_syrk_lpt2.jsonl 2>> $O/r2k_syrk_lpt2.err
  BA_SYRK_FLOOR=$f timeout 300 python tools/syrk_sweep.py --cams 30 --points 20000 --tag lpt_floor$f >> $O/r2k_syrk_lpt2.jsonl 2>> $O/r2k_syrk_lpt2.err
  BA_SYRK_FLOOR=$f timeout 300 python tools/syrk_sweep.py --cams 100 --points 5000 --tag lpt_floor$f >> $O/r2k_syrk_lpt2.jsonl 2>> $O/r2k_syrk_lpt2.err
done
