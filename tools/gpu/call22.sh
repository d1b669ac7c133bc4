#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out
: > $O/r2v_pairs_ab.txt
for v in "" "BA_PAIRS_OCC12=1" "BA_PAIRS_REG=1"; do
  echo "== $v" >> $O/r2v_pairs_ab.txt
  env $v timeout 300 python tools/time_phases.py --cams 1000 --points 200000 --vis 0.1 --iters 3 >> $O/r2v_pairs_ab.txt 2>&1
done
