#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out
: > $O/r2ag_backsolve_ab.txt
for v in "X=1" "BA_CHOL_BACKSOLVE_1CTA=1"; do
  for c in "50 10000" "200 20000" "120 2000"; do
    set -- $c
    echo "== $v $c" >> $O/r2ag_backsolve_ab.txt
    env $v timeout 300 python tools/chol_ab.py --cams $1 --points $2 >> $O/r2ag_backsolve_ab.txt 2>&1
  done
done
