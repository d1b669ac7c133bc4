#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out
: > $O/r2aa_chol_ab.txt
for v in "BA_CHOL_NO_PERSIST=1" "X=1"; do
  for c in "50 10000" "200 20000" "24 500"; do
    set -- $c
    echo "== $v $c" >> $O/r2aa_chol_ab.txt
    env $v timeout 300 python tools/chol_ab.py --cams $1 --points $2 >> $O/r2aa_chol_ab.txt 2>&1
  done
done
( time timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 ) > $O/r2aa_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> $O/r2aa_pytest_gpu.log
