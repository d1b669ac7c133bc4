#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out
: > $O/r2o_pairs.txt
for v in "BA_PAIRS_REG=1" "X=1"; do
  echo "== $v" >> $O/r2o_pairs.txt
  env $v timeout 300 python tools/time_phases.py --cams 1000 --points 200000 --vis 0.1 --iters 3 >> $O/r2o_pairs.txt 2>&1
done
( time BA_PAIRS_REG=1 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 600 -k "kernels_match or thousand or outlier or graph_loop or single_view" ) > $O/r2o_pytest_pairs.log 2>&1
echo "pytest rc=$?" >> $O/r2o_pytest_pairs.log
