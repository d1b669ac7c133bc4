#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 600 -k "kernels_match or thousand or outlier or graph_loop or single_view or full_run" ) > $O/r2n_pytest_pairs.log 2>&1
echo "pytest rc=$?" >> $O/r2n_pytest_pairs.log
: > $O/r2n_pairs.txt
for v in "X=1" "BA_PAIRS_NO_LIST=1"; do
  echo "== $v" >> $O/r2n_pairs.txt
  env $v BA_TIMING=1 timeout 300 python tools/time_phases.py --cams 1000 --points 200000 --vis 0.1 --iters 3 >> $O/r2n_pairs.txt 2>&1
done
timeout 900 python bench.py --workload c4 --extras none --steps 10 --warmup 3 --no-cpu-baseline > $O/r2n_bench_c4.json 2> $O/r2n_bench_c4.err
echo "bench rc=$?" >> $O/r2n_bench_c4.err
