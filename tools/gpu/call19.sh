#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 600 -k "kernels_match or thousand or outlier or graph_loop or single_view" ) > $O/r2s_pytest_pairs.log 2>&1
echo "pytest rc=$?" >> $O/r2s_pytest_pairs.log
timeout 600 python bench.py --workload c4 --extras none --steps 5 --warmup 3 --no-cpu-baseline > $O/r2s_bench_c4_reg.json 2> $O/r2s_bench_c4_reg.err
echo "rc=$?" >> $O/r2s_bench_c4_reg.err
BA_PAIRS_DMMA=1 timeout 600 python bench.py --workload c4 --extras none --steps 5 --warmup 3 --no-cpu-baseline > $O/r2s_bench_c4_dmma.json 2> $O/r2s_bench_c4_dmma.err
echo "rc=$?" >> $O/r2s_bench_c4_dmma.err
