#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out
: > $O/r2i_pairs.txt
timeout 300 python tools/time_phases.py --cams 1000 --points 200000 --vis 0.1 --iters 3 >> $O/r2i_pairs.txt 2>&1
( time timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 600 -k "kernels_match or thousand or outlier or graph_loop or single_view" ) > $O/r2i_pytest_pairs.log 2>&1
echo "pytest rc=$?" >> $O/r2i_pytest_pairs.log
