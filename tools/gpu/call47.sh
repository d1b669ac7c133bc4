#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 ) > $O/r2l_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> $O/r2l_pytest_gpu.log
timeout 300 python tools/chol_ab.py --cams 200 --points 100000 > $O/r2l_c3_phases.txt 2>&1
