#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 600 -k "kernels_match or full_run or matrix_free" ) > $O/r2m_pytest.log 2>&1
echo "pytest rc=$?" >> $O/r2m_pytest.log
timeout 300 python tools/chol_ab.py --cams 50 --points 10000 > $O/r2m_c2_phases.txt 2>&1
