#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/gpu/dual_probe.py > gpurun_out/r2i_dual_probe.txt 2>&1
echo "rc=$?" >> gpurun_out/r2i_dual_probe.txt
( time timeout 900 python -m pytest tests/test_projective_depth.py -m gpu -q --timeout 600 ) > gpurun_out/r2i_pytest_depth.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2i_pytest_depth.log
