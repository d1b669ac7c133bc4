#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out
( time timeout 600 python -m pytest tests/test_affine_calibration.py -m gpu -q -x --timeout 300 ) > $O/r2r_pytest_affine.log 2>&1
echo "pytest rc=$?" >> $O/r2r_pytest_affine.log
BA_SCRIPT_REPORT=$O/r2r_script_affine_report.json timeout 300 python tools/run_reference_script.py oracle/_ref/affine_reconstruction.py > $O/r2r_script_affine_run.txt 2>&1
echo "script rc=$?" >> $O/r2r_script_affine_run.txt
