#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q --timeout 900 ) > $O/r2f_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> $O/r2f_pytest_gpu.log
timeout 300 python tools/run_reference_script.py oracle/_ref/euclidiean_reconstruction.py > $O/r2f_script_euclid.txt 2>&1
echo "script rc=$?" >> $O/r2f_script_euclid.txt
timeout 300 python tools/run_reference_script.py oracle/_ref/affine_reconstruction.py > $O/r2f_script_affine.txt 2>&1
echo "script rc=$?" >> $O/r2f_script_affine.txt
timeout 900 python bench.py --workload c4 --extras none --steps 10 --warmup 3 > $O/r2f_bench_c4.json 2> $O/r2f_bench_c4.err
echo "bench rc=$?" >> $O/r2f_bench_c4.err
