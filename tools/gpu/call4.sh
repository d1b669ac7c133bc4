#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out/r2d_exit_probe.txt
: > $O
timeout 300 python -X faulthandler tools/gpu/exit_probe.py script_trace >> $O 2>&1
echo "== script_trace rc=$?" >> $O
OPENBLAS_NUM_THREADS=1 OMP_NUM_THREADS=1 timeout 300 python -X faulthandler tools/run_reference_script.py oracle/_ref/euclidiean_reconstruction.py > gpurun_out/r2d_script_1thread.txt 2>&1
echo "== one BLAS thread rc=$?" >> $O
CUDA_LAUNCH_BLOCKING=1 timeout 300 python -X faulthandler tools/run_reference_script.py oracle/_ref/euclidiean_reconstruction.py > gpurun_out/r2d_script_blocking.txt 2>&1
echo "== launch blocking rc=$?" >> $O
