#!/bin/bash
set -u
mkdir -p gpurun_out
LD_PRELOAD=$PWD/tools/gpu/libsegv_trace.so timeout 300 python tools/run_reference_script.py oracle/_ref/euclidiean_reconstruction.py > gpurun_out/r2e_script_bt.txt 2>&1
echo "rc=$?" >> gpurun_out/r2e_script_bt.txt
