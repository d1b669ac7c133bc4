#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/gpu/dual_probe.py > gpurun_out/r2h_dual_probe.txt 2>&1
echo "rc=$?" >> gpurun_out/r2h_dual_probe.txt
