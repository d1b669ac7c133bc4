#!/bin/bash
mkdir -p gpurun_out
timeout 60 ./tools/gpu/jac_test > gpurun_out/r2k_jac_test.txt 2>&1
echo "rc=$?" >> gpurun_out/r2k_jac_test.txt
