#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out
BA_TIMING=1 timeout 600 python tools/time_e2e_fine.py c3 20 > $O/r2h_e2e_fine_c3.txt 2>&1
BA_TIMING=1 timeout 600 python tools/time_e2e_fine.py c2 20 > $O/r2h_e2e_fine_c2.txt 2>&1
