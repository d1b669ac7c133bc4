#!/bin/bash
# GPU call 1 of round 2: tests, first c3-default bench lines, the reference arm, compute-sanitizer on
# the small golden cases, the unmodified script, SYRK planner experiments.  Everything lands in gpurun_out/.
set -u
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $O/r2_smi.txt 2>&1
nproc > $O/r2_host.txt; free -g >> $O/r2_host.txt
( time timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 ) > $O/r2_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> $O/r2_pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 5 > $O/r2_bench_c3.json 2> $O/r2_bench_c3.err
echo "bench rc=$?" >> $O/r2_bench_c3.err
timeout 600 python bench.py --impl reference --steps 4 --warmup 1 > $O/r2_bench_ref_c3.json 2> $O/r2_bench_ref_c3.err
timeout 300 python tools/run_reference_script.py oracle/_ref/euclidiean_reconstruction.py > $O/r2_script_run.txt 2>&1
echo "script rc=$?" >> $O/r2_script_run.txt
# SYRK: how much do the ragged edge tiles of n_pad = 1808 cost?  (M = 199 -> n_pad = 1792 = 14 x 128)
for m in 199 200; do timeout 300 python tools/syrk_sweep.py --cams $m --points 100000 --tag base >> $O/r2_syrk_sweep.jsonl 2>> $O/r2_syrk_sweep.err; done
for fl in 0.5 0.25; do BA_SYRK_FLOOR=$fl timeout 300 python tools/syrk_sweep.py --cams 200 --points 100000 --tag floor$fl >> $O/r2_syrk_sweep.jsonl 2>> $O/r2_syrk_sweep.err; done
# compute-sanitizer (ONE tool per gpurun call, B200_PROFILING.md): memcheck on the small golden cases
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -q -x -k "golden_cases or minimal_problem or camera_without" > $O/r2_sanitizer_memcheck.log 2>&1
echo "memcheck rc=$?" >> $O/r2_sanitizer_memcheck.log
ls -la $O | tail -20
