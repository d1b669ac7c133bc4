#!/bin/bash
# final N=1 evidence: tests, default bench, launch list of the same command, c4 bench, ncu of the matrix-free kernels
set -u
mkdir -p gpurun_out
O=gpurun_out
NCU="ncu --clock-control none"
( time timeout 1500 python -m pytest tests -m gpu -q --timeout 900 ) > $O/r2f_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> $O/r2f_pytest_gpu.log
timeout 900 python bench.py --steps 20 --warmup 5 > $O/r2f_bench_c3.json 2> $O/r2f_bench_c3.err
echo "bench rc=$?" >> $O/r2f_bench_c3.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/r2f_bench_ref.json 2> $O/r2f_bench_ref.err
timeout 900 python bench.py --workload c4 --extras none --steps 10 --warmup 3 > $O/r2f_bench_c4.json 2> $O/r2f_bench_c4.err
echo "bench rc=$?" >> $O/r2f_bench_c4.err
python bench.py --steps 2 --warmup 1 --no-cpu-baseline --extras none > $O/r2f_bench_short_plain.json 2> $O/r2f_bench_short_plain.err &&
$NCU --metrics gpu__time_duration.sum -c 600 --csv --log-file $O/r2f_launches_c3_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --extras none > $O/r2f_bench_short_ncu.json 2> $O/r2f_bench_short_ncu.err
python tools/prof_case.py --cams 200 --points 100000 --solves 1 > $O/r2f_prof_plain.log 2>&1 &&
$NCU --set full --import-source on -k "regex:k2a_point|camera_blocks_kernel|k2b_point|point_update" -c 4 -f -o $O/r2f_matrix_free_c3 python tools/prof_case.py --cams 200 --points 100000 --solves 1 > $O/r2f_prof_ncu.log 2>&1
echo "ncu rc=$?" >> $O/r2f_prof_ncu.log
