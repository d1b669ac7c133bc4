#!/bin/bash
# final N=1 evidence (second pass): smoke, tests, default bench, reference arm, c4 bench
set -u
mkdir -p gpurun_out
O=gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > $O/r2f_smoke.txt 2>&1
echo "smoke rc=$?" >> $O/r2f_smoke.txt
( time timeout 1500 python -m pytest tests -m gpu -q --timeout 900 ) > $O/r2f_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> $O/r2f_pytest_gpu.log
timeout 900 python bench.py --steps 20 --warmup 5 > $O/r2f_bench_c3.json 2> $O/r2f_bench_c3.err
echo "bench rc=$?" >> $O/r2f_bench_c3.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/r2f_bench_ref.json 2> $O/r2f_bench_ref.err
timeout 900 python bench.py --workload c4 --extras none --steps 10 --warmup 3 > $O/r2f_bench_c4.json 2> $O/r2f_bench_c4.err
echo "bench rc=$?" >> $O/r2f_bench_c4.err
