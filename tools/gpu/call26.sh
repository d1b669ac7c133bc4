#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out
python - > $O/r2z_fp64_occ.txt 2>&1 <<'PY'
import ba_b200
e = ba_b200.submodule("engine")
for w in (4, 8, 12, 16, 32):
    print("dfma, warps per SM", w, round(e.fp64_peak(0, 100 + w), 2), "TF/s")
PY
( time BA_PAIRS_REG=1 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 600 -k "kernels_match or thousand or outlier or graph_loop or single_view" ) > $O/r2z_pytest_pairs.log 2>&1
echo "pytest rc=$?" >> $O/r2z_pytest_pairs.log
: > $O/r2z_pairs_ab.txt
for v in "BA_PAIRS_REG=1"; do
  echo "== $v" >> $O/r2z_pairs_ab.txt
  env $v timeout 300 python tools/time_phases.py --cams 1000 --points 200000 --vis 0.1 --iters 3 >> $O/r2z_pairs_ab.txt 2>&1
done
