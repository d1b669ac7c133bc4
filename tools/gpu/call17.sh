#!/bin/bash
# final-kernel ncu captures of the SYRK (C3, C2), mixed DMMA+DFMA peak probe
set -u
mkdir -p gpurun_out
O=gpurun_out
NCU="ncu --clock-control none"
python - > $O/r2q_fp64_mixed.txt 2>&1 <<'PY'
import ba_b200
e = ba_b200.submodule("engine")
for mode, name in ((1, "dmma"), (0, "dfma"), (2, "mixed")):
    print(name, e.fp64_peak(0, mode if mode != 1 else True) if mode else e.fp64_peak(0, False))
PY
python tools/prof_case.py --cams 200 --points 100000 --solves 2 > $O/r2q_prof_c3_plain.log 2>&1 &&
$NCU --set full --import-source on -k regex:syrk_tma -s 1 -c 1 -f -o $O/r2b_syrk_tma_c3 python tools/prof_case.py --cams 200 --points 100000 --solves 2 > $O/r2q_prof_c3_ncu.log 2>&1
echo "c3 ncu rc=$?" >> $O/r2q_prof_c3_ncu.log
python tools/prof_case.py --cams 50 --points 10000 --solves 2 > $O/r2q_prof_c2_plain.log 2>&1 &&
$NCU --set full --import-source on -k "regex:syrk_tma|chol_step|syrk_reduce" -s 9 -c 10 -f -o $O/r2b_kernels_c2 python tools/prof_case.py --cams 50 --points 10000 --solves 2 > $O/r2q_prof_c2_ncu.log 2>&1
echo "c2 ncu rc=$?" >> $O/r2q_prof_c2_ncu.log
ls -la $O | grep "r2q\|r2b"
