#!/usr/bin/env bash
set -u
OUT=gpurun_out/r02_s2c28
mkdir -p "$OUT"
timeout 300 python -m pytest tests/test_gpu_parity.py -q -k "sweep or wide or elasticity or vcycle or smoother" 2>&1 | tail -3
cat > /tmp/tune_min.py <<'PY'
import os, sys
sys.path.insert(0, os.getcwd())
import ngsamg_b200 as ng
from ngsamg_b200 import synthetic as S
p = S.poisson3d_kuhn(151)
A = ng.SparseMatrix(p["n"], p["n"], 1, 1, p["rowptr"], p["col"], p["val"])
pc = ng.h1_scal(A, p["free"])
NL = pc.GetNLevels()
for st in ({}, {"tri_rm": 0}, {"tri_rm": 0, "tri_sleep_ns": 0}):
    for k, v in {"tri_rm": 1, "tri_sleep_ns": 100}.items():
        pc.SetTunable(k, st.get(k, v))
    row = []
    for l in range(1, NL - 1):
        f, _ = pc.ProfileKernel("gs_tri_fwd", level=l, reps=10); b, _ = pc.ProfileKernel("gs_tri_bwd", level=l, reps=10)
        row.append("%d:%.1f/%.1f" % (l, f * 1e3, b * 1e3))
    print(st, "  ".join(row), flush=True)
PY
timeout 300 python /tmp/tune_min.py 2>&1 | tail -4
