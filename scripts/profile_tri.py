"""small driver for ncu / experiments: builds a Poisson hierarchy and launches the triangular sweep kernels a few times"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ngsamg_b200 as ng
from ngsamg_b200 import synthetic as S
n = int(sys.argv[1]) if len(sys.argv) > 1 else 151
extra = {}
for kv in os.environ.get("NGSAMG_FLAGS", "").split(","):
    if "=" in kv:
        k, v = kv.split("=", 1)
        extra["ngs_amg_" + k.strip()] = v.strip()
import numpy as np
diri = tuple(x for x in os.environ.get("DIRI", "x0,y1").split(",") if x)
p = S.poisson3d_kuhn(n, dirichlet=diri)
val = p["val"]
if not diri:   # regularise the pure Neumann matrix
    rows = np.repeat(np.arange(p["n"]), np.diff(p["rowptr"]))
    val = val + (rows == p["col"]) * 1e-3
A = ng.SparseMatrix(p["n"], p["n"], 1, 1, p["rowptr"], p["col"], val)
print("dirichlet", diri)
pc = ng.h1_scal(A, p["free"], **extra)
for kv in os.environ.get("TUNABLES", "").split(","):
    if "=" in kv:
        k, v = kv.split("=", 1)
        pc.SetTunable(k.strip(), float(v))
for lvl in [int(x) for x in os.environ.get('LEVELS', '0').split(',')]:
    for name in os.environ.get("KERNELS", "gs_tri_fwd,gs_tri_bwd").split(","):
        ms, by = pc.ProfileKernel(name, level=lvl, reps=5)
        print(extra, lvl, name, "%.3f ms" % ms, "%.0f GB/s" % (by / ms / 1e6), "depth", pc.level_info(lvl).gs_depth)
