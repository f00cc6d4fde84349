import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import ngsamg_b200 as ng
from helpers import poisson, rand, to_oracle
from oracle import oracle as O
p, A = poisson(12)
pc = ng.h1_scal(A, p["free"], ngs_amg_max_coarse_size=20, ngs_amg_b200_tri_small_rows=0, ngs_amg_b200_tri_level_launch_depth=1000, ngs_amg_b200_cuda_graph=0)
amg = O.OracleAMG(to_oracle(A), p["free"], [to_oracle(P) for P in pc.GetMap()])
b = rand(97, p["n"])
x = pc * b
xo = amg.apply(b)
print("x nan", np.isnan(x).sum(), "of", x.size)
for l in range(pc.GetNLevels()):
    for w in ("rhs", "res", "x"):
        v = pc.GetLevelVector(w, l)
        try:
            o = amg.level_vec(w, l) if not (l == 0 and w in ("x", "rhs")) else (xo if w == "x" else b)
            err = np.linalg.norm(v - o) / max(np.linalg.norm(o), 1e-300)
        except Exception as e:
            err = -1
        print("level", l, w, "nan", int(np.isnan(v).sum()), "n", v.size, "depth", pc.level_info(l).gs_depth, "relerr %.2e" % err)
