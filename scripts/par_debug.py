"""debug driver: R ranks as threads on ONE GPU (host-staged exchange), sizes/flags from the command line"""
import faulthandler
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import ngsamg_b200 as ng
from ngsamg_b200 import parallel as par, synthetic as S

faulthandler.dump_traceback_later(int(sys.argv[4]) if len(sys.argv) > 4 else 120, exit=True)
n, R = int(sys.argv[1]), int(sys.argv[2])
flags = {}
for kv in (sys.argv[3] if len(sys.argv) > 3 else "").split(","):
    if "=" in kv:
        k, v = kv.split("=")
        flags["ngs_amg_" + k] = v
t0 = time.time()
parts = [S.slab_poisson3d(n, n, n, R, r) for r in range(R)]
print("gen %.1fs" % (time.time() - t0), flush=True)


def build(r, comm):
    p = parts[r]
    A = ng.SparseMatrix(p["n"], p["n"], 1, 1, p["rowptr"], p["col"], p["val"])
    t = time.time()
    pc = par.h1_scal_par(A, par.Halo(p["peers"], p["ex"]), comm, p["free"], ngs_amg_log_level="basic", **flags)
    print("rank %d setup %.1fs npar %d levels %s" % (r, time.time() - t, pc.GetNParallelLevels(),
                                                    [(pc.level_info(l).n, pc.level_info(l).gs_depth) for l in range(pc.GetNLevels())]), flush=True)
    return pc


pcs = par.run_ranks(R, build, timeout=600)


def solve(r, comm):
    p = parts[r]
    x = np.zeros(p["n"])
    t = time.time()
    pcs[r].Mult(p["rhs"], x)
    print("rank %d apply %.3fs" % (r, time.time() - t), flush=True)
    t = time.time()
    it, errs = pcs[r]._pcg(p["rhs"], x, 1e-8, 100)
    print("rank %d pcg %.3fs its %d" % (r, time.time() - t, it), flush=True)
    return it


print(par.run_ranks(R, solve, timeout=600))
