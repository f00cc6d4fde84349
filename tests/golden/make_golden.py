"""Generates tests/golden/*.npz: frozen outputs of the (scipy / pure-python cross-checked) oracle on small seeded problems.
The reference itself cannot run in this image (no NGSolve) and its tests hold no golden vectors, so these fixtures pin the
ORACLE against regressions; tests/test_oracle.py checks the oracle against scipy/dense/pure-python independently.
Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from helpers import host_hierarchy, poisson, elasticity, rand, to_oracle  # noqa: E402
from oracle import oracle as O  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def make_poisson():
    p, A = poisson(7)
    prols = host_hierarchy(A, p["free"], max_coarse=30)[:1]
    P = prols[0]
    amg = O.OracleAMG(to_oracle(A), p["free"], [to_oracle(P)])
    Ac = amg.level_matrix(1)
    b = rand(1234, p["n"])
    x = amg.apply(b)
    u, it, errs = amg.pcg(b, tol=1e-8, maxsteps=50)
    np.savez_compressed(os.path.join(HERE, "poisson_n7.npz"), n=p["n"], rowptr=A.rowptr, col=A.col, val=A.val, free=p["free"],
                        nc0=P.ncols, p0_rowptr=P.rowptr, p0_col=P.col, p0_val=P.val, ac_rowptr=Ac.rowptr, ac_col=Ac.col,
                        ac_val=Ac.val, b=b, vcycle_x=x, pcg_iters=it, pcg_errors=errs, pcg_u=u)
    print("poisson_n7: n=%d nc=%d nnz(Ac)=%d pcg iters=%d" % (p["n"], P.ncols, Ac.nnz, it))


def make_elasticity():
    p, A = elasticity(5, 3, 3)
    prols = host_hierarchy(A, p["free"], xyz=p["xyz"], elast=True, max_coarse=4, max_per_row=4)[:1]
    P = prols[0]
    amg = O.OracleAMG(to_oracle(A), p["free"], [to_oracle(P)], pinv=True)
    Ac = amg.level_matrix(1)
    b = rand(4321, p["n"] * 3)
    x = amg.apply(b)
    np.savez_compressed(os.path.join(HERE, "elast_5x3x3.npz"), n=p["n"], rowptr=A.rowptr, col=A.col, val=A.val, free=p["free"],
                        xyz=p["xyz"], nc0=P.ncols, p0_rowptr=P.rowptr, p0_col=P.col, p0_val=P.val, ac_rowptr=Ac.rowptr,
                        ac_col=Ac.col, ac_val=Ac.val, b=b, vcycle_x=x)
    print("elast_5x3x3: n=%d nc=%d nnz(Ac)=%d" % (p["n"], P.ncols, Ac.nnz))


def make_poisson_multirank():
    """2 ranks, 9 x 8 x 11 Poisson cut in z: hierarchy from the product's host-side class-respecting coarsening, everything else from the
    multi-rank oracle (hybrid split, modified diagonal, stage sweeps, DCC exchange, contraction, all-reduced CG)"""
    from ngsamg_b200 import synthetic as S
    from oracle import cpu_pipeline as CP
    parts = S.partition_poisson3d(9, 8, 11, grid=(1, 1, 2))
    amg, info = CP.build(parts, ctr_nv=100, max_coarse=15)
    b = [rand(777 + r, p["n"]) * p["free"] for r, p in enumerate(parts)]
    x = amg.apply(b)
    rhs = [p["rhs"] * p["free"] for p in parts]
    u, it, errs = amg.pcg(rhs, tol=1e-8, maxsteps=50)
    out = dict(pcg_iters=it, pcg_errors=errs, distributed_levels=info["distributed_levels"])
    for r in range(2):
        L0 = amg.levels[0]
        M, G = L0.M[r].tocsr(), L0.G[r].tocsr()
        M.sort_indices(); G.sort_indices()
        out.update({"b%d" % r: b[r], "vcycle_x%d" % r: x[r], "pcg_u%d" % r: u[r], "mod_diag%d" % r: L0.md[r].ravel(),
                    "m_indptr%d" % r: M.indptr, "m_indices%d" % r: M.indices, "m_data%d" % r: M.data,
                    "g_indptr%d" % r: G.indptr, "g_indices%d" % r: G.indices, "g_data%d" % r: G.data,
                    "master%d" % r: L0.master[r].astype(np.uint8)})
    np.savez_compressed(os.path.join(HERE, "poisson_2ranks_9x8x11.npz"), **out)
    print("poisson_2ranks_9x8x11: distributed levels=%d pcg iters=%d" % (info["distributed_levels"], it))


if __name__ == "__main__":
    make_poisson()
    make_elasticity()
    make_poisson_multirank()
