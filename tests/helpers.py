"""shared test helpers: problem set-up, product <-> oracle conversions, comparisons."""
import numpy as np

import ngsamg_b200 as ng
from ngsamg_b200 import synthetic as S
from oracle import oracle as O


def to_oracle(M):
    return O.Bsr(M.nrows, M.ncols, M.bh, M.bw, M.rowptr, M.col, M.val)


def to_product(M):
    return ng.SparseMatrix(M.nrows, M.ncols, M.bh, M.bw, M.rowptr, M.col, M.val)


def poisson(n, **kw):
    p = S.poisson3d_kuhn(n, **kw)
    A = ng.SparseMatrix(p["n"], p["n"], 1, 1, p["rowptr"], p["col"], p["val"])
    return p, A


def elasticity(nx, ny, nz, **kw):
    p = S.elasticity3d_kuhn(nx, ny, nz, **kw)
    A = ng.SparseMatrix(p["n"], p["n"], 3, 3, p["rowptr"], p["col"], p["val"])
    return p, A


def rel(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    d = np.linalg.norm(b)
    return np.linalg.norm(a - b) / (d if d > 0 else 1.0)


def rand(seed, n):
    return S.splitmix64(20260101 + seed, n)


def host_hierarchy(A, free, xyz=None, elast=False, max_levels=10, max_coarse=50, **copt):
    """the product's host-side coarsening driven from python with the ORACLE doing the Galerkin products
    (CPU-only stand-in for finalize(), used by the not-gpu tests)."""
    prols, cur, fm, cx = [], A, free, xyz
    while cur.nrows > max_coarse and len(prols) + 1 < max_levels:
        bc = cur.bh
        if elast and len(prols) == 0 and cur.bh == 3:
            bc = 6
        P, vmap, cxyz = ng.coarsen(cur, fm, cx, bcoarse=bc, **copt)
        if P.ncols == 0 or P.ncols > 0.8 * cur.nrows:
            break
        prols.append(P)
        Po = to_oracle(P)
        Ac = O.restrict_matrix(O.transpose(Po), to_oracle(cur), Po)
        cur, fm, cx = to_product(Ac), None, cxyz
    return prols


def assert_same_pattern(M1, M2):
    assert M1.nrows == M2.nrows and M1.ncols == M2.ncols and M1.bh == M2.bh and M1.bw == M2.bw
    assert np.array_equal(M1.rowptr, M2.rowptr), "row pointers differ"
    assert np.array_equal(M1.col, M2.col), "column indices differ"
