"""Pins the oracle (oracle/ngsamg_oracle.c) against the REFERENCE'S OWN CODE for the hot path (`-m "not gpu"`).

oracle/_ref/libngsamg_ref.so holds the bodies of TransposeSPMImpl, MatMultABImpl, RestrictMatrix, GSS3::*, BaseSmoother /
ProxySmoother, ProlMap transfers and AMGMatrix::SmoothV/W/BS cut out of /root/reference at build time and compiled verbatim
against a stand-in for the NGSolve containers they call (oracle/ref_pin/README.md says exactly what that covers).
 * live tests: oracle vs that library on seeded inputs -- bit for bit for the integer patterns and for every floating-point
   result that does not pass through the exact coarse solve (the reference uses NGSolve's sparse Cholesky there, the harness a
   dense inverse: <= 1e-13 after it).  Skipped where the library is neither present nor buildable.
 * fixture test: the same comparisons against tests/golden/refpin_*.npz, which tests/golden/make_ref_golden.py wrote from
   that library -- runs on any machine.
"""
import os

import numpy as np
import pytest
import scipy.sparse as sp

from helpers import poisson, elasticity, rand, rel, to_oracle, host_hierarchy
from oracle import oracle as O
from oracle.ref_pin import ref as R

GOLD = os.path.join(os.path.dirname(__file__), "golden")
needs_ref = pytest.mark.skipif(not R.available(), reason="oracle/_ref/libngsamg_ref.so not built and /root/reference not present")
FLAGS = [(ru, ur, xz, bw) for ru in (0, 1) for ur in (0, 1) for xz in (0, 1) for bw in (0, 1)]


def random_bsr(seed, n, m, bh, bw, density=0.2):
    rng = np.random.default_rng(seed)
    pat = sp.random(n, m, density=density, random_state=rng, format="csr")
    pat.sort_indices()
    return O.Bsr(n, m, bh, bw, pat.indptr, pat.indices, rng.standard_normal((pat.nnz, bh, bw)))


def same(M1, M2):
    assert (M1.nrows, M1.ncols, M1.bh, M1.bw) == (M2.nrows, M2.ncols, M2.bh, M2.bw)
    assert np.array_equal(M1.rowptr, M2.rowptr) and np.array_equal(M1.col, M2.col), "patterns differ"
    assert np.array_equal(M1.val, M2.val), "values differ (max %.3e)" % np.abs(M1.val - M2.val).max()


@needs_ref
def test_library_is_built_from_the_reference_sources():
    idx = R.fragment_index()
    for name, where in [("transpose", "utils_sparseMM.cpp"), ("matmult", "utils_sparseMM.cpp"), ("restrict", "utils_sparseMM.hpp"),
                        ("gss3_rhs", "gssmoother.cpp"), ("gss3_res", "gssmoother.cpp"), ("gss3_calcdiags", "gssmoother.cpp"),
                        ("proxy_smooth", "base_smoother.hpp"), ("prol_f2c", "dof_map.cpp"), ("prol_addc2f", "dof_map.cpp"),
                        ("amg_smoothv", "amg_matrix.cpp"), ("amg_smoothw", "amg_matrix.cpp"), ("amg_smoothbs", "amg_matrix.cpp")]:
        line = [ln for ln in idx.splitlines() if ln.split()[0] == name]
        assert line and where in line[0], (name, idx)
    # the cut-out text never enters the repository
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    assert "oracle/_ref/" in open(os.path.join(root, ".gitignore")).read().split()


@needs_ref
@pytest.mark.parametrize("bh,bw", [(1, 1), (2, 3), (3, 2), (3, 3), (3, 6), (6, 3), (6, 6)])
def test_transpose_bit_exact(bh, bw):
    for seed, n, m, dens in [(1, 17, 11, 0.2), (2, 40, 3, 0.5), (3, 5, 60, 0.05)]:   # incl. empty rows and columns
        A = random_bsr(seed, n, m, bh, bw, dens)
        same(O.transpose(A), R.transpose(A))


@needs_ref
@pytest.mark.parametrize("a,b,c", [(1, 1, 1), (2, 2, 3), (3, 2, 3), (3, 3, 3), (3, 3, 6), (6, 3, 3), (6, 3, 6), (3, 6, 6), (6, 6, 3), (6, 6, 6)])
def test_matmult_bit_exact(a, b, c):
    A, B = random_bsr(2, 23, 19, a, b), random_bsr(3, 19, 13, b, c)
    same(O.matmul(A, B), R.matmul(A, B))
    A, B = random_bsr(4, 9, 30, a, b, 0.7), random_bsr(5, 30, 21, b, c, 0.02)     # > 16 merged lists, many empty rows of B
    same(O.matmul(A, B), R.matmul(A, B))


@needs_ref
def test_matmult_hash_collisions_take_the_search_branch():
    """product rows that hold columns c and c + 2048 (same slot of the 2048-entry hash, utils_sparseMM.cpp:183-216)"""
    rng = np.random.default_rng(7)
    n, k, m = 12, 40, 9000
    A = random_bsr(8, n, k, 1, 1, 0.6)
    rows, cols = [], []
    for r in range(k):
        base = rng.choice(2048, size=6, replace=False)
        cs = np.unique(np.concatenate([base, base[:4] + 2048, base[:2] + 4096]))
        rows += [r] * len(cs)
        cols += list(cs)
    pat = sp.csr_matrix((np.ones(len(rows)), (rows, cols)), shape=(k, m))
    pat.sort_indices()
    B = O.Bsr(k, m, 1, 1, pat.indptr, pat.indices, rng.standard_normal(pat.nnz))
    Co, Cr = O.matmul(A, B), R.matmul(A, B)
    row_cols = Cr.col[Cr.rowptr[0]:Cr.rowptr[1]]
    assert len(set(row_cols % 2048)) < len(row_cols)          # there ARE collisions in the row
    same(Co, Cr)
    assert np.abs(Cr.to_scipy() - A.to_scipy() @ B.to_scipy()).max() < 1e-12


@needs_ref
def test_matmult_rows_wider_than_the_hash():
    """a product row with more than 1024 entries makes the reference grow its hash (utils_sparseMM.cpp:187-188)"""
    A = random_bsr(9, 6, 50, 1, 1, 0.9)
    B = random_bsr(10, 50, 3000, 1, 1, 0.05)
    Co, Cr = O.matmul(A, B), R.matmul(A, B)
    assert np.diff(Cr.rowptr).max() > 1024
    same(Co, Cr)


@needs_ref
@pytest.mark.parametrize("h,w", [(1, 1), (3, 3), (3, 6), (6, 6), (2, 3)])
def test_restrict_matrix_bit_exact(h, w):
    A = random_bsr(11, 30, 30, h, h, 0.15)
    P = random_bsr(12, 30, 8, h, w, 0.12)
    PT = O.transpose(P)
    same(O.restrict_matrix(PT, A, P), R.restrict_matrix(PT, A, P))


def hierarchy(kind):
    if kind == "poisson":
        p, A = poisson(9)
        prols = host_hierarchy(A, p["free"], max_coarse=30)
    else:
        p, A = elasticity(5, 4, 4)
        prols = host_hierarchy(A, p["free"], p["xyz"], elast=True, max_coarse=4, max_per_row=4)
    return p, to_oracle(A), [to_oracle(P) for P in prols]


def check_hierarchy(oa, get_mat, get_dinv, nlevels):
    for l in range(nlevels):
        same(oa.level_matrix(l), get_mat(l))
        if l + 1 < nlevels:
            assert np.array_equal(oa.level_dinv(l), get_dinv(l)), "dinv differs on level %d" % l


@needs_ref
@pytest.mark.parametrize("kind", ["poisson", "elasticity"])
def test_galerkin_hierarchy_and_diagonal_inverses_bit_exact(kind):
    p, A, prols = hierarchy(kind)
    oa, ra = O.OracleAMG(A, p["free"], prols), R.RefAMG(A, p["free"], prols)
    assert oa.nlevels == ra.nlevels >= 3
    check_hierarchy(oa, ra.level_matrix, ra.level_dinv, oa.nlevels)


def sweep_inputs(A, ru, xz):
    nb = A.nrows * A.bh
    x0, b0 = rand(11, nb), rand(12, nb)
    x = np.zeros(nb) if xz else x0.copy()
    res = (b0.copy() if xz else b0 - A.to_scipy() @ x0) if ru else rand(13, nb)
    return x, b0, res


@needs_ref
@pytest.mark.parametrize("kind", ["poisson", "elasticity"])
@pytest.mark.parametrize("freemask", ["given", "all", "none_set", "null"])
def test_gss3_sweeps_bit_exact_for_all_protocol_flags(kind, freemask):
    """GSS3::Smooth / SmoothBack -> SmoothRESInternal / SmoothRHSInternal (gssmoother.cpp:195-398) incl. first_free / next_free"""
    p, A, prols = hierarchy(kind)
    free = {"given": p["free"], "all": np.ones(A.nrows, np.uint8), "none_set": np.zeros(A.nrows, np.uint8), "null": None}[freemask]
    if freemask == "given":
        free = np.array(free, np.uint8)
        free[:3] = 0            # first_free > 0
        free[-5:] = 0           # next_free < n
    oa, ra = O.OracleAMG(A, free, prols[:1], clev="none"), R.RefAMG(A, free, prols[:1], coarse_inv=False)
    for ru, ur, xz, bw in FLAGS:
        x1, b, r1 = sweep_inputs(A, ru, xz)
        x2, r2 = x1.copy(), r1.copy()
        oa.smooth(0, x1, b, r1, ru, ur, xz, bw)
        ra.smooth(0, x2, b, r2, ru, ur, xz, bw, bare=True)
        assert np.array_equal(x1, x2), (ru, ur, xz, bw, rel(x1, x2))
        if ur:
            assert np.array_equal(r1, r2), (ru, ur, xz, bw, rel(r1, r2))


@needs_ref
@pytest.mark.parametrize("steps,symm", [(1, True), (2, False), (3, True)])
def test_proxy_smoother_bit_exact(steps, symm):
    p, A, prols = hierarchy("poisson")
    oa = O.OracleAMG(A, p["free"], prols[:1], sm_steps=steps, sm_symm=symm, clev="none")
    ra = R.RefAMG(A, p["free"], prols[:1], sm_steps=steps, sm_symm=symm, coarse_inv=False)
    for ru, ur, xz, bw in FLAGS:
        x1, b, r1 = sweep_inputs(A, ru, xz)
        x2, r2 = x1.copy(), r1.copy()
        oa.smooth(0, x1, b, r1, ru, ur, xz, bw)
        ra.smooth(0, x2, b, r2, ru, ur, xz, bw)
        assert np.array_equal(x1, x2), (ru, ur, xz, bw, rel(x1, x2))
        if ur:
            assert np.array_equal(r1, r2), (ru, ur, xz, bw)


def check_cycles(oa, apply_ref, level_vec_ref, b, nlevels):
    x = oa.apply(b, "V")
    # everything before the exact coarse solve is bit-identical ...
    assert np.array_equal(oa.level_vec("res", 0), level_vec_ref("res", 0))
    for l in range(1, nlevels):
        assert np.array_equal(oa.level_vec("rhs", l), level_vec_ref("rhs", l)), "rhs on level %d" % l
        if l + 1 < nlevels:
            assert np.array_equal(oa.level_vec("res", l), level_vec_ref("res", l)), "res on level %d" % l
    # ... after it the two coarse solvers (Cholesky here, dense inverse there) differ in the last bits
    for l in range(1, nlevels):
        assert rel(oa.level_vec("x", l), level_vec_ref("x", l)) < 1e-13
    assert rel(x, apply_ref("V")) < 1e-13
    assert rel(oa.apply(b, "W"), apply_ref("W")) < 1e-13
    assert rel(oa.apply(b, "BS"), apply_ref("BS")) < 1e-13


@needs_ref
@pytest.mark.parametrize("kind,steps,symm", [("poisson", 1, False), ("poisson", 2, True), ("elasticity", 1, False)])
def test_v_w_bs_cycles_against_the_reference_code(kind, steps, symm):
    """AMGMatrix::SmoothV / SmoothW / SmoothBS / SmoothVFromLevel (amg_matrix.cpp:37-374) with ProlMap transfers"""
    p, A, prols = hierarchy(kind)
    oa = O.OracleAMG(A, p["free"], prols, sm_steps=steps, sm_symm=symm)
    ra = R.RefAMG(A, p["free"], prols, sm_steps=steps, sm_symm=symm)
    b = rand(14, A.nrows * A.bh)
    xs = {c: ra.apply(b, c) for c in ("W", "BS", "V")}          # V last: the level vectors read below are the V-cycle's
    check_cycles(oa, lambda c: xs[c], ra.level_vec, b, oa.nlevels)


@needs_ref
def test_cycles_without_coarse_inverse_are_bit_exact():
    """clev = none: no dense solve anywhere, so the whole cycle must agree bit for bit"""
    p, A, prols = hierarchy("poisson")
    oa, ra = O.OracleAMG(A, p["free"], prols, clev="none"), R.RefAMG(A, p["free"], prols, coarse_inv=False)
    b = rand(15, A.nrows)
    for cyc in ("V", "W", "BS"):
        assert np.array_equal(oa.apply(b, cyc), ra.apply(b, cyc)), cyc


# ---- the same pins against the fixtures the reference library wrote (no library needed) --------------------------------
def unpack(g, prefix):
    nr, nc, bh, bw = (int(v) for v in g[prefix + "_shape"])
    return O.Bsr(nr, nc, bh, bw, g[prefix + "_rowptr"], g[prefix + "_col"], g[prefix + "_val"])


@pytest.mark.parametrize("name", ["refpin_poisson_n7", "refpin_poisson_n7_symm2", "refpin_elast_5x3x3"])
def test_oracle_against_reference_made_fixtures(name):
    g = np.load(os.path.join(GOLD, name + ".npz"))
    nlev = int(g["nlevels"])
    assert "amg_matrix.cpp" in str(g["fragments"]) and "gssmoother.cpp" in str(g["fragments"])
    A, free = unpack(g, "A0"), g["free"]
    prols = [unpack(g, "P%d" % l) for l in range(nlev - 1)]
    for l, P in enumerate(prols):
        same(O.transpose(P), unpack(g, "PT%d" % l))
    oa = O.OracleAMG(A, free, prols, sm_steps=int(g["sm_steps"]), sm_symm=bool(g["sm_symm"]))
    check_hierarchy(oa, lambda l: unpack(g, "A%d" % l), lambda l: g["dinv%d" % l], nlev)
    bare = O.OracleAMG(A, free, prols[:1], clev="none")
    for ru, ur, xz, bw in FLAGS:
        x = np.zeros_like(g["sm_x_in"]) if xz else g["sm_x_in"].copy()
        res = (g["sm_b"].copy() if xz else g["sm_res_in_true"].copy()) if ru else g["sm_res_in_junk"].copy()
        bare.smooth(0, x, g["sm_b"], res, ru, ur, xz, bw)
        key = "%d%d%d%d" % (ru, ur, xz, bw)
        assert np.array_equal(x, g["sm_x_" + key]), key
        if ur:
            assert np.array_equal(res, g["sm_res_" + key]), key
    check_cycles(oa, lambda c: g["x_" + c], lambda w, l: g["V_%s%d" % (w, l)], g["b"], nlev)


@needs_ref
def test_pcg_around_the_reference_cycle_matches_the_oracle():
    """same CG restatement on both sides, the reference's SmoothV vs the oracle's as preconditioner: identical iteration counts and
    error histories (this is what bench.py's CPU arm runs, kind "reference")"""
    p, A, _ = hierarchy("poisson")
    ra = R.RefAMG(A, p["free"])                 # level by level, like bench.py builds it
    prols, cur, fm = [], A, p["free"]
    import ngsamg_b200 as ng
    while cur.nrows > 30:
        P, _, _ = ng.coarsen(ng.SparseMatrix(cur.nrows, cur.ncols, 1, 1, cur.rowptr, cur.col, cur.val), fm)
        prols.append(to_oracle(P))
        cur, fm = ra.add_prol(prols[-1]), None
    ra.finalize()
    oa = O.OracleAMG(A, p["free"], prols)
    assert ra.nlevels == oa.nlevels
    u1, it1, e1 = oa.pcg(p["rhs"], tol=1e-8, maxsteps=50)
    u2, it2, e2 = ra.pcg(p["rhs"], tol=1e-8, maxsteps=50)
    assert it1 == it2 and rel(e1, e2) < 1e-12 and rel(u1, u2) < 1e-12


@needs_ref
def test_bench_cpu_arm_runs_the_reference_library():
    import bench
    r = bench.cpu_reference_run(15, 1, 0)
    assert r["kind"] == "reference" and r["iterations"] > 0 and r["ndof"] == 15 ** 3


@needs_ref
@pytest.mark.parametrize("kind", ["poisson", "elasticity"])
@pytest.mark.parametrize("steps,symm", [(1, False), (2, True)])
def test_jacobi_smoother_protocol_vs_reference_code(kind, steps, symm):
    """JacobiSmoother ctor + RichardsonSmoother::Smooth / SmoothBack (base_smoother.cpp:52-114): which vector feeds the update and when
    the residual is recomputed, for all 16 flag combinations.  The update itself is NGSolve's DiagonalMatrix::MultAdd; its
    association (omega * d) * r vs omega * (d * r) is not in the reference tree, so the comparison is to 1e-14, not bit for bit."""
    p, A, prols = hierarchy(kind)
    oa = O.OracleAMG(A, p["free"], prols, sm_type="jacobi", sm_steps=steps, sm_symm=symm)
    ra = R.RefAMG(A, p["free"], prols, sm_steps=steps, sm_symm=symm)
    ra.use_jacobi(0.9, steps, symm)
    for ru, ur, xz, bw in FLAGS:
        x1, b, r1 = sweep_inputs(A, ru, xz)
        x2, r2 = x1.copy(), r1.copy()
        oa.smooth(0, x1, b, r1, ru, ur, xz, bw)
        ra.smooth(0, x2, b, r2, ru, ur, xz, bw)
        assert rel(x1, x2) < 1e-14, (ru, ur, xz, bw)
        if ur:
            assert rel(r1, r2) < 1e-14, (ru, ur, xz, bw)
    b = rand(16, A.nrows * A.bh)
    for cyc in ("V", "W", "BS"):
        assert rel(oa.apply(b, cyc), ra.apply(b, cyc)) < 1e-13


def pinv_cases(n, seed=0):
    rng = np.random.default_rng(seed)
    X = rng.standard_normal((n, n))
    spd = X @ X.T + n * np.eye(n)
    yield "spd", spd, True
    z = spd.copy(); z[1, :] = 0; z[:, 1] = 0
    yield "zero row and column", z, True
    v = rng.standard_normal((n, n - 1))
    yield "rank n-1 (eigenvalue fall-back)", v @ v.T, False
    tiny = spd.copy(); tiny[0, :] *= 1e-8; tiny[:, 0] *= 1e-8
    yield "diagonal entry below RelZeroTol * max", tiny, True
    yield "zero block", np.zeros((n, n)), True
    one = np.zeros((n, n)); one[0, 0] = 2.0
    yield "single entry", one, True
    neg = spd.copy(); neg[0, 0] = -1.0
    yield "negative diagonal entry is dropped", neg, True
    if n == 6:
        e = np.zeros((6, 6)); e[:3, :3] = spd[:3, :3]
        yield "translations only (rotational dofs without stiffness)", e, True
        w = spd.copy(); w[3:, 3:] *= 1e-14; w[:3, 3:] *= 1e-7; w[3:, :3] *= 1e-7
        yield "nearly uncoupled weak rotations", w, True


def oracle_pinv(M):
    n = M.shape[0]
    return O.calc_dinv(O.Bsr(1, 1, n, n, [0, 1], [0], M.reshape(1, n, n)), None, pinv=True).reshape(n, n)


def product_pinv(M):
    import ctypes as C
    from ngsamg_b200 import _lib
    a = np.ascontiguousarray(M, np.float64).copy()
    L = _lib.lib()
    L.ngsamg_b200_block_pinv.argtypes = [C.c_int, C.c_void_p]
    assert L.ngsamg_b200_block_pinv(a.shape[0], a.ctypes.data_as(C.c_void_p)) == 0
    return a


@needs_ref
@pytest.mark.parametrize("n", [2, 3, 6])
def test_pseudo_inverse_vs_reference_code(n):
    """CalcPseudoInverseTryNormal(Mat<N,N>&): CallOnNonZeroDiagonalBlock + TryDirectInverse_simple + the eigenvalue fall-back
    (utils_denseLA.hpp:1237-1569, utils_denseLA.cpp:458-555) vs the oracle AND vs the product's host routine (dense.cpp):
    bit for bit wherever the direct inverse is taken, 1e-12 on the fall-back (LAPACK in NGSolve, Jacobi rotations everywhere here)."""
    for name, M, direct in pinv_cases(n):
        ref = R.pinv_block(M)
        for who, got in (("oracle", oracle_pinv(M)), ("product", product_pinv(M))):
            if direct:
                assert np.array_equal(ref, got), (who, n, name)
            else:
                assert np.abs(ref - got).max() <= 1e-12 * np.abs(ref).max(), (who, n, name)
        if name.startswith("spd"):
            assert np.abs(ref @ M - np.eye(n)).max() < 1e-12
    assert R.pinv_block(np.array([[1e-21]]))[0, 0] == 0.0 == oracle_pinv(np.array([[1e-21]]))[0, 0] == product_pinv(np.array([[1e-21]]))[0, 0]
    assert R.pinv_block(np.array([[4.0]]))[0, 0] == 0.25 == oracle_pinv(np.array([[4.0]]))[0, 0] == product_pinv(np.array([[4.0]]))[0, 0]


@needs_ref
def test_pinv_smoothers_through_the_hierarchy():
    """GSS3(..., pinv = true) on an elasticity hierarchy (what ngs_amg_regularize_cmats switches on): dinv and the V-cycle"""
    p, A, prols = hierarchy("elasticity")
    oa, ra = O.OracleAMG(A, p["free"], prols, pinv=True), R.RefAMG(A, p["free"], prols, pinv=True)
    for l in range(oa.nlevels - 1):
        assert np.array_equal(oa.level_dinv(l), ra.level_dinv(l)), l
    b = rand(21, A.nrows * A.bh)
    assert rel(oa.apply(b), ra.apply(b)) < 1e-13


def product_regularize(M, dim):
    import ctypes as C
    from ngsamg_b200 import _lib
    a = np.ascontiguousarray(M, np.float64).copy()
    L = _lib.lib()
    L.ngsamg_b200_block_regularize.argtypes = [C.c_int, C.c_void_p, C.c_int]
    assert L.ngsamg_b200_block_regularize(a.shape[0], a.ctypes.data_as(C.c_void_p), dim) == 0
    return a


@needs_ref
def test_coarse_regularisation_vs_reference_code():
    """RegularizeMatrix on the coarsest diagonal blocks (ngs_amg_regularize_cmats, elasticity_pc_impl.hpp:734-763 -> RegTM<0,6,6>,
    utils_denseLA.hpp:1198-1234): oracle and product (csrc/dense.cpp) vs the reference's own code"""
    rng = np.random.default_rng(3)
    Q, _ = np.linalg.qr(rng.standard_normal((6, 6)))
    X = rng.standard_normal((6, 6))
    cases = {"regular (left alone)": X @ X.T + 6 * np.eye(6),
             "three zero eigenvalues (a lone vertex: no rotational stiffness)": Q @ np.diag([0, 0, 0, 1.5, 2.0, 7.0]) @ Q.T,
             "translations only": np.diag([3.0, 2.0, 5.0, 0, 0, 0]),
             "one tiny eigenvalue": Q @ np.diag([1e-14, 1.0, 2.0, 3.0, 4.0, 5.0]) @ Q.T,
             "zero block becomes the identity": np.zeros((6, 6))}
    for name, M in cases.items():
        ref = R.regularize_block6(M)
        for who, got in (("oracle", O.regularize_block(M, 3)), ("product", product_regularize(M, 3))):
            assert np.abs(ref - got).max() <= 1e-12 * max(np.abs(ref).max(), 1.0), (who, name)
        if name.startswith("regular"):
            assert np.array_equal(ref, M)
        if name.startswith("translations"):
            assert np.allclose(ref, np.diag([3.0, 2.0, 5.0, 2.0, 2.0, 2.0]), atol=1e-12)     # smallest non-zero eigenvalue on the kernel
    # 2D: unit rotational entry (elasticity_pc_impl.hpp:721-728); no reference fragment needed for a one-liner, oracle == product
    M2 = np.diag([2.0, 3.0, 1e-9])
    assert O.regularize_block(M2, 2)[2, 2] == 1.0 == product_regularize(M2, 2)[2, 2]
    M3 = np.diag([2.0, 3.0, 1e-3])
    assert np.array_equal(O.regularize_block(M3, 2), M3) and np.array_equal(product_regularize(M3, 2), M3)


def test_regularised_coarse_solve_handles_lone_vertices():
    """an elasticity hierarchy whose coarsest level has a vertex without rotational stiffness: singular without RegularizeMatrix,
    solvable with it (what ngs_amg_regularize_cmats is for)"""
    p, A, prols = hierarchy("elasticity")
    P = prols[0]
    # cut the rotational columns of one coarse vertex out of the first prolongation -> its 6x6 diagonal block loses rank 3
    v = P.val.reshape(-1, 3, 6).copy()
    v[P.col == 0, :, 3:] = 0.0
    P0 = O.Bsr(P.nrows, P.ncols, 3, 6, P.rowptr, P.col, v)
    with pytest.raises(RuntimeError):
        O.OracleAMG(A, p["free"], [P0], pinv=True, regularize=False)
    oa = O.OracleAMG(A, p["free"], [P0], pinv=True)                     # regularize defaults to on with pinv on 6x6 coarse blocks
    x = oa.apply(rand(31, A.nrows * 3))
    assert np.isfinite(x).all() and np.abs(x).max() > 0


@needs_ref
def test_mult_quartet_vs_reference_code():
    """AMGMatrix::Mult / MultTrans / MultAdd / MultTransAdd and the Smooth dispatch on the cycle type (amg_matrix.hpp:37-43,
    amg_matrix.cpp:377-393): Mult overwrites x, MultAdd adds s * C b to x WITHOUT zeroing it, the Trans variants are aliases"""
    p, A, prols = hierarchy("poisson")
    oa, ra = O.OracleAMG(A, p["free"], prols, clev="none"), R.RefAMG(A, p["free"], prols, coarse_inv=False)
    b, x0 = rand(41, A.nrows), rand(42, A.nrows)
    for cyc in ("V", "W", "BS"):
        ref = oa.apply(b, cyc)
        for trans in (False, True):
            assert np.array_equal(ra.mult(b, x0.copy(), cyc, trans), ref)
            assert np.array_equal(ra.mult_add(-0.7, b, x0.copy(), cyc, trans), x0 + (-0.7) * ref)
    assert np.array_equal(oa.apply_add(-0.7, b, x0.copy()), ra.mult_add(-0.7, b, x0.copy()))


@needs_ref
def test_operator_complexity_accounting_of_the_reference():
    """AMGMatrix::GetOC + BaseSmoother / ProxySmoother GetNOps / GetANZE (amg_matrix.cpp:551-582, base_smoother.hpp:145-150,
    base_smoother.cpp:38-47): what the product's ngsamg_b200_operator_complexities restates (its GPU test compares the two)"""
    p, A, prols = hierarchy("poisson")
    ra = R.RefAMG(A, p["free"], prols, sm_steps=2, sm_symm=True)
    nze = [ra.level_matrix(l).nnz for l in range(ra.nlevels)]
    for cyc, fac in (("V", lambda l: 1.0), ("W", lambda l: 2.0 ** l), ("BS", lambda l: 2.0 * (1 + l))):
        occ = ra.get_oc(cyc)
        exp = [fac(l) * 2 * 2 * nze[l] / nze[0] for l in range(ra.nlevels - 1)] + [0.0]
        assert np.allclose(occ[1:], exp, rtol=1e-15) and np.isclose(occ[0], sum(exp), rtol=1e-15)
    assert len(R.RefAMG(A, p["free"], prols, coarse_inv=False).get_oc()) == ra.nlevels       # no entry for a level without inverse


@needs_ref
def test_elasticity_prolongation_blocks_are_the_reference_rigid_body_transport():
    """the piecewise prolongation of elast_3d: a fine vertex at x interpolates from its coarse vertex at c with the rigid-body transport
    Q(x - c) of the reference (EpsEpsEnergy::CalcQ, elasticity_energy_impl.hpp:8-29; CalcQHh: t = fine - coarse) -- the top three rows
    [I | -skew(t)] on the 3 -> 6 step (E_D = [I 0], elasticity_pc_impl.hpp:668-685), the full 6 x 6 block below.  Same sign convention,
    bit for bit, so prolongations dumped by the reference can be injected as they are."""
    import ngsamg_b200 as ng
    Q = R.elast_calcq([1.0, 2.0, 3.0])
    assert np.array_equal(Q[:3, 3:], -np.array([[0, -3.0, 2.0], [3.0, 0, -1.0], [-2.0, 1.0, 0]])) and np.array_equal(Q[3:, :3], np.zeros((3, 3)))
    p, A = elasticity(5, 4, 4)
    P, vmap, cxyz = ng.coarsen(A, p["free"], p["xyz"], bcoarse=6, max_per_row=4, smooth=False)
    blocks = P.val.reshape(-1, 3, 6)
    assert P.nnz > 0
    for i in range(P.nrows):
        for k in range(P.rowptr[i], P.rowptr[i + 1]):
            assert np.array_equal(blocks[k], R.elast_calcq(p["xyz"][i] - cxyz[P.col[k]])[:3, :]), (i, k)
    # 6 -> 6 step
    Po = to_oracle(P)
    A1 = O.restrict_matrix(O.transpose(Po), to_oracle(A), Po)
    A1p = ng.SparseMatrix(A1.nrows, A1.ncols, 6, 6, A1.rowptr, A1.col, A1.val)
    P1, _, cxyz1 = ng.coarsen(A1p, None, cxyz, bcoarse=6, max_per_row=4, smooth=False)
    b1 = P1.val.reshape(-1, 6, 6)
    assert P1.nnz > 0
    for i in range(P1.nrows):
        for k in range(P1.rowptr[i], P1.rowptr[i + 1]):
            assert np.array_equal(b1[k], R.elast_calcq(cxyz[i] - cxyz1[P1.col[k]])), (i, k)
    # 2D: rotation about the out-of-plane axis
    assert np.array_equal(R.elast_calcq([2.0, 5.0]), np.array([[1.0, 0, -5.0], [0, 1.0, 2.0], [0, 0, 1.0]]))


def test_dense_block_routines_against_reference_made_fixture():
    """pseudo-inverse, coarse regularisation and rigid-body transport against tests/golden/refpin_dense.npz (written by the reference library,
    tests/golden/make_ref_golden.py): oracle and product host routines, no library needed"""
    g = np.load(os.path.join(GOLD, "refpin_dense.npz"))
    assert "utils_denseLA" in str(g["fragments"])
    for k in range(int(g["pinv_count"])):
        M, ref, direct = g["pinv_in_%d" % k], g["pinv_out_%d" % k], bool(g["pinv_direct_%d" % k])
        for got in (oracle_pinv(M), product_pinv(M)):
            if direct:
                assert np.array_equal(ref, got), k
            else:
                assert np.abs(ref - got).max() <= 1e-12 * np.abs(ref).max(), k
    for i in range(int(g["reg_count"])):
        M, ref = g["reg_in_%d" % i], g["reg_out_%d" % i]
        for got in (O.regularize_block(M, 3), product_regularize(M, 3)):
            assert np.abs(ref - got).max() <= 1e-12 * max(np.abs(ref).max(), 1.0), i
    for t, Q in zip(g["calcq_t"], g["calcq_q"]):
        sk = np.array([[0, -t[2], t[1]], [t[2], 0, -t[0]], [-t[1], t[0], 0]])
        assert np.array_equal(Q[:3, :3], np.eye(3)) and np.array_equal(Q[:3, 3:], -sk) and np.array_equal(Q[3:, 3:], np.eye(3))
