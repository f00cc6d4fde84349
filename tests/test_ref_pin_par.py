"""Pins the MULTI-RANK oracle (oracle/oracle_par.py: HybridLevel) against the reference's own code (`-m "not gpu"`).

oracle/_ref/libngsamg_ref.so also holds the reference's multi-rank functions -- BasicDCCMap::CalcDOFMasters, the DCCMap exchanges,
DecomposeSparseMatrixHybrid, MyAllReduceDofData, CalcHybridSmootherRDGItGeneric, GSS3 range sweeps, GSS4, HybridGSSmoother::Finalize /
SmoothStage*, HybridBaseSmoother::SmoothImpl* / CallStageKernelsImpl -- cut out of /root/reference at build time and compiled against
a threaded MPI stand-in (R ranks = R host threads, oracle/ref_pin/ngs_standin_mpi.hpp).  Everything is compared BIT FOR BIT:
M and G of the hybrid split, master flags, m_ex / g_ex lists, the inverted modified diagonal, the split index, and x / res of every
smoother call for all 16 combinations of the protocol flags, with and without overlapped exchange, on 2, 3, 4 and 8 ranks
(dofs shared by up to 8 ranks) and on 3x3 elasticity blocks.
"""
import os

import numpy as np
import pytest

from ngsamg_b200 import synthetic as S
from oracle import oracle as O
from oracle import oracle_par as OP
from oracle.ref_pin import ref as R

GOLD = os.path.join(os.path.dirname(__file__), "golden")
needs_ref = pytest.mark.skipif(not R.available(), reason="oracle/_ref/libngsamg_ref.so not built and /root/reference not present")
FLAGS = [(ru, ur, xz, bw) for ru in (0, 1) for ur in (0, 1) for xz in (0, 1) for bw in (0, 1)]


def level_args(parts, b):
    A = [O.Bsr(p["n"], p["n"], b, b, p["rowptr"], p["col"], p["val"]) for p in parts]
    return A, [p["free"] for p in parts], [p["peers"] for p in parts], [p["ex"] for p in parts]


def smoother_inputs(HL, parts, b, seed, ru, xz):
    """x CUMULATED (consistent on shared dofs), b DISTRIBUTED, res = b - A x (DISTRIBUTED) when res_updated"""
    rng = np.random.default_rng(seed)
    nglob = 1 + max(int(p["gidx"].max()) for p in parts)
    xg = rng.standard_normal((nglob, b))
    x = [np.zeros(p["n"] * b) if xz else np.ascontiguousarray(xg[p["gidx"]].reshape(-1)) for p in parts]
    rhs = [rng.standard_normal(p["n"] * b) for p in parts]
    if ru:
        y = HL.mult(x)
        res = [rhs[r] - y[r] for r in range(len(parts))]
    else:
        res = [rng.standard_normal(p["n"] * b) for p in parts]
    return x, rhs, res


def check_setup(HL, get_M, get_G, get_info, get_lists, R_):
    for r in range(R_):
        M, G = get_M(r), get_G(r)
        assert abs(M.to_scipy() - HL.M[r]).max() == 0, "M differs on rank %d" % r
        if G is None:
            assert HL.G[r].nnz == 0
        else:
            assert abs(G.to_scipy() - HL.G[r]).max() == 0, "G differs on rank %d" % r
            Gb = HL.Gb[r]
            assert np.array_equal(G.rowptr, Gb.rowptr) and np.array_equal(G.col, Gb.col)
        split, master, dinv = get_info(r)
        assert np.array_equal(master.astype(bool), HL.master[r])
        assert np.array_equal(dinv, HL.dinv[r]), "inverted modified diagonal differs on rank %d" % r
        m1, mex, m2 = HL.masks[r]
        loc = (m1 | m2).astype(bool)
        assert not m1[split:].any() and not m2[:split].any(), "split index"
        if loc.any() and HL.free[r] is not None:
            idx = np.flatnonzero(loc)
            assert split == idx[len(idx) // 2]
        m_ex, g_ex = get_lists(r)
        for a, b_ in zip(m_ex, HL.m_ex[r]):
            assert np.array_equal(a, b_)
        for a, b_ in zip(g_ex, HL.g_ex[r]):
            assert np.array_equal(a, b_)


@needs_ref
@pytest.mark.parametrize("grid", [(1, 1, 2), (1, 1, 3), (1, 2, 2), (2, 2, 2)])
@pytest.mark.parametrize("overlap", [True, False])
def test_hybrid_level_bit_exact_vs_reference_code(grid, overlap):
    parts = S.partition_poisson3d(7, 6, 9, grid=grid)
    args = level_args(parts, 1)
    HL, RL = OP.HybridLevel(*args), R.RefHybridLevel(*args, overlap=overlap)
    check_setup(HL, RL.M, RL.G, RL.info, RL.dcc_lists, len(parts))
    for k, (ru, ur, xz, bw) in enumerate(FLAGS):
        x, rhs, res = smoother_inputs(HL, parts, 1, 100 + k, ru, xz)
        x2, rhs2, res2 = [v.copy() for v in x], [v.copy() for v in rhs], [v.copy() for v in res]
        HL.smooth(x, rhs, res, ru, ur, xz, bw)
        RL.smooth(x2, rhs2, res2, ru, ur, xz, bw)
        for r in range(len(parts)):
            assert np.array_equal(x[r], x2[r]), ("x", grid, (ru, ur, xz, bw), r)
            assert np.array_equal(rhs[r], rhs2[r]), "the right-hand side must come back untouched"
            if ur:
                assert np.array_equal(res[r], res2[r]), ("res", grid, (ru, ur, xz, bw), r)


@needs_ref
def test_hybrid_level_blocks_vs_reference_code():
    """3x3 elasticity blocks: Mat<3,3> modified diagonal (max over the block rows), GSS3/GSS4 on blocks"""
    parts = S.partition_elasticity3d(5, 4, 7, 2)
    args = level_args(parts, 3)
    HL, RL = OP.HybridLevel(*args), R.RefHybridLevel(*args)
    for r in range(2):
        assert abs(RL.M(r).to_scipy() - HL.M[r]).max() == 0 and abs(RL.G(r).to_scipy() - HL.G[r]).max() == 0
        _, master, dinv = RL.info(r)
        assert np.array_equal(master.astype(bool), HL.master[r])
        scale = np.abs(HL.dinv[r]).max()
        assert np.abs(dinv - HL.dinv[r]).max() < 1e-13 * scale      # block inverse: CalcInverse is NGSolve's (stand-in: Gauss-Jordan)
    for k, (ru, ur, xz, bw) in enumerate(FLAGS):
        x, rhs, res = smoother_inputs(HL, parts, 3, 200 + k, ru, xz)
        x2, res2 = [v.copy() for v in x], [v.copy() for v in res]
        HL.smooth(x, rhs, res, ru, ur, xz, bw)
        RL.smooth(x2, rhs, res2, ru, ur, xz, bw)
        for r in range(2):
            assert np.linalg.norm(x[r] - x2[r]) <= 1e-12 * np.linalg.norm(x[r])
            if ur:
                assert np.linalg.norm(res[r] - res2[r]) <= 1e-12 * np.linalg.norm(res[r])


@needs_ref
def test_dcc_exchanges_and_hybrid_mult_vs_reference_code():
    parts = S.partition_poisson3d(6, 7, 8, grid=(2, 2, 2))
    args = level_args(parts, 1)
    HL, RL = OP.HybridLevel(*args), R.RefHybridLevel(*args)
    rng = np.random.default_rng(5)
    v = [rng.standard_normal(p["n"]) for p in parts]
    v2 = [a.copy() for a in v]
    HL.dis2co(v)
    RL.dis2co(v2)                       # StartDIS2CO, ApplyDIS2CO, FinishDIS2CO
    assert all(np.array_equal(a, b) for a, b in zip(v, v2))
    for r, p in enumerate(parts):
        assert not v[r][~HL.master[r]].any(), "CONCENTRATED: ghosts hold zero"
    HL.co2cu(v)
    RL.co2cu(v2)                        # StartCO2CU, ApplyCO2CU, FinishCO2CU
    assert all(np.array_equal(a, b) for a, b in zip(v, v2))
    y, y2 = HL.mult(v), RL.mult(v2)     # HybridBaseMatrix::Mult
    assert all(np.array_equal(a, b) for a, b in zip(y, y2))


@needs_ref
def test_symmetric_local_stages_run():
    """symm_loc (F,-,F / -,FB,- / B,-,B stage table, gssmoother.cpp:722-745) is not restated by the oracle; it must at least be a
    convergent smoother in the reference code: the residual of A x = b drops"""
    parts = S.partition_poisson3d(7, 6, 9, grid=(1, 1, 2))
    args = level_args(parts, 1)
    HL, RL = OP.HybridLevel(*args), R.RefHybridLevel(*args, symm_loc=True)
    x = [np.zeros(p["n"]) for p in parts]
    rhs = [p["rhs"] * p["free"] for p in parts]
    res = [v.copy() for v in rhs]
    n0 = np.sqrt(sum(np.dot(HL_r, HL_r) for HL_r in _cumulate(HL, res)))
    for _ in range(5):
        RL.smooth(x, rhs, res, True, True, False, False)
    n1 = np.sqrt(sum(np.dot(v, v) for v in _cumulate(HL, res)))
    assert n1 < 0.8 * n0 and np.isfinite(n1)


def _cumulate(HL, vec):
    """the master parts of the cumulated vector (for norms)"""
    v = [a.copy() for a in vec]
    HL.dis2co(v)
    return [np.where(np.asarray(HL.free[r], bool), v[r], 0.0) if HL.free[r] is not None else v[r] for r in range(HL.R)]


# ---- fixture written by the reference library (tests/golden/make_ref_golden.py), checked on any machine -----------------------
def test_multirank_oracle_against_reference_made_fixture():
    g = np.load(os.path.join(GOLD, "refpin_hybrid_2x2x2.npz"), allow_pickle=False)
    Rn = int(g["R"])
    assert "hybrid_base_smoother.cpp" in str(g["fragments"]) and "dcc_map.cpp" in str(g["fragments"])
    A, free, peers, ex = [], [], [], []
    for r in range(Rn):
        n = int(g["n%d" % r])
        A.append(O.Bsr(n, n, 1, 1, g["rowptr%d" % r], g["col%d" % r], g["val%d" % r]))
        free.append(g["free%d" % r])
        peers.append(list(g["peers%d" % r]))
        ptr = g["exptr%d" % r]
        ex.append([g["exdofs%d" % r][ptr[k]:ptr[k + 1]] for k in range(len(peers[-1]))])
    HL = OP.HybridLevel(A, free, peers, ex)

    def bsr(prefix):
        sh = g[prefix + "_shape"]
        return O.Bsr(int(sh[0]), int(sh[1]), 1, 1, g[prefix + "_rowptr"], g[prefix + "_col"], g[prefix + "_val"])

    for r in range(Rn):
        assert abs(bsr("M%d" % r).to_scipy() - HL.M[r]).max() == 0
        assert abs(bsr("G%d" % r).to_scipy() - HL.G[r]).max() == 0
        assert np.array_equal(g["master%d" % r].astype(bool), HL.master[r])
        assert np.array_equal(g["dinv%d" % r], HL.dinv[r])
    for ru, ur, xz, bw in FLAGS:
        key = "%d%d%d%d" % (ru, ur, xz, bw)
        x = [g["in_x_%s_%d" % (key, r)].copy() for r in range(Rn)]
        rhs = [g["in_b_%s_%d" % (key, r)].copy() for r in range(Rn)]
        res = [g["in_res_%s_%d" % (key, r)].copy() for r in range(Rn)]
        HL.smooth(x, rhs, res, ru, ur, xz, bw)
        for r in range(Rn):
            assert np.array_equal(x[r], g["out_x_%s_%d" % (key, r)]), key
            if ur:
                assert np.array_equal(res[r], g["out_res_%s_%d" % (key, r)]), key


@needs_ref
@pytest.mark.parametrize("grid,steps,symm", [((1, 1, 2), 1, False), ((2, 2, 2), 1, False), ((1, 2, 2), 2, True), ((1, 1, 3), 1, False)])
def test_parallel_vcycle_and_pcg_vs_reference_code(grid, steps, symm):
    """the whole multi-rank preconditioner: AMGMatrix::SmoothV of the reference over distributed levels with its HybridGSSmoother
    (+ProxySmoother) and ProlMap, its CtrMap (DoAssembleMatrix, TransferF2C / TransferC2F) onto rank 0 and its serial cycle below
    (RefParAMG) vs OracleParAMG on the same hierarchy.  Everything before the exact coarse solve is bit-identical on every level and rank; after it the two
    dense coarse solvers differ in the last bits; identical PCG iteration counts."""
    from oracle import cpu_pipeline as CP
    from helpers import rand, rel
    parts = S.partition_poisson3d(13, 11, 15, grid=grid)
    Rn = len(parts)
    hier, info = CP.build(parts, ctr_nv=150, max_coarse=15, engine="args")
    oa = OP.OracleParAMG(*hier, sm_steps=steps, sm_symm=symm)
    ra = R.RefParAMG(*hier, sm_steps=steps, sm_symm=symm)
    assert info["distributed_levels"] >= 2
    # CtrMap::DoAssembleMatrix: the contracted matrix is a structural union (entries that cancel across ranks stay) summed in group order
    Am, Ao = ra.A_merged, oa.A_merged
    assert np.array_equal(Am.rowptr, Ao.rowptr) and np.array_equal(Am.col, Ao.col), "contracted pattern"
    assert np.array_equal(Am.val, Ao.val), "contracted values"
    b = [rand(7 + r, p["n"]) * p["free"] for r, p in enumerate(parts)]
    xo, xr = oa.apply(b), ra.apply(b)
    for l in range(oa.npar):
        for r in range(Rn):
            assert np.array_equal(ra.level_vec("res", l, r), oa.level_res[l][r]), ("res", l, r)
            assert np.array_equal(ra.level_vec("rhs", l + 1, r), oa.level_rhs[l + 1][r]), ("rhs", l + 1, r)
    for r in range(Rn):
        assert rel(xr[r], xo[r]) < 1e-13
        for l in range(1, oa.npar + 1):
            assert rel(ra.level_vec("x", l, r), oa.level_x[l][r]) < 1e-13
    rhs = [p["rhs"] * p["free"] for p in parts]
    u1, it1, e1 = oa.pcg(rhs, tol=1e-8, maxsteps=60)
    u2, it2, e2 = ra.pcg(rhs, tol=1e-8, maxsteps=60)
    assert it1 == it2 and rel(e1, e2) < 1e-10


@needs_ref
def test_parallel_elasticity_hierarchy_vs_reference_code():
    """3x3 fine / 6x6 coarse blocks on two distributed levels (HybridGSSmoother<Mat<3,3>> / <Mat<6,6>>, ProlMap<Mat<3,6>>, DCCMap with block
    size 6, CtrMap<Vec<6>>): downward leg bit for bit, contracted matrix bit for bit"""
    from oracle import cpu_pipeline as CP
    from helpers import rand, rel
    parts = S.partition_elasticity3d(13, 5, 5, 2)
    hier, info = CP.build(parts, b=3, elast=True, ctr_nv=8, max_coarse=3, engine="args")
    assert info["distributed_levels"] == 2
    oa, ra = OP.OracleParAMG(*hier, pinv=False), R.RefParAMG(*hier)
    Am, Ao = ra.A_merged, oa.A_merged
    assert Am.bh == 6 and np.array_equal(Am.rowptr, Ao.rowptr) and np.array_equal(Am.col, Ao.col) and np.array_equal(Am.val, Ao.val)
    b = [rand(7 + r, p["n"] * 3) * np.repeat(p["free"], 3) for r, p in enumerate(parts)]
    xo, xr = oa.apply(b), ra.apply(b)
    for l in range(oa.npar):
        for r in range(2):
            assert np.array_equal(ra.level_vec("res", l, r), oa.level_res[l][r])
            assert np.array_equal(ra.level_vec("rhs", l + 1, r), oa.level_rhs[l + 1][r])
    assert max(rel(xr[r], xo[r]) for r in range(2)) < 1e-12


@needs_ref
@pytest.mark.parametrize("grid", [(1, 1, 2), (1, 1, 3), (2, 2, 2)])
def test_product_host_hybrid_split_vs_reference_code(grid):
    """the PRODUCT's host set-up of a distributed level (par.cpp: cumulate_matrix, hybrid_split, hybrid_mod_diag -- what the multi-GPU path
    uploads) directly against the reference's DecomposeSparseMatrixHybrid / CalcHybridSmootherRDG / BasicDCCMap: M and G with their patterns,
    master flags and the modified diagonal, bit for bit"""
    import ngsamg_b200 as ng
    from ngsamg_b200 import parallel as par
    parts = S.partition_poisson3d(7, 6, 9, grid=grid)
    Rn = len(parts)

    def fn(r, comm):
        p = parts[r]
        A = ng.SparseMatrix(p["n"], p["n"], 1, 1, p["rowptr"], p["col"], p["val"])
        return par.hybrid_host(A, par.Halo(p["peers"], p["ex"]), comm, p["free"])

    res = par.run_ranks(Rn, fn)
    RL = R.RefHybridLevel(*level_args(parts, 1))
    for r in range(Rn):
        M, G = RL.M(r), RL.G(r)
        pm, pg = res[r]["M"], res[r]["G"]
        assert np.array_equal(pm.rowptr, M.rowptr) and np.array_equal(pm.col, M.col) and np.array_equal(pm.val, M.val), "M on rank %d" % r
        if G is None:
            assert pg.nnz == 0
        else:
            assert np.array_equal(pg.rowptr, G.rowptr) and np.array_equal(pg.col, G.col) and np.array_equal(pg.val, G.val), "G on rank %d" % r
        _, master, dinv = RL.info(r)
        assert np.array_equal(res[r]["master"], master)
        md = res[r]["mod_diag"]
        live = (np.asarray(parts[r]["free"]) != 0) & (master != 0)
        assert np.array_equal(1.0 / md[live], dinv[live]) and not md[~live].any(), "modified diagonal on rank %d" % r
