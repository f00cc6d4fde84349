"""CPU checks of the restated block Gauss-Seidel smoother (oracle/oracle_bgs.py; BSmoother2, loc_block_gssmoother_impl.hpp)"""
import numpy as np

import ngsamg_b200 as ng
from helpers import poisson, rand, rel, to_oracle
from oracle import oracle as O
from oracle import oracle_bgs as OB


def _setup(n=6):
    p, A = poisson(n)
    return p, A, A.to_scipy().tocsr()


def test_singleton_blocks_are_point_gauss_seidel():
    """blocks of one vertex each: BSBlock::RichardsonUpdate is GSS3's row update (gssmoother.cpp:195-257), forward and backward"""
    p, A, As = _setup()
    n = p["n"]
    blk = np.where(p["free"] > 0, np.cumsum(p["free"].astype(np.int64)) - 1, -1)
    g = OB.BlockGS(As, 1, blk)
    x0, b = rand(1, n), rand(2, n)
    for back in (False, True):
        x = x0.copy()
        g.smooth_simple(x, b, reverse=back)
        # independent literal point sweep
        xs = x0.copy()
        order = range(n - 1, -1, -1) if back else range(n)
        for i in order:
            if p["free"][i]:
                xs[i] += (b[i] - As[i].dot(xs)[0]) / As[i, i]
        assert rel(x, xs) < 1e-14


def test_res_form_equals_rhs_form_and_keeps_the_residual():
    """RichardsonUpdate_RES == RichardsonUpdate when res = b - A x on entry (symmetric A), and leaves res = b - A x_new"""
    p, A, As = _setup(7)
    n = p["n"]
    blk = np.where(p["free"] > 0, (np.arange(n) // 5), -1)
    # make the block ids contiguous
    ids = np.unique(blk[blk >= 0])
    remap = -np.ones(blk.max() + 2, np.int64)
    remap[ids] = np.arange(len(ids))
    blk = np.where(blk >= 0, remap[blk], -1)
    g = OB.BlockGS(As, 1, blk)
    x0, b = rand(3, n), rand(4, n)
    for back in (False, True):
        x1, r1 = x0.copy(), b - As @ x0
        g.smooth(x1, b, r1, True, True, False, back)
        x2, r2 = x0.copy(), np.zeros(n)
        g.smooth(x2, b, r2, False, True, False, back)
        assert rel(x1, x2) < 1e-13 and np.linalg.norm(r1 - r2) < 1e-13 * np.linalg.norm(b)
        assert np.linalg.norm(r1 - (b - As @ x1)) < 1e-13 * np.linalg.norm(b)
        # non-block (Dirichlet) vertices are never touched
        assert np.array_equal(x1[blk < 0], x0[blk < 0])


def test_block_sweep_is_a_point_sweep_on_the_prescaled_matrix():
    """the identity the CUDA path rests on (csrc/amg.cu setup_bgs): with A~ = DB^-1 A the block update of B is the point update of its
    rows on A~ (unit diagonal, no couplings inside a block) against DB^-1 b"""
    import scipy.sparse as sp
    p, A, As = _setup(6)
    n = p["n"]
    blk = np.where(p["free"] > 0, np.arange(n) // 4, -1)
    ids = np.unique(blk[blk >= 0])
    remap = -np.ones(blk.max() + 2, np.int64)
    remap[ids] = np.arange(len(ids))
    blk = np.where(blk >= 0, remap[blk], -1)
    g = OB.BlockGS(As, 1, blk)
    Dinv = sp.lil_matrix((n, n))
    for dofs, DB, DBinv in g.blocks:
        Dinv[np.ix_(dofs, dofs)] = DBinv
    Dinv = Dinv.tocsr()
    At = (Dinv @ As).toarray()
    bt = Dinv @ rand(5, n)
    b = rand(5, n)
    x0 = rand(6, n)
    x = x0.copy()
    g.smooth_simple(x, b)
    xs = x0.copy()
    for dofs, _, _ in g.blocks:
        new = {}
        for i in dofs:
            off = [j for j in range(n) if blk[j] != blk[i] or j == i]
            new[i] = bt[i] - sum(At[i, j] * xs[j] for j in off if j != i)
        for i in dofs:
            xs[i] = new[i]
    assert rel(x, xs) < 1e-12


# ---- pin: the restated updates against the reference's OWN code (oracle/_ref/libngsamg_ref_bgs.so) --------------------------------------
import pytest
from oracle.ref_pin import ref_bgs as RB

needs_ref = pytest.mark.skipif(not RB.available(), reason="oracle/_ref/libngsamg_ref_bgs.so not built (needs /root/reference)")


def _blocks(n, free, size):
    blk = np.where(np.asarray(free) > 0, np.arange(n) // size, -1)
    ids = np.unique(blk[blk >= 0])
    remap = -np.ones(int(blk.max()) + 2, np.int64)
    remap[ids] = np.arange(len(ids))
    return np.where(blk >= 0, remap[blk], -1)


@needs_ref
@pytest.mark.parametrize("problem", ["poisson", "elasticity"])
def test_restated_updates_match_the_reference_code(problem):
    """BSBlock::RichardsonUpdate and RichardsonUpdate_RES (loc_block_gssmoother_impl.hpp:244-268, 516-541), compiled verbatim, against
    oracle_bgs.BlockGS: both forms, forward and backward block order, scalar and 3x3 blocks, blocks of several vertices; the two sides
    share numpy's dense block inverses, so they differ by summation order only"""
    from helpers import elasticity
    if problem == "poisson":
        p, A = poisson(6)
        b = 1
    else:
        p, A = elasticity(5, 3, 4)
        b = 3
    n = p["n"]
    As = A.to_scipy().tocsr()
    blk = _blocks(n, p["free"], 5)
    g = OB.BlockGS(As, b, blk)
    x0, rhs = rand(7, n * b), rand(8, n * b)
    for back in (False, True):
        # RHS form
        xo = x0.copy()
        g.smooth_simple(xo, rhs, reverse=back)
        xr, rr = x0.copy(), rhs.copy()
        RB.sweep(A, blk, xr, rr, res_form=False, reverse=back)
        assert rel(xo, xr) < 1e-13 and np.array_equal(rr, rhs)
        # RES form
        xo, ro = x0.copy(), rhs - As @ x0
        g.smooth_res_simple(xo, ro, reverse=back)
        xr, rr = x0.copy(), rhs - As @ x0
        RB.sweep(A, blk, xr, rr, res_form=True, reverse=back)
        assert rel(xo, xr) < 1e-13
        assert np.linalg.norm(ro - rr) < 1e-13 * max(np.linalg.norm(rhs), np.linalg.norm(rhs - As @ x0))


@needs_ref
@pytest.mark.parametrize("problem", ["poisson", "elasticity"])
def test_smoother_protocol_and_block_order_match_the_reference_code(problem):
    """BSmoother2::SmoothWO with the reference's own IterateBlocks / SmoothSimple / SmoothRESSimple (loc_block_gssmoother_impl.hpp:617-706),
    compiled verbatim: every (res_updated, update_res) combination, forward and reverse block order, against oracle_bgs.BlockGS.smooth"""
    from helpers import elasticity
    if problem == "poisson":
        p, A = poisson(6)
        b = 1
    else:
        p, A = elasticity(5, 3, 4)
        b = 3
    n = p["n"]
    As = A.to_scipy().tocsr()
    blk = _blocks(n, p["free"], 4)
    g = OB.BlockGS(As, b, blk)
    x0, rhs = rand(9, n * b), rand(10, n * b)
    scale = max(np.linalg.norm(rhs), np.linalg.norm(rhs - As @ x0))
    for back in (False, True):
        for ru in (False, True):
            for ur in (False, True):
                r0 = rhs - As @ x0 if ru else rand(11, n * b)
                xo, ro = x0.copy(), r0.copy()
                g.smooth(xo, rhs, ro, ru, ur, False, back)
                xr, rr = x0.copy(), r0.copy()
                RB.smooth_wo(A, blk, xr, rhs, rr, ru, ur, False, reverse=back)
                assert rel(xo, xr) < 1e-13, (back, ru, ur)
                if ur:
                    assert np.linalg.norm(ro - rr) < 1e-13 * scale, (back, ru, ur)
                else:
                    assert np.array_equal(rr, r0)      # res is left alone when no update was asked for


@pytest.mark.parametrize("tag", ["h1", "el"])
def test_oracle_against_the_reference_made_fixture(tag):
    """tests/golden/refpin_bgs.npz (written by tests/golden/make_ref_golden.py from the reference's own SmoothWO): the oracle reproduces
    every stored smoother call -- runs on machines without the pin library"""
    import os
    import scipy.sparse as sp
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "refpin_bgs.npz"))
    b = int(g[tag + "_b"])
    rp, ci, val = g[tag + "_rowptr"], g[tag + "_col"], g[tag + "_val"]
    n = len(rp) - 1
    As = sp.bsr_matrix((val.reshape(-1, b, b), ci, rp), shape=(n * b, n * b)).tocsr()
    sm = OB.BlockGS(As, b, g[tag + "_blk"])
    x0, rhs = g[tag + "_x0"], g[tag + "_rhs"]
    scale = max(np.linalg.norm(rhs), np.linalg.norm(rhs - As @ x0))
    for back in (0, 1):
        for ru in (0, 1):
            for ur in (0, 1):
                key = "%s_%d%d%d" % (tag, back, ru, ur)
                x, r = x0.copy(), g[key + "_r0"].copy()
                sm.smooth(x, rhs, r, bool(ru), bool(ur), False, bool(back))
                assert rel(x, g[key + "_x"]) < 1e-13, key
                assert np.linalg.norm(r - g[key + "_r"]) < 1e-13 * scale, key


@needs_ref
@pytest.mark.parametrize("problem", ["poisson", "elasticity"])
def test_block_setup_by_the_reference_code(problem):
    """the blocks built by the reference's own BSBlock::SetFromSPMat (loc_block_gssmoother_impl.hpp:67-132: sorted dofnrs, off-block rows,
    dense diagonal block, its inverse) instead of the harness: the smoother calls still agree with the oracle (the dense inverse is the
    stand-in's Gauss-Jordan there and numpy's here: 1e-11)"""
    from helpers import elasticity
    if problem == "poisson":
        p, A = poisson(6)
        b = 1
    else:
        p, A = elasticity(5, 3, 4)
        b = 3
    n = p["n"]
    As = A.to_scipy().tocsr()
    blk = _blocks(n, p["free"], 4)
    g = OB.BlockGS(As, b, blk)
    x0, rhs = rand(12, n * b), rand(13, n * b)
    scale = max(np.linalg.norm(rhs), np.linalg.norm(rhs - As @ x0))
    for back in (False, True):
        for ru, ur in ((True, True), (False, True), (False, False)):
            r0 = rhs - As @ x0 if ru else np.zeros(n * b)
            xo, ro = x0.copy(), r0.copy()
            g.smooth(xo, rhs, ro, ru, ur, False, back)
            xr, rr = x0.copy(), r0.copy()
            RB.smooth_wo(A, blk, xr, rhs, rr, ru, ur, False, reverse=back, ref_setup=True)
            assert rel(xo, xr) < 1e-11, (back, ru, ur)
            if ur:
                assert np.linalg.norm(ro - rr) < 1e-11 * scale, (back, ru, ur)
