"""CPU tests of the two-level (tile) Gauss-Seidel schedule (ngsamg_b200/csrc/tiles.cpp through the host-only C entry points):
the schedule must be a topological order of the row DAG of the reference's sequential sweep (gssmoother.cpp:195-315) -- checked by the
library's own verifier AND by emulating the tile kernel's data flow (kernels_tile.cuh: out-of-tile couplings first, then the tile-local
levels, tiles in schedule order) in numpy and comparing with the oracle's sequential sweep on the original matrix."""
import ctypes as C

import numpy as np
import pytest
import scipy.sparse as sp

import ngsamg_b200 as ng
from ngsamg_b200 import _lib
from ngsamg_b200 import synthetic as S
from helpers import poisson, rand, rel, to_oracle
from oracle import oracle as O

KEYS = ["ok", "ntiles", "npad", "nonfree_pad", "tile_depth", "max_local_levels", "merged", "violations", "npred"]


def tile_schedule(A, mask=None, rank=None, rounds=5, max_rows=32):
    L = _lib.lib()
    info = np.zeros(9, np.int64)
    h = C.c_void_p()
    m = None if mask is None else np.ascontiguousarray(mask, np.uint8)
    r = None if rank is None else np.ascontiguousarray(rank, np.int32)
    abi = A._abi()
    rc = L.ngsamg_b200_tile_schedule_begin(C.byref(abi), _lib.ptr(m), _lib.ptr(r), rounds, max_rows, C.byref(h), _lib.ptr(info))
    assert rc == 0, L.ngsamg_b200_tiles_last_error()
    d = dict(zip(KEYS, [int(v) for v in info]))
    perm = np.zeros(A.nrows, np.int32)
    ts = np.zeros(d["ntiles"] + 1, np.int32)
    nl = np.zeros(max(d["ntiles"], 1), np.int32)
    rl = np.zeros(max(d["npad"], 1), np.uint8)
    pp = np.zeros(d["ntiles"] + 1, np.int64)
    pr = np.zeros(max(d["npred"], 1), np.int32)
    L.ngsamg_b200_tile_schedule_fetch(h, _lib.ptr(perm), _lib.ptr(ts), _lib.ptr(nl), _lib.ptr(rl), _lib.ptr(pp), _lib.ptr(pr))
    d.update(perm=perm, tile_slice=ts, tile_nlev=nl, row_lvl=rl, pred_ptr=pp, pred=pr)
    return d


def emulate_forward(A, free, d, res):
    """forward triangular half-sweep on the tile schedule, data flow as in k_gs_tile<.., ADD_SELF=false, WRITE_R=true>:
    delta = (L + D)^-1 res in the sweep order, rout = res - (L + D) delta;  returns (delta, rout) in the ORIGINAL numbering"""
    n = A.shape[0]
    perm = d["perm"].astype(np.int64)
    inv = np.full(d["npad"], -1, np.int64)
    inv[perm] = np.arange(n)
    Ap = A.tocsr()
    diag = Ap.diagonal()
    out = np.zeros(n)
    rout = res.copy()
    done = np.zeros(d["ntiles"], bool)
    for t in range(d["ntiles"]):
        for q in d["pred"][d["pred_ptr"][t]:d["pred_ptr"][t + 1]]:
            assert done[q], "a tile runs before a tile it waits for"
        r0, r1 = d["tile_slice"][t] * 32, d["tile_slice"][t + 1] * 32
        rows = [r for r in range(r0, r1) if d["row_lvl"][r] != 255]
        acc = {}
        for r in rows:                                   # out-of-tile couplings: lower rows of OTHER tiles (final values)
            i = inv[r]
            a = res[i]
            for k in range(Ap.indptr[i], Ap.indptr[i + 1]):
                j = Ap.indices[k]
                pj = perm[j]
                if j != i and free[j] and pj < r and not (r0 <= pj < r1):
                    a -= Ap.data[k] * out[j]
            acc[r] = a
        for s in range(d["tile_nlev"][t]):               # tile-local levels
            for r in rows:
                if d["row_lvl"][r] != s:
                    continue
                i = inv[r]
                a = acc[r]
                for k in range(Ap.indptr[i], Ap.indptr[i + 1]):
                    j = Ap.indices[k]
                    pj = perm[j]
                    if j != i and free[j] and pj < r and r0 <= pj < r1:
                        assert d["row_lvl"][pj] < s, "in-tile dependency on the same or a later local level"
                        a -= Ap.data[k] * out[j]
                out[i] = a / diag[i]
                rout[i] = a - diag[i] * out[i]
        done[t] = True
    return out, rout


@pytest.mark.parametrize("cfg", [dict(rounds=5, max_rows=32), dict(rounds=6, max_rows=64), dict(rounds=3, max_rows=32),
                                 dict(rounds=-1, max_rows=256), dict(rounds=9, max_rows=512)])
def test_tile_schedule_is_a_valid_sweep_order(cfg):
    p = S.poisson3d_kuhn(13, 11, 9)
    A = ng.SparseMatrix(p["n"], p["n"], 1, 1, p["rowptr"], p["col"], p["val"])
    d = tile_schedule(A, p["free"], None, **cfg)
    assert d["ok"] == 1 and d["violations"] == 0
    assert d["nonfree_pad"] >= int((p["free"] == 0).sum()) and d["npad"] % 32 == 0
    # the tile DAG is much shallower than the row DAG of the natural ordering (13 + 11 + 9 wavefronts)
    if cfg["rounds"] >= 5:
        assert d["tile_depth"] < 25
    # emulated tile sweep == sequential sweep of the oracle (GSS3::SmoothRESInternal with x = 0)
    As = to_oracle(A).to_scipy()
    free = p["free"].astype(bool)
    res = rand(3, p["n"]) * p["free"]
    delta, rout = emulate_forward(As, free, d, res)
    x = np.zeros(p["n"])
    r2 = res.copy()
    dinv = O.calc_dinv(to_oracle(A), p["free"])
    O.gs_res(to_oracle(A), dinv, p["free"], x, r2, False)
    assert rel(delta[free], x[free]) < 1e-13
    # the oracle's residual holds res - A delta; ours only the (L + D) part: add the U part
    U = sp.triu(As, 1).tocsr()
    full = rout - U @ delta
    assert rel(full[free], r2[free]) < 1e-12


def test_tile_schedule_with_a_custom_sweep_order():
    """non-natural sweep orders (multicolour option, hybrid stage order LOC_PART_1 | EX_PART | LOC_PART_2): still valid"""
    p = S.poisson3d_kuhn(9, 9, 9)
    A = ng.SparseMatrix(p["n"], p["n"], 1, 1, p["rowptr"], p["col"], p["val"])
    rng = np.random.default_rng(5)
    rank = rng.permutation(p["n"]).astype(np.int32)
    d = tile_schedule(A, p["free"], rank)
    assert d["ok"] == 1 and d["violations"] == 0
    # stage-like order: second half of the rows first
    n = p["n"]
    rank2 = np.concatenate([np.arange(n // 2, n), np.arange(0, n // 2)]).astype(np.int32)
    inv = np.zeros(n, np.int32)
    inv[rank2] = np.arange(n, dtype=np.int32)
    d2 = tile_schedule(A, p["free"], inv)
    assert d2["ok"] == 1 and d2["violations"] == 0


def test_tile_schedule_all_rows_smoothed_and_tiny():
    p = S.poisson3d_kuhn(5, 4, 3, dirichlet=())
    A = ng.SparseMatrix(p["n"], p["n"], 1, 1, p["rowptr"], p["col"], p["val"])
    d = tile_schedule(A, None, None)
    assert d["ok"] == 1 and d["violations"] == 0 and d["nonfree_pad"] == 0
    assert sorted(d["perm"]) == sorted(set(d["perm"])) and d["perm"].max() < d["npad"]


def test_tile_schedule_with_cluster_hints_reaches_the_box_bound():
    """caller-supplied clusters (ngsamg_b200_tile_schedule_hinted): axis-aligned 4x4x2 boxes of the structured grid give the ideal tile-DAG
    depth sum_d ceil(n_d / e_d) - 2 (every lower neighbour of a box member lies in a box with smaller-or-equal box coordinates), and the
    schedule is still a valid sweep order; the built-in pairwise clustering stays within 1.6x of that bound"""
    n = 21
    p, A = poisson(n)
    idx = np.arange(n ** 3)
    x, y, z = idx % n, (idx // n) % n, idx // (n * n)
    bx, by, bz = -(-n // 4), -(-n // 4), -(-n // 2)
    cl = np.ascontiguousarray((x // 4) + bx * ((y // 4) + by * (z // 2)), np.int32)
    L = _lib.lib()
    abi = A._abi()
    h = C.c_void_p()
    info = np.zeros(9, np.int64)
    m = np.ascontiguousarray(p["free"], np.uint8)
    rc = L.ngsamg_b200_tile_schedule_hinted(C.byref(abi), _lib.ptr(m), None, _lib.ptr(cl), 32, C.byref(h), _lib.ptr(info))
    assert rc == 0, L.ngsamg_b200_tiles_last_error()
    d = dict(zip(KEYS, [int(v) for v in info]))
    L.ngsamg_b200_tile_schedule_fetch(h, None, None, None, None, None, None)
    assert d["ok"] == 1 and d["violations"] == 0
    assert d["tile_depth"] <= bx + by + bz - 2
    d2 = tile_schedule(A, p["free"], None)
    assert d2["tile_depth"] <= 1.6 * d["tile_depth"]


def emulate_backward(A, free, d, t_in, x_old):
    """backward triangular half-sweep on the tile schedule, data flow as in k_gs_tile<.., ADD_SELF=true, WRITE_R=false> with backward = 1:
    tiles in REVERSE schedule order waiting for their SUCCESSORS, tile-local levels descending;
    x_new_i = x_old_i + (t_i - sum_{U} A_ik x_new_k) / d_i   (post-smoother: t = b - (L + D) x_old)"""
    n = A.shape[0]
    perm = d["perm"].astype(np.int64)
    inv = np.full(d["npad"], -1, np.int64)
    inv[perm] = np.arange(n)
    Ap = A.tocsr()
    diag = Ap.diagonal()
    nt = d["ntiles"]
    succ = [[] for _ in range(nt)]                       # the kernel gets these from the scheduler (transpose of the predecessor lists)
    for t in range(nt):
        for q in d["pred"][d["pred_ptr"][t]:d["pred_ptr"][t + 1]]:
            succ[q].append(t)
    out = x_old.copy()
    done = np.zeros(nt, bool)
    for t in range(nt - 1, -1, -1):
        for q in succ[t]:
            assert done[q], "a tile runs before a tile it waits for (backward)"
        r0, r1 = d["tile_slice"][t] * 32, d["tile_slice"][t + 1] * 32
        rows = [r for r in range(r0, r1) if d["row_lvl"][r] != 255]
        acc = {}
        for r in rows:                                   # couplings to HIGHER rows of other tiles: final values
            i = inv[r]
            a = t_in[i]
            for k in range(Ap.indptr[i], Ap.indptr[i + 1]):
                j = Ap.indices[k]
                pj = perm[j]
                if j != i and free[j] and pj > r and not (r0 <= pj < r1):
                    a -= Ap.data[k] * out[j]
            acc[r] = a
        for s in range(d["tile_nlev"][t] - 1, -1, -1):   # tile-local levels, descending
            for r in rows:
                if d["row_lvl"][r] != s:
                    continue
                i = inv[r]
                a = acc[r]
                for k in range(Ap.indptr[i], Ap.indptr[i + 1]):
                    j = Ap.indices[k]
                    pj = perm[j]
                    if j != i and free[j] and pj > r and r0 <= pj < r1:
                        assert d["row_lvl"][pj] > s, "in-tile dependency on the same or an earlier local level (backward)"
                        a -= Ap.data[k] * out[j]
                out[i] = x_old[i] + a / diag[i]
        done[t] = True
    return out


@pytest.mark.parametrize("cfg", [dict(rounds=5, max_rows=32), dict(rounds=6, max_rows=64), dict(rounds=-1, max_rows=512)])
def test_tile_schedule_backward_sweep(cfg):
    """the same schedule run backwards (successor lists, descending local levels) == GSS3::SmoothRHSInternal(backwards) of the oracle"""
    p = S.poisson3d_kuhn(11, 13, 9)
    A = ng.SparseMatrix(p["n"], p["n"], 1, 1, p["rowptr"], p["col"], p["val"])
    d = tile_schedule(A, p["free"], None, **cfg)
    assert d["ok"] == 1 and d["violations"] == 0
    As = to_oracle(A).to_scipy().tocsr()
    free = p["free"].astype(bool)
    # tile-major numbering keeps every dependency's orientation: perm is monotone along every edge of the sweep DAG
    x_old, b = rand(5, p["n"]) * p["free"], rand(6, p["n"]) * p["free"]
    LD = sp.tril(As, 0).tocsr()
    t_in = b - LD @ x_old                                   # the parallel half of the post-smoother (one SpMV on the device)
    # couplings to non-free columns do not enter the sweep (their x is fixed): move them to the right-hand side like the kernel's N part
    nf = sp.diags((~free).astype(float))
    t_in = t_in - sp.triu(As, 1).tocsr() @ (nf @ x_old)
    x_new = emulate_backward(As, free, d, t_in, x_old)
    xo = x_old.copy()
    dinv = O.calc_dinv(to_oracle(A), p["free"])
    O.gs_rhs(to_oracle(A), dinv, p["free"], xo, b, True)
    assert rel(x_new[free], xo[free]) < 1e-13


@pytest.mark.parametrize("dims", [(23, 23, 23), (17, 9, 26), (40, 33, 1)])
def test_grid_numbered_matrices_get_box_tiles(dims):
    """rounds < 0 (the setup default): a matrix numbered like a structured grid is detected from its pattern alone (line length = period
    of the rows without a left neighbour, plane size likewise) and tiled into near-cubic boxes -> the ideal tile DAG sum_d ceil(ext_d / e) - 2"""
    p = S.poisson3d_kuhn(*dims)
    A = ng.SparseMatrix(p["n"], p["n"], 1, 1, p["rowptr"], p["col"], p["val"])
    d = tile_schedule(A, p["free"], None, rounds=-1, max_rows=512)
    assert d["ok"] == 1 and d["violations"] == 0
    free = p["free"].astype(bool).reshape(dims[::-1])
    ext = [int(np.any(free, axis=tuple(a for a in range(3) if a != ax)).sum()) for ax in (2, 1, 0)]
    nd = sum(e > 1 for e in ext)
    edge = int(np.floor(512 ** (1.0 / nd) + 1e-9))
    ideal = sum(-(-e // edge) for e in ext if e > 1) - (nd - 1)
    assert d["tile_depth"] == ideal, (d["tile_depth"], ideal, ext)
    assert d["max_local_levels"] <= nd * edge - (nd - 1)


def test_a_scrambled_numbering_falls_back_to_the_pairwise_clustering():
    """the grid detector must refuse a matrix that is not numbered like a grid (the schedule is still valid, built from the matching)"""
    p = S.poisson3d_kuhn(9, 9, 9)
    rng = np.random.default_rng(5)
    q = rng.permutation(p["n"])
    M = sp.csr_matrix((p["val"], p["col"], p["rowptr"]), shape=(p["n"], p["n"]))
    Mp = M[q][:, q].tocsr()
    Mp.sort_indices()
    A = ng.SparseMatrix(p["n"], p["n"], 1, 1, Mp.indptr.astype(np.int64), Mp.indices.astype(np.int32), Mp.data)
    d = tile_schedule(A, p["free"][q], None, rounds=-1, max_rows=512)
    assert d["ok"] == 1 and d["violations"] == 0
