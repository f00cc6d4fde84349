"""CPU tests (`-m "not gpu"`) of the multi-rank HOST logic: DCC master/ghost lists, hybrid M/G split, modified diagonal and the
stage order of the hybrid smoother (ngsamg_b200/csrc/par.cpp through the C ABI) against the multi-rank oracle -- with the ranks as
threads (ThreadComm) and as world_size-2 gloo processes (TorchDistComm) -- plus the partitioned problem generator."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest
import scipy.sparse as sp

import ngsamg_b200 as ng
from ngsamg_b200 import parallel as par
from ngsamg_b200 import synthetic as S
from oracle import oracle as O
from oracle import oracle_par as OP

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("grid", [(1, 1, 2), (1, 2, 2), (2, 2, 2)])
def test_partition_sums_to_global(grid):
    g = S.poisson3d_kuhn(7, 6, 9)
    A = sp.csr_matrix((g["val"], g["col"], g["rowptr"]), shape=(g["n"], g["n"]))
    B = sp.csr_matrix(A.shape)
    rhs = np.zeros(g["n"])
    parts = S.partition_poisson3d(7, 6, 9, grid=grid)
    for r, p in enumerate(parts):
        Al = sp.csr_matrix((p["val"], p["col"], p["rowptr"]), shape=(p["n"], p["n"])).tocoo()
        B = B + sp.coo_matrix((Al.data, (p["gidx"][Al.row], p["gidx"][Al.col])), shape=A.shape).tocsr()
        rhs += np.bincount(p["gidx"], weights=p["rhs"], minlength=g["n"])
        assert (p["free"] == g["free"][p["gidx"]]).all()
        # halo lists are pairwise consistent: k-th shared dof here == k-th shared dof there
        for kp, q in enumerate(p["peers"]):
            kq = list(parts[q]["peers"]).index(r)
            assert np.array_equal(p["gidx"][p["ex"][kp]], parts[q]["gidx"][parts[q]["ex"][kq]])
    assert abs(A - B).max() < 1e-15 and abs(rhs - g["rhs"]).max() < 1e-18


@pytest.mark.parametrize("grid", [(1, 1, 2), (1, 1, 3), (1, 2, 2), (2, 2, 2)])
def test_hybrid_split_threads(grid):
    parts = S.partition_poisson3d(7, 6, 9, grid=grid)
    R = len(parts)

    def fn(r, comm):
        p = parts[r]
        A = ng.SparseMatrix(p["n"], p["n"], 1, 1, p["rowptr"], p["col"], p["val"])
        return par.hybrid_host(A, par.Halo(p["peers"], p["ex"]), comm, p["free"])

    res = par.run_ranks(R, fn)
    HL = OP.HybridLevel([O.Bsr(p["n"], p["n"], 1, 1, p["rowptr"], p["col"], p["val"]) for p in parts], [p["free"] for p in parts],
                        [p["peers"] for p in parts], [p["ex"] for p in parts])
    for r in range(R):
        assert abs(res[r]["M"].to_scipy() - HL.M[r]).max() < 1e-14
        assert abs(res[r]["G"].to_scipy() - HL.G[r]).max() == 0
        assert np.allclose(res[r]["mod_diag"], HL.md[r].ravel(), rtol=1e-13, atol=0)
        assert (res[r]["master"].astype(bool) == HL.master[r]).all()
        m1, mex, m2 = HL.masks[r]
        exp = np.concatenate([np.flatnonzero(m1), np.flatnonzero(mex), np.flatnonzero(m2)])
        assert np.array_equal(np.argsort(res[r]["sweep_rank"])[:len(exp)], exp), "stage order LOC_PART_1 | EX_PART | LOC_PART_2"


def test_hybrid_split_blocks_threads():
    parts = S.partition_elasticity3d(5, 4, 7, 2)

    def fn(r, comm):
        p = parts[r]
        A = ng.SparseMatrix(p["n"], p["n"], 3, 3, p["rowptr"], p["col"], p["val"])
        return par.hybrid_host(A, par.Halo(p["peers"], p["ex"]), comm, p["free"])

    res = par.run_ranks(2, fn)
    HL = OP.HybridLevel([O.Bsr(p["n"], p["n"], 3, 3, p["rowptr"], p["col"], p["val"]) for p in parts], [p["free"] for p in parts],
                        [p["peers"] for p in parts], [p["ex"] for p in parts])
    for r in range(2):
        scale = abs(HL.M[r]).max()
        assert abs(res[r]["M"].to_scipy() - HL.M[r]).max() < 1e-14 * scale
        assert abs(res[r]["G"].to_scipy() - HL.G[r]).max() == 0
        assert np.allclose(res[r]["mod_diag"], HL.md[r].ravel(), rtol=1e-12, atol=1e-14 * scale)


def test_hybrid_operator_is_the_global_operator():
    """sum over ranks of (M_r + G_r) x_r == A x for a consistent (CUMULATED) x  -- HybridBaseMatrix::Mult, hybrid_matrix.cpp:433-453"""
    g = S.poisson3d_kuhn(7, 6, 9)
    A = sp.csr_matrix((g["val"], g["col"], g["rowptr"]), shape=(g["n"], g["n"]))
    parts = S.partition_poisson3d(7, 6, 9, grid=(1, 2, 2))
    HL = OP.HybridLevel([O.Bsr(p["n"], p["n"], 1, 1, p["rowptr"], p["col"], p["val"]) for p in parts], [p["free"] for p in parts],
                        [p["peers"] for p in parts], [p["ex"] for p in parts])
    x = np.random.default_rng(3).standard_normal(g["n"])
    y = HL.mult([x[p["gidx"]].copy() for p in parts])
    tot = np.zeros(g["n"])
    for p, yr in zip(parts, y):
        tot += np.bincount(p["gidx"], weights=yr, minlength=g["n"])
    assert np.allclose(tot, A @ x, rtol=1e-12, atol=1e-13)


def test_hybrid_split_over_gloo():
    """world_size-2 torch.distributed (gloo) run of the same host path: the communicator callbacks are backed by isend/irecv"""
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "_gloo_worker.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-4000:]
    assert "rank 0 ok" in out.stdout and "rank 1 ok" in out.stdout


@pytest.mark.parametrize("grid", [(1, 1, 2), (1, 2, 2), (2, 2, 2), (1, 1, 3)])
def test_parallel_coarsening_is_consistent_across_sharers(grid):
    """the class-respecting coarsening needs NO communication for P: every sharer of a DOF must compute the bit-identical
    prolongation row (same coarse vertices, same weights) and the coarse sharing lists must pair up entry by entry"""
    parts = S.partition_poisson3d(9, 8, 11, grid=grid)
    R = len(parts)

    def fn(r, comm):
        p = parts[r]
        A = ng.SparseMatrix(p["n"], p["n"], 1, 1, p["rowptr"], p["col"], p["val"])
        return par.coarsen_par(A, par.Halo(p["peers"], p["ex"]), comm, p["free"])

    res = par.run_ranks(R, fn)
    keys = []   # global identity of a coarse vertex: smallest global fine id among its members
    for r in range(R):
        P, vmap, _, _ = res[r]
        key = np.full(P.ncols, np.iinfo(np.int64).max)
        ok = vmap >= 0
        np.minimum.at(key, vmap[ok], parts[r]["gidx"][ok])
        keys.append(key)
        # aggregates never mix sharing classes
        nshare = np.zeros(parts[r]["n"], np.int64)
        for e in parts[r]["ex"]:
            nshare[e] += 1
        for c in range(P.ncols):
            mem = np.flatnonzero(vmap == c)
            assert len(set(nshare[mem])) == 1
    checked = 0
    for r in range(R):
        P, _, _, ch = res[r]
        for kp, q in enumerate(parts[r]["peers"]):
            kq = list(parts[q]["peers"]).index(r)
            Pq, _, _, chq = res[q]
            for dr, dq in zip(parts[r]["ex"][kp], parts[q]["ex"][kq]):
                rowr = {int(keys[r][P.col[k]]): P.val[k] for k in range(P.rowptr[dr], P.rowptr[dr + 1])}
                rowq = {int(keys[q][Pq.col[k]]): Pq.val[k] for k in range(Pq.rowptr[dq], Pq.rowptr[dq + 1])}
                assert rowr == rowq, "prolongation rows of a shared DOF differ between rank %d and rank %d" % (r, q)
                checked += 1
            if q in list(ch.peers):
                a, b = ch.ex[list(ch.peers).index(q)], chq.ex[list(chq.peers).index(r)]
                assert len(a) == len(b) and np.array_equal(keys[r][a], keys[q][b])
                assert np.all(np.diff(a) > 0) and np.all(np.diff(b) > 0)
            else:
                assert r not in list(chq.peers)
    assert checked > 0


@pytest.mark.parametrize("grid", [(1, 1, 2), (2, 2, 2)])
def test_cpu_multirank_pipeline_solves_the_global_problem(grid):
    """CPU-only restatement of the reference's MPI solve (host coarsening of the product + multi-rank oracle): the PCG solution of
    the partitioned problem equals a direct solve of the assembled problem; iteration count within the reference's own ceilings"""
    import scipy.sparse.linalg as spl
    from oracle import cpu_pipeline as CP
    dims = (13, 11, 15)
    parts = S.partition_poisson3d(*dims, grid=grid)
    amg, info = CP.build(parts, ctr_nv=150, max_coarse=15)
    assert info["distributed_levels"] >= 1
    rhs = [p["rhs"] * p["free"] for p in parts]
    u, it, errs = amg.pcg(rhs, tol=1e-8, maxsteps=100)
    assert it <= 30
    g = S.poisson3d_kuhn(*dims)
    A = sp.csr_matrix((g["val"], g["col"], g["rowptr"]), shape=(g["n"], g["n"]))
    f = g["free"].astype(bool)
    xg = np.zeros(g["n"])
    xg[f] = spl.spsolve(A[f][:, f].tocsc(), g["rhs"][f])
    for r, p in enumerate(parts):
        assert np.linalg.norm(u[r] - xg[p["gidx"]]) < 1e-7 * np.linalg.norm(xg)
    # threaded execution of the simulated ranks gives the same numbers (no dependence on thread timing)
    from oracle import oracle_par as OP
    OP.set_threads(4)
    try:
        u2, it2, _ = amg.pcg(rhs, tol=1e-8, maxsteps=100)
    finally:
        OP.set_threads(0)
    assert it2 == it and all(np.array_equal(a, b) for a, b in zip(u, u2))


def test_multirank_golden_fixture():
    """tests/golden/poisson_2ranks_9x8x11.npz (tests/golden/make_golden.py) pins the multi-rank oracle + the host-side parallel coarsening:
    hybrid split, modified diagonal, V-cycle result and PCG history of a 2-rank problem must reproduce"""
    from oracle import cpu_pipeline as CP
    path = os.path.join(ROOT, "tests", "golden", "poisson_2ranks_9x8x11.npz")
    if not os.path.exists(path):
        pytest.skip("golden fixture missing")
    g = np.load(path)
    parts = S.partition_poisson3d(9, 8, 11, grid=(1, 1, 2))
    amg, info = CP.build(parts, ctr_nv=100, max_coarse=15)
    assert info["distributed_levels"] == int(g["distributed_levels"])
    L0 = amg.levels[0]
    for r in range(2):
        M = sp.csr_matrix((g["m_data%d" % r], g["m_indices%d" % r], g["m_indptr%d" % r]), shape=L0.M[r].shape)
        G = sp.csr_matrix((g["g_data%d" % r], g["g_indices%d" % r], g["g_indptr%d" % r]), shape=L0.G[r].shape)
        assert abs(M - L0.M[r]).max() < 1e-15 and abs(G - L0.G[r]).max() == 0
        assert np.array_equal(g["master%d" % r].astype(bool), L0.master[r])
        assert np.allclose(g["mod_diag%d" % r], L0.md[r].ravel(), rtol=1e-14, atol=0)
    x = amg.apply([g["b0"], g["b1"]])
    for r in range(2):
        assert np.linalg.norm(x[r] - g["vcycle_x%d" % r]) <= 1e-13 * np.linalg.norm(g["vcycle_x%d" % r])
    u, it, errs = amg.pcg([p["rhs"] * p["free"] for p in parts], tol=1e-8, maxsteps=50)
    assert it == int(g["pcg_iters"]) and np.allclose(errs, g["pcg_errors"], rtol=1e-9)
    for r in range(2):
        assert np.linalg.norm(u[r] - g["pcg_u%d" % r]) <= 1e-10 * np.linalg.norm(g["pcg_u%d" % r])


@pytest.mark.parametrize("grid", [(1, 1, 2), (1, 2, 2), (2, 2, 2)])
def test_host_contraction_is_the_reference_ctrmap(grid):
    """contract_to_root (par.cpp), the host contraction of the multi-GPU path, against oracle_par.merge_contracted -- which
    tests/test_ref_pin_par.py pins bit for bit to the reference's CtrMap::DoAssembleMatrix (dof_contract.cpp:557-727): same pattern
    (structural union: entries that cancel across ranks stay), same values (members summed in rank order)"""
    parts = S.partition_poisson3d(7, 6, 9, grid=grid)
    R = len(parts)

    def fn(r, comm):
        p = parts[r]
        A = ng.SparseMatrix(p["n"], p["n"], 1, 1, p["rowptr"], p["col"], p["val"])
        return par.contract_host(A, par.Halo(p["peers"], p["ex"]), comm, p["free"])

    res = par.run_ranks(R, fn)
    merged, maps = res[0]
    assert all(res[r] == (None, None) for r in range(1, R))
    N = merged.nrows
    g = S.poisson3d_kuhn(7, 6, 9)
    assert N == g["n"]                                              # every global dof has exactly one master
    for r, p in enumerate(parts):                                   # shared dofs of different ranks map to the same merged dof
        for kp, q in enumerate(p["peers"]):
            kq = list(parts[q]["peers"]).index(r)
            assert np.array_equal(maps[r][p["ex"][kp]], maps[q][parts[q]["ex"][kq]])
    A_loc = [O.Bsr(p["n"], p["n"], 1, 1, p["rowptr"], p["col"], p["val"]) for p in parts]
    ref = OP.merge_contracted(A_loc, maps, N, 1)
    assert np.array_equal(merged.rowptr, ref.rowptr) and np.array_equal(merged.col, ref.col), "contracted pattern"
    assert np.array_equal(merged.val, ref.val), "contracted values"
    # and it is the assembled global operator up to the renumbering
    G = sp.csr_matrix((g["val"], g["col"], g["rowptr"]), shape=(N, N))
    perm = np.zeros(N, np.int64)
    for r, p in enumerate(parts):
        perm[p["gidx"]] = maps[r]
    Pm = sp.csr_matrix((np.ones(N), (perm, np.arange(N))), shape=(N, N))
    assert abs(merged.to_scipy() - Pm @ G @ Pm.T).max() < 1e-13
