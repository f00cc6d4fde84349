"""the synthetic problem generators: stencil-based generators for large meshes against element-by-element assembly"""
import numpy as np

import ngsamg_b200 as ng
from ngsamg_b200 import synthetic as S


def test_p2_elasticity_stencil_equals_assembly():
    """BASELINE.json configs[2]: the nodal-P2 beam from the translation-invariant stencil == the element-assembled matrix"""
    a = S.elasticity3d_p2_kuhn(6, 4, 5)
    b = S.elasticity3d_p2_kuhn_stencil(6, 4, 5)
    assert a["n"] == b["n"] == 11 * 7 * 9
    assert np.array_equal(a["rowptr"], b["rowptr"]) and np.array_equal(a["col"], b["col"]) and np.array_equal(a["free"], b["free"])
    assert np.abs(a["val"] - b["val"]).max() < 1e-12 * np.abs(a["val"]).max()
    assert np.abs(a["rhs"] - b["rhs"]).max() < 1e-12 * np.abs(a["rhs"]).max()
    assert np.allclose(a["xyz"], b["xyz"])


def test_p2_elasticity_is_symmetric_and_keeps_the_rigid_body_modes():
    a = S.elasticity3d_p2_kuhn(5, 3, 4, clamp=())
    n = a["n"]
    A = ng.SparseMatrix(n, n, 3, 3, a["rowptr"], a["col"], a["val"]).to_scipy()
    assert abs(A - A.T).max() < 1e-13 * abs(A).max()
    X = a["xyz"]
    modes = [np.tile(e, n) for e in np.eye(3)]
    modes += [np.stack([-X[:, 1], X[:, 0], 0 * X[:, 0]], 1).ravel(), np.stack([0 * X[:, 0], -X[:, 2], X[:, 1]], 1).ravel(),
              np.stack([X[:, 2], 0 * X[:, 0], -X[:, 0]], 1).ravel()]
    for m in modes:
        assert np.abs(A @ m).max() < 1e-12 * abs(A).max() * np.abs(m).max()
    # the load integrates (0, x, 0): total force = int x over the beam
    lx = X[:, 0].max()
    assert abs(a["rhs"].reshape(-1, 3)[:, 1].sum() - 0.5 * lx * lx * X[:, 1].max() * X[:, 2].max()) < 1e-12


def test_jump_elasticity_stencil_equals_assembly():
    """BASELINE.json configs[4] workload: P1 elasticity with a modulus jumping by 1e4 on a checkerboard -- the cube-stencil generator
    (A = sum of E_cube * K_cube) == element assembly with the same piecewise constant modulus; uniform modulus == the uniform generator"""
    nx, ny, nz = 7, 5, 6
    h = 1.0 / (min(ny, nz) - 1)
    cm = S.checkerboard_modulus(2, 1e4)
    jump = lambda cx, cy, cz: cm(np.floor(cx / h + 1e-9).astype(int), np.floor(cy / h + 1e-9).astype(int), np.floor(cz / h + 1e-9).astype(int))
    a = S.elasticity3d_kuhn(nx, ny, nz, jump=jump)
    b = S.elasticity3d_kuhn_jump_stencil(nx, ny, nz, cm)
    assert np.array_equal(a["rowptr"], b["rowptr"]) and np.array_equal(a["col"], b["col"]) and np.array_equal(a["free"], b["free"])
    assert np.abs(a["val"] - b["val"]).max() < 1e-13 * np.abs(a["val"]).max()
    assert np.abs(a["val"]).max() > 1e3 * np.abs(S.elasticity3d_kuhn(nx, ny, nz)["val"]).max()      # the contrast is really there
    u = S.elasticity3d_kuhn_stencil(nx, ny, nz)
    c = S.elasticity3d_kuhn_jump_stencil(nx, ny, nz, lambda x, y, z: np.ones(len(x)))
    assert np.abs(u["val"] - c["val"]).max() < 1e-13 * np.abs(u["val"]).max() and np.allclose(u["rhs"], c["rhs"])


def test_box_partition_of_the_jump_problem_sums_to_the_global_matrix():
    """configs[4] on N ranks: the ranks' sub-assembled matrices (box_elasticity3d_jump) add up to the global jump problem, shared DOFs
    are listed consistently on both sides, the loads add up, the clamp is the global face x = 0"""
    import scipy.sparse as sp
    n, grid = 4, (2, 1, 2)
    px, py, pz = grid
    gd = tuple((n - 1) * q + 1 for q in grid)
    glob = S.elasticity3d_kuhn_jump_stencil(*gd, S.checkerboard_modulus(2, 1e4), h=1.0 / (max(gd) - 1))
    N = glob["n"]
    Ag = ng.SparseMatrix(N, N, 3, 3, glob["rowptr"], glob["col"], glob["val"]).to_scipy()
    acc = sp.csr_matrix((3 * N, 3 * N))
    rhs = np.zeros(3 * N)
    parts = [S.box_elasticity3d_jump(n, grid, r, box_cells=2) for r in range(px * py * pz)]
    gids = []
    for r, p in enumerate(parts):
        bx, by, bz = r % px, (r // px) % py, r // (px * py)
        ids = np.arange(n ** 3)
        gx, gy, gz = ids % n + bx * (n - 1), (ids // n) % n + by * (n - 1), ids // (n * n) + bz * (n - 1)
        g = gx + gd[0] * (gy + gd[1] * gz)
        gids.append(g)
        dof = (3 * g[:, None] + np.arange(3)[None, :]).ravel()
        Al = ng.SparseMatrix(p["n"], p["n"], 3, 3, p["rowptr"], p["col"], p["val"]).to_scipy().tocoo()
        acc = acc + sp.csr_matrix((Al.data, (dof[Al.row], dof[Al.col])), shape=(3 * N, 3 * N))
        rhs[dof] += p["rhs"]
        assert np.allclose(p["xyz"], glob["xyz"][g])
        assert np.array_equal(p["free"], glob["free"][g])
    assert abs(acc - Ag).max() < 1e-12 * abs(Ag).max()
    assert np.allclose(rhs, glob["rhs"])
    for r, p in enumerate(parts):                         # k-th shared DOF with rank q here == k-th shared DOF with rank r on q
        for q, e in zip(p["peers"], p["ex"]):
            eq = parts[q]["ex"][parts[q]["peers"].index(r)]
            assert np.array_equal(gids[r][e], gids[q][eq])
    assert sum(p["n_master"] for p in parts) == N
