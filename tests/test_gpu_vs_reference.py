"""GPU parity against the REFERENCE'S OWN CODE (`-m gpu`): the CUDA path through the C ABI vs oracle/_ref/libngsamg_ref.so (the
reference's function bodies compiled against an NGSolve container stand-in, oracle/ref_pin/README.md), no oracle in between.
The library is built where /root/reference exists and travels with the snapshot; without it these tests skip (the same checks
then run against the oracle, which tests/test_ref_pin.py pins to the library bit for bit)."""
import numpy as np
import pytest

import ngsamg_b200 as ng
from helpers import assert_same_pattern, elasticity, poisson, rand, rel, to_oracle, to_product
from oracle.ref_pin import ref as R

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not R.available(), reason="oracle/_ref/libngsamg_ref.so not present")]

TOL_VCYCLE = 1e-10   # north_star: residual-vector agreement <= 1e-10 relative per V-cycle in fp64


@pytest.fixture(scope="module")
def pois():
    p, A = poisson(13)
    pc = ng.h1_scal(A, p["free"], ngs_amg_max_coarse_size=20)
    ra = R.RefAMG(to_oracle(A), p["free"], [to_oracle(P) for P in pc.GetMap()])
    return p, A, pc, ra


def test_spgemm_and_transpose_bit_exact_vs_reference_code():
    import scipy.sparse as sp
    from oracle import oracle as O
    rng = np.random.default_rng(21)
    for ah, aw, bw in [(1, 1, 1), (6, 3, 6), (6, 6, 6)]:
        def rnd(n, m, h, w, dens):
            pat = sp.random(n, m, density=dens, random_state=rng, format="csr")
            pat.sort_indices()
            return O.Bsr(n, m, h, w, pat.indptr, pat.indices, rng.standard_normal((pat.nnz, h, w)))
        A, B = rnd(200, 150, ah, aw, 0.05), rnd(150, 180, aw, bw, 0.06)
        Cg, Cr = ng.matmul(to_product(A), to_product(B)), R.matmul(A, B)          # MatMultABImpl
        assert_same_pattern(Cg, Cr)
        assert np.array_equal(Cg.val, Cr.val)
        Tg, Tr = ng.transpose(to_product(A)), R.transpose(A)                       # TransposeSPMImpl
        assert_same_pattern(Tg, Tr)
        assert np.array_equal(Tg.val, Tr.val)


def test_galerkin_hierarchy_vs_reference_code(pois):
    p, A, pc, ra = pois
    assert pc.GetNLevels() == ra.nlevels >= 3
    for l in range(pc.GetNLevels()):
        Ag, Ar = pc.GetLevelMatrix(l), ra.level_matrix(l)                           # RestrictMatrix
        assert_same_pattern(Ag, Ar)
        assert rel(Ag.val, Ar.val) < 1e-12


@pytest.mark.parametrize("cycle", ["V", "W", "BS"])
def test_cycles_vs_reference_code(pois, cycle):
    """AMGMatrix::SmoothV / SmoothW / SmoothBS of the reference vs the CUDA cycle"""
    p, A, pc, ra = pois
    pcc = pc if cycle == "V" else ng.h1_scal(A, p["free"], ngs_amg_max_coarse_size=20, ngs_amg_mg_cycle=cycle)
    rac = ra if cycle == "V" else R.RefAMG(to_oracle(A), p["free"], [to_oracle(P) for P in pcc.GetMap()])
    for seed in (1, 2):
        b = rand(seed, p["n"])
        x = np.zeros(p["n"])
        pcc.Mult(b, x)
        assert rel(x, rac.apply(b, cycle)) < TOL_VCYCLE


@pytest.mark.parametrize("cycle,steps,symm", [("V", 1, False), ("W", 2, True), ("BS", 3, False)])
def test_operator_complexity_vs_reference_code(pois, cycle, steps, symm):
    """AMGMatrix::GetOC (amg_matrix.cpp:551-582) incl. the ProxySmoother and cycle factors"""
    p, A, pc, ra = pois
    pcc = ng.h1_scal(A, p["free"], ngs_amg_max_coarse_size=20, ngs_amg_mg_cycle=cycle, ngs_amg_sm_steps=steps, ngs_amg_sm_symm=symm)
    rac = R.RefAMG(to_oracle(A), p["free"], [to_oracle(P) for P in pcc.GetMap()], sm_steps=steps, sm_symm=symm)
    assert np.allclose(pcc.GetOC(), rac.get_oc(cycle), rtol=1e-14, atol=0)


def test_pcg_iterations_vs_reference_code(pois):
    p, A, pc, ra = pois
    cg = ng.CGSolver(mat=A, pre=pc, maxsteps=50, tol=1e-8)
    u = cg.Solve(p["rhs"])
    ur, itr, _ = ra.pcg(p["rhs"], tol=1e-8, maxsteps=50)
    assert cg.iterations == itr
    assert rel(np.asarray(u), ur) < 1e-8


def test_elasticity_3_to_6_vs_reference_code():
    p, A = elasticity(6, 4, 4)
    pc = ng.elast_3d(A, p["free"], vertex_xyz=p["xyz"], ngs_amg_max_coarse_size=4)
    assert pc.GetNLevels() >= 2 and pc.GetBlockSize(1) == 6
    # GSS3<Mat<3,3>> / GSS3<Mat<6,6>> with pinv (regularize_cmats is on by default for elast_3d), ProlMap<Mat<3,6>> / <Mat<6,6>>
    ra = R.RefAMG(to_oracle(A), p["free"], [to_oracle(P) for P in pc.GetMap()], pinv=True)
    b = rand(3, p["n"] * 3)
    x = np.zeros(p["n"] * 3)
    pc.Mult(b, x)
    assert rel(x, ra.apply(b)) < 1e-10             # the north_star bar (measured 3e-15 on hardware against the oracle)


@pytest.mark.parametrize("grid", [(1, 1, 2), (2, 2, 2)])
def test_multirank_vcycle_and_pcg_vs_reference_code(grid):
    """R ranks (threads) sharing the GPU vs RefParAMG: the reference's AMGMatrix::SmoothV per rank over its HybridGSSmoother, DCCMap
    exchanges and ProlMap (in-process MPI stand-in), fed the hierarchy read back from the product"""
    from ngsamg_b200 import parallel as par
    from ngsamg_b200 import synthetic as S
    from oracle import oracle as O
    parts = S.partition_poisson3d(13, 11, 15, grid=grid)
    Rn = len(parts)

    def build(r, comm):
        p = parts[r]
        A = ng.SparseMatrix(p["n"], p["n"], 1, 1, p["rowptr"], p["col"], p["val"])
        return par.h1_scal_par(A, par.Halo(p["peers"], p["ex"]), comm, p["free"], ngs_amg_max_coarse_size=15, ngs_amg_b200_ctr_nv=150)

    pcs = par.run_ranks(Rn, build)
    npar = pcs[0].GetNParallelLevels()
    assert npar >= 1
    prols = [[to_oracle(pcs[r].GetProlongation(l)) for r in range(Rn)] for l in range(npar)]
    halos = []
    for l in range(npar + 1):
        hs = [pcs[r].GetHalo(l) for r in range(Rn)]
        halos.append(([list(h.peers) for h in hs], [[np.asarray(e) for e in h.ex] for h in hs]))
    maps = [pcs[0].GetContractionMap(r) for r in range(Rn)]
    nprols = [to_oracle(P) for P in pcs[0].GetContracted().GetMap()]
    A0 = [O.Bsr(p["n"], p["n"], 1, 1, p["rowptr"], p["col"], p["val"]) for p in parts]
    ra = R.RefParAMG(A0, [p["free"] for p in parts], halos[0][0], halos[0][1], prols, halos, maps, nprols)
    b = [rand(40 + r, p["n"]) * p["free"] for r, p in enumerate(parts)]
    xr = ra.apply(b)

    def ap(r, comm):
        x = np.zeros(parts[r]["n"])
        pcs[r].Mult(b[r], x)
        return x

    got = par.run_ranks(Rn, ap)
    for r in range(Rn):
        assert rel(got[r], xr[r]) < TOL_VCYCLE, (r, rel(got[r], xr[r]))
    rhs = [p["rhs"] * p["free"] for p in parts]
    ur, itr, _ = ra.pcg(rhs, tol=1e-8, maxsteps=100)

    def solve(r, comm):
        x = np.zeros(parts[r]["n"])
        it, errs = pcs[r]._pcg(rhs[r], x, 1e-8, 100)
        return x, it

    sol = par.run_ranks(Rn, solve)
    for r in range(Rn):
        assert sol[r][1] == itr
        assert rel(sol[r][0], ur[r]) < 1e-8
