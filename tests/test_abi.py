"""CPU tests of the drop-in boundary (`-m "not gpu"`): the C-ABI library loads and exports every symbol include/*.h
declares, fails loudly without a device, and the host-side logic (flags, coarsening, DOF maps) behaves."""
import ctypes
import os
import re

import numpy as np
import pytest

import ngsamg_b200 as ng
from helpers import elasticity, host_hierarchy, poisson, to_oracle
from ngsamg_b200 import _lib
from oracle import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def test_header_symbols_exported():
    hdr = open(os.path.join(ROOT, "include", "ngsamg_b200.h")).read()
    declared = set(re.findall(r"\b(ngsamg_b200_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 25
    L = ctypes.CDLL(_lib.LIB_PATH)
    for sym in declared:
        assert hasattr(L, sym), "symbol %s declared in the header but not exported" % sym
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)


def test_no_torch_types_in_abi():
    hdr = open(os.path.join(ROOT, "include", "ngsamg_b200.h")).read()
    assert "torch" not in hdr and "at::" not in hdr and "#include <cuda" not in hdr


def test_product_never_imports_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "ngsamg_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                for pat in (r"^\s*(from|import)\s+oracle", r"#include.*oracle", r"libngsamg_oracle", r"dlopen.*oracle",
                            r"\borc_[a-z_]+\s*\("):
                    assert not re.search(pat, txt, re.M), (f, pat)


@pytest.mark.skipif(_has_gpu(), reason="checks the no-device failure mode")
def test_fails_loudly_without_device():
    p, A = poisson(5)
    with pytest.raises(ng.NgsAMGError, match="no CUDA device|CUDA"):
        ng.h1_scal(A, p["free"])
    with pytest.raises(ng.NgsAMGError):
        ng.rap(A, A)


def test_unknown_type_rejected_before_any_device_work():
    p, A = poisson(4)
    with pytest.raises(ng.NgsAMGError):
        ng.CreatePreconditioner("NgsAMG.no_such_pc", A, p["free"])


def test_coarsening_h1():
    p, A = poisson(13)
    P, vmap, _ = ng.coarsen(A, p["free"])
    assert (vmap[p["free"] == 0] == -1).all() and (vmap[p["free"] == 1] >= 0).all()
    sizes = np.bincount(vmap[vmap >= 0])
    assert sizes.max() <= 8 + 7 and sizes.min() >= 1          # pairs^3 (+ orphans joined)
    assert P.ncols == sizes.shape[0] and 4.0 < p["free"].sum() / P.ncols <= 9.0
    per_row = np.diff(P.rowptr)
    assert per_row.max() <= 3 and (per_row[p["free"] == 0] == 0).all()
    rows = per_row > 0
    assert np.allclose(np.asarray(P.to_scipy().sum(axis=1)).ravel()[rows], 1.0)   # constants preserved
    for r in np.flatnonzero(rows)[:200]:
        c = P.col[P.rowptr[r]:P.rowptr[r + 1]]
        assert (np.diff(c) > 0).all() and vmap[r] in c
    # deterministic
    P2, vmap2, _ = ng.coarsen(A, p["free"])
    assert np.array_equal(P.col, P2.col) and np.array_equal(P.val, P2.val) and np.array_equal(vmap, vmap2)
    # piecewise
    Pp, _, _ = ng.coarsen(A, p["free"], smooth=False)
    assert np.diff(Pp.rowptr).max() == 1 and set(np.unique(Pp.val)) == {1.0}


def test_coarsening_elasticity_rigid_body_modes():
    p, A = elasticity(7, 4, 4)
    P, vmap, cxyz = ng.coarsen(A, p["free"], p["xyz"], bcoarse=6, max_per_row=4)
    assert P.bh == 3 and P.bw == 6
    u0, w = np.array([0.3, -0.2, 0.5]), np.array([0.1, 0.7, -0.4])
    crbm = np.concatenate([u0 + np.cross(w, cxyz), np.tile(w, (len(cxyz), 1))], axis=1).ravel()
    frbm = (u0 + np.cross(w, p["xyz"])).ravel()
    rows = np.repeat(np.diff(P.rowptr) > 0, 3)
    assert np.abs((P.to_scipy() @ crbm - frbm)[rows]).max() < 1e-13     # check_kvecs analogue (base_factory.cpp:260-261)


def test_host_hierarchy_iteration_ceiling():
    """product coarsening + oracle cycle: CG iterations under the reference's own test ceiling (test_2d_lo.py:11)"""
    p, A = poisson(21)
    prols = host_hierarchy(A, p["free"])
    amg = O.OracleAMG(to_oracle(A), p["free"], [to_oracle(P) for P in prols])
    _, it, errs = amg.pcg(p["rhs"], tol=1e-8, maxsteps=50)
    assert it < 30 and len(prols) >= 2


def test_world_size_2_gloo_partition():
    """multi-process plumbing (gloo, world_size 2): independent subdomain problems, no data-path collective"""
    import subprocess
    import sys
    code = r'''
import os, sys
sys.path.insert(0, %r); sys.path.insert(0, os.path.join(%r, "tests"))
import torch, torch.distributed as dist
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:29533", rank=int(sys.argv[1]), world_size=2)
import bench
n = bench.rank_workload(64, dist.get_rank(), 2)
t = torch.tensor([float(n)], dtype=torch.float64)
dist.all_reduce(t)
mx = torch.tensor([1.0 + dist.get_rank()], dtype=torch.float64)
dist.all_reduce(mx, op=dist.ReduceOp.MAX)
assert t.item() == 2 * n and mx.item() == 2.0
dist.barrier()
print("ok", dist.get_rank())
''' % (ROOT, ROOT)
    procs = [subprocess.Popen([sys.executable, "-c", code, str(r)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT) for r in range(2)]
    outs = [p.communicate(timeout=120)[0].decode() for p in procs]
    assert all(p.returncode == 0 for p in procs), outs


def _build_cpp_test(tmp_path):
    import subprocess
    exe = os.path.join(str(tmp_path), "test_hpp")
    libdir = os.path.dirname(_lib.LIB_PATH)
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "test_hpp.cpp"),
                           "-o", exe, "-L", libdir, "-lngsamg_b200", "-Wl,-rpath," + libdir])
    return exe


def test_cpp_parallel_host_mirror(tmp_path):
    """tests/cpp/test_hpp_parallel.cpp: the multi-rank host path (amg::ParallelDofs, amg::Communicator, amg::DecomposeHybrid) driven from
    C++ with two std::thread ranks -- master lists, M/G split, assembled interface block, operator identity, stage order"""
    import subprocess
    exe = os.path.join(str(tmp_path), "test_hpp_parallel")
    libdir = os.path.dirname(_lib.LIB_PATH)
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "test_hpp_parallel.cpp"),
                           "-o", exe, "-L", libdir, "-lngsamg_b200", "-Wl,-rpath," + libdir, "-pthread"])
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, (r.returncode, r.stdout, r.stderr)
    assert "bad=0" in r.stdout


@pytest.mark.skipif(_has_gpu(), reason="checks the no-device failure mode of the C++ host mirror")
def test_cpp_host_mirror_compiles_and_fails_loudly(tmp_path):
    """include/ngsamg_b200.hpp (C++ mirror of BaseAMGPC / CGSolver / RestrictMatrix) links against the C ABI; without a
    device it raises amg::Exception("no CUDA device ...") -> exit code 3"""
    import subprocess
    exe = _build_cpp_test(tmp_path)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 3, (r.returncode, r.stdout, r.stderr)
    assert "no CUDA device" in r.stdout


@pytest.mark.gpu
def test_cpp_host_mirror_on_gpu(tmp_path):
    import subprocess
    exe = _build_cpp_test(tmp_path)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, (r.returncode, r.stdout, r.stderr)
    assert "rap_pattern_same=1" in r.stdout


def test_regularize_matrix_method_is_host_only():
    """RegularizeMatrix of the elasticity classes (python_amg.hpp:86-91) works on a matrix object without a device: it only needs the
    host-side block routine"""
    import ngsamg_b200 as ng
    import numpy as np
    rng = np.random.default_rng(0)
    X = rng.standard_normal((6, 6))
    spd = X @ X.T + 6 * np.eye(6)
    lone = np.zeros((6, 6)); lone[:3, :3] = spd[:3, :3]
    # 2 x 2 block matrix: diagonal blocks spd / lone, one off-diagonal block
    M = ng.SparseMatrix(2, 2, 6, 6, [0, 2, 3], [0, 1, 1], np.stack([spd, spd, lone]).reshape(-1))
    pc = object.__new__(ng.elast_3d)                       # no device: only the method and the library are needed
    from ngsamg_b200 import _lib
    pc._lib = _lib.lib()
    out = ng.elast_3d.RegularizeMatrix(pc, M).val.reshape(3, 6, 6)
    assert np.array_equal(out[0], spd) and np.array_equal(out[1], spd)            # regular diagonal block and the off-diagonal block: untouched
    assert np.allclose(out[2][3:, 3:], np.eye(3) * np.linalg.eigvalsh(spd[:3, :3]).min())   # lone vertex: smallest non-zero eigenvalue on the rotations
    h1 = object.__new__(ng.h1_scal)
    h1._lib = pc._lib
    before = M.val.copy()
    assert h1.RegularizeMatrix(M) is M and np.array_equal(M.val, before) and pc.GetNProcs() == 1      # H1: nothing to regularise


def test_vector_arguments_are_checked_before_they_reach_the_c_abi():
    """the C ABI reads n doubles through every vector pointer: float32 / strided / short arrays must raise, not corrupt memory"""
    from ngsamg_b200 import _lib
    ok = np.zeros(10)
    assert _lib.vec(ok, 10).value == ok.ctypes.data
    with pytest.raises(TypeError):
        _lib.vec(np.zeros(10, np.float32), 10)
    with pytest.raises(ValueError):
        _lib.vec(np.zeros(20)[::2], 10)
    with pytest.raises(ValueError):
        _lib.vec(np.zeros(9), 10)
    with pytest.raises(TypeError):
        _lib.vec([0.0] * 10, 10)
    assert _lib.vec(None, 10) is None


def test_malformed_matrices_are_rejected():
    """check_csr: row pointers monotone, columns in range and strictly ascending (host-only entry point, no device needed)"""
    import ngsamg_b200 as ng
    rp = np.array([0, 2, 4], np.int64)
    good = ng.SparseMatrix(2, 2, 1, 1, rp, np.array([0, 1, 0, 1], np.int32), np.ones(4))
    ng.coarsen(good, None)
    for col in ([1, 0, 0, 1], [0, 0, 0, 1], [0, 2, 0, 1], [0, -1, 0, 1]):
        bad = ng.SparseMatrix(2, 2, 1, 1, rp, np.array(col, np.int32), np.ones(4))
        with pytest.raises(Exception):
            ng.coarsen(bad, None)
    bad = ng.SparseMatrix(2, 2, 1, 1, np.array([0, 3, 2], np.int64), np.array([0, 1, 0, 1], np.int32), np.ones(4))
    with pytest.raises(Exception):
        ng.coarsen(bad, None)


def _build_adapter_test(tmp_path):
    import subprocess
    exe = os.path.join(str(tmp_path), "test_adapter")
    libdir = os.path.dirname(_lib.LIB_PATH)
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(ROOT, "tests", "cpp"),
                           os.path.join(ROOT, "tests", "cpp", "test_adapter.cpp"), "-o", exe, "-L", libdir, "-lngsamg_b200", "-Wl,-rpath," + libdir])
    return exe


def test_ngsolve_adapter_compiles_and_registers_the_reference_names(tmp_path):
    """include/ngsamg_b200_ngsolve.hpp -- the reference-side adapter (B200AMGPC : ngcomp::Preconditioner, B200AMGMatrix : BaseMatrix,
    RegisterPreconditioner) -- compiles against the NGSolve stand-in, links against the C ABI, and its static initialisers put
    "NgsAMG.h1_scal" / "NgsAMG.elast_3d" (amg_register.hpp:79-98, elasticity.hpp:104-140) into the preconditioner registry"""
    import subprocess
    exe = _build_adapter_test(tmp_path)
    r = subprocess.run([exe, "--list"], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0, (r.stdout, r.stderr)
    assert set(r.stdout.split()) == {"NgsAMG.h1_scal", "NgsAMG.elast_3d"}


@pytest.mark.skipif(_has_gpu(), reason="checks the no-device failure mode of the adapter")
def test_ngsolve_adapter_fails_loudly_without_a_device(tmp_path):
    import subprocess
    r = subprocess.run([_build_adapter_test(tmp_path)], capture_output=True, text=True, timeout=60)
    assert r.returncode == 2 and "no CUDA device" in r.stdout, (r.returncode, r.stdout, r.stderr)


@pytest.mark.gpu
def test_ngsolve_adapter_on_the_device(tmp_path):
    """the adapter driven like NGSolve drives a registered preconditioner (registry lookup, InitLevel, FinalizeLevel, Mult / MultAdd /
    MultTrans through BaseMatrix, PCG): identical to direct C-ABI calls, PCG converges, wrong entry type throws ngcore::Exception"""
    import subprocess
    r = subprocess.run([_build_adapter_test(tmp_path)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, (r.returncode, r.stdout, r.stderr)
    assert "adapter: levels=" in r.stdout
