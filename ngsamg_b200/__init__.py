"""ngsamg_b200 -- B200-native drop-in for NgsAMG's preconditioner-apply hot path.

Host-side mirror of the reference's Python surface (module `NgsAMG`, src/base/python/python_amg.cpp:37-63):
  * preconditioner classes `h1_scal`, `h1_2d`, `h1_3d`, `elast_2d`, `elast_3d` with the strict-algebraic
    constructor `(mat, freedofs=None, **kwargs)` (src/h1/python_h1.cpp:24-33), `ngs_amg_*` kwargs, and the
    NGSolve BaseMatrix quartet Mult / MultAdd / MultTrans / MultTransAdd (amg_matrix.cpp:377-393);
  * introspection `GetNLevels`, `GetNDof`, `GetBlockSize`, `GetSmoother`, `GetAMGMatrix`, `GetMap`
    (python_amg.hpp:30-101);
  * `CGSolver(mat, pre, maxsteps, tol)` with `.Solve(rhs)`, `.iterations`, `.errors` as the reference's tests use
    `ngsolve.krylovspace.CGSolver` (tests/h1/amg_utils.py:346-362).
Everything computes through the C ABI (include/ngsamg_b200.h) on a CUDA device; there is no CPU path.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import NgsAMGError, build  # noqa: F401

__all__ = ["SparseMatrix", "h1_scal", "h1_2d", "h1_3d", "elast_2d", "elast_3d", "Preconditioner", "CGSolver",
           "AMGMatrix", "Smoother", "NgsAMGError", "rap", "matmul", "transpose", "coarsen", "build"]


class SparseMatrix:
    """Block CSR with NGSolve SparseMatrix<Mat<bh,bw>> semantics (rowptr int64, sorted int32 cols, row-major blocks)."""

    def __init__(self, nrows, ncols, bh, bw, rowptr, col, val):
        self.nrows, self.ncols, self.bh, self.bw = int(nrows), int(ncols), int(bh), int(bw)
        self.rowptr = np.ascontiguousarray(rowptr, dtype=np.int64)
        self.col = np.ascontiguousarray(col, dtype=np.int32)
        self.val = np.ascontiguousarray(val, dtype=np.float64).reshape(-1)
        if self.rowptr.shape[0] != self.nrows + 1:
            raise ValueError("rowptr has the wrong length")
        if self.val.shape[0] != self.col.shape[0] * self.bh * self.bw:
            raise ValueError("val has the wrong length")

    height = property(lambda s: s.nrows)
    width = property(lambda s: s.ncols)

    @property
    def nnz(self):
        return int(self.rowptr[-1])

    @staticmethod
    def from_scipy(m, bh=1, bw=1):
        import scipy.sparse as sp
        if bh == 1 and bw == 1:
            m = sp.csr_matrix(m)
            m.sort_indices()
            return SparseMatrix(m.shape[0], m.shape[1], 1, 1, m.indptr, m.indices, m.data)
        m = sp.bsr_matrix(m, blocksize=(bh, bw))
        m.sort_indices()
        return SparseMatrix(m.shape[0] // bh, m.shape[1] // bw, bh, bw, m.indptr, m.indices, m.data)

    def to_scipy(self):
        import scipy.sparse as sp
        if self.bh == 1 and self.bw == 1:
            return sp.csr_matrix((self.val, self.col, self.rowptr), shape=(self.nrows, self.ncols))
        return sp.bsr_matrix((self.val.reshape(-1, self.bh, self.bw), self.col, self.rowptr),
                             shape=(self.nrows * self.bh, self.ncols * self.bw)).tocsr()

    def _abi(self):
        return _lib.Csr(self.nrows, self.ncols, self.bh, self.bw, self.rowptr.ctypes.data, self.col.ctypes.data,
                        self.val.ctypes.data)


def _flag_value(v):
    if isinstance(v, bool):
        return "1" if v else "0"
    if isinstance(v, (list, tuple)):
        return ",".join(_flag_value(x) for x in v)
    return str(v)


def _is_device(a):
    return hasattr(a, "is_cuda") and a.is_cuda


class Smoother:
    """BaseSmoother view of one level (python_smoothers.cpp:33-389): Smooth / SmoothBack with the reference flags."""

    def __init__(self, pc, level):
        self._pc, self.level = pc, level

    def Smooth(self, x, b, res=None, res_updated=False, update_res=True, x_zero=False):
        self._pc._smooth(self.level, x, b, res, res_updated, update_res, x_zero, False)

    def SmoothBack(self, x, b, res=None, res_updated=False, update_res=True, x_zero=False):
        self._pc._smooth(self.level, x, b, res, res_updated, update_res, x_zero, True)


class AMGMatrix:
    """AMGMatrix view (amg_matrix.hpp:14-87): the multigrid cycle as a BaseMatrix."""

    def __init__(self, pc):
        self._pc = pc

    def Mult(self, b, x):
        self._pc.Mult(b, x)

    def MultAdd(self, s, b, x):
        self._pc.MultAdd(s, b, x)

    MultTrans = Mult
    MultTransAdd = MultAdd

    def GetNLevels(self, rank=0):
        return self._pc.GetNLevels(rank)

    def GetNDof(self, level, rank=0):
        return self._pc.GetNDof(level, rank)

    def GetOC(self):
        return self._pc.GetOC()

    def GetBF(self, vec, level, dof, comp=0, rank=0, onLevel=0):
        return self._pc.GetBF(vec, level, dof, comp, rank, onLevel)

    def CINV(self, x, b):
        return self._pc.CINV(x, b)

    def GetSmoother(self, level=0):
        return self._pc.GetSmoother(level)


class Preconditioner:
    """BaseAMGPC (src/base/precond/amg_pc.hpp:26-228), strict-algebraic mode."""

    _type = None

    def __init__(self, mat, freedofs=None, vertex_xyz=None, prolongations=None, device=0, **kwargs):
        if not isinstance(mat, SparseMatrix):
            raise TypeError("mat must be an ngsamg_b200.SparseMatrix")
        L = _lib.lib()
        self._lib = L
        self._h = C.c_void_p()
        self.mat = mat
        self.flags = dict(kwargs)
        fm = None if freedofs is None else np.ascontiguousarray(freedofs, dtype=np.uint8)
        if fm is not None and fm.shape[0] != mat.nrows:
            raise ValueError("freedofs has the wrong length")
        xyz = None if vertex_xyz is None else np.ascontiguousarray(vertex_xyz, dtype=np.float64)
        keys = [k.encode() for k in kwargs]
        vals = [_flag_value(v).encode() for v in kwargs.values()]
        karr = (C.c_char_p * max(len(keys), 1))(*keys)
        varr = (C.c_char_p * max(len(vals), 1))(*vals)
        abi = mat._abi()
        _lib.check(L.ngsamg_b200_create(self._type.encode(), C.byref(abi), _lib.ptr(fm), _lib.ptr(xyz), karr, varr, len(keys),
                                        int(device), C.byref(self._h)))
        self._finalized = False
        self._prols = None
        if prolongations is not None:
            self.SetProlongations(prolongations)
        self.FinalizeLevel(mat)

    # -- BaseAMGPC protocol ----------------------------------------------------------------------
    def SetProlongations(self, prols):
        arr = (_lib.Csr * max(len(prols), 1))(*[p._abi() for p in prols])
        self._prols = list(prols)
        _lib.check(self._lib.ngsamg_b200_set_prolongations(self._h, len(prols), arr))

    def InitLevel(self, freedofs=None):  # freedofs are bound at construction in strict-algebraic mode
        return None

    def FinalizeLevel(self, mat=None):
        if not self._finalized:
            _lib.check(self._lib.ngsamg_b200_finalize(self._h))
            self._finalized = True

    def __del__(self):
        try:
            if self._h:
                self._lib.ngsamg_b200_destroy(self._h)
                self._h = None
        except Exception:
            pass

    # -- BaseMatrix quartet ----------------------------------------------------------------------
    def Mult(self, b, x):
        n = self.height
        _lib.check(self._lib.ngsamg_b200_apply(self._h, _lib.vec(b, n, "Mult: b"), _lib.vec(x, n, "Mult: x")))

    def MultAdd(self, s, b, x):
        n = self.height
        _lib.check(self._lib.ngsamg_b200_apply_add(self._h, float(s), _lib.vec(b, n, "MultAdd: b"), _lib.vec(x, n, "MultAdd: x")))

    PHASES = ("tri_sweeps", "parallel_halves", "transfers", "exchange", "g_spmv", "coarse", "other")

    def MultPhases(self, b, x):
        """one V-cycle run eagerly with CUDA events around every phase -> dict of device milliseconds per phase (measurement aid)"""
        n = self.height
        ms = np.zeros(8)
        _lib.check(self._lib.ngsamg_b200_apply_phases(self._h, _lib.vec(b, n, "MultPhases: b"), _lib.vec(x, n, "MultPhases: x"), _lib.ptr(ms)))
        return dict(zip(self.PHASES, [float(v) for v in ms[:7]]))

    MultTrans = Mult          # amg_matrix.cpp:381-382
    MultTransAdd = MultAdd    # amg_matrix.cpp:392-393

    def __mul__(self, b):
        x = np.zeros(self.height)
        self.Mult(np.ascontiguousarray(b, dtype=np.float64), x)
        return x

    @property
    def height(self):
        return self.mat.nrows * self.mat.bh

    width = height

    def CreateVector(self):
        return np.zeros(self.height)

    CreateRowVector = CreateColVector = CreateVector

    def IsComplex(self):
        return False

    # -- introspection ---------------------------------------------------------------------------
    def GetNLevels(self, rank=0):
        return int(self._lib.ngsamg_b200_num_levels(self._h))

    def level_info(self, level):
        info = _lib.LevelInfo()
        _lib.check(self._lib.ngsamg_b200_level_info(self._h, int(level), C.byref(info)))
        return info

    def GetNDof(self, level, rank=0):
        return int(self.level_info(level).n)

    def SweepKind(self, level=0):
        """which kernel sweeps the level: 'rows' | 'warp_tiles' | 'cta_tiles' | 'tile_images'"""
        return {0: "rows", 1: "warp_tiles", 2: "cta_tiles", 3: "tile_images", 4: "rows_rm"}[int(self._lib.ngsamg_b200_level_sweep_kind(self._h, int(level)))]

    def GetBlockSize(self, level=0):
        return int(self.level_info(level).b)

    def GetNProcs(self, level=0):
        """number of ranks holding `level` (python_amg.hpp:36-48): 1 for the single-rank preconditioner"""
        return 1

    @staticmethod
    def __flags_doc__():
        return {}

    def RegularizeMatrix(self, mat):
        """VertexAMGPC::RegularizeMatrix (python_amg.hpp:86-91; elasticity_pc_impl.hpp:711-763), local branch: regularises the diagonal blocks
        of `mat` in place (elast_3d: RegTM<0,6,6> on 6x6 blocks, elast_2d: unit rotational entry on 3x3 blocks; no-op for the H1 classes)"""
        dim = {"elast_3d": 3, "elast_2d": 2}.get(self._type, 0)
        if not dim or mat.bh != mat.bw or mat.bh != (6 if dim == 3 else 3):
            return mat
        bb = mat.bh * mat.bh
        val = mat.val.reshape(-1, bb)
        self._lib.ngsamg_b200_block_regularize.argtypes = [C.c_int, C.c_void_p, C.c_int]
        for i in range(mat.nrows):
            for k in range(mat.rowptr[i], mat.rowptr[i + 1]):
                if mat.col[k] == i:
                    blk = np.ascontiguousarray(val[k])
                    _lib.check(self._lib.ngsamg_b200_block_regularize(mat.bh, blk.ctypes.data_as(C.c_void_p), dim))
                    val[k] = blk
        mat.val = np.ascontiguousarray(val.reshape(-1))
        return mat

    def GetNDBS(self, level, rank=0):
        i = self.level_info(level)
        return int(i.n), int(i.b)

    def GetOC(self):
        """AMGMatrix::GetOC (amg_matrix.cpp:551-582): [OC, OC_l0, OC_l1, ...]"""
        n = self._lib.ngsamg_b200_operator_complexities(self._h, None, 0)
        occs = np.zeros(max(n, 1))
        self._lib.ngsamg_b200_operator_complexities(self._h, occs.ctypes.data_as(C.c_void_p), n)
        return [float(v) for v in occs[:n]]

    def GetAMGMatrix(self):
        return AMGMatrix(self)

    GetMatrix = GetAMGMatrix

    def GetAMatrix(self):
        return self.mat

    def GetSmoother(self, level=0):
        if level >= self.GetNLevels() - 1:
            raise NgsAMGError("only have %d smoothers" % (self.GetNLevels() - 1))
        return Smoother(self, level)

    def GetLevelMatrix(self, level):
        i = self.level_info(level)
        rp = np.zeros(i.n + 1, np.int64)
        ci = np.zeros(max(i.nnz, 1), np.int32)
        v = np.zeros(max(i.nnz, 1) * i.b * i.b, np.float64)
        _lib.check(self._lib.ngsamg_b200_get_level_matrix(self._h, int(level), _lib.ptr(rp), _lib.ptr(ci), _lib.ptr(v)))
        return SparseMatrix(i.n, i.n, i.b, i.b, rp, ci[:i.nnz], v[:i.nnz * i.b * i.b])

    def GetProlongation(self, level):
        """the DOF map step level+1 -> level (ProlMap::GetProl)"""
        i = self.level_info(level)
        rp = np.zeros(i.n + 1, np.int64)
        ci = np.zeros(max(i.nnz_prol, 1), np.int32)
        v = np.zeros(max(i.nnz_prol, 1) * i.b * i.bcoarse, np.float64)
        _lib.check(self._lib.ngsamg_b200_get_prolongation(self._h, int(level), _lib.ptr(rp), _lib.ptr(ci), _lib.ptr(v)))
        return SparseMatrix(i.n, i.ncoarse, i.b, i.bcoarse, rp, ci[:i.nnz_prol], v[:i.nnz_prol * i.b * i.bcoarse])

    def GetMap(self):
        return [self.GetProlongation(l) for l in range(self.GetNLevels() - 1)]

    def GetLevelVector(self, which, level):
        i = self.level_info(level)
        out = np.zeros(i.n * i.b)
        _lib.check(self._lib.ngsamg_b200_get_level_vector(self._h, int(level), {"x": 0, "rhs": 1, "res": 2}[which], _lib.ptr(out)))
        return out

    def GetSweepOrder(self, level=0):
        """rank of every row in the Gauss-Seidel sweep of `level` (identity = the reference's natural order)"""
        out = np.zeros(self.level_info(level).n, np.int32)
        _lib.check(self._lib.ngsamg_b200_get_sweep_order(self._h, int(level), _lib.ptr(out)))
        return out

    def VCycleBytes(self):
        return float(self._lib.ngsamg_b200_vcycle_bytes(self._h))

    def LastMs(self, what="apply"):
        return float(self._lib.ngsamg_b200_last_ms(self._h, {"apply": 0, "pcg": 1, "setup": 2, "rap": 3, "host": 4, "rap_bytes": 5}[what]))

    def LaunchCount(self):
        return int(self._lib.ngsamg_b200_launch_count(self._h))

    KERNELS = {"gs_tri_fwd": 0, "gs_upass": 1, "gs_lpass": 2, "gs_tri_bwd": 3, "spmv": 4, "restrict": 5, "prolong": 6, "gs_tri_fwd_rhs": 7,
               "gs_tri_bwd_res": 8, "halo_exchange": 9}

    def ProfileKernel(self, which, level=0, reps=10):
        """(avg ms per launch, algorithmic bytes per launch) of one V-cycle kernel, CUDA events on the library stream"""
        ms, by = C.c_double(), C.c_double()
        _lib.check(self._lib.ngsamg_b200_profile_kernel(self._h, int(level), self.KERNELS[which], int(reps), C.byref(ms), C.byref(by)))
        return ms.value, by.value

    def GetGSBlocks(self, level):
        """block (= coarse vertex) of every vertex of a level smoothed by block Gauss-Seidel (sm_type=bgs), -1 = in no block
        (GetGSBlocks, amg_pc_vertex_impl.hpp:1171-1269)"""
        out = np.zeros(self.GetNDof(level), np.int32)
        _lib.check(self._lib.ngsamg_b200_get_gs_blocks(self._h, int(level), _lib.ptr(out)))
        return out

    def SetTunable(self, name, value):
        """measurement aid: change a run-time tunable of the sweep kernels on the finalized hierarchy (include/ngsamg_b200.h)"""
        _lib.check(self._lib.ngsamg_b200_set_tunable(self._h, str(name).encode(), float(value)))

    # -- level operations ------------------------------------------------------------------------
    def _level_len(self, level):
        """scalar length of a vector of that level"""
        return self.GetNDof(level) * self.GetBlockSize(level)

    def _smooth(self, level, x, b, res, ru, ur, xz, back):
        n = self._level_len(level)
        _lib.check(self._lib.ngsamg_b200_smooth(self._h, int(level), _lib.vec(x, n, "smooth: x"), _lib.vec(b, n, "smooth: b"),
                                                _lib.vec(res, n, "smooth: res"), int(ru), int(ur),
                                                int(xz), int(back)))

    def LevelMultAdd(self, level, s, x, y):
        """y += s * A_level * x"""
        n = self._level_len(level)
        _lib.check(self._lib.ngsamg_b200_spmv_add(self._h, int(level), float(s), _lib.vec(x, n, "LevelMultAdd: x"), _lib.vec(y, n, "LevelMultAdd: y")))

    def TransferF2C(self, level, xf, xc):
        _lib.check(self._lib.ngsamg_b200_restrict(self._h, int(level), _lib.vec(xf, self._level_len(level), "TransferF2C: fine"),
                                                  _lib.vec(xc, self._level_len(level + 1), "TransferF2C: coarse")))

    def AddC2F(self, level, fac, xf, xc):
        _lib.check(self._lib.ngsamg_b200_prolong_add(self._h, int(level), float(fac), _lib.vec(xc, self._level_len(level + 1), "AddC2F: coarse"),
                                                     _lib.vec(xf, self._level_len(level), "AddC2F: fine")))

    def CoarseSolve(self, rhs, x):
        """crs_inv->Mult on the coarsest level (amg_matrix.cpp:228-233)"""
        n = self._level_len(self.GetNLevels() - 1)
        _lib.check(self._lib.ngsamg_b200_coarse_solve(self._h, _lib.vec(rhs, n, "CoarseSolve: rhs"), _lib.vec(x, n, "CoarseSolve: x")))
        return x

    def CINV(self, x, b):
        """AMGMatrix::CINV (amg_matrix.cpp:407-435): restrict b through all levels (TransferF2C), solve exactly on the coarsest level,
        prolongate the result back (TransferC2F) -- the coarse-grid correction without any smoothing"""
        nl = self.GetNLevels()
        r = np.ascontiguousarray(b, np.float64)
        for l in range(nl - 1):
            rc = np.zeros(self.GetNDof(l + 1) * self.GetBlockSize(l + 1))
            self.TransferF2C(l, r, rc)
            r = rc
        xc = np.zeros_like(r)
        self.CoarseSolve(r, xc)
        for l in range(nl - 2, -1, -1):
            xf = np.zeros(self.GetNDof(l) * self.GetBlockSize(l))
            self.AddC2F(l, 1.0, xf, xc)
            xc = xf
        x[:] = xc
        return x

    def GetBF(self, vec, level, dof, comp=0, rank=0, onLevel=0):
        """AMGMatrix::GetBF (amg_matrix.cpp:438-510; python_amg.hpp:30-101): the coarse basis function of (level, dof, comp) prolongated down to
        `onLevel` -- unit vector on `level`, then x_l = P_l x_{l+1} through the device transfers.  Single rank: rank must be 0."""
        if rank != 0:
            raise NgsAMGError("GetBF: rank %d does not exist (single-rank preconditioner)" % rank)
        nl = self.GetNLevels()
        if not (0 <= onLevel <= level < nl):
            raise NgsAMGError("GetBF: invalid level %d (levels 0..%d, onLevel %d)" % (level, nl - 1, onLevel))
        bs = self.GetBlockSize(level)
        if not (0 <= comp < bs and 0 <= dof < self.GetNDof(level)):
            raise NgsAMGError("GetBF: component %d / dof %d invalid on level %d" % (comp, dof, level))
        x = np.zeros(self.GetNDof(level) * bs)
        x[bs * dof + comp] = 1.0
        for l in range(level - 1, onLevel - 1, -1):
            xf = np.zeros(self.GetNDof(l) * self.GetBlockSize(l))
            self.AddC2F(l, 1.0, xf, x)
            x = xf
        vec[:] = x
        return vec

    def _pcg(self, rhs, x, tol, maxsteps):
        it = C.c_int(0)
        errs = np.zeros(maxsteps + 2)
        n = self.height
        _lib.check(self._lib.ngsamg_b200_pcg(self._h, _lib.vec(rhs, n, "pcg: rhs"), _lib.vec(x, n, "pcg: x"), float(tol), int(maxsteps), C.byref(it),
                                             _lib.ptr(errs)))
        return it.value, errs[: it.value + 1].copy()


def _make(name):
    return type(name, (Preconditioner,), {"_type": name, "__doc__": "NgsAMG.%s (registered in the reference by "
                                                                     "RegisterAMGSolver / RegisterPreconditioner)" % name})


h1_scal = _make("h1_scal")      # src/h1/h1_dim1.cpp:76
h1_2d = _make("h1_2d")          # src/h1/h1_dim2.cpp:44
h1_3d = _make("h1_3d")          # src/h1/h1_dim3.cpp:43
elast_2d = _make("elast_2d")    # src/elasticity/elasticity_2d.cpp:425
elast_3d = _make("elast_3d")    # src/elasticity/elasticity_3d.cpp:904

_REGISTRY = {"h1_scal": h1_scal, "h1_2d": h1_2d, "h1_3d": h1_3d, "elast_2d": elast_2d, "elast_3d": elast_3d}


def CreatePreconditioner(name, mat, freedofs=None, **kwargs):
    """ngsolve.Preconditioner(a, "NgsAMG.h1_scal", **flags) analogue (names as registered, amg_register.hpp:79-98)."""
    key = name
    for pre in ("NgsAMG.", "ngs_amg."):
        if key.startswith(pre):
            key = key[len(pre):]
    if key not in _REGISTRY:
        raise NgsAMGError("unknown preconditioner type '%s'" % name)
    return _REGISTRY[key](mat, freedofs, **kwargs)


class CGSolver:
    """ngsolve.krylovspace.CGSolver(mat, pre, maxsteps, tol) as used by tests/h1/amg_utils.py:346-362; runs on the device."""

    def __init__(self, mat, pre, maxsteps=100, tol=1e-12, callback=None):
        if pre.mat is not mat:
            raise ValueError("CGSolver: `mat` must be the matrix the preconditioner was built for")
        self.mat, self.pre, self.maxsteps, self.tol, self.callback = mat, pre, int(maxsteps), float(tol), callback
        self.iterations, self.errors = 0, []

    def Solve(self, rhs, sol=None):
        if sol is None:
            sol = np.zeros(self.pre.height)
        self.iterations, errs = self.pre._pcg(rhs, sol, self.tol, self.maxsteps)
        self.errors = list(errs)
        if self.callback:
            for k, e in enumerate(self.errors[1:]):
                self.callback(k, e)
        return sol


def _spm_fetch(L, h, nrows, ncols, bh, bw, nnz):
    rp = np.zeros(nrows + 1, np.int64)
    ci = np.zeros(max(nnz, 1), np.int32)
    v = np.zeros(max(nnz, 1) * bh * bw, np.float64)
    _lib.check(L.ngsamg_b200_spm_fetch(h, _lib.ptr(rp), _lib.ptr(ci), _lib.ptr(v)))
    return SparseMatrix(nrows, ncols, bh, bw, rp, ci[:nnz], v[:nnz * bh * bw])


def rap(A, P, device=0):
    """Galerkin product (P^T A) P on the device == RestrictMatrix (utils_sparseMM.hpp:93-109)."""
    L = _lib.lib()
    h, nr, nz = C.c_void_p(), C.c_int64(), C.c_int64()
    a, p = A._abi(), P._abi()
    _lib.check(L.ngsamg_b200_rap_begin(C.byref(a), C.byref(p), device, C.byref(h), C.byref(nr), C.byref(nz)))
    return _spm_fetch(L, h, nr.value, P.ncols, P.bw, P.bw, nz.value)


def matmul(A, B, device=0):
    """MatMultABImpl (utils_sparseMM.cpp:107-238) on the device."""
    L = _lib.lib()
    h, nr, nz = C.c_void_p(), C.c_int64(), C.c_int64()
    a, b = A._abi(), B._abi()
    _lib.check(L.ngsamg_b200_matmul_begin(C.byref(a), C.byref(b), device, C.byref(h), C.byref(nr), C.byref(nz)))
    return _spm_fetch(L, h, nr.value, B.ncols, A.bh, B.bw, nz.value)


def transpose(A, device=0):
    """TransposeSPMImpl (utils_sparseMM.cpp:54-93)."""
    L = _lib.lib()
    h, nr, nz = C.c_void_p(), C.c_int64(), C.c_int64()
    a = A._abi()
    _lib.check(L.ngsamg_b200_transpose_begin(C.byref(a), device, C.byref(h), C.byref(nr), C.byref(nz)))
    return _spm_fetch(L, h, nr.value, A.nrows, A.bw, A.bh, nz.value)


def coarsen(A, freedofs=None, vertex_xyz=None, bcoarse=None, max_per_row=3, min_frac=0.08, omega=1.0, smooth=True, rounds=3):
    """Host-side DOF-map construction (what finalize() runs per level): returns (P, vmap, coarse_xyz)."""
    L = _lib.lib()
    bc = A.bh if bcoarse is None else int(bcoarse)
    fm = None if freedofs is None else np.ascontiguousarray(freedofs, dtype=np.uint8)
    xyz = None if vertex_xyz is None else np.ascontiguousarray(vertex_xyz, dtype=np.float64)
    h, nc, nz = C.c_void_p(), C.c_int64(), C.c_int64()
    a = A._abi()
    _lib.check(L.ngsamg_b200_coarsen_begin(C.byref(a), _lib.ptr(fm), _lib.ptr(xyz), bc, int(max_per_row), float(min_frac),
                                           float(omega), int(bool(smooth)), int(rounds), C.byref(h), C.byref(nc), C.byref(nz)))
    rp = np.zeros(A.nrows + 1, np.int64)
    ci = np.zeros(max(nz.value, 1), np.int32)
    v = np.zeros(max(nz.value, 1) * A.bh * bc, np.float64)
    vmap = np.zeros(A.nrows, np.int32)
    cxyz = None if xyz is None else np.zeros((nc.value, 3))
    _lib.check(L.ngsamg_b200_coarsen_fetch(h, _lib.ptr(rp), _lib.ptr(ci), _lib.ptr(v), _lib.ptr(vmap), _lib.ptr(cxyz)))
    P = SparseMatrix(A.nrows, nc.value, A.bh, bc, rp, ci[:nz.value], v[:nz.value * A.bh * bc])
    return P, vmap, cxyz
