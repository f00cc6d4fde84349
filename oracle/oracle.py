"""ctypes front-end of the CPU oracle (oracle/ngsamg_oracle.c).

TEST INFRASTRUCTURE ONLY -- imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs.  The product package (ngsamg_b200/) never imports this module.

PARITY PINNED IN PART (see the header of ngsamg_oracle.c): the reference as a whole cannot be built or run in
this image and its tests carry no golden vectors, but the bodies of its functions for the single-rank path are
compiled verbatim against a stand-in for the NGSolve containers (oracle/ref_pin/ -> oracle/_ref/) and this oracle
agrees with them bit for bit (tests/test_ref_pin.py, tests/golden/refpin_*.npz).  The NGSolve-internal arithmetic,
the pseudo-inverse, Jacobi, CG and the multi-rank path stay cross-checked only against scipy / pure-python loops /
the assembled operator (tests/test_oracle.py, tests/test_parallel_host.py).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libngsamg_oracle.so")

i64p = np.ctypeslib.ndpointer(np.int64, flags="C_CONTIGUOUS")
i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")
u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")

SM_GS, SM_JACOBI = 0, 1


def build(force=False):
    src = os.path.join(_HERE, "ngsamg_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B"], stdout=subprocess.DEVNULL)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_SO):
        build()
    L = C.CDLL(_SO)
    vp, i64, ci, dbl = C.c_void_p, C.c_int64, C.c_int, C.c_double
    L.orc_spmv_add.argtypes = [i64, ci, ci, i64p, i32p, f64p, dbl, f64p, f64p]
    L.orc_transpose.argtypes = [i64, i64, ci, ci, i64p, i32p, f64p, i64p, i32p, f64p]
    L.orc_matmul_count.argtypes = [i64, i64p, i32p, i64p, i32p, i64p]
    L.orc_matmul_count.restype = i64
    L.orc_matmul_fill.argtypes = [i64, ci, ci, ci, i64p, i32p, f64p, i64p, i32p, f64p, i64p, i32p, f64p]
    L.orc_calc_dinv.argtypes = [i64, ci, i64p, i32p, f64p, vp, ci, vp, f64p]
    L.orc_calc_dinv.restype = ci
    L.orc_gs_rhs.argtypes = [i64, ci, i64p, i32p, f64p, f64p, vp, f64p, f64p, ci]
    L.orc_gs_res.argtypes = [i64, ci, i64p, i32p, f64p, f64p, vp, f64p, f64p, ci]
    L.orc_amg_new.argtypes = [ci]
    L.orc_amg_new.restype = vp
    L.orc_amg_set_matrix.argtypes = [vp, ci, i64, ci, i64p, i32p, f64p, vp]
    L.orc_amg_set_smoother.argtypes = [vp, ci, ci, ci, ci, ci, dbl]
    L.orc_amg_set_smoother.restype = ci
    L.orc_amg_set_prol.argtypes = [vp, ci, i64, ci, i64p, i32p, f64p]
    L.orc_amg_galerkin.argtypes = [vp, ci]
    L.orc_amg_regularize_coarse.argtypes = [vp, ci]
    L.orc_regularize_block.argtypes = [ci, f64p, ci]
    L.orc_amg_set_coarse_inv.argtypes = [vp]
    L.orc_amg_set_coarse_inv.restype = ci
    L.orc_amg_smooth.argtypes = [vp, ci, f64p, f64p, f64p, ci, ci, ci, ci]
    L.orc_amg_apply.argtypes = [vp, f64p, f64p]
    L.orc_amg_apply_w.argtypes = [vp, f64p, f64p]
    L.orc_amg_apply_bs.argtypes = [vp, f64p, f64p]
    L.orc_amg_apply_add.argtypes = [vp, dbl, f64p, f64p]
    L.orc_amg_pcg.argtypes = [vp, f64p, f64p, dbl, ci, f64p]
    L.orc_amg_pcg.restype = ci
    L.orc_amg_free.argtypes = [vp]
    for nm, rt in [("n", i64), ("b", ci), ("nnz", i64)]:
        f = getattr(L, "orc_amg_level_" + nm)
        f.argtypes = [vp, ci]
        f.restype = rt
    for nm in ["level_rowptr", "level_col", "level_val", "level_dinv", "level_x", "level_rhs", "level_res",
               "pt_rowptr", "pt_col", "pt_val"]:
        f = getattr(L, "orc_amg_" + nm)
        f.argtypes = [vp, ci]
        f.restype = vp
    _lib = L
    return L


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Bsr:
    """Block CSR in NGSolve SparseMatrix<Mat<bh,bw>> layout."""

    def __init__(self, nrows, ncols, bh, bw, rowptr, col, val):
        self.nrows, self.ncols, self.bh, self.bw = int(nrows), int(ncols), int(bh), int(bw)
        self.rowptr = np.ascontiguousarray(rowptr, dtype=np.int64)
        self.col = np.ascontiguousarray(col, dtype=np.int32)
        self.val = np.ascontiguousarray(val, dtype=np.float64).reshape(-1)
        assert self.rowptr.shape[0] == self.nrows + 1
        assert self.val.shape[0] == self.col.shape[0] * self.bh * self.bw

    @property
    def nnz(self):
        return int(self.rowptr[-1])

    @staticmethod
    def from_scipy(m, bh=1, bw=1):
        import scipy.sparse as sp
        if bh == 1 and bw == 1:
            m = sp.csr_matrix(m)
            m.sort_indices()
            return Bsr(m.shape[0], m.shape[1], 1, 1, m.indptr, m.indices, m.data)
        m = sp.bsr_matrix(m, blocksize=(bh, bw))
        m.sort_indices()
        return Bsr(m.shape[0] // bh, m.shape[1] // bw, bh, bw, m.indptr, m.indices, m.data)

    def to_scipy(self):
        import scipy.sparse as sp
        if self.bh == 1 and self.bw == 1:
            return sp.csr_matrix((self.val, self.col, self.rowptr), shape=(self.nrows, self.ncols))
        return sp.bsr_matrix((self.val.reshape(-1, self.bh, self.bw), self.col, self.rowptr),
                             shape=(self.nrows * self.bh, self.ncols * self.bw)).tocsr()


def spmv_add(A, s, x, y):
    lib().orc_spmv_add(A.nrows, A.bh, A.bw, A.rowptr, A.col, A.val, s, x, y)
    return y


def transpose(A):
    trp = np.zeros(A.ncols + 1, np.int64)
    tci = np.zeros(max(A.nnz, 1), np.int32)
    tv = np.zeros(max(A.nnz, 1) * A.bh * A.bw, np.float64)
    lib().orc_transpose(A.nrows, A.ncols, A.bh, A.bw, A.rowptr, A.col, A.val, trp, tci, tv)
    return Bsr(A.ncols, A.nrows, A.bw, A.bh, trp, tci[:A.nnz], tv[:A.nnz * A.bh * A.bw])


def matmul(A, B):
    assert A.ncols == B.nrows and A.bw == B.bh
    crp = np.zeros(A.nrows + 1, np.int64)
    nnz = lib().orc_matmul_count(A.nrows, A.rowptr, A.col, B.rowptr, B.col, crp)
    cci = np.zeros(max(nnz, 1), np.int32)
    cv = np.zeros(max(nnz, 1) * A.bh * B.bw, np.float64)
    lib().orc_matmul_fill(A.nrows, A.bh, A.bw, B.bw, A.rowptr, A.col, A.val, B.rowptr, B.col, B.val, crp, cci, cv)
    return Bsr(A.nrows, B.ncols, A.bh, B.bw, crp, cci[:nnz], cv[:nnz * A.bh * B.bw])


def restrict_matrix(PT, A, P):
    """RestrictMatrix (utils_sparseMM.hpp:93-109): (PT*A)*P."""
    return matmul(matmul(PT, A), P)


def calc_dinv(A, free=None, pinv=False, repl_diag=None):
    d = np.zeros(A.nrows * A.bh * A.bh, np.float64)
    fm = None if free is None else np.ascontiguousarray(free, np.uint8)
    rd = None if repl_diag is None else np.ascontiguousarray(repl_diag, np.float64)
    lib().orc_calc_dinv(A.nrows, A.bh, A.rowptr, A.col, A.val, _ptr(fm), int(pinv), _ptr(rd), d)
    return d


def regularize_block(m, dim):
    """RegularizeMatrix on one diagonal block (elasticity_pc_impl.hpp:711-763): dim 3 -> RegTM<0,6,6>, dim 2 -> unit rotational entry"""
    a = np.ascontiguousarray(m, np.float64).copy()
    lib().orc_regularize_block(a.shape[0], a.reshape(-1), int(dim))
    return a


def gs_rhs(A, dinv, free, x, rhs, backwards):
    fm = None if free is None else np.ascontiguousarray(free, np.uint8)
    lib().orc_gs_rhs(A.nrows, A.bh, A.rowptr, A.col, A.val, dinv, _ptr(fm), x, rhs, int(backwards))


def gs_res(A, dinv, free, x, res, backwards):
    fm = None if free is None else np.ascontiguousarray(free, np.uint8)
    lib().orc_gs_res(A.nrows, A.bh, A.rowptr, A.col, A.val, dinv, _ptr(fm), x, res, int(backwards))


def _as_np(ptr, n, dt):
    if n == 0:
        return np.zeros(0, dt)
    ct = {np.int64: C.c_int64, np.int32: C.c_int32, np.float64: C.c_double}[dt]
    return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(ct)), shape=(n,)).copy()


class OracleAMG:
    """AMGMatrix restatement: level matrices by Galerkin RAP from injected prolongations,
    one smoother per level (the last level only has the exact coarse solve)."""

    def __init__(self, A, free, prols, sm_type="gs", sm_steps=1, sm_symm=False, pinv=False, omega=None,
                 clev="inv", regularize=None):
        """pinv models ngs_amg_regularize_cmats for the smoothers; regularize (default: pinv on elasticity block sizes) is the other
        half of that flag, RegularizeMatrix on the coarsest diagonal blocks before the inverse (amg_pc.cpp:861-862)"""
        L = lib()
        self.nlevels = len(prols) + 1
        self.h = L.orc_amg_new(self.nlevels)
        fm = None if free is None else np.ascontiguousarray(free, np.uint8)
        self._keep = [A, fm, prols]
        L.orc_amg_set_matrix(self.h, 0, A.nrows, A.bh, A.rowptr, A.col, A.val, _ptr(fm))
        for l, P in enumerate(prols):
            assert P.nrows == L.orc_amg_level_n(self.h, l), (P.nrows, L.orc_amg_level_n(self.h, l))
            L.orc_amg_set_prol(self.h, l, P.ncols, P.bw, P.rowptr, P.col, P.val)
            L.orc_amg_galerkin(self.h, l)
        smt = {"gs": SM_GS, "jacobi": SM_JACOBI}[sm_type]
        if omega is None:
            omega = 0.9 if smt == SM_JACOBI else 1.0
        for l in range(self.nlevels - 1):
            rc = L.orc_amg_set_smoother(self.h, l, smt, int(sm_steps), int(bool(sm_symm)), int(bool(pinv)), float(omega))
            if rc:
                raise RuntimeError("oracle: singular diagonal block on level %d (rc=%d)" % (l, rc))
        self.has_cinv = False
        cb = prols[-1].bw if len(prols) else A.bh
        if regularize is None:
            regularize = bool(pinv) and (cb == 6 or (cb == 3 and A.bh == 2))   # elast_3d: 6x6 coarse blocks; elast_2d: 2 -> 3
        if regularize and clev == "inv":
            L.orc_amg_regularize_coarse(self.h, 3 if cb == 6 else 2)     # 6x6 coarse blocks: 3D, 3x3: 2D
        if clev == "inv":
            rc = L.orc_amg_set_coarse_inv(self.h)
            if rc:
                raise RuntimeError("oracle: coarsest matrix not positive definite")
            self.has_cinv = True
        self.n0 = A.nrows * A.bh

    def __del__(self):
        try:
            lib().orc_amg_free(self.h)
        except Exception:
            pass

    def level_matrix(self, l):
        L = lib()
        n, b, nnz = L.orc_amg_level_n(self.h, l), L.orc_amg_level_b(self.h, l), L.orc_amg_level_nnz(self.h, l)
        return Bsr(n, n, b, b, _as_np(L.orc_amg_level_rowptr(self.h, l), n + 1, np.int64),
                   _as_np(L.orc_amg_level_col(self.h, l), nnz, np.int32),
                   _as_np(L.orc_amg_level_val(self.h, l), nnz * b * b, np.float64))

    def level_vec(self, which, l):
        L = lib()
        n, b = L.orc_amg_level_n(self.h, l), L.orc_amg_level_b(self.h, l)
        return _as_np(getattr(L, "orc_amg_level_" + which)(self.h, l), n * b, np.float64)

    def level_dinv(self, l):
        L = lib()
        n, b = L.orc_amg_level_n(self.h, l), L.orc_amg_level_b(self.h, l)
        return _as_np(L.orc_amg_level_dinv(self.h, l), n * b * b, np.float64)

    def smooth(self, l, x, b, res, res_updated=False, update_res=True, x_zero=False, backwards=False):
        lib().orc_amg_smooth(self.h, l, x, b, res, int(res_updated), int(update_res), int(x_zero), int(backwards))

    def apply(self, b, cycle="V"):
        """AMGMatrix::Mult -> SmoothV / SmoothW / SmoothBS (amg_matrix.cpp:37-307)"""
        x = np.zeros(self.n0)
        fn = {"V": lib().orc_amg_apply, "W": lib().orc_amg_apply_w, "BS": lib().orc_amg_apply_bs}[cycle]
        fn(self.h, np.ascontiguousarray(b, np.float64), x)
        return x

    def apply_add(self, s, b, x):
        """AMGMatrix::MultAdd"""
        lib().orc_amg_apply_add(self.h, float(s), np.ascontiguousarray(b, np.float64), x)
        return x

    def pcg(self, rhs, tol=1e-8, maxsteps=200):
        u = np.zeros(self.n0)
        errs = np.zeros(maxsteps + 2)
        it = lib().orc_amg_pcg(self.h, np.ascontiguousarray(rhs, np.float64), u, float(tol), int(maxsteps), errs)
        return u, it, errs[: it + 1].copy()
