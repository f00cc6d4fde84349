"""ctypes front-end of oracle/_ref/libngsamg_ref.so: the reference's OWN hot-path functions (cut from /root/reference at
build time, compiled verbatim against a stand-in for the NGSolve containers, see ref_harness.cpp / README.md).

TEST INFRASTRUCTURE ONLY -- used by tests/test_ref_pin.py and tests/golden/make_ref_golden.py to pin the oracle.
The library can only be BUILT where /root/reference exists; the built file travels with the repository snapshot.
"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

from oracle.oracle import Bsr

_HERE = os.path.dirname(os.path.abspath(__file__))
_ORACLE = os.path.dirname(_HERE)
_SO = os.path.join(_ORACLE, "_ref", "libngsamg_ref.so")
REFERENCE = os.environ.get("NGSAMG_REFERENCE", "/root/reference")

i64p = np.ctypeslib.ndpointer(np.int64, flags="C_CONTIGUOUS")
i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")


_avail = None


def available():
    """True if the library can be used: it loads (building it first where the reference tree is present).  Never raises -- tests skip,
    bench.py falls back to the oracle port, smoke() prints a note."""
    global _avail
    if _avail is None:
        try:
            lib()
            _avail = True
        except Exception as e:          # no reference tree and no prebuilt library, no compiler, or a library that does not load here
            sys.stderr.write("[oracle/ref_pin] reference library unavailable: %s\n" % e)
            _avail = False
    return _avail


def build():
    if os.path.isdir(os.path.join(REFERENCE, "src", "base")):
        subprocess.check_call(["make", "-C", _ORACLE, "REFERENCE=" + REFERENCE, "_ref/libngsamg_ref.so"], stdout=subprocess.DEVNULL)
    if not os.path.exists(_SO):
        raise RuntimeError("oracle/_ref/libngsamg_ref.so is missing and %s is not present to build it from" % REFERENCE)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    build()
    L = C.CDLL(_SO)
    vp, i64, ci = C.c_void_p, C.c_int64, C.c_int
    L.ref_last_error.restype = C.c_char_p
    L.ref_fragment_index.restype = C.c_char_p
    L.ref_mat_new.argtypes = [i64, i64, ci, ci, i64p, i32p, f64p]
    L.ref_mat_new.restype = vp
    L.ref_mat_free.argtypes = [vp]
    L.ref_mat_info.argtypes = [vp] + [C.POINTER(i64)] * 2 + [C.POINTER(ci)] * 2 + [C.POINTER(i64)]
    L.ref_mat_fetch.argtypes = [vp, i64p, i32p, f64p]
    L.ref_mat_transpose.argtypes = [vp]
    L.ref_mat_transpose.restype = vp
    L.ref_mat_mult.argtypes = [vp, vp]
    L.ref_mat_mult.restype = vp
    L.ref_mat_restrict.argtypes = [vp, vp, vp]
    L.ref_mat_restrict.restype = vp
    L.ref_amg_new.argtypes = [ci]
    L.ref_amg_new.restype = vp
    L.ref_amg_free.argtypes = [vp]
    L.ref_amg_set_matrix.argtypes = [vp, i64, ci, i64p, i32p, f64p, vp]
    L.ref_amg_set_prol.argtypes = [vp, ci, i64, ci, i64p, i32p, f64p]
    L.ref_amg_finalize.argtypes = [vp, ci, ci, vp]
    L.ref_amg_level_matrix.argtypes = [vp, ci]
    L.ref_amg_level_matrix.restype = vp
    L.ref_amg_level_pt.argtypes = [vp, ci]
    L.ref_amg_level_pt.restype = vp
    L.ref_amg_level_dinv.argtypes = [vp, ci, f64p]
    L.ref_amg_level_vec.argtypes = [vp, ci, ci, f64p]
    L.ref_amg_smooth.argtypes = [vp, ci, f64p, f64p, f64p, ci, ci, ci, ci, ci]
    L.ref_amg_apply.argtypes = [vp, ci, f64p, f64p]
    L.ref_amg_pcg.argtypes = [vp, ci, f64p, f64p, C.c_double, ci, f64p, C.POINTER(ci)]
    _lib = L
    return L


def fragment_index():
    """the reference file:line ranges the library was built from"""
    return lib().ref_fragment_index().decode()


def _check(rc):
    if rc:
        raise RuntimeError("reference: " + lib().ref_last_error().decode())


def _ptr(h):
    if not h:
        raise RuntimeError("reference: " + lib().ref_last_error().decode())
    return h


class RefMat:
    """a SparseMatrix<Mat<bh,bw>> living in the reference library"""

    def __init__(self, handle, owned=True):
        self.h, self.owned = _ptr(handle), owned

    @staticmethod
    def from_bsr(M):
        return RefMat(lib().ref_mat_new(M.nrows, M.ncols, M.bh, M.bw, M.rowptr, M.col if M.nnz else np.zeros(1, np.int32),
                                        M.val if M.nnz else np.zeros(1, np.float64)))

    def __del__(self):
        try:
            if self.owned:
                lib().ref_mat_free(self.h)
        except Exception:
            pass

    def to_bsr(self):
        nr, nc, nnz, bh, bw = C.c_int64(), C.c_int64(), C.c_int64(), C.c_int(), C.c_int()
        _check(lib().ref_mat_info(self.h, C.byref(nr), C.byref(nc), C.byref(bh), C.byref(bw), C.byref(nnz)))
        rp = np.zeros(nr.value + 1, np.int64)
        ci = np.zeros(max(nnz.value, 1), np.int32)
        v = np.zeros(max(nnz.value, 1) * bh.value * bw.value, np.float64)
        _check(lib().ref_mat_fetch(self.h, rp, ci, v))
        return Bsr(nr.value, nc.value, bh.value, bw.value, rp, ci[:nnz.value], v[:nnz.value * bh.value * bw.value])


def transpose(A):
    """TransposeSPMImpl (utils_sparseMM.cpp:54-93)"""
    a = RefMat.from_bsr(A)
    return RefMat(lib().ref_mat_transpose(a.h)).to_bsr()


def matmul(A, B):
    """MatMultABImpl (utils_sparseMM.cpp:107-238)"""
    a, b = RefMat.from_bsr(A), RefMat.from_bsr(B)
    return RefMat(lib().ref_mat_mult(a.h, b.h)).to_bsr()


def restrict_matrix(PT, A, P):
    """RestrictMatrix (utils_sparseMM.hpp:93-109)"""
    pt, a, p = RefMat.from_bsr(PT), RefMat.from_bsr(A), RefMat.from_bsr(P)
    return RefMat(lib().ref_mat_restrict(pt.h, a.h, p.h)).to_bsr()


def pinv_block(m):
    """CalcPseudoInverseTryNormal(Mat<N,N>&) (utils_denseLA.hpp:1549-1562) on one block"""
    L = lib()
    L.ref_pinv.argtypes = [C.c_int, f64p]
    a = np.ascontiguousarray(m, np.float64).copy()
    _check(L.ref_pinv(a.shape[0], a.reshape(-1)))
    return a


def elast_calcq(t, si=1.0, sj=1.0):
    """EpsEpsEnergy<DIM>::CalcQ (elasticity_energy_impl.hpp:8-29): the rigid-body transport block for the offset t (len 2 or 3)"""
    L = lib()
    L.ref_elast_calcq.argtypes = [C.c_int, f64p, C.c_double, C.c_double, f64p]
    t = np.ascontiguousarray(t, np.float64)
    n = 6 if len(t) == 3 else 3
    q = np.zeros(n * n)
    _check(L.ref_elast_calcq(len(t), t, float(si), float(sj), q))
    return q.reshape(n, n)


def regularize_block6(m):
    """RegTM<0,6,6> (utils_denseLA.hpp:1198-1234) on one block"""
    L = lib()
    L.ref_regularize6.argtypes = [f64p]
    a = np.ascontiguousarray(m, np.float64).copy()
    _check(L.ref_regularize6(a.reshape(-1)))
    return a


class RefAMG:
    """AMGMatrix of the reference: coarse matrices by TransposeSPMImpl + RestrictMatrix from injected prolongations, GSS3 per
    level (ProxySmoother around it for sm_steps > 1 / sm_symm), cycles by AMGMatrix::SmoothV / SmoothW / SmoothBS.  The exact
    coarse solve is a dense inverse handed in by the caller (the reference calls NGSolve's sparse Cholesky here)."""

    def __init__(self, A, free, prols=None, sm_steps=1, sm_symm=False, coarse_inv=True, max_levels=32, pinv=False):
        """with prols: the whole hierarchy at once.  prols=None: level by level -- add_prol(P) returns the Galerkin matrix the
        reference's RestrictMatrix produced (input of the next coarsening step), finalize() ends the set-up."""
        L = lib()
        self.nlevels = 1
        self.h = _ptr(L.ref_amg_new(max_levels if prols is None else len(prols) + 1))
        fm = None if free is None else np.ascontiguousarray(free, np.uint8)
        _check(L.ref_amg_set_matrix(self.h, A.nrows, A.bh, A.rowptr, A.col, A.val, None if fm is None else fm.ctypes.data_as(C.c_void_p)))
        self.n0 = A.nrows * A.bh
        L.ref_amg_set_pinv.argtypes = [C.c_void_p, C.c_int]
        L.ref_amg_set_pinv(self.h, int(bool(pinv)))
        self._opts = (int(sm_steps), int(bool(sm_symm)), bool(coarse_inv))
        if prols is not None:
            for P in prols:
                self.add_prol(P, fetch=False)
            self.finalize()

    def use_jacobi(self, omega=0.9, sm_steps=1, sm_symm=False):
        """replace the Gauss-Seidel smoothers by the reference's JacobiSmoother (RichardsonSmoother with prec = diag^-1)"""
        L = lib()
        L.ref_amg_use_jacobi.argtypes = [C.c_void_p, C.c_double, C.c_int, C.c_int]
        _check(L.ref_amg_use_jacobi(self.h, float(omega), int(sm_steps), int(bool(sm_symm))))

    def add_prol(self, P, fetch=True):
        """prolongation of the current coarsest level: TransposeSPMImpl + RestrictMatrix build the next level matrix"""
        _check(lib().ref_amg_set_prol(self.h, self.nlevels - 1, P.ncols, P.bw, P.rowptr, P.col, P.val))
        self.nlevels += 1
        return self.level_matrix(self.nlevels - 1) if fetch else None

    def finalize(self):
        sm_steps, sm_symm, coarse_inv = self._opts
        cinv = None
        if coarse_inv:
            Ac = self.level_matrix(self.nlevels - 1).to_scipy().toarray()
            cinv = np.ascontiguousarray(np.linalg.inv(Ac))
        _check(lib().ref_amg_finalize(self.h, sm_steps, sm_symm, None if cinv is None else cinv.ctypes.data_as(C.c_void_p)))

    def __del__(self):
        try:
            lib().ref_amg_free(self.h)
        except Exception:
            pass

    def level_matrix(self, l):
        return RefMat(lib().ref_amg_level_matrix(self.h, l), owned=False).to_bsr()

    def level_pt(self, l):
        return RefMat(lib().ref_amg_level_pt(self.h, l), owned=False).to_bsr()

    def level_dinv(self, l):
        A = self.level_matrix(l)
        out = np.zeros(A.nrows * A.bh * A.bh)
        _check(lib().ref_amg_level_dinv(self.h, l, out))
        return out

    def level_vec(self, which, l):
        A = self.level_matrix(l)
        out = np.zeros(A.nrows * A.bh)
        _check(lib().ref_amg_level_vec(self.h, {"x": 0, "rhs": 1, "res": 2}[which], l, out))
        return out

    def smooth(self, l, x, b, res, res_updated=False, update_res=True, x_zero=False, backwards=False, bare=False):
        _check(lib().ref_amg_smooth(self.h, l, x, np.ascontiguousarray(b, np.float64), res, int(res_updated), int(update_res),
                                    int(x_zero), int(backwards), int(bare)))

    def apply(self, b, cycle="V"):
        x = np.zeros(self.n0)
        _check(lib().ref_amg_apply(self.h, {"V": 0, "W": 1, "BS": 2}[cycle], np.ascontiguousarray(b, np.float64), x))
        return x

    def get_oc(self, cycle="V"):
        """AMGMatrix::GetOC (amg_matrix.cpp:551-582): [OC, OC_l0, OC_l1, ...]"""
        L = lib()
        L.ref_amg_get_oc.argtypes = [C.c_void_p, C.c_int, f64p, C.c_int]
        out = np.zeros(64)
        n = L.ref_amg_get_oc(self.h, {"V": 0, "W": 1, "BS": 2}[cycle], out, 64)
        if n < 0:
            _check(1)
        return [float(v) for v in out[:n]]

    def mult(self, b, x, cycle="V", trans=False):
        """AMGMatrix::Mult / MultTrans: x = C b (x is overwritten)"""
        L = lib()
        L.ref_amg_mult.argtypes = [C.c_void_p, C.c_int, C.c_int, f64p, f64p]
        _check(L.ref_amg_mult(self.h, {"V": 0, "W": 1, "BS": 2}[cycle], int(trans), np.ascontiguousarray(b, np.float64), x))
        return x

    def mult_add(self, s, b, x, cycle="V", trans=False):
        """AMGMatrix::MultAdd / MultTransAdd: x += s * C b"""
        L = lib()
        L.ref_amg_mult_add.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double, f64p, f64p]
        _check(L.ref_amg_mult_add(self.h, {"V": 0, "W": 1, "BS": 2}[cycle], int(trans), float(s), np.ascontiguousarray(b, np.float64), x))
        return x

    def pcg(self, rhs, tol=1e-8, maxsteps=200, cycle="V"):
        """CG (harness glue, NGSolve's CGSolver restated like the oracle's) preconditioned with the reference's cycle"""
        u, errs, it = np.zeros(self.n0), np.zeros(maxsteps + 2), C.c_int(0)
        _check(lib().ref_amg_pcg(self.h, {"V": 0, "W": 1, "BS": 2}[cycle], np.ascontiguousarray(rhs, np.float64), u, float(tol),
                                 int(maxsteps), errs, C.byref(it)))
        return u, it.value, errs[: it.value + 1].copy()


class RefHybridLevel:
    """one distributed level in the reference library: R ranks = R host threads around the reference's BasicDCCMap, HybridMatrix
    (DecomposeSparseMatrixHybrid), HybridGSSmoother (CalcModDiag, GSS3 + GSS4, stage protocol of HybridBaseSmoother).
    Same inputs as oracle_par.HybridLevel: per rank the sub-assembled matrix, free mask, ascending peers and exchange dofs."""

    def __init__(self, A_loc, free, peers, ex, overlap=True, symm_loc=False, nsteps_loc=1):
        L = lib()
        vp, ci, i64 = C.c_void_p, C.c_int, C.c_int64
        L.ref_par_new.argtypes, L.ref_par_new.restype = [ci], vp
        L.ref_par_free.argtypes = [vp]
        L.ref_par_set_rank.argtypes = [vp, ci, i64, ci, i64p, i32p, f64p, vp, ci, i32p, i64p, i32p]
        L.ref_par_setup.argtypes = [vp, ci, ci, ci]
        L.ref_par_M.argtypes, L.ref_par_M.restype = [vp, ci], vp
        L.ref_par_G.argtypes, L.ref_par_G.restype = [vp, ci], vp
        L.ref_par_info.argtypes = [vp, ci, C.POINTER(i64), vp, f64p]
        L.ref_par_dcc_lists.argtypes = [vp, ci, ci, C.POINTER(i64), vp, C.POINTER(i64), vp]
        pp = C.POINTER(C.c_void_p)
        L.ref_par_smooth.argtypes = [vp, pp, pp, pp, ci, ci, ci, ci, ci, ci]
        L.ref_par_mult.argtypes = [vp, pp, pp]
        L.ref_par_exchange.argtypes = [vp, ci, pp]
        self.R = len(A_loc)
        self.n = [a.nrows for a in A_loc]
        self.b = A_loc[0].bh
        self.peers = peers
        self.h = _ptr(L.ref_par_new(self.R))
        for r, A in enumerate(A_loc):
            fm = None if free[r] is None else np.ascontiguousarray(free[r], np.uint8)
            pr = np.ascontiguousarray(peers[r], np.int32)
            ptr = np.zeros(len(pr) + 1, np.int64)
            for k in range(len(pr)):
                ptr[k + 1] = ptr[k] + len(ex[r][k])
            exd = np.ascontiguousarray(np.concatenate([np.asarray(e, np.int32) for e in ex[r]] + [np.zeros(0, np.int32)]), np.int32)
            _check(L.ref_par_set_rank(self.h, r, A.nrows, A.bh, A.rowptr, A.col, A.val, None if fm is None else fm.ctypes.data_as(vp),
                                      len(pr), pr if len(pr) else np.zeros(1, np.int32), ptr, exd if len(exd) else np.zeros(1, np.int32)))
        _check(L.ref_par_setup(self.h, int(overlap), int(symm_loc), int(nsteps_loc)))

    def __del__(self):
        try:
            lib().ref_par_free(self.h)
        except Exception:
            pass

    def M(self, r):
        return RefMat(lib().ref_par_M(self.h, r), owned=False).to_bsr()

    def G(self, r):
        g = lib().ref_par_G(self.h, r)
        return None if not g else RefMat(g, owned=False).to_bsr()

    def info(self, r):
        """(split_ind, master flags, inverted modified diagonal blocks held by the local smoothers)"""
        split = C.c_int64()
        master = np.zeros(self.n[r], np.uint8)
        dinv = np.zeros(self.n[r] * self.b * self.b)
        _check(lib().ref_par_info(self.h, r, C.byref(split), master.ctypes.data_as(C.c_void_p), dinv))
        return split.value, master, dinv

    def dcc_lists(self, r):
        """m_ex / g_ex of BasicDCCMap::CalcDOFMasters per neighbour"""
        m_ex, g_ex = [], []
        for k in range(len(self.peers[r])):
            nm, ng = C.c_int64(), C.c_int64()
            _check(lib().ref_par_dcc_lists(self.h, r, k, C.byref(nm), None, C.byref(ng), None))
            m, g = np.zeros(max(nm.value, 1), np.int32), np.zeros(max(ng.value, 1), np.int32)
            _check(lib().ref_par_dcc_lists(self.h, r, k, C.byref(nm), m.ctypes.data_as(C.c_void_p), C.byref(ng), g.ctypes.data_as(C.c_void_p)))
            m_ex.append(m[:nm.value].astype(np.int64))
            g_ex.append(g[:ng.value].astype(np.int64))
        return m_ex, g_ex

    def _pp(self, vecs):
        arr = (C.c_void_p * self.R)()
        for r in range(self.R):
            assert vecs[r].dtype == np.float64 and vecs[r].flags["C_CONTIGUOUS"]
            arr[r] = vecs[r].ctypes.data
        return arr

    def smooth(self, x, b, res, res_updated, update_res, x_zero, backward, status_b=0, status_res=0):
        """HybridBaseSmoother::Smooth / SmoothBack on all ranks; x CUMULATED, b / res DISTRIBUTED by default; x, res updated in place"""
        _check(lib().ref_par_smooth(self.h, self._pp(x), self._pp(b), self._pp(res), status_b, status_res, int(res_updated), int(update_res),
                                    int(x_zero), int(backward)))

    def mult(self, x):
        y = [np.zeros_like(v) for v in x]
        _check(lib().ref_par_mult(self.h, self._pp(x), self._pp(y)))
        return y

    def dis2co(self, vec):
        _check(lib().ref_par_exchange(self.h, 0, self._pp(vec)))

    def co2cu(self, vec):
        _check(lib().ref_par_exchange(self.h, 1, self._pp(vec)))


class RefParAMG:
    """The multi-rank preconditioner run by the reference's own code: per rank an AMGMatrix (SmoothV) over distributed levels with
    HybridGSSmoother and ProlMap, R ranks = R host threads.  Same arguments as oracle_par.OracleParAMG.  The local Galerkin
    products are the reference's RestrictMatrix, the merged matrix of the contracted level its CtrMap::DoAssembleMatrix, the
    gather / scatter around the serial coarse hierarchy its CtrMap::TransferF2C / TransferC2F (one group, master rank 0).
    Glue (not reference code): where the contraction sits (plugged in as the coarse solve instead of being a DOFMap step with
    ranks dropping out) and the CG loop (NGSolve's CGSolver restated)."""

    def __init__(self, A0, free0, peers0, ex0, prols, halos, ctr_maps, nested_prols, pinv=False, nested_free=None, sm_steps=1,
                 sm_symm=False, overlap=True):
        import scipy.sparse as sp
        L = lib()
        vp, ci, i64 = C.c_void_p, C.c_int, C.c_int64
        pp = C.POINTER(C.c_void_p)
        L.ref_paramg_new.argtypes, L.ref_paramg_new.restype = [ci, ci], vp
        L.ref_paramg_free.argtypes = [vp]
        L.ref_paramg_set_matrix.argtypes = [vp, ci, i64, ci, i64p, i32p, f64p, vp]
        L.ref_paramg_set_halo.argtypes = [vp, ci, ci, ci, i32p, i64p, i32p]
        L.ref_paramg_set_prol.argtypes = [vp, ci, ci, i64, ci, i64p, i32p, f64p]
        L.ref_paramg_set_contraction.argtypes = [vp, ci, i64, i64p, i64]
        L.ref_paramg_merged_matrix.argtypes, L.ref_paramg_merged_matrix.restype = [vp], vp
        L.ref_paramg_set_nested.argtypes = [vp, vp]
        L.ref_paramg_setup.argtypes = [vp, ci, ci, ci]
        L.ref_paramg_apply.argtypes = [vp, pp, pp]
        L.ref_paramg_mult.argtypes = [vp, pp, pp]
        L.ref_paramg_level_vec.argtypes = [vp, ci, ci, ci, f64p]
        L.ref_paramg_level_size.argtypes, L.ref_paramg_level_size.restype = [vp, ci, ci], i64
        assert not pinv, "pinv smoothers on distributed levels are not wired into the harness"
        self.R, self.npar = len(A0), len(prols)
        self.h = _ptr(L.ref_paramg_new(self.R, self.npar + 1))
        self.b0 = A0[0].bh
        self.n0 = [a.nrows * a.bh for a in A0]
        for r in range(self.R):
            fm = None if free0[r] is None else np.ascontiguousarray(free0[r], np.uint8)
            A = A0[r]
            _check(L.ref_paramg_set_matrix(self.h, r, A.nrows, A.bh, A.rowptr, A.col, A.val, None if fm is None else fm.ctypes.data_as(vp)))
        for l in range(self.npar + 1):
            peers, ex = (peers0, ex0) if l == 0 else halos[l]
            for r in range(self.R):
                pr = np.ascontiguousarray(peers[r], np.int32)
                ptr = np.zeros(len(pr) + 1, np.int64)
                for k in range(len(pr)):
                    ptr[k + 1] = ptr[k] + len(ex[r][k])
                exd = np.ascontiguousarray(np.concatenate([np.asarray(e, np.int32) for e in ex[r]] + [np.zeros(0, np.int32)]), np.int32)
                _check(L.ref_paramg_set_halo(self.h, r, l, len(pr), pr if len(pr) else np.zeros(1, np.int32), ptr,
                                             exd if len(exd) else np.zeros(1, np.int32)))
                if l == self.npar:
                    continue
                P = prols[l][r]
                _check(L.ref_paramg_set_prol(self.h, r, l, P.ncols, P.bw, P.rowptr, P.col if P.nnz else np.zeros(1, np.int32),
                                             P.val if P.nnz else np.zeros(1)))
        # contracted level: CtrMap of the reference (group = all ranks, master 0) merges the local matrices; serial hierarchy below
        self.nested = None
        if ctr_maps is not None:
            maps = [np.ascontiguousarray(m, np.int64) for m in ctr_maps]
            N = int(max(int(m.max()) for m in maps if len(m)) + 1)
            for r in range(self.R):
                _check(L.ref_paramg_set_contraction(self.h, r, len(maps[r]), maps[r] if len(maps[r]) else np.zeros(1, np.int64), N))
        _check(L.ref_paramg_setup(self.h, int(sm_steps), int(bool(sm_symm)), int(bool(overlap))))
        if ctr_maps is not None:
            self.A_merged = RefMat(L.ref_paramg_merged_matrix(self.h), owned=False).to_bsr()     # CtrMap::DoAssembleMatrix
            self.nested = RefAMG(self.A_merged, nested_free, nested_prols, sm_steps=sm_steps, sm_symm=sm_symm)
            _check(L.ref_paramg_set_nested(self.h, self.nested.h))

    def _coarsest_matrix(self, r):
        """local matrix of the contracted level (RestrictMatrix of the reference), fetched through a one-level-deeper handle trick:
        recomputed in python from the prolongations is avoided by asking the library"""
        L = lib()
        L.ref_paramg_level_matrix.argtypes, L.ref_paramg_level_matrix.restype = [C.c_void_p, C.c_int, C.c_int], C.c_void_p
        return RefMat(L.ref_paramg_level_matrix(self.h, r, self.npar), owned=False).to_bsr()

    def __del__(self):
        try:
            lib().ref_paramg_free(self.h)
        except Exception:
            pass

    def _pp(self, vecs):
        arr = (C.c_void_p * self.R)()
        for r in range(self.R):
            assert vecs[r].dtype == np.float64 and vecs[r].flags["C_CONTIGUOUS"]
            arr[r] = vecs[r].ctypes.data
        return arr

    def apply(self, b0):
        b = [np.ascontiguousarray(v, np.float64) for v in b0]
        x = [np.zeros_like(v) for v in b]
        _check(lib().ref_paramg_apply(self.h, self._pp(b), self._pp(x)))
        return x

    def mult(self, x):
        y = [np.zeros_like(v) for v in x]
        _check(lib().ref_paramg_mult(self.h, self._pp(x), self._pp(y)))
        return y

    def level_vec(self, which, l, r):
        n = lib().ref_paramg_level_size(self.h, r, l)
        out = np.zeros(n)
        _check(lib().ref_paramg_level_vec(self.h, r, {"x": 0, "rhs": 1, "res": 2}[which], l, out))
        return out

    def pcg(self, rhs, tol=1e-8, maxsteps=200):
        """CGSolver on parallel vectors (glue, same restatement as OracleParAMG.pcg): d DISTRIBUTED, w/s/u CUMULATED"""
        R = self.R
        dot = lambda a, b: float(sum(float(np.multiply(a[r], b[r]).sum()) for r in range(R)))
        d = [np.ascontiguousarray(v, np.float64).copy() for v in rhs]
        u = [np.zeros_like(v) for v in d]
        w = self.apply(d)
        s = [v.copy() for v in w]
        wdn = dot(w, d)
        err0 = np.sqrt(abs(wdn))
        errs, it = [err0], 0
        if wdn != 0.0:
            for it in range(1, maxsteps + 1):
                q = self.mult(s)
                wd = wdn
                alpha = wd / dot(s, q)
                for r in range(R):
                    u[r] += alpha * s[r]
                    d[r] -= alpha * q[r]
                w = self.apply(d)
                wdn = dot(w, d)
                beta = wdn / wd
                for r in range(R):
                    s[r] = w[r] + beta * s[r]
                err = np.sqrt(abs(wd))
                errs.append(err)
                if err < tol * err0:
                    break
        return u, it, np.array(errs)
