"""ctypes binding of oracle/_ref/libngsamg_ref_bgs.so: the reference's OWN block Gauss-Seidel update routines
(BSmoother2<TM>::BSBlock::RichardsonUpdate / RichardsonUpdate_RES, loc_block_gssmoother_impl.hpp:244-268, 516-541), cut out of
/root/reference at build time and compiled against oracle/ref_pin/ngs_standin_bgs.hpp.  TEST INFRASTRUCTURE ONLY."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ORACLE = os.path.dirname(_HERE)
_SO = os.path.join(_ORACLE, "_ref", "libngsamg_ref_bgs.so")
_lib = None
_avail = None


def build():
    if os.path.exists(_SO) and not os.path.isdir(os.environ.get("NGSAMG_REFERENCE", "/root/reference")):
        return
    if os.path.isdir(os.environ.get("NGSAMG_REFERENCE", "/root/reference")):
        subprocess.run(["make", "-C", _ORACLE, "_ref/libngsamg_ref_bgs.so"], check=True, capture_output=True)
    if not os.path.exists(_SO):
        raise RuntimeError("oracle/_ref/libngsamg_ref_bgs.so is not built and the reference tree is not present")


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        i64, ci, vp = C.c_int64, C.c_int, C.c_void_p
        L.ref_bgs_sweep.argtypes = [i64, ci, vp, vp, vp, i64, vp, vp, vp, vp, vp, vp, ci, ci]
        L.ref_bgs_sweep.restype = ci
        L.ref_bgs_smooth_wo.argtypes = [i64, ci, vp, vp, vp, i64, vp, vp, vp, vp, vp, vp, vp, ci, ci, ci, ci, ci, ci]
        L.ref_bgs_smooth_wo.restype = ci
        _lib = L
    return _lib


def available():
    global _avail
    if _avail is None:
        try:
            lib()
            _avail = True
        except Exception as e:
            sys.stderr.write("[oracle/ref_pin] bgs reference library unavailable: %s\n" % e)
            _avail = False
    return _avail


def _prepare(A, block_of):
    import scipy.sparse as sp
    n, b = int(A.nrows), int(A.bh)
    block_of = np.asarray(block_of).astype(np.int64)
    nb = int(block_of.max()) + 1 if n else 0
    order = np.argsort(block_of, kind="stable")
    order = order[block_of[order] >= 0]
    bptr = np.zeros(nb + 1, np.int64)
    np.add.at(bptr, block_of[order] + 1, 1)
    bptr = np.cumsum(bptr)
    bverts = order.astype(np.int32)
    S = sp.bsr_matrix((np.asarray(A.val, dtype=np.float64).reshape(-1, b, b), np.asarray(A.col), np.asarray(A.rowptr)), shape=(n * b, n * b)).tocsr()
    dinv, off = [], np.zeros(nb + 1, np.int64)
    for k in range(nb):
        verts = bverts[bptr[k]:bptr[k + 1]]
        dofs = (verts[:, None].astype(np.int64) * b + np.arange(b)[None, :]).ravel()
        D = S[dofs][:, dofs].toarray() if len(dofs) else np.zeros((0, 0))
        Di = np.linalg.inv(D) if len(dofs) else D
        dinv.append(Di.ravel())
        off[k + 1] = off[k] + Di.size
    dinv = np.ascontiguousarray(np.concatenate(dinv) if dinv else np.zeros(0))
    return (n, b, np.ascontiguousarray(A.rowptr, dtype=np.int64), np.ascontiguousarray(A.col, dtype=np.int32),
            np.ascontiguousarray(A.val, dtype=np.float64), nb, bptr, bverts, dinv, off)


def smooth_wo(A, block_of, x, b, res, res_updated, update_res, x_zero, reverse=False, steps=1, symm=False, ref_setup=False):
    """BSmoother2::SmoothWO (loc_block_gssmoother_impl.hpp:655-668) through the reference's own IterateBlocks / SmoothSimple / SmoothRESSimple /
    RichardsonUpdate[_RES], all blocks in one group; x and res are updated in place.
    ref_setup: the blocks are built by the reference's own BSBlock::SetFromSPMat (:67-132, dense inverse by the stand-in's Gauss-Jordan)
    instead of being laid out by the harness with numpy's inverses."""
    n, bs, rp, ci, av, nb, bptr, bverts, dinv, off = _prepare(A, block_of)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    if ref_setup:
        p_dinv = None
    else:
        p_dinv = p(dinv)
    b = np.ascontiguousarray(b, dtype=np.float64)
    assert x.dtype == np.float64 and res.dtype == np.float64 and x.flags.c_contiguous and res.flags.c_contiguous
    rc = lib().ref_bgs_smooth_wo(n, bs, p(rp), p(ci), p(av), nb, p(bptr), p(bverts), p_dinv, p(off), p(x), p(b), p(res), int(steps), int(res_updated),
                                 int(update_res), int(x_zero), int(reverse), int(symm))
    if rc != 0:
        raise RuntimeError("ref_bgs_smooth_wo failed (%d)" % rc)


def sweep(A, block_of, x, r, res_form, reverse=False):
    """one sweep of the reference's block updates over all blocks (ascending, or descending when `reverse`), in place.
    A: block CSR object with nrows, bh, rowptr, col, val; block_of[v] = block of vertex v (-1: none);
    res_form False: RichardsonUpdate(x, rhs = r); True: RichardsonUpdate_RES(x, res = r).  The dense block inverses are numpy's."""
    import scipy.sparse as sp
    n, b = int(A.nrows), int(A.bh)
    block_of = np.asarray(block_of).astype(np.int64)
    nb = int(block_of.max()) + 1 if n else 0
    order = np.argsort(block_of, kind="stable")
    order = order[block_of[order] >= 0]
    bptr = np.zeros(nb + 1, np.int64)
    np.add.at(bptr, block_of[order] + 1, 1)
    bptr = np.cumsum(bptr)
    bverts = order.astype(np.int32)
    S = sp.bsr_matrix((np.asarray(A.val, dtype=np.float64).reshape(-1, b, b), np.asarray(A.col), np.asarray(A.rowptr)), shape=(n * b, n * b)).tocsr()
    dinv, off = [], np.zeros(nb + 1, np.int64)
    for k in range(nb):
        verts = bverts[bptr[k]:bptr[k + 1]]
        dofs = (verts[:, None].astype(np.int64) * b + np.arange(b)[None, :]).ravel()
        D = S[dofs][:, dofs].toarray() if len(dofs) else np.zeros((0, 0))
        Di = np.linalg.inv(D) if len(dofs) else D
        dinv.append(Di.ravel())
        off[k + 1] = off[k] + Di.size
    dinv = np.ascontiguousarray(np.concatenate(dinv) if dinv else np.zeros(0))
    rp = np.ascontiguousarray(A.rowptr, dtype=np.int64)
    ci = np.ascontiguousarray(A.col, dtype=np.int32)
    av = np.ascontiguousarray(A.val, dtype=np.float64)
    assert x.dtype == np.float64 and r.dtype == np.float64 and x.flags.c_contiguous and r.flags.c_contiguous
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    rc = lib().ref_bgs_sweep(n, b, p(rp), p(ci), p(av), nb, p(bptr), p(bverts), p(dinv), p(off), p(x), p(r), 1 if res_form else 0, 1 if reverse else 0)
    if rc != 0:
        raise RuntimeError("ref_bgs_sweep failed (%d)" % rc)
