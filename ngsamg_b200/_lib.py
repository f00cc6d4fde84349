"""ctypes binding of libngsamg_b200.so (C ABI: include/ngsamg_b200.h).

There is no fallback of any kind: if the shared library is missing, fails to load, or no CUDA device is
present, every entry point raises.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "lib", "libngsamg_b200.so")
CSRC = os.path.join(_PKG, "csrc")

EXPORTS = [
    "ngsamg_b200_create", "ngsamg_b200_set_prolongations", "ngsamg_b200_finalize", "ngsamg_b200_destroy",
    "ngsamg_b200_last_error", "ngsamg_b200_apply", "ngsamg_b200_apply_add", "ngsamg_b200_spmv_add",
    "ngsamg_b200_apply_phases", "ngsamg_b200_level_sweep_kind", "ngsamg_b200_smooth", "ngsamg_b200_restrict", "ngsamg_b200_prolong_add", "ngsamg_b200_coarse_solve", "ngsamg_b200_pcg",
    "ngsamg_b200_num_levels", "ngsamg_b200_level_info", "ngsamg_b200_get_level_matrix",
    "ngsamg_b200_get_prolongation", "ngsamg_b200_get_level_vector", "ngsamg_b200_operator_complexity", "ngsamg_b200_operator_complexities",
    "ngsamg_b200_vcycle_bytes", "ngsamg_b200_last_ms", "ngsamg_b200_launch_count", "ngsamg_b200_rap_begin",
    "ngsamg_b200_matmul_begin", "ngsamg_b200_transpose_begin", "ngsamg_b200_spm_fetch",
    "ngsamg_b200_coarsen_begin", "ngsamg_b200_coarsen_fetch", "ngsamg_b200_profile_kernel",
    "ngsamg_b200_get_sweep_order", "ngsamg_b200_set_tunable", "ngsamg_b200_get_gs_blocks",
    "ngsamg_b200_create_parallel", "ngsamg_b200_nccl_unique_id", "ngsamg_b200_nccl_comm_init", "ngsamg_b200_nccl_comm_destroy",
    "ngsamg_b200_get_halo", "ngsamg_b200_get_hybrid", "ngsamg_b200_num_parallel_levels", "ngsamg_b200_halo_transport", "ngsamg_b200_get_contracted",
    "ngsamg_b200_get_contraction_map", "ngsamg_b200_hybrid_host_begin", "ngsamg_b200_hybrid_host_fetch",
    "ngsamg_b200_contract_host_begin", "ngsamg_b200_contract_host_fetch",
    "ngsamg_b200_coarsen_parallel_begin", "ngsamg_b200_coarsen_parallel_fetch",
    "ngsamg_b200_tile_schedule_begin", "ngsamg_b200_tile_schedule_hinted", "ngsamg_b200_tile_schedule_fetch", "ngsamg_b200_tiles_last_error", "ngsamg_b200_block_pinv", "ngsamg_b200_block_regularize",
]


class Csr(C.Structure):
    _fields_ = [("nrows", C.c_int64), ("ncols", C.c_int64), ("bh", C.c_int32), ("bw", C.c_int32),
                ("rowptr", C.c_void_p), ("col", C.c_void_p), ("val", C.c_void_p)]


class LevelInfo(C.Structure):
    _fields_ = [("n", C.c_int64), ("b", C.c_int32), ("nnz", C.c_int64), ("nnz_prol", C.c_int64),
                ("ncoarse", C.c_int64), ("bcoarse", C.c_int32), ("gs_depth", C.c_int32),
                ("bytes_matrix", C.c_int64), ("bytes_prol", C.c_int64), ("bytes_vec", C.c_int64)]


def build(force=False, verbose=False):
    """compile the CUDA extension in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    if force:
        subprocess.check_call(["make", "-C", CSRC, "clean"], stdout=subprocess.DEVNULL)
    out = None if verbose else subprocess.DEVNULL
    subprocess.check_call(["make", "-C", CSRC, "-j4"], stdout=out)
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError("ngsamg_b200: CUDA extension %s is missing -- run `python -c 'import __graft_entry__ as g; "
                           "g.build()'` (there is no CPU fallback)" % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    vp, i64, ci, dbl = C.c_void_p, C.c_int64, C.c_int, C.c_double
    L.ngsamg_b200_last_error.restype = C.c_char_p
    L.ngsamg_b200_create.argtypes = [C.c_char_p, C.POINTER(Csr), vp, vp, C.POINTER(C.c_char_p), C.POINTER(C.c_char_p), ci, ci,
                                     C.POINTER(vp)]
    L.ngsamg_b200_set_prolongations.argtypes = [vp, ci, C.POINTER(Csr)]
    L.ngsamg_b200_finalize.argtypes = [vp]
    L.ngsamg_b200_destroy.argtypes = [vp]
    L.ngsamg_b200_destroy.restype = None
    L.ngsamg_b200_apply.argtypes = [vp, vp, vp]
    L.ngsamg_b200_apply_add.argtypes = [vp, dbl, vp, vp]
    L.ngsamg_b200_spmv_add.argtypes = [vp, ci, dbl, vp, vp]
    L.ngsamg_b200_level_sweep_kind.argtypes = [vp, ci]
    L.ngsamg_b200_set_tunable.argtypes = [vp, C.c_char_p, dbl]
    L.ngsamg_b200_halo_transport.argtypes = [vp, ci]
    L.ngsamg_b200_get_gs_blocks.argtypes = [vp, ci, vp]
    L.ngsamg_b200_apply_phases.argtypes = [vp, vp, vp, vp]
    L.ngsamg_b200_smooth.argtypes = [vp, ci, vp, vp, vp, ci, ci, ci, ci]
    L.ngsamg_b200_restrict.argtypes = [vp, ci, vp, vp]
    L.ngsamg_b200_prolong_add.argtypes = [vp, ci, dbl, vp, vp]
    L.ngsamg_b200_coarse_solve.argtypes = [vp, vp, vp]
    L.ngsamg_b200_pcg.argtypes = [vp, vp, vp, dbl, ci, C.POINTER(ci), vp]
    L.ngsamg_b200_num_levels.argtypes = [vp]
    L.ngsamg_b200_level_info.argtypes = [vp, ci, C.POINTER(LevelInfo)]
    L.ngsamg_b200_get_level_matrix.argtypes = [vp, ci, vp, vp, vp]
    L.ngsamg_b200_get_prolongation.argtypes = [vp, ci, vp, vp, vp]
    L.ngsamg_b200_get_level_vector.argtypes = [vp, ci, ci, vp]
    L.ngsamg_b200_operator_complexity.argtypes = [vp]
    L.ngsamg_b200_operator_complexity.restype = dbl
    L.ngsamg_b200_operator_complexities.argtypes = [vp, vp, ci]
    L.ngsamg_b200_vcycle_bytes.argtypes = [vp]
    L.ngsamg_b200_vcycle_bytes.restype = dbl
    L.ngsamg_b200_last_ms.argtypes = [vp, ci]
    L.ngsamg_b200_last_ms.restype = dbl
    L.ngsamg_b200_launch_count.argtypes = [vp]
    L.ngsamg_b200_launch_count.restype = i64
    for nm in ("rap", "matmul"):
        getattr(L, "ngsamg_b200_%s_begin" % nm).argtypes = [C.POINTER(Csr), C.POINTER(Csr), ci, C.POINTER(vp), C.POINTER(i64),
                                                            C.POINTER(i64)]
    L.ngsamg_b200_transpose_begin.argtypes = [C.POINTER(Csr), ci, C.POINTER(vp), C.POINTER(i64), C.POINTER(i64)]
    L.ngsamg_b200_spm_fetch.argtypes = [vp, vp, vp, vp]
    L.ngsamg_b200_coarsen_begin.argtypes = [C.POINTER(Csr), vp, vp, ci, ci, dbl, dbl, ci, ci, C.POINTER(vp), C.POINTER(i64),
                                            C.POINTER(i64)]
    L.ngsamg_b200_coarsen_fetch.argtypes = [vp, vp, vp, vp, vp, vp]
    L.ngsamg_b200_get_sweep_order.argtypes = [vp, ci, vp]
    L.ngsamg_b200_profile_kernel.argtypes = [vp, ci, ci, ci, C.POINTER(dbl), C.POINTER(dbl)]
    # multi-rank
    L.ngsamg_b200_create_parallel.argtypes = [C.c_char_p, C.POINTER(Csr), vp, vp, vp, vp, C.POINTER(C.c_char_p), C.POINTER(C.c_char_p),
                                              ci, ci, C.POINTER(vp)]
    L.ngsamg_b200_nccl_unique_id.argtypes = [vp]
    L.ngsamg_b200_nccl_comm_init.argtypes = [vp, ci, ci, ci, C.POINTER(vp)]
    L.ngsamg_b200_nccl_comm_destroy.argtypes = [vp]
    L.ngsamg_b200_get_halo.argtypes = [vp, ci, vp, vp, vp, vp]
    L.ngsamg_b200_get_hybrid.argtypes = [vp, ci, ci, vp, vp, vp, vp, vp]
    L.ngsamg_b200_num_parallel_levels.argtypes = [vp]
    L.ngsamg_b200_get_contracted.argtypes = [vp]
    L.ngsamg_b200_get_contracted.restype = vp
    L.ngsamg_b200_get_contraction_map.argtypes = [vp, ci, vp, vp]
    L.ngsamg_b200_hybrid_host_begin.argtypes = [C.POINTER(Csr), vp, vp, vp, C.POINTER(vp), C.POINTER(i64), C.POINTER(i64)]
    L.ngsamg_b200_hybrid_host_fetch.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, vp, vp]
    L.ngsamg_b200_contract_host_begin.argtypes = [C.POINTER(Csr), vp, vp, vp, C.POINTER(vp), C.POINTER(i64), C.POINTER(i64), C.POINTER(i64)]
    L.ngsamg_b200_contract_host_fetch.argtypes = [vp, vp, vp, vp, vp, vp]
    L.ngsamg_b200_coarsen_parallel_begin.argtypes = [C.POINTER(Csr), vp, vp, vp, vp, ci, ci, dbl, dbl, ci, ci, C.POINTER(vp), C.POINTER(i64),
                                                     C.POINTER(i64), C.POINTER(C.c_int32), C.POINTER(i64)]
    L.ngsamg_b200_coarsen_parallel_fetch.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, vp]
    L.ngsamg_b200_tile_schedule_begin.argtypes = [C.POINTER(Csr), vp, vp, ci, ci, C.POINTER(vp), vp]
    L.ngsamg_b200_tile_schedule_hinted.argtypes = [C.POINTER(Csr), vp, vp, vp, ci, C.POINTER(vp), vp]
    L.ngsamg_b200_tile_schedule_fetch.argtypes = [vp, vp, vp, vp, vp, vp, vp]
    L.ngsamg_b200_tiles_last_error.restype = C.c_char_p
    _lib = L
    return L


class NgsAMGError(RuntimeError):
    """mirrors ngcore::Exception thrown by the reference (e.g. amg_pc.cpp:430)"""


def check(rc):
    if rc != 0:
        raise NgsAMGError(lib().ngsamg_b200_last_error().decode())


def vec(a, n, what="vector"):
    """a float64 vector of at least n entries -> void*.  The C ABI reads/writes n doubles through the pointer, so anything else
    (float32, a strided slice, a short array) would silently give garbage or overrun memory: raise instead.  Nothing is copied:
    output arrays must be written in place."""
    if isinstance(a, np.ndarray):
        if a.dtype != np.float64:
            raise TypeError("%s: expected float64, got %s" % (what, a.dtype))
        if not a.flags["C_CONTIGUOUS"]:
            raise ValueError("%s: array must be C-contiguous" % what)
        if a.size < n:
            raise ValueError("%s: %d entries, at least %d expected" % (what, a.size, n))
    elif hasattr(a, "data_ptr"):
        import torch
        if a.dtype != torch.float64:
            raise TypeError("%s: expected a float64 tensor, got %s" % (what, a.dtype))
        if not a.is_contiguous():
            raise ValueError("%s: tensor must be contiguous" % what)
        if a.numel() < n:
            raise ValueError("%s: %d entries, at least %d expected" % (what, a.numel(), n))
    elif a is not None and not isinstance(a, int):
        raise TypeError("%s: unsupported buffer type %r" % (what, type(a)))
    return ptr(a)


def ptr(a):
    """numpy array / torch CUDA tensor / raw int -> void*"""
    if a is None:
        return None
    if isinstance(a, int):
        return C.c_void_p(a)
    if isinstance(a, np.ndarray):
        return a.ctypes.data_as(C.c_void_p)
    if hasattr(a, "data_ptr"):  # torch tensor (device memory is used in place)
        return C.c_void_p(a.data_ptr())
    raise TypeError("unsupported buffer type %r" % type(a))
