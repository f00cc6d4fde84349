"""micro-benchmark: cost of one dependency hop of the sync-free Gauss-Seidel sweep.
Matrix: row i couples to rows i-stride and i+stride (stride multiple of 32) => dependency level = i // stride, every
level is stride/32 slices wide, depth = n/stride.  time(gs_tri_fwd)/depth = hop latency."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ngsamg_b200 as ng
import scipy.sparse as sp

def chain(n, stride, extra=0):
    i = np.arange(n)
    rows, cols, vals = [i], [i], [np.full(n, 4.0)]
    for d in (stride,):
        m = i >= d
        rows += [i[m], i[m] - d]; cols += [i[m] - d, i[m]]; vals += [np.full(m.sum(), -1.0)] * 2
    for e in range(extra):   # extra dependencies further back (already published)
        d = stride * (2 + e) + 32
        m = i >= d
        rows += [i[m], i[m] - d]; cols += [i[m] - d, i[m]]; vals += [np.full(m.sum(), -0.1)] * 2
    A = sp.coo_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(n, n)).tocsr()
    return ng.SparseMatrix.from_scipy(A)

def pairs(n):
    return ng.SparseMatrix(n, n // 2, 1, 1, np.arange(n + 1), np.arange(n) // 2, np.ones(n))

for (n, stride, extra) in [(32 * 20000, 32, 0), (32 * 20000, 256, 0), (256 * 4000, 256, 0), (8192 * 500, 8192, 0), (8192 * 500, 8192, 5)]:
    A = chain(n, stride, extra)
    pc = ng.h1_scal(A, None, prolongations=[pairs(n)], ngs_amg_clev="none",
                    **{k: v for k, v in [kv.split("=") for kv in os.environ.get("NGSAMG_FLAGS", "").split(",") if "=" in kv]})
    depth = pc.level_info(0).gs_depth
    for name in ("gs_tri_fwd", "gs_tri_bwd"):
        ms, by = pc.ProfileKernel(name, level=0, reps=5)
        print("n=%d stride=%d extra=%d depth=%d %s: %.3f ms -> %.3f us/hop  (%.0f GB/s)" % (n, stride, extra, depth, name, ms, 1e3 * ms / depth, by / ms / 1e6))
    del pc
