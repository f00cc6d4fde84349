"""A/B of the two block-SpMV designs on the 6x6 levels of the P2 elasticity hierarchy: SELL-32 thread-per-block-row (element-planar blocks,
every warp load one coalesced 256-byte segment) against row-major BSR records with one warp per block row (north_star's layout)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ngsamg_b200 as ng
from ngsamg_b200 import synthetic as S
n = int(sys.argv[1]) if len(sys.argv) > 1 else 151
ny = max(3, n // 3 + 1)
p = S.elasticity3d_p2_kuhn_stencil(n, ny, ny)
A = ng.SparseMatrix(p["n"], p["n"], 3, 3, p["rowptr"], p["col"], p["val"])
pc = ng.elast_3d(A, p["free"], vertex_xyz=p["xyz"])
NL = pc.GetNLevels()
print("levels", [(l, pc.level_info(l).n, pc.level_info(l).b, pc.level_info(l).nnz) for l in range(NL)])
for tag, rows in (("SELL-32 thread per block row (k_sell_spmv)", 0), ("row-major BSR, warp per block row (k_rm_spmv)", 1e12)):
    pc.SetTunable("rm_spmv_rows", rows)
    pc.SetTunable("spmv_small_rows", 0)
    for l in range(1, min(NL - 1, 4)):
        out = []
        for name in ("gs_upass", "gs_lpass", "spmv"):
            ms, by = pc.ProfileKernel(name, level=l, reps=10)
            out.append("%s %.3f ms %.0f GB/s" % (name, ms, by / ms / 1e6))
        print("%-48s level %d (%d rows): %s" % (tag, l, pc.level_info(l).n, "; ".join(out)), flush=True)
