"""one hierarchy, many settings: per-level times of the triangular half-sweeps under different run-time tunables (SetTunable)"""
import os, sys, itertools
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ngsamg_b200 as ng
from ngsamg_b200 import synthetic as S
n = int(sys.argv[1]) if len(sys.argv) > 1 else 151
p = S.poisson3d_kuhn(n)
A = ng.SparseMatrix(p["n"], p["n"], 1, 1, p["rowptr"], p["col"], p["val"])
pc = ng.h1_scal(A, p["free"])
NL = pc.GetNLevels()
levels = [int(x) for x in os.environ.get("LEVELS", ",".join(str(l) for l in range(1, NL - 1))).split(",")]
print("levels", [(l, pc.level_info(l).n, pc.level_info(l).gs_depth, pc.SweepKind(l)) for l in range(NL)])

def run(tag):
    row = []
    for l in levels:
        f, _ = pc.ProfileKernel("gs_tri_fwd", level=l, reps=10)
        b, _ = pc.ProfileKernel("gs_tri_bwd", level=l, reps=10)
        row.append("%d:%.1f/%.1f" % (l, f * 1e3, b * 1e3))
    print("%-60s %s  [us fwd/bwd]" % (tag, "  ".join(row)), flush=True)

settings = [
    {},
    {"tri_rm": 0},
    {"tri_prepoll": 0},
    {"tri_sleep_ns": 0}, {"tri_sleep_ns": 20}, {"tri_sleep_ns": 50}, {"tri_sleep_ns": 200}, {"tri_sleep_ns": 400},
    {"tri_rm_rows_per_warp": 1}, {"tri_rm_rows_per_warp": 2}, {"tri_rm_rows_per_warp": 8}, {"tri_rm_rows_per_warp": 16},
    {"tri_rm_rows_per_warp": 32}, {"tri_rm_rows_per_warp": 64},
    {"tri_rm_rows_per_warp": 16, "tri_sleep_ns": 20}, {"tri_rm_rows_per_warp": 16, "tri_prepoll": 0},
    {"tri_pollmode": 1}, {"tri_pollmode": 3},
]
defaults = {"tri_rm": 1, "tri_prepoll": 1, "tri_sleep_ns": 100, "tri_rm_rows_per_warp": 4, "tri_pollmode": 0}
for st in settings:
    for k, v in defaults.items():
        pc.SetTunable(k, st.get(k, v))
    run(str(st))
