"""per-level times of the triangular half-sweeps of a P2 elasticity hierarchy (3x3 / 6x6 blocks) under run-time tunables"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ngsamg_b200 as ng
from ngsamg_b200 import synthetic as S
n = int(sys.argv[1]) if len(sys.argv) > 1 else 76
ny = max(3, n // 3 + 1)
p = S.elasticity3d_p2_kuhn_stencil(n, ny, ny)
A = ng.SparseMatrix(p["n"], p["n"], 3, 3, p["rowptr"], p["col"], p["val"])
pc = ng.elast_3d(A, p["free"], vertex_xyz=p["xyz"])
NL = pc.GetNLevels()
print("levels", [(l, pc.level_info(l).n, pc.level_info(l).b, pc.level_info(l).gs_depth, pc.SweepKind(l)) for l in range(NL)])
levels = [int(x) for x in os.environ.get("LEVELS", "0,1,2,3").split(",")]

def run(tag):
    row = []
    for l in levels:
        if l >= NL - 1: continue
        f, _ = pc.ProfileKernel("gs_tri_fwd", level=l, reps=5)
        b, _ = pc.ProfileKernel("gs_tri_bwd", level=l, reps=5)
        row.append("%d:%.0f/%.0f" % (l, f * 1e3, b * 1e3))
    print("%-64s %s  [us fwd/bwd]" % (tag, "  ".join(row)), flush=True)

defaults = {"tri_gate_all": 1, "tri_prepoll": 1, "tri_sleep_ns": 100, "tri_regate": 1, "tri_ctas_per_sm": 0, "tri_rm": 1, "tri_small_rows": 1000000}
settings = [{}, {"tri_gate_all": 0}, {"tri_prepoll": 0}, {"tri_gate_all": 0, "tri_sleep_ns": 0}, {"tri_ctas_per_sm": 1}, {"tri_ctas_per_sm": 1, "tri_gate_all": 0},
            {"tri_small_rows": 0}, {"tri_small_rows": 0, "tri_gate_all": 0}, {"tri_rm": 0}]
for st in settings:
    for k, v in defaults.items():
        pc.SetTunable(k, st.get(k, v))
    run(str(st))
