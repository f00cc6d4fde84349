#!/usr/bin/env python
"""bench.py -- PCG+AMG solve throughput and V-cycle HBM bandwidth on B200 (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            # our CUDA path, one rank per GPU
  python bench.py --impl reference --gpus N --steps K ...  # CPU arm: the reference's own functions (oracle/_ref), else the oracle port

A "step" is one complete PCG+AMG solve (tol 1e-8, V(1,1)-cycle preconditioner) of the synthetic 3D Poisson P1 problem
(Kuhn tets on the unit cube, Dirichlet on x=0 and y=1, f=1).  N=1 workload = BASELINE.json configs[1]: 311^3 = 30.1 M DOFs.
value   = DOFs solved per second (whole job) with rhs/solution resident in HBM (the reference prints the same unit,
          "dofs / (sec * np)", tests/h1/amg_utils.py:358); `solve_s`, `iterations`, `vcycle_ms`, `vcycle_gbs` ride along.
e2e     = the same through the C ABI with HOST numpy buffers (H2D of the rhs and D2H of the solution inside the timed region).
roofline= the dominant kernel of the V-cycle (level-0 Gauss-Seidel triangular sweep) timed alone with CUDA events on the
          library's stream, algorithmic bytes / time against the measured HBM peak (MEASURED_PEAKS.json).
The matrix (5.5 GB at N=311) is far larger than the 126 MB L2, so no explicit L2 flush is needed between iterations.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

DEFAULT_N = 311          # 311^3 = 30,080,231 DOFs (configs[1])
CPU_SAMPLE_N = 91        # 91^3 = 753,571 DOFs: ~10-20 s of single-core CPU work for setup + solve
TOL = 1e-8


def rank_workload(n, rank, world):
    """grid size of the subdomain a rank owns: n^3 vertices per GPU (weak scaling).  At N > 1 the ranks solve ONE global problem:
    the mesh is cut into N sub-boxes of n^3 vertices (1x1x2, 1x2x2, 2x2x2 for N = 2, 4, 8: 621^3 = 239 M DOFs at N = 8 with the
    default n = 311, BASELINE.json configs[3]); interface DOFs are shared (NGSolve-style duplicated DOFs), hybrid smoothers + NCCL
    halo exchange per sweep, coarse levels contracted onto rank 0."""
    return n


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


class ClockSampler:
    """samples SM clocks / throttle reasons with nvidia-smi while the timed region runs"""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def make_problem(n, problem="poisson"):
    from ngsamg_b200 import synthetic as S
    import ngsamg_b200 as ng
    if problem == "elasticity_p2":
        # BASELINE.json configs[2]: nodal-P2 elasticity beam (examples/elasticity/beamP2.py; AMG on all P2 nodes, `subset=nodalp2`):
        # n x (n/3) x (n/3) vertices -> (2n-1) x ... P2 nodes with 3 DOFs each; n = 151 gives 3.07 M nodes = 9.2 M DOFs
        ny = max(3, n // 3 + 1)
        p = S.elasticity3d_p2_kuhn_stencil(n, ny, ny)
        A = ng.SparseMatrix(p["n"], p["n"], 3, 3, p["rowptr"], p["col"], p["val"])
        return p, A
    if problem == "elasticity_jump":
        # BASELINE.json configs[4] on ONE GPU: P1 elasticity with a jumping Young's modulus (checkerboard of 8^3-cell boxes, contrast 1e4; the
        # 3D analogue of tests/elasticity/mdim/jump/test_2d_jump_lo.py); n = 201 gives 2.05 M vertices = 6.15 M DOFs (the per-GPU share of ~50 M on 8)
        p = S.elasticity3d_kuhn_jump_stencil(n, n // 2 + 1, n // 2 + 1, S.checkerboard_modulus(8, 1e4))
        A = ng.SparseMatrix(p["n"], p["n"], 3, 3, p["rowptr"], p["col"], p["val"])
        return p, A
    if problem == "elasticity":
        # P1 elasticity beam (3x3 blocks, 6x6 on the coarse levels): nx x (nx/2) x (nx/2) vertices, clamped at x=0, body force (0,x,0)
        p = S.elasticity3d_kuhn_stencil(n, n // 2 + 1, n // 2 + 1)
        A = ng.SparseMatrix(p["n"], p["n"], 3, 3, p["rowptr"], p["col"], p["val"])
        return p, A
    p = S.poisson3d_kuhn(n)
    A = ng.SparseMatrix(p["n"], p["n"], 1, 1, p["rowptr"], p["col"], p["val"])
    return p, A


def cpu_reference_run(n, steps, warmup, tol=TOL):
    """the CPU arm, single thread like one MPI rank of the reference.
    kind "reference": oracle/_ref/libngsamg_ref.so -- the reference's OWN code for the path (RestrictMatrix / MatMultABImpl /
    TransposeSPMImpl for the Galerkin products, GSS3 sweeps, ProlMap transfers, AMGMatrix::SmoothV; cut out of the reference at build
    time and compiled against a stand-in for the NGSolve containers, oracle/ref_pin/README.md) under a CG loop; used whenever that
    library is present.  kind "port": the oracle's C restatement of the same functions (bit-identical results, tests/test_ref_pin.py).
    The hierarchy (prolongations) comes from the product's host-side coarsening in both cases."""
    import ngsamg_b200 as ng
    from oracle import oracle as O
    from oracle.ref_pin import ref as R
    O.build()
    kind = "port"
    if R.available():
        try:
            R.lib()
            kind = "reference"
        except Exception as e:            # no compiler / stale tree: fall back to the port, say so
            sys.stderr.write("[bench] oracle/_ref unusable (%s): CPU arm runs the oracle port\n" % e)
    p, A = make_problem(n)
    to_o = lambda M: O.Bsr(M.nrows, M.ncols, M.bh, M.bw, M.rowptr, M.col, M.val)
    t0 = time.time()
    prols, cur, fm = [], A, p["free"]
    amg = R.RefAMG(to_o(A), p["free"]) if kind == "reference" else None
    while cur.nrows > 50 and len(prols) + 1 < 10:
        P, _, _ = ng.coarsen(cur, fm)
        if P.ncols == 0 or P.ncols > 0.8 * cur.nrows:
            break
        prols.append(P)
        Po = to_o(P)
        Ac = amg.add_prol(Po) if kind == "reference" else O.restrict_matrix(O.transpose(Po), to_o(cur), Po)
        cur, fm = ng.SparseMatrix(Ac.nrows, Ac.ncols, 1, 1, Ac.rowptr, Ac.col, Ac.val), None
    if kind == "reference":
        amg.finalize()
    else:
        amg = O.OracleAMG(to_o(A), p["free"], [to_o(P) for P in prols])
    setup_s = time.time() - t0
    for _ in range(warmup):
        amg.pcg(p["rhs"], tol=tol, maxsteps=200)
    times, its = [], 0
    for _ in range(steps):
        t = time.time()
        _, its, errs = amg.pcg(p["rhs"], tol=tol, maxsteps=200)
        times.append(time.time() - t)
    tv = time.time()
    for _ in range(3):
        amg.apply(p["rhs"])
    vcycle_s = (time.time() - tv) / 3
    return dict(ndof=p["n"], solve_s=float(np.mean(times)), iterations=int(its), setup_s=setup_s, vcycle_s=vcycle_s,
                levels=amg.nlevels, kind=kind)


CPU_KIND_TEXT = {"reference": "the reference's own functions (oracle/_ref: RestrictMatrix, GSS3, ProlMap, AMGMatrix::SmoothV compiled against "
                              "an NGSolve container stand-in) under a CG loop",
                 "port": "the oracle port of the reference algorithm (oracle/_ref not present)"}


def cpu_reference_run_parallel(n, world, steps, warmup, tol=TOL):
    """N > 1: the reference's MPI-parallel CPU solve restated (oracle/cpu_pipeline.py): `world` ranks on `world` host threads, the
    same box partition as the GPU arm with n^3 vertices per rank, hybrid Gauss-Seidel smoothers, DCC halo exchange, CtrMap."""
    from ngsamg_b200 import synthetic as S
    from oracle import cpu_pipeline as CP
    from oracle import oracle as O
    from oracle import oracle_par as OP
    from oracle.ref_pin import ref as R
    O.build()
    kind = "port"
    if R.available():
        try:
            R.lib()
            kind = "reference"
        except Exception as e:
            sys.stderr.write("[bench] oracle/_ref unusable (%s): CPU arm runs the oracle port\n" % e)
    grid = S.bench_grid(world)
    parts = [S.box_poisson3d(n, grid, r) for r in range(world)]
    t0 = time.time()
    # kind "reference": the hierarchy runs inside oracle/_ref -- the reference's OWN HybridGSSmoother / DCCMap / ProlMap /
    # AMGMatrix::SmoothV per rank, `world` ranks = `world` host threads over an in-process MPI stand-in (oracle/ref_pin/README.md)
    amg, info = CP.build(parts, ctr_nv=20000, engine="reference" if kind == "reference" else "oracle")
    setup_s = time.time() - t0
    OP.set_threads(world)
    rhs = [p["rhs"] * p["free"] for p in parts]
    for _ in range(warmup):
        amg.pcg(rhs, tol=tol, maxsteps=200)
    times, its = [], 0
    for _ in range(steps):
        t = time.time()
        _, its, _ = amg.pcg(rhs, tol=tol, maxsteps=200)
        times.append(time.time() - t)
    tv = time.time()
    for _ in range(3):
        amg.apply(rhs)
    vcycle_s = (time.time() - tv) / 3
    ndof = int(sum(p["n_master"] for p in parts))
    return dict(ndof=ndof, solve_s=float(np.mean(times)), iterations=int(its), setup_s=setup_s, vcycle_s=vcycle_s,
                levels=info["distributed_levels"] + info["nested_levels"], grid=grid, info=info, kind=kind)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if args.problem != "poisson":
        # the CPU arm is wired for the headline workload only; say so instead of silently timing another problem
        print(json.dumps({"impl": "reference", "unavailable": "the CPU reference arm runs the Poisson workload (configs[1]/[3]) only, not --problem %s" % args.problem}), flush=True)
        return
    world = max(1, args.gpus)
    if world > 1:
        n = args.cpu_n_par
        r = cpu_reference_run_parallel(n, world, max(1, args.steps), max(0, args.warmup))   # exactly K timed solves after W warm-up solves
        val = r["ndof"] / r["solve_s"]
        line = {
            "impl": "reference", "metric": "pcg_amg_solve_dofs_per_s", "value": val, "unit": "DOF/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["solve_s"] * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "3D Poisson P1 (Kuhn tets), ONE global problem cut into %d sub-boxes (%dx%dx%d), h1_scal + CG to 1e-8, hybrid "
                                   "Gauss-Seidel; CPU sample: %d^3 vertices per rank = %d DOFs of the %d^3-per-GPU workload"
                                   % ((world,) + tuple(r["grid"]) + (n, r["ndof"], args.n)), "tol": TOL, "levels": r["levels"],
                       "multi_rank": r["info"]},
            "solve_s": r["solve_s"], "iterations": r["iterations"], "vcycle_ms": r["vcycle_s"] * 1e3, "setup_s": r["setup_s"],
            "cpu_baseline": {"value": val, "unit": "DOF/s", "cores": world, "kind": r["kind"],
                             "sample": ("the reference's own multi-rank functions (oracle/_ref: HybridGSSmoother, DCCMap exchange, ProlMap, "
                                        "AMGMatrix::SmoothV per rank over an in-process MPI stand-in; contraction and CG are glue)"
                                        if r["kind"] == "reference" else
                                        "multi-rank oracle (restated MPI path: hybrid smoothers, DCC exchange, CtrMap)") +
                                       ", %d ranks on %d host threads, %d^3 vertices per rank = %d DOFs (the reference as a whole needs "
                                       "NGSolve/MPI and cannot be built here)" % (world, world, n, r["ndof"])},
            "e2e": {"value": val, "unit": "DOF/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        }
        print(json.dumps(line), flush=True)
        return
    n = args.cpu_n
    r = cpu_reference_run(n, max(1, args.steps), max(0, args.warmup))   # exactly K timed solves after W warm-up solves
    val = r["ndof"] / r["solve_s"]
    # "all the host threads it can use": the reference is an MPI code -- on a CPU node it runs one rank per core.  The same sample size is
    # therefore also solved by R = min(8, cores) ranks of the reference's multi-rank path (hybrid smoothers, DCC exchange; R host threads);
    # the line reports the FASTER of the two as the CPU arm and keeps both under `variants`.
    variants = {"serial_1_core": {"value": val, "cores": 1, "ndof": r["ndof"], "solve_s": r["solve_s"], "iterations": r["iterations"], "kind": r["kind"]}}
    R = 1
    while R * 2 <= min(8, os.cpu_count() or 1):
        R *= 2
    if R > 1 and not args.cpu_serial_only:
        try:
            npr = max(21, int(round(n / R ** (1.0 / 3.0))))     # about the same global size as the serial sample
            rp = cpu_reference_run_parallel(npr, R, max(1, args.steps), max(0, args.warmup))
            vp = rp["ndof"] / rp["solve_s"]
            variants["ranks_%d" % R] = {"value": vp, "cores": R, "ndof": rp["ndof"], "solve_s": rp["solve_s"], "iterations": rp["iterations"], "kind": rp["kind"],
                                         "vertices_per_rank": npr ** 3}
            if vp > val:
                line = {
                    "impl": "reference", "metric": "pcg_amg_solve_dofs_per_s", "value": vp, "unit": "DOF/s", "n_gpus": args.gpus,
                    "steps": args.steps, "warmup": args.warmup, "ms_per_step": rp["solve_s"] * 1e3, "higher_is_better": True,
                    "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                    "config": {"workload": "3D Poisson P1 unit cube (Kuhn tets), h1_scal + CG to 1e-8; CPU sample: %d ranks (one per host thread) x %d^3 vertices = %d DOFs "
                                           "of the %d^3 workload, hybrid Gauss-Seidel + DCC exchange between the ranks" % (R, npr, rp["ndof"], args.n), "tol": TOL, "levels": rp["levels"]},
                    "solve_s": rp["solve_s"], "iterations": rp["iterations"], "vcycle_ms": rp["vcycle_s"] * 1e3, "setup_s": rp["setup_s"],
                    "cpu_baseline": {"value": vp, "unit": "DOF/s", "cores": R, "kind": rp["kind"],
                                     "sample": "PCG+AMG solve by the reference's own multi-rank functions (oracle/_ref) on %d host threads, %d DOFs; the single-core serial "
                                               "run of the same sample size reaches %.3g DOF/s" % (R, rp["ndof"], val)},
                    "variants": variants,
                    "e2e": {"value": vp, "unit": "DOF/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                }
                print(json.dumps(line), flush=True)
                return
        except Exception as e:      # the multi-rank CPU pipeline is optional: the serial arm stands on its own
            sys.stderr.write("[bench] multi-rank CPU variant failed (%s): reporting the serial arm\n" % e)
    line = {
        "impl": "reference", "metric": "pcg_amg_solve_dofs_per_s", "value": val, "unit": "DOF/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["solve_s"] * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "3D Poisson P1 unit cube (Kuhn tets), h1_scal + CG to 1e-8; CPU sample %d^3 = %d DOFs of the "
                               "%d^3 workload" % (n, r["ndof"], args.n), "tol": TOL, "levels": r["levels"]},
        "solve_s": r["solve_s"], "iterations": r["iterations"], "vcycle_ms": r["vcycle_s"] * 1e3,
        "cpu_baseline": {"value": val, "unit": "DOF/s", "cores": 1, "kind": r["kind"],
                         "sample": "PCG+AMG solve by %s, %d^3 = %d DOFs, 1 thread (the reference as a whole needs NGSolve/MPI and cannot be "
                                   "built here)" % (CPU_KIND_TEXT[r["kind"]], n, r["ndof"])},
        "variants": variants,
        "e2e": {"value": val, "unit": "DOF/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--size", dest="n", type=int, default=int(os.environ.get("NGSAMG_BENCH_N", DEFAULT_N)))
    ap.add_argument("--cpu-n", type=int, default=int(os.environ.get("NGSAMG_BENCH_CPU_N", CPU_SAMPLE_N)))
    ap.add_argument("--cpu-n-par", type=int, default=int(os.environ.get("NGSAMG_BENCH_CPU_N_PAR", 61)),
                    help="vertices per axis and rank of the multi-rank CPU reference arm (N > 1)")
    ap.add_argument("--problem", default="poisson", choices=["poisson", "elasticity", "elasticity_p2", "elasticity_jump"],
                    help="poisson = BASELINE.json configs[1] (headline); elasticity = P1 beam with elast_3d; elasticity_p2 = BASELINE.json "
                         "configs[2], nodal-P2 beam (--size 151 = 9.2 M DOFs); elasticity_jump = the configs[4] workload on one GPU (P1, modulus jumping by "
                         "1e4 on a checkerboard, --size 201 = 6.15 M DOFs); all secondary, reported on request")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-serial-only", action="store_true", help="reference arm at N=1: skip the multi-rank CPU variant")
    ap.add_argument("--no-multicolor", action="store_true", help="skip the separately reported multicolour-smoother variant")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
        return

    import faulthandler
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # NCCL's banner / warnings go to stderr: stdout carries ONE JSON line
    import torch
    import ngsamg_b200 as ng

    # a stuck collective must not burn the whole job: dump the Python stacks and exit after NGSAMG_BENCH_WATCHDOG_S seconds
    faulthandler.dump_traceback_later(int(os.environ.get("NGSAMG_BENCH_WATCHDOG_S", "1500")), exit=True)
    t_start = time.time()
    # stdout carries ONE JSON line: whatever native libraries print to fd 1 meanwhile (NCCL's version banner) is sent to stderr
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)

    def progress(msg):
        if os.environ.get("NGSAMG_BENCH_VERBOSE"):
            import resource
            print("[bench r%s %.1fs, peak rss %.1f GB] %s" % (os.environ.get("RANK", "0"), time.time() - t_start,
                                                             resource.getrusage(resource.RUSAGE_SELF).ru_maxrss / 1048576.0, msg), file=sys.stderr, flush=True)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        import torch.distributed as dist
        # torch.distributed is plumbing only: gloo carries the setup-phase host callbacks and the unique-id broadcast, the
        # library opens its own NCCL communicator for the halo exchange / dot products / coarse gather
        dist.init_process_group("cpu:gloo,cuda:nccl")

    def barrier():
        if world > 1:
            dist.all_reduce(torch.zeros(1, device="cuda"))   # NCCL all-reduce as the barrier (the group mixes gloo and nccl)
        torch.cuda.synchronize()

    n = rank_workload(args.n, rank, world)
    t0 = time.time()
    comm = None
    if world > 1:
        if args.problem not in ("poisson", "elasticity_jump"):
            raise SystemExit("bench.py: the multi-GPU bench runs the Poisson workload (configs[3]) or elasticity_jump (configs[4])")
        from ngsamg_b200 import parallel as par
        from ngsamg_b200 import synthetic as S
        grid = S.bench_grid(world)
        if args.problem == "elasticity_jump":
            # BASELINE.json configs[4]: n^3 vertices per GPU (--size 128 on 8 GPUs: 255^3 vertices = 49.7 M DOFs), modulus jumping by 1e4
            p = S.box_elasticity3d_jump(n, grid, rank)
            A = ng.SparseMatrix(p["n"], p["n"], 3, 3, p["rowptr"], p["col"], p["val"])
        else:
            p = S.box_poisson3d(n, grid, rank)
            A = ng.SparseMatrix(p["n"], p["n"], 1, 1, p["rowptr"], p["col"], p["val"])
        comm = par.TorchDistComm(use_nccl=os.environ.get("NGSAMG_BENCH_TRANSPORT", "nccl") == "nccl", device=local_rank)
    else:
        p, A = make_problem(n, args.problem)
    gen_s = time.time() - t0
    progress("problem generated")
    elast = args.problem in ("elasticity", "elasticity_p2", "elasticity_jump")
    if elast:
        args.no_cpu_baseline = True
        args.no_multicolor = True
    t0 = time.time()
    extra = {}
    for kv in os.environ.get("NGSAMG_FLAGS", "").split(","):
        if "=" in kv:
            k, v = kv.split("=", 1)
            extra["ngs_amg_" + k.strip()] = v.strip()
    tol = 1e-6 if elast else TOL       # the reference's elasticity tests solve to 1e-6 (tests/elasticity/amg_utils.py:439)
    if world > 1:
        args.no_multicolor = True
        if elast:
            pc = par.elast_3d_par(A, par.Halo(p["peers"], p["ex"]), comm, p["free"], vertex_xyz=p["xyz"], device=local_rank, defer_finalize=True, **extra)
        else:
            pc = par.h1_scal_par(A, par.Halo(p["peers"], p["ex"]), comm, p["free"], device=local_rank, defer_finalize=True, **extra)
        # the library holds its own copy now: release the generator's matrix arrays before the (memory-hungry) host setup
        empty_i, empty_d = np.zeros(0, np.int32), np.zeros(0)
        p["col"], p["val"], A.col, A.val = empty_i, empty_d, empty_i, empty_d
        pc.FinalizeLevel()
    elif elast:
        pc = ng.elast_3d(A, p["free"], vertex_xyz=p["xyz"], device=local_rank, **extra)
    else:
        pc = ng.h1_scal(A, p["free"], device=local_rank, **extra)
    setup_s = time.time() - t0
    progress("hierarchy built")
    ndof = p["n"] * A.bh
    rhs_h = np.ascontiguousarray(p["rhs"])
    x_h = np.zeros(ndof)
    rhs_d = torch.from_numpy(rhs_h).cuda()
    x_d = torch.zeros_like(rhs_d)
    cg = ng.CGSolver(mat=A, pre=pc, maxsteps=200, tol=tol)

    # ---- device-resident solve (value) ---------------------------------------------------------------
    for _ in range(args.warmup):
        cg.Solve(rhs_d, x_d)
        progress("warm-up solve done (%d iterations)" % cg.iterations)
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    l0 = pc.LaunchCount()
    wall0 = time.time()
    dev_ms = 0.0
    for _ in range(args.steps):
        cg.Solve(rhs_d, x_d)
        dev_ms += pc.LastMs("pcg")        # CUDA events on the library stream around the whole solve
    barrier()
    wall = time.time() - wall0
    launches = pc.LaunchCount() - l0
    iters = cg.iterations
    tmax = torch.tensor([dev_ms / 1e3, wall], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    dev_s, wall_s = tmax[0].item(), tmax[1].item()
    solve_s = dev_s / args.steps

    # ---- end-to-end through the C ABI with host buffers (e2e) ----------------------------------------
    progress("timed solves done")
    cg.Solve(rhs_h, x_h)
    barrier()
    e0 = time.time()
    for _ in range(args.steps):
        cg.Solve(rhs_h, x_h)
    barrier()
    e2e = torch.tensor([time.time() - e0], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(e2e, op=dist.ReduceOp.MAX)
    e2e_s = e2e.item() / args.steps

    # ---- V-cycle alone + per-kernel roofline (level 0) -------------------------------------------------
    progress("e2e solves done")
    for _ in range(3):
        pc.Mult(rhs_d, x_d)
    vms = []
    for _ in range(20):
        pc.Mult(rhs_d, x_d)
        vms.append(pc.LastMs("apply"))
    clocks = sampler.stop()
    vcycle_ms = float(np.mean(vms))
    # where the cycle spends its device time: one eager V-cycle with events around every phase (max over ranks per phase)
    pc.MultPhases(rhs_d, x_d)
    ph = pc.MultPhases(rhs_d, x_d)
    pht = torch.tensor([ph[k] for k in pc.PHASES], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(pht, op=dist.ReduceOp.MAX)
    phases = {k: round(float(v), 4) for k, v in zip(pc.PHASES, pht.tolist())}
    phases["note"] = "eager launch (no CUDA graph): launch gaps are exposed, compare shares"
    vbytes = pc.VCycleBytes()
    ndof_global = ndof
    if world > 1:
        # whole-job figures: bytes of all ranks, V-cycle time = max over ranks, global DOF count = master DOFs
        agg = torch.tensor([vbytes, float(p["n_master"]) * A.bh], dtype=torch.float64, device="cuda")   # master vertices x DOFs per vertex
        dist.all_reduce(agg)
        vt = torch.tensor([vcycle_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(vt, op=dist.ReduceOp.MAX)
        vbytes, ndof_global, vcycle_ms = agg[0].item(), int(agg[1].item()), vt.item()
    peak_scale = world
    peak, peak_src = measured_peak()
    kern = {}
    for name in ("gs_tri_fwd", "gs_upass", "gs_lpass", "gs_tri_bwd", "spmv", "restrict", "prolong"):
        ms, by = pc.ProfileKernel(name, level=0, reps=10)
        kern[name] = {"ms": ms, "gbs": by / ms / 1e6, "bytes": by}
    by_level = []
    for l in range(pc.GetNLevels() - 1):
        row = {}
        for name in ("gs_tri_fwd", "gs_upass", "restrict", "prolong", "gs_lpass", "gs_tri_bwd"):
            ms, by = pc.ProfileKernel(name, level=l, reps=5)
            row[name] = round(ms, 4)
        by_level.append(row)
    dom = max(("gs_tri_fwd", "gs_tri_bwd", "gs_upass", "gs_lpass"), key=lambda k: kern[k]["ms"])
    traffic = None
    sweep_kind = pc.SweepKind(0)
    tiled = sweep_kind in ("warp_tiles", "cta_tiles", "tile_images")
    sweep_kernel = {"rows": "k_gs_tri", "rows_rm": "k_gs_tri_rm", "warp_tiles": "k_gs_tile", "cta_tiles": "k_gs_ctile", "tile_images": "k_gs_itile"}[sweep_kind]
    try:
        # dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu --set full capture of the same kernel and size
        with open(os.path.join(ROOT, "profiles", "r02_traffic.json" if tiled else "r01_traffic.json")) as f:
            tj = json.load(f)
        if tj.get("n") == n:
            traffic = tj.get(dom)
    except Exception:
        pass
    roof = {"bound": "hbm", "kernel": "%s (%s, level 0)" % (sweep_kernel, dom), "achieved": kern[dom]["gbs"], "peak": peak,
            "unit": "GB/s", "frac": kern[dom]["gbs"] / peak, "traffic": traffic, "peak_source": peak_src,
            "ms_per_launch": kern[dom]["ms"], "algorithmic_bytes": kern[dom]["bytes"]}

    levels = []
    for l in range(pc.GetNLevels()):
        i = pc.level_info(l)
        levels.append({"n": int(i.n), "b": int(i.b), "nnz": int(i.nnz), "gs_depth": int(i.gs_depth)})
    par_info = None
    if world > 1:
        npar = pc.GetNParallelLevels()
        for l in range(len(levels)):
            levels[l]["distributed"] = l < npar
            levels[l]["rank0_local"] = True
        nested = pc.GetContracted()
        if nested is not None:
            for l in range(nested.GetNLevels()):
                i = nested.level_info(l)
                levels.append({"n": int(i.n), "b": int(i.b), "nnz": int(i.nnz), "gs_depth": int(i.gs_depth), "contracted_on_rank0": True})
        # one DIS2CO + one CO2CU of every distributed level, timed back to back (collective; no skew from the sweeps in between)
        ex_ms = []
        for l in range(npar):
            ms, _ = pc.ProfileKernel("halo_exchange", level=l, reps=20)
            ex_ms.append(round(ms, 4))
        par_info = {"distributed_levels": npar, "transport": {"peer_memory": "NVLink peer memory (IPC-mapped receive buffers, push/pull kernels); NCCL for the coarse gather and the dot products",
                                                              "nccl": "nccl p2p", "host": "host-staged (gloo callbacks)"}[pc.HaloTransport(0)],
                    "halo_exchange_pair_ms_by_level": ex_ms,
                    "host_exchanges_setup": comm.n_exchange, "box_grid": list(grid), "neighbours_rank0": len(p["peers"]),
                    "shared_dofs_rank0": int(sum(len(e) for e in p["ex"])), "global_dims": list(p["global_dims"])}

    # ---- optional variant, reported separately: multicolour Gauss-Seidel on the fine level ------------------------------
    variant = None
    if world == 1 and not args.no_multicolor and "ngs_amg_b200_sm_order" not in extra:
        del cg
        t0 = time.time()
        pcm = ng.h1_scal(A, p["free"], device=local_rank, ngs_amg_b200_sm_order="multicolor", **extra)
        cgm = ng.CGSolver(mat=A, pre=pcm, maxsteps=200, tol=TOL)
        setup_m = time.time() - t0
        for _ in range(2):
            cgm.Solve(rhs_d, x_d)
        ms = 0.0
        for _ in range(args.steps):
            cgm.Solve(rhs_d, x_d)
            ms += pcm.LastMs("pcg")
        vm = []
        for _ in range(10):
            pcm.Mult(rhs_d, x_d)
            vm.append(pcm.LastMs("apply"))
        vb = pcm.VCycleBytes()
        variant = {"smoother": "multicolour Gauss-Seidel on the fine level (ngs_amg_b200_sm_order=multicolor); NOT the reference's "
                               "natural-order sweep", "solve_s": ms / 1e3 / args.steps, "iterations": cgm.iterations,
                   "dofs_per_s": ndof / (ms / 1e3 / args.steps), "vcycle_ms": float(np.mean(vm)), "vcycle_gbs": vb / np.mean(vm) / 1e6,
                   "vcycle_frac_of_peak": vb / np.mean(vm) / 1e6 / peak, "gs_depth_level0": int(pcm.level_info(0).gs_depth),
                   "setup_s": setup_m}
        del pcm, cgm

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        r = cpu_reference_run(args.cpu_n, 1, 0)
        cpu = {"value": r["ndof"] / r["solve_s"], "unit": "DOF/s", "cores": 1, "kind": r["kind"],
               "sample": "PCG+AMG solve of the same problem at %d^3 = %d DOFs by %s (1 thread, %d its, %.2f s; V-cycle %.1f ms)"
                         % (args.cpu_n, r["ndof"], CPU_KIND_TEXT[r["kind"]], r["iterations"], r["solve_s"], r["vcycle_s"] * 1e3)}

    if rank == 0:
        line = {
            "metric": "pcg_amg_solve_dofs_per_s", "value": ndof_global / solve_s, "unit": "DOF/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": solve_s * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": ("3D linear elasticity P1 (Kuhn tets), Young's modulus jumping by 1e4 on a checkerboard of 8^3-cell boxes (BASELINE configs[4]), ONE global problem of "
                                    "%d x %d x %d vertices = %d DOFs cut into %d sub-boxes (%dx%dx%d) of %d^3 vertices (one per GPU, interface DOFs shared), elast_3d + CG to 1e-6, "
                                    "hybrid Gauss-Seidel + NCCL halo exchange" % (tuple(p["global_dims"]) + (ndof_global, world) + tuple(grid) + (n,))) if (world > 1 and elast) else
                                   ("3D Poisson P1 (Kuhn tets), ONE global problem of %d x %d x %d = %d DOFs cut into %d sub-boxes (%dx%dx%d) of %d^3 vertices (one per GPU, "
                                    "interface DOFs shared), h1_scal + CG to 1e-8, hybrid Gauss-Seidel + NCCL halo exchange" % (tuple(p["global_dims"]) + (ndof_global, world) + tuple(grid) + (n,))) if world > 1 else
                                   ("3D Poisson P1 unit cube (Kuhn tets), %d^3 = %d DOFs per GPU, h1_scal + CG to 1e-8" % (n, ndof)) if not elast else
                                   ("3D linear elasticity %s beam (Kuhn tets), %d nodes = %d DOFs per GPU, elast_3d (3x3 fine / 6x6 coarse blocks) + CG to 1e-6"
                                    % ({"elasticity_p2": "nodal-P2 (BASELINE configs[2]; AMG on all P2 nodes)", "elasticity_jump": "P1, Young's modulus jumping by 1e4 on a checkerboard of 8^3-cell boxes (BASELINE configs[4] workload, one GPU),"}.get(args.problem, "P1"), p["n"], ndof)),
                       "tol": tol, "levels": levels, "operator_complexity": pc.GetOC()[0],
                       "parallelism": "1 GPU" if world == 1 else "%d subdomains, one per GPU; DIS2CO/CO2CU halo exchange per sweep, all-reduced CG dot products, coarse levels contracted onto rank 0" % world,
                       "multi_gpu": par_info,
                       "l2": "inputs larger than L2 (level-0 matrix %.1f GB)" % (levels[0]["nnz"] * (8 * levels[0]["b"] ** 2 + 4) / 1e9)},
            "solve_s": solve_s, "iterations": iters, "setup_s": setup_s, "setup_rap_ms": pc.LastMs("rap"),
            "setup_host_ms": pc.LastMs("host"), "gen_s": gen_s,
            # Galerkin products as measured device work: transpose (counting/radix sort) + (P^T A) + (P^T A) P of every level, CUDA events;
            # compulsory bytes = M_f + 2 P + M_c per level -- reported, not gated
            "rap": {"ms": pc.LastMs("rap"), "bytes_compulsory": pc.LastMs("rap_bytes"),
                    "gbs": pc.LastMs("rap_bytes") / max(pc.LastMs("rap"), 1e-9) / 1e6,
                    "frac": pc.LastMs("rap_bytes") / max(pc.LastMs("rap"), 1e-9) / 1e6 / peak}, "wall_s_timed_region": wall_s,
            "vcycle_ms": vcycle_ms, "vcycle_bytes": vbytes, "vcycle_gbs": vbytes / vcycle_ms / 1e6,
            "vcycle_frac_of_peak": vbytes / vcycle_ms / 1e6 / (peak * peak_scale),
            "vcycle_phases_ms": phases, "kernels_level0": kern, "kernel_ms_by_level": by_level, "roofline": roof, "variant_multicolor": variant, "cpu_baseline": cpu,
            "e2e": {"value": ndof_global / e2e_s, "unit": "DOF/s", "h2d_bytes_per_step": 8 * ndof, "d2h_bytes_per_step": 8 * ndof,
                    "solve_s": e2e_s},
            "gpu_launches": int(launches), "clocks": clocks, "flags": extra,
        }
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        # the captured V-cycle graph pins the NCCL communicator: destroy the hierarchy first, then the communicator
        nested = None
        pc.close()
        del cg, pc
        barrier()
        comm.close()
        dist.destroy_process_group()
    faulthandler.cancel_dump_traceback_later()


if __name__ == "__main__":
    main()
