"""Generates tests/golden/refpin_*.npz: outputs of the REFERENCE's own functions on small seeded problems.

The producer is oracle/_ref/libngsamg_ref.so -- the bodies of TransposeSPMImpl, MatMultABImpl, RestrictMatrix, GSS3::*,
ProxySmoother, ProlMap transfers and AMGMatrix::SmoothV/W/BS cut out of /root/reference at build time and compiled verbatim
against a stand-in for the NGSolve containers (oracle/ref_pin/).  It can only be built where /root/reference exists, so the
outputs are frozen here; tests/test_ref_pin.py::test_oracle_against_reference_made_fixtures checks the oracle against them on
any machine.  The inputs (matrix, free mask, prolongations, vectors) are stored too, so the fixtures do not depend on the
product's coarsening staying the same.
Run from the repo root:  python tests/golden/make_ref_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from helpers import host_hierarchy, poisson, elasticity, rand, to_oracle  # noqa: E402
from oracle.ref_pin import ref as R  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
FLAGS = [(ru, ur, xz, bw) for ru in (0, 1) for ur in (0, 1) for xz in (0, 1) for bw in (0, 1)]


def pack(prefix, M, out):
    out[prefix + "_shape"] = np.array([M.nrows, M.ncols, M.bh, M.bw], np.int64)
    out[prefix + "_rowptr"], out[prefix + "_col"], out[prefix + "_val"] = M.rowptr, M.col, M.val


def make(name, p, A, elast, sm_steps, sm_symm, **copt):
    prols = [to_oracle(P) for P in host_hierarchy(A, p["free"], p.get("xyz"), elast=elast, **copt)]
    Ao = to_oracle(A)
    ra = R.RefAMG(Ao, p["free"], prols, sm_steps=sm_steps, sm_symm=sm_symm)
    out = {"free": np.asarray(p["free"], np.uint8), "nlevels": ra.nlevels, "sm_steps": sm_steps, "sm_symm": int(sm_symm),
           "fragments": np.array(R.fragment_index())}
    pack("A0", Ao, out)
    for l, P in enumerate(prols):
        pack("P%d" % l, P, out)
        pack("PT%d" % l, ra.level_pt(l), out)           # TransposeSPMImpl
        pack("A%d" % (l + 1), ra.level_matrix(l + 1), out)  # RestrictMatrix
        out["dinv%d" % l] = ra.level_dinv(l)            # GSS3::CalcDiags
    nb = Ao.nrows * Ao.bh
    # GSS3::Smooth / SmoothBack on level 0 for every combination of the protocol flags (bare smoother, no proxy)
    x0, b0 = rand(11, nb), rand(12, nb)
    out["sm_x_in"], out["sm_b"] = x0, b0
    r_true = b0 - Ao.to_scipy() @ x0
    for ru, ur, xz, bw in FLAGS:
        x = np.zeros(nb) if xz else x0.copy()
        res = (b0.copy() if xz else r_true.copy()) if ru else rand(13, nb)
        ra.smooth(0, x, b0, res, ru, ur, xz, bw, bare=True)
        out["sm_x_%d%d%d%d" % (ru, ur, xz, bw)] = x
        out["sm_res_%d%d%d%d" % (ru, ur, xz, bw)] = res
    out["sm_res_in_true"], out["sm_res_in_junk"] = r_true, rand(13, nb)
    # the cycles
    b = rand(14, nb)
    out["b"] = b
    for cyc in ("V", "W", "BS"):
        out["x_" + cyc] = ra.apply(b, cyc)
        if cyc == "V":
            for l in range(ra.nlevels):
                for w in ("x", "rhs", "res"):
                    out["V_%s%d" % (w, l)] = ra.level_vec(w, l)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print("%s: levels %s, %d arrays" % (name, [ra.level_matrix(l).nrows for l in range(ra.nlevels)], len(out)))




def make_hybrid():
    """2 x 2 x 2 ranks (dofs shared by up to 8 ranks): the reference's hybrid split, modified diagonal and every smoother call"""
    from ngsamg_b200 import synthetic as S
    from oracle import oracle as O
    from oracle import oracle_par as OP
    parts = S.partition_poisson3d(6, 5, 7, grid=(2, 2, 2))
    Rn = len(parts)
    A = [O.Bsr(p["n"], p["n"], 1, 1, p["rowptr"], p["col"], p["val"]) for p in parts]
    args = (A, [p["free"] for p in parts], [p["peers"] for p in parts], [p["ex"] for p in parts])
    RL = R.RefHybridLevel(*args)
    HL = OP.HybridLevel(*args)          # only used to produce a consistent residual input (b - A x); outputs come from RL
    out = {"R": Rn, "fragments": np.array(R.fragment_index())}
    for r, p in enumerate(parts):
        out["n%d" % r] = p["n"]
        out["rowptr%d" % r], out["col%d" % r], out["val%d" % r] = A[r].rowptr, A[r].col, A[r].val
        out["free%d" % r] = np.asarray(p["free"], np.uint8)
        out["peers%d" % r] = np.asarray(p["peers"], np.int32)
        ptr = np.zeros(len(p["peers"]) + 1, np.int64)
        for k in range(len(p["peers"])):
            ptr[k + 1] = ptr[k] + len(p["ex"][k])
        out["exptr%d" % r] = ptr
        out["exdofs%d" % r] = np.concatenate([np.asarray(e, np.int32) for e in p["ex"]])
        pack("M%d" % r, RL.M(r), out)
        G = RL.G(r)
        pack("G%d" % r, G if G is not None else O.Bsr(p["n"], p["n"], 1, 1, np.zeros(p["n"] + 1, np.int64), np.zeros(0, np.int32), np.zeros(0)), out)
        _, master, dinv = RL.info(r)
        out["master%d" % r], out["dinv%d" % r] = master, dinv
    nglob = 1 + max(int(p["gidx"].max()) for p in parts)
    for k, (ru, ur, xz, bw) in enumerate(FLAGS):
        rng = np.random.default_rng(300 + k)
        xg = rng.standard_normal(nglob)
        x = [np.zeros(p["n"]) if xz else xg[p["gidx"]].copy() for p in parts]
        rhs = [rng.standard_normal(p["n"]) for p in parts]
        if ru:
            y = RL.mult(x)
            res = [rhs[r] - y[r] for r in range(Rn)]
        else:
            res = [rng.standard_normal(p["n"]) for p in parts]
        key = "%d%d%d%d" % (ru, ur, xz, bw)
        for r in range(Rn):
            out["in_x_%s_%d" % (key, r)], out["in_b_%s_%d" % (key, r)], out["in_res_%s_%d" % (key, r)] = x[r].copy(), rhs[r].copy(), res[r].copy()
        RL.smooth(x, rhs, res, ru, ur, xz, bw)
        for r in range(Rn):
            out["out_x_%s_%d" % (key, r)], out["out_res_%s_%d" % (key, r)] = x[r], res[r]
    np.savez_compressed(os.path.join(HERE, "refpin_hybrid_2x2x2.npz"), **out)
    print("refpin_hybrid_2x2x2: %d ranks, %s dofs, %d arrays" % (Rn, [p["n"] for p in parts], len(out)))


def make_dense():
    """pseudo-inverse (CalcPseudoInverseTryNormal), coarse regularisation (RegTM<0,6,6>) and rigid-body transport (CalcQ) on seeded blocks"""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from test_ref_pin import pinv_cases
    out = {"fragments": np.array(R.fragment_index())}
    k = 0
    for n in (2, 3, 6):
        for name, M, direct in pinv_cases(n):
            out["pinv_in_%d" % k], out["pinv_out_%d" % k], out["pinv_direct_%d" % k] = M, R.pinv_block(M), np.array(direct)
            k += 1
    out["pinv_count"] = k
    rng = np.random.default_rng(3)
    Q, _ = np.linalg.qr(rng.standard_normal((6, 6)))
    X = rng.standard_normal((6, 6))
    regs = [X @ X.T + 6 * np.eye(6), Q @ np.diag([0, 0, 0, 1.5, 2.0, 7.0]) @ Q.T, np.diag([3.0, 2.0, 5.0, 0, 0, 0]), np.zeros((6, 6))]
    for i, M in enumerate(regs):
        out["reg_in_%d" % i], out["reg_out_%d" % i] = M, R.regularize_block6(M)
    out["reg_count"] = len(regs)
    ts = rng.standard_normal((5, 3))
    out["calcq_t"] = ts
    out["calcq_q"] = np.stack([R.elast_calcq(t) for t in ts])
    np.savez_compressed(os.path.join(HERE, "refpin_dense.npz"), **out)
    print("refpin_dense: %d pinv cases, %d regularisation cases, %d transport blocks" % (k, len(regs), len(ts)))


def make_bgs():
    """block Gauss-Seidel: inputs and outputs of the reference's own SmoothWO (oracle/_ref/libngsamg_ref_bgs.so) for every flag combination,
    forward and reverse block order, scalar and 3x3 blocks"""
    from helpers import elasticity, poisson, rand
    from oracle.ref_pin import ref_bgs as RB
    out = {}
    for tag, (p, A), b in (("h1", poisson(5), 1), ("el", elasticity(4, 3, 3), 3)):
        n = p["n"]
        blk = np.where(np.asarray(p["free"]) > 0, np.arange(n) // 3, -1)
        ids = np.unique(blk[blk >= 0])
        remap = -np.ones(int(blk.max()) + 2, np.int64)
        remap[ids] = np.arange(len(ids))
        blk = np.where(blk >= 0, remap[blk], -1)
        As = A.to_scipy().tocsr()
        x0, rhs = rand(31, n * b), rand(32, n * b)
        out[tag + "_rowptr"], out[tag + "_col"], out[tag + "_val"], out[tag + "_b"] = A.rowptr, A.col, A.val, np.int64(b)
        out[tag + "_blk"], out[tag + "_x0"], out[tag + "_rhs"] = blk, x0, rhs
        for back in (0, 1):
            for ru in (0, 1):
                for ur in (0, 1):
                    r0 = rhs - As @ x0 if ru else rand(33, n * b)
                    x, r = x0.copy(), r0.copy()
                    RB.smooth_wo(A, blk, x, rhs, r, ru, ur, False, reverse=bool(back))
                    key = "%s_%d%d%d" % (tag, back, ru, ur)
                    out[key + "_r0"], out[key + "_x"], out[key + "_r"] = r0, x, r
    np.savez_compressed(os.path.join(HERE, "refpin_bgs.npz"), **out)
    print("refpin_bgs.npz written")


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "bgs":
        make_bgs()
        sys.exit(0)
    p, A = poisson(7)
    make("refpin_poisson_n7", p, A, False, 1, False, max_coarse=20)
    make("refpin_poisson_n7_symm2", p, A, False, 2, True, max_coarse=20)
    p, A = elasticity(5, 3, 3)
    make("refpin_elast_5x3x3", p, A, True, 1, False, max_coarse=4, max_per_row=4)
    make_hybrid()
    make_dense()
    make_bgs()
