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


if __name__ == "__main__":
    p, A = poisson(7)
    make("refpin_poisson_n7", p, A, False, 1, False, max_coarse=20)
    make("refpin_poisson_n7_symm2", p, A, False, 2, True, max_coarse=20)
    p, A = elasticity(5, 3, 3)
    make("refpin_elast_5x3x3", p, A, True, 1, False, max_coarse=4, max_per_row=4)
