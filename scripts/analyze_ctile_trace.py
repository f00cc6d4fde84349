"""offline analysis of the per-tile trace of the CTA-per-tile Gauss-Seidel sweep (NGSAMG_B200_TRACE_FILE, ngsamg_b200_profile_kernel):
where a tile spends its time, how long tiles wait for their dependencies and what the tile-DAG critical path costs."""
import sys

import numpy as np


def load(fn):
    raw = np.fromfile(fn, dtype=np.uint8)
    nt, npred = np.frombuffer(raw[:16], dtype=np.int64)
    o = 16
    pp = np.frombuffer(raw[o:o + 8 * (nt + 1)], dtype=np.int64); o += 8 * (nt + 1)
    pl = np.frombuffer(raw[o:o + 4 * npred], dtype=np.int32); o += 4 * npred
    tr = np.frombuffer(raw[o:o + 128 * nt], dtype=np.uint64).reshape(nt, 16)
    return int(nt), pp, pl, tr


def main(fn, backward=False):
    nt, pp, pl, tr = load(fn)
    t = tr[:, :6].astype(np.int64)
    t0 = t[:, 0].min()
    t = (t - t0) / 1e3                                       # us
    smid = (tr[:, 6] & np.uint64(0xffff)).astype(np.int64)
    cta = ((tr[:, 6] >> np.uint64(16)) & np.uint64(0xffffff)).astype(np.int64)
    nlev = ((tr[:, 6] >> np.uint64(40)) & np.uint64(0xfff)).astype(np.int64)
    ns = (tr[:, 6] >> np.uint64(52)).astype(np.int64)
    t_first = (tr[:, 7].astype(np.int64) - t0) / 1e3
    span = t[:, 5].max()
    ph = {"hint wait": t[:, 1] - t[:, 0], "slab wait": t[:, 2] - t[:, 1], "gather": t[:, 3] - t[:, 2], "levels": t[:, 4] - t[:, 3],
          "tail": t[:, 5] - t[:, 4], "whole tile": t[:, 5] - t[:, 0]}
    print("%s: %d tiles, %d CTAs on %d SMs, span %.1f us" % (fn, nt, len(np.unique(cta)), len(np.unique(smid)), span))
    for k, v in ph.items():
        print("  %-10s mean %7.2f  median %7.2f  p90 %7.2f  max %8.2f us   (sum/CTA-time %.2f)" % (k, v.mean(), np.median(v), np.percentile(v, 90), v.max(),
                                                                                                 v.sum() / (span * len(np.unique(cta)))))
    ok = (nlev > 1) & (tr[:, 7] > 0)
    if ok.any():
      print("  first level incl. gather skew: mean %.2f us; later levels: %.3f us each" % ((t_first - t[:, 3])[ok].mean(), ((t[:, 4] - t_first)[ok] / (nlev[ok] - 1)).mean()))
    cyc = tr[:, 8:12].astype(np.float64)
    if cyc[:, 3].max() > 0:
      print("  warp 0 cycle accounting per tile: loop %.0f cyc (%.2f us at 1.965 GHz) = active blocks %.0f (%.1f visits, %.0f cyc each) + barriers %.0f (%.0f cyc per level)"
          % (cyc[:, 3].mean(), cyc[:, 3].mean() / 1965.0, cyc[:, 0].mean(), cyc[:, 1].mean(), cyc[:, 0].sum() / max(cyc[:, 1].sum(), 1), cyc[:, 2].mean(),
             cyc[:, 2].sum() / max(nlev.sum(), 1)))
    print("  per local level: %.3f us (levels / nlev, mean nlev %.1f, mean slices %.1f)" % ((ph["levels"] / np.maximum(nlev, 1)).mean(), nlev.mean(), ns.mean()))
    # tile DAG levels and the dependency slack: when did the last dependency finish vs when did the tile pass its hint wait / gather
    lvl = np.zeros(nt, np.int64)
    order = range(nt - 1, -1, -1) if backward else range(nt)
    last_dep_end = np.zeros(nt)
    for q in order:
        d = pl[pp[q]:pp[q + 1]]
        if len(d):
            lvl[q] = lvl[d].max() + 1
            last_dep_end[q] = t[d, 4].max()              # dependency values are stored during its level loop: done at t4
    depth = lvl.max() + 1
    has = pp[1:] > pp[:-1]
    react = t[has, 3] - last_dep_end[has]                # from "last dependency finished" to "gathered everything"
    print("  tile DAG depth %d -> %.2f us per tile level" % (depth, span / depth))
    print("  reaction (last dependency's levels done -> own gather done): mean %.2f median %.2f p90 %.2f us" % (react.mean(), np.median(react), np.percentile(react, 90)))
    idle = (t[has, 0] > last_dep_end[has]).mean()
    print("  tiles whose dependencies were all finished before the CTA even picked them up: %.1f %%" % (100 * idle))
    # critical path through the measured per-tile costs
    own = t[:, 5] - np.maximum(t[:, 2], 0)
    # per DAG level: first start / last end
    lv_end = np.zeros(depth)
    np.maximum.at(lv_end, lvl, t[:, 5])
    d_end = np.diff(lv_end)
    print("  advance of the level front: mean %.2f us/level, median %.2f" % (d_end.mean(), np.median(d_end)))
    busy = ph["whole tile"].sum() / (span * len(np.unique(cta)))
    print("  CTA occupancy by tiles: %.2f;  time inside gather+levels+tail: %.2f" % (busy, (ph["gather"] + ph["levels"] + ph["tail"]).sum() / (span * len(np.unique(cta)))))


if __name__ == "__main__":
    main(sys.argv[1], backward=sys.argv[1].endswith(".bwd"))
