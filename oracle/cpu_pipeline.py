"""CPU-only multi-rank pipeline: "the reference's MPI CPU solve" restated end to end for R ranks on R host threads.

TEST / BASELINE INFRASTRUCTURE ONLY (tests/, bench.py --impl reference and its cpu_baseline leg).

  hierarchy : the product's HOST-side class-respecting coarsening per rank (ngsamg_b200.parallel.coarsen_par, no device), Galerkin
              products by the C oracle (RestrictMatrix restatement), contraction maps of the coarsest distributed level, serial
              hierarchy below it (ngsamg_b200.coarsen + oracle RAP)
  solve     : oracle_par.OracleParAMG (hybrid smoothers, DCC exchange, CtrMap, all-reduced CG)
The ranks run as threads (ThreadComm); the sequential Gauss-Seidel sweeps of the ranks run concurrently in the C oracle.
"""
import numpy as np

import ngsamg_b200 as ng
from ngsamg_b200 import parallel as par

from . import oracle as O
from . import oracle_par as OP


def _to_o(M):
    return O.Bsr(M.nrows, M.ncols, M.bh, M.bw, M.rowptr, M.col, M.val)


def _to_p(M):
    return ng.SparseMatrix(M.nrows, M.ncols, M.bh, M.bw, M.rowptr, M.col, M.val)


def contraction_maps(ns, peers, ex):
    """merged numbering of a distributed level: masters of rank 0, masters of rank 1, ...; ghosts map to their master's dof"""
    R = len(ns)
    mrank = [np.full(ns[r], r, np.int64) for r in range(R)]
    mindex = [np.arange(ns[r], dtype=np.int64) for r in range(R)]
    # lowest sharer wins; process neighbours in DEscending order so that the lowest one is written last
    for r in range(R):
        for kp in sorted(range(len(peers[r])), key=lambda k: -peers[r][k]):
            q = peers[r][kp]
            if q >= r:
                continue
            kq = list(peers[q]).index(r)
            e, eo = np.asarray(ex[r][kp], np.int64), np.asarray(ex[q][kq], np.int64)
            mrank[r][e] = q
            mindex[r][e] = eo
    # a dof shared by >2 ranks: the neighbour's copy may itself be a ghost of a still lower rank -> follow the chain once more
    for _ in range(R):
        for r in range(R):
            for d in np.flatnonzero(mrank[r] != r):
                m, i = mrank[r][d], mindex[r][d]
                if mrank[m][i] != m:
                    mrank[r][d], mindex[r][d] = mrank[m][i], mindex[m][i]
    offs, ordinal = [0], []
    for r in range(R):
        own = mrank[r] == r
        o = np.full(ns[r], -1, np.int64)
        o[own] = np.arange(int(own.sum()))
        ordinal.append(o)
        offs.append(offs[-1] + int(own.sum()))
    return [np.array([offs[mrank[r][d]] + ordinal[mrank[r][d]][mindex[r][d]] for d in range(ns[r])], np.int64) for r in range(R)]


def build(parts, b=1, elast=False, ctr_nv=2000, max_coarse=50, max_levels=10, pinv=False, engine="oracle"):
    """parts[r]: dict(n, rowptr, col, val, free, peers, ex[, xyz]).  Returns (OracleParAMG, info); engine="reference" returns the same
    hierarchy inside the reference library (RefParAMG), engine="both" returns the pair."""
    R = len(parts)
    A = [ng.SparseMatrix(p["n"], p["n"], b, b, p["rowptr"], p["col"], p["val"]) for p in parts]
    free = [p["free"] for p in parts]
    xyz = [p.get("xyz") if elast else None for p in parts]
    peers = [list(p["peers"]) for p in parts]
    ex = [[np.asarray(e, np.int32) for e in p["ex"]] for p in parts]
    A0 = [_to_o(a) for a in A]
    prols, halos = [], [(peers, ex)]
    cur, cfree, cxyz = A, free, xyz
    nglob = None
    for lvl in range(max_levels - 1):
        masters = 0
        for r in range(R):
            m, _, _, _ = OP.dcc_lists(r, cur[r].nrows, halos[-1][0][r], halos[-1][1][r])
            masters += int(m.sum())
        nglob = masters
        if lvl > 0 and masters <= ctr_nv:
            break
        bc = cur[0].bh
        if elast and lvl == 0 and bc == 3:
            bc = 6

        def step(r, comm):
            return par.coarsen_par(cur[r], par.Halo(halos[-1][0][r], halos[-1][1][r]), comm, cfree[r], cxyz[r], bcoarse=bc,
                                   max_per_row=(4 if elast else 3))

        res = par.run_ranks(R, step)
        Pl = [_to_o(res[r][0]) for r in range(R)]
        prols.append(Pl)
        nxt = [_to_p(O.restrict_matrix(O.transpose(Pl[r]), _to_o(cur[r]), Pl[r])) for r in range(R)]
        halos.append(([list(res[r][3].peers) for r in range(R)], [[np.asarray(e) for e in res[r][3].ex] for r in range(R)]))
        cur, cfree, cxyz = nxt, [None] * R, [res[r][2] for r in range(R)]
    maps = contraction_maps([c.nrows for c in cur], halos[-1][0], halos[-1][1])
    # merged level + serial hierarchy below it
    amg_probe = OP.OracleParAMG(A0, free, peers, ex, prols, halos, maps, [], pinv=pinv, nested_free=None) if False else None
    import scipy.sparse as sp
    bb = cur[0].bh
    N = int(max(int(m.max()) for m in maps if len(m)) + 1)
    mxyz = None
    for r in range(R):
        if elast and cxyz[r] is not None:
            if mxyz is None:
                mxyz = np.zeros((N, 3))
            mxyz[maps[r]] = cxyz[r]
    merged = OP.merge_contracted([_to_o(c) for c in cur], [np.asarray(m, np.int64) for m in maps], N, bb)
    nested, curm, cx = [], _to_p(merged), mxyz
    while curm.nrows > max_coarse and len(nested) + len(prols) + 2 < max_levels + 1:
        P, vmap, cxn = ng.coarsen(curm, None, cx, bcoarse=curm.bh, max_per_row=(4 if elast else 3))
        if P.ncols == 0 or P.ncols > 0.8 * curm.nrows:
            break
        nested.append(_to_o(P))
        curm = _to_p(O.restrict_matrix(O.transpose(nested[-1]), _to_o(curm), nested[-1]))
        cx = cxn
    args = (A0, free, peers, ex, prols, halos, maps, nested)
    info = dict(distributed_levels=len(prols), nested_levels=len(nested) + 1, n_contracted=N, n_global=nglob)
    if engine == "args":           # only the hierarchy: (A0, free, peers, ex, prols, halos, ctr_maps, nested_prols) for OracleParAMG / RefParAMG
        return args, info
    if engine == "reference":
        # the same hierarchy run by the reference's OWN smoothers / transfers / cycle (oracle/_ref, oracle/ref_pin/README.md)
        from oracle.ref_pin import ref as RP
        return RP.RefParAMG(*args, pinv=pinv), info
    amg = OP.OracleParAMG(*args, pinv=pinv)
    if engine == "both":
        from oracle.ref_pin import ref as RP
        return (amg, RP.RefParAMG(*args, pinv=pinv)), info
    return amg, info
