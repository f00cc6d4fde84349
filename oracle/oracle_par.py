"""Multi-rank CPU oracle: the reference's MPI-parallel preconditioner apply restated for R simulated ranks in ONE process.

TEST INFRASTRUCTURE ONLY (like oracle.py): imported by tests/ and the cpu_baseline / --impl reference legs of bench.py.

PARITY PINNED for the hybrid smoother level (HybridLevel): the reference's own functions -- BasicDCCMap, DecomposeSparseMatrixHybrid,
MyAllReduceDofData, CalcHybridSmootherRDG*, GSS3/GSS4, HybridGSSmoother, HybridBaseSmoother -- are cut out of /root/reference at build
time and compiled against a threaded MPI stand-in (oracle/ref_pin/ -> oracle/_ref/libngsamg_ref.so); tests/test_ref_pin_par.py compares
M, G, master flags, DCC lists, the inverted modified diagonal and x / res of every smoother call bit for bit.  UNPINNED: CtrMap /
contraction, the parallel V-cycle driver and CG around it (restated from the sources; checked against the assembled global
operator in tests/test_parallel_host.py).

What is restated (reference paths relative to /root/reference):
  * BasicDCCMap::CalcDOFMasters            src/base/linalg/dcc_map.cpp:494-543   -> dcc_lists
  * DCCMap DIS2CO / CO2CU                  src/base/linalg/dcc_map.cpp:76-302    -> dis2co / co2cu
  * DecomposeSparseMatrixHybrid            src/base/linalg/hybrid_matrix.cpp:17-307 -> hybrid_split
  * CalcHybridSmootherRDGItGeneric         src/base/smoothers/hybrid_smoother_utils.hpp:11-143 -> mod_diag
  * HybridGSSmoother::Finalize / stages    src/base/smoothers/gssmoother.cpp:616-861 -> HybridLevel.stage_masks
  * HybridBaseSmoother::SmoothImplRES/RHS, CallStageKernelsImpl  src/base/smoothers/hybrid_base_smoother.cpp:294-574
  * AMGMatrix::SmoothV on parallel vectors src/base/solve/amg_matrix.cpp:160-307
  * CtrMap transfers + matrix contraction  src/base/coarsening/dof_contract.cpp:49-228, 557-727
  * CGSolver with all-reduced inner products (ngsolve.krylovspace.CGSolver on ParallelVectors)
The sequential Gauss-Seidel sweeps themselves run in the C oracle (GSS3 / GSS4 arithmetic, gssmoother.cpp:195-315, 406-570).
The hierarchy (local prolongations, coarse sharing lists, contraction maps) is an INPUT -- it is read back from the product,
exactly like the single-rank parity tests inject the product's prolongations into OracleAMG.
"""
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import scipy.sparse as sp

from . import oracle as O

_POOL = None


def set_threads(n):
    """run the per-rank stages of the simulated ranks on n host threads (the C oracle releases the GIL); 0/1 = sequential"""
    global _POOL
    _POOL = ThreadPoolExecutor(n) if n and n > 1 else None


def _each(R, fn):
    """fn(r) for every rank -- the ranks are independent between two exchanges, exactly like MPI ranks"""
    if _POOL is None:
        for r in range(R):
            fn(r)
    else:
        list(_POOL.map(fn, range(R)))


def _expand(A):
    """Bsr -> scipy csr on scalar dofs, keeping structural zeros out of the way (values only matter here)"""
    return A.to_scipy().tocsr()


def dcc_lists(rank, n, peers, ex):
    """master flags + m_ex / g_ex per neighbour.  master of a shared dof = lowest rank sharing it (dps[0] < rank -> ghost)."""
    sharers = [[] for _ in range(n)]
    for kp, p in enumerate(peers):
        for d in ex[kp]:
            sharers[d].append(p)
    master = np.ones(n, bool)
    master_rank = np.full(n, rank, np.int64)
    for d in range(n):
        if sharers[d] and min(sharers[d]) < rank:
            master[d] = False
            master_rank[d] = min(sharers[d])
    m_ex, g_ex = [], []
    for kp, p in enumerate(peers):
        e = np.asarray(ex[kp], np.int64)
        m_ex.append(e[master[e]] if len(e) else e)
        g_ex.append(e[master_rank[e] == p] if len(e) else e)
    nshare = np.array([len(s) for s in sharers])
    return master, m_ex, g_ex, nshare


class HybridLevel:
    """one distributed level: all ranks' local matrices + sharing lists; builds M, G, mod diag, stage masks, dinv per rank"""

    def __init__(self, A_loc, free, peers, ex, pinv=False, sm_steps=1, sm_symm=False):
        self.sm_steps, self.sm_symm = int(sm_steps), bool(sm_symm)
        self.R = len(A_loc)
        self.A = A_loc
        self.b = A_loc[0].bh
        self.n = [a.nrows for a in A_loc]
        self.free = [None if f is None else np.asarray(f, np.uint8) for f in free]
        self.peers, self.ex = peers, ex
        self.pinv = pinv
        R, b = self.R, self.b
        self.master, self.m_ex, self.g_ex, self.nshare = [], [], [], []
        for r in range(R):
            m, me, ge, ns = dcc_lists(r, self.n[r], peers[r], ex[r])
            self.master.append(m); self.m_ex.append(me); self.g_ex.append(ge); self.nshare.append(ns)
        self._split()
        self._mod_diag()
        self._smoothers()

    def _sdofs(self, dofs):
        d = np.asarray(dofs, np.int64)
        return (d[:, None] * self.b + np.arange(self.b)[None, :]).ravel()

    # ---- DecomposeSparseMatrixHybrid -------------------------------------------------------------------
    def _split(self):
        R, b = self.R, self.b
        S = [_expand(a) for a in self.A]
        self.S = S
        # every rank ships the diagonal block on the ghost dofs of neighbour kp (master = that neighbour) to it
        shipped = {}
        for r in range(R):
            for kp, p in enumerate(self.peers[r]):
                g = self.g_ex[r][kp]
                if len(g):
                    sd = self._sdofs(g)
                    shipped[(r, p)] = S[r][sd][:, sd].tocoo()
        self.M, self.G = [], []
        for r in range(R):
            n = self.n[r] * b
            msk = np.repeat(self.master[r], b)
            D = sp.diags(msk.astype(float))
            M = (D @ S[r] @ D).tocsr()       # master x master part of the own matrix
            for kp, p in enumerate(self.peers[r]):   # ascending neighbour order, like the reference's merge
                if (p, r) in shipped:
                    blk = shipped[(p, r)]
                    sd = self._sdofs(self.m_ex[r][kp])
                    M = M + sp.coo_matrix((blk.data, (sd[blk.row], sd[blk.col])), shape=(n, n)).tocsr()
            # G: entries whose row- and column-masters differ
            mr = np.full(self.n[r], -1, np.int64)
            for kp, p in enumerate(self.peers[r]):
                if len(self.g_ex[r][kp]):
                    mr[self.g_ex[r][kp]] = kp
            mrs = np.repeat(mr, b)
            C = S[r].tocoo()
            keep = mrs[C.row] != mrs[C.col]
            G = sp.coo_matrix((C.data[keep], (C.row[keep], C.col[keep])), shape=(n, n)).tocsr()
            self.M.append(M); self.G.append(G)

    # ---- CalcHybridSmootherRDGItGeneric ----------------------------------------------------------------
    def _allreduce_dofdata(self, data):
        """MyAllReduceDofData: every sharer ends up with the sum over all sharers (data[r]: n_r x k)"""
        out = [d.copy() for d in data]
        for r in range(self.R):
            for kp, p in enumerate(self.peers[r]):
                e = np.asarray(self.ex[r][kp], np.int64)
                kq = list(self.peers[p]).index(r)
                eo = np.asarray(self.ex[p][kq], np.int64)
                if len(e):
                    out[r][e] += data[p][eo]
        return out

    def _mod_diag(self):
        R, b = self.R, self.b
        od = []
        for r in range(R):
            d = np.zeros((self.n[r], b, b))
            Md = self.M[r]
            for i in range(b):
                for j in range(b):
                    d[:, i, j] = np.asarray(Md[np.arange(self.n[r]) * b + i, np.arange(self.n[r]) * b + j]).ravel()
            od.append(d.reshape(self.n[r], b * b))
        od = self._allreduce_dofdata(od)
        self.orig_diag = od
        ad = []
        for r in range(R):
            dd = od[r].reshape(self.n[r], b, b)
            sq = np.sqrt(np.abs(np.einsum("kii->ki", dd))).reshape(-1)
            G = self.G[r].tocoo()
            with np.errstate(divide="ignore", invalid="ignore"):
                w = np.abs(G.data) / (sq[G.row] * sq[G.col])
            a = np.bincount(G.row, weights=w, minlength=self.n[r] * b).astype(np.float64).reshape(self.n[r], b)
            if self.free[r] is not None:
                a[self.free[r] == 0] = 0.0
            ad.append(a)
        ad = self._allreduce_dofdata(ad)
        self.md = []
        for r in range(R):
            fac = np.maximum(1.0, 0.51 * (1.0 + ad[r])).max(axis=1)
            md = od[r] * fac[:, None]
            dead = ~self.master[r]
            if self.free[r] is not None:
                dead |= self.free[r] == 0
            md[dead] = 0.0
            self.md.append(md)

    # ---- HybridGSSmoother::Finalize: loc / ex subsets, split index, GSS3(M, mod_diag, loc), GSS4(M, mod_diag, ex) ----
    def _smoothers(self):
        R = self.R
        self.Mb, self.Gb, self.dinv, self.masks = [], [], [], []
        for r in range(R):
            n = self.n[r]
            loc = self.master[r] & (self.nshare[r] == 0)
            exm = self.master[r] & (self.nshare[r] > 0)
            if self.free[r] is not None:
                loc &= self.free[r] != 0
                exm &= self.free[r] != 0
                cnt, half, split = 0, int(loc.sum()) // 2, 0
                idx = np.flatnonzero(loc)
                if len(idx) > half:
                    split = int(idx[half])
            else:
                split = n // 2
            ar = np.arange(n)
            m1 = (loc & (ar < split)).astype(np.uint8)
            m2 = (loc & (ar >= split)).astype(np.uint8)
            Mb = O.Bsr.from_scipy(self.M[r], self.b, self.b) if self.M[r].nnz else O.Bsr(n, n, self.b, self.b, np.zeros(n + 1, np.int64), np.zeros(0, np.int32), np.zeros(0))
            if Mb.nrows != n:
                raise RuntimeError("block conversion changed the size")
            Gb = O.Bsr.from_scipy(self.G[r], self.b, self.b) if self.G[r].nnz else O.Bsr(n, n, self.b, self.b, np.zeros(n + 1, np.int64), np.zeros(0, np.int32), np.zeros(0))
            sm = (loc | exm).astype(np.uint8)
            dinv = O.calc_dinv(Mb, sm, self.pinv, self.md[r].reshape(-1))
            self.Mb.append(Mb); self.Gb.append(Gb); self.dinv.append(dinv)
            self.masks.append((m1, exm.astype(np.uint8), m2))

    # ---- DCCMap ----------------------------------------------------------------------------------------
    def dis2co_start(self, vec):
        """StartDIS2CO: BufferG (pack + zero the ghosts) and send to the master; returns the messages in flight"""
        b = self.b
        buf = {}
        for r in range(self.R):
            v = vec[r].reshape(-1, b)
            for kp, p in enumerate(self.peers[r]):
                g = self.g_ex[r][kp]
                if len(g):
                    buf[(r, p)] = v[g].copy()
                    v[g] = 0.0
        return buf

    def dis2co_finish(self, vec, buf):
        """FinishDIS2CO: ApplyM (the master adds what it received, neighbours ascending)"""
        b = self.b
        for r in range(self.R):
            v = vec[r].reshape(-1, b)
            for kp, p in enumerate(self.peers[r]):
                m = self.m_ex[r][kp]
                if len(m):
                    v[m] += buf[(p, r)]

    def dis2co(self, vec):
        """DISTRIBUTED -> CONCENTRATED in one go (DCCMap::StartDIS2CO, ApplyDIS2CO, FinishDIS2CO)"""
        self.dis2co_finish(vec, self.dis2co_start(vec))

    def co2cu(self, vec):
        """BufferM, send to the ghosts, ApplyG (ghost values are overwritten)"""
        b = self.b
        buf = {}
        for r in range(self.R):
            v = vec[r].reshape(-1, b)
            for kp, p in enumerate(self.peers[r]):
                m = self.m_ex[r][kp]
                if len(m):
                    buf[(r, p)] = v[m].copy()
        for r in range(self.R):
            v = vec[r].reshape(-1, b)
            for kp, p in enumerate(self.peers[r]):
                g = self.g_ex[r][kp]
                if len(g):
                    v[g] = buf[(p, r)]

    # ---- HybridBaseSmoother::SmoothImplRES / SmoothImplRHS with CallStageKernelsImpl --------------------
    def _stages(self, r, backward):
        m1, mex, m2 = self.masks[r]
        return (m2, mex, m1) if backward else (m1, mex, m2)

    def smooth_res(self, x, res, backward, x_zero):
        """x CUMULATED, res DISTRIBUTED (b - A x); afterwards x CUMULATED (new), res DISTRIBUTED"""
        R = self.R
        gx = None
        if not x_zero:
            gx = [O.spmv_add(self.Gb[r], 1.0, x[r], np.zeros_like(x[r])) for r in range(R)]
        # CallStageKernelsImpl (hybrid_base_smoother.cpp:498-574): StartDIS2CO | first local part | FinishDIS2CO (ApplyM: the received
        # values are added AFTER the first local part has already scattered into the master rows) | EX_PART | StartCO2CU | second
        # local part | FinishCO2CU.  With or without overlap the order of these additions is the same.
        inflight = self.dis2co_start(res)
        st = [self._stages(r, backward) for r in range(R)]
        _each(R, lambda r: O.gs_res(self.Mb[r], self.dinv[r], st[r][0], x[r], res[r], backward))
        self.dis2co_finish(res, inflight)

        def sweep(r):
            # EX_PART, then (after StartCO2CU, which only reads the exchange rows) the second local part
            for mask in st[r][1:]:
                O.gs_res(self.Mb[r], self.dinv[r], mask, x[r], res[r], backward)
        _each(R, sweep)
        self.co2cu(x)

        def fix(r):
            if gx is not None:
                res[r] += gx[r]
            O.spmv_add(self.Gb[r], -1.0, x[r], res[r])
        _each(R, fix)

    def smooth_rhs(self, x, rhs, backward, x_zero):
        """x CUMULATED, rhs DISTRIBUTED (not modified); afterwards x CUMULATED"""
        R = self.R
        t = [rhs[r].copy() for r in range(R)]
        if not x_zero:
            for r in range(R):
                O.spmv_add(self.Gb[r], -1.0, x[r], t[r])
        self.dis2co(t)

        def sweep(r):
            for mask in self._stages(r, backward):
                O.gs_rhs(self.Mb[r], self.dinv[r], mask, x[r], t[r], backward)
        _each(R, sweep)
        self.co2cu(x)

    def smooth(self, x, b, res, res_updated, update_res, x_zero, backward):
        """HybridBaseSmoother::SmoothImpl (hybrid_base_smoother.cpp:242-290): picks the RES or the RHS form by cost"""
        R = self.R
        if update_res:
            if not res_updated:
                if not x_zero:
                    self.smooth_rhs(x, b, backward, False)

                    def residual(r):        # res = b; res -= A x  ->  HybridBaseMatrix::MultAdd(-1): M first, then G
                        res[r][:] = b[r]
                        O.spmv_add(self.Mb[r], -1.0, x[r], res[r])
                        O.spmv_add(self.Gb[r], -1.0, x[r], res[r])
                    _each(R, residual)
                else:
                    for r in range(R):
                        res[r][:] = b[r]
                    self.smooth_res(x, res, backward, True)
            else:
                self.smooth_res(x, res, backward, x_zero)
        else:
            self.smooth_rhs(x, b, backward, x_zero)

    def level_smooth(self, x, b, res, ru, ur, xz, backward):
        """ProxySmoother around the hybrid smoother: SmoothK / SmoothBackK / SmoothSymmK (base_smoother.hpp:79-112, 169-229)"""
        k = max(1, self.sm_steps)
        if self.sm_symm:
            self.smooth(x, b, res, ru, ur, xz, False)
            self.smooth(x, b, res, ur, ur, False, True)
            for _ in range(k - 1):
                self.smooth(x, b, res, ur, ur, False, False)
                self.smooth(x, b, res, ur, ur, False, True)
        else:
            self.smooth(x, b, res, ru, ur, xz, backward)
            for _ in range(k - 1):
                self.smooth(x, b, res, ur, ur, False, backward)

    def mult(self, x):
        """HybridBaseMatrix::Mult: y = (M + G) x, x CUMULATED, y DISTRIBUTED"""
        out = [np.zeros_like(x[r]) for r in range(self.R)]

        def mv(r):
            O.spmv_add(self.Mb[r], 1.0, x[r], out[r])
            O.spmv_add(self.Gb[r], 1.0, x[r], out[r])
        _each(self.R, mv)
        return out


def merge_contracted(A, maps, N, b):
    """CtrMap::DoAssembleMatrix (dof_contract.cpp:557-727): the master remaps the members' rows through the dof maps, merges the
    (sorted) column lists -- a STRUCTURAL union, entries that cancel to zero stay in the pattern -- and sums coinciding entries,
    members in group order.  (scipy's csr + csr would drop the cancelled entries, hence the single COO pass.)"""
    rows, cols, vals = [], [], []
    for r in range(len(A)):
        C = A[r].to_scipy().tocoo() if A[r].nnz else None
        if C is None:
            continue
        # to_scipy keeps every stored scalar of every stored block, explicit zeros included
        rp, ci, bb = A[r].rowptr, A[r].col, A[r].bh
        br = np.repeat(np.arange(A[r].nrows), np.diff(rp))
        sd_r = (maps[r][br][:, None, None] * bb + np.arange(bb)[None, :, None])
        sd_c = (maps[r][ci][:, None, None] * bb + np.arange(bb)[None, None, :])
        rows.append(np.broadcast_to(sd_r, (len(ci), bb, bb)).ravel())
        cols.append(np.broadcast_to(sd_c, (len(ci), bb, bb)).ravel())
        vals.append(A[r].val.reshape(-1))
    if not rows:
        return O.Bsr(N, N, b, b, np.zeros(N + 1, np.int64), np.zeros(0, np.int32), np.zeros(0))
    nb = N * b
    keys = [rw.astype(np.int64) * nb + cl for rw, cl in zip(rows, cols)]
    ukeys = np.unique(np.concatenate(keys))                     # the merged pattern: sorted union, nothing dropped
    acc = np.zeros(len(ukeys))
    for k, v in zip(keys, vals):                                # rvs = 0; rvs[pos] += member values, members in group (= rank) order
        acc[np.searchsorted(ukeys, k)] += v                     # a member's dof map is injective: no duplicate positions inside one member
    m = sp.csr_matrix((acc, (ukeys // nb, ukeys % nb)), shape=(nb, nb))
    m.sort_indices()
    return O.Bsr.from_scipy(m, b, b)


class OracleParAMG:
    """AMGMatrix on a distributed hierarchy.
    levels[l] = dict(A=[Bsr per rank], free=[mask or None per rank], peers=[...], ex=[...], P=[Bsr per rank]) for l < npar
    ctr       = dict(maps=[local -> merged dof per rank]) for the contracted level npar
    nested    = dict(prols=[Bsr ...], sm kwargs) : the serial hierarchy on the merged level (OracleAMG)"""

    def __init__(self, A0, free0, peers0, ex0, prols, halos, ctr_maps, nested_prols, pinv=False, nested_free=None, sm_steps=1,
                 sm_symm=False):
        self.R = len(A0)
        self.npar = len(prols)
        self.levels = []
        A, free, peers, ex = A0, free0, peers0, ex0
        self.P, self.PT = [], []
        for l in range(self.npar):
            self.levels.append(HybridLevel(A, free, peers, ex, pinv, sm_steps, sm_symm))
            Pl = prols[l]
            PTl = [O.transpose(p) for p in Pl]
            self.P.append(Pl); self.PT.append(PTl)
            A = [O.restrict_matrix(PTl[r], A[r], Pl[r]) for r in range(self.R)]     # local Galerkin products (DISTRIBUTED sum)
            free = [None] * self.R
            peers, ex = halos[l + 1]
        self.A_ctr = A
        self.b_ctr = A[0].bh
        # CtrMap::DoAssembleMatrix: the master sums the members' local matrices through the dof maps
        self.maps = [np.asarray(m, np.int64) for m in ctr_maps]
        N = int(max(int(m.max()) for m in self.maps if len(m)) + 1)
        b = self.b_ctr
        self.A_merged = merge_contracted(A, self.maps, N, b)
        self.N = N
        self.nested = O.OracleAMG(self.A_merged, nested_free, nested_prols, pinv=pinv, sm_steps=sm_steps, sm_symm=sm_symm)

    def contracted_solve(self, rhs):
        b = self.b_ctr
        g = np.zeros(self.N * b)
        for r in range(self.R):                                       # CtrMap::TransferF2C: master adds the members' values
            np.add.at(g.reshape(-1, b), self.maps[r], rhs[r].reshape(-1, b))
        xg = self.nested.apply(g)
        return [xg.reshape(-1, b)[self.maps[r]].reshape(-1).copy() for r in range(self.R)]   # TransferC2F: CUMULATED values

    def apply(self, b0):
        """x = C b : b0[r] DISTRIBUTED local vectors -> CUMULATED x per rank; keeps the level vectors for inspection"""
        R = self.R
        rhs = [[np.ascontiguousarray(v, np.float64).copy() for v in b0]]
        xs, ress = [], []
        for l in range(self.npar):
            L = self.levels[l]
            x = [np.zeros_like(v) for v in rhs[l]]
            res = [v.copy() for v in rhs[l]]
            L.level_smooth(x, rhs[l], res, True, True, True, False)
            xs.append(x); ress.append(res)
            nxt = []
            for r in range(R):
                PT = self.PT[l][r]
                nxt.append(O.spmv_add(PT, 1.0, res[r], np.zeros(PT.nrows * PT.bh)))
            rhs.append(nxt)
        xc = self.contracted_solve(rhs[self.npar])
        self.level_x = [None] * self.npar + [xc]
        for l in range(self.npar - 1, -1, -1):
            L = self.levels[l]
            x = xs[l]
            for r in range(R):
                O.spmv_add(self.P[l][r], 1.0, xc[r], x[r])
            L.level_smooth(x, rhs[l], [np.zeros_like(v) for v in x], False, False, False, True)
            xc = x
            self.level_x[l] = x
        self.level_rhs, self.level_res = rhs, ress
        return xc

    def pcg(self, rhs, tol=1e-8, maxsteps=200):
        """CGSolver on parallel vectors: d DISTRIBUTED, w/s/u CUMULATED; inner products all-reduced"""
        R = self.R
        L0 = self.levels[0]
        dot = lambda a, b: float(sum(float(np.multiply(a[r], b[r]).sum()) for r in range(R)))   # local dots + all-reduce (no BLAS threads)
        d = [np.ascontiguousarray(v, np.float64).copy() for v in rhs]
        u = [np.zeros_like(v) for v in d]
        w = self.apply(d)
        s = [v.copy() for v in w]
        wdn = dot(w, d)
        err0 = np.sqrt(abs(wdn))
        errs = [err0]
        it = 0
        if wdn != 0.0:
            for it in range(1, maxsteps + 1):
                q = L0.mult(s)
                wd = wdn
                alpha = wd / dot(s, q)
                for r in range(R):
                    u[r] += alpha * s[r]
                    d[r] -= alpha * q[r]
                w = self.apply(d)
                wdn = dot(w, d)
                beta = wdn / wd
                for r in range(R):
                    s[r] = w[r] + beta * s[r]
                err = np.sqrt(abs(wd))
                errs.append(err)
                if err < tol * err0:
                    break
        return u, it, np.array(errs)
