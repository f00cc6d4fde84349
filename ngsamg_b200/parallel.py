"""Multi-rank (one rank == one GPU) front-end: communicators, the parallel preconditioner class and NCCL plumbing.

The reference runs one MPI rank per subdomain (NGSolve ParallelMatrix / ParallelDofs; hybrid smoothers and the DCC halo
exchange, src/base/linalg/dcc_map.cpp, src/base/smoothers/hybrid_base_smoother.cpp).  Here one process (or, in the 1-GPU
parity tests, one thread) per rank drives one handle of the C ABI:

  * `Halo`            == ParallelDofs: neighbour ranks + the shared local DOFs per neighbour (ascending, pairwise consistent)
  * `ThreadComm`      ranks are threads of one process (several ranks may share one GPU; device traffic is staged through host)
  * `TorchDistComm`   ranks are torch.distributed processes (gloo for the host callbacks; optional NCCL communicator created by
                      the library itself for the device data path -- torch only broadcasts the 128-byte unique id)
  * `ParallelPreconditioner` / `h1_scal_par` ... == the reference classes on a ParallelMatrix; Mult / pcg are collective.

Setup-phase host traffic goes through the two callbacks of `ngsamg_comm` (include/ngsamg_b200.h); an NGSolve adapter would
implement them with its MPI communicator.
"""
import ctypes as C
import queue
import threading

import numpy as np

from . import _lib
from . import Preconditioner, SparseMatrix, NgsAMGError, _flag_value  # noqa: F401

_EXCH = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int32, C.POINTER(C.c_int32), C.POINTER(C.c_void_p), C.POINTER(C.c_int64),
                    C.POINTER(C.c_void_p), C.POINTER(C.c_int64))
_ARED = C.CFUNCTYPE(C.c_int, C.c_void_p, C.POINTER(C.c_double), C.c_int32)


class CommStruct(C.Structure):
    _fields_ = [("rank", C.c_int32), ("size", C.c_int32), ("ctx", C.c_void_p), ("exchange", _EXCH), ("allreduce_sum", _ARED),
                ("nccl", C.c_void_p)]


class HaloStruct(C.Structure):
    _fields_ = [("npeers", C.c_int32), ("peers", C.c_void_p), ("ex_ptr", C.c_void_p), ("ex_dofs", C.c_void_p)]


class Halo:
    """ParallelDofs of the fine level: `peers` ascending neighbour ranks, `ex[k]` the local DOFs shared with peers[k]."""

    def __init__(self, peers, ex):
        order = np.argsort(np.asarray(peers, dtype=np.int64), kind="stable") if len(peers) else []
        self.peers = np.ascontiguousarray([peers[i] for i in order], dtype=np.int32)
        self.ex = [np.ascontiguousarray(ex[i], dtype=np.int32) for i in order]
        self.ex_ptr = np.zeros(len(self.peers) + 1, np.int64)
        for k, e in enumerate(self.ex):
            self.ex_ptr[k + 1] = self.ex_ptr[k] + len(e)
        self.ex_dofs = np.ascontiguousarray(np.concatenate(self.ex) if len(self.ex) else np.zeros(0), dtype=np.int32)

    def _abi(self):
        return HaloStruct(len(self.peers), self.peers.ctypes.data, self.ex_ptr.ctypes.data, self.ex_dofs.ctypes.data)


class BaseComm:
    """builds the ngsamg_comm callback table; subclasses implement _exchange(peers, payloads, recv_sizes) and _allreduce(array)"""

    def __init__(self, rank, size, nccl=None):
        self.rank, self.size, self.nccl = int(rank), int(size), nccl
        self.error = None
        self.n_exchange = 0
        self.bytes_sent = 0

        def exch(ctx, npeers, peers, sendbuf, sendbytes, recvbuf, recvbytes):
            try:
                pl = [int(peers[k]) for k in range(npeers)]
                payload = [C.string_at(sendbuf[k], sendbytes[k]) if sendbytes[k] else b"" for k in range(npeers)]
                sizes = [int(recvbytes[k]) for k in range(npeers)]
                self.n_exchange += 1
                self.bytes_sent += sum(len(p) for p in payload)
                got = self._exchange(pl, payload, sizes)
                for k in range(npeers):
                    if len(got[k]) != sizes[k]:
                        raise RuntimeError("rank %d: expected %d bytes from rank %d, got %d" % (self.rank, sizes[k], pl[k], len(got[k])))
                    if sizes[k]:
                        C.memmove(recvbuf[k], got[k], sizes[k])
                return 0
            except BaseException as e:  # noqa: BLE001 -- must not propagate through the C frame
                self.error = e
                return 1

        def ared(ctx, vals, n):
            try:
                a = np.ctypeslib.as_array(vals, shape=(n,))
                a[:] = self._allreduce(a.copy())
                return 0
            except BaseException as e:  # noqa: BLE001
                self.error = e
                return 1

        self._cb = (_EXCH(exch), _ARED(ared))   # keep the thunks alive
        self.struct = CommStruct(self.rank, self.size, None, self._cb[0], self._cb[1], C.c_void_p(nccl) if nccl else None)

    def barrier(self):
        self._allreduce(np.zeros(1))


class _ThreadWorld:
    def __init__(self, size, timeout):
        self.size, self.timeout = size, timeout
        self.q = {(s, d): queue.Queue() for s in range(size) for d in range(size) if s != d}
        self.barrier = threading.Barrier(size, timeout=timeout)
        self.slots = [None] * size


class ThreadComm(BaseComm):
    """ranks == threads of this process; messages travel through queues.  Create the set with ThreadComm.world(n)."""

    def __init__(self, world, rank):
        super().__init__(rank, world.size)
        self.w = world

    @staticmethod
    def world(size, timeout=300.0):
        w = _ThreadWorld(size, timeout)
        return [ThreadComm(w, r) for r in range(size)]

    def _exchange(self, peers, payload, sizes):
        for p, data in zip(peers, payload):
            self.w.q[(self.rank, p)].put(data)
        return [self.w.q[(p, self.rank)].get(timeout=self.w.timeout) for p in peers]

    def _allreduce(self, a):
        self.w.slots[self.rank] = a
        self.w.barrier.wait()
        out = np.zeros_like(a)
        for r in range(self.size):   # rank order: identical result on every rank
            out += self.w.slots[r]
        self.w.barrier.wait()
        return out


class TorchDistComm(BaseComm):
    """ranks == torch.distributed processes.  Host callbacks use CPU tensors (gloo); pass use_nccl=True on a GPU box to let the
    library create its own NCCL communicator for the device data path (the unique id is broadcast with torch.distributed)."""

    def __init__(self, group=None, use_nccl=False, device=0):
        import torch.distributed as dist
        self.dist, self.group = dist, group
        rank, size = dist.get_rank(group), dist.get_world_size(group)
        nccl = None
        if use_nccl and size > 1:
            nccl = nccl_comm_init(self._bcast_id(rank), rank, size, device)
        super().__init__(rank, size, nccl)

    def _bcast_id(self, rank):
        obj = [nccl_unique_id() if rank == 0 else None]
        self.dist.broadcast_object_list(obj, src=0, group=self.group)
        return obj[0]

    def _exchange(self, peers, payload, sizes):
        import torch
        dist = self.dist
        ops, recv = [], []
        keep = []
        for p, data, n in zip(peers, payload, sizes):
            if len(data):
                t = torch.frombuffer(bytearray(data), dtype=torch.uint8)
                keep.append(t)
                ops.append(dist.P2POp(dist.isend, t, p, self.group))
            r = torch.empty(n, dtype=torch.uint8)
            recv.append(r)
            if n:
                ops.append(dist.P2POp(dist.irecv, r, p, self.group))
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()
        return [r.numpy().tobytes() for r in recv]

    def _allreduce(self, a):
        import torch
        t = torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64).copy())
        self.dist.all_reduce(t, group=self.group)
        return t.numpy()

    def close(self):
        if self.nccl:
            _lib.check(_lib.lib().ngsamg_b200_nccl_comm_destroy(C.c_void_p(self.nccl)))
            self.nccl = None


def nccl_unique_id():
    buf = C.create_string_buffer(128)
    _lib.check(_lib.lib().ngsamg_b200_nccl_unique_id(buf))
    return buf.raw


def nccl_comm_init(uid, rank, size, device=0):
    out = C.c_void_p()
    _lib.check(_lib.lib().ngsamg_b200_nccl_comm_init(C.create_string_buffer(uid, 128), int(rank), int(size), int(device), C.byref(out)))
    return out.value


class _ContractedView(Preconditioner):
    """the serial hierarchy below the contracted level (rank 0): borrowed handle, never destroyed here"""

    def __init__(self, parent, handle):  # noqa: super().__init__ is not called on purpose -- nothing is created
        self._lib, self._h, self._parent = parent._lib, C.c_void_p(handle), parent
        self._type, self.mat, self.flags, self._finalized = parent._type, None, parent.flags, True

    def __del__(self):
        self._h = None


class ParallelPreconditioner(Preconditioner):
    """BaseAMGPC on a ParallelMatrix: `mat` is this rank's sub-assembled local matrix, `halo` its ParallelDofs.
    Mult(b, x): b DISTRIBUTED (local load vector), x CUMULATED.  All ranks must call collectively."""

    def __init__(self, mat, halo, comm, freedofs=None, vertex_xyz=None, device=0, defer_finalize=False, **kwargs):
        if not isinstance(mat, SparseMatrix):
            raise TypeError("mat must be an ngsamg_b200.SparseMatrix")
        L = _lib.lib()
        self._lib = L
        self._h = C.c_void_p()
        self.mat, self.halo, self.comm = mat, halo, comm
        self.flags = dict(kwargs)
        fm = None if freedofs is None else np.ascontiguousarray(freedofs, dtype=np.uint8)
        xyz = None if vertex_xyz is None else np.ascontiguousarray(vertex_xyz, dtype=np.float64)
        keys = [k.encode() for k in kwargs]
        vals = [_flag_value(v).encode() for v in kwargs.values()]
        karr = (C.c_char_p * max(len(keys), 1))(*keys)
        varr = (C.c_char_p * max(len(vals), 1))(*vals)
        abi, habi = mat._abi(), halo._abi()
        self._check(L.ngsamg_b200_create_parallel(self._type.encode(), C.byref(abi), _lib.ptr(fm), _lib.ptr(xyz), C.byref(habi),
                                                  C.byref(comm.struct), karr, varr, len(keys), int(device), C.byref(self._h)))
        self._finalized = False
        if not defer_finalize:   # the library copied the host arrays at create: a caller short of host memory may drop its own first
            self.FinalizeLevel()

    def FinalizeLevel(self, mat=None):
        if not self._finalized:
            self._check(self._lib.ngsamg_b200_finalize(self._h))
            self._finalized = True

    def _check(self, rc):
        if rc != 0:
            msg = self._lib.ngsamg_b200_last_error().decode()
            if self.comm.error is not None:
                msg += " [callback: %r]" % (self.comm.error,)
            raise NgsAMGError(msg)

    def close(self):
        """destroy the device hierarchy now (a captured V-cycle graph pins the NCCL communicator until it is gone)"""
        if self._h:
            self._lib.ngsamg_b200_destroy(self._h)
            self._h = None

    def Mult(self, b, x):
        self._check(self._lib.ngsamg_b200_apply(self._h, _lib.ptr(b), _lib.ptr(x)))

    def MultAdd(self, s, b, x):
        self._check(self._lib.ngsamg_b200_apply_add(self._h, float(s), _lib.ptr(b), _lib.ptr(x)))

    MultTrans = Mult
    MultTransAdd = MultAdd

    def _pcg(self, rhs, x, tol, maxsteps):
        it = C.c_int(0)
        errs = np.zeros(maxsteps + 2)
        self._check(self._lib.ngsamg_b200_pcg(self._h, _lib.ptr(rhs), _lib.ptr(x), float(tol), int(maxsteps), C.byref(it), _lib.ptr(errs)))
        return it.value, errs[: it.value + 1].copy()

    # -- multi-rank introspection ------------------------------------------------------------------
    def GetNParallelLevels(self):
        return int(self._lib.ngsamg_b200_num_parallel_levels(self._h))

    def HaloTransport(self, level=0):
        """'host' | 'nccl' | 'peer_memory' (kernels_p2p.cuh) -- how the DIS2CO / CO2CU exchanges of a distributed level travel"""
        return {0: "host", 1: "nccl", 2: "peer_memory", -1: "none"}[int(self._lib.ngsamg_b200_halo_transport(self._h, int(level)))]

    def GetHalo(self, level):
        npeers = C.c_int32()
        _lib.check(self._lib.ngsamg_b200_get_halo(self._h, int(level), C.byref(npeers), None, None, None))
        peers = np.zeros(max(npeers.value, 1), np.int32)
        ptr = np.zeros(npeers.value + 1, np.int64)
        _lib.check(self._lib.ngsamg_b200_get_halo(self._h, int(level), None, _lib.ptr(peers), _lib.ptr(ptr), None))
        dofs = np.zeros(max(int(ptr[-1]), 1), np.int32)
        _lib.check(self._lib.ngsamg_b200_get_halo(self._h, int(level), None, None, None, _lib.ptr(dofs)))
        return Halo(list(peers[:npeers.value]), [dofs[ptr[k]:ptr[k + 1]] for k in range(npeers.value)])

    def GetHybrid(self, level):
        """(M, G, mod_diag) of the hybrid split of a distributed level (local numbering)"""
        i = self.level_info(level)
        nnz = (C.c_int64 * 2)()
        _lib.check(self._lib.ngsamg_b200_get_hybrid(self._h, int(level), 2, nnz, None, None, None, None))
        out = []
        md = np.zeros(i.n * i.b * i.b)
        for which in (0, 1):
            rp = np.zeros(i.n + 1, np.int64)
            ci = np.zeros(max(nnz[which], 1), np.int32)
            v = np.zeros(max(nnz[which], 1) * i.b * i.b)
            one = C.c_int64()
            _lib.check(self._lib.ngsamg_b200_get_hybrid(self._h, int(level), which, C.byref(one), _lib.ptr(rp), _lib.ptr(ci), _lib.ptr(v),
                                                        _lib.ptr(md)))
            out.append(SparseMatrix(i.n, i.n, i.b, i.b, rp, ci[:nnz[which]], v[:nnz[which] * i.b * i.b]))
        return out[0], out[1], md

    def GetContracted(self):
        """rank 0: view of the serial hierarchy below the contracted level, else None"""
        self._lib.ngsamg_b200_get_contracted.restype = C.c_void_p
        self._lib.ngsamg_b200_get_contracted.argtypes = [C.c_void_p]
        h = self._lib.ngsamg_b200_get_contracted(self._h)
        return _ContractedView(self, h) if h else None

    def GetContractionMap(self, rank):
        n = C.c_int64()
        _lib.check(self._lib.ngsamg_b200_get_contraction_map(self._h, int(rank), C.byref(n), None))
        m = np.zeros(max(n.value, 1), np.int32)
        _lib.check(self._lib.ngsamg_b200_get_contraction_map(self._h, int(rank), None, _lib.ptr(m)))
        return m[:n.value]


def _make_par(name):
    return type(name + "_par", (ParallelPreconditioner,), {"_type": name})


h1_scal_par = _make_par("h1_scal")
h1_3d_par = _make_par("h1_3d")
elast_3d_par = _make_par("elast_3d")
elast_2d_par = _make_par("elast_2d")


def hybrid_host(mat, halo, comm, freedofs=None):
    """host-only hybrid split of one level (no device needed): returns dict(M, G, mod_diag, sweep_rank, master)"""
    L = _lib.lib()
    fm = None if freedofs is None else np.ascontiguousarray(freedofs, dtype=np.uint8)
    h, nm, ng = C.c_void_p(), C.c_int64(), C.c_int64()
    abi, habi = mat._abi(), halo._abi()
    rc = L.ngsamg_b200_hybrid_host_begin(C.byref(abi), _lib.ptr(fm), C.byref(habi), C.byref(comm.struct), C.byref(h), C.byref(nm), C.byref(ng))
    if rc:
        raise NgsAMGError(L.ngsamg_b200_last_error().decode() + (" [callback: %r]" % (comm.error,) if comm.error else ""))
    n, bs = mat.nrows, mat.bh * mat.bw
    mrp, grp = np.zeros(n + 1, np.int64), np.zeros(n + 1, np.int64)
    mci, gci = np.zeros(max(nm.value, 1), np.int32), np.zeros(max(ng.value, 1), np.int32)
    mv, gv = np.zeros(max(nm.value, 1) * bs), np.zeros(max(ng.value, 1) * bs)
    md, sw, ma = np.zeros(n * bs), np.zeros(n, np.int32), np.zeros(n, np.uint8)
    _lib.check(L.ngsamg_b200_hybrid_host_fetch(h, _lib.ptr(mrp), _lib.ptr(mci), _lib.ptr(mv), _lib.ptr(grp), _lib.ptr(gci), _lib.ptr(gv),
                                               _lib.ptr(md), _lib.ptr(sw), _lib.ptr(ma)))
    return dict(M=SparseMatrix(n, n, mat.bh, mat.bw, mrp, mci[:nm.value], mv[:nm.value * bs]),
                G=SparseMatrix(n, n, mat.bh, mat.bw, grp, gci[:ng.value], gv[:ng.value * bs]), mod_diag=md, sweep_rank=sw, master=ma)


def contract_host(mat, halo, comm, freedofs=None):
    """host-only contraction of one distributed level onto rank 0 (collective): rank 0 gets (merged SparseMatrix, [dof map per rank]),
    the other ranks (None, None)"""
    L = _lib.lib()
    fm = None if freedofs is None else np.ascontiguousarray(freedofs, dtype=np.uint8)
    h, n, nz, nmap = C.c_void_p(), C.c_int64(), C.c_int64(), C.c_int64()
    abi, habi = mat._abi(), halo._abi()
    rc = L.ngsamg_b200_contract_host_begin(C.byref(abi), _lib.ptr(fm), C.byref(habi), C.byref(comm.struct), C.byref(h), C.byref(n), C.byref(nz),
                                           C.byref(nmap))
    if rc:
        raise NgsAMGError(L.ngsamg_b200_last_error().decode() + (" [callback: %r]" % (comm.error,) if comm.error else ""))
    if comm.rank != 0:
        _lib.check(L.ngsamg_b200_contract_host_fetch(h, None, None, None, None, None))
        return None, None
    bs = mat.bh * mat.bw
    rp, ci, v = np.zeros(n.value + 1, np.int64), np.zeros(max(nz.value, 1), np.int32), np.zeros(max(nz.value, 1) * bs)
    mp, dm = np.zeros(comm.size + 1, np.int64), np.zeros(max(nmap.value, 1), np.int32)
    _lib.check(L.ngsamg_b200_contract_host_fetch(h, _lib.ptr(rp), _lib.ptr(ci), _lib.ptr(v), _lib.ptr(mp), _lib.ptr(dm)))
    maps = [dm[mp[r]:mp[r + 1]].astype(np.int64) for r in range(comm.size)]
    return SparseMatrix(n.value, n.value, mat.bh, mat.bw, rp, ci[:nz.value], v[:nz.value * bs]), maps


def coarsen_par(mat, halo, comm, freedofs=None, vertex_xyz=None, bcoarse=None, max_per_row=3, min_frac=0.08, omega=1.0, smooth=True,
                rounds=3):
    """host-only: one class-respecting coarsening step of a distributed level (collective).  Returns (P, vmap, coarse_xyz, coarse Halo)."""
    L = _lib.lib()
    bc = mat.bh if bcoarse is None else int(bcoarse)
    fm = None if freedofs is None else np.ascontiguousarray(freedofs, dtype=np.uint8)
    xyz = None if vertex_xyz is None else np.ascontiguousarray(vertex_xyz, dtype=np.float64)
    h, nc, nz, npc, nsh = C.c_void_p(), C.c_int64(), C.c_int64(), C.c_int32(), C.c_int64()
    abi, habi = mat._abi(), halo._abi()
    rc = L.ngsamg_b200_coarsen_parallel_begin(C.byref(abi), _lib.ptr(fm), _lib.ptr(xyz), C.byref(habi), C.byref(comm.struct), bc,
                                              int(max_per_row), float(min_frac), float(omega), int(bool(smooth)), int(rounds), C.byref(h),
                                              C.byref(nc), C.byref(nz), C.byref(npc), C.byref(nsh))
    if rc:
        raise NgsAMGError(L.ngsamg_b200_last_error().decode() + (" [callback: %r]" % (comm.error,) if comm.error else ""))
    rp = np.zeros(mat.nrows + 1, np.int64)
    ci = np.zeros(max(nz.value, 1), np.int32)
    v = np.zeros(max(nz.value, 1) * mat.bh * bc)
    vmap = np.zeros(mat.nrows, np.int32)
    cxyz = None if xyz is None else np.zeros((nc.value, 3))
    peers = np.zeros(max(npc.value, 1), np.int32)
    exp = np.zeros(npc.value + 1, np.int64)
    exd = np.zeros(max(nsh.value, 1), np.int32)
    _lib.check(L.ngsamg_b200_coarsen_parallel_fetch(h, _lib.ptr(rp), _lib.ptr(ci), _lib.ptr(v), _lib.ptr(vmap), _lib.ptr(cxyz), _lib.ptr(peers),
                                                    _lib.ptr(exp), _lib.ptr(exd)))
    P = SparseMatrix(mat.nrows, nc.value, mat.bh, bc, rp, ci[:nz.value], v[:nz.value * mat.bh * bc])
    return P, vmap, cxyz, Halo(list(peers[:npc.value]), [exd[exp[k]:exp[k + 1]] for k in range(npc.value)])


def run_ranks(nranks, fn, timeout=300.0):
    """run fn(rank, comm) on `nranks` threads (ThreadComm world); returns the list of results, re-raises the first failure"""
    comms = ThreadComm.world(nranks, timeout)
    res, err = [None] * nranks, [None] * nranks

    def work(r):
        try:
            res[r] = fn(r, comms[r])
        except BaseException as e:  # noqa: BLE001
            err[r] = e
            try:
                comms[r].w.barrier.abort()
            except Exception:
                pass
    th = [threading.Thread(target=work, args=(r,), daemon=True) for r in range(nranks)]
    for t in th:
        t.start()
    for t in th:
        t.join(timeout * 4)
    for e in err:
        if e is not None and not isinstance(e, (threading.BrokenBarrierError, queue.Empty)):
            raise e
    for e in err:
        if e is not None:
            raise e
    return res
