"""Multi-GPU parity check of the NCCL transport (run under torch.distributed.run on N >= 2 GPUs, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/run_nccl_check.py

Every rank builds its part of ONE partitioned Poisson problem with the library's own NCCL communicator for the device data path
(halo exchange, dot products, coarse gather), gathers the hierarchy of all ranks, runs the multi-rank CPU oracle for the whole
problem and compares its own V-cycle / PCG results with it (<= 1e-10 relative, identical iteration counts).  Also run once with
CUDA-graph capture of the whole distributed V-cycle (ngs_amg_b200_cuda_graph_par=1)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    import torch.distributed as dist
    import ngsamg_b200 as ng
    from helpers import rand, rel, to_oracle
    from ngsamg_b200 import parallel as par, synthetic as S
    from oracle import oracle as O
    from oracle import oracle_par as OP

    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("cpu:gloo,cuda:nccl")
    rank, size = dist.get_rank(), dist.get_world_size()
    comm = par.TorchDistComm(use_nccl=True, device=local)
    assert comm.nccl, "the NCCL communicator was not created"
    dims = (21, 17, 8 * size + 1)
    parts = S.partition_poisson3d(*dims, grid=(1, 1, size))
    p = parts[rank]
    A = ng.SparseMatrix(p["n"], p["n"], 1, 1, p["rowptr"], p["col"], p["val"])
    ok = True
    # graph: the whole distributed V-cycle captured into a CUDA graph; p2p: halo exchange through NVLink peer memory
    # (kernels_p2p.cuh, IPC-mapped receive buffers) instead of ncclSend/ncclRecv
    # (the peer-memory variants run on request -- NGSAMG_CHECK_P2P=1 -- the transport is opt-in, `ngs_amg_b200_halo_p2p`)
    combos = ((0, 0), (1, 0), (0, 1), (1, 1)) if os.environ.get("NGSAMG_CHECK_P2P") else ((0, 0), (1, 0))
    for graph, p2p in combos:
        pc = par.h1_scal_par(A, par.Halo(p["peers"], p["ex"]), comm, p["free"], device=local, ngs_amg_max_coarse_size=15,
                             ngs_amg_b200_ctr_nv=400, ngs_amg_b200_cuda_graph_par=graph, ngs_amg_b200_halo_p2p=p2p)
        npar = pc.GetNParallelLevels()
        if not p2p:
            assert pc.HaloTransport(0) == "nccl", pc.HaloTransport(0)      # with p2p = 1 the library falls back to NCCL where IPC is unavailable
        mine = dict(prols=[pc.GetProlongation(l) for l in range(npar)], halos=[(list(pc.GetHalo(l).peers), [np.asarray(e) for e in pc.GetHalo(l).ex]) for l in range(npar + 1)])
        if rank == 0:
            mine["maps"] = [pc.GetContractionMap(r) for r in range(size)]
            mine["nested"] = pc.GetContracted().GetMap()
        allh = [None] * size
        dist.all_gather_object(allh, mine)
        prols = [[to_oracle(allh[r]["prols"][l]) for r in range(size)] for l in range(npar)]
        halos = [([allh[r]["halos"][l][0] for r in range(size)], [allh[r]["halos"][l][1] for r in range(size)]) for l in range(npar + 1)]
        amg = OP.OracleParAMG([O.Bsr(q["n"], q["n"], 1, 1, q["rowptr"], q["col"], q["val"]) for q in parts], [q["free"] for q in parts],
                              halos[0][0], halos[0][1], prols, halos, allh[0]["maps"], [to_oracle(P) for P in allh[0]["nested"]])
        b = [rand(40 + r, q["n"]) * q["free"] for r, q in enumerate(parts)]
        xo = amg.apply(b)
        x = np.zeros(p["n"])
        pc.Mult(b[rank], x)
        e1 = rel(x, xo[rank])
        rhs = [q["rhs"] * q["free"] for q in parts]
        uo, ito, _ = amg.pcg(rhs, tol=1e-8, maxsteps=100)
        xd = torch.zeros(p["n"], dtype=torch.float64, device="cuda")
        it, errs = pc._pcg(torch.from_numpy(rhs[rank]).cuda(), xd, 1e-8, 100)
        e2 = rel(xd.cpu().numpy(), uo[rank])
        good = e1 < 1e-10 and it == ito and e2 < 1e-8
        ok &= good
        print("rank %d graph=%d halo=%s: distributed levels %d, V-cycle rel err %.2e, PCG its %d (oracle %d), solution rel err %.2e -> %s"
              % (rank, graph, pc.HaloTransport(0), npar, e1, it, ito, e2, "ok" if good else "FAIL"), flush=True)
        del pc
    flag = torch.tensor([0.0 if ok else 1.0], device="cuda")
    dist.all_reduce(flag)
    comm.close()
    dist.destroy_process_group()
    sys.exit(0 if flag.item() == 0 else 1)


if __name__ == "__main__":
    main()
