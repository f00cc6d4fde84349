"""worker of tests/test_parallel_host.py::test_hybrid_split_over_gloo -- launched with torch.distributed.run (gloo, CPU only).
Every process is one rank: it builds its own sub-assembled local problem, runs the library's host-side hybrid split through
the torch.distributed-backed communicator callbacks and checks its part against the multi-rank oracle."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch.distributed as dist
    import ngsamg_b200 as ng
    from ngsamg_b200 import parallel as par, synthetic as S
    from oracle import oracle as O
    from oracle import oracle_par as OP

    dist.init_process_group("gloo")
    rank, size = dist.get_rank(), dist.get_world_size()
    comm = par.TorchDistComm()
    dims, grid = (9, 8, 11), (1, 1, size)
    p = S.partition_poisson3d(*dims, grid=grid, rank=rank)
    A = ng.SparseMatrix(p["n"], p["n"], 1, 1, p["rowptr"], p["col"], p["val"])
    got = par.hybrid_host(A, par.Halo(p["peers"], p["ex"]), comm, p["free"])
    # oracle: all ranks simulated locally
    parts = S.partition_poisson3d(*dims, grid=grid)
    HL = OP.HybridLevel([O.Bsr(q["n"], q["n"], 1, 1, q["rowptr"], q["col"], q["val"]) for q in parts], [q["free"] for q in parts],
                        [q["peers"] for q in parts], [q["ex"] for q in parts])
    assert abs(got["M"].to_scipy() - HL.M[rank]).max() < 1e-14
    assert abs(got["G"].to_scipy() - HL.G[rank]).max() == 0
    assert np.allclose(got["mod_diag"], HL.md[rank].ravel(), rtol=1e-13, atol=0)
    assert (got["master"].astype(bool) == HL.master[rank]).all()
    # the BASELINE configs[4] workload in small (3x3 blocks, modulus jumping by 1e4 on a checkerboard): bench.py's box generator
    q = S.box_elasticity3d_jump(4, grid, rank, box_cells=2)
    Aq = ng.SparseMatrix(q["n"], q["n"], 3, 3, q["rowptr"], q["col"], q["val"])
    gq = par.hybrid_host(Aq, par.Halo(q["peers"], q["ex"]), comm, q["free"])
    qs = [S.box_elasticity3d_jump(4, grid, r, box_cells=2) for r in range(size)]
    HQ = OP.HybridLevel([O.Bsr(t["n"], t["n"], 3, 3, t["rowptr"], t["col"], t["val"]) for t in qs], [t["free"] for t in qs],
                        [t["peers"] for t in qs], [t["ex"] for t in qs])
    assert abs(gq["M"].to_scipy() - HQ.M[rank]).max() < 1e-9 * abs(HQ.M[rank]).max()
    assert abs(gq["G"].to_scipy() - HQ.G[rank]).max() == 0
    assert np.allclose(gq["mod_diag"], HQ.md[rank].ravel(), rtol=1e-12, atol=0)
    # all-reduce callback
    v = np.array([rank + 1.0, 2.0])
    assert np.allclose(comm._allreduce(v), [size * (size + 1) / 2, 2.0 * size])
    dist.barrier()
    dist.destroy_process_group()
    print("rank %d ok" % rank)


if __name__ == "__main__":
    main()
