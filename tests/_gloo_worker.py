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
    # all-reduce callback
    v = np.array([rank + 1.0, 2.0])
    assert np.allclose(comm._allreduce(v), [size * (size + 1) / 2, 2.0 * size])
    dist.barrier()
    dist.destroy_process_group()
    print("rank %d ok" % rank)


if __name__ == "__main__":
    main()
