// par.hpp -- host-side structures of the multi-rank (one rank == one GPU) path:
//   * Comm        : thin wrapper over the caller-supplied communication table (ngsamg_comm, include/ngsamg_b200.h)
//   * ParDofs     : NGSolve-ParallelDofs-like sharing information of one level + the DCC master/ghost lists
//                   (BasicDCCMap::CalcDOFMasters, src/base/linalg/dcc_map.cpp:494-543)
//   * cumulate / hybrid split / modified diagonal (DecomposeSparseMatrixHybrid, src/base/linalg/hybrid_matrix.cpp:17-307;
//     CalcHybridSmootherRDGItGeneric, src/base/smoothers/hybrid_smoother_utils.hpp:11-143)
//   * contraction of a distributed level onto rank 0 (CtrMap, src/base/coarsening/dof_contract.cpp:49-228, 557-727)
// No CUDA in here: everything is testable on a CPU-only machine through the host-only C entry points.
#pragma once
#include "../../include/ngsamg_b200.h"
#include "common.hpp"

namespace ngb {

struct Comm {
  ngsamg_comm c{};
  bool active() const { return c.size > 1; }
  int rank() const { return c.rank; }
  int size() const { return c.size; }
  // variable-size neighbour exchange (sizes travel first through the same callback); collective over the peers
  void exchange(const std::vector<i32> &peers, const std::vector<std::vector<char>> &send, std::vector<std::vector<char>> &recv) const;
  // sizes known on both sides
  void exchange_fixed(const std::vector<i32> &peers, const std::vector<const void *> &send, const std::vector<i64> &sbytes,
                      const std::vector<void *> &recv, const std::vector<i64> &rbytes) const;
  void allreduce_sum(double *v, int n) const;
  i64 allreduce_sum(i64 v) const;
};

struct ParDofs {
  i64 n = 0;
  std::vector<i32> peers;            // neighbour ranks, ascending
  std::vector<std::vector<i32>> ex;  // per neighbour: the shared local dofs; k-th entry here == k-th entry on the neighbour
  // ---- derived (derive())
  std::vector<std::vector<i32>> sharers;  // eqc id -> the OTHER ranks sharing the dofs of that class (ascending); class 0 = local
  std::vector<i32> eqc;                   // dof -> class id
  std::vector<i32> master_of;             // -1: this rank is master, else index into peers of the master rank (lowest sharer)
  std::vector<std::vector<i32>> m_ex, g_ex;  // DCC lists per neighbour: dofs I am master of / dofs whose master is that neighbour
  std::vector<i64> canon;                 // canonical position of a dof: identical relative order on every sharer
  void derive(int rank);
  bool is_master(i64 d) const { return master_of[d] < 0; }
  bool shared(i64 d) const { return eqc[d] != 0; }
  // class a may interpolate from class b:  sharers(b) is a superset of sharers(a)
  bool finer_or_equal(i32 a, i32 b) const;
};

// sum over the sharers of per-dof data (bs doubles per dof), identical bit pattern on every sharer (contributions are added
// in ascending rank order) -- MyAllReduceDofData with a deterministic order.
void allreduce_dof_data(const Comm &comm, const ParDofs &pd, int bs, std::vector<double> &data);

// A_cum(i,j) = sum over all ranks holding both i and j of their local contribution (entries between two dofs of this rank,
// fully assembled); rows of unshared dofs are returned unchanged.  Bit-identical on every rank that holds (i,j).
void cumulate_matrix(const Comm &comm, const ParDofs &pd, const HostBsr &A, HostBsr &Acum);

// M = master x master block of the assembled matrix, G = local entries whose row- and column-masters differ.
void hybrid_split(const ParDofs &pd, const HostBsr &A, const HostBsr &Acum, HostBsr &M, HostBsr &G);

// l1-style modified diagonal of the hybrid smoother: md_k = max(1, 0.51 (1 + ad_k)) d_k on master & free dofs, 0 elsewhere.
void hybrid_mod_diag(const Comm &comm, const ParDofs &pd, const HostBsr &Acum, const HostBsr &G, const uint8_t *free_mask,
                     std::vector<double> &md);

// sweep position of every row for HybridGSSmoother's stage order LOC_PART_1 | EX_PART | LOC_PART_2 (gssmoother.cpp:616-700,
// hybrid_base_smoother.cpp:508-573); rows that are not smoothed on this rank come last.  smoothed[i] = master & free.
void hybrid_sweep_order(const ParDofs &pd, const uint8_t *free_mask, std::vector<i32> &sweep_rank, std::vector<uint8_t> &smoothed,
                        i64 *split_ind);

// coarse-level sharing information from the vertex map of a consistent (class-respecting) coarsening
void coarse_pardofs(const ParDofs &fine, const std::vector<i32> &vmap, i64 ncoarse, ParDofs &coarse, int rank);

// renumber the dofs of a level: new index = perm[old]
void permute_pardofs(ParDofs &pd, const std::vector<i32> &perm);

// ---- contraction onto rank 0 ---------------------------------------------------------------------------------
struct Contraction {
  // on rank 0: for every rank r the map local dof -> dof of the merged level; merged matrix / masks
  std::vector<std::vector<i32>> dof_maps;
  HostBsr A;                       // merged (fully assembled) matrix
  std::vector<uint8_t> free_mask;  // empty = all free
  std::vector<double> xyz;
  std::vector<i64> n_local;        // local sizes of all ranks (rank 0)
};
void contract_to_root(const Comm &comm, const ParDofs &pd, const HostBsr &A, const std::vector<uint8_t> &free_mask,
                      const std::vector<double> &xyz, Contraction &out);

// ---- class-respecting coarsening (coarsen.cpp) -----------------------------------------------------------------
struct ParCoarsen {
  const ParDofs *pd = nullptr;
  const std::vector<double> *rowsum = nullptr;  // assembled row sums (bs doubles per dof), identical on all sharers
  const std::vector<i32> *cls = nullptr;        // (internal) sharing class per vertex in the canonical labelling
  int rank = 0;
};
// assembled row sums of a distributed matrix (per dof bh*bw doubles), bit-identical on all sharers
void assembled_row_sums(const Comm &comm, const ParDofs &pd, const HostBsr &A, std::vector<double> &rs);

}  // namespace ngb
