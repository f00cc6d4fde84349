// C++ host-side mirror test of the MULTI-RANK host path (no device needed): two ranks = two std::threads with an in-process
// Communicator, a 1D-partitioned 3D 7-point Poisson problem, amg::DecomposeHybrid on each rank, then checks that
//   * M only couples master DOFs and holds the assembled master-master block (own entries + the block shipped by the ghost side),
//   * G holds exactly the couplings between DOFs with different masters,
//   * sum_r (M_r + G_r) x_r == A x for a consistent x   (HybridBaseMatrix::Mult, hybrid_matrix.cpp:433-453),
//   * the stage order is LOC_PART_1 | EX_PART | LOC_PART_2.
// Exit code 0 = pass.
#include <cmath>
#include <condition_variable>
#include <cstdio>
#include <cstring>
#include <map>
#include <mutex>
#include <thread>
#include <vector>

#include "ngsamg_b200.hpp"

struct World {
  int size;
  std::mutex mu;
  std::condition_variable cv;
  std::map<std::pair<int, int>, std::vector<std::vector<char>>> box;   // (src, dst) -> queue of messages
  std::vector<std::vector<double>> slots;
  int arrived = 0, generation = 0;
};

class ThreadComm : public amg::Communicator {
  World &w; int rank;
public:
  ThreadComm(World &aw, int r) : w(aw), rank(r) {}
  int Rank() const override { return rank; }
  int Size() const override { return w.size; }
  void Exchange(int np, const int32_t *peers, const void *const *sb, const int64_t *sn, void *const *rb, const int64_t *rn) override
  {
    {
      std::lock_guard<std::mutex> g(w.mu);
      for (int k = 0; k < np; k++) w.box[{rank, peers[k]}].emplace_back((const char *)sb[k], (const char *)sb[k] + sn[k]);
    }
    w.cv.notify_all();
    for (int k = 0; k < np; k++) {
      std::unique_lock<std::mutex> g(w.mu);
      auto &q = w.box[{peers[k], rank}];
      w.cv.wait(g, [&] { return !q.empty(); });
      if ((int64_t)q.front().size() != rn[k]) throw amg::Exception("message size mismatch");
      std::memcpy(rb[k], q.front().data(), (size_t)rn[k]);
      q.erase(q.begin());
    }
  }
  void AllReduceSum(double *v, int n) override
  {
    std::unique_lock<std::mutex> g(w.mu);
    w.slots[rank].assign(v, v + n);
    const int gen = w.generation;
    if (++w.arrived == w.size) { w.arrived = 0; w.generation++; w.cv.notify_all(); }
    else w.cv.wait(g, [&] { return w.generation != gen; });
    for (int i = 0; i < n; i++) { double s = 0; for (int r = 0; r < w.size; r++) s += w.slots[r][i]; v[i] = s; }
  }
};

// sub-assembled 7-point "finite difference as element sum" matrix of the slab z in [z0, z1] of an N x N x NZ grid: every grid edge
// contributes [[1,-1],[-1,1]] * (share), edges inside the interface plane are split half/half between the two slabs
static amg::SparseMat slab_matrix(int N, int z0, int z1, int NZ, std::vector<double> &diag_check)
{
  const int nz = z1 - z0 + 1;
  const int64_t n = (int64_t)N * N * nz;
  std::vector<std::map<int32_t, double>> rows(n);
  auto id = [&](int x, int y, int z) { return (int32_t)(x + N * (y + N * (z - z0))); };
  auto edge = [&](int32_t a, int32_t b, double w) { rows[a][a] += w; rows[b][b] += w; rows[a][b] -= w; rows[b][a] -= w; };
  for (int z = z0; z <= z1; z++)
    for (int y = 0; y < N; y++)
      for (int x = 0; x < N; x++) {
        const bool iface = (z == z0 && z0 > 0) || (z == z1 && z1 < NZ - 1);
        const double w = iface ? 0.5 : 1.0;
        if (x + 1 < N) edge(id(x, y, z), id(x + 1, y, z), w);
        if (y + 1 < N) edge(id(x, y, z), id(x, y + 1, z), w);
        if (z + 1 <= z1) edge(id(x, y, z), id(x, y, z + 1), 1.0);
        rows[id(x, y, z)][id(x, y, z)] += (iface ? 0.5 : 1.0) * 0.1;   // mass-like shift: SPD without Dirichlet rows
      }
  std::vector<int64_t> rp(n + 1, 0);
  std::vector<int32_t> ci;
  std::vector<double> v;
  diag_check.assign(n, 0.0);
  for (int64_t i = 0; i < n; i++) {
    for (auto &e : rows[i]) { ci.push_back(e.first); v.push_back(e.second); if (e.first == i) diag_check[i] = e.second; }
    rp[i + 1] = (int64_t)ci.size();
  }
  return amg::SparseMat(n, n, 1, 1, rp, ci, v);
}

int main()
{
  const int N = 6, NZ = 9, ZC = 4;    // rank 0: planes 0..4, rank 1: planes 4..8, plane 4 shared (master = rank 0)
  World w;
  w.size = 2; w.slots.resize(2);
  std::vector<double> d0, d1;
  amg::SparseMat A[2] = {slab_matrix(N, 0, ZC, NZ, d0), slab_matrix(N, ZC, NZ - 1, NZ, d1)};
  const int plane = N * N;
  std::vector<int32_t> top(plane), bottom(plane);
  for (int k = 0; k < plane; k++) { top[k] = plane * ZC + k; bottom[k] = k; }
  amg::ParallelDofs pd[2] = {amg::ParallelDofs(A[0].nrows, {1}, {top}), amg::ParallelDofs(A[1].nrows, {0}, {bottom})};
  amg::HybridMatrix H[2];
  bool failed = false;
  auto work = [&](int r) {
    try { ThreadComm c(w, r); H[r] = amg::DecomposeHybrid(A[r], pd[r], c); }
    catch (const amg::Exception &e) { std::printf("rank %d: amg::Exception: %s\n", r, e.what()); failed = true; }
  };
  std::thread t0(work, 0), t1(work, 1);
  t0.join(); t1.join();
  if (failed) return 2;
  int bad = 0;
  // masters: rank 0 owns everything it has, rank 1 everything but its bottom plane
  for (int64_t i = 0; i < A[0].nrows; i++) bad += H[0].master[i] != 1;
  for (int64_t i = 0; i < A[1].nrows; i++) bad += H[1].master[i] != (i >= plane ? 1 : 0);
  // rank 1: M has no entry touching the ghost plane, G has exactly the ghost <-> master couplings
  for (int64_t i = 0; i < A[1].nrows; i++) {
    for (int64_t k = H[1].M.rowptr[i]; k < H[1].M.rowptr[i + 1]; k++) bad += (i < plane || H[1].M.col[k] < plane);
    for (int64_t k = H[1].G.rowptr[i]; k < H[1].G.rowptr[i + 1]; k++) bad += ((i < plane) == (H[1].G.col[k] < plane));
  }
  bad += H[0].G.NZE() != 0;
  // rank 0: the interface block of M is fully assembled (both halves of the in-plane edges)
  for (int k = 0; k < plane; k++) {
    const int64_t i = top[k];
    for (int64_t e = H[0].M.rowptr[i]; e < H[0].M.rowptr[i + 1]; e++)
      if (H[0].M.col[e] == i && std::fabs(H[0].M.val[e] - (d0[i] + d1[bottom[k]])) > 1e-14) bad++;
  }
  // operator identity against the assembled global matrix, x = a smooth consistent vector
  const int64_t ng = (int64_t)N * N * NZ;
  std::vector<double> xg(ng), yg(ng, 0.0), ysum(ng, 0.0);
  for (int64_t i = 0; i < ng; i++) xg[i] = std::sin(0.37 * i) + 0.01 * i;
  auto gid = [&](int r, int64_t i) { return r == 0 ? i : i + (int64_t)plane * ZC; };
  for (int r = 0; r < 2; r++)
    for (int64_t i = 0; i < A[r].nrows; i++) {
      for (int64_t k = A[r].rowptr[i]; k < A[r].rowptr[i + 1]; k++) yg[gid(r, i)] += A[r].val[k] * xg[gid(r, A[r].col[k])];
      for (const amg::SparseMat *S : {&H[r].M, &H[r].G})
        for (int64_t k = S->rowptr[i]; k < S->rowptr[i + 1]; k++) ysum[gid(r, i)] += S->val[k] * xg[gid(r, S->col[k])];
    }
  double err = 0;
  for (int64_t i = 0; i < ng; i++) err = std::fmax(err, std::fabs(yg[i] - ysum[i]));
  // stage order on rank 0: local rows below split, the shared (exchange) rows, local rows from split on
  const int64_t n0 = A[0].nrows, split = n0 / 2;
  for (int64_t i = 0; i < n0; i++) {
    const bool ex = i >= (int64_t)plane * ZC;
    const int stage_i = ex ? 1 : (i < split ? 0 : 2);
    for (int64_t j = i + 1; j < n0 && j < i + 40; j++) {
      const bool exj = j >= (int64_t)plane * ZC;
      const int stage_j = exj ? 1 : (j < split ? 0 : 2);
      if (stage_i < stage_j || (stage_i == stage_j)) bad += !(H[0].sweep_rank[i] < H[0].sweep_rank[j]);
      else bad += !(H[0].sweep_rank[i] > H[0].sweep_rank[j]);
    }
  }
  // modified diagonal: >= assembled diagonal on master rows, 0 on ghosts
  for (int64_t i = 0; i < plane; i++) bad += H[1].mod_diag[i] != 0.0;
  for (int k = 0; k < plane; k++) bad += !(H[0].mod_diag[top[k]] >= (d0[top[k]] + d1[bottom[k]]) * (1 - 1e-14));
  std::printf("hybrid split: nnz(M0)=%lld nnz(G1)=%lld operator_err=%.2e bad=%d\n", (long long)H[0].M.NZE(), (long long)H[1].G.NZE(), err, bad);
  return (bad == 0 && err < 1e-12) ? 0 : 1;
}
