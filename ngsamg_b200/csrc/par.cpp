// par.cpp -- host side of the multi-rank path (see par.hpp).  Our own restatement of what the reference does with MPI:
// DCC master/ghost lists, hybrid M/G split, modified diagonal, stage order of the hybrid Gauss-Seidel smoother, coarse-level
// sharing information and the contraction of a coarse level onto one rank.
#include "par.hpp"

#include <numeric>

namespace ngb {

// ---------------------------------------------------------------------------------------------------------------
// Comm
// ---------------------------------------------------------------------------------------------------------------
void Comm::exchange_fixed(const std::vector<i32> &peers, const std::vector<const void *> &send, const std::vector<i64> &sbytes,
                          const std::vector<void *> &recv, const std::vector<i64> &rbytes) const
{
  if (peers.empty()) return;
  if (!c.exchange) throw Error("communicator has no exchange callback");
  const int rc = c.exchange(c.ctx, (i32)peers.size(), peers.data(), send.data(), sbytes.data(), recv.data(), rbytes.data());
  if (rc) throw Error("communicator exchange callback failed (rc=" + std::to_string(rc) + ")");
}

void Comm::exchange(const std::vector<i32> &peers, const std::vector<std::vector<char>> &send, std::vector<std::vector<char>> &recv) const
{
  const size_t np = peers.size();
  recv.assign(np, {});
  if (np == 0) return;
  std::vector<i64> ssz(np), rsz(np, 0), eight(np, 8);
  std::vector<const void *> sp(np);
  std::vector<void *> rp(np);
  for (size_t k = 0; k < np; k++) { ssz[k] = (i64)send[k].size(); sp[k] = &ssz[k]; rp[k] = &rsz[k]; }
  exchange_fixed(peers, sp, eight, rp, eight);
  for (size_t k = 0; k < np; k++) { recv[k].resize(rsz[k]); sp[k] = send[k].data(); rp[k] = recv[k].data(); }
  exchange_fixed(peers, sp, ssz, rp, rsz);
}

void Comm::allreduce_sum(double *v, int n) const
{
  if (!active() || n == 0) return;
  if (!c.allreduce_sum) throw Error("communicator has no allreduce callback");
  const int rc = c.allreduce_sum(c.ctx, v, n);
  if (rc) throw Error("communicator allreduce callback failed (rc=" + std::to_string(rc) + ")");
}

i64 Comm::allreduce_sum(i64 v) const
{
  double d = (double)v;
  allreduce_sum(&d, 1);
  return (i64)std::llround(d);
}

// ---------------------------------------------------------------------------------------------------------------
// ParDofs
// ---------------------------------------------------------------------------------------------------------------
void ParDofs::derive(int rank)
{
  // classes of dofs with the same set of sharers ("equivalence classes" of the reference's EQCHierarchy)
  eqc.assign(n, 0);
  sharers.assign(1, {});
  std::map<std::pair<i32, i32>, i32> trans;  // (class, neighbour index) -> class with that neighbour added
  for (size_t kp = 0; kp < peers.size(); kp++) {
    if (kp > 0 && peers[kp] <= peers[kp - 1]) throw Error("halo: neighbour ranks must be ascending and distinct");
    if (peers[kp] == rank) throw Error("halo: a rank cannot be its own neighbour");
    for (i32 d : ex[kp]) {
      if (d < 0 || d >= n) throw Error("halo: shared dof index out of range");
      auto key = std::make_pair(eqc[d], (i32)kp);
      auto it = trans.find(key);
      if (it == trans.end()) {
        std::vector<i32> s = sharers[eqc[d]];
        s.push_back(peers[kp]);
        sharers.push_back(std::move(s));
        it = trans.emplace(key, (i32)sharers.size() - 1).first;
      }
      eqc[d] = it->second;
    }
  }
  std::vector<i32> cls_master(sharers.size(), -1);
  for (size_t c = 1; c < sharers.size(); c++)
    if (sharers[c][0] < rank) cls_master[c] = (i32)(std::lower_bound(peers.begin(), peers.end(), sharers[c][0]) - peers.begin());
  master_of.resize(n);
  for (i64 d = 0; d < n; d++) master_of[d] = cls_master[eqc[d]];
  m_ex.assign(peers.size(), {});
  g_ex.assign(peers.size(), {});
  for (size_t kp = 0; kp < peers.size(); kp++)
    for (i32 d : ex[kp]) {
      if (master_of[d] < 0) m_ex[kp].push_back(d);
      else if (master_of[d] == (i32)kp) g_ex[kp].push_back(d);
    }
  if ((i64)canon.size() != n) { canon.resize(n); std::iota(canon.begin(), canon.end(), (i64)0); }
}

bool ParDofs::finer_or_equal(i32 a, i32 b) const
{
  if (a == b) return true;
  const auto &sa = sharers[a], &sb = sharers[b];
  return std::includes(sb.begin(), sb.end(), sa.begin(), sa.end());
}

void permute_pardofs(ParDofs &pd, const std::vector<i32> &perm)
{
  // lists stay ascending; the renumbering must keep the relative order of two shared dofs the same on every sharer, then the sorted
  // lists still pair up entry by entry
  for (auto *ll : {&pd.ex, &pd.m_ex, &pd.g_ex})
    for (auto &l : *ll) { for (i32 &d : l) d = perm[d]; std::sort(l.begin(), l.end()); }
  auto mv = [&](auto &v) {
    auto t = v;
    for (i64 i = 0; i < pd.n; i++) t[perm[i]] = v[i];
    v.swap(t);
  };
  if ((i64)pd.eqc.size() == pd.n) mv(pd.eqc);
  if ((i64)pd.master_of.size() == pd.n) mv(pd.master_of);
  if ((i64)pd.canon.size() == pd.n) mv(pd.canon);
}

void coarse_pardofs(const ParDofs &fine, const std::vector<i32> &vmap, i64 ncoarse, ParDofs &coarse, int rank)
{
  coarse = ParDofs();
  coarse.n = ncoarse;
  // canonical key of a coarse vertex: the smallest canonical key among its members (same vertex on every sharer)
  coarse.canon.assign(ncoarse, std::numeric_limits<i64>::max());
  for (i64 v = 0; v < fine.n; v++)
    if (vmap[v] >= 0) coarse.canon[vmap[v]] = std::min(coarse.canon[vmap[v]], fine.canon[v]);
  std::vector<i32> stamp(ncoarse, -1);
  for (size_t kp = 0; kp < fine.peers.size(); kp++) {
    std::vector<i32> lst;
    for (i32 d : fine.ex[kp]) {
      const i32 c = vmap[d];
      if (c < 0 || stamp[c] == (i32)kp) continue;
      stamp[c] = (i32)kp;
      lst.push_back(c);
    }
    // ascending in the coarse number: the coarse numbering of build_prolongation is class-major / canonical inside a class, so the
    // relative order of two shared coarse vertices is the same on both sides and the sorted lists still pair up entry by entry
    std::sort(lst.begin(), lst.end());
    if (!lst.empty()) { coarse.peers.push_back(fine.peers[kp]); coarse.ex.push_back(std::move(lst)); }
  }
  coarse.derive(rank);
}

// ---------------------------------------------------------------------------------------------------------------
// per-dof all-reduce with a fixed (rank-ascending) summation order
// ---------------------------------------------------------------------------------------------------------------
void allreduce_dof_data(const Comm &comm, const ParDofs &pd, int bs, std::vector<double> &data)
{
  const size_t np = pd.peers.size();
  if (np == 0) return;
  std::vector<std::vector<double>> sb(np), rb(np);
  std::vector<const void *> sp(np);
  std::vector<void *> rp(np);
  std::vector<i64> bytes(np);
  for (size_t kp = 0; kp < np; kp++) {
    const auto &l = pd.ex[kp];
    sb[kp].resize(l.size() * bs);
    rb[kp].resize(l.size() * bs);
    for (size_t k = 0; k < l.size(); k++) std::memcpy(&sb[kp][k * bs], &data[(i64)l[k] * bs], sizeof(double) * bs);
    sp[kp] = sb[kp].data(); rp[kp] = rb[kp].data(); bytes[kp] = (i64)(sizeof(double) * l.size() * bs);
  }
  comm.exchange_fixed(pd.peers, sp, bytes, rp, bytes);
  // ascending rank order, own contribution at its place: lower ranks first, then self, then higher ranks
  const int me = comm.rank();
  std::vector<double> own = data;
  std::vector<uint8_t> started(pd.n, 0);
  auto add_from = [&](size_t kp) {
    const auto &l = pd.ex[kp];
    for (size_t k = 0; k < l.size(); k++) {
      const i64 d = l[k];
      double *dst = &data[d * bs];
      if (!started[d]) { std::memcpy(dst, &rb[kp][k * bs], sizeof(double) * bs); started[d] = 1; }
      else for (int e = 0; e < bs; e++) dst[e] += rb[kp][k * bs + e];
    }
  };
  size_t kp = 0;
  for (; kp < np && pd.peers[kp] < me; kp++) add_from(kp);
  for (i64 d = 0; d < pd.n; d++)
    if (started[d]) for (int e = 0; e < bs; e++) data[d * bs + e] += own[d * bs + e];
    else started[d] = 1;   // data[d] already holds the own value
  for (; kp < np; kp++) add_from(kp);
}

void assembled_row_sums(const Comm &comm, const ParDofs &pd, const HostBsr &A, std::vector<double> &rs)
{
  const int bs = A.bs();
  rs.assign((size_t)A.nrows * bs, 0.0);
  parallel_for(A.nrows, [&](i64 lo, i64 hi) {
    for (i64 i = lo; i < hi; i++)
      for (i64 e = A.rowptr[i]; e < A.rowptr[i + 1]; e++)
        for (int q = 0; q < bs; q++) rs[i * bs + q] += A.val[e * bs + q];
  });
  allreduce_dof_data(comm, pd, bs, rs);
}

// ---------------------------------------------------------------------------------------------------------------
// assembled ("cumulated") local matrix
// ---------------------------------------------------------------------------------------------------------------
namespace {

template <class T> void put(std::vector<char> &b, const T *p, size_t n)
{
  const size_t o = b.size();
  b.resize(o + sizeof(T) * n);
  if (n) std::memcpy(b.data() + o, p, sizeof(T) * n);
}
template <class T> const T *take(const std::vector<char> &b, size_t &off, size_t n)
{
  const T *p = reinterpret_cast<const T *>(b.data() + off);
  off += sizeof(T) * n;
  if (off > b.size()) throw Error("truncated message in neighbour exchange");
  return p;
}

}  // namespace

void cumulate_matrix(const Comm &comm, const ParDofs &pd, const HostBsr &A, HostBsr &Acum)
{
  const size_t np = pd.peers.size();
  const int bs = A.bs();
  if (np == 0) { Acum = A; return; }
  // sub-matrix on the dofs shared with each neighbour, in the pair-local numbering (SPM_DIAG, hybrid_matrix.cpp:57-100)
  std::vector<std::vector<char>> sb(np), rb;
  std::vector<i32> pos(A.nrows, -1);
  for (size_t kp = 0; kp < np; kp++) {
    const auto &l = pd.ex[kp];
    for (size_t k = 0; k < l.size(); k++) pos[l[k]] = (i32)k;
    std::vector<i64> rp(l.size() + 1, 0);
    std::vector<i32> ci;
    std::vector<double> va;
    for (size_t k = 0; k < l.size(); k++) {
      const i64 i = l[k];
      for (i64 e = A.rowptr[i]; e < A.rowptr[i + 1]; e++) {
        const i32 pj = pos[A.col[e]];
        if (pj < 0) continue;
        ci.push_back(pj);
        va.insert(va.end(), &A.val[e * bs], &A.val[e * bs] + bs);
      }
      rp[k + 1] = (i64)ci.size();
    }
    for (i32 d : l) pos[d] = -1;
    const i64 hdr[2] = {(i64)l.size(), (i64)ci.size()};
    put(sb[kp], hdr, 2);
    put(sb[kp], rp.data(), rp.size());
    put(sb[kp], ci.data(), ci.size());
    put(sb[kp], va.data(), va.size());
  }
  comm.exchange(pd.peers, sb, rb);
  // rebuild the shared rows: entries (column, source rank, values), merged per column in ascending rank order
  struct Ent { i32 col; i32 rank; const double *v; };
  std::vector<std::vector<Ent>> rows;     // per shared row
  std::vector<i32> rowslot(A.nrows, -1);
  std::vector<i64> shared_rows;
  for (i64 i = 0; i < A.nrows; i++)
    if (pd.eqc[i] != 0) { rowslot[i] = (i32)shared_rows.size(); shared_rows.push_back(i); }
  rows.resize(shared_rows.size());
  const int me = comm.rank();
  for (size_t s = 0; s < shared_rows.size(); s++) {
    const i64 i = shared_rows[s];
    for (i64 e = A.rowptr[i]; e < A.rowptr[i + 1]; e++) rows[s].push_back(Ent{A.col[e], (i32)me, &A.val[e * bs]});
  }
  for (size_t kp = 0; kp < np; kp++) {
    const auto &l = pd.ex[kp];
    size_t off = 0;
    const i64 *hdr = take<i64>(rb[kp], off, 2);
    if (hdr[0] != (i64)l.size()) throw Error("halo lists of rank " + std::to_string(me) + " and rank " + std::to_string(pd.peers[kp]) + " differ in length");
    const i64 *rp = take<i64>(rb[kp], off, l.size() + 1);
    const i32 *ci = take<i32>(rb[kp], off, (size_t)hdr[1]);
    const double *va = take<double>(rb[kp], off, (size_t)hdr[1] * bs);
    for (size_t k = 0; k < l.size(); k++) {
      auto &row = rows[rowslot[l[k]]];
      for (i64 e = rp[k]; e < rp[k + 1]; e++) row.push_back(Ent{l[ci[e]], pd.peers[kp], va + e * bs});
    }
  }
  std::vector<i64> newlen(A.nrows);
  for (i64 i = 0; i < A.nrows; i++) newlen[i] = A.rowptr[i + 1] - A.rowptr[i];
  for (size_t s = 0; s < rows.size(); s++) {
    auto &row = rows[s];
    std::sort(row.begin(), row.end(), [](const Ent &a, const Ent &b) { return a.col < b.col || (a.col == b.col && a.rank < b.rank); });
    i64 u = 0;
    for (size_t q = 0; q < row.size(); q++) if (q == 0 || row[q].col != row[q - 1].col) u++;
    newlen[shared_rows[s]] = u;
  }
  Acum = HostBsr();
  Acum.nrows = A.nrows; Acum.ncols = A.ncols; Acum.bh = A.bh; Acum.bw = A.bw;
  Acum.rowptr.assign(A.nrows + 1, 0);
  for (i64 i = 0; i < A.nrows; i++) Acum.rowptr[i + 1] = Acum.rowptr[i] + newlen[i];
  Acum.col.resize(Acum.rowptr[A.nrows]);
  Acum.val.assign((size_t)Acum.rowptr[A.nrows] * bs, 0.0);
  parallel_for(A.nrows, [&](i64 lo, i64 hi) {
    for (i64 i = lo; i < hi; i++) {
      i64 p = Acum.rowptr[i];
      if (rowslot[i] < 0) {
        const i64 len = A.rowptr[i + 1] - A.rowptr[i];
        std::memcpy(&Acum.col[p], &A.col[A.rowptr[i]], sizeof(i32) * len);
        std::memcpy(&Acum.val[p * bs], &A.val[A.rowptr[i] * bs], sizeof(double) * len * bs);
        continue;
      }
      const auto &row = rows[rowslot[i]];
      for (size_t q = 0; q < row.size(); q++) {
        if (q > 0 && row[q].col == row[q - 1].col) {
          for (int e = 0; e < bs; e++) Acum.val[(p - 1) * bs + e] += row[q].v[e];
        } else {
          Acum.col[p] = row[q].col;
          std::memcpy(&Acum.val[p * bs], row[q].v, sizeof(double) * bs);
          p++;
        }
      }
    }
  });
}

void hybrid_split(const ParDofs &pd, const HostBsr &A, const HostBsr &Acum, HostBsr &M, HostBsr &G)
{
  const i64 n = A.nrows;
  const int bs = A.bs();
  auto filter = [&](const HostBsr &S, HostBsr &D, auto keep) {
    D = HostBsr();
    D.nrows = n; D.ncols = n; D.bh = S.bh; D.bw = S.bw;
    D.rowptr.assign(n + 1, 0);
    parallel_for(n, [&](i64 lo, i64 hi) {
      for (i64 i = lo; i < hi; i++) {
        i64 c = 0;
        for (i64 e = S.rowptr[i]; e < S.rowptr[i + 1]; e++) c += keep(i, S.col[e]) ? 1 : 0;
        D.rowptr[i + 1] = c;
      }
    });
    for (i64 i = 0; i < n; i++) D.rowptr[i + 1] += D.rowptr[i];
    D.col.resize(D.rowptr[n]);
    D.val.resize((size_t)D.rowptr[n] * bs);
    parallel_for(n, [&](i64 lo, i64 hi) {
      for (i64 i = lo; i < hi; i++) {
        i64 p = D.rowptr[i];
        for (i64 e = S.rowptr[i]; e < S.rowptr[i + 1]; e++)
          if (keep(i, S.col[e])) {
            D.col[p] = S.col[e];
            std::memcpy(&D.val[p * bs], &S.val[e * bs], sizeof(double) * bs);
            p++;
          }
      }
    });
  };
  // M: master x master part of the assembled matrix (own entries + the ghost-ghost blocks the other sharers ship to the master,
  //    hybrid_matrix.cpp:57-245)
  filter(Acum, M, [&](i64 i, i32 j) { return pd.master_of[i] < 0 && pd.master_of[j] < 0; });
  // G: local entries whose row- and column-masters differ (hybrid_matrix.cpp:252-290)
  filter(A, G, [&](i64 i, i32 j) { return pd.master_of[i] != pd.master_of[j]; });
}

void hybrid_mod_diag(const Comm &comm, const ParDofs &pd, const HostBsr &Acum, const HostBsr &G, const uint8_t *free_mask,
                     std::vector<double> &md)
{
  const i64 n = Acum.nrows;
  const int b = Acum.bh, bs = b * b;
  // origDiag: assembled diagonal block (the reference all-reduces M(k,k), which holds it on the master)
  std::vector<double> od((size_t)n * bs, 0.0), sq((size_t)n * b, 0.0);
  parallel_for(n, [&](i64 lo, i64 hi) {
    for (i64 i = lo; i < hi; i++) {
      for (i64 e = Acum.rowptr[i]; e < Acum.rowptr[i + 1]; e++)
        if (Acum.col[e] == i) { std::memcpy(&od[i * bs], &Acum.val[e * bs], sizeof(double) * bs); break; }
      for (int l = 0; l < b; l++) sq[i * b + l] = std::sqrt(od[i * bs + l * b + l]);
    }
  });
  std::vector<double> ad((size_t)n * b, 0.0);
  parallel_for(n, [&](i64 lo, i64 hi) {
    for (i64 k = lo; k < hi; k++) {
      if (free_mask && !free_mask[k]) continue;
      for (i64 e = G.rowptr[k]; e < G.rowptr[k + 1]; e++) {
        const i64 j = G.col[e];
        for (int l = 0; l < b; l++)
          for (int m = 0; m < b; m++) ad[k * b + l] += std::fabs(G.val[e * bs + l * b + m]) / (sq[k * b + l] * sq[j * b + m]);
      }
    }
  });
  allreduce_dof_data(comm, pd, b, ad);
  md.assign((size_t)n * bs, 0.0);
  for (i64 k = 0; k < n; k++) {
    if (pd.master_of[k] >= 0 || (free_mask && !free_mask[k])) continue;
    double fac = 1.0;
    for (int l = 0; l < b; l++) fac = std::max(fac, 0.51 * (1.0 + ad[k * b + l]));
    for (int e = 0; e < bs; e++) md[k * bs + e] = fac * od[k * bs + e];
  }
}

void hybrid_sweep_order(const ParDofs &pd, const uint8_t *free_mask, std::vector<i32> &sweep_rank, std::vector<uint8_t> &smoothed,
                        i64 *split_out)
{
  const i64 n = pd.n;
  smoothed.assign(n, 0);
  std::vector<uint8_t> stage(n, 3);
  i64 nloc = 0;
  for (i64 k = 0; k < n; k++) {
    if (pd.master_of[k] >= 0 || (free_mask && !free_mask[k])) continue;
    smoothed[k] = 1;
    if (pd.eqc[k] == 0) nloc++;
  }
  // split_ind (gssmoother.cpp:664-678): N/2 without a freedofs mask, else the index of the median local dof
  i64 split = 0;
  if (!free_mask) split = n / 2;
  else {
    i64 cnt = 0;
    const i64 half = nloc / 2;
    for (i64 k = 0; k < n; k++)
      if (smoothed[k] && pd.eqc[k] == 0 && cnt++ == half) { split = k; break; }
  }
  for (i64 k = 0; k < n; k++) {
    if (!smoothed[k]) continue;
    stage[k] = pd.eqc[k] != 0 ? 1 : (k < split ? 0 : 2);
  }
  i64 cnt[4] = {0, 0, 0, 0}, start[4];
  for (i64 k = 0; k < n; k++) cnt[stage[k]]++;
  start[0] = 0;
  for (int s = 1; s < 4; s++) start[s] = start[s - 1] + cnt[s - 1];
  sweep_rank.resize(n);
  for (i64 k = 0; k < n; k++) sweep_rank[k] = (i32)(start[stage[k]]++);
  if (split_out) *split_out = split;
}

// ---------------------------------------------------------------------------------------------------------------
// contraction onto rank 0
// ---------------------------------------------------------------------------------------------------------------
void contract_to_root(const Comm &comm, const ParDofs &pd, const HostBsr &A, const std::vector<uint8_t> &free_mask,
                      const std::vector<double> &xyz, Contraction &out)
{
  const i64 n = A.nrows;
  const int bs = A.bs();
  const int me = comm.rank(), R = comm.size();
  // every dof gets the key (master rank, local index on the master); ghosts learn it from their master
  std::vector<i32> mrank(n, me), mindex(n);
  for (i64 d = 0; d < n; d++) mindex[d] = (i32)d;
  {
    const size_t np = pd.peers.size();
    std::vector<std::vector<i32>> sb(np), rb(np);
    std::vector<const void *> sp(np);
    std::vector<void *> rp(np);
    std::vector<i64> sbytes(np), rbytes(np);
    for (size_t kp = 0; kp < np; kp++) {
      sb[kp].assign(pd.m_ex[kp].begin(), pd.m_ex[kp].end());
      rb[kp].resize(pd.g_ex[kp].size());
      sp[kp] = sb[kp].data(); rp[kp] = rb[kp].data();
      sbytes[kp] = (i64)(sizeof(i32) * sb[kp].size()); rbytes[kp] = (i64)(sizeof(i32) * rb[kp].size());
    }
    comm.exchange_fixed(pd.peers, sp, sbytes, rp, rbytes);
    for (size_t kp = 0; kp < np; kp++)
      for (size_t k = 0; k < pd.g_ex[kp].size(); k++) { mrank[pd.g_ex[kp][k]] = pd.peers[kp]; mindex[pd.g_ex[kp][k]] = rb[kp][k]; }
  }
  // message to the root: sizes, keys, matrix, mask, coordinates
  std::vector<char> msg;
  const i64 hdr[4] = {n, A.nnz(), (i64)(free_mask.empty() ? 0 : 1), (i64)(xyz.empty() ? 0 : 1)};
  put(msg, hdr, 4);
  put(msg, mrank.data(), (size_t)n);
  put(msg, mindex.data(), (size_t)n);
  put(msg, A.rowptr.data(), (size_t)n + 1);
  put(msg, A.col.data(), (size_t)A.nnz());
  put(msg, A.val.data(), (size_t)A.nnz() * bs);
  if (hdr[2]) put(msg, free_mask.data(), (size_t)n);
  if (hdr[3]) put(msg, xyz.data(), (size_t)n * 3);
  std::vector<i32> peers;
  std::vector<std::vector<char>> sb, rb;
  if (me == 0) { for (int r = 1; r < R; r++) { peers.push_back(r); sb.emplace_back(); } }
  else { peers.push_back(0); sb.push_back(std::move(msg)); }
  comm.exchange(peers, sb, rb);
  out = Contraction();
  if (me != 0) return;
  // ---- root: merged numbering = masters of rank 0, masters of rank 1, ... in their local order
  struct Part { i64 n, nnz; const i32 *mrank, *mindex; const i64 *rp; const i32 *ci; const double *va; const uint8_t *fm; const double *xyz; };
  std::vector<Part> parts(R);
  parts[0] = Part{n, A.nnz(), mrank.data(), mindex.data(), A.rowptr.data(), A.col.data(), A.val.data(),
                  free_mask.empty() ? nullptr : free_mask.data(), xyz.empty() ? nullptr : xyz.data()};
  for (int r = 1; r < R; r++) {
    size_t off = 0;
    const i64 *h = take<i64>(rb[r - 1], off, 4);
    Part &p = parts[r];
    p.n = h[0]; p.nnz = h[1];
    p.mrank = take<i32>(rb[r - 1], off, (size_t)p.n);
    p.mindex = take<i32>(rb[r - 1], off, (size_t)p.n);
    p.rp = take<i64>(rb[r - 1], off, (size_t)p.n + 1);
    p.ci = take<i32>(rb[r - 1], off, (size_t)p.nnz);
    p.va = take<double>(rb[r - 1], off, (size_t)p.nnz * bs);
    p.fm = h[2] ? take<uint8_t>(rb[r - 1], off, (size_t)p.n) : nullptr;
    p.xyz = h[3] ? take<double>(rb[r - 1], off, (size_t)p.n * 3) : nullptr;
  }
  // ordinal of every master dof on its own rank, and the rank offsets
  std::vector<std::vector<i32>> ordinal(R);
  std::vector<i64> offset(R + 1, 0);
  for (int r = 0; r < R; r++) {
    ordinal[r].assign(parts[r].n, -1);
    i32 c = 0;
    for (i64 d = 0; d < parts[r].n; d++)
      if (parts[r].mrank[d] == r) ordinal[r][d] = c++;
    offset[r + 1] = offset[r] + c;
  }
  const i64 N = offset[R];
  out.n_local.resize(R);
  out.dof_maps.resize(R);
  for (int r = 0; r < R; r++) {
    out.n_local[r] = parts[r].n;
    out.dof_maps[r].resize(parts[r].n);
    for (i64 d = 0; d < parts[r].n; d++) {
      const i32 mr = parts[r].mrank[d];
      const i32 o = ordinal[mr][parts[r].mindex[d]];
      if (o < 0) throw Error("contraction: inconsistent master information");
      out.dof_maps[r][d] = (i32)(offset[mr] + o);
    }
  }
  // merged matrix: sum of the local matrices (ascending rank order per entry), DoAssembleMatrix dof_contract.cpp:557-727
  struct Ent { i32 col; i32 rank; const double *v; };
  std::vector<i64> cnt(N + 1, 0);
  for (int r = 0; r < R; r++)
    for (i64 i = 0; i < parts[r].n; i++) cnt[out.dof_maps[r][i] + 1] += parts[r].rp[i + 1] - parts[r].rp[i];
  for (i64 i = 0; i < N; i++) cnt[i + 1] += cnt[i];
  std::vector<Ent> ent(cnt[N]);
  {
    std::vector<i64> pos(cnt.begin(), cnt.end() - 1);
    for (int r = 0; r < R; r++)
      for (i64 i = 0; i < parts[r].n; i++) {
        const i64 gi = out.dof_maps[r][i];
        for (i64 e = parts[r].rp[i]; e < parts[r].rp[i + 1]; e++) ent[pos[gi]++] = Ent{out.dof_maps[r][parts[r].ci[e]], (i32)r, parts[r].va + e * bs};
      }
  }
  HostBsr &C = out.A;
  C.nrows = N; C.ncols = N; C.bh = A.bh; C.bw = A.bw;
  C.rowptr.assign(N + 1, 0);
  parallel_for(N, [&](i64 lo, i64 hi) {
    for (i64 i = lo; i < hi; i++) {
      std::sort(ent.begin() + cnt[i], ent.begin() + cnt[i + 1], [](const Ent &a, const Ent &b) { return a.col < b.col || (a.col == b.col && a.rank < b.rank); });
      i64 u = 0;
      for (i64 q = cnt[i]; q < cnt[i + 1]; q++) if (q == cnt[i] || ent[q].col != ent[q - 1].col) u++;
      C.rowptr[i + 1] = u;
    }
  }, 256);
  for (i64 i = 0; i < N; i++) C.rowptr[i + 1] += C.rowptr[i];
  C.col.resize(C.rowptr[N]);
  C.val.assign((size_t)C.rowptr[N] * bs, 0.0);
  parallel_for(N, [&](i64 lo, i64 hi) {
    for (i64 i = lo; i < hi; i++) {
      i64 p = C.rowptr[i];
      for (i64 q = cnt[i]; q < cnt[i + 1]; q++) {
        if (q > cnt[i] && ent[q].col == ent[q - 1].col) { for (int e = 0; e < bs; e++) C.val[(p - 1) * bs + e] += ent[q].v[e]; }
        else { C.col[p] = ent[q].col; std::memcpy(&C.val[p * bs], ent[q].v, sizeof(double) * bs); p++; }
      }
    }
  }, 256);
  bool any_mask = false, any_xyz = false;
  for (int r = 0; r < R; r++) { any_mask |= parts[r].fm != nullptr; any_xyz |= parts[r].xyz != nullptr; }
  if (any_mask) {
    out.free_mask.assign(N, 1);
    for (int r = 0; r < R; r++)
      if (parts[r].fm) for (i64 d = 0; d < parts[r].n; d++) if (!parts[r].fm[d]) out.free_mask[out.dof_maps[r][d]] = 0;
  }
  if (any_xyz) {
    out.xyz.assign((size_t)N * 3, 0.0);
    for (int r = 0; r < R; r++)
      if (parts[r].xyz) for (i64 d = 0; d < parts[r].n; d++) if (parts[r].mrank[d] == r) std::memcpy(&out.xyz[(i64)out.dof_maps[r][d] * 3], &parts[r].xyz[d * 3], sizeof(double) * 3);
  }
}

}  // namespace ngb
