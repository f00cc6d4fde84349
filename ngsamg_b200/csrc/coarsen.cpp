// coarsen.cpp -- host-side construction of the level DOF maps (prolongations).
//
// The reference builds P with SPW pairwise agglomeration + "semi-auxiliary" smoothed prolongation
// (src/base/coarsening/spw_agg_impl.hpp:1417-1831, src/base/factory/vertex_factory_impl.hpp:1599-1659,
// 1834-2433).  That construction is order- and quicksort-tie-dependent and is the
// "next" row of the scope table (§8f #1/#2); the hot path only CONSUMES P.  This file is our own,
// simplified builder in the same spirit so the library is usable stand-alone:
//   * successive pairwise matching rounds on the strength graph (aggregates of <= 2^rounds vertices,
//     reverse vertex order, scalar strength w_ij/sqrt(maxOD_i maxOD_j) with a 0.25 relative threshold,
//     orphan round) -- cf. SPWAgglomerator::FormAgglomerates;
//   * piecewise prolongation with rigid-body transport blocks, then one weighted-Jacobi-like smoothing
//     step on the auxiliary (edge weight) graph restricted to <= max_per_row coarse neighbours with
//     the sp_min_frac threshold -- cf. SemiAuxSProlMap's aux path.
// It does NOT claim bit-exact agreement with the reference's aggregates.  Users who need the
// reference's own DOF maps inject them with ngsamg_b200_set_prolongations().
#include <cstdlib>
#include "common.hpp"
#include "par.hpp"
#include <chrono>
#include <numeric>

namespace ngb {

namespace {

struct Graph {
  i64 n = 0;
  std::vector<i64> ptr;
  std::vector<i32> adj;
  std::vector<double> w;
  std::vector<double> vwt;    // "L2"/ground weight: |row sum| (h1_impl.hpp:383-431)
  std::vector<double> maxod;  // max(max incident edge weight, vwt)
  std::vector<i32> size;      // number of level-0 vertices represented
};

double entry_weight(const double *blk, int b)
{
  if (b == 1) return std::fabs(blk[0]);
  double tr = 0;
  for (int k = 0; k < b; k++) tr += blk[k * b + k];
  return std::fabs(tr) / b;  // trace-based weight for vector problems
}

// strength graph of the level matrix; `drop[v]` marks Dirichlet vertices (not part of any aggregate)
void graph_from_matrix(const HostBsr &A, const std::vector<uint8_t> &drop, Graph &G, const double *rowsum = nullptr,
                       const i32 *cls = nullptr)
{
  // rowsum: externally supplied (assembled) row sums; cls: only couplings inside one sharing class become graph edges
  const i64 n = A.nrows;
  const int b = A.bh, bb = b * b;
  G.n = n;
  G.ptr.assign(n + 1, 0);
  G.vwt.assign(n, 0.0);
  G.maxod.assign(n, 0.0);
  G.size.assign(n, 1);
  parallel_for(n, [&](i64 lo, i64 hi) {
    for (i64 i = lo; i < hi; i++) {
      i64 c = 0;
      if (!drop[i])
        for (i64 k = A.rowptr[i]; k < A.rowptr[i + 1]; k++) {
          i32 j = A.col[k];
          if (j != i && !drop[j] && (!cls || cls[j] == cls[i])) c++;
        }
      G.ptr[i + 1] = c;
    }
  });
  for (i64 i = 0; i < n; i++) G.ptr[i + 1] += G.ptr[i];
  G.adj.resize(G.ptr[n]);
  G.w.resize(G.ptr[n]);
  parallel_for(n, [&](i64 lo, i64 hi) {
    std::vector<double> rs(bb);
    for (i64 i = lo; i < hi; i++) {
      if (drop[i]) continue;
      i64 p = G.ptr[i];
      std::fill(rs.begin(), rs.end(), 0.0);
      double mx = 0;
      for (i64 k = A.rowptr[i]; k < A.rowptr[i + 1]; k++) {
        i32 j = A.col[k];
        const double *blk = &A.val[k * bb];
        for (int e = 0; e < bb; e++) rs[e] += blk[e];
        if (j == i || drop[j] || (cls && cls[j] != cls[i])) continue;
        double w = entry_weight(blk, b);
        G.adj[p] = j;
        G.w[p] = w;
        p++;
        mx = std::max(mx, w);
      }
      double vw = entry_weight(rowsum ? rowsum + i * bb : rs.data(), b);
      G.vwt[i] = vw;
      G.maxod[i] = std::max(mx, vw);
    }
  });
}

// one pairwise matching round.  cmap[v] = coarse id (numbered ascending by smallest member), -1 if dropped.
// forward = false: vertices are visited in descending order and ties go to the highest neighbour (the coarsening of the hierarchy);
// forward = true: ascending order, ties to the lowest neighbour (the tiling of a sweep: leftovers end up at the END of the sweep order)
i64 pairing_round(const Graph &G, const std::vector<uint8_t> &drop, double soc_thresh, std::vector<i32> &cmap, bool forward = false)
{
  const i64 n = G.n;
  std::vector<i32> mate(n, -2);  // -2 unhandled, -1 single, >=0 partner
  for (i64 vv = 0; vv < n; vv++) {
    const i64 v = forward ? vv : n - 1 - vv;
    if (mate[v] != -2) continue;
    if (drop[v]) { mate[v] = -1; continue; }
    double maxsoc = 0;
    const double mv = G.maxod[v];
    for (i64 e = G.ptr[v]; e < G.ptr[v + 1]; e++) {
      i32 j = G.adj[e];
      double d = mv * G.maxod[j];
      double soc = d > 0 ? G.w[e] / std::sqrt(d) : 0.0;
      maxsoc = std::max(maxsoc, soc);
    }
    const double thr = soc_thresh * maxsoc;
    i32 best = -1;
    double bestsoc = 0;
    for (i64 e = G.ptr[v]; e < G.ptr[v + 1]; e++) {
      i32 j = G.adj[e];
      if (mate[j] != -2 || drop[j]) continue;
      double d = mv * G.maxod[j];
      double soc = d > 0 ? G.w[e] / std::sqrt(d) : 0.0;
      if (soc <= 0 || soc < thr) continue;
      // strongest connection wins; ties -> the neighbour with the highest index (closest in numbering)
      if (soc > bestsoc || (soc == bestsoc && (forward ? (best < 0 || j < best) : j > best))) { bestsoc = soc; best = j; }
    }
    if (best >= 0) { mate[v] = best; mate[best] = (i32)v; }
    else mate[v] = -1;
  }
  cmap.assign(n, -1);
  i64 nc = 0;
  for (i64 v = 0; v < n; v++) {
    if (drop[v] || cmap[v] >= 0) continue;
    cmap[v] = (i32)nc;
    if (mate[v] >= 0) cmap[mate[v]] = (i32)nc;
    nc++;
  }
  return nc;
}

// coarse graph: weights summed, vwt summed, maxod = max(coarse incident weights, members' maxod)
void coarsen_graph(const Graph &G, const std::vector<i32> &cmap, i64 nc, Graph &C)
{
  const i64 n = G.n;
  std::vector<i64> mptr(nc + 1, 0);
  for (i64 v = 0; v < n; v++)
    if (cmap[v] >= 0) mptr[cmap[v] + 1]++;
  for (i64 c = 0; c < nc; c++) mptr[c + 1] += mptr[c];
  std::vector<i32> mem(mptr[nc]);
  {
    std::vector<i64> pos(mptr.begin(), mptr.end() - 1);
    for (i64 v = 0; v < n; v++)
      if (cmap[v] >= 0) mem[pos[cmap[v]]++] = (i32)v;
  }
  C.n = nc;
  C.ptr.assign(nc + 1, 0);
  C.vwt.assign(nc, 0.0);
  C.maxod.assign(nc, 0.0);
  C.size.assign(nc, 0);
  // two passes (count, fill) so that rows can be produced in parallel chunks
  const int nt = host_threads();
  const i64 chunk = (nc + nt - 1) / std::max(nt, 1);
  std::vector<std::vector<i32>> cadj(nt);
  std::vector<std::vector<double>> cw(nt);
  std::vector<std::thread> th;
  auto work = [&](int t) {
    i64 lo = t * chunk, hi = std::min(nc, lo + chunk);
    std::vector<std::pair<i32, double>> row;
    for (i64 c = lo; c < hi; c++) {
      row.clear();
      double vw = 0, mo = 0;
      i32 sz = 0;
      for (i64 m = mptr[c]; m < mptr[c + 1]; m++) {
        i32 v = mem[m];
        vw += G.vwt[v];
        mo = std::max(mo, G.maxod[v]);
        sz += G.size[v];
        for (i64 e = G.ptr[v]; e < G.ptr[v + 1]; e++) {
          i32 cj = cmap[G.adj[e]];
          if (cj < 0 || cj == c) continue;
          row.emplace_back(cj, G.w[e]);
        }
      }
      // sort by coarse neighbour, then merge duplicates (weights summed in a fixed order)
      std::stable_sort(row.begin(), row.end(),
                       [](const std::pair<i32, double> &a, const std::pair<i32, double> &b) { return a.first < b.first; });
      size_t u = 0;
      for (size_t q = 0; q < row.size(); q++) {
        if (u > 0 && row[u - 1].first == row[q].first) row[u - 1].second += row[q].second;
        else row[u++] = row[q];
      }
      row.resize(u);
      for (auto &pr : row) { cadj[t].push_back(pr.first); cw[t].push_back(pr.second); mo = std::max(mo, pr.second); }
      C.ptr[c + 1] = (i64)row.size();
      C.vwt[c] = vw;
      C.maxod[c] = mo;
      C.size[c] = sz;
    }
  };
  if (nt > 1 && nc > 20000) {
    for (int t = 0; t < nt; t++) th.emplace_back(work, t);
    for (auto &x : th) x.join();
  } else {
    for (int t = 0; t < nt; t++) work(t);
  }
  for (i64 c = 0; c < nc; c++) C.ptr[c + 1] += C.ptr[c];
  C.adj.resize(C.ptr[nc]);
  C.w.resize(C.ptr[nc]);
  for (int t = 0; t < nt; t++) {
    i64 lo = t * chunk;
    if (lo >= nc) break;
    std::copy(cadj[t].begin(), cadj[t].end(), C.adj.begin() + C.ptr[lo]);
    std::copy(cw[t].begin(), cw[t].end(), C.w.begin() + C.ptr[lo]);
  }
}

// orphan round: aggregates that still consist of one level-0 vertex join the most strongly connected
// neighbouring aggregate (JoiningIteration, spw_agg_impl.hpp:1265-1361).  returns new count, updates cmap
i64 orphan_round(const Graph &G, double soc_thresh, std::vector<i32> &join)
{
  const i64 n = G.n;
  join.assign(n, -1);
  for (i64 v = 0; v < n; v++) {
    if (G.size[v] != 1) continue;
    double maxsoc = 0;
    for (i64 e = G.ptr[v]; e < G.ptr[v + 1]; e++) {
      double d = G.maxod[v] * G.maxod[G.adj[e]];
      maxsoc = std::max(maxsoc, d > 0 ? G.w[e] / std::sqrt(d) : 0.0);
    }
    i32 best = -1;
    double bw = 0;
    for (i64 e = G.ptr[v]; e < G.ptr[v + 1]; e++) {
      i32 j = G.adj[e];
      if (G.size[j] <= 1) continue;
      double d = G.maxod[v] * G.maxod[j];
      double soc = d > 0 ? G.w[e] / std::sqrt(d) : 0.0;
      if (soc <= 0 || soc < soc_thresh * maxsoc) continue;
      if (soc > bw) { bw = soc; best = j; }
    }
    join[v] = best;
  }
  std::vector<i32> newid(n, -1);
  i64 nc = 0;
  for (i64 v = 0; v < n; v++)
    if (join[v] < 0) newid[v] = (i32)nc++;
  for (i64 v = 0; v < n; v++) join[v] = (join[v] < 0) ? newid[v] : newid[join[v]];
  return nc;
}

inline void skew_neg(const double *t, double *S /*3x3 = -skew(t)*/)
{
  // skew(t) w = t x w ;  -skew(t):
  S[0] = 0;      S[1] = t[2];   S[2] = -t[1];
  S[3] = -t[2];  S[4] = 0;      S[5] = t[0];
  S[6] = t[1];   S[7] = -t[0];  S[8] = 0;
}

// transport block Q (bf x bc) from coarse vertex at xc to fine vertex at xv
void transport_block(int bf, int bc, const double *xv, const double *xc, double *Q)
{
  std::fill(Q, Q + bf * bc, 0.0);
  if (bf == bc && bf != 6) { for (int k = 0; k < bf; k++) Q[k * bc + k] = 1.0; return; }
  double t[3] = {0, 0, 0};
  if (xv && xc) for (int k = 0; k < 3; k++) t[k] = xv[k] - xc[k];
  double S[9];
  skew_neg(t, S);
  if (bf == 3 && bc == 6) {  // [I | -skew(t)] : displacement of the rigid body motion (u, w) at xv
    for (int r = 0; r < 3; r++) {
      Q[r * 6 + r] = 1.0;
      for (int c = 0; c < 3; c++) Q[r * 6 + 3 + c] = S[r * 3 + c];
    }
  } else if (bf == 6 && bc == 6) {  // [[I, -skew(t)], [0, I]]
    for (int r = 0; r < 3; r++) {
      Q[r * 6 + r] = 1.0;
      Q[(3 + r) * 6 + 3 + r] = 1.0;
      for (int c = 0; c < 3; c++) Q[r * 6 + 3 + c] = S[r * 3 + c];
    }
  } else if (bf == 2 && bc == 3) {  // 2d elasticity: (u, w) -> u + w * (-t_y, t_x)
    Q[0] = 1; Q[1] = 0; Q[2] = -t[1];
    Q[3] = 0; Q[4] = 1; Q[5] = t[0];
  } else if (bf == 3 && bc == 3) {
    for (int k = 0; k < 3; k++) Q[k * 3 + k] = 1.0;
  } else {
    throw Error("transport_block: unsupported block shape " + std::to_string(bf) + "x" + std::to_string(bc));
  }
}

}  // namespace

void host_transpose(const HostBsr &A, HostBsr &T)
{
  // counting-sort transpose with per-block transposition (cf. TransposeSPMImpl, utils_sparseMM.cpp:54-93)
  T.nrows = A.ncols; T.ncols = A.nrows; T.bh = A.bw; T.bw = A.bh;
  T.rowptr.assign(T.nrows + 1, 0);
  const i64 nnz = A.nnz();
  for (i64 k = 0; k < nnz; k++) T.rowptr[A.col[k] + 1]++;
  for (i64 r = 0; r < T.nrows; r++) T.rowptr[r + 1] += T.rowptr[r];
  T.col.resize(nnz);
  T.val.resize(nnz * A.bs());
  std::vector<i64> pos(T.rowptr.begin(), T.rowptr.end() - 1);
  const int bh = A.bh, bw = A.bw;
  for (i64 i = 0; i < A.nrows; i++)
    for (i64 k = A.rowptr[i]; k < A.rowptr[i + 1]; k++) {
      i64 p = pos[A.col[k]]++;
      T.col[p] = (i32)i;
      const double *s = &A.val[k * bh * bw];
      double *d = &T.val[p * bh * bw];
      for (int r = 0; r < bh; r++)
        for (int c = 0; c < bw; c++) d[c * bh + r] = s[r * bw + c];
    }
}

// The builder proper.  `pc` (multi-rank mode, canonical vertex labels): pc->cls[v] = sharing class of v, pc->rowsum = assembled row
// sums; `A` then only holds the couplings a vertex may use (see build_prolongation below) and aggregation is restricted to
// vertices of the same class.
static void build_prolongation_canonical(const HostBsr &A, const uint8_t *free_mask, int bc, const std::vector<double> &xyz,
                                         const CoarsenOptions &opt, HostBsr &P, std::vector<i32> &vmap, std::vector<double> &cxyz,
                                         const ParCoarsen *pc)
{
  const i64 n = A.nrows;
  const int bf = A.bh;
  std::vector<uint8_t> drop(n, 0);
  if (free_mask) for (i64 i = 0; i < n; i++) drop[i] = free_mask[i] ? 0 : 1;
  const bool timing = std::getenv("NGSAMG_B200_TIMING") != nullptr;
  auto t_last = std::chrono::steady_clock::now();
  auto lap = [&](const char *what) {
    if (!timing) return;
    auto now = std::chrono::steady_clock::now();
    std::fprintf(stderr, "[coarsen] %-28s %.3f s\n", what, std::chrono::duration<double>(now - t_last).count());
    t_last = now;
  };
  Graph G0;
  graph_from_matrix(A, drop, G0, pc ? pc->rowsum->data() : nullptr, pc ? pc->cls->data() : nullptr);
  lap("graph_from_matrix");
  // isolated vertices are not aggregated (spw_agg_impl.hpp:1599-1614)
  for (i64 i = 0; i < n; i++) {
    if (drop[i] || G0.ptr[i + 1] != G0.ptr[i]) continue;
    if (!pc) { drop[i] = 1; continue; }
    // multi-rank: the graph only holds same-class edges; a vertex is isolated if it is unshared and has no neighbour at all
    if ((*pc->cls)[i] != 0) continue;
    bool any = false;
    for (i64 k = A.rowptr[i]; k < A.rowptr[i + 1] && !any; k++) any = (A.col[k] != i && !(free_mask && !free_mask[A.col[k]]));
    if (!any) drop[i] = 1;
  }

  // ---- aggregation: `rounds` pairwise matching rounds + orphan round
  vmap.assign(n, -1);
  std::vector<i32> cmap;
  Graph Gc, Gn;
  const Graph *cur = &G0;
  std::vector<uint8_t> nodrop;
  i64 nc = 0;
  for (int r = 0; r < opt.rounds; r++) {
    const std::vector<uint8_t> &dr = (r == 0) ? drop : nodrop;
    if (r > 0) nodrop.assign(cur->n, 0);
    nc = pairing_round(*cur, (r == 0) ? drop : nodrop, opt.soc_thresh, cmap);
    lap("pairing_round");
    (void)dr;
    if (r == 0) { for (i64 v = 0; v < n; v++) vmap[v] = cmap[v]; }
    else { for (i64 v = 0; v < n; v++) if (vmap[v] >= 0) vmap[v] = cmap[vmap[v]]; }
    coarsen_graph(*cur, cmap, nc, Gn);
    lap("coarsen_graph");
    std::swap(Gc, Gn);
    cur = &Gc;
  }
  {
    std::vector<i32> join;
    i64 nc2 = orphan_round(*cur, opt.soc_thresh, join);
    if (nc2 != nc) {
      for (i64 v = 0; v < n; v++) if (vmap[v] >= 0) vmap[v] = join[vmap[v]];
      nc = nc2;
    }
  }
  // renumber coarse vertices ascending by smallest member (keeps the coarse numbering local)
  {
    std::vector<i32> newid(nc, -1);
    i64 k = 0;
    for (i64 v = 0; v < n; v++)
      if (vmap[v] >= 0 && newid[vmap[v]] < 0) newid[vmap[v]] = (i32)k++;
    for (i64 v = 0; v < n; v++) if (vmap[v] >= 0) vmap[v] = newid[vmap[v]];
    nc = k;
  }
  // coarse vertex positions = centroid of the aggregate
  const bool have_xyz = !xyz.empty();
  cxyz.clear();
  if (have_xyz) {
    cxyz.assign(nc * 3, 0.0);
    std::vector<i32> cnt(nc, 0);
    for (i64 v = 0; v < n; v++)
      if (vmap[v] >= 0) { for (int k = 0; k < 3; k++) cxyz[vmap[v] * 3 + k] += xyz[v * 3 + k]; cnt[vmap[v]]++; }
    for (i64 c = 0; c < nc; c++) for (int k = 0; k < 3; k++) cxyz[c * 3 + k] /= std::max(cnt[c], 1);
  }

  lap("orphans + renumber");
  // ---- prolongation rows
  const int maxpr = std::max(1, opt.max_per_row);
  P.nrows = n; P.ncols = nc; P.bh = bf; P.bw = bc;
  P.rowptr.assign(n + 1, 0);
  std::vector<i32> rcol(n * maxpr, -1);
  std::vector<double> rw(n * maxpr, 0.0);
  const int bb = bf * bf;
  parallel_for(n, [&](i64 lo, i64 hi) {
    std::vector<std::pair<i32, double>> nb;
    for (i64 v = lo; v < hi; v++) {
      const i32 C = vmap[v];
      if (C < 0) continue;  // empty row (PWProlMap: perow = 0 when vmap == -1, vertex_factory_impl.hpp:1624-1626)
      i32 *oc = &rcol[v * maxpr];
      double *ow = &rw[v * maxpr];
      oc[0] = C; ow[0] = 1.0;
      int cnt = 1;
      if (opt.smooth && maxpr > 1) {
        nb.clear();
        double in_w = 0, tot = 0;
        for (i64 k = A.rowptr[v]; k < A.rowptr[v + 1]; k++) {
          i32 j = A.col[k];
          if (j == v || vmap[j] < 0) continue;
          double w = entry_weight(&A.val[k * bb], bf);
          if (w <= 0) continue;
          tot += w;
          i32 cj = vmap[j];
          if (cj == C) { in_w += w; continue; }
          bool found = false;
          for (auto &pr : nb) if (pr.first == cj) { pr.second += w; found = true; break; }
          if (!found) nb.emplace_back(cj, w);
        }
        // rank coarse neighbours by summed weight (desc), ties by coarse index (asc)
        std::sort(nb.begin(), nb.end(), [](const std::pair<i32, double> &a, const std::pair<i32, double> &b) {
          return a.second > b.second || (a.second == b.second && a.first < b.first);
        });
        double ws = in_w;
        for (auto &pr : nb) {
          if (cnt >= maxpr) break;
          if (pr.second < opt.min_frac * tot) break;
          oc[cnt] = pr.first; ow[cnt] = pr.second; ws += pr.second; cnt++;
        }
        if (cnt > 1 && ws > 0) {
          double others = 0;
          for (int k = 1; k < cnt; k++) { ow[k] = opt.omega * ow[k] / ws; others += ow[k]; }
          ow[0] = 1.0 - others;
        }
        // ascending column order
        for (int a = 1; a < cnt; a++)
          for (int q = a; q > 0 && oc[q] < oc[q - 1]; q--) { std::swap(oc[q], oc[q - 1]); std::swap(ow[q], ow[q - 1]); }
      }
      P.rowptr[v + 1] = cnt;
    }
  });
  for (i64 v = 0; v < n; v++) P.rowptr[v + 1] += P.rowptr[v];
  lap("prolongation rows");
  P.col.resize(P.rowptr[n]);
  P.val.assign(P.rowptr[n] * (i64)bf * bc, 0.0);
  parallel_for(n, [&](i64 lo, i64 hi) {
    std::vector<double> Q(bf * bc);
    for (i64 v = lo; v < hi; v++) {
      i64 p = P.rowptr[v];
      int cnt = (int)(P.rowptr[v + 1] - p);
      for (int k = 0; k < cnt; k++) {
        i32 c = rcol[v * maxpr + k];
        P.col[p + k] = c;
        transport_block(bf, bc, have_xyz ? &xyz[v * 3] : nullptr, have_xyz ? &cxyz[(i64)c * 3] : nullptr, Q.data());
        double a = rw[v * maxpr + k];
        for (int e = 0; e < bf * bc; e++) P.val[(p + k) * bf * bc + e] = a * Q[e];
      }
    }
  });
}

// Multi-rank mode (par != nullptr): `A` is the ASSEMBLED local matrix (cumulate_matrix) and the coarsening respects the sharing
// classes the way the reference's EQC-wise agglomeration does: vertices are only merged with vertices shared by
// the same set of ranks, and a vertex interpolates only from coarse vertices shared by at least its own set of ranks
// (vertex_factory_impl.hpp:1845-1848), so every sharer computes bit-identical aggregates and prolongation rows from data it
// holds: the class-restricted graph is built in a canonical vertex order that is the same on every sharer.
void build_prolongation(const HostBsr &A_in, const uint8_t *free_mask, int bc, const std::vector<double> &xyz_in,
                        const CoarsenOptions &opt, HostBsr &P, std::vector<i32> &vmap, std::vector<double> &cxyz,
                        const ParCoarsen *par)
{
  if (par) {
    const ParDofs &pd = *par->pd;
    const i64 n = A_in.nrows;
    const int bs = A_in.bs();
    // consistent order of the classes: lexicographic on the full rank set (sharers + this rank)
    const size_t ncls = pd.sharers.size();
    std::vector<std::vector<i32>> full(ncls);
    for (size_t c = 0; c < ncls; c++) { full[c] = pd.sharers[c]; full[c].push_back(par->rank); std::sort(full[c].begin(), full[c].end()); }
    std::vector<i32> cord(ncls), cls_order(ncls);
    std::iota(cord.begin(), cord.end(), 0);
    std::sort(cord.begin(), cord.end(), [&](i32 a, i32 b) { return full[a] < full[b]; });
    for (size_t q = 0; q < ncls; q++) cls_order[cord[q]] = (i32)q;
    // canonical labels: class-major (unshared vertices first -- nobody else sees them --, then the shared classes in their consistent
    // order), canonical key inside a class.  Counting sort by class; a class is only sorted by key when the keys are not already
    // ascending in the local number (they are on the fine level).
    std::vector<i32> order(n), lab(n);
    {
      std::vector<i32> slot(ncls);
      for (size_t c = 0; c < ncls; c++) slot[c] = (c == 0) ? 0 : 1 + cls_order[c] - (cls_order[c] > cls_order[0] ? 1 : 0);
      std::vector<i64> start(ncls + 1, 0);
      for (i64 v = 0; v < n; v++) start[slot[pd.eqc[v]] + 1]++;
      for (size_t c = 0; c < ncls; c++) start[c + 1] += start[c];
      std::vector<i64> pos(start.begin(), start.end() - 1);
      for (i64 v = 0; v < n; v++) order[pos[slot[pd.eqc[v]]]++] = (i32)v;
      for (size_t c = 0; c < ncls; c++) {
        bool sorted = true;
        for (i64 q = start[c] + 1; q < start[c + 1] && sorted; q++) sorted = pd.canon[order[q - 1]] < pd.canon[order[q]];
        if (!sorted) std::sort(order.begin() + start[c], order.begin() + start[c + 1], [&](i32 a, i32 b) { return pd.canon[a] < pd.canon[b]; });
      }
    }
    for (i64 q = 0; q < n; q++) lab[order[q]] = (i32)q;
    // class a may use couplings to class b: table instead of a set comparison per matrix entry
    std::vector<uint8_t> fe(ncls * ncls);
    for (size_t a = 0; a < ncls; a++) for (size_t b = 0; b < ncls; b++) fe[a * ncls + b] = pd.finer_or_equal((i32)a, (i32)b) ? 1 : 0;
    auto usable = [&](i64 i, i32 j) { return fe[(size_t)pd.eqc[i] * ncls + pd.eqc[j]] != 0; };
    // relabelled matrix restricted to the couplings a vertex may use: columns of a class that is shared by at least the row's ranks
    HostBsr B;
    B.nrows = n; B.ncols = n; B.bh = A_in.bh; B.bw = A_in.bw;
    B.rowptr.assign(n + 1, 0);
    parallel_for(n, [&](i64 lo, i64 hi) {
      for (i64 q = lo; q < hi; q++) {
        const i64 i = order[q];
        i64 c = 0;
        if (pd.eqc[i] == 0) c = A_in.rowptr[i + 1] - A_in.rowptr[i];   // an unshared vertex may use all its couplings
        else for (i64 e = A_in.rowptr[i]; e < A_in.rowptr[i + 1]; e++) c += usable(i, A_in.col[e]) ? 1 : 0;
        B.rowptr[q + 1] = c;
      }
    });
    for (i64 q = 0; q < n; q++) B.rowptr[q + 1] += B.rowptr[q];
    B.col.resize(B.rowptr[n]);
    B.val.resize((size_t)B.rowptr[n] * bs);
    parallel_for(n, [&](i64 lo, i64 hi) {
      std::vector<std::pair<i32, i64>> ent;
      for (i64 q = lo; q < hi; q++) {
        const i64 i = order[q];
        ent.clear();
        for (i64 e = A_in.rowptr[i]; e < A_in.rowptr[i + 1]; e++)
          if (pd.eqc[i] == 0 || usable(i, A_in.col[e])) ent.emplace_back(lab[A_in.col[e]], e);
        std::sort(ent.begin(), ent.end());
        i64 p = B.rowptr[q];
        for (auto &x : ent) { B.col[p] = x.first; std::memcpy(&B.val[p * bs], &A_in.val[x.second * bs], sizeof(double) * bs); p++; }
      }
    });
    std::vector<uint8_t> fm;
    if (free_mask) { fm.resize(n); for (i64 q = 0; q < n; q++) fm[q] = free_mask[order[q]]; }
    std::vector<double> xyz, rs((size_t)n * bs);
    if (!xyz_in.empty()) { xyz.resize((size_t)n * 3); for (i64 q = 0; q < n; q++) std::memcpy(&xyz[q * 3], &xyz_in[(i64)order[q] * 3], sizeof(double) * 3); }
    for (i64 q = 0; q < n; q++) std::memcpy(&rs[q * bs], &(*par->rowsum)[(i64)order[q] * bs], sizeof(double) * bs);
    std::vector<i32> cls(n);
    for (i64 q = 0; q < n; q++) cls[q] = pd.eqc[order[q]];
    ParCoarsen inner;
    inner.rowsum = &rs; inner.cls = &cls; inner.rank = par->rank;
    HostBsr Pq;
    std::vector<i32> vq;
    build_prolongation_canonical(B, free_mask ? fm.data() : nullptr, bc, xyz, opt, Pq, vq, cxyz, &inner);
    // back to the local numbering (rows only; coarse vertices keep the numbering of the canonical run)
    vmap.resize(n);
    for (i64 i = 0; i < n; i++) vmap[i] = vq[lab[i]];
    P = HostBsr();
    P.nrows = n; P.ncols = Pq.ncols; P.bh = Pq.bh; P.bw = Pq.bw;
    P.rowptr.assign(n + 1, 0);
    for (i64 i = 0; i < n; i++) P.rowptr[i + 1] = P.rowptr[i] + (Pq.rowptr[lab[i] + 1] - Pq.rowptr[lab[i]]);
    P.col.resize(Pq.nnz());
    P.val.resize(Pq.val.size());
    const int pbs = Pq.bs();
    parallel_for(n, [&](i64 lo, i64 hi) {
      for (i64 i = lo; i < hi; i++) {
        const i64 s0 = Pq.rowptr[lab[i]], len = Pq.rowptr[lab[i] + 1] - s0;
        if (!len) continue;
        std::memcpy(&P.col[P.rowptr[i]], &Pq.col[s0], sizeof(i32) * len);
        std::memcpy(&P.val[P.rowptr[i] * pbs], &Pq.val[s0 * pbs], sizeof(double) * len * pbs);
      }
    });
    return;
  }
  build_prolongation_canonical(A_in, free_mask, bc, xyz_in, opt, P, vmap, cxyz, nullptr);
}

// Clustering only (no orphan round, no prolongation): `rounds` pairwise matching rounds on the strength graph of the rows with
// keep[i] != 0 -> cluster id per row (-1 for the others; isolated kept rows become singleton clusters), clusters hold at most 2^rounds
// rows.  Used to tile a level for the two-level scheduled Gauss-Seidel sweep (tiles.cpp), there with forward = true: matching in ascending
// order with ties to the lowest neighbour lines the clusters up from the START of the sweep, so the incomplete leftover clusters sit at its end
// (measured on 61^3 Poisson, 64-row tiles: tile-DAG depth 44 forward against 74 reverse; axis-aligned 4x4x4 boxes give 46).
i64 cluster_rows(const HostBsr &A, const uint8_t *keep, int rounds, double soc_thresh, std::vector<i32> &cluster, bool forward)
{
  const i64 n = A.nrows;
  std::vector<uint8_t> drop(n, 0);
  if (keep) for (i64 i = 0; i < n; i++) drop[i] = keep[i] ? 0 : 1;
  Graph G0;
  graph_from_matrix(A, drop, G0);
  cluster.assign(n, -1);
  std::vector<i32> cmap;
  Graph Gc, Gn;
  const Graph *cur = &G0;
  std::vector<uint8_t> nodrop;
  i64 nc = 0;
  for (int r = 0; r < rounds; r++) {
    if (r > 0) nodrop.assign(cur->n, 0);
    nc = pairing_round(*cur, (r == 0) ? drop : nodrop, soc_thresh, cmap, forward);
    if (r == 0) { for (i64 v = 0; v < n; v++) cluster[v] = cmap[v]; }
    else { for (i64 v = 0; v < n; v++) if (cluster[v] >= 0) cluster[v] = cmap[cluster[v]]; }
    if (r + 1 < rounds) {
      coarsen_graph(*cur, cmap, nc, Gn);
      std::swap(Gc, Gn);
      cur = &Gc;
    }
  }
  return nc;
}

}  // namespace ngb
