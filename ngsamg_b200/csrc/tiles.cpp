// tiles.cpp -- host construction of the two-level (tile DAG x tile-local levels) Gauss-Seidel schedule, see tiles.hpp.
#include "tiles.hpp"

#include <atomic>
#include <cstdlib>
#include <numeric>
#include <queue>

namespace ngb {

namespace {

// strongly connected components of the tile graph (edges pred -> tile), iterative Tarjan; comp ids are assigned in
// reverse topological order of the condensation (a component gets its id when it is popped)
[[maybe_unused]] i64 tarjan_scc(i64 nt, const std::vector<i64> &succ_ptr, const std::vector<i32> &succ, std::vector<i32> &comp)
{
  comp.assign(nt, -1);
  std::vector<i32> index(nt, -1), low(nt, 0), stack, callstack;
  std::vector<i64> edge_pos(nt, 0);
  std::vector<uint8_t> onstack(nt, 0);
  i32 next_index = 0;
  i64 ncomp = 0;
  for (i64 root = 0; root < nt; root++) {
    if (index[root] >= 0) continue;
    callstack.push_back((i32)root);
    index[root] = low[root] = next_index++;
    stack.push_back((i32)root);
    onstack[root] = 1;
    edge_pos[root] = succ_ptr[root];
    while (!callstack.empty()) {
      const i32 v = callstack.back();
      if (edge_pos[v] < succ_ptr[v + 1]) {
        const i32 w = succ[edge_pos[v]++];
        if (index[w] < 0) {
          index[w] = low[w] = next_index++;
          stack.push_back(w);
          onstack[w] = 1;
          edge_pos[w] = succ_ptr[w];
          callstack.push_back(w);
        } else if (onstack[w]) low[v] = std::min(low[v], index[w]);
      } else {
        if (low[v] == index[v]) {
          for (;;) {
            const i32 w = stack.back();
            stack.pop_back();
            onstack[w] = 0;
            comp[w] = (i32)ncomp;
            if (w == v) break;
          }
          ncomp++;
        }
        callstack.pop_back();
        if (!callstack.empty()) low[callstack.back()] = std::min(low[callstack.back()], low[v]);
      }
    }
  }
  return ncomp;
}

// CSR of unique (from -> to) pairs given per-"to" predecessor lists
void transpose_lists(i64 nt, const std::vector<i64> &pptr, const std::vector<i32> &pl, std::vector<i64> &sptr, std::vector<i32> &sl)
{
  sptr.assign(nt + 1, 0);
  for (i32 p : pl) sptr[p + 1]++;
  for (i64 t = 0; t < nt; t++) sptr[t + 1] += sptr[t];
  sl.resize(pl.size());
  std::vector<i64> pos(sptr.begin(), sptr.end() - 1);
  for (i64 t = 0; t < nt; t++)
    for (i64 k = pptr[t]; k < pptr[t + 1]; k++) sl[pos[pl[k]]++] = (i32)t;
}

}  // namespace

void build_tile_schedule(const HostBsr &A, const std::vector<uint8_t> &mask, const std::vector<i32> &sweep_rank, int rounds, int max_rows,
                         TileSchedule &ts, const std::vector<i32> *cluster_hint)
{
  ts = TileSchedule();
  const i64 n = A.nrows;
  ts.n = n;
  const bool hm = !mask.empty();
  const bool natural = sweep_rank.empty();
  auto smoothed = [&](i64 i) { return !hm || mask[i]; };
  auto rank_of = [&](i64 i) -> i64 { return natural ? i : (i64)sweep_rank[i]; };
  if (max_rows % 32 || max_rows < 32 || max_rows > 1024) throw Error("tile capacity must be a multiple of 32 rows, at most 1024");   // the warp kernel takes <= 64, a CTA-per-tile kernel more

  // ---- 1. clusters of graph-neighbouring smoothed rows: only a HINT for which rows should share a tile
  std::vector<i32> agg;
  i64 nagg = 0;
  if (cluster_hint && (i64)cluster_hint->size() == n) {
    agg = *cluster_hint;
    for (i64 i = 0; i < n; i++) {
      if (!smoothed(i)) agg[i] = -1;
      else if (agg[i] < 0) throw Error("tile schedule: a smoothed row has no cluster in the hint");
      nagg = std::max<i64>(nagg, (i64)agg[i] + 1);
    }
  } else
    nagg = cluster_rows(A, hm ? mask.data() : nullptr, rounds, 0.25, agg, true);   // stricter strength thresholds (0.6 .. 0.99) were tried: deeper tile DAGs

  // ---- 2. tiles = consecutive chunks of a topological order of the row DAG (=> the tile graph is acyclic by construction).
  // The order is produced by list scheduling that stays inside one cluster as long as that cluster has executable rows: a cluster
  // whose outside dependencies are all done becomes one tile; a cluster that is not "convex" in the DAG is split where it has to be.
  std::vector<i32> indeg(n, 0), cl(n, -1);
  parallel_for(n, [&](i64 lo, i64 hi) {
    for (i64 i = lo; i < hi; i++) {
      if (!smoothed(i)) continue;
      const i64 ri = rank_of(i);
      i32 d = 0;
      for (i64 k = A.rowptr[i]; k < A.rowptr[i + 1]; k++) {
        const i32 jj = A.col[k];
        if (jj != i && smoothed(jj) && rank_of(jj) < ri) d++;
      }
      indeg[i] = d;
    }
  });
  using RK = std::pair<i64, i32>;   // (sweep rank, row)
  std::priority_queue<RK, std::vector<RK>, std::greater<RK>> heap, work;
  std::vector<std::vector<i32>> ready(nagg);
  std::vector<uint8_t> emitted(n, 0);
  // members of every cluster in sweep order, and the number of not yet satisfied dependencies that leave the cluster
  std::vector<i64> aptr(nagg + 1, 0);
  std::vector<i32> amem;
  std::vector<i64> ext(nagg, 0);
  {
    for (i64 i = 0; i < n; i++) if (agg[i] >= 0) aptr[agg[i] + 1]++;
    for (i64 a = 0; a < nagg; a++) aptr[a + 1] += aptr[a];
    amem.resize(aptr[nagg]);
    std::vector<i64> pos(aptr.begin(), aptr.end() - 1);
    for (i64 i = 0; i < n; i++) if (agg[i] >= 0) amem[pos[agg[i]]++] = (i32)i;
    parallel_for(nagg, [&](i64 lo, i64 hi) {
      for (i64 a = lo; a < hi; a++) {
        if (!natural) std::sort(amem.begin() + aptr[a], amem.begin() + aptr[a + 1], [&](i32 x, i32 y) { return sweep_rank[x] < sweep_rank[y]; });
        i64 e = 0;
        for (i64 m = aptr[a]; m < aptr[a + 1]; m++) {
          const i64 i = amem[m];
          const i64 ri = rank_of(i);
          for (i64 k = A.rowptr[i]; k < A.rowptr[i + 1]; k++) {
            const i32 jj = A.col[k];
            if (jj != i && smoothed(jj) && agg[jj] != a && rank_of(jj) < ri) e++;
          }
        }
        ext[a] = e;
      }
    }, 256);
  }
  // clusters whose outside dependencies are all done, in the order of their first row
  std::priority_queue<RK, std::vector<RK>, std::greater<RK>> full;
  for (i64 a = 0; a < nagg; a++)
    if (ext[a] == 0 && aptr[a + 1] > aptr[a]) full.emplace(rank_of(amem[aptr[a]]), (i32)a);
  for (i64 i = 0; i < n; i++)
    if (smoothed(i) && indeg[i] == 0) { heap.emplace(rank_of(i), (i32)i); ready[agg[i]].push_back((i32)i); }
  const int min_fill = 1;           // every emission closes its tile: sharing a tile between unrelated fragments would chain distant regions
  std::vector<i64> mptr{0};
  std::vector<i32> mem;
  mem.reserve(n);
  i64 nt = 0;
  i64 cur_rows = 0;
  auto close_tile = [&]() {
    if (cur_rows == 0) return;
    mptr.push_back((i64)mem.size());
    nt++;
    cur_rows = 0;
  };
  auto emit = [&](i32 x, i32 a, bool to_work) {
    if (cur_rows == max_rows) close_tile();
    emitted[x] = 1;
    cl[x] = (i32)nt;
    mem.push_back(x);
    cur_rows++;
    const i64 rx = rank_of(x);
    for (i64 k = A.rowptr[x]; k < A.rowptr[x + 1]; k++) {
      const i32 y = A.col[k];
      if (y == x || !smoothed(y) || rank_of(y) < rx) continue;
      const i32 ay = agg[y];
      if (ay != a && --ext[ay] == 0) {
        // first not yet emitted member of that cluster keys the queue
        i64 m = aptr[ay];
        while (m < aptr[ay + 1] && emitted[amem[m]]) m++;
        if (m < aptr[ay + 1]) full.emplace(rank_of(amem[m]), ay);
      }
      if (--indeg[y] == 0) {
        if (ay == a && to_work) work.emplace(rank_of(y), y);
        else { ready[ay].push_back(y); heap.emplace(rank_of(y), y); }
      }
    }
  };
  for (;;) {
    if (!full.empty()) {
      // a cluster that can run to completion: all its remaining rows, in sweep order, form (the rest of) a tile
      const i32 a = full.top().second;
      full.pop();
      i64 rem = 0;
      for (i64 m = aptr[a]; m < aptr[a + 1]; m++) rem += emitted[amem[m]] ? 0 : 1;
      if (rem == 0) continue;
      if (cur_rows > 0 && cur_rows + rem > max_rows) close_tile();
      for (i64 m = aptr[a]; m < aptr[a + 1]; m++)
        if (!emitted[amem[m]]) emit(amem[m], a, false);
      ready[a].clear();
      if (cur_rows >= min_fill) close_tile();
      continue;
    }
    // no cluster is complete-able (they wait for each other): run the executable part of the cluster owning the first ready row
    bool progressed = false;
    while (!heap.empty()) {
      const i32 r = heap.top().second;
      heap.pop();
      if (emitted[r]) continue;
      const i32 a = agg[r];
      for (i32 x : ready[a]) if (!emitted[x]) work.emplace(rank_of(x), x);
      ready[a].clear();
      while (!work.empty()) {
        const i32 x = work.top().second;
        work.pop();
        if (emitted[x]) continue;
        emit(x, a, true);
      }
      if (cur_rows >= min_fill) close_tile();
      progressed = true;
      break;
    }
    if (!progressed) break;
  }
  close_tile();
  {
    i64 nsm = 0;
    for (i64 i = 0; i < n; i++) nsm += smoothed(i) ? 1 : 0;
    if ((i64)mem.size() != nsm) return;   // not every row became executable: the matrix pattern is not symmetric -> no tiling
  }
  // predecessor tiles of every tile: tiles holding a row that some row of the tile depends on
  std::vector<i64> pptr, sptr;
  std::vector<i32> pl, sl;
  {
    std::vector<std::vector<i32>> per(nt);
    parallel_for(nt, [&](i64 lo, i64 hi) {
      std::vector<i32> tmp;
      for (i64 t = lo; t < hi; t++) {
        tmp.clear();
        for (i64 m = mptr[t]; m < mptr[t + 1]; m++) {
          const i64 i = mem[m];
          const i64 ri = rank_of(i);
          for (i64 k = A.rowptr[i]; k < A.rowptr[i + 1]; k++) {
            const i32 jj = A.col[k];
            if (jj == i || cl[jj] < 0 || cl[jj] == t) continue;
            if (rank_of(jj) < ri) tmp.push_back(cl[jj]);
          }
        }
        std::sort(tmp.begin(), tmp.end());
        tmp.erase(std::unique(tmp.begin(), tmp.end()), tmp.end());
        per[t] = tmp;
      }
    }, 256);
    pptr.assign(nt + 1, 0);
    for (i64 t = 0; t < nt; t++) pptr[t + 1] = pptr[t] + (i64)per[t].size();
    pl.resize(pptr[nt]);
    for (i64 t = 0; t < nt; t++) std::copy(per[t].begin(), per[t].end(), pl.begin() + pptr[t]);
  }
  transpose_lists(nt, pptr, pl, sptr, sl);
  // members of a tile in sweep order (the emission order inside a tile already is one; sort for a canonical layout)
  parallel_for(nt, [&](i64 lo, i64 hi) {
    for (i64 t = lo; t < hi; t++) std::sort(mem.begin() + mptr[t], mem.begin() + mptr[t + 1], [&](i32 x, i32 y) { return rank_of(x) < rank_of(y); });
  }, 1024);

  // ---- 3. levels of the tile DAG (longest path), by Kahn's algorithm
  std::vector<i32> tlevel(nt, 0), tdeg(nt);
  for (i64 t = 0; t < nt; t++) tdeg[t] = (i32)(pptr[t + 1] - pptr[t]);
  std::vector<i32> queue;
  queue.reserve(nt);
  for (i64 t = 0; t < nt; t++) if (!tdeg[t]) queue.push_back((i32)t);
  for (size_t q = 0; q < queue.size(); q++) {
    const i32 t = queue[q];
    for (i64 k = sptr[t]; k < sptr[t + 1]; k++) {
      const i32 s = sl[k];
      tlevel[s] = std::max(tlevel[s], tlevel[t] + 1);
      if (--tdeg[s] == 0) queue.push_back(s);
    }
  }
  if ((i64)queue.size() != nt) throw Error("tile schedule: tile graph has a cycle");
  int depth = 0;
  for (i64 t = 0; t < nt; t++) depth = std::max(depth, tlevel[t] + 1);
  ts.tile_depth = depth;

  // ---- 4. schedule order: tile level major, then by the first row of the tile in the sweep
  std::vector<i32> order(nt), newid(nt);
  std::iota(order.begin(), order.end(), 0);
  std::vector<i64> firstrank(nt, 0);
  for (i64 t = 0; t < nt; t++) firstrank[t] = mptr[t + 1] > mptr[t] ? rank_of(mem[mptr[t]]) : 0;
  std::sort(order.begin(), order.end(), [&](i32 a, i32 b) {
    return tlevel[a] < tlevel[b] || (tlevel[a] == tlevel[b] && firstrank[a] < firstrank[b]);
  });
  for (i64 q = 0; q < nt; q++) newid[order[q]] = (i32)q;

  // ---- 5. tile-local dependency levels and the row placement
  i64 nnon = 0;
  for (i64 i = 0; i < n; i++) nnon += smoothed(i) ? 0 : 1;
  ts.nonfree_pad = (nnon + 31) / 32 * 32;
  ts.ntiles = nt;
  ts.tile_slice.assign(nt + 1, 0);
  ts.tile_nlev.assign(nt, 0);
  ts.tile_slice[0] = (i32)(ts.nonfree_pad / 32);
  for (i64 q = 0; q < nt; q++) {
    const i64 t = order[q];
    const i64 rows = mptr[t + 1] - mptr[t];
    ts.tile_slice[q + 1] = ts.tile_slice[q] + (i32)((rows + 31) / 32);
  }
  ts.npad = std::max<i64>((i64)ts.tile_slice[nt] * 32, 32);
  ts.perm.assign(n, -1);
  ts.row_lvl.assign(ts.npad, 255);
  {
    i64 p = 0;
    for (i64 i = 0; i < n; i++) if (!smoothed(i)) ts.perm[i] = (i32)p++;
  }
  std::vector<i32> lvl(n, 0);
  int maxlev = 0;
  parallel_for(nt, [&](i64 lo, i64 hi) {
    std::vector<std::pair<i32, i32>> key;
    for (i64 q = lo; q < hi; q++) {
      const i64 t = order[q];
      int nl = 0;
      // members are in sweep order: a row's in-tile dependencies precede it
      for (i64 m = mptr[t]; m < mptr[t + 1]; m++) {
        const i64 i = mem[m];
        const i64 ri = rank_of(i);
        i32 l = 0;
        for (i64 k = A.rowptr[i]; k < A.rowptr[i + 1]; k++) {
          const i32 j = A.col[k];
          if (j == i || cl[j] != t) continue;
          if (rank_of(j) < ri) l = std::max(l, lvl[j] + 1);
        }
        lvl[i] = l;
        nl = std::max(nl, l + 1);
      }
      ts.tile_nlev[q] = nl;
      key.clear();
      for (i64 m = mptr[t]; m < mptr[t + 1]; m++) key.emplace_back(lvl[mem[m]], (i32)(m - mptr[t]));
      std::sort(key.begin(), key.end());
      const i64 r0 = (i64)ts.tile_slice[q] * 32;
      for (size_t k = 0; k < key.size(); k++) {
        const i64 i = mem[mptr[t] + key[k].second];
        ts.perm[i] = (i32)(r0 + (i64)k);
        ts.row_lvl[r0 + (i64)k] = (uint8_t)key[k].first;
      }
    }
  }, 256);
  for (i64 q = 0; q < nt; q++) maxlev = std::max(maxlev, (int)ts.tile_nlev[q]);
  ts.max_local_levels = maxlev;
  if (maxlev > 250) return;   // local level does not fit the byte array (cannot happen with <= 128 rows)

  // ---- 6. wait lists in the new tile numbering
  ts.pred_ptr.assign(nt + 1, 0);
  ts.succ_ptr.assign(nt + 1, 0);
  for (i64 q = 0; q < nt; q++) {
    const i64 t = order[q];
    ts.pred_ptr[q + 1] = ts.pred_ptr[q] + (pptr[t + 1] - pptr[t]);
    ts.succ_ptr[q + 1] = ts.succ_ptr[q] + (sptr[t + 1] - sptr[t]);
  }
  ts.pred.resize(ts.pred_ptr[nt]);
  ts.succ.resize(ts.succ_ptr[nt]);
  for (i64 q = 0; q < nt; q++) {
    const i64 t = order[q];
    for (i64 k = 0; k < pptr[t + 1] - pptr[t]; k++) ts.pred[ts.pred_ptr[q] + k] = newid[pl[pptr[t] + k]];
    for (i64 k = 0; k < sptr[t + 1] - sptr[t]; k++) ts.succ[ts.succ_ptr[q] + k] = newid[sl[sptr[t] + k]];
  }
  ts.ok = true;
}

// ---------------------------------------------------------------------------------------------------------------------------
// Cluster hint for matrices numbered like a structured grid (lexicographic x-fastest numbering of an N1 x N2 x N3 box -- what a
// structured mesher, NGSolve's MakeStructured3DMesh or any "i + N1 (j + N2 k)" assembly loop produces).  Purely algebraic: the line
// length N1 is the period of the rows that have no entry at column i-1, the plane size N1*N2 the first row >= N1 without an entry at
// column i-N1; both are verified on every row.  The hint groups the smoothed rows into near-cubic boxes (edges as equal as the free
// extents allow); build_tile_schedule turns them into tiles and splits whatever is not convex in the sweep DAG, so a wrong guess costs
// depth, never correctness.  Returns false if the pattern is not of that kind (the caller falls back to the pairwise clustering).
// ---------------------------------------------------------------------------------------------------------------------------
bool grid_box_hint(const HostBsr &A, const std::vector<uint8_t> &mask, int max_rows, std::vector<i32> &hint, i64 dims[3])
{
  const i64 n = A.nrows;
  if (n < 64) return false;
  auto has = [&](i64 i, i64 j) {
    if (j < 0) return false;
    const i32 *b = A.col.data() + A.rowptr[i], *e = A.col.data() + A.rowptr[i + 1];
    const i32 *q = std::lower_bound(b, e, (i32)j);
    return q != e && *q == (i32)j;
  };
  // line length: rows 1 .. N1-1 are coupled to their predecessor, row N1 is not
  i64 n1 = 0;
  for (i64 i = 1; i < n; i++) if (!has(i, i - 1)) { n1 = i; break; }
  if (n1 < 2) n1 = n;                     // a single line (1D) -- or no x-coupling at all
  if (n % n1) return false;
  i64 n12 = 0;
  for (i64 i = n1; i < n; i += n1) if (!has(i, i - n1)) { n12 = i; break; }
  if (n12 == 0) n12 = n;                  // a single plane (2D)
  if (n12 % n1 || n % n12) return false;
  const i64 N1 = n1, N2 = n12 / n1, N3 = n / n12;
  // verify on all rows: every entry is a grid neighbour (|dx|,|dy|,|dz| <= 1) and the line / plane starts are where they should be
  std::atomic<int> bad{0};
  parallel_for(n, [&](i64 lo, i64 hi) {
    for (i64 i = lo; i < hi && !bad.load(std::memory_order_relaxed); i++) {
      const i64 x = i % N1, y = (i / N1) % N2, z = i / n12;
      for (i64 k = A.rowptr[i]; k < A.rowptr[i + 1]; k++) {
        const i64 j = A.col[k];
        const i64 dx = j % N1 - x, dy = (j / N1) % N2 - y, dz = j / n12 - z;
        if (dx < -1 || dx > 1 || dy < -1 || dy > 1 || dz < -1 || dz > 1) { bad = 1; return; }
      }
    }
  }, 1 << 16);
  if (bad) return false;
  dims[0] = N1; dims[1] = N2; dims[2] = N3;
  // extents of the smoothed rows
  const bool hm = !mask.empty();
  i64 lo[3] = {N1, N2, N3}, hi[3] = {-1, -1, -1};
  for (i64 i = 0; i < n; i++) {
    if (hm && !mask[i]) continue;
    const i64 c[3] = {i % N1, (i / N1) % N2, i / n12};
    for (int d = 0; d < 3; d++) { lo[d] = std::min(lo[d], c[d]); hi[d] = std::max(hi[d], c[d]); }
  }
  if (hi[0] < 0) return false;
  i64 ext[3], nb[3];
  int nd = 0;
  for (int d = 0; d < 3; d++) { ext[d] = hi[d] - lo[d] + 1; nd += ext[d] > 1; }
  if (nd == 0) return false;
  // box edge: the largest e with e^nd <= max_rows; extents are cut into ceil(ext / e) nearly equal parts
  i64 e = 1;
  for (;;) {
    i64 v = 1;
    for (int d = 0; d < nd; d++) v *= (e + 1);
    if (v > max_rows) break;
    e++;
  }
  for (int d = 0; d < 3; d++) nb[d] = ext[d] > 1 ? (ext[d] + e - 1) / e : 1;
  if ((double)nb[0] * nb[1] * nb[2] > 2.0e9) return false;
  hint.assign(n, -1);
  parallel_for(n, [&](i64 a, i64 b) {
    for (i64 i = a; i < b; i++) {
      if (hm && !mask[i]) continue;
      const i64 c[3] = {i % N1, (i / N1) % N2, i / n12};
      i64 bx[3];
      for (int d = 0; d < 3; d++) bx[d] = (c[d] - lo[d]) * nb[d] / ext[d];
      hint[i] = (i32)((bx[2] * nb[1] + bx[1]) * nb[0] + bx[0]);
    }
  }, 1 << 16);
  return true;
}

i64 check_tile_schedule(const HostBsr &A, const std::vector<uint8_t> &mask, const std::vector<i32> &sweep_rank, const TileSchedule &ts)
{
  const i64 n = A.nrows;
  const bool hm = !mask.empty(), natural = sweep_rank.empty();
  auto rank_of = [&](i64 i) -> i64 { return natural ? i : (i64)sweep_rank[i]; };
  // tile of a (new) row
  std::vector<i32> tile_of_row(ts.npad, -1);
  for (i64 q = 0; q < ts.ntiles; q++)
    for (i64 r = (i64)ts.tile_slice[q] * 32; r < (i64)ts.tile_slice[q + 1] * 32; r++) tile_of_row[r] = (i32)q;
  i64 bad = 0;
  std::vector<uint8_t> seen(ts.npad, 0);
  for (i64 i = 0; i < n; i++) {
    const i64 pi = ts.perm[i];
    if (pi < 0 || pi >= ts.npad || seen[pi]) { bad++; continue; }
    seen[pi] = 1;
    const bool sm = !hm || mask[i];
    if (!sm) { bad += (pi >= ts.nonfree_pad) ? 1 : 0; continue; }
    if (pi < ts.nonfree_pad || ts.row_lvl[pi] == 255) { bad++; continue; }
    const i32 ti = tile_of_row[pi];
    for (i64 k = A.rowptr[i]; k < A.rowptr[i + 1]; k++) {
      const i64 j = A.col[k];
      if (j == i || (hm && !mask[j])) continue;
      const i64 pj = ts.perm[j];
      const i32 tj = tile_of_row[pj];
      if (rank_of(j) < rank_of(i)) {
        // dependency: must have a lower row number, and be either in an awaited earlier tile or on a lower local level
        if (!(pj < pi)) bad++;
        if (tj == ti) bad += (ts.row_lvl[pj] < ts.row_lvl[pi]) ? 0 : 1;
        else {
          bad += (tj < ti) ? 0 : 1;
          bool listed = false;
          for (i64 q = ts.pred_ptr[ti]; q < ts.pred_ptr[ti + 1]; q++) listed |= ts.pred[q] == tj;
          bad += listed ? 0 : 1;
          bool listed2 = false;
          for (i64 q = ts.succ_ptr[tj]; q < ts.succ_ptr[tj + 1]; q++) listed2 |= ts.succ[q] == ti;
          bad += listed2 ? 0 : 1;
        }
      } else if (!(pj > pi)) bad++;
    }
  }
  return bad;
}

}  // namespace ngb
