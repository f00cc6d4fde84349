// tiles_abi.cpp -- host-only C entry points of the tile scheduler (tiles.hpp), for the CPU tests and offline analysis.
#include "../../include/ngsamg_b200.h"
#include "tiles.hpp"

using namespace ngb;

struct ngsamg_b200_tiles { TileSchedule ts; i64 violations = 0; };

static thread_local std::string g_tile_err;

extern "C" {

const char *ngsamg_b200_tiles_last_error(void) { return g_tile_err.c_str(); }

static int tile_schedule_impl(const ngsamg_csr *A, const uint8_t *smoothed_mask, const int32_t *sweep_rank, const int32_t *cluster_hint, int rounds,
                              int max_rows, ngsamg_b200_tiles **out, int64_t *info);

int ngsamg_b200_tile_schedule_begin(const ngsamg_csr *A, const uint8_t *smoothed_mask, const int32_t *sweep_rank, int rounds, int max_rows,
                                    ngsamg_b200_tiles **out, int64_t *info /* 9 values, see header */)
{
  return tile_schedule_impl(A, smoothed_mask, sweep_rank, nullptr, rounds, max_rows, out, info);
}

int ngsamg_b200_tile_schedule_hinted(const ngsamg_csr *A, const uint8_t *smoothed_mask, const int32_t *sweep_rank, const int32_t *cluster, int max_rows,
                                     ngsamg_b200_tiles **out, int64_t *info)
{
  if (!cluster) { g_tile_err = "null cluster hint"; return 1; }
  return tile_schedule_impl(A, smoothed_mask, sweep_rank, cluster, 0, max_rows, out, info);
}

static int tile_schedule_impl(const ngsamg_csr *A, const uint8_t *smoothed_mask, const int32_t *sweep_rank, const int32_t *cluster_hint, int rounds,
                              int max_rows, ngsamg_b200_tiles **out, int64_t *info)
{
  try {
    if (!A || !out || !A->rowptr) throw Error("null argument");
    HostBsr h;
    h.nrows = A->nrows; h.ncols = A->ncols; h.bh = A->bh; h.bw = A->bw;
    h.rowptr.assign(A->rowptr, A->rowptr + A->nrows + 1);
    const i64 nnz = h.rowptr.back();
    h.col.assign(A->col, A->col + nnz);
    h.val.assign(A->val, A->val + nnz * h.bs());
    std::vector<uint8_t> mask;
    if (smoothed_mask) mask.assign(smoothed_mask, smoothed_mask + A->nrows);
    std::vector<i32> rank;
    if (sweep_rank) rank.assign(sweep_rank, sweep_rank + A->nrows);
    auto r = std::make_unique<ngsamg_b200_tiles>();
    std::vector<i32> hint;
    if (cluster_hint) hint.assign(cluster_hint, cluster_hint + A->nrows);
    bool hinted = cluster_hint != nullptr;
    if (!hinted && rounds < 0) {
      // rounds < 0: what the library does at setup -- box-shaped clusters if the matrix is numbered like a structured grid, else the
      // pairwise clustering with as many rounds as the capacity asks for
      i64 dims[3];
      hinted = grid_box_hint(h, mask, max_rows, hint, dims);
      rounds = 5;
      while ((1 << rounds) < max_rows) rounds++;
    }
    build_tile_schedule(h, mask, rank, rounds, max_rows, r->ts, hinted ? &hint : nullptr);
    if (r->ts.ok) r->violations = check_tile_schedule(h, mask, rank, r->ts);
    if (info) {
      const TileSchedule &t = r->ts;
      const int64_t v[9] = {t.ok ? 1 : 0, t.ntiles, t.npad, t.nonfree_pad, t.tile_depth, t.max_local_levels, t.merged_tiles, r->violations,
                            (int64_t)t.pred.size()};
      std::memcpy(info, v, sizeof(v));
    }
    *out = r.release();
  } catch (const std::exception &e) { g_tile_err = e.what(); return 1; }
  return 0;
}

int ngsamg_b200_tile_schedule_fetch(ngsamg_b200_tiles *m, int32_t *perm, int32_t *tile_slice, int32_t *tile_nlev, uint8_t *row_lvl,
                                    int64_t *pred_ptr, int32_t *pred)
{
  if (!m) { g_tile_err = "null handle"; return 1; }
  const TileSchedule &t = m->ts;
  if (t.ok) {
    if (perm) std::memcpy(perm, t.perm.data(), sizeof(i32) * t.perm.size());
    if (tile_slice) std::memcpy(tile_slice, t.tile_slice.data(), sizeof(i32) * t.tile_slice.size());
    if (tile_nlev) std::memcpy(tile_nlev, t.tile_nlev.data(), sizeof(i32) * t.tile_nlev.size());
    if (row_lvl) std::memcpy(row_lvl, t.row_lvl.data(), t.row_lvl.size());
    if (pred_ptr) std::memcpy(pred_ptr, t.pred_ptr.data(), sizeof(i64) * t.pred_ptr.size());
    if (pred) std::memcpy(pred, t.pred.data(), sizeof(i32) * t.pred.size());
  }
  delete m;
  return 0;
}

}  // extern "C"
