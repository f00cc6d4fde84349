// tiles.hpp -- two-level schedule of a sequential Gauss-Seidel sweep: the rows are grouped into compact TILES (clusters of <= 128
// graph-neighbouring rows); a tile is processed by one warp, its rows in the order of their tile-local dependency levels (shared
// memory / registers carry the dependencies), and tiles synchronise through one flag per tile.  The number of global
// store -> load hops on the critical path of a sweep drops from the depth of the row DAG (~3 n^(1/3) in the natural ordering of a
// grid) to the depth of the TILE DAG (~3 n^(1/3) / tile edge), which is what bounds the sync-free row-level sweep today.
// Any topological order of the row DAG reproduces the reference's sequential sweep; executing whole tiles atomically is
// such an order iff the tile graph is acyclic -- cyclic groups of tiles are merged.
// Host side only (testable without a device); the kernel that consumes the schedule is k_gs_tile (kernels_tile.cuh).
#pragma once
#include "common.hpp"

namespace ngb {

struct TileSchedule {
  bool ok = false;                 // false: no valid tiling (a merged tile would exceed max_rows) -> caller keeps the row-level sweep
  i64 n = 0, npad = 0, nonfree_pad = 0;
  std::vector<i32> perm;           // original row -> row in the tile-major numbering
  i64 ntiles = 0;
  std::vector<i32> tile_slice;     // ntiles + 1: first 32-row slice of every tile, tiles in schedule order (tile DAG level-major)
  std::vector<i32> tile_nlev;      // number of tile-local dependency levels
  std::vector<uint8_t> row_lvl;    // per new row: tile-local level, 255 = padding / non-smoothed
  std::vector<i64> pred_ptr, succ_ptr;   // ntiles + 1
  std::vector<i32> pred, succ;     // tiles a tile waits for in the forward / backward sweep
  int tile_depth = 0;              // levels of the tile DAG  (critical path in tiles)
  int max_local_levels = 0;
  i64 merged_tiles = 0;            // tiles that had to be merged because of cyclic dependencies
};

// A: level matrix (original numbering); mask: rows that are smoothed (empty = all); sweep_rank: position of each row in the sweep
// (empty = row number); max_rows: capacity of a tile (multiple of 32, <= 1024; the warp-per-tile kernel handles <= 64); rounds: pairwise clustering rounds (tiles of <= 2^rounds rows).
// cluster_hint (optional): caller-supplied cluster id per row (-1 = not smoothed) instead of the pairwise clustering -- e.g. boxes of a structured grid.
void build_tile_schedule(const HostBsr &A, const std::vector<uint8_t> &mask, const std::vector<i32> &sweep_rank, int rounds, int max_rows,
                         TileSchedule &out, const std::vector<i32> *cluster_hint = nullptr);

// self-check: every dependency of every row is scheduled before the row (other tile listed as predecessor and earlier in the order, or
// same tile and lower local level).  Returns the number of violations.
i64 check_tile_schedule(const HostBsr &A, const std::vector<uint8_t> &mask, const std::vector<i32> &sweep_rank, const TileSchedule &ts);

// cluster hint for matrices numbered like a structured grid (algebraic detection of the line / plane strides); false = not such a matrix
bool grid_box_hint(const HostBsr &A, const std::vector<uint8_t> &mask, int max_rows, std::vector<i32> &hint, i64 dims[3]);

i64 cluster_rows(const HostBsr &A, const uint8_t *keep, int rounds, double soc_thresh, std::vector<i32> &cluster, bool forward = false);

}  // namespace ngb
