// kernels_tile.cuh -- EXPERIMENTAL (flag ngs_amg_b200_tile_sweep, default off): triangular half-sweep of the Gauss-Seidel smoother on
// the two-level tile schedule of tiles.hpp.  One warp executes one tile (MAXS 32-row slices): it waits for the flags of the tiles it
// depends on, folds every out-of-tile coupling into its accumulators (those values are final), then walks the tile-local dependency
// levels with the in-tile couplings served from shared memory -- no global round trip inside a tile -- and finally publishes the tile
// with one release store.  Critical path = depth of the TILE DAG (x one flag hop) instead of the depth of the row DAG.
// Same contract as k_gs_tri (kernels.cuh):  out = (ADD_SELF ? self : 0) + dinv * (rin - T out),  rout = rin - (T + diag) * delta.
// Scalar matrices (B = 1) only.  NOT yet validated on hardware -- written in round 1 after the GPU budget was spent; see DESIGN.md §9.
#pragma once
#include "kernels.cuh"

namespace ngb {

struct TileParams {
  i32 ntiles;
  int backward;
  const i32 *tile_slice;     // ntiles + 1: first slice of every tile (schedule order)
  const i32 *tile_nlev;      // tile-local dependency levels
  const uint8_t *row_lvl;    // per row: local level, 255 = padding
  const i64 *dep_ptr;        // tiles to wait for: predecessors (forward) / successors (backward)
  const i32 *dep;
  int *done;                 // one flag per tile, zeroed before the launch
  unsigned sleep_ns;
  int *err;                  // watchdog
};

__device__ __forceinline__ int ld_acquire_i32(const int *p)
{
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_i32(int *p, int v)
{
  asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

template <int MAXS, bool ADD_SELF, bool WRITE_R>
__global__ void __launch_bounds__(256) k_gs_tile(SellView T, const double *__restrict__ diag, const double *__restrict__ dinv,
                                                const double *rin, const double *__restrict__ self, double *out, double *rout,
                                                TileParams p)
{
  constexpr int PRE = 8;                       // slots cached in registers per slice; wider rows re-read the rest from global
  __shared__ double sh[8][MAXS * 32];
  const int lane = threadIdx.x & 31;
  double *mysh = sh[threadIdx.x >> 5];
  const i64 gw = ((i64)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const i64 nw = ((i64)gridDim.x * blockDim.x) >> 5;
  for (i64 q = gw; q < p.ntiles; q += nw) {
    const i32 t = p.backward ? (i32)(p.ntiles - 1 - q) : (i32)q;
    const i32 s0 = p.tile_slice[t];
    const int ns = p.tile_slice[t + 1] - s0;
    const i64 r0 = (i64)s0 * 32, r1 = r0 + (i64)ns * 32;
    // ---- everything that does not depend on `out`
    double acc[MAXS], dv[MAXS], dg[MAXS], sv[MAXS];
    int lv[MAXS], wd[MAXS];
    i64 base[MAXS];
    i32 pc[MAXS][PRE];
    double pv[MAXS][PRE];
#pragma unroll
    for (int k = 0; k < MAXS; k++) {
      lv[k] = 255; wd[k] = 0; base[k] = 0; acc[k] = 0.0; dv[k] = 0.0; dg[k] = 0.0; sv[k] = 0.0;
#pragma unroll
      for (int e = 0; e < PRE; e++) { pc[k][e] = -1; pv[k][e] = 0.0; }
      if (k < ns) {
        const i64 slice = s0 + k, row = slice * 32 + lane;
        base[k] = T.slice_ptr[slice];
        wd[k] = (int)(T.slice_ptr[slice + 1] - base[k]);
        lv[k] = p.row_lvl[row];
        acc[k] = rin[row];
        dv[k] = dinv[slice * 32 + lane];
        if (WRITE_R) dg[k] = diag[slice * 32 + lane];
        if (ADD_SELF) sv[k] = self[row];
#pragma unroll
        for (int e = 0; e < PRE; e++)
          if (e < wd[k]) { pc[k][e] = T.col[(base[k] + e) * 32 + lane]; pv[k][e] = T.val[(base[k] + e) * 32 + lane]; }
      }
    }
    // ---- wait for the tiles this one depends on (one lane per dependency)
    {
      const i64 d0 = p.dep_ptr[t], d1 = p.dep_ptr[t + 1];
      for (i64 k = d0 + lane; k < d1; k += 32) {
        const int *f = p.done + p.dep[k];
        unsigned spins = 0;
        while (ld_acquire_i32(f) == 0) {
          if (p.sleep_ns) __nanosleep(p.sleep_ns);
          if (spin_fail(spins, p.err, 1u << 24)) break;
        }
      }
      __syncwarp();
    }
    // ---- couplings to rows outside the tile: their values are final now
#pragma unroll
    for (int k = 0; k < MAXS; k++) {
      if (k < ns) {
#pragma unroll
        for (int e = 0; e < PRE; e++) {
          const i32 c = pc[k][e];
          if (c >= 0 && (c < r0 || c >= r1)) { acc[k] = fma(-pv[k][e], __ldcg(out + c), acc[k]); pc[k][e] = -1; }
        }
        for (int e = PRE; e < wd[k]; e++) {
          const i32 c = T.col[(base[k] + e) * 32 + lane];
          if (c >= 0 && (c < r0 || c >= r1)) acc[k] = fma(-T.val[(base[k] + e) * 32 + lane], __ldcg(out + c), acc[k]);
        }
      }
    }
    // ---- the tile itself, local level by local level (ascending forward, descending backward)
    const int nlev = p.tile_nlev[t];
    for (int it = 0; it < nlev; it++) {
      const int s = p.backward ? (nlev - 1 - it) : it;
#pragma unroll
      for (int k = 0; k < MAXS; k++) {
        if (k < ns && lv[k] == s) {
          double a = acc[k];
#pragma unroll
          for (int e = 0; e < PRE; e++) {
            const i32 c = pc[k][e];
            if (c >= 0) a = fma(-pv[k][e], mysh[c - r0], a);
          }
          for (int e = PRE; e < wd[k]; e++) {
            const i32 c = T.col[(base[k] + e) * 32 + lane];
            if (c >= r0 && c < r1) a = fma(-T.val[(base[k] + e) * 32 + lane], mysh[c - r0], a);
          }
          const double d = dv[k] * a;
          const double res = ADD_SELF ? sv[k] + d : d;
          const i64 row = r0 + (i64)k * 32 + lane;
          mysh[k * 32 + lane] = res;
          out[row] = res;
          if (WRITE_R) rout[row] = fma(-dg[k], d, a);
        }
      }
      __syncwarp();
    }
    // ---- publish: every lane's stores are fenced, then one release store of the tile flag
    __threadfence();
    __syncwarp();
    if (lane == 0) st_release_i32(p.done + t, 1);
  }
}

// non-smoothed prefix rows (Dirichlet / ghost rows): never updated
template <bool ADD_SELF, bool WRITE_R>
__global__ void k_gs_tile_prefix(i64 nonfree, const double *rin, const double *__restrict__ self, double *out, double *rout)
{
  const i64 row = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= nonfree) return;
  out[row] = ADD_SELF ? self[row] : 0.0;
  if (WRITE_R) rout[row] = rin[row];
}

}  // namespace ngb
