// kernels_ctile.cuh -- triangular half-sweep of the Gauss-Seidel smoother on the two-level tile schedule of tiles.hpp, ONE CTA PER TILE.
//
// Why: the row-level sync-free sweep (k_gs_tri) pays one L2 store -> poll hop per dependency level of the ROW DAG (930 levels at 311^3,
// ~2.6 us each under load = 2.4 ms for 0.57 ms worth of bytes).  Here a tile of up to MAXS*32 rows (512: an 8x8x8 box of a grid) is swept
// by one CTA: its in-tile dependencies travel through shared memory (one __syncthreads per tile-local level), only values of OTHER tiles
// come from L2 -- the critical path is the depth of the TILE DAG (~3 n^(1/3) / 8).
//
// Data movement (sm_100a): the tile's slab of the SELL matrix (slices of one tile are contiguous: one run of column indices, one run of
// values) is fetched by ONE elected thread with two 1-D bulk copies (cp.async.bulk.shared::cluster.global, SASS UBLKCP) that complete on an
// mbarrier -- ~45 KB in flight per CTA without a single register staging the matrix; the copy is issued before the CTA starts waiting for
// its dependencies, so HBM latency is off the dependency chain.  Several CTAs per SM keep > 100 KB in flight per SM.
//
// Protocol: `out` is sentinel-filled (all-ones NaN) before the launch; a value of another tile is polled until it is no longer the
// sentinel -- the data is the flag, no fence on the critical path.  A per-tile hint flag (plain store, NOT a release) only tells waiting
// CTAs when polling the data is worth it, so that a CTA far down the schedule costs one sector per poll round instead of ~100.
// Tiles are dealt to the CTAs round-robin in schedule order (tile-DAG level major); every CTA walks its tiles in that order, so the
// earliest unfinished tile is always being worked on with all its dependencies done: no dead-lock while the grid is co-resident
// (cooperative launch, see launch_resident in amg.cu).
//
// Same contract as k_gs_tri:  out = (ADD_SELF ? self : 0) + dinv * (rin - T out),   rout = rin - (T + diag) * delta   [WRITE_R].
// Scalar matrices (B = 1).
#pragma once
#include "kernels.cuh"

namespace ngb {

struct CTileParams {
  i32 ntiles;
  int backward;
  const i32 *tile_slice;     // ntiles + 1: first slice of every tile (schedule order)
  const i32 *tile_nlev;      // tile-local dependency levels
  const uint8_t *row_lvl;    // per row: tile-local level, 255 = padding
  const i64 *dep_ptr;        // tiles to wait for: predecessors (forward) / successors (backward)
  const i32 *dep;
  int *done;                 // hint flag per tile, zeroed before the launch
  unsigned sleep_ns;         // back-off of the hint polls
  unsigned repoll_ns;        // back-off of the data re-polls (stragglers)
  int pollmode;              // flavour of the polling load (ld_poll)
  int cap_slots;             // capacity of the shared-memory slab in SELL slots (a slot = 32 entries)
  int *err;                  // watchdog
  unsigned long long *trace; // debug: 8 words per tile {begin, hints passed, slab landed, gathered, levels done, end, smid|cta<<32, nlev|slices<<16} (NULL = off)
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, int count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}" ::"r"(bar), "r"(parity)
      : "memory");
}
// 1-D bulk copy global -> shared (TMA engine, no tensor map); bytes and both addresses are multiples of 16
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar)
{
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes),
               "r"(bar)
               : "memory");
}
__device__ __forceinline__ int ld_relaxed_i32(const int *p)
{
  int v;
  asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed_i32(int *p, int v) { asm volatile("st.relaxed.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }

constexpr int CTILE_HDR = 128;   // bytes in front of the slab: the mbarrier
__host__ __device__ inline size_t ctile_smem_bytes(int maxs, int cap_slots) { return (size_t)CTILE_HDR + (size_t)maxs * 32 * 8 + (size_t)cap_slots * 32 * 12; }

template <int NT, int MAXS, bool ADD_SELF, bool WRITE_R>
__global__ void __launch_bounds__(NT, (MAXS >= 16 ? 3 : 4)) k_gs_ctile(SellView T, const double *__restrict__ diag, const double *__restrict__ dinv,
                                                const double *rin, const double *__restrict__ self, double *out, double *rout,
                                                CTileParams p)
{
  constexpr int NW = NT / 32;
  static_assert(MAXS % NW == 0, "every warp owns the same number of slices");
  constexpr int NR = MAXS / NW;                 // rows per thread: tile-local rows tid, tid + NT, ...
  constexpr int CH = (NR <= 1) ? 8 : 4;         // slots gathered per round and row
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double *xs = (double *)(smem_raw + CTILE_HDR);                      // the tile's part of `out`
  double *vals_s = xs + MAXS * 32;                                    // slab: values, then column indices
  i32 *cols_s = (i32 *)(vals_s + (size_t)p.cap_slots * 32);
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const uint32_t bar = smem_u32(smem_raw);
  if (tid == 0) { mbar_init(bar, 1); mbar_fence_init(); }
  __syncthreads();
  uint32_t phase = 0;
  for (i32 q = blockIdx.x; q < p.ntiles; q += gridDim.x) {
    const i32 t = p.backward ? (p.ntiles - 1 - q) : q;
    unsigned long long *tr = (p.trace && tid == 0) ? p.trace + (size_t)t * 8 : nullptr;
    if (tr) tr[0] = gtimer();
    const i32 s0 = p.tile_slice[t];
    const int ns = p.tile_slice[t + 1] - s0;
    const i32 r0 = s0 * 32;
    const unsigned nrow = (unsigned)ns * 32u;
    const i64 base0 = T.slice_ptr[s0];
    const int nslots = (int)(T.slice_ptr[s0 + ns] - base0);
    const bool has = nslots > 0;
    if (has && tid == 0) {
      mbar_expect_tx(bar, (uint32_t)nslots * 32u * 12u);
      bulk_g2s(smem_u32(vals_s), T.val + base0 * 32, (uint32_t)nslots * 256u, bar);
      bulk_g2s(smem_u32(cols_s), T.col + base0 * 32, (uint32_t)nslots * 128u, bar);
    }
    // ---- per-row data that does not depend on `out`
    double acc[NR], dv[NR], sv[NR], dg[NR];
    int lv[NR], sb[NR], wd[NR];
#pragma unroll
    for (int j = 0; j < NR; j++) {
      const int sl = w + j * NW;
      lv[j] = 256; sb[j] = 0; wd[j] = 0; acc[j] = 0.0; dv[j] = 0.0; sv[j] = 0.0; dg[j] = 0.0;
      if (sl < ns) {
        const i64 slice = (i64)s0 + sl, row = slice * 32 + lane;
        const i64 b = T.slice_ptr[slice];
        sb[j] = (int)(b - base0);
        wd[j] = (int)(T.slice_ptr[slice + 1] - b);
        lv[j] = p.row_lvl[row];
        acc[j] = rin[row];
        dv[j] = dinv[row];
        if (ADD_SELF) sv[j] = self[row];
        if (WRITE_R) dg[j] = diag[row];
      }
    }
    // ---- hint flags of the tiles this one depends on (one thread per dependency), then the slab
    {
      const i64 d0 = p.dep_ptr[t], d1 = p.dep_ptr[t + 1];
      for (i64 k = d0 + tid; k < d1; k += NT) {
        const int *f = p.done + p.dep[k];
        unsigned spins = 0;
        while (ld_relaxed_i32(f) == 0) {
          if (p.sleep_ns) __nanosleep(p.sleep_ns);
          if (spin_fail(spins, p.err)) break;
        }
      }
    }
    if (tr) tr[1] = gtimer();
    if (has) { mbar_wait(bar, phase); phase ^= 1u; }
    __syncthreads();
    if (tr) tr[2] = gtimer();
    // ---- couplings to rows of other tiles: poll the data itself (sentinel), a round of CH slots per owned row at a time
    int maxw = 0;
#pragma unroll
    for (int j = 0; j < NR; j++) maxw = max(maxw, wd[j]);
    for (int k0 = 0; k0 < maxw; k0 += CH) {
      i32 cc[NR][CH];
      double xv[NR][CH];
#pragma unroll
      for (int j = 0; j < NR; j++)
#pragma unroll
        for (int e = 0; e < CH; e++) {
          const int k = k0 + e;
          i32 c = (k < wd[j]) ? cols_s[(sb[j] + k) * 32 + lane] : -1;
          if (c >= 0 && (unsigned)(c - r0) < nrow) c = -1;       // in-tile: served from shared memory below
          cc[j][e] = c;
          xv[j][e] = (c >= 0) ? ld_poll(out + c, p.pollmode) : 0.0;
        }
#pragma unroll
      for (int j = 0; j < NR; j++)
#pragma unroll
        for (int e = 0; e < CH; e++)
          if (cc[j][e] >= 0) {
            unsigned spins = 0;
            while (is_sentinel(xv[j][e])) {
              if (p.repoll_ns) __nanosleep(p.repoll_ns);
              xv[j][e] = ld_poll(out + cc[j][e], p.pollmode);
              if (spin_fail(spins, p.err)) break;
            }
            acc[j] = fma(-vals_s[(sb[j] + k0 + e) * 32 + lane], xv[j][e], acc[j]);
          }
    }
    if (tr) tr[3] = gtimer();
    // ---- the tile itself, local level by local level (ascending forward, descending backward)
    const int nlev = p.tile_nlev[t];
    for (int it = 0; it < nlev; it++) {
      const int s = p.backward ? (nlev - 1 - it) : it;
#pragma unroll
      for (int j = 0; j < NR; j++) {
        if (lv[j] == s) {
          double a = acc[j], a2 = 0.0;
          // rounds of 8 slots: all shared-memory loads of a round are issued before the first FMA (two FMA chains)
          for (int k0 = 0; k0 < wd[j]; k0 += 8) {
            i32 c[8];
            double v[8], x[8];
#pragma unroll
            for (int e = 0; e < 8; e++) {
              const bool in = k0 + e < wd[j];
              c[e] = in ? cols_s[(sb[j] + k0 + e) * 32 + lane] : -1;
              v[e] = in ? vals_s[(sb[j] + k0 + e) * 32 + lane] : 0.0;
            }
#pragma unroll
            for (int e = 0; e < 8; e++) {
              const unsigned lc = (unsigned)(c[e] - r0);
              x[e] = (c[e] >= 0 && lc < nrow) ? xs[lc] : 0.0;
            }
#pragma unroll
            for (int e = 0; e < 8; e += 2) { a = fma(-v[e], x[e], a); a2 = fma(-v[e + 1], x[e + 1], a2); }
          }
          a += a2;
          const double d = dv[j] * a;
          const double r = ADD_SELF ? sv[j] + d : d;
          const int lr = tid + j * NT;
          xs[lr] = r;
          __stcg(out + (i64)r0 + lr, r);
          if (WRITE_R) rout[(i64)r0 + lr] = fma(-dg[j], d, a);
        }
      }
      __syncthreads();
    }
    if (tr) tr[4] = gtimer();
    // padding rows of the tile (no level): never updated, but `out` must not keep the sentinel
#pragma unroll
    for (int j = 0; j < NR; j++)
      if (lv[j] == 255) {
        const int lr = tid + j * NT;
        __stcg(out + (i64)r0 + lr, ADD_SELF ? sv[j] : 0.0);
        if (WRITE_R) rout[(i64)r0 + lr] = acc[j];
      }
    if (tid == 0) st_relaxed_i32(p.done + t, 1);
    if (tr) {
      unsigned smid;
      asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
      tr[5] = gtimer();
      tr[6] = (unsigned long long)smid | ((unsigned long long)blockIdx.x << 32);
      tr[7] = (unsigned long long)nlev | ((unsigned long long)ns << 16);
    }
    __syncthreads();     // the slab and xs are reused by the next tile
  }
}

}  // namespace ngb
