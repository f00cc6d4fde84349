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

// everything a CTA needs to know about a tile before it touches the matrix, in one 32-byte record (one broadcast load per warp);
// one array per sweep direction (the slab of L or U, predecessors or successors)
struct __align__(16) CTileMeta {
  i64 base;      // first SELL slot of the tile's slab in this part (L forward, U backward)
  i32 s0;        // first 32-row slice of the tile
  i32 ns;        // slices
  i32 nslots;    // slots of the slab
  i32 nlev;      // tile-local dependency levels (low 16 bits) | rows of the tile that are not padding << 16
  i32 d0;        // first entry of the tile's wait list in `dep`
  i32 nd;        // tiles to wait for
};

struct CTileParams {
  i32 ntiles;
  int backward;
  const CTileMeta *meta;     // per tile, this direction
  const uint8_t *row_lvl;    // per row: tile-local level, 255 = padding
  const i32 *dep;            // wait lists: predecessors (forward) / successors (backward)
  int *done;                 // hint flag per tile, zeroed before the launch
  unsigned sleep_ns;         // back-off of the hint polls
  unsigned repoll_ns;        // back-off of the data re-polls (stragglers)
  int pollmode;              // flavour of the polling load (ld_poll)
  int cap_slots;             // capacity of ONE shared-memory slab in SELL slots (a slot = 32 entries)
  int *err;                  // watchdog
  unsigned long long *trace; // debug: 8 words per tile {begin, hints passed, slab landed, gathered, levels done, end, smid|cta<<16|nlev<<40|slices<<52, first level done} (NULL = off)
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

// shared-memory accesses by 32-bit shared-window address: keeps the address arithmetic in plain integer registers (through C++ pointers
// the compiler re-derives the shared window base from SR_CgaCtaId in front of every access of the level loop -- S2R on the critical path)
__device__ __forceinline__ double lds_f64(uint32_t a)
{
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ i32 lds_i32(uint32_t a)
{
  i32 v;
  asm volatile("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ void sts_f64(uint32_t a, double v) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(v) : "memory"); }
__device__ __forceinline__ double ld_poll_relaxed(const double *p)
{
  double v;
  asm volatile("ld.relaxed.gpu.global.f64 %0, [%1];" : "=d"(v) : "l"(p));
  return v;
}

__device__ __forceinline__ unsigned lds_u16(uint32_t a)
{
  unsigned short v;
  asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ void sts_u16(uint32_t a, unsigned v) { asm volatile("st.shared.u16 [%0], %1;" ::"r"(a), "h"((unsigned short)v) : "memory"); }
__device__ __forceinline__ void sts_i32(uint32_t a, i32 v) { asm volatile("st.shared.s32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }

__device__ __forceinline__ uint4 lds_v4(uint32_t a)
{
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ void sts_v4(uint32_t a, uint4 v)
{
  asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// shared memory of a CTA: [ header | level starts | gather indices | row bases | xs (+ zero slot) | acc | dinv | aux | NBUF slabs ]
//   header        : the mbarriers (one per slab)
//   level starts  : first tile-local row of every tile-local level (+ end), u16
//   gather indices: per row 8 x u16 (one 16-byte word): tile-local row of the in-tile column in slot e, or the zero slot
//   row bases     : per row  (first slot of the row's slice in the slab) * 32 + lane  |  slice width << 24
//   xs            : the tile's part of `out`, followed by a slot that always holds 0.0
//   acc           : per row  rin - (couplings to other tiles);  dinv;  aux: self (ADD_SELF) or diag (WRITE_R)
constexpr int CTILE_HDR = 128;
constexpr int CTILE_MAXLEV = 256;                                   // tile-local levels fit a byte (tiles.cpp)
constexpr int CTILE_LS_BYTES = 640;                                 // (CTILE_MAXLEV + 8) u16, rounded to 128 bytes
__host__ __device__ inline size_t ctile_fixed_bytes(int maxs)
{
  const size_t rows = (size_t)maxs * 32;
  return (size_t)CTILE_HDR + CTILE_LS_BYTES + rows * 16 + rows * 4 + (rows + 16) * 8 + 3 * rows * 8;
}
__host__ __device__ inline size_t ctile_smem_bytes(int maxs, int cap_slots, int nbuf)
{
  return ((ctile_fixed_bytes(maxs) + 127) / 128) * 128 + (size_t)nbuf * (size_t)cap_slots * 32 * 12;
}

// One CTA per tile.  All warps fetch, wait, fold the couplings to other tiles into the rows' accumulators and PREPARE the solver's
// operands; then ONE warp (the solver) walks the tile-local levels with __syncwarp between them -- a CTA-wide barrier per level costs
// ~0.35 us with one late warp (measured, scripts/micro/), a warp walking the levels alone ~0.12 us per 32 rows -- and all warps publish.
// Small CTAs (NT = 128 for 256-row tiles) so that 6+ tiles per SM are in flight: the phases of different tiles overlap.
// NBUF = slabs in shared memory: 1 = the next tile's slab is fetched when the current tile is finished, 2 = while the current tile is
// being swept (HBM latency never exposed; twice the shared memory).
template <int NT, int MAXS, int NBUF, bool ADD_SELF, bool WRITE_R, int MINB>
__global__ void __launch_bounds__(NT, MINB) k_gs_ctile(SellView T, const double *__restrict__ diag, const double *__restrict__ dinv,
                                                       const double *rin, const double *__restrict__ self, double *out, double *rout,
                                                       CTileParams p)
{
  constexpr int NW = NT / 32;
  static_assert(MAXS % NW == 0, "every warp owns the same number of slices");
  constexpr int NR = MAXS / NW;                 // rows per thread: tile-local rows tid, tid + NT, ...
  constexpr int CH = (NR <= 2) ? 8 : 4;         // slots gathered per round and row
  constexpr unsigned ZERO = MAXS * 32;          // gather index of the zero slot
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const size_t slab_bytes = (size_t)p.cap_slots * 32 * 12;
  int tid;                                                            // (opaque: S2R SR_TID.X would be re-issued inside the loops)
  asm volatile("mov.u32 %0, %%tid.x;" : "=r"(tid));
  const int lane = tid & 31, w = tid >> 5;
  // the shared-window base goes through an opaque move: otherwise the compiler re-derives it from SR_CgaCtaId (S2R, slow) in front of
  // every shared-memory access instead of keeping it in a register
  uint32_t smem_base;
  asm volatile("mov.u32 %0, %1;" : "=r"(smem_base) : "r"(smem_u32(smem_raw)));
  const uint32_t bar0 = smem_base;                                    // mbarrier b at bar0 + 8 b
  const uint32_t ls_a = smem_base + CTILE_HDR;                        // level starts
  const uint32_t ix_a = ls_a + CTILE_LS_BYTES;                        // gather indices
  const uint32_t rb_a = ix_a + MAXS * 32 * 16;                        // row bases
  const uint32_t xs_a = rb_a + MAXS * 32 * 4;
  const uint32_t acc_a = xs_a + (MAXS * 32 + 16) * 8, dv_a = acc_a + MAXS * 32 * 8, aux_a = dv_a + MAXS * 32 * 8;
  const uint32_t slab_a0 = smem_base + (uint32_t)(((ctile_fixed_bytes(MAXS) + 127) / 128) * 128);
  if (tid == 0) {
    for (int b = 0; b < NBUF; b++) mbar_init(bar0 + 8 * b, 1);
    mbar_fence_init();
    sts_f64(xs_a + ZERO * 8u, 0.0);
  }
  __syncthreads();
  auto tile_of = [&](i32 q) { return p.backward ? (p.ntiles - 1 - q) : q; };
  auto fetch = [&](const CTileMeta &m, int buf) {                     // one thread: both bulk copies of a tile's slab
    if (m.nslots <= 0) return;
    const uint32_t sl = slab_a0 + (uint32_t)buf * (uint32_t)slab_bytes;
    const uint32_t bar = bar0 + 8 * buf;
    mbar_expect_tx(bar, (uint32_t)m.nslots * 32u * 12u);
    bulk_g2s(sl, T.val + m.base * 32, (uint32_t)m.nslots * 256u, bar);
    bulk_g2s(sl + (uint32_t)p.cap_slots * 256u, T.col + m.base * 32, (uint32_t)m.nslots * 128u, bar);
  };
  const CTileMeta none{0, 0, 0, 0, 0, 0, 0};
  uint32_t phases = 0;                                                // bit b: parity the next wait on mbarrier b expects
  i32 q = blockIdx.x;
  if (q >= p.ntiles) return;
  // software pipeline: the record of the NEXT tile (and the first tiles it waits for) is loaded while the current tile is swept
  CTileMeta cur = p.meta[tile_of(q)];
  i32 cur_dep = (tid < cur.nd) ? p.dep[cur.d0 + tid] : -1;
  if (tid == 0) fetch(cur, 0);
  int buf = 0;
  int solver = (int)(blockIdx.x % NW);
  for (; q < p.ntiles; q += gridDim.x) {
    const i32 t = tile_of(q);
    unsigned long long *tr = (p.trace && tid == 0) ? p.trace + (size_t)t * 16 : nullptr;
    if (tr) tr[0] = gtimer();
    const i32 qn = q + (i32)gridDim.x;
    const bool more = qn < p.ntiles;
    const CTileMeta nxt = more ? p.meta[tile_of(qn)] : none;          // in flight until the end of this iteration
    if (NBUF == 2 && more && tid == 0) fetch(nxt, buf ^ 1);           // the other slab is free: its tile finished an iteration ago
    const i32 s0 = cur.s0;
    const int ns = cur.ns;
    const i32 r0 = s0 * 32;
    const unsigned nrow = (unsigned)ns * 32u;
    const int nlev = cur.nlev & 0xffff, nreal = cur.nlev >> 16;       // tile-local levels, rows that are not padding
    const uint32_t vals_a = slab_a0 + (uint32_t)buf * (uint32_t)slab_bytes;      // values of the slab, then its column indices
    const uint32_t cols_a = vals_a + (uint32_t)p.cap_slots * 256u;
    // ---- the first hint flag is requested before anything else (one L2 round trip that everything below overlaps)
    int flag0 = 1;
    if (tid < cur.nd) flag0 = ld_relaxed_i32(p.done + cur_dep);
    // ---- per-row data that does not depend on `out`
    double acc[NR], dv[NR], aux[NR];
    int lv[NR], lvp[NR], sb[NR], wd[NR];
#pragma unroll
    for (int j = 0; j < NR; j++) {
      const int sl = w + j * NW;
      lv[j] = 256; lvp[j] = 256; sb[j] = 0; wd[j] = 0; acc[j] = 0.0; dv[j] = 0.0; aux[j] = 0.0;
      if (sl < ns) {
        const i64 slice = (i64)s0 + sl, row = slice * 32 + lane;
        const i64 b = T.slice_ptr[slice];
        sb[j] = (int)(b - cur.base);
        wd[j] = (int)(T.slice_ptr[slice + 1] - b);
        lv[j] = p.row_lvl[row];
        lvp[j] = (sl == 0 && lane == 0) ? -1 : (int)p.row_lvl[row - 1];
        acc[j] = rin[row];
        dv[j] = dinv[row];
        aux[j] = ADD_SELF ? self[row] : (WRITE_R ? diag[row] : 0.0);
      }
    }
    // ---- hint flags of the tiles this one depends on (one thread per dependency; the first NT ids were prefetched)
    for (int k = tid; k < cur.nd; k += NT) {
      const int *f = p.done + ((k < NT) ? cur_dep : p.dep[cur.d0 + k]);
      unsigned spins = 0;
      int fl = (k < NT) ? flag0 : ld_relaxed_i32(f);
      while (fl == 0) {
        if (p.sleep_ns) __nanosleep(p.sleep_ns);
        if (spin_fail(spins, p.err)) break;
        fl = ld_relaxed_i32(f);
      }
    }
    if (tr) tr[1] = gtimer();
    if (cur.nslots > 0) { mbar_wait(bar0 + 8 * buf, (phases >> buf) & 1u); phases ^= 1u << buf; }
    __syncthreads();
    if (tr) tr[2] = gtimer();
    // ---- couplings to rows of other tiles: poll the data itself (sentinel).  All polls of a round (CH slots of every owned row) are in
    // flight together; a round with a straggler is simply polled again as a whole.
    int maxw = 0;
#pragma unroll
    for (int j = 0; j < NR; j++) maxw = max(maxw, wd[j]);
#pragma unroll 1
    for (int k0 = 0; k0 < maxw; k0 += CH) {
      double xv[NR][CH];
      unsigned spins = 0;
      bool missing;
#pragma unroll 1
      do {
        missing = false;
#pragma unroll
        for (int j = 0; j < NR; j++)
#pragma unroll
          for (int e = 0; e < CH; e++) {
            const int k = k0 + e;
            const i32 c = (k < wd[j]) ? lds_i32(cols_a + (uint32_t)((sb[j] + k) * 32 + lane) * 4u) : -1;
            xv[j][e] = (c >= 0 && (unsigned)(c - r0) >= nrow) ? ld_poll_relaxed(out + c) : 0.0;   // in-tile columns: the solver's business
          }
#pragma unroll
        for (int j = 0; j < NR; j++)
#pragma unroll
          for (int e = 0; e < CH; e++) missing |= is_sentinel(xv[j][e]);
        if (missing) {
          if (p.repoll_ns) __nanosleep(p.repoll_ns);
          if (spin_fail(spins, p.err)) break;
        }
      } while (missing);
#pragma unroll
      for (int j = 0; j < NR; j++)
#pragma unroll
        for (int e = 0; e < CH; e++) {
          const int k = k0 + e;
          // xv == 0 for in-tile / padding slots; the value slot is read only where the row has one
          if (k < wd[j]) acc[j] = fma(-lds_f64(vals_a + (uint32_t)((sb[j] + k) * 32 + lane) * 8u), xv[j][e], acc[j]);
        }
    }
    // ---- hand the rows to the solver: per-row scalars, gather indices of the first 8 slots, row base, first row of every level
#pragma unroll
    for (int j = 0; j < NR; j++) {
      const int sl = w + j * NW;
      if (sl < ns) {
        const uint32_t lr = (uint32_t)(tid + j * NT);
        sts_f64(acc_a + lr * 8u, acc[j]);
        sts_f64(dv_a + lr * 8u, dv[j]);
        sts_f64(aux_a + lr * 8u, aux[j]);
        unsigned h[8];
#pragma unroll
        for (int e = 0; e < 8; e++) {
          const i32 c = (e < wd[j]) ? lds_i32(cols_a + (uint32_t)((sb[j] + e) * 32 + lane) * 4u) : -1;
          const unsigned lc = (unsigned)(c - r0);
          h[e] = (c >= 0 && lc < nrow) ? lc : ZERO;
        }
        sts_v4(ix_a + lr * 16u, make_uint4(h[0] | (h[1] << 16), h[2] | (h[3] << 16), h[4] | (h[5] << 16), h[6] | (h[7] << 16)));
        sts_i32(rb_a + lr * 4u, (sb[j] * 32 + lane) | (min(wd[j], 255) << 24));
        if (lv[j] != lvp[j] && lv[j] < 255) sts_u16(ls_a + (uint32_t)lv[j] * 2u, lr);
      }
    }
    if (tid == 0) sts_u16(ls_a + (uint32_t)nlev * 2u, (unsigned)nreal);
    if (tr) tr[3] = gtimer();
    // the next tile's first dependency ids: its record has arrived by now
    const i32 nxt_dep = (more && tid < nxt.nd) ? p.dep[nxt.d0 + tid] : -1;
    const int wide = __syncthreads_or(maxw > 8);                      // rows with more than 8 slots: the solver's second round
    // ---- the tile itself: the solver warp walks the local levels (ascending forward, descending backward); in-tile couplings come from xs
    // (the solver role rotates over the warps: warp w of every CTA sits on scheduler w % 4 -- a fixed solver warp would put the
    // latency-critical chains of all resident CTAs on the same scheduler)
    if (w == solver) {
#pragma unroll 1
      for (int it = 0; it < nlev; it++) {
        const int s = p.backward ? (nlev - 1 - it) : it;
        const int rb = (int)lds_u16(ls_a + (uint32_t)s * 2u), re = (int)lds_u16(ls_a + (uint32_t)(s + 1) * 2u);
#pragma unroll 1
        for (int rr = rb; rr < re; rr += 32) {
          const int r = rr + lane;
          const bool act = r < re;
          const uint32_t rc = (uint32_t)(act ? r : rb);                 // idle lanes shadow the first row (results discarded)
          const uint4 iw = lds_v4(ix_a + rc * 16u);
          const i32 rbw = lds_i32(rb_a + rc * 4u);
          const uint32_t vb = vals_a + (uint32_t)(rbw & 0xffffff) * 8u;
          const int wdt = (int)((unsigned)rbw >> 24);
          double a = lds_f64(acc_a + rc * 8u), a2 = 0.0;
          const double dvv = lds_f64(dv_a + rc * 8u), ax = lds_f64(aux_a + rc * 8u);
          const unsigned ix[4] = {iw.x, iw.y, iw.z, iw.w};
          double v[8], x[8];
#pragma unroll
          for (int e = 0; e < 8; e++) v[e] = (e < wdt) ? lds_f64(vb + (uint32_t)e * 256u) : 0.0;
#pragma unroll
          for (int e = 0; e < 8; e++) x[e] = lds_f64(xs_a + ((ix[e >> 1] >> ((e & 1) * 16)) & 0xffffu) * 8u);
#pragma unroll
          for (int e = 0; e < 8; e += 2) { a = fma(-v[e], x[e], a); a2 = fma(-v[e + 1], x[e + 1], a2); }
          if (wide) {
            // slots 8.. of wide rows: column indices straight from the slab (rare: coarse-like matrices)
            const uint32_t cb = cols_a + (uint32_t)(rbw & 0xffffff) * 4u;
            for (int k = 8; k < wdt; k++) {
              const i32 c = lds_i32(cb + (uint32_t)k * 128u);
              const unsigned lc = (unsigned)(c - r0);
              if (c >= 0 && lc < nrow) a = fma(-lds_f64(vb + (uint32_t)k * 256u), lds_f64(xs_a + lc * 8u), a);
            }
          }
          a += a2;
          const double d = dvv * a;
          if (act) {
            sts_f64(xs_a + (uint32_t)r * 8u, ADD_SELF ? ax + d : d);
            if (WRITE_R) sts_f64(acc_a + (uint32_t)r * 8u, fma(-ax, d, a));   // the row's new residual
          }
        }
        __syncwarp();
      }
    }
    __syncthreads();
    if (tr) tr[4] = gtimer();
    // ---- publish the tile: coalesced stores (padding rows: never updated, but `out` must not keep the sentinel), then the hint flag
#pragma unroll
    for (int j = 0; j < NR; j++)
      if (lv[j] <= 255) {
        const uint32_t lr = (uint32_t)(tid + j * NT);
        const i64 row = (i64)r0 + lr;
        if (lv[j] == 255) {
          __stcg(out + row, ADD_SELF ? aux[j] : 0.0);
          if (WRITE_R) rout[row] = acc[j];
        } else {
          __stcg(out + row, lds_f64(xs_a + lr * 8u));
          if (WRITE_R) rout[row] = lds_f64(acc_a + lr * 8u);
        }
      }
    if (tid == 0) st_relaxed_i32(p.done + t, 1);
    if (tr) {
      unsigned smid;
      asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
      tr[5] = gtimer();
      tr[6] = (unsigned long long)smid | ((unsigned long long)blockIdx.x << 16) | ((unsigned long long)nlev << 40) | ((unsigned long long)ns << 52);
    }
    __syncthreads();     // slab, xs and the per-row arrays are free again
    if (NBUF == 1) { if (more && tid == 0) fetch(nxt, 0); }
    else buf ^= 1;
    cur = nxt;
    cur_dep = nxt_dep;
    solver = (solver + 1) % NW;
  }
}

}  // namespace ngb
