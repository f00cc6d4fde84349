// kernels_rm.cuh -- the triangular Gauss-Seidel half-sweep on SMALL levels (coarse levels: few rows, wide rows, a dependency level
// holds far fewer rows than the machine has warps), on a ROW-MAJOR copy of the triangle.
//
// k_gs_tri_small (kernels.cuh) walks the SELL-32 layout with one warp per row: the 32 lanes of a row read 32 different 128-byte
// lines (4 useful bytes each), and every dependency level pays three dependent memory latencies (slice pointer -> entries -> poll)
// plus a fourth for the row's right-hand side and diagonal after the reduction.  Here
//   * the triangle is stored row by row (ptr / col / val, entries of a row contiguous, SELL slot order kept): a warp reads a row
//     with fully coalesced 128 / 256-byte requests;
//   * everything that does not depend on the sweep -- the row pointers (two rows ahead), the columns, the (scalar) values, the row's
//     right-hand side, `self`, and its blocks of dinv / diag (one row ahead) -- is staged in shared memory with cp.async BEFORE the
//     warp starts polling.  (Prefetching into registers with plain loads was measured 1.5x SLOWER than no prefetch at all: a polling
//     load shares its scoreboard with the outstanding prefetch loads and waits for their DRAM latency; cp.async completion is tracked
//     by its own wait-group counter.)
//   * the polls of all staged chunks of a row (up to 128 entries for scalar matrices) are in flight TOGETHER; k_gs_tri_small spins
//     chunk after chunk, one L2 round trip each.  The dependency chain of a level is poll -> shuffle tree -> one 8-byte store;
//   * the per-row epilogue is spread over the lanes (lane p owns component p of a block row).
// Same protocol as k_gs_tri: `out` sentinel-filled, rows dealt round-robin to the resident warps in sweep order, the data is the flag.
#pragma once
#include "kernels_ctile.cuh"

namespace ngb {

struct RmView {
  const i64 *ptr;      // [nrows_pad + 1], in entries
  const i32 *col;
  const double *val;   // [entry][bh*bw] row-major blocks
  const i32 *gate;     // per row: the dependency that is published last (largest column of L / smallest of U), -1 = none
};

__global__ void k_rm_count(i64 nrows_pad, SellView T, i64 *cnt)
{
  const i64 row = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (row > nrows_pad) return;
  if (row == nrows_pad) { cnt[row] = 0; return; }
  const i64 slice = row >> 5;
  const int lane = (int)(row & 31);
  const i64 base = T.slice_ptr[slice];
  const int width = (int)(T.slice_ptr[slice + 1] - base);
  int c = 0;
  for (int k = 0; k < width; k++) c += T.col[(base + k) * 32 + lane] >= 0;
  cnt[row] = c;
}

__global__ void k_rm_fill(i64 nrows_pad, int bs, SellView T, const i64 *__restrict__ ptr, i32 *col, double *val, i32 *gate, int upper, i32 nonfree)
{
  const i64 row = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= nrows_pad) return;
  const i64 slice = row >> 5;
  const int lane = (int)(row & 31);
  const i64 base = T.slice_ptr[slice];
  const int width = (int)(T.slice_ptr[slice + 1] - base);
  i64 o = ptr[row];
  // rows are numbered level-major and dealt to the warps in sweep order: the dependency that is published last is (almost always)
  // the one closest to the row itself.  Couplings between two non-free rows are skipped by the sweep (see k_gs_tri).
  const i32 cut = (row < nonfree) ? nonfree : 0;
  i32 g = upper ? 0x7fffffff : -1;
  for (int k = 0; k < width; k++) {
    const i32 c = T.col[(base + k) * 32 + lane];
    if (c < 0) continue;
    if (c >= cut) g = upper ? min(g, c) : max(g, c);
    col[o] = c;
    for (int e = 0; e < bs; e++) val[o * bs + e] = T.val[((base + k) * (i64)bs + e) * 32 + lane];
    o++;
  }
  gate[row] = (g == 0x7fffffff) ? -1 : g;
}

__device__ __forceinline__ void cp_async4(uint32_t dst, const void *src) { asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory"); }
__device__ __forceinline__ void cp_async8(uint32_t dst, const void *src) { asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory"); }
__device__ __forceinline__ long long lds_i64(uint32_t a)
{
  long long v;
  asm volatile("ld.shared.s64 %0, [%1];" : "=l"(v) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

constexpr int RM_THREADS = 64;   // two warps per CTA: the polling warps of a small level spread over all SMs (a spinning warp occupies
                                 // the SM's memory pipeline, which the shuffles / shared-memory loads of the critical warp share)

template <int B>
struct RmStage {
  static constexpr int BS = B * B;
  static constexpr int NCH = (B == 1) ? 4 : 2;              // staged chunks of 32 entries
  static constexpr int NSC = 2 * B + 2 * BS;                // rin | self | dinv | diag
  static constexpr int COLS = 0;                            // byte offsets inside one stage
  static constexpr int VALS = COLS + NCH * 32 * 4;          // scalar matrices only
  static constexpr int SCAL = VALS + (B == 1 ? NCH * 32 * 8 : 0);
  static constexpr int BYTES = SCAL + ((NSC * 8 + 15) / 16) * 16;
  static constexpr int WARP_BYTES = 2 * BYTES + 32 + 16;    // two stages + two row-pointer pairs + two gate columns
};

template <int B, bool ADD_SELF, bool WRITE_R>
__global__ void __launch_bounds__(RM_THREADS) k_gs_tri_rm(RmView T, const double *__restrict__ diag, const double *__restrict__ dinv,
                                                  const double *rin, const double *__restrict__ self, double *out, double *rout,
                                                  TriParams prm)
{
  using St = RmStage<B>;
  constexpr int BS = B * B, NCH = St::NCH;
  constexpr unsigned FULL = 0xffffffffu;
  __shared__ __align__(16) unsigned char smem[(RM_THREADS / 32) * St::WARP_BYTES];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const i64 gw = ((i64)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const i64 nw = ((i64)gridDim.x * blockDim.x) >> 5;
  const i64 nrows = prm.nslices * 32;
  if (gw >= nrows) return;
  const i64 rlast = nrows - 1;
  auto rowof = [&](i64 r) { return prm.backward ? (nrows - 1 - r) : r; };
  const uint32_t wbase = smem_u32(smem) + (uint32_t)w * St::WARP_BYTES;
  const uint32_t ptr_a = wbase + 2 * St::BYTES;               // [2][2] i64
  const uint32_t gate_a = ptr_a + 32;                         // [2] i32

  // stage the entries and scalars of `row` (entries [p0, p1)) into stage s; nothing here depends on the sweep
  auto stage_row = [&](int s, i64 row, i64 p0, i64 p1) {
    const uint32_t sa = wbase + (uint32_t)s * St::BYTES;
    const int cnt = (int)min((i64)(NCH * 32), p1 - p0);
#pragma unroll
    for (int ch = 0; ch < NCH; ch++) {
      const int k = ch * 32 + lane;
      if (k < cnt) {
        cp_async4(sa + St::COLS + (uint32_t)k * 4u, T.col + p0 + k);
        if (B == 1) cp_async8(sa + St::VALS + (uint32_t)k * 8u, T.val + p0 + k);
      }
    }
    if (B > 1 && prm.gate_all) {
      // block values are loaded at use (36 doubles per entry do not fit the stage): on the latency-bound (small) levels pull the row's
      // lines into L2 now, so that the loads behind the polls are L2 hits instead of DRAM misses on the critical path (a bandwidth-bound
      // level loses with the extra requests: the host sets the flag below tri_rm_max_rows only)
      const char *vb = (const char *)(T.val + p0 * BS);
      const i64 nbytes = (p1 - p0) * (i64)(BS * 8);
      for (i64 o = (i64)lane * 128; o < nbytes; o += 32 * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(vb + o));
    }
    if (lane == 31) cp_async4(gate_a + (uint32_t)s * 4u, T.gate + row);
    const i64 slice = row >> 5;
    const int lr = (int)(row & 31);
#pragma unroll
    for (int j0 = 0; j0 < St::NSC; j0 += 32) {
      const int j = j0 + lane;
      if (j < St::NSC) {
        const double *src;
        if (j < B) src = rin + row * B + j;
        else if (j < 2 * B) src = ADD_SELF ? self + row * B + (j - B) : nullptr;
        else if (j < 2 * B + BS) src = dinv + slice * (i64)BS * 32 + (i64)(j - 2 * B) * 32 + lr;
        else src = WRITE_R ? diag + slice * (i64)BS * 32 + (i64)(j - 2 * B - BS) * 32 + lr : nullptr;
        if (src) cp_async8(sa + St::SCAL + (uint32_t)j * 8u, src);
      }
    }
  };
  auto stage_ptr = [&](int s, i64 row) {
    if (lane < 2) cp_async8(ptr_a + (uint32_t)(s * 2 + lane) * 8u, T.ptr + row + lane);
  };

  i64 r = gw;
  i64 p0, p1;
  {
    const i64 row = rowof(r);
    p0 = T.ptr[row]; p1 = T.ptr[row + 1];
    stage_row(0, row, p0, p1);
    stage_ptr(1, rowof(min(r + nw, rlast)));
  }
  int it = 0;
  for (; r < nrows; r += nw, it ^= 1) {
    const i64 row = rowof(r);
    unsigned long long *tr = (prm.trace && lane == 0) ? prm.trace + row * 12 : nullptr;
    if (tr) tr[0] = gtimer();
    cp_async_wait_all();
    __syncwarp();
    if (tr) tr[4] = gtimer();
    // ---- prefetch: entries and scalars of this warp's next row, row pointers of the one after
    const i64 rn = rowof(min(r + nw, rlast));
    const i64 q0 = (i64)lds_i64(ptr_a + (uint32_t)((it ^ 1) * 2) * 8u), q1 = (i64)lds_i64(ptr_a + (uint32_t)((it ^ 1) * 2 + 1) * 8u);
    if (tr) tr[5] = gtimer() + (q0 & 0);
    stage_row(it ^ 1, rn, q0, q1);
    if (tr) tr[6] = gtimer();
    stage_ptr(it, rowof(min(r + 2 * nw, rlast)));
    if (tr) tr[1] = gtimer();
    // ---- gate: the warp first polls its newest dependency only: a spinning warp costs one sector per poll instead of up to 32 per chunk
    // Control flow below is kept WARP-UNIFORM (votes instead of per-lane spin loops): after a divergent spin loop the warp reaches
    // the shuffle tree split into groups and every SHFL takes the slow collective path (measured 1.3 us for the 5-stage butterfly
    // instead of 0.09 us).
    if (prm.prepoll) {
      const i32 g = lds_i32(gate_a + (uint32_t)it * 4u);     // same address in every lane: one sector per poll
      if (g >= 0) {
        const double *gp = out + (i64)g * B + (B - 1);
        unsigned spins = 0;
        while (is_sentinel(ld_poll_relaxed(gp))) {
          if (prm.sleep_ns) __nanosleep(prm.sleep_ns);
          if (spin_fail(spins, prm.err)) break;
        }
      }
      __syncwarp();
    }
    if (tr) tr[7] = gtimer();
    // ---- this row: all polls of the staged chunks in flight together
    const uint32_t sa = wbase + (uint32_t)it * St::BYTES;
    const i32 cut = (row < prm.nonfree) ? (i32)prm.nonfree : 0;
    const int cnt = (int)(p1 - p0);
    i32 c[NCH];
    double xv[NCH][B];
#pragma unroll
    for (int ch = 0; ch < NCH; ch++) {
      const int k = ch * 32 + lane;
      c[ch] = (k < cnt) ? lds_i32(sa + St::COLS + (uint32_t)k * 4u) : -1;
      if (c[ch] < cut) c[ch] = -1;
#pragma unroll
      for (int q = 0; q < B; q++) xv[ch][q] = (c[ch] >= 0) ? ld_poll_relaxed(out + (i64)c[ch] * B + q) : 0.0;
    }
    {
      unsigned spins = 0;
      for (;;) {
        bool miss = false;
#pragma unroll
        for (int ch = 0; ch < NCH; ch++)
#pragma unroll
          for (int q = 0; q < B; q++) miss |= is_sentinel(xv[ch][q]);
        if (!__any_sync(FULL, miss)) break;
        if (prm.sleep_ns) __nanosleep(prm.sleep_ns);
#pragma unroll
        for (int ch = 0; ch < NCH; ch++)
#pragma unroll
          for (int q = 0; q < B; q++)
            if (is_sentinel(xv[ch][q])) xv[ch][q] = ld_poll_relaxed(out + (i64)c[ch] * B + q);
        const bool fail = spin_fail(spins, prm.err);
        if (__any_sync(FULL, fail)) break;
      }
    }
    double acc[B];
#pragma unroll
    for (int p = 0; p < B; p++) acc[p] = 0.0;
#pragma unroll
    for (int ch = 0; ch < NCH; ch++) {
      if (B == 1) {
        const double a = lds_f64(sa + St::VALS + (uint32_t)(ch * 32 + lane) * 8u);
        if (c[ch] >= 0) acc[0] = fma(-a, xv[ch][0], acc[0]);
      } else if (c[ch] >= 0) {
        double a[BS];
#pragma unroll
        for (int e = 0; e < BS; e++) a[e] = T.val[(p0 + ch * 32 + lane) * BS + e];
#pragma unroll
        for (int p = 0; p < B; p++) {
          double t = 0.0;
#pragma unroll
          for (int q = 0; q < B; q++) t = fma(a[p * B + q], xv[ch][q], t);
          acc[p] -= t;
        }
      }
    }
    if (cnt > NCH * 32) {                                  // rows wider than the staged window (rare); warp-uniform trip count
      for (int k0 = NCH * 32; k0 < cnt; k0 += 32) {
        const int k = k0 + lane;
        i32 cc = (k < cnt) ? T.col[p0 + k] : -1;
        if (cc < cut) cc = -1;
        double xw[B];
#pragma unroll
        for (int q = 0; q < B; q++) xw[q] = (cc >= 0) ? ld_poll_relaxed(out + (i64)cc * B + q) : 0.0;
        unsigned spins = 0;
        for (;;) {
          bool miss = false;
#pragma unroll
          for (int q = 0; q < B; q++) miss |= is_sentinel(xw[q]);
          if (!__any_sync(FULL, miss)) break;
          if (prm.sleep_ns) __nanosleep(prm.sleep_ns);
#pragma unroll
          for (int q = 0; q < B; q++)
            if (is_sentinel(xw[q])) xw[q] = ld_poll_relaxed(out + (i64)cc * B + q);
          const bool fail = spin_fail(spins, prm.err);
          if (__any_sync(FULL, fail)) break;
        }
        if (cc >= 0) {
          double a[BS];
#pragma unroll
          for (int e = 0; e < BS; e++) a[e] = T.val[(p0 + k) * BS + e];
#pragma unroll
          for (int p = 0; p < B; p++) {
            double t = 0.0;
#pragma unroll
            for (int q = 0; q < B; q++) t = fma(a[p * B + q], xw[q], t);
            acc[p] -= t;
          }
        }
      }
    }
    __syncwarp();
    if (tr) tr[2] = gtimer();
    // xor butterfly: every lane ends with the same bits (a + b == b + a)
#pragma unroll
    for (int p = 0; p < B; p++)
#pragma unroll
      for (int o = 16; o; o >>= 1) acc[p] += __shfl_xor_sync(FULL, acc[p], o);
    if (tr) tr[8] = gtimer() + (__double_as_longlong(acc[0]) & 0);
    // ---- epilogue: lane p owns component p.  af = rin - sum ; dl = dinv af ; out = self + dl ; rout = af - diag dl
    const uint32_t sc = sa + St::SCAL;
    double mine = 0.0, selfv = 0.0;
    if (lane < B) {
#pragma unroll
      for (int p = 0; p < B; p++)
        if (lane == p) mine = acc[p];
      mine += lds_f64(sc + (uint32_t)lane * 8u);
      if (ADD_SELF) selfv = lds_f64(sc + (uint32_t)(B + lane) * 8u);
    }
    if (tr) tr[9] = gtimer() + (__double_as_longlong(mine) & 0);
    if (B == 1) {
      if (lane == 0) {
        const double dlp = fma(lds_f64(sc + 2u * 8u), mine, 0.0);
        __stcg(out + row, ADD_SELF ? selfv + dlp : dlp);
        if (WRITE_R) rout[row] = fma(-lds_f64(sc + 3u * 8u), dlp, mine);
      }
    } else {
      double af[B], dl[B];
#pragma unroll
      for (int q = 0; q < B; q++) af[q] = __shfl_sync(FULL, mine, q);
      double dlp = 0.0;
      const int lp = lane < B ? lane : 0;
#pragma unroll
      for (int q = 0; q < B; q++) dlp = fma(lds_f64(sc + (uint32_t)(2 * B + lp * B + q) * 8u), af[q], dlp);
#pragma unroll
      for (int q = 0; q < B; q++) dl[q] = __shfl_sync(FULL, dlp, q);
      if (lane < B) {
        __stcg(out + row * B + lane, ADD_SELF ? selfv + dlp : dlp);
        if (WRITE_R) {
          double t = mine;
#pragma unroll
          for (int q = 0; q < B; q++) t = fma(-lds_f64(sc + (uint32_t)(2 * B + BS + lane * B + q) * 8u), dl[q], t);
          rout[row * B + lane] = t;
        }
      }
    }
    if (tr) tr[3] = gtimer();
    p0 = q0; p1 = q1;
  }
  cp_async_wait_all();
}

// y_out = beta * y_in + alpha * (S1 v [+ S2 v] [+ D v]) on the row-major copies, one WARP per block row: the lanes split the row's entries
// and read each block as 8*B*B contiguous bytes (full sectors; the SELL walk of k_sell_spmv_small uses a quarter of every sector for
// blocks), fixed shuffle tree, lane p writes component p.  For block levels with too few rows for the thread-per-row kernel.
template <int B, bool HAS_S2, bool HAS_D>
__global__ void __launch_bounds__(256) k_rm_spmv(i64 nrows_pad, RmView S1, RmView S2, const double *__restrict__ diag, const double *__restrict__ v,
                                                const double *y_in, double *y_out, double alpha, double beta)
{
  constexpr int BS = B * B;
  constexpr unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const i64 gw = ((i64)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const i64 nw = ((i64)gridDim.x * blockDim.x) >> 5;
  for (i64 row = gw; row < nrows_pad; row += nw) {
    double acc[B];
#pragma unroll
    for (int p = 0; p < B; p++) acc[p] = 0.0;
    auto part = [&](const RmView &S) {
      const i64 p0 = S.ptr[row], p1 = S.ptr[row + 1];
      for (i64 k = p0 + lane; k < p1; k += 32) {
        const i32 c = S.col[k];
        double xv[B];
#pragma unroll
        for (int q = 0; q < B; q++) xv[q] = v[(i64)c * B + q];
#pragma unroll
        for (int p = 0; p < B; p++) {
          double t = 0.0;
#pragma unroll
          for (int q = 0; q < B; q++) t = fma(S.val[k * BS + p * B + q], xv[q], t);
          acc[p] += t;
        }
      }
    };
    part(S1);
    if (HAS_S2) part(S2);
#pragma unroll
    for (int p = 0; p < B; p++)
#pragma unroll
      for (int o = 16; o; o >>= 1) acc[p] += __shfl_xor_sync(FULL, acc[p], o);
    if (lane < B) {
      double a = 0.0;
#pragma unroll
      for (int p = 0; p < B; p++)
        if (lane == p) a = acc[p];
      if (HAS_D) {
        const double *dp = diag + (row >> 5) * (i64)BS * 32 + (row & 31);
#pragma unroll
        for (int q = 0; q < B; q++) a = fma(dp[(lane * B + q) * 32], v[row * B + q], a);
      }
      double y = alpha * a;
      if (beta != 0.0) y = fma(beta, y_in[row * B + lane], y);
      y_out[row * B + lane] = y;
    }
  }
}

}  // namespace ngb
