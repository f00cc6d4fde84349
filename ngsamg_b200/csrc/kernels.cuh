// kernels.cuh -- sm_100a device code of the V-cycle hot path.
//
// Data layout.  Every level lives on the device in a *level-scheduled numbering*:
// rows are sorted by their Gauss-Seidel dependency level (longest path in the DAG "j < i and A_ij != 0"),
// each level padded to a multiple of 32 rows.  A sequential sweep over that numbering is the same
// sweep as the reference's (gssmoother.cpp:195-315) because every DAG edge keeps its orientation.
// Matrices are SELL-32 with element-planar blocks: a slice = 32 consecutive rows = one warp, thread per
// block row; slot k of a slice stores col[(base+k)*32 + lane] and, for each of the bh*bw block
// entries e, val[((base+k)*bh*bw + e)*32 + lane] -- every load instruction of a warp is one
// fully coalesced 128 B (cols) / 256 B (values) segment.  A is split into strictly-lower L, diagonal D
// and strictly-upper U so that each half-sweep streams exactly the bytes it needs.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ngb {

using i64 = int64_t;
using i32 = int32_t;

struct SellView {
  const i64 *slice_ptr;  // [nslices+1], in slots
  const i32 *col;        // -1 = padding
  const double *val;
};

__device__ __forceinline__ double ld_cg(const double *p) { return __ldcg(p); }

__device__ __forceinline__ int ld_acquire(const int *p)
{
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_add(int *p, int v)
{
  asm volatile("red.release.gpu.global.add.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// acc[BH] -= / += sum_k blk_k * x[col_k]   over one SELL row.  CG = bypass L1 for the gathered vector
// (needed inside the triangular kernels where other CTAs produce those values during the launch).
template <int BH, int BW, bool CG>
__device__ __forceinline__ void sell_row_mac(const SellView &S, i64 slice, int lane, const double *__restrict__ x,
                                             double (&acc)[BH], double sign)
{
  const i64 base = S.slice_ptr[slice];
  const int width = (int)(S.slice_ptr[slice + 1] - base);
  const i32 *cp = S.col + base * 32 + lane;
  const double *vp = S.val + base * (i64)(BH * BW) * 32 + lane;
#pragma unroll 4
  for (int k = 0; k < width; k++) {
    const i32 c = cp[(i64)k * 32];
    if (c >= 0) {
      double xv[BW];
#pragma unroll
      for (int q = 0; q < BW; q++) xv[q] = CG ? ld_cg(x + (i64)c * BW + q) : x[(i64)c * BW + q];
#pragma unroll
      for (int p = 0; p < BH; p++) {
        double s = 0.0;
#pragma unroll
        for (int q = 0; q < BW; q++) s = fma(vp[((i64)k * (BH * BW) + p * BW + q) * 32], xv[q], s);
        acc[p] = fma(sign, s, acc[p]);
      }
    }
  }
}

// scalar fast path: the row is walked in blocks of U slots; ALL loads of a block (columns, values, then the gathered vector
// entries) are issued before its FMAs, the tail block is masked instead of falling back to one slot at a time -- a slice of width 7
// costs one round of dependent latencies (column -> gather), not four; a prolongation row (width <= 3) one instead of three.
// predicated loads as volatile asm: the compiler keeps volatile asm statements in program order, so the loads of a block are ISSUED as
// a batch (with plain C++ loads ptxas interleaves them with the FMAs to stay within 32 registers and the batch degenerates into a chain)
__device__ __forceinline__ i32 ldp_nc_i32(const i32 *p, bool pred)
{
  i32 v;
  asm volatile("{ .reg .pred q; setp.ne.b32 q, %2, 0; mov.b32 %0, -1; @q ld.global.nc.s32 %0, [%1]; }" : "=r"(v) : "l"(p), "r"((int)pred));
  return v;
}
__device__ __forceinline__ double ldp_nc_f64(const double *p, bool pred)
{
  double v;
  asm volatile("{ .reg .pred q; setp.ne.b32 q, %2, 0; mov.b64 %0, 0; @q ld.global.nc.f64 %0, [%1]; }" : "=d"(v) : "l"(p), "r"((int)pred));
  return v;
}
__device__ __forceinline__ double ldp_f64(const double *p, bool pred, bool cg)
{
  double v;
  if (cg) asm volatile("{ .reg .pred q; setp.ne.b32 q, %2, 0; mov.b64 %0, 0; @q ld.global.cg.f64 %0, [%1]; }" : "=d"(v) : "l"(p), "r"((int)pred));
  else asm volatile("{ .reg .pred q; setp.ne.b32 q, %2, 0; mov.b64 %0, 0; @q ld.global.f64 %0, [%1]; }" : "=d"(v) : "l"(p), "r"((int)pred));
  return v;
}

// legacy walk: blocks of 4 slots, then one slot at a time (the compiler pipelines the blocks of wide rows; best for wide slices and for
// the register-heavy kernel variants)
template <bool CG>
__device__ __forceinline__ double sell_row_dot1_legacy(const SellView &S, i64 slice, int lane, const double *__restrict__ x)
{
  const i64 base = S.slice_ptr[slice];
  const int width = (int)(S.slice_ptr[slice + 1] - base);
  const i32 *cp = S.col + base * 32 + lane;
  const double *vp = S.val + base * 32 + lane;
  double s0 = 0.0, s1 = 0.0;
  int k = 0;
  for (; k + 4 <= width; k += 4) {
    i32 c0 = cp[(i64)(k + 0) * 32], c1 = cp[(i64)(k + 1) * 32], c2 = cp[(i64)(k + 2) * 32], c3 = cp[(i64)(k + 3) * 32];
    double v0 = vp[(i64)(k + 0) * 32], v1 = vp[(i64)(k + 1) * 32], v2 = vp[(i64)(k + 2) * 32], v3 = vp[(i64)(k + 3) * 32];
    double x0 = c0 >= 0 ? (CG ? ld_cg(x + c0) : x[c0]) : 0.0;
    double x1 = c1 >= 0 ? (CG ? ld_cg(x + c1) : x[c1]) : 0.0;
    double x2 = c2 >= 0 ? (CG ? ld_cg(x + c2) : x[c2]) : 0.0;
    double x3 = c3 >= 0 ? (CG ? ld_cg(x + c3) : x[c3]) : 0.0;
    s0 = fma(v0, x0, s0); s1 = fma(v1, x1, s1); s0 = fma(v2, x2, s0); s1 = fma(v3, x3, s1);
  }
  for (; k < width; k++) {
    i32 c = cp[(i64)k * 32];
    double v = vp[(i64)k * 32];
    double xv = c >= 0 ? (CG ? ld_cg(x + c) : x[c]) : 0.0;
    s0 = fma(v, xv, s0);
  }
  return s0 + s1;
}

// U = 0 selects the legacy walk
template <bool CG, int U = 0>
__device__ __forceinline__ double sell_row_dot1(const SellView &S, i64 slice, int lane, const double *__restrict__ x)
{
  if constexpr (U == 0) return sell_row_dot1_legacy<CG>(S, slice, lane, x);
  constexpr int UU = U ? U : 4;
  const i64 base = S.slice_ptr[slice];
  const int width = (int)(S.slice_ptr[slice + 1] - base);
  const i32 *cp = S.col + base * 32 + lane;
  const double *vp = S.val + base * 32 + lane;
  double s0 = 0.0, s1 = 0.0;
  for (int k = 0; k < width; k += UU) {
    i32 c[UU];
    double v[UU], xv[UU];
#pragma unroll
    for (int j = 0; j < UU; j++) c[j] = ldp_nc_i32(cp + (i64)(k + j) * 32, k + j < width);
#pragma unroll
    for (int j = 0; j < UU; j++) v[j] = ldp_nc_f64(vp + (i64)(k + j) * 32, k + j < width);
#pragma unroll
    for (int j = 0; j < UU; j++) xv[j] = ldp_f64(x + (c[j] >= 0 ? c[j] : 0), c[j] >= 0, CG);
#pragma unroll
    for (int j = 0; j < UU; j += 2) { s0 = fma(v[j], xv[j], s0); if (j + 1 < UU) s1 = fma(v[j + 1], xv[j + 1], s1); }
  }
  return s0 + s1;
}

// ------------------------------------------------------------------------------------------------
// K1/K2b/K3a/K5/K6: y_out = beta*y_in + alpha*(S1 v [+ S2 v] [+ D v]);  optionally xadd += v (own rows)
//   plain SpMV (S1=L,S2=U,D), the U-pass of the forward sweep, the (L+D)-pass of the backward sweep,
//   restriction (S1=PT) and prolongation-add (S1=P).  Thread per block row, warp per slice.
// ------------------------------------------------------------------------------------------------
template <int BH, int BW, bool HAS_S2, bool HAS_D, int U = 0>
__global__ void __launch_bounds__(256) k_sell_spmv(i64 nrows_pad, SellView S1, SellView S2, const double *__restrict__ diag,
                                                  const double *__restrict__ v, const double *y_in, double *y_out,
                                                  double alpha, double beta, double *xadd, SellView S3, const i32 *__restrict__ rowmap)
{
  // rowmap (restriction only): the SELL rows are stored sorted by length; rowmap[row] is the output row (-1 = padding)
  const i64 row = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= nrows_pad) return;
  const i64 slice = row >> 5;
  const int lane = threadIdx.x & 31;
  double acc[BH];
#pragma unroll
  for (int p = 0; p < BH; p++) acc[p] = 0.0;
  // scalar case without a row map: y_in is requested up front (it is only needed by the epilogue, after two dependent latencies)
  double yin0 = 0.0;
  if (U != 0 && BH == 1 && BW == 1 && !rowmap && beta != 0.0) yin0 = y_in[row];
  if (BH == 1 && BW == 1) {
    acc[0] = sell_row_dot1<false, U>(S1, slice, lane, v);
    if (HAS_S2) acc[0] += sell_row_dot1<false, U>(S2, slice, lane, v);
  } else {
    sell_row_mac<BH, BW, false>(S1, slice, lane, v, acc, 1.0);
    if (HAS_S2) sell_row_mac<BH, BW, false>(S2, slice, lane, v, acc, 1.0);
  }
  if (HAS_D && S3.slice_ptr) {   // couplings to non-free rows travel with the diagonal (see k_layout_count)
    if (BH == 1 && BW == 1) acc[0] += sell_row_dot1<false, 0>(S3, slice, lane, v);
    else sell_row_mac<BH, BW, false>(S3, slice, lane, v, acc, 1.0);
  }
  if (HAS_D || xadd) {
    double vi[BW];
#pragma unroll
    for (int q = 0; q < BW; q++) vi[q] = v[row * BW + q];
    if (HAS_D) {
      const double *dp = diag + slice * (i64)(BH * BW) * 32 + lane;
#pragma unroll
      for (int p = 0; p < BH; p++)
#pragma unroll
        for (int q = 0; q < BW; q++) acc[p] = fma(dp[(p * BW + q) * 32], vi[q], acc[p]);
    }
    if (xadd) {
#pragma unroll
      for (int q = 0; q < BW; q++) xadd[row * BW + q] += vi[q];
    }
  }
  i64 orow = row;
  if (rowmap) { orow = rowmap[row]; if (orow < 0) return; }
#pragma unroll
  for (int p = 0; p < BH; p++) {
    double y = alpha * acc[p];
    if (beta != 0.0) y = fma(beta, (U != 0 && BH == 1 && BW == 1 && !rowmap) ? yin0 : y_in[orow * BH + p], y);
    y_out[orow * BH + p] = y;
  }
}

// Same contract as k_sell_spmv for SMALL levels (coarse levels: few rows, wide rows): one WARP per row, the lanes split
// the row's entries (all parts), fixed shuffle tree.  Removes the long per-thread chains that make the thread-per-row kernel
// latency bound when a level has fewer rows than the machine has threads.
template <int BH, int BW, bool HAS_S2, bool HAS_D>
__global__ void __launch_bounds__(256) k_sell_spmv_small(i64 nrows_pad, SellView S1, SellView S2, const double *__restrict__ diag,
                                                        const double *__restrict__ v, const double *y_in, double *y_out,
                                                        double alpha, double beta, double *xadd, SellView S3,
                                                        const i32 *__restrict__ rowmap)
{
  const int lane = threadIdx.x & 31;
  const i64 gw = ((i64)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const i64 nw = ((i64)gridDim.x * blockDim.x) >> 5;
  for (i64 row = gw; row < nrows_pad; row += nw) {
    const i64 slice = row >> 5;
    const int lr = (int)(row & 31);
    double acc[BH];
#pragma unroll
    for (int p = 0; p < BH; p++) acc[p] = 0.0;
    auto part = [&](const SellView &S) {
      const i64 base = S.slice_ptr[slice];
      const int width = (int)(S.slice_ptr[slice + 1] - base);
      for (int k = lane; k < width; k += 32) {
        const i32 c = S.col[(base + k) * 32 + lr];
        if (c < 0) continue;
        double xv[BW];
#pragma unroll
        for (int q = 0; q < BW; q++) xv[q] = v[(i64)c * BW + q];
#pragma unroll
        for (int p = 0; p < BH; p++) {
          double t = 0.0;
#pragma unroll
          for (int q = 0; q < BW; q++) t = fma(S.val[((base + k) * (i64)(BH * BW) + p * BW + q) * 32 + lr], xv[q], t);
          acc[p] += t;
        }
      }
    };
    part(S1);
    if (HAS_S2) part(S2);
    if (HAS_D && S3.slice_ptr) part(S3);
#pragma unroll
    for (int p = 0; p < BH; p++)
#pragma unroll
      for (int o = 16; o; o >>= 1) acc[p] += __shfl_xor_sync(0xffffffffu, acc[p], o);
    if (lane != 0) continue;
    if (HAS_D || xadd) {
      double vi[BW];
#pragma unroll
      for (int q = 0; q < BW; q++) vi[q] = v[row * BW + q];
      if (HAS_D) {
        const double *dp = diag + slice * (i64)(BH * BW) * 32 + lr;
#pragma unroll
        for (int p = 0; p < BH; p++)
#pragma unroll
          for (int q = 0; q < BW; q++) acc[p] = fma(dp[(p * BW + q) * 32], vi[q], acc[p]);
      }
      if (xadd) {
#pragma unroll
        for (int q = 0; q < BW; q++) xadd[row * BW + q] += vi[q];
      }
    }
    i64 orow = row;
    if (rowmap) { orow = rowmap[row]; if (orow < 0) continue; }
#pragma unroll
    for (int p = 0; p < BH; p++) {
      double y = alpha * acc[p];
      if (beta != 0.0) y = fma(beta, y_in[orow * BH + p], y);
      y_out[orow * BH + p] = y;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// K2/K3: triangular half-sweep of Gauss-Seidel in the reference's row order -- sync-free, "data is the flag".
//   out_i  = (ADD_SELF ? self_i : 0) + dinv_i * ( rin_i - sum_{k in T_i} A_ik out_k )      [T = L forward, U backward]
//   rout_i = ( rin_i - sum_T A_ik out_k ) - D_ii * delta_i                               [RES form only]
// `out` is filled with a sentinel (all-ones NaN) before the launch; a row polls the entries of `out` it depends on
// until they are no longer the sentinel, computes, and publishes its own value with a plain 8-byte store: no
// barriers, counters, atomics or fences anywhere.  Rows are stored in dependency-level order, so the rows of a
// slice (= warp) are mutually independent and their dependencies lie in earlier slices.  Slices are dealt to the
// resident warps round-robin in sweep order; every warp walks its slices in that order, hence the earliest
// unfinished slice always has all its dependencies finished and the sweep cannot deadlock as long as the whole grid
// is resident (the host sizes it from the occupancy calculator).  Each warp loads its matrix entries BEFORE it starts
// polling, so HBM latency is off the dependency chain; the chain costs one L2 store->load hop per dependency level.
// ------------------------------------------------------------------------------------------------
// m = flavour of the polling load (flag ngs_amg_b200_tri_pollmode; a per-handle kernel parameter)
__device__ __forceinline__ double ld_poll(const double *p, int m)
{
  double v;
  if (m == 0) asm volatile("ld.relaxed.gpu.global.f64 %0, [%1];" : "=d"(v) : "l"(p));
  else if (m == 1) asm volatile("ld.volatile.global.f64 %0, [%1];" : "=d"(v) : "l"(p));
  else if (m == 2) asm volatile("ld.acquire.gpu.global.f64 %0, [%1];" : "=d"(v) : "l"(p));
  else asm volatile("ld.global.cv.f64 %0, [%1];" : "=d"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ bool is_sentinel(double v) { return __double_as_longlong(v) == -1LL; }
// watchdog of every spin loop: gives up after `limit` polls, or as soon as ANY waiter of the launch has given up (the flag is read
// every 1024 polls), so one timeout bounds the whole kernel instead of every dependent row spinning to its own limit.
__device__ __forceinline__ bool spin_fail(unsigned &spins, int *err, unsigned limit = (1u << 26))
{
  if ((++spins & 1023u) != 0) return false;
  if (spins > limit || *(volatile int *)err != 0) { atomicExch(err, 1); return true; }
  return false;
}

struct TriParams {
  i64 nslices;
  int backward;
  unsigned sleep_ns;  // back-off of the single-address pre-poll
  int prepoll;        // 1: gate on the latest dependency with one polling lane before the gather
  int gate_all;       // 1: gate every chunk on its newest entry, 0: only the last chunk of the row
  i64 gate_gap;       // > 0: early gate (see k_gs_tri), in rows
  unsigned repoll_ns; // back-off of the per-lane straggler polls
  int regate;         // 1: stragglers are waited for with a single polling lane (re-gate) instead of all lanes spinning
  i64 nonfree;        // rows [0, nonfree) are the non-free rows (dependency level 0)
  int *err;           // watchdog flag (set if a wait exceeds ~2^26 polls; never in a healthy run)
  int pollmode;       // flavour of the polling load (ld_poll)
  const i32 *bnd;     // split gate: per slice, first row of the previous level (forward) / end row of the next level (backward)
  unsigned long long *trace;  // debug: per slice {pick-up, gate passed, published} globaltimer stamps (NULL = off)
};

__device__ __forceinline__ unsigned long long gtimer()
{
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

template <int B, bool ADD_SELF, bool WRITE_R, int PRE>
__global__ void __launch_bounds__(256) k_gs_tri(SellView T, const double *__restrict__ diag, const double *__restrict__ dinv,
                                               const double *rin, const double *__restrict__ self, double *out,
                                               double *rout, TriParams prm)
{
  // PRE = slots cached in registers across the wait (scalar case only; the host picks 8, 12 or 16 from the slice widths)
  static_assert(B == 1 || PRE == 0, "register slot cache is implemented for scalar matrices");
  const int lane = threadIdx.x & 31;
  const i64 gw = ((i64)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const i64 nw = ((i64)gridDim.x * blockDim.x) >> 5;
  for (i64 s = gw; s < prm.nslices; s += nw) {
    const i64 slice = prm.backward ? (prm.nslices - 1 - s) : s;
    const i64 row = slice * 32 + lane;
    const i64 base = T.slice_ptr[slice];
    const int width = (int)(T.slice_ptr[slice + 1] - base);
    const i32 *cp = T.col + base * 32 + lane;
    const double *vp = T.val + base * (i64)(B * B) * 32 + lane;
    // a non-free row is never updated (dinv = 0) and the increments of non-free rows are zero, so couplings between two
    // non-free rows contribute nothing: skip them (they would otherwise chain the boundary rows one after another)
    const i32 cut = (row < prm.nonfree) ? (i32)prm.nonfree : 0;
    if (prm.trace && lane == 0) prm.trace[slice * 3 + 0] = gtimer();
    // ---- everything that does not depend on `out`: own rhs, diagonal blocks, matrix entries
    double acc[B];
#pragma unroll
    for (int p = 0; p < B; p++) acc[p] = rin[row * B + p];
    double sv[B];
    if (ADD_SELF) {
#pragma unroll
      for (int p = 0; p < B; p++) sv[p] = self[row * B + p];
    }
    const double *dp = dinv + slice * (i64)(B * B) * 32 + lane;
    double di0 = 0.0;
    if (B == 1) di0 = dp[0];
    i32 pc[PRE > 0 ? PRE : 1];
    double pv[PRE > 0 ? PRE : 1];
    if (PRE > 0) {
#pragma unroll
      for (int k = 0; k < PRE; k++) {
        pc[k] = (k < width) ? cp[(i64)k * 32] : -1;
        if (pc[k] < cut) pc[k] = -1;
        pv[k] = (k < width) ? vp[(i64)k * 32] : 0.0;
      }
    }
    if (width > PRE) {
      // entries that do not fit the register cache: pull their lines into L1 now, they are consumed after the wait
      const char *cb = (const char *)(T.col + (base + PRE) * 32);
      const char *vb = (const char *)(T.val + (base + PRE) * (i64)(B * B) * 32);
      const int nlc = width - PRE, nlv = (width - PRE) * B * B * 2;  // 128 B lines
      for (int l = lane; l < nlc; l += 32) asm volatile("prefetch.global.L1 [%0];" ::"l"(cb + (i64)l * 128));
      for (int l = lane; l < nlv; l += 32) asm volatile("prefetch.global.L1 [%0];" ::"l"(vb + (i64)l * 128));
    }
    constexpr int CH = (B == 1) ? 8 : (B == 2 ? 4 : (B == 3 ? 2 : 1));
    // Entries are sorted by age (oldest dependency first), so within every chunk of slots the dependency published last
    // is the chunk's last valid entry.  Before a chunk is gathered ONE lane gates on the warp's newest entry of that
    // chunk with back-off (1 sector per poll instead of 32 x chunk); after the gate the chunk's polls are issued
    // together and a straggler is simply re-polled.  Old chunks pass their gate at once, so a wide row is gathered
    // while the sweep is still waiting for the row's newest dependencies, without hammering L2 during the wait.
    const int last0 = (width <= PRE) ? 0 : PRE + ((width - PRE - 1) / CH) * CH;   // first slot of the last chunk
    // gate(mylast, k0, pick): one lane polls (with back-off) a single entry before the chunk's polls are issued.
    // gate_gap == 0: the entry is the warp's newest dependency of the chunk.  gate_gap > 0 ("early gate"): the newest entry
    // that is at least gate_gap rows older than that one, i.e. about one dependency level older; the chunk's own polls then
    // spin (per lane) for the last hop, which removes one L2 round trip from the dependency chain.
    auto gate = [&](i32 mylast, int k0, auto pick) {
      if (!prm.prepoll || (!prm.gate_all && k0 != last0)) return;
      i32 f = (mylast >= cut && (mylast >> 5) != slice) ? mylast : -1;
      if (prm.backward) { if (f < 0) f = 0x7fffffff; }
#pragma unroll
      for (int o = 16; o; o >>= 1) {
        const i32 g = __shfl_xor_sync(0xffffffffu, f, o);
        f = prm.backward ? min(f, g) : max(f, g);
      }
      if (prm.backward && f == 0x7fffffff) f = -1;
      if (f >= 0 && prm.gate_gap > 0) {
        i32 e = pick(f);
        if (prm.backward) { if (e < 0) e = 0x7fffffff; }
#pragma unroll
        for (int o = 16; o; o >>= 1) {
          const i32 g = __shfl_xor_sync(0xffffffffu, e, o);
          e = prm.backward ? min(e, g) : max(e, g);
        }
        if (prm.backward && e == 0x7fffffff) e = -1;
        if (e >= 0) f = e;
      }
      if (f >= 0 && lane == 0) {
        unsigned spins = 0;
        while (is_sentinel(ld_poll(out + (i64)f * B, prm.pollmode))) {
          if (prm.sleep_ns) __nanosleep(prm.sleep_ns);
          if (spin_fail(spins, prm.err)) break;
        }
      }
      __syncwarp();
    };
    // is entry c at least gate_gap rows older than the warp's newest dependency f ?
    auto old_enough = [&](i32 c, i32 f) {
      return c >= cut && (prm.backward ? ((i64)c >= (i64)f + prm.gate_gap) : ((i64)c <= (i64)f - prm.gate_gap));
    };
    if (PRE > 0) {
      // split gate (prm.bnd): entries in the dependency level processed just before this one ("new") are polled directly by
      // all lanes -- their poll result IS the data, one L2 round trip per hop -- while the single-lane gate with back-off
      // waits for the newest OLDER entry, which is published about one hop earlier.
      const bool split = prm.bnd != nullptr && width <= PRE;
      const i32 bnd = split ? prm.bnd[slice] : 0;
      auto is_new = [&](i32 c) { return split && c >= 0 && (prm.backward ? (c < bnd) : (c >= bnd)); };
      {
        i32 mylast = -1;
#pragma unroll
        for (int k = 0; k < PRE; k++) if (pc[k] >= 0 && !is_new(pc[k])) mylast = pc[k];
        gate(mylast, 0, [&](i32 f) {
          i32 e = -1;
#pragma unroll
          for (int k = 0; k < PRE; k++) if (old_enough(pc[k], f)) e = pc[k];   // sorted by age: the last hit is the newest
          return e;
        });
      }
      if (prm.trace && lane == 0) prm.trace[slice * 3 + 1] = gtimer();
      double xk[PRE > 0 ? PRE : 1];
#pragma unroll
      for (int k = 0; k < PRE; k++) xk[k] = (pc[k] >= 0) ? ld_poll(out + pc[k], prm.pollmode) : 0.0;
      if (prm.regate) {
        // stragglers (the wavefront is not perfectly ordered): instead of spinning with all lanes, gate again with one
        // lane on an entry that is still unpublished, then re-poll only the missing entries
        for (unsigned round = 0;; round++) {
          i32 u = -1;
#pragma unroll
          for (int k = 0; k < PRE; k++) if (pc[k] >= 0 && !is_new(pc[k]) && is_sentinel(xk[k])) u = pc[k];
          if (prm.backward && u < 0) u = 0x7fffffff;
#pragma unroll
          for (int o = 16; o; o >>= 1) {
            const i32 g = __shfl_xor_sync(0xffffffffu, u, o);
            u = prm.backward ? min(u, g) : max(u, g);
          }
          if (prm.backward && u == 0x7fffffff) u = -1;
          if (u < 0) break;
          if (lane == 0 && (u >> 5) != slice) {
            unsigned spins = 0;
            while (is_sentinel(ld_poll(out + u, prm.pollmode))) {
              if (prm.sleep_ns) __nanosleep(prm.sleep_ns);
              if (spin_fail(spins, prm.err)) break;
            }
          }
          __syncwarp();
#pragma unroll
          for (int k = 0; k < PRE; k++) if (pc[k] >= 0 && is_sentinel(xk[k])) xk[k] = ld_poll(out + pc[k], prm.pollmode);
          if (round > (1u << 22) || *(volatile int *)prm.err) { atomicExch(prm.err, 1); break; }
        }
      }
      if (split) {
        for (unsigned it = 0;; it++) {
          bool un = false;
#pragma unroll
          for (int k = 0; k < PRE; k++)
            if (pc[k] >= 0 && is_sentinel(xk[k])) { xk[k] = ld_poll(out + pc[k], prm.pollmode); un |= is_sentinel(xk[k]); }
          if (!__any_sync(0xffffffffu, un)) break;
          if ((it & 1023u) == 1023u && (it > (1u << 26) || *(volatile int *)prm.err)) { atomicExch(prm.err, 1); break; }
        }
      }
#pragma unroll
      for (int k = 0; k < PRE; k++) {
        if (pc[k] >= 0) {
          unsigned spins = 0;
          while (is_sentinel(xk[k])) {
            if (prm.repoll_ns) __nanosleep(prm.repoll_ns);
            xk[k] = ld_poll(out + pc[k], prm.pollmode);
            if (spin_fail(spins, prm.err)) break;
          }
          acc[0] = fma(-pv[k], xk[k], acc[0]);
        }
      }
    }
    for (int k0 = PRE; k0 < width; k0 += CH) {
      i32 c[CH];
      double xv[CH][B];
#pragma unroll
      for (int j = 0; j < CH; j++) {
        c[j] = (k0 + j < width) ? cp[(i64)(k0 + j) * 32] : -1;
        if (c[j] < cut) c[j] = -1;
      }
      double av[CH][B * B];   // the chunk's matrix blocks, loaded before the polls so their latency overlaps
#pragma unroll
      for (int j = 0; j < CH; j++)
#pragma unroll
        for (int e = 0; e < B * B; e++) av[j][e] = (k0 + j < width) ? vp[((i64)(k0 + j) * (B * B) + e) * 32] : 0.0;
      {
        i32 mylast = -1;
#pragma unroll
        for (int j = 0; j < CH; j++) if (c[j] >= 0) mylast = c[j];
        gate(mylast, k0, [&](i32 f) {
          i32 e = -1;
#pragma unroll
          for (int j = 0; j < CH; j++) if (old_enough(c[j], f)) e = c[j];
          return e;
        });
      }
#pragma unroll
      for (int j = 0; j < CH; j++)
        if (c[j] >= 0) {
#pragma unroll
          for (int q = 0; q < B; q++) xv[j][q] = ld_poll(out + (i64)c[j] * B + q, prm.pollmode);
        }
#pragma unroll
      for (int j = 0; j < CH; j++)
        if (c[j] >= 0) {
#pragma unroll
          for (int q = 0; q < B; q++) {
            unsigned spins = 0;
            while (is_sentinel(xv[j][q])) {
              if (prm.repoll_ns) __nanosleep(prm.repoll_ns);
              xv[j][q] = ld_poll(out + (i64)c[j] * B + q, prm.pollmode);
              if (spin_fail(spins, prm.err)) break;
            }
          }
#pragma unroll
          for (int p = 0; p < B; p++) {
            double t = 0.0;
#pragma unroll
            for (int q = 0; q < B; q++) t = fma(av[j][p * B + q], xv[j][q], t);
            acc[p] -= t;
          }
        }
    }
    double dl[B];
    if (B == 1) dl[0] = di0 * acc[0];
    else {
#pragma unroll
      for (int p = 0; p < B; p++) {
        double t = 0.0;
#pragma unroll
        for (int q = 0; q < B; q++) t = fma(dp[(p * B + q) * 32], acc[q], t);
        dl[p] = t;
      }
    }
#pragma unroll
    for (int p = 0; p < B; p++) __stcg(out + row * B + p, ADD_SELF ? sv[p] + dl[p] : dl[p]);
    if (prm.trace && lane == 0) prm.trace[slice * 3 + 2] = gtimer();
    if (WRITE_R) {
      const double *gp = diag + slice * (i64)(B * B) * 32 + lane;
#pragma unroll
      for (int p = 0; p < B; p++) {
        double t = acc[p];
#pragma unroll
        for (int q = 0; q < B; q++) t = fma(-gp[(p * B + q) * 32], dl[q], t);
        rout[row * B + p] = t;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Level-by-level variant of the triangular half-sweep for SHALLOW dependency DAGs (a handful of colours, many rows per
// colour): one launch per dependency level over the rows [row0, row1) of that level; plain cached gathers, no polling.
// Used when the depth is small enough that ~depth launches cost less than the polling traffic of the sync-free sweep.
// ------------------------------------------------------------------------------------------------
// PDL = launched with programmatic stream serialization behind the previous colour's launch: the kernel starts while that one is
// still running, lets ITS successor start (launch_dependents), pulls everything that does not depend on the sweep into registers / L1
// (row pointers, the first 8 entries, the rest of the slice by prefetch, rhs, dinv) and only then waits for the previous colour
// (griddepcontrol.wait = cudaGridDependencySynchronize).  A colour of a shallow level is a 10-20 us kernel: without the overlap half
// of it is launch latency and the dependent load chain in front of the first gather.
template <int B, bool ADD_SELF, bool WRITE_R, bool PDL>
__global__ void __launch_bounds__(256) k_gs_level(SellView T, const double *__restrict__ diag, const double *__restrict__ dinv,
                                                 const double *rin, const double *__restrict__ self, double *out, double *rout,
                                                 i64 row0, i64 row1, i64 nonfree)
{
  if (PDL) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const i64 row = row0 + (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= row1) return;
  if (ADD_SELF && row < nonfree) {
    // non-free rows are never updated (dinv = 0); their couplings to other non-free rows of the same launch must not be read
#pragma unroll
    for (int p = 0; p < B; p++) out[row * B + p] = self[row * B + p];
    return;
  }
  const i64 slice = row >> 5;
  const int lane = (int)(row & 31);
  double acc[B];
#pragma unroll
  for (int p = 0; p < B; p++) acc[p] = rin[row * B + p];
  constexpr int PRE = (B == 1) ? 8 : 0;
  i32 pc[PRE > 0 ? PRE : 1];
  double pv[PRE > 0 ? PRE : 1];
  int width = 0;
  i64 base = 0;
  if (PDL) {
    base = T.slice_ptr[slice];
    width = (int)(T.slice_ptr[slice + 1] - base);
    if (B == 1) {
#pragma unroll
      for (int k = 0; k < PRE; k++) {
        pc[k] = ldp_nc_i32(T.col + (base + k) * 32 + lane, k < width);
        pv[k] = ldp_nc_f64(T.val + (base + k) * 32 + lane, k < width);
      }
    }
    // the rest of the slice: one 128-byte line of columns and B*B*2 lines of values per slot
    const char *cb = (const char *)(T.col + (base + PRE) * 32);
    const char *vb = (const char *)(T.val + (base + PRE) * (i64)(B * B) * 32);
    const int nlc = max(width - PRE, 0), nlv = nlc * B * B * 2;
    for (int l = lane; l < nlc; l += 32) asm volatile("prefetch.global.L1 [%0];" ::"l"(cb + (i64)l * 128));
    for (int l = lane; l < nlv; l += 32) asm volatile("prefetch.global.L1 [%0];" ::"l"(vb + (i64)l * 128));
  }
  const double *dp = dinv + slice * (i64)(B * B) * 32 + lane;
  double dv[B * B];
#pragma unroll
  for (int e = 0; e < B * B; e++) dv[e] = dp[e * 32];
  double sv[B];
#pragma unroll
  for (int p = 0; p < B; p++) sv[p] = ADD_SELF ? self[row * B + p] : 0.0;
  if (PDL) asm volatile("griddepcontrol.wait;" ::: "memory");
  if (PDL && B == 1) {
    double s0 = 0.0, s1 = 0.0;
    double xk[PRE > 0 ? PRE : 1];
#pragma unroll
    for (int k = 0; k < PRE; k++) xk[k] = ldp_f64(out + (pc[k] >= 0 ? pc[k] : 0), pc[k] >= 0, true);
#pragma unroll
    for (int k = 0; k < PRE; k += 2) { s0 = fma(pv[k], xk[k], s0); if (k + 1 < PRE) s1 = fma(pv[k + 1], xk[k + 1], s1); }
    const i32 *cp = T.col + base * 32 + lane;
    const double *vp = T.val + base * 32 + lane;
    for (int k = PRE; k < width; k += 4) {
      i32 c[4];
      double v[4], xv[4];
#pragma unroll
      for (int j = 0; j < 4; j++) c[j] = ldp_nc_i32(cp + (i64)(k + j) * 32, k + j < width);
#pragma unroll
      for (int j = 0; j < 4; j++) v[j] = ldp_nc_f64(vp + (i64)(k + j) * 32, k + j < width);
#pragma unroll
      for (int j = 0; j < 4; j++) xv[j] = ldp_f64(out + (c[j] >= 0 ? c[j] : 0), c[j] >= 0, true);
      s0 = fma(v[0], xv[0], s0); s1 = fma(v[1], xv[1], s1); s0 = fma(v[2], xv[2], s0); s1 = fma(v[3], xv[3], s1);
    }
    acc[0] -= s0 + s1;
  } else if (B == 1) acc[0] -= sell_row_dot1<false>(T, slice, lane, out);
  else sell_row_mac<B, B, PDL>(T, slice, lane, out, acc, -1.0);
  double dl[B];
#pragma unroll
  for (int p = 0; p < B; p++) {
    double t = 0.0;
#pragma unroll
    for (int q = 0; q < B; q++) t = fma(dv[p * B + q], acc[q], t);
    dl[p] = t;
  }
#pragma unroll
  for (int p = 0; p < B; p++) out[row * B + p] = ADD_SELF ? sv[p] + dl[p] : dl[p];
  if (WRITE_R) {
    const double *gp = diag + slice * (i64)(B * B) * 32 + lane;
#pragma unroll
    for (int p = 0; p < B; p++) {
      double t = acc[p];
#pragma unroll
      for (int q = 0; q < B; q++) t = fma(-gp[(p * B + q) * 32], dl[q], t);
      rout[row * B + p] = t;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// K4-like variant of the triangular half-sweep for SMALL levels (coarse levels, wide rows, few rows per dependency level):
// one WARP per row, the lanes split the row's entries, a fixed shuffle tree sums the partial products (deterministic).
// The chain then costs one L2 poll + one warp reduction per dependency level, independent of the row width.
// Same sync-free protocol as k_gs_tri (sentinel-filled output, rows dealt round-robin in sweep order).
// ------------------------------------------------------------------------------------------------
template <int B, bool ADD_SELF, bool WRITE_R>
__global__ void __launch_bounds__(256) k_gs_tri_small(SellView T, const double *__restrict__ diag, const double *__restrict__ dinv,
                                                     const double *rin, const double *__restrict__ self, double *out, double *rout,
                                                     TriParams prm)
{
  const int lane = threadIdx.x & 31;
  const i64 gw = ((i64)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const i64 nw = ((i64)gridDim.x * blockDim.x) >> 5;
  const i64 nrows = prm.nslices * 32;
  for (i64 r = gw; r < nrows; r += nw) {
    const i64 row = prm.backward ? (nrows - 1 - r) : r;
    const i64 slice = row >> 5;
    const int lr = (int)(row & 31);
    const i64 base = T.slice_ptr[slice];
    const int width = (int)(T.slice_ptr[slice + 1] - base);
    const i32 cut = (row < prm.nonfree) ? (i32)prm.nonfree : 0;
    double acc[B];
#pragma unroll
    for (int p = 0; p < B; p++) acc[p] = 0.0;
    // warp-uniform control flow (uniform trip count, votes instead of per-lane spin loops): after a divergent spin loop the warp reaches
    // the shuffle tree split into groups and every SHFL takes the slow collective path (measured in k_gs_tri_rm: 1.3 us instead of 0.09 us)
    for (int k0 = 0; k0 < width; k0 += 32) {
      const int k = k0 + lane;
      i32 c = (k < width) ? T.col[(base + k) * 32 + lr] : -1;
      if (c < cut) c = -1;
      double a[B * B];
#pragma unroll
      for (int e = 0; e < B * B; e++) a[e] = (c >= 0) ? T.val[((base + k) * (i64)(B * B) + e) * 32 + lr] : 0.0;
      double xv[B];
#pragma unroll
      for (int q = 0; q < B; q++) xv[q] = (c >= 0) ? ld_poll(out + (i64)c * B + q, prm.pollmode) : 0.0;
      unsigned spins = 0;
      for (;;) {
        bool miss = false;
#pragma unroll
        for (int q = 0; q < B; q++) miss |= is_sentinel(xv[q]);
        if (!__any_sync(0xffffffffu, miss)) break;
        if (prm.sleep_ns) __nanosleep(prm.sleep_ns);
#pragma unroll
        for (int q = 0; q < B; q++)
          if (is_sentinel(xv[q])) xv[q] = ld_poll(out + (i64)c * B + q, prm.pollmode);
        const bool fail = spin_fail(spins, prm.err);
        if (__any_sync(0xffffffffu, fail)) break;
      }
#pragma unroll
      for (int p = 0; p < B; p++) {
        double t = 0.0;
#pragma unroll
        for (int q = 0; q < B; q++) t = fma(a[p * B + q], xv[q], t);
        acc[p] -= t;
      }
    }
    __syncwarp();
#pragma unroll
    for (int p = 0; p < B; p++)
#pragma unroll
      for (int o = 16; o; o >>= 1) acc[p] += __shfl_xor_sync(0xffffffffu, acc[p], o);
    if (lane == 0) {
#pragma unroll
      for (int p = 0; p < B; p++) acc[p] += rin[row * B + p];
      const double *dp = dinv + slice * (i64)(B * B) * 32 + lr;
      double dl[B];
#pragma unroll
      for (int p = 0; p < B; p++) {
        double t = 0.0;
#pragma unroll
        for (int q = 0; q < B; q++) t = fma(dp[(p * B + q) * 32], acc[q], t);
        dl[p] = t;
      }
#pragma unroll
      for (int p = 0; p < B; p++) __stcg(out + row * B + p, ADD_SELF ? self[row * B + p] + dl[p] : dl[p]);
      if (WRITE_R) {
        const double *gp = diag + slice * (i64)(B * B) * 32 + lr;
#pragma unroll
        for (int p = 0; p < B; p++) {
          double t = acc[p];
#pragma unroll
          for (int q = 0; q < B; q++) t = fma(-gp[(p * B + q) * 32], dl[q], t);
          rout[row * B + p] = t;
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// layout construction (K15): plain CSR (original numbering) -> permuted SELL-32 split L / D / U
// ------------------------------------------------------------------------------------------------
// pass 1: per permuted row, number of entries going to S1 (lower, or everything if !SPLIT) and S2 (upper)
__global__ void k_layout_count(i64 n, const i64 *__restrict__ rowptr, const i32 *__restrict__ col,
                               const i32 *__restrict__ rperm, const i32 *__restrict__ cperm, int split, i32 *len1, i32 *len2,
                               i32 nonfree, i32 *len3)
{
  // split: S1 = strictly lower, S2 = strictly upper, S3 = couplings of a free row to non-free rows (rows [0,nonfree) of the
  // level-scheduled numbering).  Non-free values never change during a sweep, so S3 is applied with the parallel half.
  const i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const i32 pi = rperm[i];
  int n1 = 0, n2 = 0, n3 = 0;
  for (i64 k = rowptr[i]; k < rowptr[i + 1]; k++) {
    const i32 c = col[k];
    if (split) {
      if (c == i) continue;
      const i32 pc = cperm[c];
      if (pi >= nonfree && pc < nonfree) n3++;
      else if (pc < pi) n1++;
      else n2++;
    } else n1++;
  }
  len1[pi] = n1;
  if (split) { len2[pi] = n2; if (len3) len3[pi] = n3; }
}

// pass 2: slice width = max row length in the slice (one warp per slice)
__global__ void k_layout_width(i64 nslices, const i32 *__restrict__ len, i64 *width)
{
  const i64 s = ((i64)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (s >= nslices) return;
  int w = len[s * 32 + (threadIdx.x & 31)];
  for (int o = 16; o; o >>= 1) w = max(w, __shfl_xor_sync(0xffffffffu, w, o));
  if ((threadIdx.x & 31) == 0) width[s] = w;
}

// pass 3: scatter entries (original column order is kept inside each part)
__global__ void k_layout_fill(i64 n, int bs, const i64 *__restrict__ rowptr, const i32 *__restrict__ col,
                              const double *__restrict__ val, const i32 *__restrict__ rperm, const i32 *__restrict__ cperm,
                              int split, const i64 *__restrict__ sp1, i32 *col1, double *val1, const i64 *__restrict__ sp2,
                              i32 *col2, double *val2, double *diag, i32 nonfree, const i64 *__restrict__ sp3, i32 *col3, double *val3)
{
  const i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const i32 pi = rperm[i];
  const i64 slice = pi >> 5;
  const int lane = pi & 31;
  i64 k1 = sp1[slice], k2 = split ? sp2[slice] : 0, k3 = (split && sp3) ? sp3[slice] : 0;
  for (i64 k = rowptr[i]; k < rowptr[i + 1]; k++) {
    const i32 c = col[k];
    const double *src = val + k * bs;
    if (split && c == i) {
      for (int e = 0; e < bs; e++) diag[(slice * bs + e) * 32 + lane] = src[e];
      continue;
    }
    const i32 pc = cperm[c];
    if (split && sp3 && pi >= nonfree && pc < nonfree) {
      col3[k3 * 32 + lane] = pc;
      for (int e = 0; e < bs; e++) val3[(k3 * bs + e) * 32 + lane] = src[e];
      k3++;
    } else if (!split || pc < pi) {
      col1[k1 * 32 + lane] = pc;
      for (int e = 0; e < bs; e++) val1[(k1 * bs + e) * 32 + lane] = src[e];
      k1++;
    } else {
      col2[k2 * 32 + lane] = pc;
      for (int e = 0; e < bs; e++) val2[(k2 * bs + e) * 32 + lane] = src[e];
      k2++;
    }
  }
}

// pass 4 (L and U only): order the entries of every row by the age of the dependency in the sweep that uses the part --
// ascending row number for L (forward sweep), descending for U (backward sweep) -- so that the dependency published
// LAST sits in the last valid slot and everything before it can be gathered while the sweep is still waiting for it.
// In-place insertion sort over the slots of the row (rows are short; setup only).
__global__ void k_sell_sort_rows(i64 nrows_pad, int bs, const i64 *__restrict__ sp, i32 *col, double *val, int descending)
{
  const i64 row = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= nrows_pad) return;
  const i64 slice = row >> 5;
  const int lane = row & 31;
  const i64 base = sp[slice];
  const int width = (int)(sp[slice + 1] - base);
  i32 *cp = col + base * 32 + lane;
  double *vp = val + base * (i64)bs * 32 + lane;
  int len = 0;
  while (len < width && cp[(i64)len * 32] >= 0) len++;
  for (int a = 1; a < len; a++) {
    const i32 ck = cp[(i64)a * 32];
    int q = a;
    while (q > 0) {
      const i32 cq = cp[(i64)(q - 1) * 32];
      const bool before = descending ? (cq < ck) : (cq > ck);   // entry q-1 must move behind the key
      if (!before) break;
      q--;
    }
    if (q == a) continue;
    // rotate slots [q, a] right by one, element by element (values are planar with stride 32)
    for (int e = 0; e < bs; e++) {
      const double key = vp[((i64)a * bs + e) * 32];
      for (int m = a; m > q; m--) vp[((i64)m * bs + e) * 32] = vp[((i64)(m - 1) * bs + e) * 32];
      vp[((i64)q * bs + e) * 32] = key;
    }
    for (int m = a; m > q; m--) cp[(i64)m * 32] = cp[(i64)(m - 1) * 32];
    cp[(i64)q * 32] = ck;
  }
}

// K13: dinv = inverse of the diagonal block (GSS3::CalcDiags, gssmoother.cpp:142-170); 0 on non-free / padding rows.
// Gauss-Jordan with partial pivoting in registers, one thread per block row.  err[0] set on a singular block.
template <int B>
__global__ void k_calc_dinv(i64 nrows_pad, const double *__restrict__ diag, const uint8_t *__restrict__ free_p, double *dinv, int *err)
{
  const i64 row = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= nrows_pad) return;
  const i64 slice = row >> 5;
  const int lane = row & 31;
  const double *dp = diag + slice * (i64)(B * B) * 32 + lane;
  double *op = dinv + slice * (i64)(B * B) * 32 + lane;
  if (!free_p[row]) {
#pragma unroll
    for (int e = 0; e < B * B; e++) op[e * 32] = 0.0;
    return;
  }
  double a[B][B], inv[B][B];
#pragma unroll
  for (int p = 0; p < B; p++)
#pragma unroll
    for (int q = 0; q < B; q++) { a[p][q] = dp[(p * B + q) * 32]; inv[p][q] = (p == q) ? 1.0 : 0.0; }
  bool sing = false;
#pragma unroll
  for (int c = 0; c < B; c++) {
    int piv = c;
    double best = fabs(a[c][c]);
#pragma unroll
    for (int r2 = c + 1; r2 < B; r2++)
      if (fabs(a[r2][c]) > best) { best = fabs(a[r2][c]); piv = r2; }
    if (best == 0.0) { sing = true; break; }
#pragma unroll
    for (int r2 = c + 1; r2 < B; r2++)
      if (r2 == piv) {
#pragma unroll
        for (int q = 0; q < B; q++) {
          double t = a[c][q]; a[c][q] = a[r2][q]; a[r2][q] = t;
          t = inv[c][q]; inv[c][q] = inv[r2][q]; inv[r2][q] = t;
        }
      }
    const double pinv = 1.0 / a[c][c];
#pragma unroll
    for (int q = 0; q < B; q++) { a[c][q] *= pinv; inv[c][q] *= pinv; }
#pragma unroll
    for (int r2 = 0; r2 < B; r2++) {
      if (r2 == c) continue;
      const double f = a[r2][c];
#pragma unroll
      for (int q = 0; q < B; q++) { a[r2][q] -= f * a[c][q]; inv[r2][q] -= f * inv[c][q]; }
    }
  }
  if (sing) {
    atomicExch(err, 1);
#pragma unroll
    for (int e = 0; e < B * B; e++) op[e * 32] = 0.0;
    return;
  }
#pragma unroll
  for (int p = 0; p < B; p++)
#pragma unroll
    for (int q = 0; q < B; q++) op[(p * B + q) * 32] = inv[p][q];
}

// planar <-> AoS block copies used for the host pseudo-inverse path and introspection
__global__ void k_planar_to_aos(i64 nrows_pad, int bs, const double *__restrict__ planar, double *aos)
{
  const i64 row = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= nrows_pad) return;
  for (int e = 0; e < bs; e++) aos[row * bs + e] = planar[((row >> 5) * bs + e) * 32 + (row & 31)];
}
__global__ void k_aos_to_planar(i64 nrows_pad, int bs, const double *__restrict__ aos, double *planar)
{
  const i64 row = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= nrows_pad) return;
  for (int e = 0; e < bs; e++) planar[((row >> 5) * bs + e) * 32 + (row & 31)] = aos[row * bs + e];
}

// ------------------------------------------------------------------------------------------------
// vector kernels (K7): permutation in/out of the level-scheduled numbering, axpy-type updates, dots
// ------------------------------------------------------------------------------------------------
__global__ void k_permute_in(i64 n, int b, const i32 *__restrict__ perm, const double *__restrict__ x, double *xp)
{
  const i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const i64 p = perm[i];
  for (int q = 0; q < b; q++) xp[p * b + q] = x[i * b + q];
}
// x = xp(perm)  or  x += s * xp(perm)
__global__ void k_permute_out(i64 n, int b, const i32 *__restrict__ perm, const double *__restrict__ xp, double *x, double s, int add)
{
  const i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const i64 p = perm[i];
  for (int q = 0; q < b; q++) {
    const double v = xp[p * b + q];
    x[i * b + q] = add ? fma(s, v, x[i * b + q]) : v;
  }
}

// y = a*x + b*y
__global__ void k_axpby(i64 n, double a, const double *__restrict__ x, double b, double *y)
{
  const i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) y[i] = (b == 0.0) ? a * x[i] : fma(a, x[i], b * y[i]);
}
// CG update: u += alpha*s ; d -= alpha*w   (one pass over four vectors)
__global__ void k_cg_update(i64 n, double alpha, const double *__restrict__ s, const double *__restrict__ w, double *u, double *d)
{
  const i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { u[i] = fma(alpha, s[i], u[i]); d[i] = fma(-alpha, w[i], d[i]); }
}
// x += omega * dinv * src   (Jacobi / Richardson step, base_smoother.cpp:61-74)
template <int B>
__global__ void k_jacobi_update(i64 nrows_pad, double omega, const double *__restrict__ dinv, const double *__restrict__ src, double *x)
{
  const i64 row = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= nrows_pad) return;
  const double *dp = dinv + (row >> 5) * (i64)(B * B) * 32 + (row & 31);
  double sv[B];
#pragma unroll
  for (int q = 0; q < B; q++) sv[q] = src[row * B + q];
#pragma unroll
  for (int p = 0; p < B; p++) {
    double s = 0.0;
#pragma unroll
    for (int q = 0; q < B; q++) s = fma(dp[(p * B + q) * 32], sv[q], s);
    x[row * B + p] = fma(omega, s, x[row * B + p]);
  }
}

// deterministic dot product: fixed grid, fixed-order tree inside a block, partials summed by k_dot_final
constexpr int DOT_BLOCKS = 592;  // 4 per SM
constexpr int DOT_THREADS = 256;
__global__ void __launch_bounds__(DOT_THREADS) k_dot_partial(i64 n, const double *__restrict__ a, const double *__restrict__ b, double *partial)
{
  __shared__ double sh[DOT_THREADS];
  double s = 0.0;
  for (i64 i = (i64)blockIdx.x * DOT_THREADS + threadIdx.x; i < n; i += (i64)gridDim.x * DOT_THREADS) s = fma(a[i], b[i], s);
  sh[threadIdx.x] = s;
  __syncthreads();
  for (int o = DOT_THREADS / 2; o; o >>= 1) {
    if ((int)threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) partial[blockIdx.x] = sh[0];
}
__global__ void __launch_bounds__(DOT_THREADS) k_dot_final(int nparts, const double *__restrict__ partial, double *out)
{
  __shared__ double sh[DOT_THREADS];
  double s = 0.0;
  for (int i = threadIdx.x; i < nparts; i += DOT_THREADS) s += partial[i];
  sh[threadIdx.x] = s;
  __syncthreads();
  for (int o = DOT_THREADS / 2; o; o >>= 1) {
    if ((int)threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) *out = sh[0];
}

// K9: coarsest-level exact solve as a dense GEMV with the explicit inverse (n <= a few thousand scalars).
// One warp per row, fixed-order shuffle reduction.
__global__ void k_dense_gemv(int n, const double *__restrict__ M, const double *__restrict__ x, double *y)
{
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= n) return;
  double s = 0.0;
  for (int c = lane; c < n; c += 32) s = fma(M[(i64)row * n + c], x[c], s);
  for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) y[row] = s;
}

// ------------------------------------------------------------------------------------------------
// multi-rank halo exchange (DCCMap, src/base/linalg/dcc_map.cpp:225-300): pack / unpack of the shared dofs.
// Index lists are in the level-scheduled numbering; one thread per (dof, component).
// ------------------------------------------------------------------------------------------------
// BufferG (zero = 1: buf = v, v = 0) / BufferM (zero = 0: buf = v)
__global__ void k_halo_pack(i64 cnt, int b, const i32 *__restrict__ idx, double *v, double *__restrict__ buf, int zero)
{
  const i64 t = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= cnt * b) return;
  const i64 k = t / b;
  const int q = (int)(t - k * b);
  const i64 p = (i64)idx[k] * b + q;
  buf[t] = v[p];
  if (zero) v[p] = 0.0;
}
// ApplyM: v[dof] += received values, neighbours in ascending order (a master dof can have several ghosts)
__global__ void k_halo_add(i64 nuniq, int b, const i32 *__restrict__ dof, const i64 *__restrict__ src_ptr,
                           const i64 *__restrict__ src_pos, const double *__restrict__ buf, double *v)
{
  const i64 t = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nuniq * b) return;
  const i64 k = t / b;
  const int q = (int)(t - k * b);
  double s = v[(i64)dof[k] * b + q];
  for (i64 e = src_ptr[k]; e < src_ptr[k + 1]; e++) s += buf[src_pos[e] * b + q];
  v[(i64)dof[k] * b + q] = s;
}
// ApplyG: v[dof] = received value
__global__ void k_halo_set(i64 cnt, int b, const i32 *__restrict__ idx, const double *__restrict__ buf, double *v)
{
  const i64 t = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= cnt * b) return;
  const i64 k = t / b;
  const int q = (int)(t - k * b);
  v[(i64)idx[k] * b + q] = buf[t];
}
// CtrMap transfers (dof_contract.cpp:49-228): dst[perm[map[i]]] += src[i]  /  dst[i] = src[perm[map[i]]]
__global__ void k_ctr_scatter_add(i64 n, int b, const i32 *__restrict__ map, const i32 *__restrict__ perm, const double *__restrict__ src, double *dst)
{
  const i64 t = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n * b) return;
  const i64 i = t / b;
  const int q = (int)(t - i * b);
  dst[(i64)perm[map[i]] * b + q] += src[t];
}
__global__ void k_ctr_gather(i64 n, int b, const i32 *__restrict__ map, const i32 *__restrict__ perm, const double *__restrict__ src, double *dst)
{
  const i64 t = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n * b) return;
  const i64 i = t / b;
  const int q = (int)(t - i * b);
  dst[t] = src[(i64)perm[map[i]] * b + q];
}
// AoS blocks in the original numbering -> planar blocks in the level-scheduled numbering (replacement diagonal of the hybrid smoother)
__global__ void k_aos_perm_to_planar(i64 n, int bs, const i32 *__restrict__ perm, const double *__restrict__ aos, double *planar)
{
  const i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const i64 row = perm[i];
  for (int e = 0; e < bs; e++) planar[((row >> 5) * bs + e) * 32 + (row & 31)] = aos[i * bs + e];
}

}  // namespace ngb
