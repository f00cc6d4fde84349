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

// scalar fast path: issue all loads of the row before the FMAs (more bytes in flight per thread)
template <bool CG>
__device__ __forceinline__ double sell_row_dot1(const SellView &S, i64 slice, int lane, const double *__restrict__ x)
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

// ------------------------------------------------------------------------------------------------
// K1/K2b/K3a/K5/K6: y_out = beta*y_in + alpha*(S1 v [+ S2 v] [+ D v]);  optionally xadd += v (own rows)
//   plain SpMV (S1=L,S2=U,D), the U-pass of the forward sweep, the (L+D)-pass of the backward sweep,
//   restriction (S1=PT) and prolongation-add (S1=P).  Thread per block row, warp per slice.
// ------------------------------------------------------------------------------------------------
template <int BH, int BW, bool HAS_S2, bool HAS_D>
__global__ void __launch_bounds__(256) k_sell_spmv(i64 nrows_pad, SellView S1, SellView S2, const double *__restrict__ diag,
                                                  const double *__restrict__ v, const double *y_in, double *y_out,
                                                  double alpha, double beta, double *xadd)
{
  const i64 row = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= nrows_pad) return;
  const i64 slice = row >> 5;
  const int lane = threadIdx.x & 31;
  double acc[BH];
#pragma unroll
  for (int p = 0; p < BH; p++) acc[p] = 0.0;
  if (BH == 1 && BW == 1) {
    acc[0] = sell_row_dot1<false>(S1, slice, lane, v);
    if (HAS_S2) acc[0] += sell_row_dot1<false>(S2, slice, lane, v);
  } else {
    sell_row_mac<BH, BW, false>(S1, slice, lane, v, acc, 1.0);
    if (HAS_S2) sell_row_mac<BH, BW, false>(S2, slice, lane, v, acc, 1.0);
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
#pragma unroll
  for (int p = 0; p < BH; p++) {
    double y = alpha * acc[p];
    if (beta != 0.0) y = fma(beta, y_in[row * BH + p], y);
    y_out[row * BH + p] = y;
  }
}

// ------------------------------------------------------------------------------------------------
// K2/K3: level-scheduled triangular half-sweep of Gauss-Seidel (persistent, dependency counters).
//   out_i = (ADD_SELF ? out_i : 0) + dinv_i * ( rin_i - sum_{k in T_i} A_ik out_k )        [T = L fwd, U bwd]
//   rout_i = ( rin_i - sum_T A_ik out_k ) - D_ii * delta_i                               [RES form only]
// Tiles = up to TILE_ROWS rows of ONE dependency level, taken in sweep order from a ticket counter;
// a tile waits until the previous level's tiles have all published (`done[level]` counters, release/acquire),
// so every claimed tile only waits on tiles that are already held by running CTAs (deadlock-free for any grid).
// The tile's matrix rows do not depend on `out`, so they are pulled towards L2 before the wait.
// ------------------------------------------------------------------------------------------------
struct TriSchedule {
  const i32 *tile_row0;   // first (padded) row of the tile
  const i32 *tile_rows;   // rows in the tile (multiple of 32)
  const i32 *tile_level;  // dependency level of the tile
  const i32 *level_tiles; // [nlevels] tiles per level
  i32 ntiles, nlevels;
  int *err;               // set to 1 if a dependency wait timed out (watchdog; never in a healthy run)
};

template <int B, bool ADD_SELF, bool WRITE_R>
__global__ void __launch_bounds__(256) k_gs_tri(SellView T, const double *__restrict__ diag, const double *__restrict__ dinv,
                                               const double *rin, double *out, double *rout, TriSchedule sch,
                                               int backward, int *counters /* [0]=ticket, [1+l]=done[l] */)
{
  __shared__ int s_tile;
  int *ticket = counters;
  int *done = counters + 1;
  const int lane = threadIdx.x & 31;
  for (;;) {
    if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1);
    __syncthreads();
    const int tk = s_tile;
    if (tk >= sch.ntiles) break;
    const int t = backward ? (sch.ntiles - 1 - tk) : tk;
    const int lvl = sch.tile_level[t];
    const i64 row = (i64)sch.tile_row0[t] + threadIdx.x;
    const bool active = (int)threadIdx.x < sch.tile_rows[t];
    const i64 slice = row >> 5;
    double r[B], self[B];
    if (active) {
      // warm L2 with this slice's entries while the dependency is still pending
      const i64 base = T.slice_ptr[slice];
      const int width = (int)(T.slice_ptr[slice + 1] - base);
      const char *cb = (const char *)(T.col + base * 32);
      const char *vb = (const char *)(T.val + base * (i64)(B * B) * 32);
      const int nlc = width;                 // 128 B lines of column indices
      const int nlv = width * B * B * 2;     // 128 B lines of values
      for (int l = lane; l < nlc; l += 32) asm volatile("prefetch.global.L2 [%0];" ::"l"(cb + (i64)l * 128));
      for (int l = lane; l < nlv; l += 32) asm volatile("prefetch.global.L2 [%0];" ::"l"(vb + (i64)l * 128));
#pragma unroll
      for (int p = 0; p < B; p++) r[p] = rin[row * B + p];
      if (ADD_SELF) {
#pragma unroll
        for (int p = 0; p < B; p++) self[p] = ld_cg(out + row * B + p);
      }
    }
    const int dep = backward ? lvl + 1 : lvl - 1;
    if (threadIdx.x == 0 && dep >= 0 && dep < sch.nlevels) {
      const int need = sch.level_tiles[dep];
      unsigned spins = 0;
      while (ld_acquire(done + dep) < need) {
        __nanosleep(20);
        if (++spins > (1u << 27)) { atomicExch(sch.err, 1); break; }  // watchdog: report instead of hanging the GPU
      }
    }
    __syncthreads();
    if (active) {
      double acc[B];
#pragma unroll
      for (int p = 0; p < B; p++) acc[p] = r[p];
      if (B == 1) acc[0] -= sell_row_dot1<true>(T, slice, lane, out);
      else sell_row_mac<B, B, true>(T, slice, lane, out, acc, -1.0);
      const double *dp = dinv + slice * (i64)(B * B) * 32 + lane;
      double dl[B];
#pragma unroll
      for (int p = 0; p < B; p++) {
        double s = 0.0;
#pragma unroll
        for (int q = 0; q < B; q++) s = fma(dp[(p * B + q) * 32], acc[q], s);
        dl[p] = s;
      }
#pragma unroll
      for (int p = 0; p < B; p++) out[row * B + p] = ADD_SELF ? self[p] + dl[p] : dl[p];
      if (WRITE_R) {
        const double *gp = diag + slice * (i64)(B * B) * 32 + lane;
#pragma unroll
        for (int p = 0; p < B; p++) {
          double s = acc[p];
#pragma unroll
          for (int q = 0; q < B; q++) s = fma(-gp[(p * B + q) * 32], dl[q], s);
          rout[row * B + p] = s;
        }
      }
    }
    __syncthreads();  // all stores of the tile issued (and s_tile free for reuse)
    if (threadIdx.x == 0) {
      __threadfence();
      red_release_add(done + lvl, 1);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// layout construction (K15): plain CSR (original numbering) -> permuted SELL-32 split L / D / U
// ------------------------------------------------------------------------------------------------
// pass 1: per permuted row, number of entries going to S1 (lower, or everything if !SPLIT) and S2 (upper)
__global__ void k_layout_count(i64 n, const i64 *__restrict__ rowptr, const i32 *__restrict__ col,
                               const i32 *__restrict__ rperm, const i32 *__restrict__ cperm, int split, i32 *len1, i32 *len2)
{
  const i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const i32 pi = rperm[i];
  int n1 = 0, n2 = 0;
  for (i64 k = rowptr[i]; k < rowptr[i + 1]; k++) {
    const i32 c = col[k];
    if (split) {
      if (c == i) continue;
      if (cperm[c] < pi) n1++; else n2++;
    } else n1++;
  }
  len1[pi] = n1;
  if (split) len2[pi] = n2;
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
                              i32 *col2, double *val2, double *diag)
{
  const i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const i32 pi = rperm[i];
  const i64 slice = pi >> 5;
  const int lane = pi & 31;
  i64 k1 = sp1[slice], k2 = split ? sp2[slice] : 0;
  for (i64 k = rowptr[i]; k < rowptr[i + 1]; k++) {
    const i32 c = col[k];
    const double *src = val + k * bs;
    if (split && c == i) {
      for (int e = 0; e < bs; e++) diag[(slice * bs + e) * 32 + lane] = src[e];
      continue;
    }
    const i32 pc = cperm[c];
    if (!split || pc < pi) {
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

}  // namespace ngb
