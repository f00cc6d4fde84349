// spgemm.cu -- Galerkin-product building block on the device: C = A * B for block-CSR matrices.
//
// Restates MatMultABImpl (src/base/linalg/utils_sparseMM.cpp:107-238) as a two-pass hash SpGEMM:
//   symbolic count  -> exclusive scan -> symbolic fill (+ per-row ascending sort) -> numeric.
// * The pattern of a row is the sorted set union of the B-rows named by the A-row, i.e. exactly what the
//   reference's k-way MergeArrays emits (structural: numerical zeros are kept) -> bit-exact integers.
// * The numeric phase walks the A-row sequentially and spreads each B-row over the lanes of a warp; columns
//   inside one B-row are distinct, so no two lanes ever touch the same output block at the same time and every
//   output block receives its contributions in the reference's order (A-row order, then B-row order).  Products
//   and sums are formed with explicit round-to-nearest mul/add (no FMA contraction), so the values agree
//   bit-for-bit with a non-contracted CPU evaluation of `C(i,col) += vala * valb`.
#include <cub/cub.cuh>

#include "device.hpp"

namespace ngb {

namespace {

constexpr int SPG_WARPS = 4;      // warps (rows) per CTA
constexpr int SPG_TABLE = 2048;   // shared-memory hash slots per warp
constexpr int SPG_EMPTY = -1;

__device__ __forceinline__ unsigned spg_hash(i32 c, unsigned mask) { return (((unsigned)c * 0x9E3779B1u) >> 11) & mask; }

// upper bound of the row size and the (power of two) size of the global hash table for rows that do not
// fit the shared-memory table
__global__ void k_spg_bound(i64 nA, const i64 *__restrict__ a_rp, const i32 *__restrict__ a_ci, const i64 *__restrict__ b_rp,
                            i64 *ub, i64 *gsize)
{
  const i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nA) return;
  i64 s = 0;
  for (i64 j = a_rp[i]; j < a_rp[i + 1]; j++) s += b_rp[a_ci[j] + 1] - b_rp[a_ci[j]];
  ub[i] = s;
  i64 g = 0;
  if (2 * s > SPG_TABLE) { g = 64; while (g < 2 * s) g <<= 1; }
  gsize[i] = g;
}

// symbolic pass.  FILL=false: count distinct columns of every row.  FILL=true: also emit them, ascending.
template <bool FILL>
__global__ void __launch_bounds__(SPG_WARPS * 32) k_spg_symbolic(i64 nA, const i64 *__restrict__ a_rp, const i32 *__restrict__ a_ci,
                                                                const i64 *__restrict__ b_rp, const i32 *__restrict__ b_ci,
                                                                const i64 *__restrict__ gsize, const i64 *__restrict__ goff,
                                                                i32 *gtable, i64 *cnt, const i64 *__restrict__ c_rp, i32 *c_ci)
{
  __shared__ i32 s_table[SPG_WARPS][SPG_TABLE];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const i64 i = (i64)blockIdx.x * SPG_WARPS + w;
  if (i >= nA) return;
  i32 *table;
  unsigned tsize;
  if (gsize[i] == 0) { table = s_table[w]; tsize = SPG_TABLE; }
  else { table = gtable + goff[i]; tsize = (unsigned)gsize[i]; }
  const unsigned mask = tsize - 1;
  for (unsigned q = lane; q < tsize; q += 32) table[q] = SPG_EMPTY;
  __syncwarp();
  int mycnt = 0;
  for (i64 j = a_rp[i]; j < a_rp[i + 1]; j++) {
    const i64 rb = a_ci[j];
    for (i64 k = b_rp[rb] + lane; k < b_rp[rb + 1]; k += 32) {
      const i32 c = b_ci[k];
      unsigned h = spg_hash(c, mask);
      for (;;) {
        const i32 old = atomicCAS(&table[h], SPG_EMPTY, c);
        if (old == SPG_EMPTY) { mycnt++; break; }
        if (old == c) break;
        h = (h + 1) & mask;
      }
    }
  }
  __syncwarp();
  for (int o = 16; o; o >>= 1) mycnt += __shfl_xor_sync(0xffffffffu, mycnt, o);
  if (!FILL) {
    if (lane == 0) cnt[i] = mycnt;
    return;
  }
  // in-place stream compaction of the occupied slots to the front of the table
  unsigned outp = 0;
  for (unsigned base = 0; base < tsize; base += 32) {
    const i32 v = table[base + lane];
    const unsigned bal = __ballot_sync(0xffffffffu, v != SPG_EMPTY);
    __syncwarp();
    if (v != SPG_EMPTY) table[outp + __popc(bal & ((1u << lane) - 1u))] = v;
    outp += __popc(bal);
    __syncwarp();
  }
  // pad to a power of two and bitonic-sort ascending
  unsigned m = 1;
  while (m < outp) m <<= 1;
  for (unsigned q = outp + lane; q < m; q += 32) table[q] = 0x7fffffff;
  __syncwarp();
  for (unsigned k = 2; k <= m; k <<= 1)
    for (unsigned j = k >> 1; j > 0; j >>= 1) {
      for (unsigned idx = lane; idx < m; idx += 32) {
        const unsigned ixj = idx ^ j;
        if (ixj > idx) {
          const i32 a = table[idx], b = table[ixj];
          const bool up = ((idx & k) == 0);
          if ((a > b) == up) { table[idx] = b; table[ixj] = a; }
        }
      }
      __syncwarp();
    }
  i32 *dst = c_ci + c_rp[i];
  for (unsigned q = lane; q < outp; q += 32) dst[q] = table[q];
}

template <int AH, int AW, int BW>
__global__ void __launch_bounds__(SPG_WARPS * 32) k_spg_numeric(i64 nA, const i64 *__restrict__ a_rp, const i32 *__restrict__ a_ci,
                                                               const double *__restrict__ a_v, const i64 *__restrict__ b_rp,
                                                               const i32 *__restrict__ b_ci, const double *__restrict__ b_v,
                                                               const i64 *__restrict__ c_rp, const i32 *__restrict__ c_ci, double *c_v)
{
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const i64 i = (i64)blockIdx.x * SPG_WARPS + w;
  if (i >= nA) return;
  const i64 cbeg = c_rp[i];
  const int ncol = (int)(c_rp[i + 1] - cbeg);
  const i32 *cc = c_ci + cbeg;
  for (i64 j = a_rp[i]; j < a_rp[i + 1]; j++) {
    const i64 rb = a_ci[j];
    double a[AH * AW];
#pragma unroll
    for (int e = 0; e < AH * AW; e++) a[e] = a_v[j * (AH * AW) + e];
    for (i64 k = b_rp[rb] + lane; k < b_rp[rb + 1]; k += 32) {
      const i32 c = b_ci[k];
      int lo = 0, hi = ncol - 1, pos = 0;
      while (lo <= hi) {
        const int mid = (lo + hi) >> 1;
        const i32 cm = cc[mid];
        if (cm == c) { pos = mid; break; }
        if (cm < c) lo = mid + 1; else hi = mid - 1;
      }
      double bb[AW * BW];
#pragma unroll
      for (int e = 0; e < AW * BW; e++) bb[e] = b_v[k * (AW * BW) + e];
      double *dst = c_v + (cbeg + pos) * (i64)(AH * BW);
#pragma unroll
      for (int p = 0; p < AH; p++)
#pragma unroll
        for (int q = 0; q < BW; q++) {
          double s = 0.0;
#pragma unroll
          for (int l = 0; l < AW; l++) s = __dadd_rn(s, __dmul_rn(a[p * AW + l], bb[l * BW + q]));
          __stcg(dst + p * BW + q, __dadd_rn(__ldcg(dst + p * BW + q), s));
        }
    }
    __syncwarp();
  }
}

template <int AH, int AW, int BW>
void launch_numeric(const DevCsr &A, const DevCsr &B, DevCsr &C, cudaStream_t st)
{
  const i64 nb = (A.nrows + SPG_WARPS - 1) / SPG_WARPS;
  k_spg_numeric<AH, AW, BW><<<(unsigned)nb, SPG_WARPS * 32, 0, st>>>(A.nrows, A.rowptr, A.col, A.val, B.rowptr, B.col, B.val,
                                                                    C.rowptr, C.col, C.val);
}

}  // namespace

void dev_csr_upload_raw(i64 nrows, i64 ncols, int bh, int bw, const i64 *rowptr, const i32 *col, const double *val,
                        DevCsr &d, cudaStream_t st)
{
  d.nrows = nrows; d.ncols = ncols; d.bh = bh; d.bw = bw; d.nnz = rowptr[nrows];
  d.rowptr = dev_alloc<i64>(nrows + 1);
  d.col = dev_alloc<i32>(d.nnz);
  d.val = dev_alloc<double>(d.nnz * bh * bw);
  NGB_CUDA(cudaMemcpyAsync(d.rowptr, rowptr, sizeof(i64) * (nrows + 1), cudaMemcpyHostToDevice, st));
  if (d.nnz) {
    NGB_CUDA(cudaMemcpyAsync(d.col, col, sizeof(i32) * d.nnz, cudaMemcpyHostToDevice, st));
    NGB_CUDA(cudaMemcpyAsync(d.val, val, sizeof(double) * d.nnz * bh * bw, cudaMemcpyHostToDevice, st));
  }
  NGB_CUDA(cudaStreamSynchronize(st));
}

void dev_csr_upload(const HostBsr &h, DevCsr &d, cudaStream_t st)
{
  dev_csr_upload_raw(h.nrows, h.ncols, h.bh, h.bw, h.rowptr.data(), h.col.data(), h.val.data(), d, st);
}

void dev_csr_download(const DevCsr &d, HostBsr &h, cudaStream_t st, bool with_values)
{
  h.nrows = d.nrows; h.ncols = d.ncols; h.bh = d.bh; h.bw = d.bw;
  h.rowptr.resize(d.nrows + 1);
  h.col.resize(d.nnz);
  NGB_CUDA(cudaMemcpyAsync(h.rowptr.data(), d.rowptr, sizeof(i64) * (d.nrows + 1), cudaMemcpyDeviceToHost, st));
  if (d.nnz) NGB_CUDA(cudaMemcpyAsync(h.col.data(), d.col, sizeof(i32) * d.nnz, cudaMemcpyDeviceToHost, st));
  if (with_values) {
    h.val.resize(d.nnz * d.bs());
    if (d.nnz) NGB_CUDA(cudaMemcpyAsync(h.val.data(), d.val, sizeof(double) * d.nnz * d.bs(), cudaMemcpyDeviceToHost, st));
  }
  NGB_CUDA(cudaStreamSynchronize(st));
}

void dev_csr_free(DevCsr &d)
{
  dev_free(d.rowptr);
  dev_free(d.col);
  dev_free(d.val);
  d.nnz = 0;
}

void dev_spgemm(const DevCsr &A, const DevCsr &B, DevCsr &C, cudaStream_t st, i64 *launches)
{
  if (A.ncols != B.nrows || A.bw != B.bh) throw Error("dev_spgemm: shape mismatch");
  const i64 nA = A.nrows;
  C.nrows = nA; C.ncols = B.ncols; C.bh = A.bh; C.bw = B.bw;
  C.rowptr = dev_alloc<i64>(nA + 1);
  i64 *ub = dev_alloc<i64>(nA + 1), *gsize = dev_alloc<i64>(nA + 1), *goff = dev_alloc<i64>(nA + 1), *cnt = dev_alloc<i64>(nA + 1);
  NGB_CUDA(cudaMemsetAsync(gsize, 0, sizeof(i64) * (nA + 1), st));
  NGB_CUDA(cudaMemsetAsync(cnt, 0, sizeof(i64) * (nA + 1), st));
  const int TB = 256;
  if (nA) k_spg_bound<<<(unsigned)((nA + TB - 1) / TB), TB, 0, st>>>(nA, A.rowptr, A.col, B.rowptr, ub, gsize);
  // offsets of the global hash tables of the large rows
  size_t tmp_bytes = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, gsize, goff, nA + 1, st);
  void *tmp = dev_alloc<char>(tmp_bytes);
  cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, gsize, goff, nA + 1, st);
  i64 gtotal = 0;
  NGB_CUDA(cudaMemcpyAsync(&gtotal, goff + nA, sizeof(i64), cudaMemcpyDeviceToHost, st));
  NGB_CUDA(cudaStreamSynchronize(st));
  i32 *gtable = dev_alloc<i32>(gtotal);
  const i64 nb = (nA + SPG_WARPS - 1) / SPG_WARPS;
  if (nA)
    k_spg_symbolic<false><<<(unsigned)nb, SPG_WARPS * 32, 0, st>>>(nA, A.rowptr, A.col, B.rowptr, B.col, gsize, goff, gtable, cnt,
                                                                  nullptr, nullptr);
  cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, cnt, C.rowptr, nA + 1, st);
  NGB_CUDA(cudaMemcpyAsync(&C.nnz, C.rowptr + nA, sizeof(i64), cudaMemcpyDeviceToHost, st));
  NGB_CUDA(cudaStreamSynchronize(st));
  C.col = dev_alloc<i32>(C.nnz);
  C.val = dev_alloc<double>(C.nnz * C.bs());
  NGB_CUDA(cudaMemsetAsync(C.val, 0, sizeof(double) * std::max<i64>(C.nnz * C.bs(), 1), st));
  if (nA) {
    k_spg_symbolic<true><<<(unsigned)nb, SPG_WARPS * 32, 0, st>>>(nA, A.rowptr, A.col, B.rowptr, B.col, gsize, goff, gtable, cnt,
                                                                 C.rowptr, C.col);
    const int key = A.bh * 100 + A.bw * 10 + B.bw;
    switch (key) {
      case 111: launch_numeric<1, 1, 1>(A, B, C, st); break;
      case 222: launch_numeric<2, 2, 2>(A, B, C, st); break;
      case 333: launch_numeric<3, 3, 3>(A, B, C, st); break;
      case 666: launch_numeric<6, 6, 6>(A, B, C, st); break;
      // elasticity level 0 (3 -> 6): PT(6x3)*A(3x3), (PTA)(6x3)*P(3x6); prolongation concatenation P(3x6)*P(6x6)
      case 633: launch_numeric<6, 3, 3>(A, B, C, st); break;
      case 636: launch_numeric<6, 3, 6>(A, B, C, st); break;
      case 366: launch_numeric<3, 6, 6>(A, B, C, st); break;
      case 336: launch_numeric<3, 3, 6>(A, B, C, st); break;
      // 2d elasticity (2 -> 3)
      case 322: launch_numeric<3, 2, 2>(A, B, C, st); break;
      case 323: launch_numeric<3, 2, 3>(A, B, C, st); break;
      case 233: launch_numeric<2, 3, 3>(A, B, C, st); break;
      case 223: launch_numeric<2, 2, 3>(A, B, C, st); break;
      default: throw Error("dev_spgemm: block shape " + std::to_string(key) + " not instantiated");
    }
  }
  NGB_CUDA(cudaStreamSynchronize(st));
  NGB_CUDA(cudaGetLastError());
  if (launches) *launches += 4 + 3;
  dev_free(ub); dev_free(gsize); dev_free(goff); dev_free(cnt); dev_free(gtable);
  cudaFree(tmp);
}

// ------------------------------------------------------------------------------------------------
// T = A^T on the device (TransposeSPMImpl, src/base/linalg/utils_sparseMM.cpp:54-93: counting sort by column, every block transposed,
// rows of T ascending).  A stable radix sort of the entries by column number keeps the row-major order of A inside every column,
// i.e. ascending row numbers -- the same integer arrays as the reference's counting sort + per-row BubbleSort, bit for bit.
// ------------------------------------------------------------------------------------------------
namespace {
__global__ void k_tr_rows(i64 nrows, const i64 *__restrict__ rp, i32 *row_of, i32 *cnt, const i32 *__restrict__ ci)
{
  const i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nrows) return;
  for (i64 k = rp[i]; k < rp[i + 1]; k++) { row_of[k] = (i32)i; atomicAdd(cnt + ci[k], 1); }
}
__global__ void k_tr_iota(i64 n, i32 *v)
{
  const i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) v[i] = (i32)i;
}
__global__ void k_tr_widen(i64 n, const i32 *__restrict__ in, i64 *out)
{
  const i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = in[i];
}
__global__ void k_tr_fill(i64 nnz, int bh, int bw, const i32 *__restrict__ src, const i32 *__restrict__ row_of, const double *__restrict__ aval,
                          i32 *tcol, double *tval)
{
  const i64 p = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= nnz) return;
  const i64 k = src[p];
  tcol[p] = row_of[k];
  const int bs = bh * bw;
  for (int r = 0; r < bh; r++)
    for (int c = 0; c < bw; c++) tval[p * bs + c * bh + r] = aval[k * bs + r * bw + c];
}
}  // namespace

void dev_transpose(const DevCsr &A, DevCsr &T, cudaStream_t st, i64 *launches)
{
  if (A.nnz >= (i64)2147483647) throw Error("dev_transpose: more than 2^31 entries");
  T.nrows = A.ncols; T.ncols = A.nrows; T.bh = A.bw; T.bw = A.bh; T.nnz = A.nnz;
  T.rowptr = dev_alloc<i64>(T.nrows + 1);
  T.col = dev_alloc<i32>(T.nnz);
  T.val = dev_alloc<double>(T.nnz * T.bs());
  const int TB = 256;
  auto nb = [&](i64 n) { return (unsigned)((n + TB - 1) / TB); };
  i32 *row_of = dev_alloc<i32>(A.nnz), *cnt = dev_alloc<i32>(T.nrows + 1), *idx = dev_alloc<i32>(A.nnz), *idx2 = dev_alloc<i32>(A.nnz),
      *keys2 = dev_alloc<i32>(A.nnz);
  i64 *cnt64 = dev_alloc<i64>(T.nrows + 1);
  NGB_CUDA(cudaMemsetAsync(cnt, 0, sizeof(i32) * (T.nrows + 1), st));
  if (A.nrows) k_tr_rows<<<nb(A.nrows), TB, 0, st>>>(A.nrows, A.rowptr, row_of, cnt, A.col);
  if (A.nnz) k_tr_iota<<<nb(A.nnz), TB, 0, st>>>(A.nnz, idx);
  k_tr_widen<<<nb(T.nrows + 1), TB, 0, st>>>(T.nrows + 1, cnt, cnt64);
  size_t b1 = 0, b2 = 0;
  int end_bit = 1;
  while (end_bit < 31 && ((i64)1 << end_bit) < std::max<i64>(A.ncols, 2)) end_bit++;
  cub::DeviceScan::ExclusiveSum(nullptr, b1, cnt64, T.rowptr, T.nrows + 1, st);
  cub::DeviceRadixSort::SortPairs(nullptr, b2, A.col, keys2, idx, idx2, A.nnz, 0, end_bit, st);
  void *tmp = dev_alloc<char>(std::max(b1, b2));
  cub::DeviceScan::ExclusiveSum(tmp, b1, cnt64, T.rowptr, T.nrows + 1, st);
  if (A.nnz) {
    cub::DeviceRadixSort::SortPairs(tmp, b2, A.col, keys2, idx, idx2, A.nnz, 0, end_bit, st);
    k_tr_fill<<<nb(A.nnz), TB, 0, st>>>(A.nnz, A.bh, A.bw, idx2, row_of, A.val, T.col, T.val);
  }
  NGB_CUDA(cudaStreamSynchronize(st));
  NGB_CUDA(cudaGetLastError());
  if (launches) *launches += 6;
  dev_free(row_of); dev_free(cnt); dev_free(idx); dev_free(idx2); dev_free(keys2); dev_free(cnt64);
  cudaFree(tmp);
}

}  // namespace ngb
