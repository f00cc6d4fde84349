// kernels_itile.cuh -- the CTA-per-tile Gauss-Seidel sweep on PREPARED TILE IMAGES (the production path for matrices with at most 7
// entries per row and triangle, e.g. P1 tets; everything else runs k_gs_ctile, kernels_ctile.cuh).
//
// What k_gs_ctile does per sweep -- scanning column indices, turning them into tile-local gather indices, separating the couplings to
// other tiles, loading dinv / diag / level numbers row by row -- does not depend on the vectors.  It is done ONCE at setup (k_img_count /
// k_img_fill below): every tile gets a contiguous image in global memory, already in the layout the kernel wants in shared memory
//
//   header (16 B) | level starts (u16) | records: per row { 8 x u16 gather index | 7 values | dinv } = 80 B | [diag per row, forward only]
//                 | external rows: { row, first entry, count } | external entries: { column, byte offset of the value }
//                 | overflow entries { gather index, value } of rows with 8 .. 22 entries (hybrid stage order: rows next to an interface)
//
// so that a sweep fetches a tile with ONE bulk copy (+ one for the tile's rows of the right-hand side, + one for `self`), all on one
// mbarrier, and no thread touches the matrix through the LSU.  Rows are sorted by tile-local level; a record is row-major (5 x LDS.128 per
// row for the solver; 80 B stride = conflict-free quarter-warps); gather index 8*MAXS*... = ZERO addresses a slot that always holds 0.0
// (padding slots and couplings to other tiles, which the gather phase folds into the accumulator).
// HBM bytes per row: 80 + 8 (diag) + ~18 (external lists) = ~106 for the forward sweep against 100 of the plain SELL triangle.
#pragma once
#include "kernels_ctile.cuh"

namespace ngb {

constexpr int IT_NV = 7;            // value slots per record
constexpr int IT_REC = 80;          // bytes per record
constexpr int IT_MAXOVF = 15;       // overflow entries per row (slot 7 of the index word: first overflow entry << 4 | count, 0xffff = none)

struct __align__(16) ITileMeta {
  i64 img;       // byte offset of the tile's image
  i32 r0;        // first row of the tile
  i32 nrow;      // rows incl. padding (multiple of 32)
  i32 bytes;     // image size (multiple of 16)
  i32 pad;
  i32 d0;        // first entry of the tile's wait list in `dep`
  i32 nd;        // tiles to wait for
};

struct ITileHeader {   // 16 bytes
  unsigned short nlev, nreal, nrow, n_ext_rows;
  unsigned short n_ext, ls_bytes, has_diag, n_ovf;
};

__host__ __device__ inline i64 itile_image_bytes(int nlev, int nrow, int n_ext_rows, int n_ext, int n_ovf, bool with_diag)
{
  const i64 ls = ((2 * (i64)(nlev + 1) + 15) / 16) * 16;
  i64 b = 16 + ls + (i64)IT_REC * nrow + (with_diag ? 8 * (i64)nrow : 0) + 8 * (i64)n_ext_rows + 8 * (i64)n_ext;
  b = ((b + 15) / 16) * 16 + 16 * (i64)n_ovf;
  return ((b + 127) / 128) * 128;
}

// ---- setup: pass 1 -- external rows / entries per tile ------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_img_count(i32 ntiles, const i32 *__restrict__ tile_slice, SellView T, i32 *n_ext_rows, i32 *n_ext, i32 *n_ovf, i32 *maxlen)
{
  __shared__ int s_rows, s_ent, s_max, s_ovf;
  const i32 t = blockIdx.x;
  if (t >= ntiles) return;
  if (threadIdx.x == 0) { s_rows = 0; s_ent = 0; s_max = 0; s_ovf = 0; }
  __syncthreads();
  const i32 s0 = tile_slice[t], s1 = tile_slice[t + 1];
  const i32 r0 = s0 * 32;
  const unsigned nrow = (unsigned)(s1 - s0) * 32u;
  int rows = 0, ent = 0, mx = 0, ov = 0;
  for (unsigned lr = threadIdx.x; lr < nrow; lr += blockDim.x) {
    const i64 slice = s0 + (lr >> 5);
    const int lane = lr & 31;
    const i64 base = T.slice_ptr[slice];
    const int wd = (int)(T.slice_ptr[slice + 1] - base);
    int e = 0, len = 0;
    for (int k = 0; k < wd; k++) {
      const i32 c = T.col[(base + k) * 32 + lane];
      if (c < 0) continue;
      len++;
      if ((unsigned)(c - r0) >= nrow) e++;
    }
    rows += e > 0; ent += e; mx = max(mx, len); ov += max(0, len - IT_NV);
  }
  atomicAdd(&s_rows, rows); atomicAdd(&s_ent, ent); atomicMax(&s_max, mx); atomicAdd(&s_ovf, ov);
  __syncthreads();
  if (threadIdx.x == 0) { n_ext_rows[t] = s_rows; n_ext[t] = s_ent; n_ovf[t] = s_ovf; atomicMax(maxlen, s_max); }
}

// ---- setup: pass 2 -- write the images.  One CTA per tile, thread per row; external lists in (row, slot) order (deterministic). ------
template <int MAXROWS>
__global__ void __launch_bounds__(128) k_img_fill(i32 ntiles, const i32 *__restrict__ tile_slice, const i32 *__restrict__ tile_nlev,
                                                 const i32 *__restrict__ tile_nreal, const uint8_t *__restrict__ row_lvl, SellView T,
                                                 const double *__restrict__ dinv, const double *__restrict__ diag, int with_diag,
                                                 const i64 *__restrict__ img_off, const i32 *__restrict__ n_ext_rows, const i32 *__restrict__ n_ext,
                                                 const i32 *__restrict__ n_ovf, unsigned char *img)
{
  constexpr int NT = 128, PER = MAXROWS / NT;
  __shared__ int s_cnt[MAXROWS + 1];       // external entries per row -> exclusive prefix
  __shared__ int s_rowid[MAXROWS + 1];     // external row flags -> exclusive prefix
  __shared__ int s_ovf[MAXROWS + 1];       // overflow entries per row -> exclusive prefix
  const i32 t = blockIdx.x;
  if (t >= ntiles) return;
  const i32 s0 = tile_slice[t], s1 = tile_slice[t + 1];
  const i32 r0 = s0 * 32;
  const unsigned nrow = (unsigned)(s1 - s0) * 32u;
  const int nlev = tile_nlev[t], nreal = tile_nreal[t];
  unsigned char *base_p = img + img_off[t];
  const i64 ls_bytes = ((2 * (i64)(nlev + 1) + 15) / 16) * 16;
  unsigned short *ls = (unsigned short *)(base_p + 16);
  unsigned char *recs = base_p + 16 + ls_bytes;
  double *dg = (double *)(recs + (i64)IT_REC * nrow);
  unsigned short *xrows = (unsigned short *)((unsigned char *)dg + (with_diag ? 8 * (i64)nrow : 0));
  i32 *xent = (i32 *)((unsigned char *)xrows + 8 * (i64)n_ext_rows[t]);
  // overflow entries start at the next 16-byte boundary (counted from the image base, which is 128-byte aligned)
  const i64 ovf_off = ((((unsigned char *)xent + 8 * (i64)n_ext[t]) - base_p + 15) / 16) * 16;
  unsigned char *ovf = base_p + ovf_off;
  const i64 ovf_from_rec = ovf - recs;
  if (threadIdx.x == 0) {
    ITileHeader h;
    h.nlev = (unsigned short)nlev; h.nreal = (unsigned short)nreal; h.nrow = (unsigned short)nrow; h.n_ext_rows = (unsigned short)n_ext_rows[t];
    h.n_ext = (unsigned short)n_ext[t]; h.ls_bytes = (unsigned short)ls_bytes; h.has_diag = (unsigned short)with_diag; h.n_ovf = (unsigned short)n_ovf[t];
    *(ITileHeader *)base_p = h;
    ls[nlev] = (unsigned short)nreal;
  }
  // per-row external / overflow counts, level starts
#pragma unroll
  for (int j = 0; j < PER; j++) {
    const unsigned lr = threadIdx.x + j * NT;
    int e = 0, len = 0;
    if (lr < nrow) {
      const i64 slice = s0 + (lr >> 5), row = (i64)r0 + lr;
      const int lane = lr & 31;
      const i64 b = T.slice_ptr[slice];
      const int wd = (int)(T.slice_ptr[slice + 1] - b);
      for (int k = 0; k < wd; k++) {
        const i32 c = T.col[(b + k) * 32 + lane];
        if (c < 0) continue;
        len++;
        if ((unsigned)(c - r0) >= nrow) e++;
      }
      const int lv = row_lvl[row], lvp = lr ? (int)row_lvl[row - 1] : -1;
      if (lv != lvp && lv < 255) ls[lv] = (unsigned short)lr;
    }
    if (lr < MAXROWS) { s_cnt[lr] = e; s_rowid[lr] = e > 0; s_ovf[lr] = max(0, len - IT_NV); }
  }
  __syncthreads();
  if (threadIdx.x == 0) {   // tiles are small: a serial prefix is fine at setup
    int a = 0, b2 = 0, o2 = 0;
    for (unsigned i = 0; i < nrow; i++) {
      const int c = s_cnt[i], f = s_rowid[i], o = s_ovf[i];
      s_cnt[i] = a; s_rowid[i] = b2; s_ovf[i] = o2;
      a += c; b2 += f; o2 += o;
    }
    s_cnt[nrow] = a; s_rowid[nrow] = b2; s_ovf[nrow] = o2;
  }
  __syncthreads();
  // records, overflow entries, external lists
#pragma unroll
  for (int j = 0; j < PER; j++) {
    const unsigned lr = threadIdx.x + j * NT;
    if (lr >= nrow) continue;
    const i64 slice = s0 + (lr >> 5), row = (i64)r0 + lr;
    const int lane = lr & 31;
    const i64 b = T.slice_ptr[slice];
    const int wd = (int)(T.slice_ptr[slice + 1] - b);
    const int xfirst = s_cnt[lr], xcnt = s_cnt[lr + 1] - xfirst;
    const int ofirst = s_ovf[lr], ocnt = s_ovf[lr + 1] - ofirst;
    unsigned short ix[8];
    double v[IT_NV];
#pragma unroll
    for (int k = 0; k < 8; k++) ix[k] = (unsigned short)MAXROWS;
#pragma unroll
    for (int k = 0; k < IT_NV; k++) v[k] = 0.0;
    ix[7] = ocnt ? (unsigned short)((ofirst << 4) | ocnt) : (unsigned short)0xffff;
    int slot = 0, xo = xfirst;
    for (int k = 0; k < wd; k++) {
      const i32 c = T.col[(b + k) * 32 + lane];
      if (c < 0) continue;
      const double val = T.val[(b + k) * 32 + lane];
      const unsigned lc = (unsigned)(c - r0);
      const bool in = lc < nrow;
      i64 voff;
      if (slot < IT_NV) {
        v[slot] = val;
        if (in) ix[slot] = (unsigned short)lc;
        voff = (i64)IT_REC * lr + 16 + 8 * slot;
      } else {
        unsigned char *oe = ovf + 16 * (i64)(ofirst + slot - IT_NV);
        *(unsigned *)oe = in ? lc : (unsigned)MAXROWS;
        *(unsigned *)(oe + 4) = 0u;
        *(double *)(oe + 8) = val;
        voff = ovf_from_rec + 16 * (i64)(ofirst + slot - IT_NV) + 8;
      }
      if (!in) { xent[2 * xo] = c; xent[2 * xo + 1] = (i32)voff; xo++; }
      slot++;
    }
    unsigned char *rec = recs + (i64)IT_REC * lr;
    *(uint4 *)rec = make_uint4(ix[0] | (ix[1] << 16), ix[2] | (ix[3] << 16), ix[4] | (ix[5] << 16), ix[6] | (ix[7] << 16));
    double *rv = (double *)(rec + 16);
#pragma unroll
    for (int k = 0; k < IT_NV; k++) rv[k] = v[k];
    rv[IT_NV] = dinv[row];
    if (with_diag) dg[lr] = diag[row];
    if (xcnt) {
      const int xr = s_rowid[lr];
      xrows[xr * 4 + 0] = (unsigned short)lr; xrows[xr * 4 + 1] = (unsigned short)xfirst; xrows[xr * 4 + 2] = (unsigned short)xcnt; xrows[xr * 4 + 3] = 0;
    }
  }
}

// bulk L2 prefetch (no shared-memory destination): the next tile's image is on its way to L2 while the current tile is swept
__device__ __forceinline__ void bulk_prefetch_l2(const void *src, uint32_t bytes)
{
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}

struct ITileParams {
  i32 ntiles;
  int backward;
  const ITileMeta *meta;
  const unsigned char *img;
  const i32 *dep;
  int *done;
  unsigned sleep_ns, repoll_ns;
  int cap_bytes;             // largest image (bytes)
  int *err;
  unsigned long long *trace;
};

// shared memory: [ header 128: mbarrier | xs (+ zero slot) | acc | aux | image ]
__host__ __device__ inline size_t itile_smem_bytes(int maxrows, int cap_bytes) { return 128 + ((size_t)maxrows + 16) * 8 + 2 * (size_t)maxrows * 8 + (size_t)cap_bytes; }

template <int NT, int MAXROWS, bool ADD_SELF, bool WRITE_R, int MINB>
__global__ void __launch_bounds__(NT, MINB) k_gs_itile(const double *rin, const double *__restrict__ self, double *out, double *rout, ITileParams p)
{
  constexpr int NW = NT / 32;
  constexpr unsigned ZERO = MAXROWS;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  int tid;
  asm volatile("mov.u32 %0, %%tid.x;" : "=r"(tid));
  const int lane = tid & 31, w = tid >> 5;
  uint32_t smem_base;
  asm volatile("mov.u32 %0, %1;" : "=r"(smem_base) : "r"(smem_u32(smem_raw)));
  const uint32_t bar = smem_base;
  const uint32_t xs_a = smem_base + 128, acc_a = xs_a + (MAXROWS + 16) * 8, aux_a = acc_a + MAXROWS * 8, img_a = aux_a + MAXROWS * 8;
  if (tid == 0) { mbar_init(bar, 1); mbar_fence_init(); sts_f64(xs_a + ZERO * 8u, 0.0); }
  __syncthreads();
  auto tile_of = [&](i32 q) { return p.backward ? (p.ntiles - 1 - q) : q; };
  auto fetch = [&](const ITileMeta &m) {   // one thread: the image, the tile's rows of rin (-> acc) and of self (-> aux)
    const uint32_t vb = (uint32_t)m.nrow * 8u;
    mbar_expect_tx(bar, (uint32_t)m.bytes + vb + (ADD_SELF ? vb : 0u));
    bulk_g2s(img_a, p.img + m.img, (uint32_t)m.bytes, bar);
    bulk_g2s(acc_a, rin + m.r0, vb, bar);
    if (ADD_SELF) bulk_g2s(aux_a, self + m.r0, vb, bar);
  };
  const ITileMeta none{0, 0, 0, 0, 0, 0, 0};
  uint32_t phase = 0;
  i32 q = blockIdx.x;
  if (q >= p.ntiles) return;
  ITileMeta cur = p.meta[tile_of(q)];
  i32 cur_dep = (tid < cur.nd) ? p.dep[cur.d0 + tid] : -1;
  if (tid == 0) fetch(cur);
  int solver = (int)(blockIdx.x % NW);
  for (; q < p.ntiles; q += gridDim.x) {
    const i32 t = tile_of(q);
    unsigned long long *tr = (p.trace && tid == 0) ? p.trace + (size_t)t * 16 : nullptr;
    if (tr) tr[0] = gtimer();
    const i32 qn = q + (i32)gridDim.x;
    const bool more = qn < p.ntiles;
    const ITileMeta nxt = more ? p.meta[tile_of(qn)] : none;
    const i32 r0 = cur.r0;
    if (more && tid == 0) {
      // the slab in shared memory is busy until this tile is done: stage the next tile's bytes in L2 meanwhile
      bulk_prefetch_l2(p.img + nxt.img, (uint32_t)nxt.bytes);
      bulk_prefetch_l2(rin + nxt.r0, (uint32_t)nxt.nrow * 8u);
      if (ADD_SELF) bulk_prefetch_l2(self + nxt.r0, (uint32_t)nxt.nrow * 8u);
    }
    // ---- the first hint flag is requested right away; the image is usually there already: parse it and preload this thread's external
    // row (columns, values) so that nothing but the data polls themselves follows the flags on the critical path
    int flag0 = 1;
    if (tid < cur.nd) flag0 = ld_relaxed_i32(p.done + cur_dep);
    mbar_wait(bar, phase);
    phase ^= 1u;
    if (tr) tr[1] = gtimer();
    const uint4 hw = lds_v4(img_a);
    const int nlev = hw.x & 0xffff, nreal = hw.x >> 16, nrow = hw.y & 0xffff, nxr = hw.y >> 16;
    const uint32_t ls_a = img_a + 16, rec_a = ls_a + (hw.z >> 16);
    const uint32_t dg_a = rec_a + (uint32_t)nrow * IT_REC;
    const uint32_t xr_a = dg_a + (WRITE_R ? (uint32_t)nrow * 8u : 0u);
    const uint32_t xe_a = xr_a + (uint32_t)nxr * 8u;
    const uint32_t ov_a = img_a + (((xe_a - img_a) + (hw.z & 0xffffu) * 8u + 15u) & ~15u);        // overflow entries
    uint32_t row0 = 0, first0 = 0, cnt0 = 0;
    i32 c0[IT_NV];
    double v0[IT_NV], a0 = 0.0;
#pragma unroll
    for (int i = 0; i < IT_NV; i++) { c0[i] = -1; v0[i] = 0.0; }
    if (tid < nxr) {
      const uint32_t e0 = lds_i32(xr_a + (uint32_t)tid * 8u), e1 = lds_i32(xr_a + (uint32_t)tid * 8u + 4u);
      row0 = e0 & 0xffffu; first0 = e0 >> 16; cnt0 = e1 & 0xffffu;
      a0 = lds_f64(acc_a + row0 * 8u);
#pragma unroll
      for (int i = 0; i < IT_NV; i++)
        if (i < (int)cnt0) { c0[i] = lds_i32(xe_a + (first0 + i) * 8u); v0[i] = lds_f64(rec_a + (uint32_t)lds_i32(xe_a + (first0 + i) * 8u + 4u)); }
    }
    // ---- hint flags of the tiles this one depends on
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
    __syncthreads();
    if (tr) tr[2] = gtimer();
    // ---- couplings to rows of other tiles: one thread per external row, all its polls in flight together; the data is the flag
    auto poll_row = [&](const i32 (&c)[IT_NV], const double (&v)[IT_NV], double a) {
      double x[IT_NV];
      unsigned spins = 0;
      bool missing;
#pragma unroll 1
      do {
        missing = false;
#pragma unroll
        for (int i = 0; i < IT_NV; i++) x[i] = (c[i] >= 0) ? ld_poll_relaxed(out + c[i]) : 0.0;
#pragma unroll
        for (int i = 0; i < IT_NV; i++) missing |= is_sentinel(x[i]);
        if (missing) {
          if (p.repoll_ns) __nanosleep(p.repoll_ns);
          if (spin_fail(spins, p.err)) break;
        }
      } while (missing);
#pragma unroll
      for (int i = 0; i < IT_NV; i++) a = fma(-v[i], x[i], a);      // v = 0 where there is no entry
      return a;
    };
    if (tid < nxr) {
      a0 = poll_row(c0, v0, a0);
      // rows with more than IT_NV couplings to other tiles (rare), and further external rows of this thread (tiles with > NT of them)
#pragma unroll 1
      for (uint32_t b0 = IT_NV; b0 < cnt0; b0 += IT_NV) {
        i32 c[IT_NV];
        double v[IT_NV];
#pragma unroll
        for (int i = 0; i < IT_NV; i++) {
          const bool in = b0 + i < cnt0;
          c[i] = in ? lds_i32(xe_a + (first0 + b0 + i) * 8u) : -1;
          v[i] = in ? lds_f64(rec_a + (uint32_t)lds_i32(xe_a + (first0 + b0 + i) * 8u + 4u)) : 0.0;
        }
        a0 = poll_row(c, v, a0);
      }
      sts_f64(acc_a + row0 * 8u, a0);
    }
#pragma unroll 1
    for (int er = tid + NT; er < nxr; er += NT) {
      const uint32_t e0 = lds_i32(xr_a + (uint32_t)er * 8u), e1 = lds_i32(xr_a + (uint32_t)er * 8u + 4u);
      const uint32_t row = e0 & 0xffffu, first = e0 >> 16, cnt = e1 & 0xffffu;
      double a = lds_f64(acc_a + row * 8u);
#pragma unroll 1
      for (uint32_t b0 = 0; b0 < cnt; b0 += IT_NV) {
        i32 c[IT_NV];
        double v[IT_NV];
#pragma unroll
        for (int i = 0; i < IT_NV; i++) {
          const bool in = b0 + i < cnt;
          c[i] = in ? lds_i32(xe_a + (first + b0 + i) * 8u) : -1;
          v[i] = in ? lds_f64(rec_a + (uint32_t)lds_i32(xe_a + (first + b0 + i) * 8u + 4u)) : 0.0;
        }
        a = poll_row(c, v, a);
      }
      sts_f64(acc_a + row * 8u, a);
    }
    if (tr) tr[3] = gtimer();
    const i32 nxt_dep = (more && tid < nxt.nd) ? p.dep[nxt.d0 + tid] : -1;
    __syncthreads();
    // ---- the tile itself: the solver warp (role rotates over the warps = schedulers) walks the local levels
    // (a software-pipelined variant -- next item's record loaded during the FMA chain -- was measured SLOWER: 1.60 vs 1.35 ms, registers)
    if (w == solver) {
#pragma unroll 1
      for (int it = 0; it < nlev; it++) {
        const int s = p.backward ? (nlev - 1 - it) : it;
        const int rb = (int)lds_u16(ls_a + (uint32_t)s * 2u), re = (int)lds_u16(ls_a + (uint32_t)(s + 1) * 2u);
#pragma unroll 1
        for (int rr = rb; rr < re; rr += 32) {
          const int r = rr + lane;
          const bool act = r < re;
          const uint32_t rc = (uint32_t)(act ? r : rb);
          const uint32_t ra = rec_a + rc * IT_REC;
          const uint4 iw = lds_v4(ra), q1 = lds_v4(ra + 16), q2 = lds_v4(ra + 32), q3 = lds_v4(ra + 48), q4 = lds_v4(ra + 64);
          double a = lds_f64(acc_a + rc * 8u), a2 = 0.0;
          const double ax = ADD_SELF ? lds_f64(aux_a + rc * 8u) : (WRITE_R ? lds_f64(dg_a + rc * 8u) : 0.0);
          const double v0 = __hiloint2double(q1.y, q1.x), v1 = __hiloint2double(q1.w, q1.z), v2 = __hiloint2double(q2.y, q2.x),
                       v3 = __hiloint2double(q2.w, q2.z), v4 = __hiloint2double(q3.y, q3.x), v5 = __hiloint2double(q3.w, q3.z),
                       v6 = __hiloint2double(q4.y, q4.x), dvv = __hiloint2double(q4.w, q4.z);
          const double x0 = lds_f64(xs_a + (iw.x & 0xffffu) * 8u), x1 = lds_f64(xs_a + (iw.x >> 16) * 8u), x2 = lds_f64(xs_a + (iw.y & 0xffffu) * 8u),
                       x3 = lds_f64(xs_a + (iw.y >> 16) * 8u), x4 = lds_f64(xs_a + (iw.z & 0xffffu) * 8u), x5 = lds_f64(xs_a + (iw.z >> 16) * 8u),
                       x6 = lds_f64(xs_a + (iw.w & 0xffffu) * 8u);
          a = fma(-v0, x0, a); a2 = fma(-v1, x1, a2); a = fma(-v2, x2, a); a2 = fma(-v3, x3, a2);
          a = fma(-v4, x4, a); a2 = fma(-v5, x5, a2); a = fma(-v6, x6, a);
          const unsigned o7 = iw.w >> 16;
          if (o7 != 0xffffu) {                         // rows with more than 7 entries (rare: next to an interface of a distributed level)
            const uint32_t oa = ov_a + (o7 >> 4) * 16u;
            for (unsigned k = 0; k < (o7 & 15u); k++) {
              const uint4 oe = lds_v4(oa + k * 16u);
              a = fma(-__hiloint2double(oe.w, oe.z), lds_f64(xs_a + oe.x * 8u), a);
            }
          }
          a += a2;
          const double d = dvv * a;
          if (act) {
            sts_f64(xs_a + (uint32_t)r * 8u, ADD_SELF ? ax + d : d);
            if (WRITE_R) sts_f64(acc_a + (uint32_t)r * 8u, fma(-ax, d, a));
          }
        }
        __syncwarp();
      }
    }
    __syncthreads();
    // ---- publish: coalesced stores; padding rows are never updated, but `out` must not keep the sentinel.  (Letting the solver store
    // every row to global memory as it goes was measured slower: +0.85 us of solver time per tile for nothing on the critical path.)
    for (int lr = tid; lr < nrow; lr += NT) {
      const i64 row = (i64)r0 + lr;
      const bool real = lr < nreal;
      __stcg(out + row, real ? lds_f64(xs_a + (uint32_t)lr * 8u) : (ADD_SELF ? lds_f64(aux_a + (uint32_t)lr * 8u) : 0.0));
      if (WRITE_R) rout[row] = lds_f64(acc_a + (uint32_t)lr * 8u);
    }
    if (tid == 0) st_relaxed_i32(p.done + t, 1);
    if (tr) tr[4] = gtimer();
    if (tr) {
      unsigned smid;
      asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
      tr[5] = gtimer();
      tr[6] = (unsigned long long)smid | ((unsigned long long)blockIdx.x << 16) | ((unsigned long long)nlev << 40) | ((unsigned long long)(nrow / 32) << 52);
    }
    __syncthreads();     // image, xs, acc, aux are free again
    if (more && tid == 0) fetch(nxt);
    cur = nxt;
    cur_dep = nxt_dep;
    solver = (solver + 1) % NW;
  }
}

}  // namespace ngb
