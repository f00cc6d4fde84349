// ref_bgs_harness.cpp -- compiles the reference's OWN block Gauss-Seidel update routines (cut out of /root/reference at build time into
// oracle/_ref/frag_bgs/, never stored in this repository) against ngs_standin_bgs.hpp and drives them over a block CSR matrix:
//   BSmoother2<TM>::BSBlock::RichardsonUpdate      loc_block_gssmoother_impl.hpp:244-268
//   BSmoother2<TM>::BSBlock::RichardsonUpdate_RES  loc_block_gssmoother_impl.hpp:516-541
// The per-block data (dofnrs, off-block rows firsti/cols/vals, diag, diag_inv) is laid out here the way BSBlock::SetFromSPMat (:67-132)
// stores it (LU = false, md = false); the dense inverse is handed in by the caller, so both sides of the comparison use the same one.
// Block order: IterateBlocks (:618-651) with one group -- ascending, or descending when `reverse`.  TEST INFRASTRUCTURE ONLY.
#include <cstdint>
#include <algorithm>
#include <cstring>
#include <vector>

#include "ngs_standin_bgs.hpp"

namespace amg {
template <class TM> class BSmoother2 {
public:
  using TV = typename bgs_vec_of<TM>::type;
  class BSBlock {
  public:
    bool LU = false, md = false;
    FlatArray<int> dofnrs;
    FlatArray<int> firsti;
    FlatArray<int> cols;
    FlatMatrix<TM> diag, diag_inv;
    FlatArray<TM> vals, mdadd;
    INLINE void RichardsonUpdate(double omega, FlatVector<TV> smallsol, FlatVector<TV> bigsol, FlatVector<TV> smallrhs, FlatVector<TV> bigrhs) const;
    INLINE void RichardsonUpdate_RES(double omega, FlatVector<TV> smallupdate, FlatVector<TV> bigsol, FlatVector<TV> smallres, FlatVector<TV> bigres) const;
    INLINE void Prefetch() const {}     // cache hint only (loc_block_gssmoother_impl.hpp:222-240)
    INLINE void SetFromSPMat(const SparseMatrixTM<TM> &A, FlatArray<int> dofs, LocalHeap &lh, bool pinv, FlatArray<TM> md);
  };
  // the members the cut-out methods below use (loc_block_gssmoother.hpp:78-99)
  Array<BSBlock> blocks;
  Array<size_t> fi_blocks;
  size_t maxbs = 0;
  BgsCsr A;
  const BgsCsr *GetAMatrix() const { return &A; }
  template <class TLAM> INLINE void IterateBlocks(FlatArray<int> groups, bool reverse, TLAM smooth_block) const;
  INLINE void SmoothWO(FlatArray<int> groups, BaseVector &x, const BaseVector &b, BaseVector &res, int steps, bool res_updated, bool update_res,
                       bool x_zero, bool reverse, bool symm) const;
  INLINE void SmoothSimple(FlatArray<int> groups, BaseVector &x, const BaseVector &b, int steps, bool reverse, bool symm) const;
  INLINE void SmoothRESSimple(FlatArray<int> groups, BaseVector &x, BaseVector &res, int steps, bool reverse, bool symm) const;
};
#define LAMBDA_INLINE
#include "../_ref/frag_bgs/bgs_richardson.inc"
#include "../_ref/frag_bgs/bgs_richardson_res.inc"
#include "../_ref/frag_bgs/bgs_setfromspmat.inc"
#include "../_ref/frag_bgs/bgs_iterate.inc"
#include "../_ref/frag_bgs/bgs_smoothsimple.inc"
#include "../_ref/frag_bgs/bgs_smoothressimple.inc"
#include "../_ref/frag_bgs/bgs_smoothwo.inc"
}  // namespace amg

namespace {
template <class TM> void set_block(TM &m, const double *v, int b);
template <> void set_block<double>(double &m, const double *v, int) { m = v[0]; }
template <int N> void set_block_mat(Mat<N, N> &m, const double *v) { for (int i = 0; i < N * N; i++) m.v[i] = v[i]; }
template <> void set_block<Mat<2, 2>>(Mat<2, 2> &m, const double *v, int) { set_block_mat<2>(m, v); }
template <> void set_block<Mat<3, 3>>(Mat<3, 3> &m, const double *v, int) { set_block_mat<3>(m, v); }
template <> void set_block<Mat<6, 6>>(Mat<6, 6> &m, const double *v, int) { set_block_mat<6>(m, v); }

// one sweep over all blocks.  mode 0: RichardsonUpdate (x, b); mode 1: RichardsonUpdate_RES (x, res)
template <class TM>
int sweep(int64_t n, int b, const int64_t *rp, const int32_t *ci, const double *av, int64_t nblocks, const int64_t *bptr, const int32_t *bverts,
          const double *dinv, const int64_t *dinv_off, double *x, double *r, int mode, int reverse)
{
  using S = amg::BSmoother2<TM>;
  using TV = typename S::TV;
  std::vector<int32_t> blk_of(n, -1), pos_in(n, -1);
  for (int64_t k = 0; k < nblocks; k++)
    for (int64_t q = bptr[k]; q < bptr[k + 1]; q++) { blk_of[bverts[q]] = (int32_t)k; pos_in[bverts[q]] = (int32_t)(q - bptr[k]); }
  std::vector<TM> A(rp[n]);
  for (int64_t e = 0; e < rp[n]; e++) set_block<TM>(A[e], av + e * b * b, b);
  FlatVector<TV> bigx(n, reinterpret_cast<TV *>(x)), bigr(n, reinterpret_cast<TV *>(r));
  for (int64_t kk = 0; kk < nblocks; kk++) {
    const int64_t k = reverse ? nblocks - 1 - kk : kk;
    const int m = (int)(bptr[k + 1] - bptr[k]);
    if (m == 0) continue;
    std::vector<int> dofnrs(m), firsti(m + 1, 0), cols;
    std::vector<TM> vals, diag((size_t)m * m), dinvm((size_t)m * m);
    for (int q = 0; q < m; q++) {
      const int v = bverts[bptr[k] + q];
      dofnrs[q] = v;
      for (int64_t e = rp[v]; e < rp[v + 1]; e++) {
        if (blk_of[ci[e]] == (int32_t)k) diag[(size_t)q * m + pos_in[ci[e]]] = A[e];
        else { cols.push_back(ci[e]); vals.push_back(A[e]); }
      }
      firsti[q + 1] = (int)cols.size();
    }
    const double *di = dinv + dinv_off[k];   // (m*b) x (m*b) row-major scalar matrix -> m x m blocks
    for (int qi = 0; qi < m; qi++)
      for (int qj = 0; qj < m; qj++) {
        double blk[36];
        for (int p = 0; p < b; p++) for (int q2 = 0; q2 < b; q2++) blk[p * b + q2] = di[(size_t)(qi * b + p) * (m * b) + qj * b + q2];
        set_block<TM>(dinvm[(size_t)qi * m + qj], blk, b);
      }
    typename S::BSBlock B;
    B.dofnrs.Assign(FlatArray<int>(m, dofnrs.data()));
    B.firsti.Assign(FlatArray<int>(m + 1, firsti.data()));
    B.cols.Assign(FlatArray<int>(cols.size(), cols.data()));
    B.vals.Assign(FlatArray<TM>(vals.size(), vals.data()));
    B.diag.AssignMemory(m, m, diag.data());
    B.diag_inv.AssignMemory(m, m, dinvm.data());
    std::vector<TV> h1(m), h2(m);
    FlatVector<TV> hx(m, h1.data()), hb(m, h2.data());
    if (mode == 0) B.RichardsonUpdate(1.0, hx, bigx, hb, bigr);
    else B.RichardsonUpdate_RES(1.0, hx, bigx, hb, bigr);
  }
  return 0;
}
}  // namespace

namespace {
// BSmoother2::SmoothWO (loc_block_gssmoother_impl.hpp:655-668) with the reference's own IterateBlocks / SmoothSimple / SmoothRESSimple:
// all blocks in ONE group (the serial smoother), `steps` sweeps, optional symmetric sweeps
template <class TM>
int smooth_wo(int64_t n, int b, const int64_t *rp, const int32_t *ci, const double *av, int64_t nblocks, const int64_t *bptr, const int32_t *bverts,
              const double *dinv, const int64_t *dinv_off, double *x, const double *rhs, double *res, int steps, int res_updated, int update_res, int x_zero,
              int reverse, int symm)
{
  using S = amg::BSmoother2<TM>;
  S sm;
  sm.A = BgsCsr{n, b, rp, ci, av};
  std::vector<int32_t> blk_of(n, -1), pos_in(n, -1);
  for (int64_t k = 0; k < nblocks; k++)
    for (int64_t q = bptr[k]; q < bptr[k + 1]; q++) { blk_of[bverts[q]] = (int32_t)k; pos_in[bverts[q]] = (int32_t)(q - bptr[k]); }
  std::vector<TM> A(rp[n]);
  for (int64_t e = 0; e < rp[n]; e++) set_block<TM>(A[e], av + e * b * b, b);
  SparseMatrixTM<TM> spm;
  spm.n = n; spm.rp = rp; spm.ci = ci;
  spm.cols.assign(ci, ci + rp[n]);
  spm.vals = A;
  struct Store { std::vector<int> dofnrs, firsti, cols; std::vector<TM> vals, diag, dinvm; };
  std::vector<Store> st(nblocks);
  sm.blocks.d.resize(nblocks);
  for (int64_t k = 0; k < nblocks; k++) {
    const int m = (int)(bptr[k + 1] - bptr[k]);
    Store &s = st[k];
    s.dofnrs.resize(m); s.firsti.assign(m + 1, 0); s.diag.resize((size_t)m * m); s.dinvm.resize((size_t)m * m);
    for (int q = 0; q < m; q++) {
      const int v = bverts[bptr[k] + q];
      s.dofnrs[q] = v;
      for (int64_t e = rp[v]; e < rp[v + 1]; e++) {
        if (blk_of[ci[e]] == (int32_t)k) s.diag[(size_t)q * m + pos_in[ci[e]]] = A[e];
        else { s.cols.push_back(ci[e]); s.vals.push_back(A[e]); }
      }
      s.firsti[q + 1] = (int)s.cols.size();
    }
    auto &B = sm.blocks[k];
    if (dinv == nullptr) {
      // the reference's OWN block set-up: BSBlock::SetFromSPMat (loc_block_gssmoother_impl.hpp:67-132) fills dofnrs (sorted), the off-block
      // rows, diag and diag_inv = CalcInverse(diag) (the stand-in's Gauss-Jordan; NGSolve's routine is third-party).  The vertices are
      // handed over in DESCENDING order so that its QuickSort has something to do.
      const size_t noff = s.cols.size();
      std::vector<int> dofs_in(s.dofnrs.rbegin(), s.dofnrs.rend());
      std::fill(s.dofnrs.begin(), s.dofnrs.end(), -7);
      std::fill(s.firsti.begin(), s.firsti.end(), -7);
      s.cols.assign(noff, -7);
      s.vals.assign(noff, TM(0.0));
      std::fill(s.diag.begin(), s.diag.end(), TM(-7.0));
      B.dofnrs.Assign(FlatArray<int>(m, s.dofnrs.data()));
      B.firsti.Assign(FlatArray<int>(m + 1, s.firsti.data()));
      B.cols.Assign(FlatArray<int>(noff, s.cols.data()));
      B.vals.Assign(FlatArray<TM>(noff, s.vals.data()));
      B.diag.AssignMemory(m, m, s.diag.data());
      B.diag_inv.AssignMemory(m, m, s.dinvm.data());
      LocalHeap lh;
      B.SetFromSPMat(spm, FlatArray<int>(m, dofs_in.data()), lh, false, FlatArray<TM>(0, nullptr));
      sm.maxbs = std::max<size_t>(sm.maxbs, (size_t)m);
      continue;
    }
    const double *di = dinv + dinv_off[k];
    for (int qi = 0; qi < m; qi++)
      for (int qj = 0; qj < m; qj++) {
        double blk[36];
        for (int p = 0; p < b; p++) for (int q2 = 0; q2 < b; q2++) blk[p * b + q2] = di[(size_t)(qi * b + p) * (m * b) + qj * b + q2];
        set_block<TM>(s.dinvm[(size_t)qi * m + qj], blk, b);
      }
    B.dofnrs.Assign(FlatArray<int>(m, s.dofnrs.data()));
    B.firsti.Assign(FlatArray<int>(m + 1, s.firsti.data()));
    B.cols.Assign(FlatArray<int>(s.cols.size(), s.cols.data()));
    B.vals.Assign(FlatArray<TM>(s.vals.size(), s.vals.data()));
    B.diag.AssignMemory(m, m, s.diag.data());
    B.diag_inv.AssignMemory(m, m, s.dinvm.data());
    sm.maxbs = std::max<size_t>(sm.maxbs, (size_t)m);
  }
  sm.fi_blocks.d = {0, (size_t)nblocks};
  int group0 = 0;
  BaseVector vx{x, (size_t)n * b}, vb{const_cast<double *>(rhs), (size_t)n * b}, vr{res, (size_t)n * b};
  sm.SmoothWO(FlatArray<int>(1, &group0), vx, vb, vr, steps, res_updated != 0, update_res != 0, x_zero != 0, reverse != 0, symm != 0);
  return 0;
}
}  // namespace

extern "C" {
int ref_bgs_smooth_wo(int64_t n, int b, const int64_t *rp, const int32_t *ci, const double *av, int64_t nblocks, const int64_t *bptr, const int32_t *bverts,
                      const double *dinv, const int64_t *dinv_off, double *x, const double *rhs, double *res, int steps, int res_updated, int update_res,
                      int x_zero, int reverse, int symm)
{
  try {
    switch (b) {
      case 1: return smooth_wo<double>(n, b, rp, ci, av, nblocks, bptr, bverts, dinv, dinv_off, x, rhs, res, steps, res_updated, update_res, x_zero, reverse, symm);
      case 2: return smooth_wo<Mat<2, 2>>(n, b, rp, ci, av, nblocks, bptr, bverts, dinv, dinv_off, x, rhs, res, steps, res_updated, update_res, x_zero, reverse, symm);
      case 3: return smooth_wo<Mat<3, 3>>(n, b, rp, ci, av, nblocks, bptr, bverts, dinv, dinv_off, x, rhs, res, steps, res_updated, update_res, x_zero, reverse, symm);
      case 6: return smooth_wo<Mat<6, 6>>(n, b, rp, ci, av, nblocks, bptr, bverts, dinv, dinv_off, x, rhs, res, steps, res_updated, update_res, x_zero, reverse, symm);
      default: return 2;
    }
  } catch (...) { return 1; }
}
// A: block CSR (n block rows, b x b blocks, row-major); blocks: bptr / bverts (vertices ascending per block); dinv: per block the dense
// (m b) x (m b) inverse, row-major, at dinv_off[k].  x and r (= rhs for mode 0, residual for mode 1) are updated in place.
int ref_bgs_sweep(int64_t n, int b, const int64_t *rp, const int32_t *ci, const double *av, int64_t nblocks, const int64_t *bptr, const int32_t *bverts,
                  const double *dinv, const int64_t *dinv_off, double *x, double *r, int mode, int reverse)
{
  try {
    switch (b) {
      case 1: return sweep<double>(n, b, rp, ci, av, nblocks, bptr, bverts, dinv, dinv_off, x, r, mode, reverse);
      case 2: return sweep<Mat<2, 2>>(n, b, rp, ci, av, nblocks, bptr, bverts, dinv, dinv_off, x, r, mode, reverse);
      case 3: return sweep<Mat<3, 3>>(n, b, rp, ci, av, nblocks, bptr, bverts, dinv, dinv_off, x, r, mode, reverse);
      case 6: return sweep<Mat<6, 6>>(n, b, rp, ci, av, nblocks, bptr, bverts, dinv, dinv_off, x, r, mode, reverse);
      default: return 2;
    }
  } catch (...) { return 1; }
}
}
