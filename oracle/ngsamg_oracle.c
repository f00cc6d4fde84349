/*
 * ngsamg_oracle.c -- CPU restatement of NgsAMG's preconditioner-apply hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (ngsamg_b200/, include/) may
 * include, link or call this file; only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs use it, and there only as the checker / CPU baseline.
 *
 * PARITY: PINNED IN PART.  The reference (LukasKogler/NgsAMG) is an NGSolve add-on; NGSolve, netgen,
 * MPI, METIS and LAPACK are absent from this image, so the reference as a whole cannot be built
 * or run here and its tests hold no golden vectors (only CG iteration ceilings).  What IS pinned:
 * the bodies of the reference's own functions for the single-rank path -- TransposeSPMImpl,
 * MatMultABImpl, RestrictMatrix, GSS3::SetUp/CalcDiags/SmoothRHSInternal/SmoothRESInternal/
 * Smooth/SmoothBack, CalcPseudoInverseTryNormal (with CallOnNonZeroDiagonalBlock, TryDirectInverse_simple),
 * RichardsonSmoother/JacobiSmoother, BaseSmoother::SmoothSymm/SmoothK/SmoothBackK/SmoothSymmK/CalcResiduum,
 * ProxySmoother::Smooth/SmoothBack, ProlMap::TransferF2C/AddC2F, AMGMatrix::SmoothV/SmoothW/
 * SmoothBS/SmoothVFromLevel -- are cut out of /root/reference at build time and compiled verbatim
 * against a stand-in for the NGSolve containers (oracle/ref_pin/ -> oracle/_ref/libngsamg_ref.so);
 * tests/test_ref_pin.py compares this file with that library bit for bit (patterns, values, sweeps,
 * level vectors) and against fixtures it wrote (tests/golden/refpin_*.npz).
 * STILL UNPINNED: the arithmetic that lives inside NGSolve and is therefore restated on both
 * sides (SparseMatrix::RowTimesVector / AddRowTransToVector / MultAdd summation order, Mat*Vec
 * evaluation order, CalcInverse, MergeArrays, SparseCholesky, krylovspace.CGSolver; NGSolve is only
 * lower-bounded, `ngsolve>=6.2.2403.post68.dev0`, pyproject.toml:2), the LAPACK eigen-solver
 * inside the pseudo-inverse fall-back, DiagonalMatrix::MultAdd (Jacobi update).  The multi-rank oracle (oracle_par.py) is pinned
 * the same way for the hybrid smoother level (tests/test_ref_pin_par.py); its contraction step and
 * V-cycle / CG drivers are cross-checked against the assembled global operator only.
 *
 * Citations are relative to /root/reference/.
 *
 * Storage: block-CSR exactly like NGSolve's SparseMatrix<Mat<H,W,double>>: rowptr[n+1] (int64),
 * col[nnz] (int32, ascending in each row), val[nnz*bh*bw] (row-major bh x bw blocks).
 * Vectors are AoS: entry i occupies x[i*b .. i*b+b).
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef int64_t i64;
typedef int32_t i32;

#define ORC_MAXB 6

/* ------------------------------------------------------------------------------------------
 * dense helpers
 * ---------------------------------------------------------------------------------------- */

/* c(hxw) = a(hxk) * b(kxw), plain triple loop, sum formed first (expression `vala * valb`) */
static void blk_mul(int h, int k, int w, const double *a, const double *b, double *c)
{
  for (int i = 0; i < h; i++)
    for (int j = 0; j < w; j++) {
      double s = 0.0;
      for (int l = 0; l < k; l++) s += a[i * k + l] * b[l * w + j];
      c[i * w + j] = s;
    }
}

/* in-place inverse of an n x n row-major matrix, Gauss-Jordan with partial pivoting.
   restates NGSolve CalcInverse(Mat<N,N>) (gssmoother.cpp:164); any exact inverse agrees to rounding.
   returns 0 ok, 1 singular */
static int dense_inverse(int n, double *a)
{
  double w[ORC_MAXB * 2 * ORC_MAXB];
  if (n > ORC_MAXB) return 2;
  if (n == 1) { if (a[0] == 0.0) return 1; a[0] = 1.0 / a[0]; return 0; }
  int n2 = 2 * n;
  for (int i = 0; i < n; i++)
    for (int j = 0; j < n; j++) {
      w[i * n2 + j] = a[i * n + j];
      w[i * n2 + n + j] = (i == j) ? 1.0 : 0.0;
    }
  for (int c = 0; c < n; c++) {
    int p = c;
    double best = fabs(w[c * n2 + c]);
    for (int r = c + 1; r < n; r++)
      if (fabs(w[r * n2 + c]) > best) { best = fabs(w[r * n2 + c]); p = r; }
    if (best == 0.0) return 1;
    if (p != c)
      for (int j = 0; j < n2; j++) { double t = w[c * n2 + j]; w[c * n2 + j] = w[p * n2 + j]; w[p * n2 + j] = t; }
    double piv = 1.0 / w[c * n2 + c];
    for (int j = 0; j < n2; j++) w[c * n2 + j] *= piv;
    for (int r = 0; r < n; r++) {
      if (r == c) continue;
      double f = w[r * n2 + c];
      if (f == 0.0) continue;
      for (int j = 0; j < n2; j++) w[r * n2 + j] -= f * w[c * n2 + j];
    }
  }
  for (int i = 0; i < n; i++)
    for (int j = 0; j < n; j++) a[i * n + j] = w[i * n2 + n + j];
  return 0;
}

/* symmetric eigen-decomposition by cyclic Jacobi rotations: a (n x n, symmetric, destroyed) ->
   evals[n], evecs rows = eigenvectors.  Stands in for LapackEigenValuesSymmetricLH
   (utils_denseLA.hpp:1486). */
static void sym_eig(int n, double *a, double *evals, double *evecs)
{
  for (int i = 0; i < n; i++)
    for (int j = 0; j < n; j++) evecs[i * n + j] = (i == j) ? 1.0 : 0.0;
  for (int sweep = 0; sweep < 100; sweep++) {
    double off = 0.0;
    for (int i = 0; i < n; i++)
      for (int j = i + 1; j < n; j++) off += a[i * n + j] * a[i * n + j];
    if (off < 1e-300) break;
    for (int p = 0; p < n; p++)
      for (int q = p + 1; q < n; q++) {
        double apq = a[p * n + q];
        if (fabs(apq) < 1e-300) continue;
        double theta = (a[q * n + q] - a[p * n + p]) / (2.0 * apq);
        double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < n; k++) {
          double akp = a[k * n + p], akq = a[k * n + q];
          a[k * n + p] = c * akp - s * akq;
          a[k * n + q] = s * akp + c * akq;
        }
        for (int k = 0; k < n; k++) {
          double apk = a[p * n + k], aqk = a[q * n + k];
          a[p * n + k] = c * apk - s * aqk;
          a[q * n + k] = s * apk + c * aqk;
        }
        for (int k = 0; k < n; k++) {
          double vpk = evecs[p * n + k], vqk = evecs[q * n + k];
          evecs[p * n + k] = c * vpk - s * vqk;
          evecs[q * n + k] = s * vpk + c * vqk;
        }
      }
  }
  for (int i = 0; i < n; i++) evals[i] = a[i * n + i];
}

/* pseudo inverse with relative tolerance, utils_denseLA.hpp:1474-1519
   (CalcPseudoInverseWithTolNonZeroBlock): eigenvalues <= relTol*mean(evals) are treated as kernel. */
static void dense_pinv_tol(int n, double *m, double reltol)
{
  double a[ORC_MAXB * ORC_MAXB], ev[ORC_MAXB], V[ORC_MAXB * ORC_MAXB];
  memcpy(a, m, sizeof(double) * n * n);
  sym_eig(n, a, ev, V);
  double tol = 0;
  for (int i = 0; i < n; i++) tol += ev[i];
  tol = reltol * tol / n;
  if (tol < 1e-20) tol = 1e-20;
  for (int i = 0; i < n; i++) ev[i] = (ev[i] > tol) ? 1.0 / ev[i] : 0.0;
  for (int i = 0; i < n; i++)
    for (int j = 0; j < n; j++) {
      double s = 0;
      for (int k = 0; k < n; k++) s += V[k * n + i] * ev[k] * V[k * n + j];
      m[i * n + j] = s;
    }
}

/* TryDirectInverse_simple (utils_denseLA.cpp:458-555): in-place Gauss-Jordan with COLUMN pivoting (the pivot of step j is
   searched along row j).  Fails -- a untouched, returns 0 -- when a pivot is smaller than max(AbsZeroTol * rest, eps) with
   eps = RelZeroTol * max_j a(j,j), rest = sum_{i>j} |inv(r,i)|.  RelZeroTol = 1e-12, AbsZeroTol = 1e-20 (utils_denseLA.hpp:93-117). */
static int try_direct_inverse_simple(int n, double *a)
{
  if (n == 0) return 0;
  double eps = 0;
  for (int j = 0; j < n; j++) if (a[j * n + j] > eps) eps = a[j * n + j];
  eps = 1e-12 * eps;
  double inv[ORC_MAXB * ORC_MAXB], hv[ORC_MAXB];
  int p[ORC_MAXB];
  memcpy(inv, a, sizeof(double) * n * n);
  for (int j = 0; j < n; j++) p[j] = j;
  for (int j = 0; j < n; j++) {
    double maxval = fabs(inv[j * n + j]);
    int r = j;
    for (int i = j + 1; i < n; i++)
      if (fabs(inv[j * n + i]) > maxval) { r = i; maxval = fabs(inv[j * n + i]); }
    double rest = 0.0;
    for (int i = j + 1; i < n; i++) rest += fabs(inv[r * n + i]);
    double lim = 1e-20 * rest;
    if (eps > lim) lim = eps;
    if (maxval < lim) return 0;
    if (r > j) {
      for (int k = 0; k < n; k++) { double t = inv[k * n + j]; inv[k * n + j] = inv[k * n + r]; inv[k * n + r] = t; }
      int t = p[j]; p[j] = p[r]; p[r] = t;
    }
    double hr = 1.0 / inv[j * n + j];
    for (int i = 0; i < n; i++) inv[j * n + i] = hr * inv[j * n + i];
    inv[j * n + j] = hr;
    for (int k = 0; k < n; k++)
      if (k != j) {
        double help = inv[k * n + j];
        double h = help * hr;
        for (int i = 0; i < n; i++) inv[k * n + i] -= help * inv[j * n + i];
        inv[k * n + j] = -h;
      }
  }
  for (int i = 0; i < n; i++) { /* row exchange */
    for (int k = 0; k < n; k++) hv[p[k]] = inv[k * n + i];
    for (int k = 0; k < n; k++) inv[k * n + i] = hv[k];
  }
  memcpy(a, inv, sizeof(double) * n * n);
  return 1;
}

/* CalcPseudoInverseTryNormal(Mat<N,N>&) (utils_denseLA.hpp:1549-1562) through CallOnNonZeroDiagonalBlock<1> (:1237-1405):
   rows whose diagonal entry is not above max(AbsZeroTol, RelZeroTol * max diagonal) are dropped (and come back as zero rows /
   columns); on the remaining block the direct inverse is tried, the eigenvalue pseudo inverse is the fall-back.
   Scalar overload (:1564-1569): 1/x if |x| > AbsZeroTol, else 0. */
static void dense_pinv_try_normal(int n, double *m)
{
  if (n == 1) { m[0] = (fabs(m[0]) > 1e-20) ? 1.0 / m[0] : 0.0; return; }
  double maxd = 0.0;
  for (int i = 0; i < n; i++) if (m[i * n + i] > maxd) maxd = m[i * n + i];
  double thresh = 1e-12 * maxd;
  if (thresh < 1e-20) thresh = 1e-20;
  int idx[ORC_MAXB], k = 0;
  for (int i = 0; i < n; i++)
    if (m[i * n + i] > thresh) idx[k++] = i;
  double sub[ORC_MAXB * ORC_MAXB];
  for (int i = 0; i < k; i++)
    for (int j = 0; j < k; j++) sub[i * k + j] = m[idx[i] * n + idx[j]];
  if (k > 0 && !try_direct_inverse_simple(k, sub)) dense_pinv_tol(k, sub, 1e-12);
  for (int i = 0; i < n * n; i++) m[i] = 0.0;
  for (int i = 0; i < k; i++)
    for (int j = 0; j < k; j++) m[idx[i] * n + idx[j]] = sub[i * k + j];
}

/* y += s * A * x  (SparseMatrix::MultAdd; used by base_smoother.hpp:140, dof_map.cpp:651,708) */
void orc_spmv_add(i64 n, int bh, int bw, const i64 *rp, const i32 *ci, const double *v,
                  double s, const double *x, double *y)
{
  for (i64 i = 0; i < n; i++) {
    double acc[ORC_MAXB] = {0};
    for (i64 k = rp[i]; k < rp[i + 1]; k++) {
      const double *blk = v + k * bh * bw;
      const double *xj = x + (i64)ci[k] * bw;
      for (int r = 0; r < bh; r++) { /* sum += data[j] * x(col[j]): the block product is evaluated first, then added */
        double t = blk[r * bw] * xj[0];
        for (int c = 1; c < bw; c++) t += blk[r * bw + c] * xj[c];
        acc[r] += t;
      }
    }
    for (int r = 0; r < bh; r++) y[i * bh + r] += s * acc[r];
  }
}

/* TransposeSPMImpl, utils_sparseMM.cpp:54-93: counting sort over columns, rows visited ascending
   (so every transposed row is already ascending; the trailing BubbleSort is a no-op), each block
   transposed (Trans(...), :79). */
void orc_transpose(i64 n, i64 m, int bh, int bw, const i64 *rp, const i32 *ci, const double *v,
                   i64 *trp, i32 *tci, double *tv)
{
  i64 *cnt = (i64 *)calloc((size_t)m + 1, sizeof(i64));
  for (i64 i = 0; i < n; i++)
    for (i64 k = rp[i]; k < rp[i + 1]; k++) cnt[ci[k]]++;
  trp[0] = 0;
  for (i64 c = 0; c < m; c++) trp[c + 1] = trp[c] + cnt[c];
  for (i64 c = 0; c < m; c++) cnt[c] = 0;
  for (i64 i = 0; i < n; i++)
    for (i64 k = rp[i]; k < rp[i + 1]; k++) {
      i32 c = ci[k];
      i64 pos = trp[c] + cnt[c]++;
      tci[pos] = (i32)i;
      const double *src = v + k * bh * bw;
      double *dst = tv + pos * bh * bw;
      for (int r = 0; r < bh; r++)
        for (int q = 0; q < bw; q++) dst[q * bh + r] = src[r * bw + q];
    }
  free(cnt);
}

static int cmp_i32(const void *a, const void *b)
{
  i32 x = *(const i32 *)a, y = *(const i32 *)b;
  return (x > y) - (x < y);
}

/* sorted set union of the B-rows named by A-row i == what MergeArrays emits
   (utils_sparseMM.cpp:140,168).  returns the number of distinct columns, written to out. */
static i64 merged_row(const i64 *a_rp, const i32 *a_ci, const i64 *b_rp, const i32 *b_ci, i64 i,
                      i32 **buf, i64 *cap)
{
  i64 tot = 0;
  for (i64 k = a_rp[i]; k < a_rp[i + 1]; k++) tot += b_rp[a_ci[k] + 1] - b_rp[a_ci[k]];
  if (tot > *cap) { *cap = 2 * tot + 64; *buf = (i32 *)realloc(*buf, sizeof(i32) * (size_t)*cap); }
  i64 p = 0;
  for (i64 k = a_rp[i]; k < a_rp[i + 1]; k++) {
    i64 rb = a_ci[k];
    for (i64 q = b_rp[rb]; q < b_rp[rb + 1]; q++) (*buf)[p++] = b_ci[q];
  }
  if (p == 0) return 0;
  qsort(*buf, (size_t)p, sizeof(i32), cmp_i32);
  i64 u = 1;
  for (i64 q = 1; q < p; q++)
    if ((*buf)[q] != (*buf)[u - 1]) (*buf)[u++] = (*buf)[q];
  return u;
}

/* MatMultABImpl symbolic phase 1, utils_sparseMM.cpp:122-146: per-row count of the merged pattern.
   Structural: numerical zeros are never dropped.  Fills c_rp[nA+1], returns nnz(C). */
i64 orc_matmul_count(i64 nA, const i64 *a_rp, const i32 *a_ci, const i64 *b_rp, const i32 *b_ci, i64 *c_rp)
{
  i32 *buf = NULL;
  i64 cap = 0;
  c_rp[0] = 0;
  for (i64 i = 0; i < nA; i++) c_rp[i + 1] = c_rp[i] + merged_row(a_rp, a_ci, b_rp, b_ci, i, &buf, &cap);
  free(buf);
  return c_rp[nA];
}

/* MatMultABImpl phases 2+3, utils_sparseMM.cpp:150-224: column fill (ascending) and the numeric
   phase: 2048-slot direct-mapped hash (col & (nhash-1)), binary-search fallback on a collision;
   accumulation order = A-row order, then B-row order; C(i,col) += vala * valb. */
void orc_matmul_fill(i64 nA, int ah, int aw, int bw, const i64 *a_rp, const i32 *a_ci, const double *a_v,
                     const i64 *b_rp, const i32 *b_ci, const double *b_v, const i64 *c_rp, i32 *c_ci,
                     double *c_v)
{
  i32 *buf = NULL;
  i64 cap = 0;
  i64 maxci = 0;
  for (i64 i = 0; i < nA; i++) {
    i64 u = merged_row(a_rp, a_ci, b_rp, b_ci, i, &buf, &cap);
    memcpy(c_ci + c_rp[i], buf, sizeof(i32) * (size_t)u);
    if (u > maxci) maxci = u;
  }
  free(buf);
  memset(c_v, 0, sizeof(double) * (size_t)c_rp[nA] * ah * bw);
  i64 nhash = 2048;
  while (nhash < 2 * maxci) nhash *= 2;
  i32 *hidx = (i32 *)malloc(sizeof(i32) * (size_t)nhash);
  i32 *hpos = (i32 *)malloc(sizeof(i32) * (size_t)nhash);
  for (i64 q = 0; q < nhash; q++) { hidx[q] = -1; hpos[q] = 0; }
  double prod[ORC_MAXB * ORC_MAXB];
  const int cbs = ah * bw;
  for (i64 i = 0; i < nA; i++) {
    const i32 *cci = c_ci + c_rp[i];
    double *cv = c_v + c_rp[i] * cbs;
    const i64 ncol = c_rp[i + 1] - c_rp[i];
    for (i64 k = 0; k < ncol; k++) {
      i64 h = (i64)((uint32_t)cci[k]) & (nhash - 1);
      hpos[h] = (i32)k;
      hidx[h] = cci[k];
    }
    for (i64 j = a_rp[i]; j < a_rp[i + 1]; j++) {
      const double *va = a_v + j * ah * aw;
      i64 rowb = a_ci[j];
      for (i64 k = b_rp[rowb]; k < b_rp[rowb + 1]; k++) {
        i32 colb = b_ci[k];
        blk_mul(ah, aw, bw, va, b_v + k * aw * bw, prod);
        i64 h = (i64)((uint32_t)colb) & (nhash - 1);
        i64 pos;
        if (hidx[h] == colb) pos = hpos[h];
        else { /* binary search, (*prod)(i,colb) */
          i64 lo = 0, hi = ncol - 1;
          pos = -1;
          while (lo <= hi) {
            i64 mid = (lo + hi) / 2;
            if (cci[mid] == colb) { pos = mid; break; }
            if (cci[mid] < colb) lo = mid + 1; else hi = mid - 1;
          }
        }
        double *dst = cv + pos * cbs;
        for (int e = 0; e < cbs; e++) dst[e] += prod[e];
      }
    }
  }
  free(hidx);
  free(hpos);
}

/* GSS3::CalcDiags, gssmoother.cpp:142-170: dinv[i] = inv(A(i,i)) (or repl_diag[i]), pseudo inverse
   if pinv, 0 on non-free rows. */
int orc_calc_dinv(i64 n, int b, const i64 *rp, const i32 *ci, const double *v, const uint8_t *freed,
                  int pinv, const double *repl_diag, double *dinv)
{
  int bb = b * b, rc = 0;
  for (i64 i = 0; i < n; i++) {
    double *d = dinv + i * bb;
    if (freed && !freed[i]) { for (int e = 0; e < bb; e++) d[e] = 0.0; continue; }
    if (repl_diag) memcpy(d, repl_diag + i * bb, sizeof(double) * bb);
    else {
      int found = 0;
      for (i64 k = rp[i]; k < rp[i + 1]; k++)
        if (ci[k] == i) { memcpy(d, v + k * bb, sizeof(double) * bb); found = 1; break; }
      if (!found) { for (int e = 0; e < bb; e++) d[e] = 0.0; rc = 3; continue; }
    }
    if (pinv) dense_pinv_try_normal(b, d);
    else if (dense_inverse(b, d)) rc = 1;
  }
  return rc;
}

/* ------------------------------------------------------------------------------------------
 * GSS3 sweeps
 * ---------------------------------------------------------------------------------------- */

typedef struct {
  i64 n;
  int b;
  const i64 *rp;
  const i32 *ci;
  const double *v;
  const double *dinv;
  const uint8_t *freed; /* NULL = all free */
  i64 first_free, next_free;
} gss3;

/* GSS3::SetUp, gssmoother.cpp:110-139 */
static void gss3_setup(gss3 *g)
{
  g->first_free = 0;
  g->next_free = g->n;
  if (g->freed) {
    i64 nset = 0;
    for (i64 i = 0; i < g->n; i++) nset += g->freed[i] ? 1 : 0;
    if (nset == 0) g->next_free = 0;
    else if (nset != g->n) {
      for (i64 c = 0; c < g->n; c++) if (g->freed[c]) { g->first_free = c; break; }
      for (i64 c = g->n - 1; c >= 0; c--) if (g->freed[c]) { g->next_free = c + 1; break; }
    }
  }
}

/* updateRow of SmoothRHSInternal, gssmoother.cpp:209-212:
     r = A.RowTimesVector(row, x);  x(row) += dinv[row] * (b(row) - r)                        */
static inline void gss3_row_rhs(const gss3 *g, i64 i, double *x, const double *rhs)
{
  const int b = g->b, bb = b * b;
  double r[ORC_MAXB] = {0};
  for (i64 k = g->rp[i]; k < g->rp[i + 1]; k++) {
    const double *blk = g->v + k * bb;
    const double *xj = x + (i64)g->ci[k] * b;
    for (int p = 0; p < b; p++) { /* sum += A(row,col) * x(col): the block product is evaluated first, then added */
      double t = blk[p * b] * xj[0];
      for (int q = 1; q < b; q++) t += blk[p * b + q] * xj[q];
      r[p] += t;
    }
  }
  double d[ORC_MAXB];
  for (int p = 0; p < b; p++) d[p] = rhs[i * b + p] - r[p];
  const double *di = g->dinv + i * bb;
  for (int p = 0; p < b; p++) {
    double s = 0;
    for (int q = 0; q < b; q++) s += di[p * b + q] * d[q];
    x[i * b + p] += s;
  }
}

/* up_row of SmoothRESInternal, gssmoother.cpp:274-278:
     w = -dinv[row]*res(row);  A.AddRowTransToVector(row, w, res);  x(row) -= w
   AddRowTransToVector: res(col_k) += Trans(A(row,col_k)) * w  for every entry of the row. */
static inline void gss3_row_res(const gss3 *g, i64 i, double *x, double *res)
{
  const int b = g->b, bb = b * b;
  const double *di = g->dinv + i * bb;
  double w[ORC_MAXB];
  for (int p = 0; p < b; p++) {
    double s = 0;
    for (int q = 0; q < b; q++) s += di[p * b + q] * res[i * b + q];
    w[p] = -s;
  }
  for (i64 k = g->rp[i]; k < g->rp[i + 1]; k++) {
    const double *blk = g->v + k * bb;
    double *rj = res + (i64)g->ci[k] * b;
    for (int q = 0; q < b; q++) {
      double s = 0;
      for (int p = 0; p < b; p++) s += blk[p * b + q] * w[p];
      rj[q] += s;
    }
  }
  for (int p = 0; p < b; p++) x[i * b + p] -= w[p];
}

/* SmoothRHSInternal, gssmoother.cpp:195-257 */
static void gss3_smooth_rhs(const gss3 *g, i64 first, i64 next, double *x, const double *rhs, int backwards)
{
  i64 uf = first > g->first_free ? first : g->first_free;
  i64 un = next < g->next_free ? next : g->next_free;
  if (!backwards) {
    for (i64 i = uf; i < un; i++)
      if (!g->freed || g->freed[i]) gss3_row_rhs(g, i, x, rhs);
  } else {
    for (i64 i = un - 1; i >= uf; i--)
      if (!g->freed || g->freed[i]) gss3_row_rhs(g, i, x, rhs);
  }
}

/* SmoothRESInternal, gssmoother.cpp:260-315 */
static void gss3_smooth_res(const gss3 *g, i64 first, i64 next, double *x, double *res, int backwards)
{
  i64 uf = first > g->first_free ? first : g->first_free;
  i64 un = next < g->next_free ? next : g->next_free;
  if (!backwards) {
    for (i64 i = uf; i < un; i++)
      if (!g->freed || g->freed[i]) gss3_row_res(g, i, x, res);
  } else {
    for (i64 i = un - 1; i >= uf; i--)
      if (!g->freed || g->freed[i]) gss3_row_res(g, i, x, res);
  }
}

/* standalone entry points for unit tests */
void orc_gs_rhs(i64 n, int b, const i64 *rp, const i32 *ci, const double *v, const double *dinv,
                const uint8_t *freed, double *x, const double *rhs, int backwards)
{
  gss3 g = {n, b, rp, ci, v, dinv, freed, 0, 0};
  gss3_setup(&g);
  gss3_smooth_rhs(&g, 0, n, x, rhs, backwards);
}
void orc_gs_res(i64 n, int b, const i64 *rp, const i32 *ci, const double *v, const double *dinv,
                const uint8_t *freed, double *x, double *res, int backwards)
{
  gss3 g = {n, b, rp, ci, v, dinv, freed, 0, 0};
  gss3_setup(&g);
  gss3_smooth_res(&g, 0, n, x, res, backwards);
}

/* ------------------------------------------------------------------------------------------
 * hierarchy: levels, smoothers, transfers, V-cycle, PCG
 * ---------------------------------------------------------------------------------------- */

enum { ORC_SM_GS = 0, ORC_SM_JACOBI = 1 };

typedef struct {
  i64 n;
  int b;
  i64 *rp;
  i32 *ci;
  double *v;
  uint8_t *freed; /* NULL = all free */
  double *dinv;
  gss3 gs;
  int sm_type, sm_steps, sm_symm, pinv;
  double omega;
  /* prolongation to level l+1 (absent on the last level) */
  i64 nc;
  int bc;
  i64 *p_rp; i32 *p_ci; double *p_v;    /* P : n x nc, blocks b x bc */
  i64 *pt_rp; i32 *pt_ci; double *pt_v; /* PT: nc x n, blocks bc x b */
  double *x, *rhs, *res, *tmp;
} orc_level;

typedef struct {
  int nlevels;
  orc_level *lev;
  /* coarsest exact solve: dense Cholesky factor of the free sub-matrix */
  i64 cn;          /* number of free scalar dofs */
  i64 *cdofs;      /* scalar dof numbers of the free dofs */
  double *cchol;   /* cn x cn lower factor L, row-major */
  int has_cinv;
} orc_amg;

orc_amg *orc_amg_new(int nlevels)
{
  orc_amg *a = (orc_amg *)calloc(1, sizeof(orc_amg));
  a->nlevels = nlevels;
  a->lev = (orc_level *)calloc((size_t)nlevels, sizeof(orc_level));
  return a;
}

static void *dupmem(const void *p, size_t bytes)
{
  void *q = malloc(bytes ? bytes : 1);
  if (bytes) memcpy(q, p, bytes);
  return q;
}

static void level_alloc_vecs(orc_level *L)
{
  size_t nb = (size_t)L->n * L->b;
  free(L->x); free(L->rhs); free(L->res); free(L->tmp);
  L->x = (double *)calloc(nb ? nb : 1, sizeof(double));
  L->rhs = (double *)calloc(nb ? nb : 1, sizeof(double));
  L->res = (double *)calloc(nb ? nb : 1, sizeof(double));
  L->tmp = (double *)calloc(nb ? nb : 1, sizeof(double));
}

/* set the (copied) matrix of a level; level 0 comes from the caller, others from orc_amg_galerkin */
void orc_amg_set_matrix(orc_amg *a, int l, i64 n, int b, const i64 *rp, const i32 *ci, const double *v,
                        const uint8_t *freed)
{
  orc_level *L = &a->lev[l];
  L->n = n; L->b = b;
  free(L->rp); free(L->ci); free(L->v); free(L->freed);
  L->rp = (i64 *)dupmem(rp, sizeof(i64) * (size_t)(n + 1));
  L->ci = (i32 *)dupmem(ci, sizeof(i32) * (size_t)rp[n]);
  L->v = (double *)dupmem(v, sizeof(double) * (size_t)rp[n] * b * b);
  L->freed = freed ? (uint8_t *)dupmem(freed, (size_t)n) : NULL;
  level_alloc_vecs(L);
}

/* smoother selection, amg_pc.cpp:1033-1138 (gs -> GSS3 serial; jacobi -> JacobiSmoother, omega .9
   default base_smoother.hpp:256) wrapped into a ProxySmoother iff symm || steps>1 (:1079-1082) */
int orc_amg_set_smoother(orc_amg *a, int l, int sm_type, int sm_steps, int sm_symm, int pinv, double omega)
{
  orc_level *L = &a->lev[l];
  L->sm_type = sm_type; L->sm_steps = sm_steps; L->sm_symm = sm_symm; L->pinv = pinv; L->omega = omega;
  free(L->dinv);
  L->dinv = (double *)malloc(sizeof(double) * (size_t)(L->n ? L->n : 1) * L->b * L->b);
  /* JacobiSmoother ctor (base_smoother.cpp:87-114) always uses CalcInverse */
  int rc = orc_calc_dinv(L->n, L->b, L->rp, L->ci, L->v, L->freed, sm_type == ORC_SM_GS ? pinv : 0, NULL, L->dinv);
  gss3 g = {L->n, L->b, L->rp, L->ci, L->v, L->dinv, L->freed, 0, 0};
  gss3_setup(&g);
  L->gs = g;
  return rc;
}

/* ProlMap: store P, build PT (ProlMap::BuildPT, dof_map.cpp:807-814) */
void orc_amg_set_prol(orc_amg *a, int l, i64 nc, int bc, const i64 *rp, const i32 *ci, const double *v)
{
  orc_level *L = &a->lev[l];
  L->nc = nc; L->bc = bc;
  i64 nnz = rp[L->n];
  free(L->p_rp); free(L->p_ci); free(L->p_v); free(L->pt_rp); free(L->pt_ci); free(L->pt_v);
  L->p_rp = (i64 *)dupmem(rp, sizeof(i64) * (size_t)(L->n + 1));
  L->p_ci = (i32 *)dupmem(ci, sizeof(i32) * (size_t)nnz);
  L->p_v = (double *)dupmem(v, sizeof(double) * (size_t)nnz * L->b * bc);
  L->pt_rp = (i64 *)malloc(sizeof(i64) * (size_t)(nc + 1));
  L->pt_ci = (i32 *)malloc(sizeof(i32) * (size_t)(nnz ? nnz : 1));
  L->pt_v = (double *)malloc(sizeof(double) * (size_t)(nnz ? nnz : 1) * L->b * bc);
  orc_transpose(L->n, nc, L->b, bc, L->p_rp, L->p_ci, L->p_v, L->pt_rp, L->pt_ci, L->pt_v);
}

/* ProlMap::AssembleMatrix -> RestrictMatrix (dof_map.cpp:817-834, utils_sparseMM.hpp:93-109):
   A_{l+1} = (PT * A) * P -- in this association.  Coarse levels have no Dirichlet dofs. */
void orc_amg_galerkin(orc_amg *a, int l)
{
  orc_level *F = &a->lev[l];
  orc_level *C = &a->lev[l + 1];
  const i64 nc = F->nc;
  const int bf = F->b, bc = F->bc;
  i64 *t_rp = (i64 *)malloc(sizeof(i64) * (size_t)(nc + 1));
  i64 tnnz = orc_matmul_count(nc, F->pt_rp, F->pt_ci, F->rp, F->ci, t_rp);
  i32 *t_ci = (i32 *)malloc(sizeof(i32) * (size_t)(tnnz ? tnnz : 1));
  double *t_v = (double *)malloc(sizeof(double) * (size_t)(tnnz ? tnnz : 1) * bc * bf);
  orc_matmul_fill(nc, bc, bf, bf, F->pt_rp, F->pt_ci, F->pt_v, F->rp, F->ci, F->v, t_rp, t_ci, t_v);
  i64 *c_rp = (i64 *)malloc(sizeof(i64) * (size_t)(nc + 1));
  i64 cnnz = orc_matmul_count(nc, t_rp, t_ci, F->p_rp, F->p_ci, c_rp);
  i32 *c_ci = (i32 *)malloc(sizeof(i32) * (size_t)(cnnz ? cnnz : 1));
  double *c_v = (double *)malloc(sizeof(double) * (size_t)(cnnz ? cnnz : 1) * bc * bc);
  orc_matmul_fill(nc, bc, bf, bc, t_rp, t_ci, t_v, F->p_rp, F->p_ci, F->p_v, c_rp, c_ci, c_v);
  free(t_rp); free(t_ci); free(t_v);
  free(C->rp); free(C->ci); free(C->v); free(C->freed);
  C->n = nc; C->b = bc; C->rp = c_rp; C->ci = c_ci; C->v = c_v; C->freed = NULL;
  level_alloc_vecs(C);
}

/* coarsest level exact solve, amg_pc.cpp:843-928: cspm->InverseMatrix(free) with SPARSECHOLESKY
   (serial).  Restated as a dense Cholesky of the scalar-expanded free sub-matrix (exact up to
   rounding, like any direct solver).  returns 0 ok, 1 not positive definite */
/* RegTM<0,6,6> (utils_denseLA.hpp:1198-1234): eigenvalues <= max(1e-15, 1e-12 * largest) count as zero, the smallest non-zero
   eigenvalue is added along every zero eigenvector, an all-zero block becomes the identity. */
static void reg_tm6(double *m)
{
  const int n = 6;
  double a[36], ev[6], V[36];
  memcpy(a, m, sizeof(a));
  sym_eig(n, a, ev, V);
  int order[6];
  for (int i = 0; i < n; i++) order[i] = i;
  for (int i = 0; i < n; i++)
    for (int j = i + 1; j < n; j++)
      if (ev[order[j]] < ev[order[i]]) { int t = order[i]; order[i] = order[j]; order[j] = t; }
  double eps = 1e-12 * ev[order[n - 1]];
  if (eps < 1e-15) eps = 1e-15;
  double min_nzev = 0.0;
  int nzero = 0;
  for (int k = 0; k < n; k++) {
    if (ev[order[k]] > eps) { min_nzev = ev[order[k]]; break; }
    nzero++;
  }
  if (nzero < n) {
    for (int l = 0; l < nzero; l++) {
      const double *v = V + order[l] * n;
      for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++) m[i * n + j] += min_nzev * v[i] * v[j];
    }
  } else
    for (int i = 0; i < n; i++)
      for (int j = 0; j < n; j++) m[i * n + j] = (i == j) ? 1.0 : 0.0;
}

/* one diagonal block, VertexAMGPC<ElasticityAMGFactory<DIM>>::RegularizeMatrix, local branch (elasticity_pc_impl.hpp:711-763) */
void orc_regularize_block(int n, double *m, int dim)
{
  if (dim == 2) { if (n == 3 && fabs(m[8]) < 1e-8) m[8] = 1.0; }
  else if (dim == 3 && n == 6) reg_tm6(m);
}

/* RegularizeMatrix on the coarsest matrix before it is inverted (amg_pc.cpp:861-862), O.regularize_cmats */
void orc_amg_regularize_coarse(orc_amg *a, int dim)
{
  orc_level *L = &a->lev[a->nlevels - 1];
  const int bb = L->b * L->b;
  for (i64 i = 0; i < L->n; i++)
    for (i64 k = L->rp[i]; k < L->rp[i + 1]; k++)
      if (L->ci[k] == i) orc_regularize_block(L->b, L->v + k * bb, dim);
}

int orc_amg_set_coarse_inv(orc_amg *a)
{
  orc_level *L = &a->lev[a->nlevels - 1];
  const int b = L->b;
  i64 cn = 0;
  i64 *glob2loc = (i64 *)malloc(sizeof(i64) * (size_t)(L->n * b + 1));
  free(a->cdofs);
  a->cdofs = (i64 *)malloc(sizeof(i64) * (size_t)(L->n * b + 1));
  for (i64 i = 0; i < L->n; i++)
    for (int p = 0; p < b; p++) {
      if (!L->freed || L->freed[i]) { glob2loc[i * b + p] = cn; a->cdofs[cn++] = i * b + p; }
      else glob2loc[i * b + p] = -1;
    }
  double *M = (double *)calloc((size_t)(cn * cn + 1), sizeof(double));
  for (i64 i = 0; i < L->n; i++)
    for (i64 k = L->rp[i]; k < L->rp[i + 1]; k++)
      for (int p = 0; p < b; p++)
        for (int q = 0; q < b; q++) {
          i64 r = glob2loc[i * b + p], c = glob2loc[(i64)L->ci[k] * b + q];
          if (r >= 0 && c >= 0) M[r * cn + c] = L->v[k * b * b + p * b + q];
        }
  int rc = 0;
  for (i64 j = 0; j < cn && !rc; j++) {
    double d = M[j * cn + j];
    for (i64 k = 0; k < j; k++) d -= M[j * cn + k] * M[j * cn + k];
    if (!(d > 0.0)) { rc = 1; break; }
    d = sqrt(d);
    M[j * cn + j] = d;
    for (i64 i = j + 1; i < cn; i++) {
      double s = M[i * cn + j];
      for (i64 k = 0; k < j; k++) s -= M[i * cn + k] * M[j * cn + k];
      M[i * cn + j] = s / d;
    }
  }
  free(glob2loc);
  free(a->cchol);
  a->cchol = M; a->cn = cn; a->has_cinv = (rc == 0);
  return rc;
}

static void coarse_solve(const orc_amg *a, const double *rhs, double *x)
{
  const orc_level *L = &a->lev[a->nlevels - 1];
  i64 nb = L->n * L->b, cn = a->cn;
  for (i64 i = 0; i < nb; i++) x[i] = 0.0;
  if (!a->has_cinv) return; /* clev=none: x_L = 0, amg_matrix.cpp:228-229 */
  double *y = (double *)malloc(sizeof(double) * (size_t)(cn + 1));
  const double *M = a->cchol;
  for (i64 i = 0; i < cn; i++) {
    double s = rhs[a->cdofs[i]];
    for (i64 k = 0; k < i; k++) s -= M[i * cn + k] * y[k];
    y[i] = s / M[i * cn + i];
  }
  for (i64 i = cn - 1; i >= 0; i--) {
    double s = y[i];
    for (i64 k = i + 1; k < cn; k++) s -= M[k * cn + i] * y[k];
    y[i] = s / M[i * cn + i];
  }
  for (i64 i = 0; i < cn; i++) x[a->cdofs[i]] = y[i];
  free(y);
}

/* BaseSmoother::CalcResiduum, base_smoother.hpp:132-142 */
static void calc_residuum(const orc_level *L, const double *x, const double *b, double *res, int x_zero)
{
  i64 nb = L->n * L->b;
  memcpy(res, b, sizeof(double) * (size_t)nb);
  if (!x_zero) orc_spmv_add(L->n, L->b, L->b, L->rp, L->ci, L->v, -1.0, x, res);
}

/* one (unwrapped) smoother step; GSS3::Smooth / SmoothBack (gssmoother.cpp:349-398) and
   RichardsonSmoother::Smooth with prec = diag^-1 (base_smoother.cpp:61-83) */
static void smooth_once(orc_level *L, double *x, const double *b, double *res, int res_updated,
                        int update_res, int x_zero, int backwards)
{
  i64 nb = L->n * L->b;
  if (L->sm_type == ORC_SM_GS) {
    if (res_updated) {
      if (update_res) gss3_smooth_res(&L->gs, 0, L->n, x, res, backwards);
      else gss3_smooth_rhs(&L->gs, 0, L->n, x, b, backwards);
    } else {
      if (update_res) {
        calc_residuum(L, x, b, res, x_zero);
        gss3_smooth_res(&L->gs, 0, L->n, x, res, backwards);
      } else gss3_smooth_rhs(&L->gs, 0, L->n, x, b, backwards);
    }
  } else { /* Jacobi == Richardson with diagonal inverse; SmoothBack == Smooth */
    const double *src;
    if (!res_updated && x_zero) src = b;
    else {
      if (!res_updated) calc_residuum(L, x, b, res, 0);
      src = res;
    }
    const int bs = L->b, bb = bs * bs;
    for (i64 i = 0; i < L->n; i++)
      for (int p = 0; p < bs; p++) {
        double s = 0;
        for (int q = 0; q < bs; q++) s += L->dinv[i * bb + p * bs + q] * src[i * bs + q];
        x[i * bs + p] += L->omega * s;
      }
    (void)nb;
    if (update_res) calc_residuum(L, x, b, res, 0);
  }
}

/* BaseSmoother::SmoothSymm / SmoothK / SmoothBackK / SmoothSymmK (base_smoother.hpp:79-112) and
   ProxySmoother::Smooth / SmoothBack (:181-196) */
static void smooth_symm(orc_level *L, double *x, const double *b, double *res, int ru, int ur, int xz)
{
  smooth_once(L, x, b, res, ru, ur, xz, 0);
  smooth_once(L, x, b, res, ur, ur, 0, 1);
}
static void level_smooth(orc_level *L, double *x, const double *b, double *res, int ru, int ur, int xz, int back)
{
  int k = L->sm_steps < 1 ? 1 : L->sm_steps;
  if (L->sm_symm) {
    smooth_symm(L, x, b, res, ru, ur, xz);
    for (int j = 0; j < k - 1; j++) smooth_symm(L, x, b, res, ur, ur, 0);
  } else {
    smooth_once(L, x, b, res, ru, ur, xz, back);
    for (int j = 0; j < k - 1; j++) smooth_once(L, x, b, res, ur, ur, 0, back);
  }
}

/* smoother-only entry (python_smoothers.cpp:144-194 analogue) on the smoother of level l */
void orc_amg_smooth(orc_amg *a, int l, double *x, const double *b, double *res, int res_updated,
                    int update_res, int x_zero, int backwards)
{
  level_smooth(&a->lev[l], x, b, res, res_updated, update_res, x_zero, backwards);
}

/* AMGMatrix::SmoothV, amg_matrix.cpp:160-307 (serial: Distribute/Cumulate are no-ops) */
void orc_amg_apply(orc_amg *a, const double *b, double *x)
{
  const int NL = a->nlevels;
  for (int l = 0; l + 1 < NL; l++) {
    orc_level *L = &a->lev[l];
    double *xl = (l == 0) ? x : L->x;
    const double *bl = (l == 0) ? b : L->rhs;
    i64 nb = L->n * L->b;
    memset(xl, 0, sizeof(double) * (size_t)nb);            /* :193 */
    memcpy(L->res, bl, sizeof(double) * (size_t)nb);       /* :201 */
    level_smooth(L, xl, bl, L->res, 1, 1, 1, 0);           /* :206 */
    /* TransferF2C, dof_map.cpp:633-654: rhs_{l+1} = PT * res_l */
    orc_level *C = &a->lev[l + 1];
    memset(C->rhs, 0, sizeof(double) * (size_t)(C->n * C->b));
    orc_spmv_add(L->nc, L->bc, L->b, L->pt_rp, L->pt_ci, L->pt_v, 1.0, L->res, C->rhs);
  }
  {
    orc_level *L = &a->lev[NL - 1];
    double *xl = (NL == 1) ? x : L->x;
    const double *bl = (NL == 1) ? b : L->rhs;
    coarse_solve(a, bl, xl);                               /* :217-247 */
  }
  for (int l = NL - 2; l >= 0; l--) {
    orc_level *L = &a->lev[l];
    orc_level *C = &a->lev[l + 1];
    double *xl = (l == 0) ? x : L->x;
    const double *bl = (l == 0) ? b : L->rhs;
    /* AddC2F, dof_map.cpp:694-709: x_l += P * x_{l+1} */
    orc_spmv_add(L->n, L->b, L->bc, L->p_rp, L->p_ci, L->p_v, 1.0, C->x, xl);
    level_smooth(L, xl, bl, L->res, 0, 0, 0, 1);           /* :302 */
  }
}

/* ---- W and BS cycles (amg_matrix.cpp:37-157, 307-374) ---------------------------------------------------- */
static void restrict_res(orc_amg *a, int l, const double *res)
{
  orc_level *L = &a->lev[l], *C = &a->lev[l + 1];
  memset(C->rhs, 0, sizeof(double) * (size_t)(C->n * C->b));
  orc_spmv_add(L->nc, L->bc, L->b, L->pt_rp, L->pt_ci, L->pt_v, 1.0, res, C->rhs);
}
static void prolong_add(orc_amg *a, int l, double *x)
{
  orc_level *L = &a->lev[l], *C = &a->lev[l + 1];
  orc_spmv_add(L->n, L->b, L->bc, L->p_rp, L->p_ci, L->p_v, 1.0, C->x, x);
}

/* the recursion of AMGMatrix::SmoothW (:46-104).  On level 0 the reference first runs a V-type visit and then the W-type visit
   below, which starts from x = 0 and res = b again: the first visit leaves no trace in the result, so only the second is run. */
static void w_visit(orc_amg *a, int l, double *xl, const double *bl)
{
  const int NL = a->nlevels;
  if (l + 1 >= NL) { coarse_solve(a, bl, xl); return; }
  orc_level *L = &a->lev[l], *C = &a->lev[l + 1];
  i64 nb = L->n * L->b;
  memset(xl, 0, sizeof(double) * (size_t)nb);
  memcpy(L->res, bl, sizeof(double) * (size_t)nb);
  level_smooth(L, xl, bl, L->res, 1, 1, 1, 0);
  restrict_res(a, l, L->res);
  w_visit(a, l + 1, C->x, C->rhs);
  prolong_add(a, l, xl);
  level_smooth(L, xl, bl, L->res, 0, 1, 0, 1);   /* SmoothBack(x, b, res, false, true, false)  :82 */
  level_smooth(L, xl, bl, L->res, 1, 1, 0, 0);   /* Smooth(x, b, res, true, true, false)       :83 */
  restrict_res(a, l, L->res);
  w_visit(a, l + 1, C->x, C->rhs);
  prolong_add(a, l, xl);
  level_smooth(L, xl, bl, L->res, 0, 0, 0, 1);   /* :89 */
}
void orc_amg_apply_w(orc_amg *a, const double *b, double *x) { w_visit(a, 0, x, b); }

/* AMGMatrix::SmoothVFromLevel (:310-374) */
static void v_from_level(orc_amg *a, int s, double *x, const double *b, double *res, int ru, int ur, int xz)
{
  const int NL = a->nlevels;
  level_smooth(&a->lev[s], x, b, res, ru, 1, xz, 0);
  restrict_res(a, s, res);
  for (int l = s + 1; l + 1 < NL; l++) {
    orc_level *L = &a->lev[l];
    i64 nb = L->n * L->b;
    memset(L->x, 0, sizeof(double) * (size_t)nb);
    memcpy(L->res, L->rhs, sizeof(double) * (size_t)nb);
    level_smooth(L, L->x, L->rhs, L->res, 1, 1, 1, 0);
    restrict_res(a, l, L->res);
  }
  coarse_solve(a, a->lev[NL - 1].rhs, a->lev[NL - 1].x);
  for (int l = NL - 2; l > s; l--) {
    orc_level *L = &a->lev[l];
    prolong_add(a, l, L->x);
    level_smooth(L, L->x, L->rhs, L->res, 0, 0, 0, 1);
  }
  prolong_add(a, s, x);
  level_smooth(&a->lev[s], x, b, res, 0, ur, 0, 1);
}

/* AMGMatrix::SmoothBS (:107-157): every level is "smoothed" by a V-cycle that starts there */
void orc_amg_apply_bs(orc_amg *a, const double *b, double *x)
{
  const int NL = a->nlevels;
  for (int l = 0; l + 1 < NL; l++) {
    orc_level *L = &a->lev[l];
    double *xl = (l == 0) ? x : L->x;
    const double *bl = (l == 0) ? b : L->rhs;
    i64 nb = L->n * L->b;
    memset(xl, 0, sizeof(double) * (size_t)nb);
    memcpy(L->res, bl, sizeof(double) * (size_t)nb);
    v_from_level(a, l, xl, bl, L->res, 1, 1, 1);
    restrict_res(a, l, L->res);
  }
  {
    orc_level *L = &a->lev[NL - 1];
    coarse_solve(a, (NL == 1) ? b : L->rhs, (NL == 1) ? x : L->x);
  }
  for (int l = NL - 2; l >= 0; l--) {
    orc_level *L = &a->lev[l];
    double *xl = (l == 0) ? x : L->x;
    const double *bl = (l == 0) ? b : L->rhs;
    prolong_add(a, l, xl);
    v_from_level(a, l, xl, bl, L->res, 0, 0, 0);
  }
}

/* AMGMatrix::MultAdd, amg_matrix.cpp:385-389: x += s * V(b) */
void orc_amg_apply_add(orc_amg *a, double s, const double *b, double *x)
{
  orc_level *L = &a->lev[0];
  i64 nb = L->n * L->b;
  double *t = (double *)malloc(sizeof(double) * (size_t)(nb + 1));
  orc_amg_apply(a, b, t);
  for (i64 i = 0; i < nb; i++) x[i] += s * t[i];
  free(t);
}

static double dot(i64 n, const double *a, const double *b)
{
  double s = 0;
  for (i64 i = 0; i < n; i++) s += a[i] * b[i];
  return s;
}

/* PCG == ngsolve.krylovspace.CGSolver(mat, pre, maxsteps, tol) as called by the reference tests
   (tests/h1/amg_utils.py:346-349).  NGSolve is not under /root/reference; restated from its
   published algorithm: u=0, d=rhs, w=C d, s=w, wdn=(w,d), err0=sqrt|wdn|; loop: w=A s, wd=wdn,
   alpha=wd/(s,w), u+=alpha s, d-=alpha w, w=C d, wdn=(w,d), beta=wdn/wd, s=beta s+w,
   err=sqrt|wd| (the value from BEFORE this update), stop when err < tol*err0.
   errors[0]=err0, errors[k]=err of iteration k.  returns the iteration count. */
int orc_amg_pcg(orc_amg *a, const double *rhs, double *u, double tol, int maxsteps, double *errors)
{
  orc_level *L = &a->lev[0];
  i64 nb = L->n * L->b;
  double *d = (double *)malloc(sizeof(double) * (size_t)(nb + 1));
  double *w = (double *)malloc(sizeof(double) * (size_t)(nb + 1));
  double *s = (double *)malloc(sizeof(double) * (size_t)(nb + 1));
  memset(u, 0, sizeof(double) * (size_t)nb);
  memcpy(d, rhs, sizeof(double) * (size_t)nb);
  orc_amg_apply(a, d, w);
  memcpy(s, w, sizeof(double) * (size_t)nb);
  double wdn = dot(nb, w, d);
  double err0 = sqrt(fabs(wdn));
  if (errors) errors[0] = err0;
  int it = 0;
  if (wdn != 0.0)
    for (it = 1; it <= maxsteps; it++) {
      memset(w, 0, sizeof(double) * (size_t)nb);
      orc_spmv_add(L->n, L->b, L->b, L->rp, L->ci, L->v, 1.0, s, w);
      double wd = wdn;
      double as_s = dot(nb, s, w);
      double alpha = wd / as_s;
      for (i64 i = 0; i < nb; i++) u[i] += alpha * s[i];
      for (i64 i = 0; i < nb; i++) d[i] -= alpha * w[i];
      orc_amg_apply(a, d, w);
      wdn = dot(nb, w, d);
      double beta = wdn / wd;
      for (i64 i = 0; i < nb; i++) s[i] = beta * s[i] + w[i];
      double err = sqrt(fabs(wd));
      if (errors) errors[it] = err;
      if (err < tol * err0) break;
    }
  if (it > maxsteps) it = maxsteps;
  free(d); free(w); free(s);
  return it;
}

/* ------------------------------------------------------------------------------------------
 * getters (tests read level matrices, DOF maps and work vectors through these)
 * ---------------------------------------------------------------------------------------- */
i64 orc_amg_level_n(orc_amg *a, int l) { return a->lev[l].n; }
int orc_amg_level_b(orc_amg *a, int l) { return a->lev[l].b; }
i64 orc_amg_level_nnz(orc_amg *a, int l) { return a->lev[l].rp[a->lev[l].n]; }
const i64 *orc_amg_level_rowptr(orc_amg *a, int l) { return a->lev[l].rp; }
const i32 *orc_amg_level_col(orc_amg *a, int l) { return a->lev[l].ci; }
const double *orc_amg_level_val(orc_amg *a, int l) { return a->lev[l].v; }
const double *orc_amg_level_dinv(orc_amg *a, int l) { return a->lev[l].dinv; }
const double *orc_amg_level_x(orc_amg *a, int l) { return a->lev[l].x; }
const double *orc_amg_level_rhs(orc_amg *a, int l) { return a->lev[l].rhs; }
const double *orc_amg_level_res(orc_amg *a, int l) { return a->lev[l].res; }
const i64 *orc_amg_pt_rowptr(orc_amg *a, int l) { return a->lev[l].pt_rp; }
const i32 *orc_amg_pt_col(orc_amg *a, int l) { return a->lev[l].pt_ci; }
const double *orc_amg_pt_val(orc_amg *a, int l) { return a->lev[l].pt_v; }

void orc_amg_free(orc_amg *a)
{
  if (!a) return;
  for (int l = 0; l < a->nlevels; l++) {
    orc_level *L = &a->lev[l];
    free(L->rp); free(L->ci); free(L->v); free(L->freed); free(L->dinv);
    free(L->p_rp); free(L->p_ci); free(L->p_v); free(L->pt_rp); free(L->pt_ci); free(L->pt_v);
    free(L->x); free(L->rhs); free(L->res); free(L->tmp);
  }
  free(a->lev); free(a->cdofs); free(a->cchol);
  free(a);
}
