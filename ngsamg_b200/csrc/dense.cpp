// dense.cpp -- host-side pseudo-inverse of one diagonal block of a level matrix (setup only; the smoother with ngs_amg_regularize_cmats).
// Restates the reference's CalcPseudoInverseTryNormal(Mat<N,N>&) chain exactly, thresholds included:
//   CallOnNonZeroDiagonalBlock<1>      src/base/utils/utils_denseLA.hpp:1237-1405   rows with diagonal <= max(1e-20, 1e-12 * max diagonal) are dropped
//   TryDirectInverse_simple            src/base/utils/utils_denseLA.cpp:458-555     Gauss-Jordan with column pivoting, gives up on a small pivot
//   CalcPseudoInverseWithTolNonZeroBlock  utils_denseLA.hpp:1474-1519               eigenvalues <= max(1e-12 * mean, 1e-20) are kernel
// tests/test_ref_pin.py compares ngsamg_b200_block_pinv with the reference's own code (oracle/_ref) on regular, rank-deficient and
// zero-row blocks: bit for bit where the direct inverse is taken, 1e-12 on the eigenvalue fall-back (LAPACK there, Jacobi rotations here).
#include <cmath>
#include <cstring>
#include <vector>

#include "../../include/ngsamg_b200.h"
#include "common.hpp"

namespace ngb {

namespace {
constexpr double REL_ZERO_TOL = 1e-12;   // RelZeroTol<double>, utils_denseLA.hpp:93-104
constexpr double ABS_ZERO_TOL = 1e-20;   // AbsZeroTol<double>, utils_denseLA.hpp:106-117

// in place; false (a untouched) when a pivot falls below max(ABS_ZERO_TOL * rest, REL_ZERO_TOL * max diagonal)
bool try_direct_inverse(int n, double *a)
{
  if (n == 0) return false;
  double eps = 0.0;
  for (int j = 0; j < n; j++) eps = std::max(eps, a[j * n + j]);
  eps *= REL_ZERO_TOL;
  std::vector<double> inv(a, a + n * n), hv(n);
  std::vector<int> p(n);
  for (int j = 0; j < n; j++) p[j] = j;
  for (int j = 0; j < n; j++) {
    double maxval = std::fabs(inv[j * n + j]);       // pivot search along row j
    int r = j;
    for (int i = j + 1; i < n; i++)
      if (std::fabs(inv[j * n + i]) > maxval) { r = i; maxval = std::fabs(inv[j * n + i]); }
    double rest = 0.0;
    for (int i = j + 1; i < n; i++) rest += std::fabs(inv[r * n + i]);
    if (maxval < std::max(ABS_ZERO_TOL * rest, eps)) return false;
    if (r > j) {
      for (int k = 0; k < n; k++) std::swap(inv[k * n + j], inv[k * n + r]);
      std::swap(p[j], p[r]);
    }
    const double hr = 1.0 / inv[j * n + j];
    for (int i = 0; i < n; i++) inv[j * n + i] = hr * inv[j * n + i];
    inv[j * n + j] = hr;
    for (int k = 0; k < n; k++) {
      if (k == j) continue;
      const double help = inv[k * n + j], h = help * hr;
      for (int i = 0; i < n; i++) inv[k * n + i] -= help * inv[j * n + i];
      inv[k * n + j] = -h;
    }
  }
  for (int i = 0; i < n; i++) {                       // undo the column exchanges
    for (int k = 0; k < n; k++) hv[p[k]] = inv[k * n + i];
    for (int k = 0; k < n; k++) inv[k * n + i] = hv[k];
  }
  std::memcpy(a, inv.data(), sizeof(double) * n * n);
  return true;
}

// symmetric eigen-decomposition by cyclic Jacobi rotations (the reference calls LAPACK): m <- sum over eigenvalues above the tolerance of v v^T / ev
void eig_pinv(int k, double *m)
{
  std::vector<double> a(m, m + k * k), V((size_t)k * k, 0.0);
  for (int i = 0; i < k; i++) V[i * k + i] = 1.0;
  for (int sweep = 0; sweep < 100; sweep++) {
    double off = 0;
    for (int i = 0; i < k; i++) for (int j = i + 1; j < k; j++) off += a[i * k + j] * a[i * k + j];
    if (off < 1e-300) break;
    for (int p = 0; p < k; p++) for (int q = p + 1; q < k; q++) {
      const double apq = a[p * k + q];
      if (std::fabs(apq) < 1e-300) continue;
      const double theta = (a[q * k + q] - a[p * k + p]) / (2.0 * apq);
      const double tt = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
      const double c = 1.0 / std::sqrt(tt * tt + 1.0), s = tt * c;
      for (int r = 0; r < k; r++) { double x = a[r * k + p], y = a[r * k + q]; a[r * k + p] = c * x - s * y; a[r * k + q] = s * x + c * y; }
      for (int r = 0; r < k; r++) { double x = a[p * k + r], y = a[q * k + r]; a[p * k + r] = c * x - s * y; a[q * k + r] = s * x + c * y; }
      for (int r = 0; r < k; r++) { double x = V[p * k + r], y = V[q * k + r]; V[p * k + r] = c * x - s * y; V[q * k + r] = s * x + c * y; }
    }
  }
  double tol = 0;
  for (int i = 0; i < k; i++) tol += a[i * k + i];
  tol = std::max(REL_ZERO_TOL * tol / k, ABS_ZERO_TOL);
  for (int i = 0; i < k; i++) for (int j = 0; j < k; j++) {
    double s = 0;
    for (int e = 0; e < k; e++) { const double ev = a[e * k + e]; if (ev > tol) s += V[e * k + i] * V[e * k + j] / ev; }
    m[i * k + j] = s;
  }
}
}  // namespace

void block_pinv(int n, double *m)
{
  if (n == 1) { m[0] = std::fabs(m[0]) > ABS_ZERO_TOL ? 1.0 / m[0] : 0.0; return; }   // scalar overload, utils_denseLA.hpp:1564-1569
  double maxd = 0.0;
  for (int i = 0; i < n; i++) maxd = std::max(maxd, m[i * n + i]);
  const double thresh = std::max(ABS_ZERO_TOL, REL_ZERO_TOL * maxd);
  std::vector<int> idx;
  for (int i = 0; i < n; i++) if (m[i * n + i] > thresh) idx.push_back(i);
  const int k = (int)idx.size();
  std::vector<double> sub((size_t)k * k);
  for (int i = 0; i < k; i++) for (int j = 0; j < k; j++) sub[i * k + j] = m[idx[i] * n + idx[j]];
  if (k > 0 && !try_direct_inverse(k, sub.data())) eig_pinv(k, sub.data());
  for (int i = 0; i < n * n; i++) m[i] = 0.0;
  for (int i = 0; i < k; i++) for (int j = 0; j < k; j++) m[idx[i] * n + idx[j]] = sub[i * k + j];
}

}  // namespace ngb

extern "C" int ngsamg_b200_block_pinv(int n, double *m)
{
  if (!m || n < 1 || n > 64) return 1;
  ngb::block_pinv(n, m);
  return 0;
}
