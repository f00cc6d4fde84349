// dense.cpp -- host-side pseudo-inverse of one diagonal block of a level matrix (setup only; the smoother with ngs_amg_regularize_cmats).
// Restates the reference's CalcPseudoInverseTryNormal(Mat<N,N>&) chain exactly, thresholds included:
//   CallOnNonZeroDiagonalBlock<1>      src/base/utils/utils_denseLA.hpp:1237-1405   rows with diagonal <= max(1e-20, 1e-12 * max diagonal) are dropped
//   TryDirectInverse_simple            src/base/utils/utils_denseLA.cpp:458-555     Gauss-Jordan with column pivoting, gives up on a small pivot
//   CalcPseudoInverseWithTolNonZeroBlock  utils_denseLA.hpp:1474-1519               eigenvalues <= max(1e-12 * mean, 1e-20) are kernel
// tests/test_ref_pin.py compares ngsamg_b200_block_pinv with the reference's own code (oracle/_ref) on regular, rank-deficient and
// zero-row blocks: bit for bit where the direct inverse is taken, 1e-12 on the eigenvalue fall-back (LAPACK there, Jacobi rotations here).
#include <algorithm>
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

// RegularizeMatrix of the elasticity preconditioners, local branch (src/elasticity/elasticity_pc_impl.hpp:711-763), applied to every
// diagonal block of the COARSEST matrix before it is inverted (amg_pc.cpp:861-862) when ngs_amg_regularize_cmats is set:
//   3D, 6x6 blocks: RegTM<0,6,6> (utils_denseLA.hpp:1198-1234) -- eigenvalues <= max(1e-15, 1e-12 * largest) count as zero; the smallest
//                   non-zero eigenvalue is added along every zero eigenvector; an entirely zero block becomes the identity
//   2D, 3x3 blocks: a rotational diagonal entry with |m(2,2)| < 1e-8 is set to 1
void block_regularize(int n, double *m, int dim)
{
  if (dim == 2) {
    if (n == 3 && std::fabs(m[2 * 3 + 2]) < 1e-8) m[2 * 3 + 2] = 1.0;
    return;
  }
  if (dim != 3 || n != 6) return;
  // eigen-decomposition (cyclic Jacobi rotations; LAPACK in the reference), eigenvalues ascending, V rows = eigenvectors
  std::vector<double> a(m, m + n * n), V((size_t)n * n, 0.0);
  for (int i = 0; i < n; i++) V[i * n + i] = 1.0;
  for (int sweep = 0; sweep < 100; sweep++) {
    double off = 0;
    for (int i = 0; i < n; i++) for (int j = i + 1; j < n; j++) off += a[i * n + j] * a[i * n + j];
    if (off < 1e-300) break;
    for (int p = 0; p < n; p++) for (int q = p + 1; q < n; q++) {
      const double apq = a[p * n + q];
      if (std::fabs(apq) < 1e-300) continue;
      const double theta = (a[q * n + q] - a[p * n + p]) / (2.0 * apq);
      const double tt = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
      const double c = 1.0 / std::sqrt(tt * tt + 1.0), s = tt * c;
      for (int r = 0; r < n; r++) { double x = a[r * n + p], y = a[r * n + q]; a[r * n + p] = c * x - s * y; a[r * n + q] = s * x + c * y; }
      for (int r = 0; r < n; r++) { double x = a[p * n + r], y = a[q * n + r]; a[p * n + r] = c * x - s * y; a[q * n + r] = s * x + c * y; }
      for (int r = 0; r < n; r++) { double x = V[p * n + r], y = V[q * n + r]; V[p * n + r] = c * x - s * y; V[q * n + r] = s * x + c * y; }
    }
  }
  std::vector<int> order(n);
  for (int i = 0; i < n; i++) order[i] = i;
  std::sort(order.begin(), order.end(), [&](int x, int y) { return a[x * n + x] < a[y * n + y]; });
  const double evmax = a[order[n - 1] * n + order[n - 1]];
  const double eps = std::max(1e-15, 1e-12 * evmax);
  double min_nzev = 0.0;
  int nzero = 0;
  for (int k = 0; k < n; k++) {
    const double ev = a[order[k] * n + order[k]];
    if (ev > eps) { min_nzev = ev; break; }
    nzero++;
  }
  if (nzero < n) {
    for (int l = 0; l < nzero; l++) {
      const double *v = &V[(size_t)order[l] * n];
      for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) m[i * n + j] += min_nzev * v[i] * v[j];
    }
  } else {
    for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) m[i * n + j] = (i == j) ? 1.0 : 0.0;
  }
}

}  // namespace ngb

extern "C" int ngsamg_b200_block_regularize(int n, double *m, int dim)
{
  if (!m || n < 1 || n > 64) return 1;
  ngb::block_regularize(n, m, dim);
  return 0;
}

extern "C" int ngsamg_b200_block_pinv(int n, double *m)
{
  if (!m || n < 1 || n > 64) return 1;
  ngb::block_pinv(n, m);
  return 0;
}
