// Stand-in for the dense part of ngbla / ngstd that the reference's pseudo-inverse code (utils_denseLA.hpp / .cpp) calls:
// FlatMatrix, LocalHeap-backed arrays, the symmetric eigenvalue routine.  TEST INFRASTRUCTURE ONLY (oracle/).
// LapackEigenValuesSymmetricLH is LAPACK's dsyev in NGSolve; here a cyclic Jacobi rotation solver with the same contract
// (eigenvalues ascending, eigenvectors as ROWS of evecs) -- the only path it feeds is the fall-back for singular blocks.
#pragma once
#include <list>

#include "ngs_standin.hpp"

namespace ngbla {
template <class T> class FlatMatrix;
template <class T> struct is_scalar_type { static constexpr bool value = false; };
template <> struct is_scalar_type<double> { static constexpr bool value = true; };
template <> struct is_scalar_type<float> { static constexpr bool value = true; };
template <> struct is_scalar_type<int> { static constexpr bool value = true; };
template <class TM> using TScal = typename mat_traits<TM>::TSCAL;

template <class T> struct OwnedMatrix {   // result of a matrix product, assignable to a FlatMatrix of the same shape
  size_t h, w;
  std::vector<T> d;
};
template <class T> struct TransView { const FlatMatrix<T> *m; };

template <class T> class FlatMatrix {
  size_t h = 0, w = 0;
  T *d = nullptr;

public:
  FlatMatrix() = default;
  FlatMatrix(size_t ah, size_t aw, T *p) : h(ah), w(aw), d(p) {}
  FlatMatrix(size_t ah, size_t aw, ngcore::LocalHeap &lh) : h(ah), w(aw), d((T *)ngcore::heap_alloc(lh, sizeof(T) * ah * aw)) {}
  template <int N> FlatMatrix(Mat<N, N, T> &m) : h(N), w(N), d(&m(0, 0)) {}
  FlatMatrix(const FlatMatrix &) = default;
  size_t Height() const { return h; }
  size_t Width() const { return w; }
  T &operator()(size_t i, size_t j) const { return d[i * w + j]; }
  T &operator()(size_t i) const { return d[i]; }      // linear (row-major) index
  void Assign(const FlatMatrix &o) { h = o.h; w = o.w; d = o.d; }
  FlatMatrix Rows(size_t a, size_t b) const { return FlatMatrix(b - a, w, d + a * w); }
  const FlatMatrix &operator=(const FlatMatrix &o) const { for (size_t i = 0; i < h * w; i++) d[i] = o.d[i]; return *this; }
  const FlatMatrix &operator=(const T &s) const { for (size_t i = 0; i < h * w; i++) d[i] = s; return *this; }
  const FlatMatrix &operator=(const OwnedMatrix<T> &o) const {
    if (o.h != h || o.w != w) throw ngcore::Exception("FlatMatrix: shape mismatch in assignment");
    for (size_t i = 0; i < h * w; i++) d[i] = o.d[i];
    return *this;
  }
};
template <class T> INLINE TransView<T> Trans(const FlatMatrix<T> &m) { return TransView<T>{&m}; }
// (A^T B)(i,j) = sum_k A(k,i) B(k,j), ascending k
template <class T> INLINE OwnedMatrix<T> operator*(const TransView<T> &a, const FlatMatrix<T> &b) {
  OwnedMatrix<T> r{a.m->Width(), b.Width(), std::vector<T>(a.m->Width() * b.Width())};
  for (size_t i = 0; i < r.h; i++)
    for (size_t j = 0; j < r.w; j++) {
      T s = 0;
      for (size_t k = 0; k < b.Height(); k++) s += (*a.m)(k, i) * b(k, j);
      r.d[i * r.w + j] = s;
    }
  return r;
}
INLINE void CalcInverse(double x, double &inv) { inv = 1.0 / x; }

template <int N, class T> class VectorMem {
  std::vector<T> d;

public:
  explicit VectorMem(size_t n) : d(n) {}
  T &operator()(size_t i) { return d[i]; }
};

// symmetric eigenvalue problem: evals ascending, row i of evecs = eigenvector of evals(i)
INLINE void LapackEigenValuesSymmetricLH(ngcore::LocalHeap &, FlatMatrix<double> M, FlatVector<double> evals, FlatMatrix<double> evecs) {
  const int n = int(M.Height());
  std::vector<double> a(n * n), V(n * n, 0.0);
  for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) a[i * n + j] = M(i, j);
  for (int i = 0; i < n; i++) V[i * n + i] = 1.0;
  for (int sweep = 0; sweep < 100; sweep++) {
    double off = 0;
    for (int i = 0; i < n; i++) for (int j = i + 1; j < n; j++) off += a[i * n + j] * a[i * n + j];
    if (off < 1e-300) break;
    for (int p = 0; p < n; p++)
      for (int q = p + 1; q < n; q++) {
        const double apq = a[p * n + q];
        if (std::fabs(apq) < 1e-300) continue;
        const double theta = (a[q * n + q] - a[p * n + p]) / (2.0 * apq);
        const double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
        const double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < n; k++) { const double x = a[k * n + p], y = a[k * n + q]; a[k * n + p] = c * x - s * y; a[k * n + q] = s * x + c * y; }
        for (int k = 0; k < n; k++) { const double x = a[p * n + k], y = a[q * n + k]; a[p * n + k] = c * x - s * y; a[q * n + k] = s * x + c * y; }
        for (int k = 0; k < n; k++) { const double x = V[p * n + k], y = V[q * n + k]; V[p * n + k] = c * x - s * y; V[q * n + k] = s * x + c * y; }
      }
  }
  std::vector<int> order(n);
  for (int i = 0; i < n; i++) order[i] = i;
  std::sort(order.begin(), order.end(), [&](int x, int y) { return a[x * n + x] < a[y * n + y]; });
  for (int i = 0; i < n; i++) {
    evals(i) = a[order[i] * n + order[i]];
    for (int k = 0; k < n; k++) evecs(i, k) = V[order[i] * n + k];
  }
}
// owning dense matrix / vector (RegTM keeps static work arrays of these)
template <class T> class Matrix : public FlatMatrix<T> {
  std::vector<T> store;

public:
  Matrix(size_t h, size_t w) : FlatMatrix<T>(), store(h * w) { this->Assign(FlatMatrix<T>(h, w, store.data())); }
};
template <class T> class Vector : public FlatVector<T> {
  std::vector<T> store;

public:
  explicit Vector(size_t n) : FlatVector<T>(), store(n) { this->AssignMemory(n, store.data()); }
};
INLINE void TimedLapackEigenValuesSymmetric(FlatMatrix<double> M, FlatVector<double> evals, FlatMatrix<double> evecs) {
  ngcore::LocalHeap lh(0, "eig");
  LapackEigenValuesSymmetricLH(lh, M, evals, evecs);
}
template <int N> INLINE void SetIdentity(Mat<N, N> &m) { for (int i = 0; i < N; i++) for (int j = 0; j < N; j++) m(i, j) = (i == j) ? 1.0 : 0.0; }
template <int H, int W> INLINE Mat<H, W> &operator*=(Mat<H, W> &m, double s) { for (int i = 0; i < H * W; i++) m.v[i] *= s; return m; }
}  // namespace ngbla

namespace ngstd {
using ngcore::ArrayMem;
}
