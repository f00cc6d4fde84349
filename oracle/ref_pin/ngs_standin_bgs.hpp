// ngs_standin_bgs.hpp -- the handful of NGSolve container / expression features that the reference's block Gauss-Seidel update
// routines use (BSmoother2<TM>::BSBlock::RichardsonUpdate / RichardsonUpdate_RES, loc_block_gssmoother_impl.hpp:244-268, 516-541),
// written for this repository.  TEST INFRASTRUCTURE ONLY.  Independent of ngs_standin.hpp so that the main pin library is untouched.
// Evaluation order of the dense expressions: every matrix-vector product is a plain row-wise sum over ascending column index
// (with omega = 1 -- the only value the smoother passes -- scaling is exact, so the grouping of `omega * M * v` does not matter).
#pragma once
#include <algorithm>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <ostream>
#include <string>
#include <tuple>
#include <vector>

#define INLINE inline
using std::ostream;
using std::string;
using std::tuple;

template <class T> struct T_Range {
  T first, next;
  struct It { T v; T operator*() const { return v; } It &operator++() { ++v; return *this; } bool operator!=(const It &o) const { return v != o.v; } };
  It begin() const { return It{first}; }
  It end() const { return It{next}; }
};
// NGSolve semantics: operator= copies the ELEMENTS (or fills with a scalar), Assign re-seats the view
template <class T> struct FlatArray {
  size_t n = 0;
  T *d = nullptr;
  FlatArray() = default;
  FlatArray(size_t an, T *p) : n(an), d(p) {}
  FlatArray(const FlatArray &) = default;
  size_t Size() const { return n; }
  T &operator[](size_t i) const { return d[i]; }
  T *Data() const { return d; }
  void Assign(const FlatArray &o) { n = o.n; d = o.d; }
  const FlatArray &operator=(const FlatArray &o) const { for (size_t i = 0; i < n; i++) d[i] = o.d[i]; return *this; }
  const FlatArray &operator=(const T &s) const { for (size_t i = 0; i < n; i++) d[i] = s; return *this; }
  T *begin() const { return d; }
  T *end() const { return d + n; }
};
INLINE T_Range<int> Range(int a, int b) { return T_Range<int>{a, b}; }
template <class T> INLINE T_Range<int> Range(const FlatArray<T> &a) { return T_Range<int>{0, (int)a.Size()}; }

template <int N> struct Vec {
  double v[N];
  Vec() { for (int i = 0; i < N; i++) v[i] = 0.0; }
  Vec(double s) { for (int i = 0; i < N; i++) v[i] = s; }
  double &operator()(int i) { return v[i]; }
  double operator()(int i) const { return v[i]; }
  Vec &operator+=(const Vec &o) { for (int i = 0; i < N; i++) v[i] += o.v[i]; return *this; }
  Vec &operator-=(const Vec &o) { for (int i = 0; i < N; i++) v[i] -= o.v[i]; return *this; }
};
template <int N> INLINE Vec<N> operator*(double s, const Vec<N> &a) { Vec<N> r; for (int i = 0; i < N; i++) r.v[i] = s * a.v[i]; return r; }
template <int H, int W> struct Mat {
  double v[H * W];
  Mat() { for (int i = 0; i < H * W; i++) v[i] = 0.0; }
  Mat(double s) { for (int i = 0; i < H * W; i++) v[i] = s; }
  double &operator()(int i, int j) { return v[i * W + j]; }
  double operator()(int i, int j) const { return v[i * W + j]; }
};
template <int H, int W> INLINE Vec<H> operator*(const Mat<H, W> &a, const Vec<W> &x) {
  Vec<H> r;
  for (int i = 0; i < H; i++) { double s = 0.0; for (int j = 0; j < W; j++) s += a(i, j) * x(j); r(i) = s; }
  return r;
}
template <int H, int W> INLINE Mat<W, H> Trans(const Mat<H, W> &a) { Mat<W, H> r; for (int i = 0; i < H; i++) for (int j = 0; j < W; j++) r(j, i) = a(i, j); return r; }
template <int H, int W> INLINE Mat<H, W> operator-(const Mat<H, W> &a, const Mat<H, W> &b) { Mat<H, W> r; for (int i = 0; i < H * W; i++) r.v[i] = a.v[i] - b.v[i]; return r; }
template <int H, int W> INLINE Mat<H, W> operator*(double s, const Mat<H, W> &a) { Mat<H, W> r; for (int i = 0; i < H * W; i++) r.v[i] = s * a.v[i]; return r; }
INLINE double Trans(double a) { return a; }

template <class TV> struct OwnedVec { std::vector<TV> d; };
template <class TV> struct ScaledVecView { double s; const TV *d; size_t n; };
template <class TV> class FlatVector;
template <class TV> struct IndirectVec {
  TV *d;
  FlatArray<int> idx;
  size_t Size() const { return idx.Size(); }
  TV &operator()(size_t i) const { return d[idx[i]]; }
  void operator+=(const ScaledVecView<TV> &e) const { for (size_t i = 0; i < idx.Size(); i++) d[idx[i]] += e.s * e.d[i]; }
  void operator+=(const FlatVector<TV> &e) const;
  void operator-=(const OwnedVec<TV> &e) const { for (size_t i = 0; i < idx.Size(); i++) d[idx[i]] -= e.d[i]; }
};
template <class TV> class FlatVector {
  size_t n = 0;
  TV *d = nullptr;

public:
  FlatVector() = default;
  FlatVector(size_t an, TV *p) : n(an), d(p) {}
  size_t Size() const { return n; }
  TV *Data() const { return d; }
  TV &operator()(size_t i) const { return d[i]; }
  TV &operator[](size_t i) const { return d[i]; }
  IndirectVec<TV> operator()(FlatArray<int> ind) const { return IndirectVec<TV>{d, ind}; }
  const FlatVector &operator=(const OwnedVec<TV> &o) const { for (size_t i = 0; i < n; i++) d[i] = o.d[i]; return *this; }
  const FlatVector &operator-=(const OwnedVec<TV> &o) const { for (size_t i = 0; i < n; i++) d[i] -= o.d[i]; return *this; }
};
template <class TV> void IndirectVec<TV>::operator+=(const FlatVector<TV> &e) const { for (size_t i = 0; i < idx.Size(); i++) d[idx[i]] += e(i); }
template <class TV> INLINE ScaledVecView<TV> operator*(double s, const FlatVector<TV> &v) { return ScaledVecView<TV>{s, v.Data(), v.Size()}; }

template <class TM> class FlatMatrix {
  size_t h = 0, w = 0;
  TM *d = nullptr;

public:
  FlatMatrix() = default;
  FlatMatrix(size_t ah, size_t aw, TM *p) : h(ah), w(aw), d(p) {}
  size_t Height() const { return h; }
  size_t Width() const { return w; }
  TM &operator()(size_t i, size_t j) const { return d[i * w + j]; }
  void AssignMemory(size_t ah, size_t aw, TM *p) { h = ah; w = aw; d = p; }
  const FlatMatrix &operator=(const FlatMatrix &o) const { for (size_t i = 0; i < h * w; i++) d[i] = o.d[i]; return *this; }
  const FlatMatrix &operator=(double s) const { for (size_t i = 0; i < h * w; i++) d[i] = TM(s); return *this; }
};
template <class TM> struct ScaledMatView { double s; const FlatMatrix<TM> *m; };
template <class TM> INLINE ScaledMatView<TM> operator*(double s, const FlatMatrix<TM> &m) { return ScaledMatView<TM>{s, &m}; }
// y_i = sum_j A(i,j) x_j, ascending j, starting from A(i,0) x_0
template <class TM, class TV, class X> INLINE OwnedVec<TV> matvec(double s, const FlatMatrix<TM> &A, const X &x) {
  OwnedVec<TV> r;
  r.d.resize(A.Height());
  for (size_t i = 0; i < A.Height(); i++) {
    TV acc = A(i, 0) * x(0);
    for (size_t j = 1; j < A.Width(); j++) acc += A(i, j) * x(j);
    r.d[i] = s * acc;
  }
  return r;
}
template <class TM, class TV> INLINE OwnedVec<TV> operator*(const FlatMatrix<TM> &A, const FlatVector<TV> &x) { return matvec<TM, TV>(1.0, A, x); }
template <class TM, class TV> INLINE OwnedVec<TV> operator*(const FlatMatrix<TM> &A, const IndirectVec<TV> &x) { return matvec<TM, TV>(1.0, A, x); }
template <class TM, class TV> INLINE OwnedVec<TV> operator*(const ScaledMatView<TM> &A, const FlatVector<TV> &x) { return matvec<TM, TV>(A.s, *A.m, x); }

// VectorMem<N, T>: a vector with its own storage; Range(a, b) gives a view
template <int N, class TV> class VectorMem {
  std::vector<TV> d;

public:
  explicit VectorMem(size_t n) : d(n) {}
  FlatVector<TV> Range(size_t a, size_t b) { return FlatVector<TV>(b - a, d.data() + a); }
};
template <class T> struct Array {
  std::vector<T> d;
  size_t Size() const { return d.size(); }
  T &operator[](size_t i) { return d[i]; }
  const T &operator[](size_t i) const { return d[i]; }
};
INLINE T_Range<int> Range(int n) { return T_Range<int>{0, n}; }
INLINE T_Range<int> Range(size_t a, size_t b) { return T_Range<int>{(int)a, (int)b}; }

// BaseVector / BaseMatrix: just enough for  x.FV<TV>()  and  res = b - A * x  (evaluated as a block-CSR SpMV; this is glue, not pinned)
struct BgsCsr { int64_t n; int b; const int64_t *rp; const int32_t *ci; const double *v; };
struct BaseVector;
struct BgsMatVec { const BgsCsr *A; const BaseVector *x; };
struct BgsResid { const BaseVector *b; BgsMatVec e; };
struct BaseVector {
  double *d;
  size_t n;   // scalars
  template <class TV> FlatVector<TV> FV() const { return FlatVector<TV>(n * sizeof(double) / sizeof(TV), reinterpret_cast<TV *>(d)); }
  BaseVector &operator=(const BgsResid &r) {
    const BgsCsr &A = *r.e.A;
    const int b = A.b;
    for (int64_t i = 0; i < A.n; i++)
      for (int p = 0; p < b; p++) {
        double s = r.b->d[i * b + p];
        for (int64_t e = A.rp[i]; e < A.rp[i + 1]; e++)
          for (int q = 0; q < b; q++) s -= A.v[e * b * b + p * b + q] * r.e.x->d[(int64_t)A.ci[e] * b + q];
        d[i * b + p] = s;
      }
    return *this;
  }
};
INLINE BgsMatVec operator*(const BgsCsr &A, const BaseVector &x) { return BgsMatVec{&A, &x}; }
INLINE BgsResid operator-(const BaseVector &b, const BgsMatVec &e) { return BgsResid{&b, e}; }

// what BSBlock::SetFromSPMat needs: row access of a sparse matrix with TM entries, find_in_sorted_array, QuickSort, LocalHeap / HeapReset,
// CalcInverse of a matrix of TM blocks (NGSolve's own routine is third-party: here a Gauss-Jordan on the scalar matrix, partial pivoting)
template <class TM> struct SparseMatrixTM {
  int64_t n;
  const int64_t *rp;
  const int32_t *ci;
  std::vector<int> cols;      // int copy of the column indices
  std::vector<TM> vals;
  FlatArray<int> GetRowIndices(int i) const { return FlatArray<int>(rp[i + 1] - rp[i], const_cast<int *>(cols.data()) + rp[i]); }
  FlatVector<TM> GetRowValues(int i) const { return FlatVector<TM>(rp[i + 1] - rp[i], const_cast<TM *>(vals.data()) + rp[i]); }
};
template <class T> INLINE T_Range<int> Range(const FlatVector<T> &a) { return T_Range<int>{0, (int)a.Size()}; }
INLINE int find_in_sorted_array(int x, FlatArray<int> a) {
  size_t lo = 0, hi = a.Size();
  while (lo < hi) { const size_t mid = (lo + hi) / 2; if (a[mid] < x) lo = mid + 1; else hi = mid; }
  return (lo < a.Size() && a[lo] == x) ? (int)lo : -1;
}
INLINE void QuickSort(FlatArray<int> a) { std::sort(a.begin(), a.end()); }     // distinct keys: every correct sort gives the same array
struct LocalHeap {};
struct HeapReset { explicit HeapReset(LocalHeap &) {} };
template <class TM> struct bgs_block_dim { static constexpr int value = 1; };
template <int N> struct bgs_block_dim<Mat<N, N>> { static constexpr int value = N; };
INLINE double &bgs_entry(double &m, int, int) { return m; }
template <int N> INLINE double &bgs_entry(Mat<N, N> &m, int p, int q) { return m(p, q); }
template <class TM> INLINE void CalcInverse(const FlatMatrix<TM> &M) {
  constexpr int B = bgs_block_dim<TM>::value;
  const int m = (int)M.Height(), N = m * B;
  std::vector<double> a((size_t)N * N), inv((size_t)N * N, 0.0);
  for (int i = 0; i < m; i++) for (int j = 0; j < m; j++) for (int p = 0; p < B; p++) for (int q = 0; q < B; q++)
    a[(size_t)(i * B + p) * N + j * B + q] = bgs_entry(M(i, j), p, q);
  for (int i = 0; i < N; i++) inv[(size_t)i * N + i] = 1.0;
  for (int c = 0; c < N; c++) {
    int piv = c;
    for (int r = c + 1; r < N; r++) if (std::fabs(a[(size_t)r * N + c]) > std::fabs(a[(size_t)piv * N + c])) piv = r;
    if (a[(size_t)piv * N + c] == 0.0) throw 1;
    if (piv != c) for (int q = 0; q < N; q++) { std::swap(a[(size_t)c * N + q], a[(size_t)piv * N + q]); std::swap(inv[(size_t)c * N + q], inv[(size_t)piv * N + q]); }
    const double f = 1.0 / a[(size_t)c * N + c];
    for (int q = 0; q < N; q++) { a[(size_t)c * N + q] *= f; inv[(size_t)c * N + q] *= f; }
    for (int r = 0; r < N; r++) {
      if (r == c) continue;
      const double g = a[(size_t)r * N + c];
      if (g == 0.0) continue;
      for (int q = 0; q < N; q++) { a[(size_t)r * N + q] -= g * a[(size_t)c * N + q]; inv[(size_t)r * N + q] -= g * inv[(size_t)c * N + q]; }
    }
  }
  for (int i = 0; i < m; i++) for (int j = 0; j < m; j++) for (int p = 0; p < B; p++) for (int q = 0; q < B; q++)
    bgs_entry(M(i, j), p, q) = inv[(size_t)(i * B + p) * N + j * B + q];
}
template <class TM> INLINE void CalcPseudoInverseTryNormal(const FlatMatrix<TM> &, LocalHeap &) { throw 2; }   // pinv blocks are not part of this pin

template <class TM> struct bgs_vec_of { using type = double; };
template <int N> struct bgs_vec_of<Mat<N, N>> { using type = Vec<N>; };
