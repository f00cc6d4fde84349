// Stand-in for the part of NGSolve's public API (ngcore / ngbla / ngla) that the hot-path functions of the reference call.
//
// TEST INFRASTRUCTURE ONLY (oracle/).  NGSolve is not in this image and not in /root/reference; the reference's functions
// (oracle/_ref/frag/*.inc, cut from /root/reference at build time) are compiled VERBATIM against this header so that the
// oracle's restatement can be compared with the reference's own code.  Everything here is written for this repository from
// the documented behaviour of the NGSolve classes; it is deliberately minimal and strictly sequential.  What this header
// decides itself (and the pin therefore does NOT cover) is listed in oracle/ref_pin/README.md: the summation order inside
// SparseMatrix::RowTimesVector / AddRowTransToVector / MultAdd, the Mat*Mat / Mat*Vec evaluation order, CalcInverse,
// MergeArrays (sorted unique union) and BubbleSort.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <functional>
#include <initializer_list>
#include <iostream>
#include <limits>
#include <memory>
#include <numeric>
#include <stdexcept>
#include <string>
#include <tuple>
#include <type_traits>
#include <vector>

#define INLINE inline
#define LAMBDA_INLINE

namespace ngcore {
using std::shared_ptr;
using std::make_shared;
using std::string;
using std::to_string;

struct Exception : std::runtime_error {
  explicit Exception(const std::string &s) : std::runtime_error(s) {}
};

template <class T> INLINE T min2(T a, T b) { return a < b ? a : b; }
template <class T> INLINE T max2(T a, T b) { return a > b ? a : b; }

// ---- timers: no-ops ------------------------------------------------------------------------------------------
struct TTracing {};
struct TTiming {};
template <class A = TTracing, class B = TTiming> struct Timer {
  Timer(const std::string &) {}
  void Start() {}
  void Stop() {}
  double GetTime() const { return 0.0; }
};
struct RegionTimer {
  template <class T> explicit RegionTimer(T &) {}
};

// ---- ranges --------------------------------------------------------------------------------------------------
template <class T> struct T_Range {
  T first, next;
  T_Range(T a, T b) : first(a), next(b) {}
  explicit T_Range(T n) : first(T(0)), next(n) {}
  struct It {
    T v;
    T operator*() const { return v; }
    It &operator++() { ++v; return *this; }
    bool operator!=(const It &o) const { return v != o.v; }
  };
  It begin() const { return It{first}; }
  It end() const { return It{next > first ? next : first}; }
  T First() const { return first; }
  T Next() const { return next; }
  size_t Size() const { return next > first ? size_t(next - first) : 0; }
};
using IntRange = T_Range<size_t>;

template <class T, class = std::enable_if_t<std::is_integral<T>::value>> INLINE T_Range<T> Range(T n) { return T_Range<T>(T(0), n); }
template <class T, class = std::enable_if_t<std::is_integral<T>::value>> INLINE T_Range<T> Range(T a, T b) { return T_Range<T>(a, b); }

// ---- arrays --------------------------------------------------------------------------------------------------
struct LocalHeap;
template <class T> class FlatArray {
protected:
  size_t size = 0;
  T *data = nullptr;

public:
  FlatArray() = default;
  FlatArray(size_t n, T *p) : size(n), data(p) {}
  FlatArray(size_t n, LocalHeap &lh);
  FlatArray Range(size_t a, size_t b) const { return FlatArray(b - a, data + a); }
  size_t Size() const { return size; }
  T *Data() const { return data; }
  T *Addr(size_t i) const { return data + i; }
  T &operator[](size_t i) const { return data[i]; }
  T &Last() const { return data[size - 1]; }
  T *begin() const { return data; }
  T *end() const { return data + size; }
  IntRange Range() const { return IntRange(0, size); }
  const FlatArray &operator=(const T &v) const {
    for (size_t i = 0; i < size; i++) data[i] = v;
    return *this;
  }
  // assignment copies the ELEMENTS (a FlatArray is a view); Assign re-seats the view
  FlatArray(const FlatArray &) = default;
  const FlatArray &operator=(const FlatArray &o) const {
    for (size_t i = 0; i < size; i++) data[i] = o.data[i];
    return *this;
  }
  void Assign(const FlatArray &o) { size = o.size; data = o.data; }
};
template <class T> INLINE IntRange Range(const FlatArray<T> &a) { return IntRange(0, a.Size()); }

template <class T> class Array : public FlatArray<T> {
  std::vector<T> store;  // T = bool is never used here
  void sync() { this->size = store.size(); this->data = store.empty() ? nullptr : store.data(); }

public:
  Array() = default;
  explicit Array(size_t n) : store(n) { sync(); }
  Array(std::initializer_list<T> l) : store(l) { sync(); }
  Array(const Array &o) : FlatArray<T>(), store(o.store) { sync(); }
  Array(Array &&o) noexcept : FlatArray<T>(), store(std::move(o.store)) { sync(); o.sync(); }
  Array &operator=(const Array &o) { store = o.store; sync(); return *this; }
  Array &operator=(Array &&o) noexcept { store = std::move(o.store); sync(); o.sync(); return *this; }
  Array &operator=(const T &v) { for (auto &e : store) e = v; return *this; }
  void SetSize(size_t n) { store.resize(n); sync(); }
  void SetSize0() { store.clear(); sync(); }
  void Append(const T &v) { store.push_back(v); sync(); }
};
template <int N, class T = int> struct IVec {
  T v[N];
  IVec() {}
  IVec(std::initializer_list<T> l) { int i = 0; for (auto e : l) v[i++] = e; }
  T &operator[](int i) { return v[i]; }
  const T &operator[](int i) const { return v[i]; }
};
template <class T, int N> class ArrayMem : public Array<T> {
public:
  ArrayMem() = default;
  explicit ArrayMem(size_t n) : Array<T>(n) {}
};

class BitArray {
  std::vector<unsigned char> bits;

public:
  explicit BitArray(size_t n) : bits(n, 0) {}
  size_t Size() const { return bits.size(); }
  bool Test(size_t i) const { return bits[i] != 0; }
  void SetBit(size_t i) { bits[i] = 1; }
  void Clear(size_t i) { bits[i] = 0; }
  void Clear() { for (auto &b : bits) b = 0; }
  void And(const BitArray &o) { for (size_t i = 0; i < bits.size(); i++) bits[i] = bits[i] && o.bits[i]; }
  size_t NumSet() const { size_t c = 0; for (auto b : bits) c += b; return c; }
};

// ---- task manager: sequential, but split into several ranges like a task pool would --------------------------
struct TasksPerThread { int n; explicit TasksPerThread(int a) : n(a) {} };
template <class F> INLINE void ParallelForRange(IntRange r, F f, TasksPerThread tpt = TasksPerThread(1)) {
  const size_t n = r.Size(), parts = std::max<size_t>(1, std::min<size_t>(n, size_t(tpt.n) * 3));
  for (size_t p = 0; p < parts; p++) {
    IntRange sub(r.First() + n * p / parts, r.First() + n * (p + 1) / parts);
    if (sub.Size()) f(sub);
  }
}
template <class F> INLINE void ParallelForRange(size_t n, F f, TasksPerThread tpt = TasksPerThread(1)) { ParallelForRange(IntRange(0, n), f, tpt); }

struct HeapStore;
struct LocalHeap {   // hands out memory that lives until a HeapReset taken earlier goes out of scope, or the heap (and its splits) dies
  std::shared_ptr<HeapStore> store;
  LocalHeap(size_t, const char *);
  LocalHeap Split() { return *this; }
};
struct HeapStore { std::vector<std::unique_ptr<char[]>> chunks; };
struct HeapReset {   // everything allocated after construction is released on destruction
  LocalHeap &lh;
  size_t mark;
  explicit HeapReset(LocalHeap &alh) : lh(alh), mark(alh.store->chunks.size()) {}
  ~HeapReset() { lh.store->chunks.resize(mark); }
};
inline LocalHeap::LocalHeap(size_t, const char *) : store(std::make_shared<HeapStore>()) {}
INLINE void *heap_alloc(LocalHeap &lh, size_t bytes) {
  lh.store->chunks.emplace_back(new char[bytes ? bytes : 1]());
  return lh.store->chunks.back().get();
}
template <class T> FlatArray<T>::FlatArray(size_t n, LocalHeap &lh) : size(n), data((T *)heap_alloc(lh, sizeof(T) * n)) {}
}  // namespace ngcore

namespace ngbla {
using namespace ngcore;

// ---- small dense blocks ----------------------------------------------------------------------------------------
template <int H, int W, class T = double> struct Mat {
  enum { HEIGHT = H, WIDTH = W };
  T v[H * W];
  Mat() {}
  Mat(T s) { for (int i = 0; i < H * W; i++) v[i] = s; }
  static constexpr int Height() { return H; }
  static constexpr int Width() { return W; }
  T &operator()(int i, int j) { return v[i * W + j]; }
  const T &operator()(int i, int j) const { return v[i * W + j]; }
  Mat &operator=(T s) { for (int i = 0; i < H * W; i++) v[i] = s; return *this; }
  Mat &operator+=(const Mat &o) { for (int i = 0; i < H * W; i++) v[i] += o.v[i]; return *this; }
  Mat &operator-=(const Mat &o) { for (int i = 0; i < H * W; i++) v[i] -= o.v[i]; return *this; }
  Mat operator-() const { Mat r; for (int i = 0; i < H * W; i++) r.v[i] = -v[i]; return r; }
};
template <int N, class T = double> struct Vec {
  enum { HEIGHT = N, WIDTH = 1 };
  T v[N];
  Vec() {}
  Vec(T s) { for (int i = 0; i < N; i++) v[i] = s; }
  T &operator()(int i) { return v[i]; }
  const T &operator()(int i) const { return v[i]; }
  T &operator[](int i) { return v[i]; }
  const T &operator[](int i) const { return v[i]; }
  Vec &operator=(T s) { for (int i = 0; i < N; i++) v[i] = s; return *this; }
  Vec &operator+=(const Vec &o) { for (int i = 0; i < N; i++) v[i] += o.v[i]; return *this; }
  Vec &operator-=(const Vec &o) { for (int i = 0; i < N; i++) v[i] -= o.v[i]; return *this; }
  Vec operator-() const { Vec r; for (int i = 0; i < N; i++) r.v[i] = -v[i]; return r; }
};
template <int N> INLINE Vec<N> operator+(const Vec<N> &a, const Vec<N> &b) { Vec<N> r; for (int i = 0; i < N; i++) r.v[i] = a.v[i] + b.v[i]; return r; }
template <int N> INLINE Vec<N> operator-(const Vec<N> &a, const Vec<N> &b) { Vec<N> r; for (int i = 0; i < N; i++) r.v[i] = a.v[i] - b.v[i]; return r; }
template <int N> INLINE Vec<N> operator*(double s, const Vec<N> &a) { Vec<N> r; for (int i = 0; i < N; i++) r.v[i] = s * a.v[i]; return r; }
template <int H, int W> INLINE Mat<H, W> operator*(double s, const Mat<H, W> &a) { Mat<H, W> r; for (int i = 0; i < H * W; i++) r.v[i] = s * a.v[i]; return r; }
// entry (i,j) of a product = sum over k in increasing k, first term assigned (no zero start)
template <int H, int K, int W> INLINE Mat<H, W> operator*(const Mat<H, K> &a, const Mat<K, W> &b) {
  Mat<H, W> r;
  for (int i = 0; i < H; i++)
    for (int j = 0; j < W; j++) {
      double s = a(i, 0) * b(0, j);
      for (int k = 1; k < K; k++) s += a(i, k) * b(k, j);
      r(i, j) = s;
    }
  return r;
}
template <int H, int W> INLINE Vec<H> operator*(const Mat<H, W> &a, const Vec<W> &x) {
  Vec<H> r;
  for (int i = 0; i < H; i++) {
    double s = a(i, 0) * x(0);
    for (int k = 1; k < W; k++) s += a(i, k) * x(k);
    r(i) = s;
  }
  return r;
}
template <int H, int W> INLINE Mat<W, H> Trans(const Mat<H, W> &a) {
  Mat<W, H> r;
  for (int i = 0; i < H; i++) for (int j = 0; j < W; j++) r(j, i) = a(i, j);
  return r;
}
INLINE double Trans(double a) { return a; }

template <class T> struct mat_traits { typedef double TSCAL; typedef Vec<T::HEIGHT> TV_COL; };
template <> struct mat_traits<double> { typedef double TSCAL; typedef double TV_COL; };
template <int I> struct IC { static constexpr int value = I; constexpr operator int() const { return I; } };
template <int N, int I = 0, class F> INLINE void Iterate(F f) { if constexpr (I < N) { f(IC<I>()); Iterate<N, I + 1>(f); } }
template <class T> constexpr int Height() { return T::HEIGHT; }
template <> constexpr int Height<double>() { return 1; }
template <class T> constexpr int Width() { return T::WIDTH; }
template <> constexpr int Width<double>() { return 1; }

// inverse of a diagonal block.  NGSolve: closed forms for N <= 3, pivoted elimination above; here: Gauss-Jordan with partial
// pivoting for every N > 1 (results agree to rounding, tests compare block inverses with a tolerance, scalars exactly)
INLINE void CalcInverse(double &d) { d = 1.0 / d; }
template <int N> INLINE void CalcInverse(Mat<N, N> &m) {
  double a[N][2 * N];
  for (int i = 0; i < N; i++) for (int j = 0; j < N; j++) { a[i][j] = m(i, j); a[i][N + j] = (i == j) ? 1.0 : 0.0; }
  for (int c = 0; c < N; c++) {
    int p = c;
    for (int r = c + 1; r < N; r++) if (std::fabs(a[r][c]) > std::fabs(a[p][c])) p = r;
    if (a[p][c] == 0.0) throw Exception("CalcInverse: singular block");
    if (p != c) for (int j = 0; j < 2 * N; j++) std::swap(a[p][j], a[c][j]);
    const double piv = 1.0 / a[c][c];
    for (int j = 0; j < 2 * N; j++) a[c][j] *= piv;
    for (int r = 0; r < N; r++) if (r != c) { const double f = a[r][c]; if (f != 0.0) for (int j = 0; j < 2 * N; j++) a[r][j] -= f * a[c][j]; }
  }
  for (int i = 0; i < N; i++) for (int j = 0; j < N; j++) m(i, j) = a[i][N + j];
}

template <class T> class FlatVector {
  size_t n = 0;
  T *d = nullptr;

public:
  FlatVector() = default;
  FlatVector(size_t an, T *p) : n(an), d(p) {}
  FlatVector(size_t an, ngcore::LocalHeap &lh) : n(an), d((T *)ngcore::heap_alloc(lh, sizeof(T) * an)) {}
  FlatVector(const FlatVector &) = default;
  void AssignMemory(size_t an, T *p) { n = an; d = p; }
  size_t Size() const { return n; }
  T *Data() const { return d; }
  T &operator()(size_t i) const { return d[i]; }
  T &operator[](size_t i) const { return d[i]; }
  T *begin() const { return d; }
  T *end() const { return d + n; }
  // assignment copies the ELEMENTS (a FlatVector is a view)
  const FlatVector &operator=(const FlatVector &o) const { for (size_t i = 0; i < n; i++) d[i] = o.d[i]; return *this; }
  const FlatVector &operator=(const T &s) const { for (size_t i = 0; i < n; i++) d[i] = s; return *this; }
};
}  // namespace ngbla

namespace ngla {
using namespace ngbla;

enum PARALLEL_STATUS { DISTRIBUTED, CUMULATED, NOT_PARALLEL };

class BaseMatrix;
class BaseVector;
// the few vector expressions the smoothers write:  A * x,  s * A * x,  b - A * x,  s * x.  Evaluated like NGSolve does:
// `v += s*A*x` is A.MultAdd(s, x, v);  `v = b - A*x` is v = b followed by A.MultAdd(-1, x, v)
struct MatVecExpr { const BaseMatrix *m; const BaseVector *x; double s = 1.0; };
struct ScaledMatExpr { double s; const BaseMatrix *m; };
struct VecMinusMatVecExpr { const BaseVector *b; MatVecExpr e; };
struct ScaledVecExpr { double s; const BaseVector *x; };

// vector of n entries of `es` doubles each (es = block size); strictly single-rank: Cumulate/Distribute only flip the status
class BaseVector {
  std::vector<double> store;
  size_t es = 1;
  mutable PARALLEL_STATUS stat = NOT_PARALLEL;

public:
  BaseVector(size_t n, size_t entry) : store(n * entry, 0.0), es(entry) {}
  virtual ~BaseVector() = default;
  size_t Size() const { return store.size() / es; }
  FlatVector<double> FVDouble() const { return FlatVector<double>(store.size(), const_cast<double *>(store.data())); }
  template <class TV> FlatVector<TV> FV() const {
    return FlatVector<TV>(store.size() * sizeof(double) / sizeof(TV), reinterpret_cast<TV *>(const_cast<double *>(store.data())));
  }
  // single rank: only the status flips.  Vectors of the multi-rank harness are marked `parallel`: a status change that would need
  // communication (or zeroing of ghost entries) is an error there -- the hybrid smoother must be handed the right status
  bool parallel = false;
  void Cumulate() const {
    if (stat == DISTRIBUTED) { if (parallel) throw Exception("BaseVector::Cumulate would need communication"); stat = CUMULATED; }
  }
  void Distribute() const {
    if (stat == CUMULATED) { if (parallel) throw Exception("BaseVector::Distribute on a cumulated parallel vector"); stat = DISTRIBUTED; }
  }
  BaseVector *GetLocalVector() const { return const_cast<BaseVector *>(this); }
  void *Memory() const { return const_cast<double *>(store.data()); }
  BaseVector &operator+=(const BaseVector &o) { for (size_t i = 0; i < store.size(); i++) store[i] += o.store[i]; return *this; }
  void SetParallelStatus(PARALLEL_STATUS s) const { stat = s; }
  PARALLEL_STATUS GetParallelStatus() const { return stat; }
  BaseVector &operator=(double s) { for (auto &e : store) e = s; return *this; }
  BaseVector &operator=(const BaseVector &o) { store = o.store; stat = o.stat; return *this; }
  BaseVector(const BaseVector &) = default;
  BaseVector &operator-=(const MatVecExpr &e);
  BaseVector &operator+=(const MatVecExpr &e);
  BaseVector &operator=(const VecMinusMatVecExpr &e);
  BaseVector &operator+=(const ScaledVecExpr &e) { for (size_t i = 0; i < store.size(); i++) store[i] += e.s * e.x->store[i]; return *this; }
};
INLINE ScaledVecExpr operator*(double s, const BaseVector &x) { return ScaledVecExpr{s, &x}; }

class BaseMatrix {
public:
  virtual ~BaseMatrix() = default;
  virtual int VHeight() const = 0;
  virtual int VWidth() const = 0;
  size_t Height() const { return VHeight(); }
  size_t Width() const { return VWidth(); }
  virtual void Mult(const BaseVector &x, BaseVector &y) const { y = 0.0; MultAdd(1.0, x, y); }
  virtual size_t ScalNZE() const { return 0; }      // stored entries * entry size (what the reference's GetScalNZE reports for sparse matrices)
  virtual void MultAdd(double s, const BaseVector &x, BaseVector &y) const = 0;
  MatVecExpr operator*(const BaseVector &x) const { return MatVecExpr{this, &x, 1.0}; }
};
INLINE ScaledMatExpr operator*(double s, const BaseMatrix &m) { return ScaledMatExpr{s, &m}; }
INLINE MatVecExpr operator*(const ScaledMatExpr &sm, const BaseVector &x) { return MatVecExpr{sm.m, &x, sm.s}; }
INLINE VecMinusMatVecExpr operator-(const BaseVector &b, const MatVecExpr &e) { return VecMinusMatVecExpr{&b, e}; }
INLINE BaseVector &BaseVector::operator-=(const MatVecExpr &e) { e.m->MultAdd(-e.s, *e.x, *this); return *this; }
INLINE BaseVector &BaseVector::operator+=(const MatVecExpr &e) { e.m->MultAdd(e.s, *e.x, *this); return *this; }
INLINE BaseVector &BaseVector::operator=(const VecMinusMatVecExpr &e) {
  if (this != e.b) *this = *e.b;
  e.e.m->MultAdd(-e.e.s, *e.e.x, *this);
  return *this;
}

// y(i) += s * d(i) * x(i)
template <class TM> class DiagonalMatrix : public BaseMatrix {
  std::vector<TM> d;

public:
  explicit DiagonalMatrix(size_t n) : d(n) {}
  int VHeight() const override { return int(d.size()); }
  int VWidth() const override { return int(d.size()); }
  TM &operator()(size_t i) { return d[i]; }
  const TM &operator()(size_t i) const { return d[i]; }
  void MultAdd(double s, const BaseVector &x, BaseVector &y) const override;
};

template <class TM> struct vec_of_rows { typedef Vec<TM::HEIGHT> type; };
template <> struct vec_of_rows<double> { typedef double type; };
template <class TM> struct vec_of_cols { typedef Vec<TM::WIDTH> type; };
template <> struct vec_of_cols<double> { typedef double type; };

template <class TM> void DiagonalMatrix<TM>::MultAdd(double s, const BaseVector &x, BaseVector &y) const {
  auto fx = x.FV<typename vec_of_cols<TM>::type>();
  auto fy = y.FV<typename vec_of_rows<TM>::type>();
  for (size_t i = 0; i < d.size(); i++) fy(i) += (s * d[i]) * fx(i);
}

// CSR with block entries, columns ascending inside a row
template <class TM> class SparseMatrix : public BaseMatrix {
  size_t h = 0, w = 0;
  std::vector<size_t> firsti;
  std::vector<int> colnr;
  std::vector<TM> data;

public:
  typedef TM TENTRY;
  typedef typename vec_of_cols<TM>::type TVX;
  typedef typename vec_of_rows<TM>::type TVY;
  SparseMatrix(const FlatArray<int> &elsperrow, size_t awidth) : h(elsperrow.Size()), w(awidth), firsti(elsperrow.Size() + 1, 0) {
    for (size_t i = 0; i < h; i++) firsti[i + 1] = firsti[i] + size_t(elsperrow[i]);
    colnr.assign(firsti[h], 0);
    data.resize(firsti[h]);
  }
  explicit SparseMatrix(const FlatArray<int> &elsperrow) : SparseMatrix(elsperrow, elsperrow.Size()) {}
  int VHeight() const override { return int(h); }
  int VWidth() const override { return int(w); }
  size_t NZE() const { return firsti[h]; }
  size_t ScalNZE() const override { return firsti[h] * (sizeof(TM) / sizeof(double)); }
  FlatArray<int> GetRowIndices(size_t i) const { return FlatArray<int>(firsti[i + 1] - firsti[i], const_cast<int *>(colnr.data()) + firsti[i]); }
  FlatVector<TM> GetRowValues(size_t i) const { return FlatVector<TM>(firsti[i + 1] - firsti[i], const_cast<TM *>(data.data()) + firsti[i]); }
  size_t First(size_t i) const { return firsti[i]; }
  size_t GetPosition(size_t i, int c) const {
    auto b = colnr.begin() + firsti[i], e = colnr.begin() + firsti[i + 1];
    auto it = std::lower_bound(b, e, c);
    if (it == e || *it != c) throw Exception("SparseMatrix: entry (" + std::to_string(i) + "," + std::to_string(c) + ") not in the pattern");
    return size_t(it - colnr.begin());
  }
  TM &operator()(size_t i, int c) { return data[GetPosition(i, c)]; }
  // const access to an entry outside the pattern reads as zero (GetPositionTest)
  const TM &operator()(size_t i, int c) const {
    static const TM nul = TM(0.0);
    auto b = colnr.begin() + firsti[i], e = colnr.begin() + firsti[i + 1];
    auto it = std::lower_bound(b, e, c);
    return (it == e || *it != c) ? nul : data[size_t(it - colnr.begin())];
  }
  FlatVector<TM> AsVector() { return FlatVector<TM>(data.size(), data.data()); }
  void PrefetchRow(size_t) const {}
  // sum over the stored entries of the row in storage (= ascending column) order, starting from zero
  TVY RowTimesVector(size_t row, FlatVector<TVX> x) const {
    TVY sum = TVY(0.0);
    for (size_t j = firsti[row]; j < firsti[row + 1]; j++) sum += data[j] * x(colnr[j]);
    return sum;
  }
  void AddRowTransToVector(size_t row, TVY el, FlatVector<TVX> x) const {
    for (size_t j = firsti[row]; j < firsti[row + 1]; j++) x(colnr[j]) += Trans(data[j]) * el;
  }
  void Mult(const BaseVector &x, BaseVector &y) const override {
    auto fx = x.FV<TVX>();
    auto fy = y.FV<TVY>();
    for (size_t i = 0; i < h; i++) fy(i) = RowTimesVector(i, fx);
  }
  void MultAdd(double s, const BaseVector &x, BaseVector &y) const override {
    auto fx = x.FV<TVX>();
    auto fy = y.FV<TVY>();
    for (size_t i = 0; i < h; i++) fy(i) += s * RowTimesVector(i, fx);
  }
};

// sorted unique union of several ascending index lists, f called once per value in ascending order
template <class F> INLINE void MergeArrays(FlatArray<int *> ptrs, FlatArray<int> sizes, F f) {
  const int big = std::numeric_limits<int>::max();
  for (;;) {
    int m = big;
    for (size_t i = 0; i < ptrs.Size(); i++) if (sizes[i] > 0 && *ptrs[i] < m) m = *ptrs[i];
    if (m == big) return;
    f(m);
    for (size_t i = 0; i < ptrs.Size(); i++) if (sizes[i] > 0 && *ptrs[i] == m) { ptrs[i]++; sizes[i]--; }
  }
}

// sort keys ascending, move vals along
template <class T> INLINE void QuickSort(FlatArray<T> a) { std::sort(a.begin(), a.end()); }
// index sort: afterwards a[idx[0]] <= a[idx[1]] <= ... (a itself is untouched)
template <class T> INLINE void QuickSortI(FlatArray<T> a, FlatArray<int> idx) { std::stable_sort(idx.begin(), idx.end(), [&](int x, int y) { return a[x] < a[y]; }); }
template <class T, class S> INLINE void BubbleSort(FlatArray<T> keys, FlatArray<S> vals) {
  for (size_t i = 0; i + 1 < keys.Size(); i++)
    for (size_t j = i + 1; j < keys.Size(); j++)
      if (keys[j] < keys[i]) { std::swap(keys[i], keys[j]); std::swap(vals[i], vals[j]); }
}
}  // namespace ngla
