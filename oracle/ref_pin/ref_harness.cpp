// oracle/_ref/libngsamg_ref.so: the reference's own hot-path functions behind a C ABI.
//
// TEST INFRASTRUCTURE ONLY (oracle/).  The function bodies come from /root/reference: oracle/ref_pin/extract_ref.py cuts them
// out at build time into oracle/_ref/frag/*.inc (git-ignored, never stored in this repository) and this file #includes them
// verbatim -- free functions at namespace scope, member functions into class declarations that repeat the reference's member
// names (the reference's headers cannot be included: they pull in NGSolve's comp.hpp).  The NGSolve containers they call are
// the stand-ins of ngs_standin.hpp.  Everything in THIS file is glue: type aliases, class shells, the C ABI.
//
// Pinned by this library (tests/test_ref_pin.py compares oracle/ngsamg_oracle.c with it):
//   TransposeSPMImpl, MatMultABImpl, RestrictMatrix              utils_sparseMM.cpp:54-238, utils_sparseMM.hpp:93-109
//   GSS3::SetUp/CalcDiags/SmoothRHSInternal/SmoothRESInternal/Smooth/SmoothBack   gssmoother.cpp:110-398
//   BaseSmoother::SmoothSymm/SmoothK/SmoothBackK/SmoothSymmK/CalcResiduum, ProxySmoother::Smooth/SmoothBack  base_smoother.hpp:79-197
//   ProlMap::TransferF2C/AddC2F                                   dof_map.cpp:633-709
//   AMGMatrix::SmoothV/SmoothW/SmoothBS/SmoothVFromLevel          amg_matrix.cpp:37-374
#define PARALLEL 1
#include "ngs_standin_mpi.hpp"
#include "ngs_standin_dense.hpp"

#include <cstdint>
#include <cstring>

namespace amg {
using namespace std;
using namespace ngla;

// ---- aliases the fragments expect (the reference defines them in utils_sparseLA.hpp) ---------------------------------
template <int H, int W> struct spm_entry { typedef Mat<H, W, double> type; };
template <> struct spm_entry<1, 1> { typedef double type; };
template <int H, int W> using SparseMatTM = SparseMatrix<typename spm_entry<H, W>::type>;
template <int H, int W> using SparseMat = SparseMatTM<H, W>;
template <class SPM> struct TM_OF_SPM { typedef typename SPM::TENTRY type; };
template <class V> struct strip_vec { typedef V type; };
template <> struct strip_vec<Vec<1, double>> { typedef double type; };

// ---- pseudo inverse of a diagonal block (utils_denseLA.hpp:93-117, 1237-1569, utils_denseLA.cpp:458-555) --------------
#include "../_ref/frag/la_reltol.inc"
#include "../_ref/frag/la_abstol.inc"
#include "../_ref/frag/la_regtm.inc"
#include "../_ref/frag/la_nzblock.inc"
#include "../_ref/frag/la_nzblock_mat.inc"
template <class TSCAL> bool TryDirectInverse_Lapack(FlatMatrix<TSCAL>, LocalHeap &) { throw Exception("ref harness: blocks of 50+ rows need LAPACK"); }
template <class TSCAL> bool TryDirectInverse_simple(FlatMatrix<TSCAL> A, LocalHeap &lh);
#include "../_ref/frag/la_trydirect.inc"
#include "../_ref/frag/la_trydirect_simple.inc"
#include "../_ref/frag/la_pinv_tol.inc"
#include "../_ref/frag/la_pinv_mat.inc"
#include "../_ref/frag/la_pinv_scal.inc"

// ---- sparse products ---------------------------------------------------------------------------------------------
#include "../_ref/frag/timer_transpose.inc"
#include "../_ref/frag/transpose.inc"
#include "../_ref/frag/timer_matmult.inc"
#include "../_ref/frag/matmult.inc"
#include "../_ref/frag/timer_restrict.inc"
#include "../_ref/frag/restrict.inc"

// GetScalNZE (utils_sparseLA.cpp:324-342) on one rank: NZE * entry size of the local sparse matrix
INLINE size_t GetScalNZE(BaseMatrix const *m) { return m->ScalNZE(); }

// ---- smoothers ---------------------------------------------------------------------------------------------------
class BaseSmoother : public BaseMatrix {
protected:
  shared_ptr<BaseMatrix> sysmat;

public:
  explicit BaseSmoother(shared_ptr<BaseMatrix> m) : sysmat(m) {}
  virtual void Smooth(BaseVector &x, const BaseVector &b, BaseVector &res, bool res_updated = false, bool update_res = true, bool x_zero = false) const = 0;
  virtual void SmoothBack(BaseVector &x, const BaseVector &b, BaseVector &res, bool res_updated = false, bool update_res = true, bool x_zero = false) const = 0;
#include "../_ref/frag/bs_smoothsymm.inc"
#include "../_ref/frag/bs_smoothk.inc"
#include "../_ref/frag/bs_smoothbackk.inc"
#include "../_ref/frag/bs_smoothsymmk.inc"
  virtual shared_ptr<BaseMatrix> GetAMatrix() const { return sysmat; }
#include "../_ref/frag/bs_getnops.inc"
#include "../_ref/frag/bs_getanze.inc"
  int VHeight() const override { return sysmat->VHeight(); }
  int VWidth() const override { return sysmat->VWidth(); }
  void MultAdd(double, const BaseVector &, BaseVector &) const override { throw Exception("BaseSmoother :: MultAdd not overloaded!"); }
#include "../_ref/frag/bs_calcresiduum.inc"
};

class ProxySmoother : public BaseSmoother {
protected:
  shared_ptr<BaseSmoother> sm;
  int nsteps;
  bool symm;

public:
  ProxySmoother(shared_ptr<BaseSmoother> _sm, int _nsteps, bool _symm) : BaseSmoother(_sm->GetAMatrix()), sm(_sm), nsteps(_nsteps), symm(_symm) {}
#include "../_ref/frag/proxy_smooth.inc"
#include "../_ref/frag/proxy_smoothback.inc"
  size_t GetNOps() const override;
  size_t GetANZE() const override;
};
#include "../_ref/frag/proxy_getnops.inc"
#include "../_ref/frag/proxy_getanze.inc"

// Richardson / Jacobi (base_smoother.hpp:250-285, base_smoother.cpp:52-114)
template <class TM> INLINE shared_ptr<SparseMatrix<TM>> GetLocalTMM(shared_ptr<BaseMatrix> A) { return dynamic_pointer_cast<SparseMatrix<TM>>(A); }
class RichardsonSmoother : public BaseSmoother {
public:
  RichardsonSmoother(shared_ptr<BaseMatrix> _mat, shared_ptr<BaseMatrix> _prec, double _omega = 1.0);
  void Smooth(BaseVector &x, const BaseVector &b, BaseVector &res, bool res_updated, bool update_res, bool x_zero) const override;
  void SmoothBack(BaseVector &x, const BaseVector &b, BaseVector &res, bool res_updated, bool update_res, bool x_zero) const override;

protected:
  shared_ptr<BaseMatrix> prec;
  double omega;
};
#include "../_ref/frag/rich_ctor.inc"
#include "../_ref/frag/rich_smooth.inc"
#include "../_ref/frag/rich_smoothback.inc"
template <class TMAT> class JacobiSmoother : public RichardsonSmoother {
public:
  using TM = TMAT;
  JacobiSmoother(shared_ptr<BaseMatrix> _mat, shared_ptr<BitArray> _freedofs, double _omega = 0.9);
  shared_ptr<DiagonalMatrix<TMAT>> DiagInv() const { return diagInv; }

protected:
  shared_ptr<DiagonalMatrix<TMAT>> diagInv;
  shared_ptr<BitArray> freedofs;
};
#include "../_ref/frag/jacobi_ctor.inc"

template <class TM> class GSS3 : public BaseSmoother {
protected:
  size_t H;
  shared_ptr<SparseMatrix<TM>> spmat;
  shared_ptr<BitArray> freedofs;
  bool pinv = false;
  Array<TM> dinv;
  size_t first_free, next_free;
  virtual void SmoothRESInternal(size_t first, size_t next, BaseVector &x, BaseVector &res, bool backwards) const;
  virtual void SmoothRHSInternal(size_t first, size_t next, BaseVector &x, const BaseVector &b, bool backwards) const;

public:
  using TSCAL = double;
  static constexpr int BS() { return ngbla::Height<TM>(); }
  using TV = typename strip_vec<Vec<BS(), TSCAL>>::type;
  GSS3(shared_ptr<SparseMatrix<TM>> mat, shared_ptr<BitArray> subset, bool _pinv) : BaseSmoother(mat), pinv(_pinv) {
    SetUp(mat, subset);
    CalcDiags(Array<TM>());
  }
  GSS3(shared_ptr<SparseMatrix<TM>> mat, FlatArray<TM> repl_diag, shared_ptr<BitArray> subset, bool _pinv) : BaseSmoother(mat), pinv(_pinv) {
    SetUp(mat, subset);
    CalcDiags(repl_diag);
  }
#include "../_ref/frag/gss3_r_smooth.inc"
#include "../_ref/frag/gss3_r_smoothback.inc"
#include "../_ref/frag/gss3_r_smoothres.inc"
#include "../_ref/frag/gss3_r_smoothbackres.inc"
  void SetUp(shared_ptr<SparseMatrix<TM>> mat, shared_ptr<BitArray> subset);
  virtual void CalcDiags(FlatArray<TM> repl_diag);
  void Smooth(BaseVector &x, const BaseVector &b, BaseVector &res, bool res_updated, bool update_res, bool x_zero) const override;
  void SmoothBack(BaseVector &x, const BaseVector &b, BaseVector &res, bool res_updated, bool update_res, bool x_zero) const override;
  FlatArray<TM> DiagInverses() const { return dinv; }
  size_t FirstFree() const { return first_free; }
  size_t NextFree() const { return next_free; }
};
#include "../_ref/frag/gss3_setup.inc"
#include "../_ref/frag/gss3_calcdiags.inc"
#include "../_ref/frag/gss3_rhs.inc"
#include "../_ref/frag/gss3_res.inc"
#include "../_ref/frag/gss3_smooth.inc"
#include "../_ref/frag/gss3_smoothback.inc"

// ---- rigid-body transport Q(t) of the elasticity energy (elasticity_energy.hpp:19-30, 823; elasticity_energy_impl.hpp:8-29) ----------
template <int DIM, class TVD, class TED> class EpsEpsEnergy {
public:
  static constexpr int DISPPV = DIM;
  static constexpr int ROTPV = (DIM == 2) ? 1 : 3;
  static constexpr int DPV = DISPPV + ROTPV;
  typedef Mat<DPV, DPV, double> TM;
  static INLINE void CalcQ(const Vec<DIM> &t, TM &Q, double si, double sj);
};
#include "../_ref/frag/el_calcq.inc"

// ---- grid transfer -------------------------------------------------------------------------------------------------
INLINE Timer<> &timer_hack_prol_f2c() { static Timer t("ProlMap::TransferF2C"); return t; }
INLINE Timer<> &timer_hack_prol_c2f() { static Timer t("ProlMap::TransferC2F"); return t; }

class BaseDOFMapStep {
public:
  virtual ~BaseDOFMapStep() = default;
  virtual void TransferF2C(BaseVector const *x_fine, BaseVector *x_coarse) const = 0;
  virtual void AddC2F(double fac, BaseVector *x_fine, BaseVector const *x_coarse) const = 0;
};
template <class TM> struct trans_entry { typedef Mat<TM::WIDTH, TM::HEIGHT, double> type; };
template <> struct trans_entry<double> { typedef double type; };
template <class TM> class ProlMap : public BaseDOFMapStep {
protected:
  shared_ptr<SparseMatrix<TM>> _prol;
  shared_ptr<SparseMatrix<typename trans_entry<TM>::type>> _prolT;

public:
  ProlMap(shared_ptr<SparseMatrix<TM>> p, shared_ptr<SparseMatrix<typename trans_entry<TM>::type>> pt) : _prol(p), _prolT(pt) {}
  void TransferF2C(BaseVector const *x_fine, BaseVector *x_coarse) const override;
  void AddC2F(double fac, BaseVector *x_fine, BaseVector const *x_coarse) const override;
};
#include "../_ref/frag/prol_f2c.inc"
#include "../_ref/frag/prol_addc2f.inc"

enum { NG_MPI_MAX = 1001 };
struct OneRankComm {      // the communicator of a single-rank AMGMatrix: reductions are the identity
  template <class T, class OP> T AllReduce(T v, OP) const { return v; }
  template <class T> void AllReduceFA(FlatArray<T>) const {}
};
struct DummyUDofs { OneRankComm GetCommunicator() const { return OneRankComm(); } };
class DOFMap {
  Array<shared_ptr<BaseDOFMapStep>> steps;
  DummyUDofs ud;

public:
  void AddStep(shared_ptr<BaseDOFMapStep> s) { steps.Append(s); }
  const DummyUDofs &GetUDofs() const { return ud; }
  void TransferF2C(int level, const BaseVector *fine, BaseVector *coarse) const { steps[level]->TransferF2C(fine, coarse); }
  void AddC2F(int level, double fac, BaseVector *fine, const BaseVector *coarse) const { steps[level]->AddC2F(fac, fine, coarse); }
};

// ---- multigrid cycles ------------------------------------------------------------------------------------------------
class AMGMatrix {
public:
  shared_ptr<DOFMap> map;
  Array<shared_ptr<BaseSmoother>> smoothers;
  Array<shared_ptr<BaseVector>> x_level, rhs_level, res_level;
  int n_levels = 0;
  bool drops_out = false, has_crs_inv = false;
  shared_ptr<BaseMatrix> crs_inv, crs_mat;
  int vwb = 0;
#include "../_ref/frag/amg_smooth.inc"
  void Mult(const BaseVector &b, BaseVector &x) const;
  void MultTrans(const BaseVector &b, BaseVector &x) const;
  void MultAdd(double s, const BaseVector &b, BaseVector &x) const;
  void MultTransAdd(double s, const BaseVector &b, BaseVector &x) const;
  Array<double> GetOC() const;
  void SmoothV(BaseVector &x, const BaseVector &b) const;
  void SmoothW(BaseVector &x, const BaseVector &b) const;
  void SmoothBS(BaseVector &x, const BaseVector &b) const;
  void SmoothVFromLevel(int startlevel, BaseVector &x, const BaseVector &b, BaseVector &res, bool res_updated, bool update_res, bool x_zero) const;
};
#include "../_ref/frag/amg_smoothw.inc"
#include "../_ref/frag/amg_smoothbs.inc"
#include "../_ref/frag/amg_smoothv.inc"
#include "../_ref/frag/amg_smoothvfrom.inc"
#include "../_ref/frag/amg_mult.inc"
#include "../_ref/frag/amg_multtrans.inc"
#include "../_ref/frag/amg_multadd.inc"
#include "../_ref/frag/amg_multtransadd.inc"
#include "../_ref/frag/amg_getoc.inc"

// dense coarse inverse handed in by the caller (the reference uses NGSolve's sparse Cholesky, which is not in its tree)
class DenseInverse : public BaseMatrix {
  size_t n;
  std::vector<double> inv;

public:
  DenseInverse(size_t an, const double *p) : n(an), inv(p, p + an * an) {}
  int VHeight() const override { return int(n); }
  int VWidth() const override { return int(n); }
  void Mult(const BaseVector &b, BaseVector &x) const override {
    auto fb = b.FVDouble(), fx = x.FVDouble();
    for (size_t i = 0; i < n; i++) {
      double s = 0.0;
      for (size_t j = 0; j < n; j++) s += inv[i * n + j] * fb(j);
      fx(i) = s;
    }
  }
  void MultAdd(double s, const BaseVector &b, BaseVector &x) const override {
    auto fb = b.FVDouble(), fx = x.FVDouble();
    for (size_t i = 0; i < n; i++) {
      double t = 0.0;
      for (size_t j = 0; j < n; j++) t += inv[i * n + j] * fb(j);
      fx(i) += s * t;
    }
  }
};
}  // namespace amg

#include "ref_par.hpp"   // the multi-rank path: class shells + fragments, C ABI ref_par_*

// =====================================================================================================================
// C ABI
// =====================================================================================================================
using namespace amg;
typedef int64_t i64;
typedef int32_t i32;

namespace {
thread_local std::string g_err;

struct MatH {
  int bh, bw;
  shared_ptr<BaseMatrix> m;
};

template <int H, int W> shared_ptr<SparseMatTM<H, W>> as(const MatH *h) {
  auto p = dynamic_pointer_cast<SparseMatTM<H, W>>(h->m);
  if (!p) throw Exception("ref harness: block size mismatch");
  return p;
}

template <int H, int W> MatH *make_mat(i64 nr, i64 nc, const i64 *rp, const i32 *ci, const double *v) {
  Array<int> cnt((size_t)nr);
  for (i64 i = 0; i < nr; i++) cnt[i] = int(rp[i + 1] - rp[i]);
  auto m = make_shared<SparseMatTM<H, W>>(cnt, (size_t)nc);
  for (i64 i = 0; i < nr; i++) {
    auto ri = m->GetRowIndices(i);
    auto rv = m->GetRowValues(i);
    for (i64 k = rp[i]; k < rp[i + 1]; k++) {
      ri[k - rp[i]] = ci[k];
      std::memcpy((void *)&rv[k - rp[i]], v + k * H * W, sizeof(double) * H * W);
    }
  }
  return new MatH{H, W, m};
}

template <int H, int W> void fetch_mat(const MatH *h, i64 *rp, i32 *ci, double *v) {
  auto m = as<H, W>(h);
  i64 k = 0;
  rp[0] = 0;
  for (size_t i = 0; i < m->Height(); i++) {
    auto ri = m->GetRowIndices(i);
    auto rv = m->GetRowValues(i);
    for (size_t j = 0; j < ri.Size(); j++, k++) {
      ci[k] = ri[j];
      std::memcpy(v + k * H * W, (const void *)&rv[j], sizeof(double) * H * W);
    }
    rp[i + 1] = k;
  }
}

template <int H, int W> i64 nze(const MatH *h) { return (i64)as<H, W>(h)->NZE(); }

// block shapes of the h1 / elasticity hierarchies: scalar, 2D (2 -> 3), 3D (3 -> 6)
#define REF_FOR_SHAPES(X) X(1, 1) X(2, 2) X(2, 3) X(3, 2) X(3, 3) X(3, 6) X(6, 3) X(6, 6)
#define REF_FOR_TRIPLES(X) \
  X(1, 1, 1) X(2, 2, 2) X(2, 2, 3) X(2, 3, 2) X(2, 3, 3) X(3, 2, 2) X(3, 2, 3) X(3, 3, 2) X(3, 3, 3) \
  X(3, 3, 6) X(3, 6, 3) X(3, 6, 6) X(6, 3, 3) X(6, 3, 6) X(6, 6, 3) X(6, 6, 6)

struct LevelH {
  int b = 0;
  MatH *A = nullptr;        // not owned for level 0 copies: always owned here
  shared_ptr<BitArray> free;
  MatH *P = nullptr, *PT = nullptr;
  shared_ptr<BaseSmoother> gs;   // the bare GSS3
};

struct AmgH {
  bool pinv = false;
  int nlevels;
  std::vector<LevelH> lev;
  AMGMatrix amg;
};

template <int B> shared_ptr<BaseSmoother> make_gss3(LevelH &L, bool pinv) {
  return make_shared<GSS3<typename spm_entry<B, B>::type>>(as<B, B>(L.A), L.free, pinv);
}
template <int N> void pinv_block(double *m) {
  typename spm_entry<N, N>::type blk;
  std::memcpy((void *)&blk, m, sizeof(double) * N * N);
  LocalHeap lh(1024 * 1024, "pinv");
  CalcPseudoInverseTryNormal(blk, lh);
  std::memcpy(m, (const void *)&blk, sizeof(double) * N * N);
}
template <int B> shared_ptr<BaseSmoother> make_jacobi(LevelH &L, double omega) {
  return make_shared<JacobiSmoother<typename spm_entry<B, B>::type>>(as<B, B>(L.A), L.free, omega);
}

template <int B> void dinv_out(LevelH &L, double *out) {
  auto g = dynamic_pointer_cast<GSS3<typename spm_entry<B, B>::type>>(L.gs);
  auto d = g->DiagInverses();
  for (size_t i = 0; i < d.Size(); i++) std::memcpy(out + i * B * B, (const void *)&d[i], sizeof(double) * B * B);
}
template <int BF, int BC> void add_step(AmgH *a, LevelH &L) {
  typedef typename spm_entry<BF, BC>::type TM;
  a->amg.map->AddStep(make_shared<ProlMap<TM>>(as<BF, BC>(L.P), as<BC, BF>(L.PT)));
}
template <int BF, int BC> MatH *rap(LevelH &L) {
  auto pt = TransposeSPMImpl<BF, BC>(*as<BF, BC>(L.P));
  L.PT = new MatH{BC, BF, pt};
  auto ac = RestrictMatrix<BF, BC>(*pt, *as<BF, BF>(L.A), *as<BF, BC>(L.P));
  return new MatH{BC, BC, ac};
}

template <class F> int guarded(F f) {
  try {
    f();
    return 0;
  } catch (const std::exception &e) {
    g_err = e.what();
    return 1;
  }
}
}  // namespace

extern "C" {

const char *ref_last_error() { return g_err.c_str(); }

// which reference lines this library was built from (oracle/_ref/frag/INDEX.txt, embedded at build time)
const char *ref_fragment_index() {
  return
#include "../_ref/frag/INDEX.inc"
      ;
}

void *ref_mat_new(i64 nr, i64 nc, int bh, int bw, const i64 *rp, const i32 *ci, const double *v) {
  MatH *out = nullptr;
  int rc = guarded([&] {
#define X(H, W) if (bh == H && bw == W) out = make_mat<H, W>(nr, nc, rp, ci, v);
    REF_FOR_SHAPES(X)
#undef X
    if (!out) throw Exception("ref_mat_new: unsupported block shape");
  });
  return rc ? nullptr : out;
}

void ref_mat_free(void *h) { delete (MatH *)h; }

int ref_mat_info(const void *hv, i64 *nr, i64 *nc, int *bh, int *bw, i64 *nnz) {
  const MatH *h = (const MatH *)hv;
  return guarded([&] {
    *nr = (i64)h->m->Height(); *nc = (i64)h->m->Width(); *bh = h->bh; *bw = h->bw;
#define X(H, W) if (h->bh == H && h->bw == W) *nnz = nze<H, W>(h);
    REF_FOR_SHAPES(X)
#undef X
  });
}

int ref_mat_fetch(const void *hv, i64 *rp, i32 *ci, double *v) {
  const MatH *h = (const MatH *)hv;
  return guarded([&] {
#define X(H, W) if (h->bh == H && h->bw == W) fetch_mat<H, W>(h, rp, ci, v);
    REF_FOR_SHAPES(X)
#undef X
  });
}

// TransposeSPMImpl
void *ref_mat_transpose(const void *hv) {
  const MatH *h = (const MatH *)hv;
  MatH *out = nullptr;
  int rc = guarded([&] {
#define X(H, W) if (h->bh == H && h->bw == W) out = new MatH{W, H, TransposeSPMImpl<H, W>(*as<H, W>(h))};
    REF_FOR_SHAPES(X)
#undef X
  });
  return rc ? nullptr : out;
}

// MatMultABImpl
void *ref_mat_mult(const void *av, const void *bv) {
  const MatH *a = (const MatH *)av, *b = (const MatH *)bv;
  MatH *out = nullptr;
  int rc = guarded([&] {
    if (a->bw != b->bh) throw Exception("ref_mat_mult: inner block sizes differ");
#define X(A, B, C) if (a->bh == A && a->bw == B && b->bw == C) out = new MatH{A, C, MatMultABImpl<A, B, C>(*as<A, B>(a), *as<B, C>(b))};
    REF_FOR_TRIPLES(X)
#undef X
    if (!out) throw Exception("ref_mat_mult: unsupported block shapes");
  });
  return rc ? nullptr : out;
}

// RestrictMatrix(PT, A, P)
void *ref_mat_restrict(const void *ptv, const void *av, const void *pv) {
  const MatH *pt = (const MatH *)ptv, *a = (const MatH *)av, *p = (const MatH *)pv;
  MatH *out = nullptr;
  int rc = guarded([&] {
#define X(H, W) if (p->bh == H && p->bw == W) out = new MatH{W, W, RestrictMatrix<H, W>(*as<W, H>(pt), *as<H, H>(a), *as<H, W>(p))};
    REF_FOR_SHAPES(X)
#undef X
  });
  return rc ? nullptr : out;
}

// ---- hierarchy: level matrices by the reference's RAP, GSS3 (+ProxySmoother) per level, cycles by AMGMatrix ------------
void *ref_amg_new(int nlevels) {
  // nlevels is an upper bound: the hierarchy ends with the last level ref_amg_set_prol produced
  AmgH *a = new AmgH;
  a->nlevels = 1;
  a->lev.resize(nlevels);
  a->amg.map = make_shared<DOFMap>();
  a->amg.n_levels = 1;
  return a;
}

void ref_amg_free(void *hv) {
  AmgH *a = (AmgH *)hv;
  for (auto &L : a->lev) { delete L.A; delete L.P; delete L.PT; }
  delete a;
}

int ref_amg_set_matrix(void *hv, i64 n, int b, const i64 *rp, const i32 *ci, const double *v, const uint8_t *freed) {
  AmgH *a = (AmgH *)hv;
  return guarded([&] {
    LevelH &L = a->lev[0];
    L.b = b;
    L.A = (MatH *)ref_mat_new(n, n, b, b, rp, ci, v);
    if (!L.A) throw Exception(g_err);
    if (freed) {
      L.free = make_shared<BitArray>((size_t)n);
      for (i64 i = 0; i < n; i++) if (freed[i]) L.free->SetBit(i);
    }
  });
}

// prolongation of level l; builds PT (TransposeSPMImpl) and A_{l+1} (RestrictMatrix)
int ref_amg_set_prol(void *hv, int l, i64 nc, int bc, const i64 *rp, const i32 *ci, const double *v) {
  AmgH *a = (AmgH *)hv;
  return guarded([&] {
    if (l + 1 != a->nlevels || l + 1 >= (int)a->lev.size()) throw Exception("ref_amg_set_prol: levels must be added in order, below the bound given to ref_amg_new");
    LevelH &L = a->lev[l], &C = a->lev[l + 1];
    L.P = (MatH *)ref_mat_new((i64)L.A->m->Height(), nc, L.b, bc, rp, ci, v);
    if (!L.P) throw Exception(g_err);
    MatH *ac = nullptr;
#define X(H, W) if (L.b == H && bc == W) { ac = rap<H, W>(L); add_step<H, W>(a, L); }
    REF_FOR_SHAPES(X)
#undef X
    if (!ac) throw Exception("ref_amg_set_prol: unsupported block shapes");
    C.b = bc;
    C.A = ac;
    a->nlevels = a->amg.n_levels = l + 2;
  });
}

// CalcPseudoInverseTryNormal on one n x n block (row-major, in place); n in {1, 2, 3, 6}
int ref_pinv(int n, double *m) {
  return guarded([&] {
    if (n == 1) pinv_block<1>(m);
    else if (n == 2) pinv_block<2>(m);
    else if (n == 3) pinv_block<3>(m);
    else if (n == 6) pinv_block<6>(m);
    else throw Exception("ref_pinv: unsupported block size");
  });
}

// RegTM<0,6,6> on one 6 x 6 block: what RegularizeMatrix (elasticity_pc_impl.hpp:734-763) applies to the coarsest diagonal blocks
int ref_regularize6(double *m) {
  return guarded([&] {
    Mat<6, 6> blk;
    std::memcpy((void *)&blk, m, sizeof(double) * 36);
    RegTM<0, 6, 6>(blk);
    std::memcpy(m, (const void *)&blk, sizeof(double) * 36);
  });
}

// EpsEpsEnergy<DIM>::CalcQ(t, Q, si, sj): dim 3 -> q is 6 x 6, dim 2 -> 3 x 3 (row-major)
int ref_elast_calcq(int dim, const double *t, double si, double sj, double *q) {
  return guarded([&] {
    if (dim == 3) {
      Vec<3> tv; for (int i = 0; i < 3; i++) tv(i) = t[i];
      Mat<6, 6> Q;
      EpsEpsEnergy<3, int, int>::CalcQ(tv, Q, si, sj);
      std::memcpy(q, (const void *)&Q, sizeof(double) * 36);
    } else if (dim == 2) {
      Vec<2> tv; for (int i = 0; i < 2; i++) tv(i) = t[i];
      Mat<3, 3> Q;
      EpsEpsEnergy<2, int, int>::CalcQ(tv, Q, si, sj);
      std::memcpy(q, (const void *)&Q, sizeof(double) * 9);
    } else throw Exception("ref_elast_calcq: dim must be 2 or 3");
  });
}

// GSS3(..., pinv): pseudo-inverted diagonal blocks (ngs_amg_regularize_cmats); call before ref_amg_finalize
void ref_amg_set_pinv(void *hv, int pinv) { ((AmgH *)hv)->pinv = pinv != 0; }

// smoothers (GSS3, wrapped into a ProxySmoother when sm_steps > 1 or sm_symm), level vectors, optional dense coarse inverse
// Jacobi instead of Gauss-Seidel on every level (call after ref_amg_finalize): JacobiSmoother(A, free, omega)
int ref_amg_use_jacobi(void *hv, double omega, int sm_steps, int sm_symm) {
  AmgH *a = (AmgH *)hv;
  return guarded([&] {
    for (int l = 0; l + 1 < a->nlevels; l++) {
      LevelH &L = a->lev[l];
      if (L.b == 1) L.gs = make_jacobi<1>(L, omega);
      else if (L.b == 2) L.gs = make_jacobi<2>(L, omega);
      else if (L.b == 3) L.gs = make_jacobi<3>(L, omega);
      else if (L.b == 6) L.gs = make_jacobi<6>(L, omega);
      else throw Exception("ref_amg_use_jacobi: unsupported block size");
      a->amg.smoothers[l] = (sm_steps > 1 || sm_symm) ? shared_ptr<BaseSmoother>(make_shared<ProxySmoother>(L.gs, sm_steps, sm_symm != 0)) : L.gs;
    }
  });
}

int ref_amg_finalize(void *hv, int sm_steps, int sm_symm, const double *coarse_inv) {
  AmgH *a = (AmgH *)hv;
  return guarded([&] {
    AMGMatrix &M = a->amg;
    M.smoothers.SetSize(a->nlevels - 1);
    M.x_level.SetSize(a->nlevels); M.rhs_level.SetSize(a->nlevels); M.res_level.SetSize(a->nlevels);
    for (int l = 0; l < a->nlevels; l++) {
      LevelH &L = a->lev[l];
      const size_t n = L.A->m->Height();
      M.x_level[l] = make_shared<BaseVector>(n, L.b);
      M.rhs_level[l] = make_shared<BaseVector>(n, L.b);
      M.res_level[l] = make_shared<BaseVector>(n, L.b);
      if (l + 1 == a->nlevels) break;
      if (L.b == 1) L.gs = make_gss3<1>(L, a->pinv);
      else if (L.b == 2) L.gs = make_gss3<2>(L, a->pinv);
      else if (L.b == 3) L.gs = make_gss3<3>(L, a->pinv);
      else if (L.b == 6) L.gs = make_gss3<6>(L, a->pinv);
      else throw Exception("ref_amg_finalize: unsupported block size");
      M.smoothers[l] = (sm_steps > 1 || sm_symm) ? shared_ptr<BaseSmoother>(make_shared<ProxySmoother>(L.gs, sm_steps, sm_symm != 0)) : L.gs;
    }
    if (coarse_inv) {
      LevelH &L = a->lev[a->nlevels - 1];
      M.crs_inv = make_shared<DenseInverse>(L.A->m->Height() * (size_t)L.b, coarse_inv);
      M.has_crs_inv = true;
    }
  });
}

const void *ref_amg_level_matrix(const void *hv, int l) { return ((const AmgH *)hv)->lev[l].A; }
const void *ref_amg_level_pt(const void *hv, int l) { return ((const AmgH *)hv)->lev[l].PT; }

int ref_amg_level_dinv(void *hv, int l, double *out) {
  AmgH *a = (AmgH *)hv;
  return guarded([&] {
    LevelH &L = a->lev[l];
    if (L.b == 1) dinv_out<1>(L, out);
    else if (L.b == 2) dinv_out<2>(L, out);
    else if (L.b == 3) dinv_out<3>(L, out);
    else dinv_out<6>(L, out);
  });
}

// which: 0 = x_level, 1 = rhs_level, 2 = res_level
int ref_amg_level_vec(const void *hv, int which, int l, double *out) {
  const AmgH *a = (const AmgH *)hv;
  return guarded([&] {
    const auto &arr = which == 0 ? a->amg.x_level : which == 1 ? a->amg.rhs_level : a->amg.res_level;
    auto fv = arr[l]->FVDouble();
    std::memcpy(out, fv.Data(), sizeof(double) * fv.Size());
  });
}

// one call of the level's smoother with the reference's flag protocol; bare != 0 bypasses the ProxySmoother
int ref_amg_smooth(void *hv, int l, double *x, const double *b, double *res, int res_updated, int update_res, int x_zero, int backwards, int bare) {
  AmgH *a = (AmgH *)hv;
  return guarded([&] {
    LevelH &L = a->lev[l];
    const size_t n = L.A->m->Height(), nb = n * (size_t)L.b;
    BaseVector vx(n, L.b), vb(n, L.b), vr(n, L.b);
    std::memcpy(vx.FVDouble().Data(), x, sizeof(double) * nb);
    std::memcpy(vb.FVDouble().Data(), b, sizeof(double) * nb);
    std::memcpy(vr.FVDouble().Data(), res, sizeof(double) * nb);
    const BaseSmoother &S = bare ? *L.gs : *a->amg.smoothers[l];
    if (backwards) S.SmoothBack(vx, vb, vr, res_updated != 0, update_res != 0, x_zero != 0);
    else S.Smooth(vx, vb, vr, res_updated != 0, update_res != 0, x_zero != 0);
    std::memcpy(x, vx.FVDouble().Data(), sizeof(double) * nb);
    std::memcpy(res, vr.FVDouble().Data(), sizeof(double) * nb);
  });
}

// AMGMatrix::MultAdd / MultTransAdd (amg_matrix.cpp:385-393) with SetVWB(cycle): x += s * C b; trans != 0 takes the Trans entry
int ref_amg_mult_add(void *hv, int cycle, int trans, double s, const double *b, double *x) {
  AmgH *a = (AmgH *)hv;
  return guarded([&] {
    LevelH &L = a->lev[0];
    const size_t n = L.A->m->Height(), nb = n * (size_t)L.b;
    BaseVector vx(n, L.b), vb(n, L.b);
    std::memcpy(vb.FVDouble().Data(), b, sizeof(double) * nb);
    std::memcpy(vx.FVDouble().Data(), x, sizeof(double) * nb);
    a->amg.vwb = cycle;
    if (trans) a->amg.MultTransAdd(s, vb, vx);
    else a->amg.MultAdd(s, vb, vx);
    a->amg.vwb = 0;
    std::memcpy(x, vx.FVDouble().Data(), sizeof(double) * nb);
  });
}

// AMGMatrix::GetOC (amg_matrix.cpp:551-582) with SetVWB(cycle): [OC, OC_l0, ...]; returns the number of entries written (<= cap)
int ref_amg_get_oc(void *hv, int cycle, double *out, int cap) {
  AmgH *a = (AmgH *)hv;
  int n = 0;
  int rc = guarded([&] {
    a->amg.vwb = cycle;
    auto occs = a->amg.GetOC();
    a->amg.vwb = 0;
    n = (int)occs.Size();
    for (int i = 0; i < n && i < cap; i++) out[i] = occs[i];
  });
  return rc ? -1 : n;
}

// AMGMatrix::Mult / MultTrans (amg_matrix.cpp:377-383)
int ref_amg_mult(void *hv, int cycle, int trans, const double *b, double *x) {
  AmgH *a = (AmgH *)hv;
  return guarded([&] {
    LevelH &L = a->lev[0];
    const size_t n = L.A->m->Height(), nb = n * (size_t)L.b;
    BaseVector vx(n, L.b), vb(n, L.b);
    std::memcpy(vb.FVDouble().Data(), b, sizeof(double) * nb);
    std::memcpy(vx.FVDouble().Data(), x, sizeof(double) * nb);       // Mult must overwrite whatever is in x
    a->amg.vwb = cycle;
    if (trans) a->amg.MultTrans(vb, vx);
    else a->amg.Mult(vb, vx);
    a->amg.vwb = 0;
    std::memcpy(x, vx.FVDouble().Data(), sizeof(double) * nb);
  });
}

// cycle: 0 = V (SmoothV), 1 = W (SmoothW), 2 = BS (SmoothBS)
int ref_amg_apply(void *hv, int cycle, const double *b, double *x) {
  AmgH *a = (AmgH *)hv;
  return guarded([&] {
    LevelH &L = a->lev[0];
    const size_t n = L.A->m->Height(), nb = n * (size_t)L.b;
    BaseVector vx(n, L.b), vb(n, L.b);
    std::memcpy(vb.FVDouble().Data(), b, sizeof(double) * nb);
    if (cycle == 0) a->amg.SmoothV(vx, vb);
    else if (cycle == 1) a->amg.SmoothW(vx, vb);
    else if (cycle == 2) a->amg.SmoothBS(vx, vb);
    else throw Exception("ref_amg_apply: unknown cycle");
    std::memcpy(x, vx.FVDouble().Data(), sizeof(double) * nb);
  });
}

// PCG around the reference's cycle.  GLUE, not reference code: ngsolve.krylovspace.CGSolver lives in NGSolve; the loop is the same
// restatement as orc_amg_pcg (oracle/ngsamg_oracle.c): the preconditioner C is AMGMatrix::SmoothV/W/BS, A*s is SparseMatrix::MultAdd.
int ref_amg_pcg(void *hv, int cycle, const double *rhs, double *u, double tol, int maxsteps, double *errors, int *iterations) {
  AmgH *a = (AmgH *)hv;
  return guarded([&] {
    LevelH &L = a->lev[0];
    const size_t n = L.A->m->Height(), nb = n * (size_t)L.b;
    BaseVector vd(n, L.b), vw(n, L.b), vs(n, L.b);
    auto d = vd.FVDouble(), w = vw.FVDouble(), sv = vs.FVDouble();
    auto precond = [&](const BaseVector &in, BaseVector &out) {
      if (cycle == 0) a->amg.SmoothV(out, in);
      else if (cycle == 1) a->amg.SmoothW(out, in);
      else a->amg.SmoothBS(out, in);
    };
    auto dot = [&](FlatVector<double> p, FlatVector<double> q) { double t = 0; for (size_t i = 0; i < nb; i++) t += p(i) * q(i); return t; };
    for (size_t i = 0; i < nb; i++) { u[i] = 0.0; d(i) = rhs[i]; }
    precond(vd, vw);
    for (size_t i = 0; i < nb; i++) sv(i) = w(i);
    double wdn = dot(w, d);
    const double err0 = std::sqrt(std::fabs(wdn));
    if (errors) errors[0] = err0;
    int it = 0;
    if (wdn != 0.0)
      for (it = 1; it <= maxsteps; it++) {
        L.A->m->Mult(vs, vw);
        const double wd = wdn, alpha = wd / dot(sv, w);
        for (size_t i = 0; i < nb; i++) u[i] += alpha * sv(i);
        for (size_t i = 0; i < nb; i++) d(i) -= alpha * w(i);
        precond(vd, vw);
        wdn = dot(w, d);
        const double beta = wdn / wd;
        for (size_t i = 0; i < nb; i++) sv(i) = beta * sv(i) + w(i);
        const double err = std::sqrt(std::fabs(wd));
        if (errors) errors[it] = err;
        if (err < tol * err0) break;
      }
    if (it > maxsteps) it = maxsteps;
    *iterations = it;
  });
}
}  // extern "C"

#include "ref_par_abi.hpp"
