// ngsamg_b200_ngsolve.hpp -- the reference-side adapter: NgsAMG's preconditioner registration and NGSolve BaseMatrix interface on top of
// the C ABI (include/ngsamg_b200.h).  A maintainer adds this one header to the reference tree (e.g. as src/base/precond/b200_pc.hpp),
// includes it from one translation unit and links -lngsamg_b200; "NgsAMG.h1_scal" / "NgsAMG.elast_3d" then resolve to the B200 path and
// every caller of ngcomp::Preconditioner / BaseMatrix::Mult / MultAdd keeps working unchanged.
//
// Replaces, for the hot path only:
//   AMGMatrix   : src/base/solve/amg_matrix.hpp:14-87  (Mult/MultAdd/MultTrans/MultTransAdd quartet, :55-66; amg_matrix.cpp:377-393)
//   BaseAMGPC   : src/base/precond/amg_pc.hpp:26-228   (InitLevel amg_pc.cpp:398-410, FinalizeLevel :413-434, GetMatrix/GetAMatrix)
//   registration: src/base/utils/amg_register.hpp:79-98, src/h1/h1_dim1.cpp:76, src/elasticity/elasticity.hpp:104-140
//
// Compiles against NGSolve (#include <comp.hpp>).  The repo's tests compile it against a stand-in of the few NGSolve types it touches
// (tests/cpp/ngcomp_standin.hpp on top of oracle/ref_pin/ngs_standin.hpp): define NGSAMG_B200_NGSOLVE_STANDIN and include the stand-in first.
#pragma once
#ifndef NGSAMG_B200_NGSOLVE_STANDIN
#include <comp.hpp>            // NGSolve, as in src/base/base.hpp:4
#endif

#include <cstdint>
#include <memory>
#include <string>
#include <vector>

#include "ngsamg_b200.h"

#ifndef NGSAMG_B200_BASEMATRIX
#define NGSAMG_B200_BASEMATRIX ngla::BaseMatrix   // (the tests' stand-in keeps the full virtual interface in a derived class)
#endif

namespace amg {

using ngcore::Flags;
using std::shared_ptr;
using std::string;

inline void b200_check(int rc)
{
  if (rc) throw ngcore::Exception(string("ngsamg_b200: ") + ngsamg_b200_last_error());   // the reference throws ngcore::Exception (amg_pc.cpp:430)
}

// AMGMatrix on the device.  Not re-entrant, like the reference's (shared work vectors mutated inside const Mult, amg_matrix.cpp:187-189).
class B200AMGMatrix : public NGSAMG_B200_BASEMATRIX
{
  ngsamg_b200_t *h;
  size_t n;
public:
  B200AMGMatrix(ngsamg_b200_t *ah, size_t an) : h(ah), n(an) {}
  B200AMGMatrix(const B200AMGMatrix &) = delete;
  ~B200AMGMatrix() override { ngsamg_b200_destroy(h); }
  ngsamg_b200_t *Handle() const { return h; }
  int VHeight() const override { return int(n); }
  int VWidth() const override { return int(n); }
  void Mult(const ngla::BaseVector &b, ngla::BaseVector &x) const override                         // amg_matrix.cpp:377-378
  { b200_check(ngsamg_b200_apply(h, b.FVDouble().Data(), x.FVDouble().Data())); }
  void MultAdd(double s, const ngla::BaseVector &b, ngla::BaseVector &x) const override             // amg_matrix.cpp:385-389
  { b200_check(ngsamg_b200_apply_add(h, s, b.FVDouble().Data(), x.FVDouble().Data())); }
  void MultTrans(const ngla::BaseVector &b, ngla::BaseVector &x) const override { Mult(b, x); }    // :381-382 (the cycle is symmetric)
  void MultTransAdd(double s, const ngla::BaseVector &b, ngla::BaseVector &x) const override { MultAdd(s, b, x); }   // :392-393
  ngla::AutoVector CreateRowVector() const override { return std::make_unique<ngla::VVector<double>>(n); }
  ngla::AutoVector CreateColVector() const override { return std::make_unique<ngla::VVector<double>>(n); }
  // introspection the reference exposes to Python (python_amg.hpp:30-101)
  int GetNLevels(int /*rank*/) const { return ngsamg_b200_num_levels(h); }
  size_t GetNDof(int level, int /*rank*/) const
  {
    ngsamg_level_info info;
    b200_check(ngsamg_b200_level_info(h, level, &info));
    return size_t(info.n);
  }
  double GetOC() const { return ngsamg_b200_operator_complexity(h); }
};

// BS = entries per block row of the fine matrix: 1 (h1_scal), 3 (elast_3d, displacement formulation), 6 (elast_3d with rotations)
template <int BS>
class B200AMGPC : public ngcomp::Preconditioner
{
protected:
  string type;
  Flags pcflags;
  shared_ptr<ngcore::BitArray> freedofs;
  shared_ptr<B200AMGMatrix> amg_mat;
  const ngla::BaseMatrix *finest = nullptr;       // the caller keeps the matrix alive (NOOP_Deleter in the reference, amg_pc.cpp:413-418)
  std::vector<double> vertex_xyz;                 // elasticity: one point per block row (ElastVData::pos), set by SetVertexCoordinates
  int device = 0;
public:
  B200AMGPC(shared_ptr<ngcomp::BilinearForm> bfa, const Flags &f, const string name, string atype)
    : ngcomp::Preconditioner(bfa, f, name), type(std::move(atype)), pcflags(f)
  { device = int(f.GetNumFlag("ngs_amg_b200_device", 0)); }

  // the reference reads the vertex positions from the mesh (amg_pc_vertex_impl.hpp:941-1103); callers without a mesh hand them in
  void SetVertexCoordinates(const double *xyz, size_t nvert) { vertex_xyz.assign(xyz, xyz + 3 * nvert); }

  void InitLevel(shared_ptr<ngcore::BitArray> afreedofs) override { freedofs = afreedofs; }         // amg_pc.cpp:398-410

  void FinalizeLevel(const ngla::BaseMatrix *mat) override                                            // amg_pc.cpp:413-434
  {
    using TM = typename std::conditional<BS == 1, double, ngbla::Mat<BS, BS, double>>::type;
    auto *spm = dynamic_cast<const ngla::SparseMatrix<TM> *>(mat);
    if (!spm) throw ngcore::Exception("B200AMGPC::FinalizeLevel: need a SparseMatrix with " + std::to_string(BS) + "x" + std::to_string(BS) + " entries");
    finest = mat;
    const size_t n = spm->Height();
    std::vector<int64_t> rowptr(n + 1);                                  // NGSolve's firsti is size_t
    for (size_t i = 0; i <= n; i++) rowptr[i] = int64_t(spm->First(i));
    const int *cols = n ? spm->GetRowIndices(0).Data() : nullptr;       // colnr / data are contiguous over the rows
    const double *vals = n ? reinterpret_cast<const double *>(spm->GetRowValues(0).Data()) : nullptr;
    ngsamg_csr A{int64_t(n), int64_t(n), BS, BS, rowptr.data(), cols, vals};
    std::vector<uint8_t> fm(n, 1);
    if (freedofs)
      for (size_t i = 0; i < n; i++) fm[i] = freedofs->Test(i) ? 1 : 0;
    if (BS > 1 && vertex_xyz.size() != 3 * n) throw ngcore::Exception("B200AMGPC: " + type + " needs the vertex coordinates (SetVertexCoordinates)");
    // every flag travels as a (key, value) string pair; unknown keys are ignored by the library like NGSolve's Flags ignores them
    std::vector<string> keys, vals_s;
    for (int i = 0; i < pcflags.GetNNumFlags(); i++) { string k; const double v = pcflags.GetNumFlag(i, k); keys.push_back(k); vals_s.push_back(num_to_string(v)); }
    for (int i = 0; i < pcflags.GetNStringFlags(); i++) { string k; const string v = pcflags.GetStringFlag(i, k); keys.push_back(k); vals_s.push_back(v); }
    for (int i = 0; i < pcflags.GetNDefineFlags(); i++) { string k; const bool v = pcflags.GetDefineFlag(i, k); keys.push_back(k); vals_s.push_back(v ? "1" : "0"); }
    std::vector<const char *> kp, vp;
    for (size_t i = 0; i < keys.size(); i++) { kp.push_back(keys[i].c_str()); vp.push_back(vals_s[i].c_str()); }
    ngsamg_b200_t *h = nullptr;
    b200_check(ngsamg_b200_create(type.c_str(), &A, fm.data(), vertex_xyz.empty() ? nullptr : vertex_xyz.data(), kp.data(), vp.data(), int(kp.size()),
                                  device, &h));
    if (int rc = ngsamg_b200_finalize(h)) { ngsamg_b200_destroy(h); b200_check(rc); }
    amg_mat = std::make_shared<B200AMGMatrix>(h, n * BS);
  }

  void Update() override {}
  const ngla::BaseMatrix &GetMatrix() const override { need(); return *amg_mat; }
  const ngla::BaseMatrix &GetAMatrix() const override
  {
    if (!finest) throw ngcore::Exception("B200AMGPC: FinalizeLevel has not been called");
    return *finest;
  }
  shared_ptr<B200AMGMatrix> GetAMGMatrix() const { need(); return amg_mat; }
  // Preconditioner is a BaseMatrix: the quartet forwards to the cycle (amg_pc.hpp:147-176)
  void Mult(const ngla::BaseVector &b, ngla::BaseVector &x) const override { need(); amg_mat->Mult(b, x); }
  void MultAdd(double s, const ngla::BaseVector &b, ngla::BaseVector &x) const override { need(); amg_mat->MultAdd(s, b, x); }
  void MultTrans(const ngla::BaseVector &b, ngla::BaseVector &x) const override { need(); amg_mat->MultTrans(b, x); }
  void MultTransAdd(double s, const ngla::BaseVector &b, ngla::BaseVector &x) const override { need(); amg_mat->MultTransAdd(s, b, x); }
  int VHeight() const override { need(); return amg_mat->VHeight(); }
  int VWidth() const override { need(); return amg_mat->VWidth(); }
  bool IsComplex() const override { return false; }

  // the whole PCG on the device (replaces ngsolve.krylovspace.CGSolver as called by tests/h1/amg_utils.py:346-349); returns the iteration count
  int SolveCG(const ngla::BaseVector &rhs, ngla::BaseVector &sol, double tol, int maxsteps, std::vector<double> *errors = nullptr) const
  {
    need();
    int its = 0;
    std::vector<double> err(size_t(maxsteps) + 2, 0.0);
    b200_check(ngsamg_b200_pcg(amg_mat->Handle(), rhs.FVDouble().Data(), sol.FVDouble().Data(), tol, maxsteps, &its, err.data()));
    if (errors) errors->assign(err.begin(), err.begin() + its + 1);
    return its;
  }

private:
  void need() const { if (!amg_mat) throw ngcore::Exception("B200AMGPC: FinalizeLevel has not been called"); }
  static string num_to_string(double v)
  {
    char buf[64];
    std::snprintf(buf, sizeof(buf), "%.17g", v);
    return buf;
  }
};

// same registered names as the reference (amg_register.hpp:85,97; elasticity.hpp:104-140)
struct H1ScalB200 : B200AMGPC<1> {
  H1ScalB200(shared_ptr<ngcomp::BilinearForm> b, const Flags &f, const string n) : B200AMGPC<1>(b, f, n, "h1_scal") {}
};
struct Elast3dB200 : B200AMGPC<3> {
  Elast3dB200(shared_ptr<ngcomp::BilinearForm> b, const Flags &f, const string n) : B200AMGPC<3>(b, f, n, "elast_3d") {}
};

}  // namespace amg

// one translation unit of the host application expands this (static-initialiser registration, amg_register.hpp:79-98)
#define NGSAMG_B200_REGISTER_PRECONDITIONERS()                                                   \
  static ngcomp::RegisterPreconditioner<amg::H1ScalB200> ngsamg_b200_reg_h1_scal("NgsAMG.h1_scal"); \
  static ngcomp::RegisterPreconditioner<amg::Elast3dB200> ngsamg_b200_reg_elast_3d("NgsAMG.elast_3d")
