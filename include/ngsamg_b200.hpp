// ngsamg_b200.hpp -- C++ host-side mirror of the reference's operator interface on top of the C ABI (ngsamg_b200.h).
//
// The reference is compiled C++ (an NGSolve add-on); NGSolve itself is absent from this image, so this header mirrors the
// interface of the path with plain std:: types -- same class / method names, argument meaning and error behaviour
// (exceptions, like ngcore::Exception) -- so that code written against the reference reads the same:
//
//   amg::SparseMat A(n, n, 1, 1, rowptr, col, val);                 // ngla::SparseMatrix<TM> layout
//   amg::BaseAMGPC pc("NgsAMG.h1_scal", A, freedofs, {{"ngs_amg_max_coarse_size", "20"}});   // amg_pc.hpp:26-228
//   pc.FinalizeLevel();                                             // amg_pc.cpp:413-434
//   pc.Mult(b, x);  pc.MultAdd(s, b, x);                            // amg_matrix.cpp:377-393
//   amg::CGSolver cg(A, pc, 100, 1e-8);  cg.Solve(rhs, sol);        // ngsolve.krylovspace.CGSolver, tests/h1/amg_utils.py:346
//   amg::ParallelAMGPC ppc("NgsAMG.h1_scal", Aloc, pardofs, comm, freedofs);                // the same on a ParallelMatrix (one rank = one GPU)
//   auto H = amg::DecomposeHybrid(Aloc, pardofs, comm);              // HybridMatrix: M, G, modified diagonal (host only)
//
// Header-only; link with -lngsamg_b200.  No torch, no NGSolve.
#pragma once
#include <cstdint>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "ngsamg_b200.h"

namespace amg {

struct Exception : std::runtime_error {   // stands in for ngcore::Exception (amg_pc.cpp:430, 446)
  explicit Exception(const std::string &m) : std::runtime_error(m) {}
};

inline void check(int rc)
{
  if (rc != 0) throw Exception(ngsamg_b200_last_error());
}

// ngla::SparseMatrix<Mat<H,W,double>>: firsti / colnr (ascending per row) / row-major blocks
class SparseMat {
public:
  int64_t nrows = 0, ncols = 0;
  int bh = 1, bw = 1;
  std::vector<int64_t> rowptr;
  std::vector<int32_t> col;
  std::vector<double> val;
  SparseMat() = default;
  SparseMat(int64_t h, int64_t w, int abh, int abw, std::vector<int64_t> rp, std::vector<int32_t> ci, std::vector<double> v)
    : nrows(h), ncols(w), bh(abh), bw(abw), rowptr(std::move(rp)), col(std::move(ci)), val(std::move(v))
  {
    if ((int64_t)rowptr.size() != nrows + 1 || val.size() != col.size() * (size_t)bh * bw) throw Exception("SparseMat: inconsistent arrays");
  }
  int64_t Height() const { return nrows; }
  int64_t Width() const { return ncols; }
  int64_t NZE() const { return rowptr.empty() ? 0 : rowptr.back(); }
  ngsamg_csr abi() const { return ngsamg_csr{nrows, ncols, bh, bw, rowptr.data(), col.data(), val.data()}; }
};

namespace detail {
inline SparseMat fetch(ngsamg_b200_spm *h, int64_t nrows, int64_t ncols, int bh, int bw, int64_t nnz)
{
  SparseMat C;
  C.nrows = nrows; C.ncols = ncols; C.bh = bh; C.bw = bw;
  C.rowptr.resize(nrows + 1); C.col.resize(nnz); C.val.resize((size_t)nnz * bh * bw);
  check(ngsamg_b200_spm_fetch(h, C.rowptr.data(), C.col.data(), C.val.data()));
  return C;
}
}  // namespace detail

// MatMultABImpl (src/base/linalg/utils_sparseMM.cpp:107-238)
inline SparseMat MatMultAB(const SparseMat &A, const SparseMat &B, int device = 0)
{
  ngsamg_b200_spm *h = nullptr; int64_t nr = 0, nnz = 0;
  ngsamg_csr a = A.abi(), b = B.abi();
  check(ngsamg_b200_matmul_begin(&a, &b, device, &h, &nr, &nnz));
  return detail::fetch(h, nr, B.ncols, A.bh, B.bw, nnz);
}
// TransposeSPMImpl (utils_sparseMM.cpp:54-93)
inline SparseMat TransposeSPM(const SparseMat &A, int device = 0)
{
  ngsamg_b200_spm *h = nullptr; int64_t nr = 0, nnz = 0;
  ngsamg_csr a = A.abi();
  check(ngsamg_b200_transpose_begin(&a, device, &h, &nr, &nnz));
  return detail::fetch(h, nr, A.nrows, A.bw, A.bh, nnz);
}
// RestrictMatrix (utils_sparseMM.hpp:93-109): (P^T A) P
inline SparseMat RestrictMatrix(const SparseMat &A, const SparseMat &P, int device = 0)
{
  ngsamg_b200_spm *h = nullptr; int64_t nr = 0, nnz = 0;
  ngsamg_csr a = A.abi(), p = P.abi();
  check(ngsamg_b200_rap_begin(&a, &p, device, &h, &nr, &nnz));
  return detail::fetch(h, nr, P.ncols, P.bw, P.bw, nnz);
}

using Flags = std::map<std::string, std::string>;   // NGSolve Flags: "ngs_amg_<key>" -> value (lists comma separated)

class BaseAMGPC;

// BaseSmoother (src/base/smoothers/base_smoother.hpp:43-156): the smoother of one level
class BaseSmoother {
  ngsamg_b200_t *h; int level;
  friend class BaseAMGPC;
  BaseSmoother(ngsamg_b200_t *ah, int l) : h(ah), level(l) {}
public:
  void Smooth(double *x, const double *b, double *res, bool res_updated = false, bool update_res = true, bool x_zero = false) const
  { check(ngsamg_b200_smooth(h, level, x, b, res, res_updated, update_res, x_zero, 0)); }
  void SmoothBack(double *x, const double *b, double *res, bool res_updated = false, bool update_res = true, bool x_zero = false) const
  { check(ngsamg_b200_smooth(h, level, x, b, res, res_updated, update_res, x_zero, 1)); }
};

// BaseAMGPC (src/base/precond/amg_pc.hpp:26-228) in strict-algebraic mode + the AMGMatrix it owns (amg_matrix.hpp:14-87)
class BaseAMGPC {
  ngsamg_b200_t *h = nullptr;
  int64_t n = 0;
  bool finalized = false;
protected:
  BaseAMGPC() = default;                                   // for ParallelAMGPC, which creates the handle itself
  void adopt(ngsamg_b200_t *ah, int64_t an) { h = ah; n = an; }
public:
  BaseAMGPC(const std::string &type, const SparseMat &A, const std::vector<uint8_t> *freedofs = nullptr, const Flags &flags = {},
            const std::vector<double> *vertex_xyz = nullptr, int device = 0)
  {
    std::vector<const char *> k, v;
    for (auto &kv : flags) { k.push_back(kv.first.c_str()); v.push_back(kv.second.c_str()); }
    ngsamg_csr a = A.abi();
    check(ngsamg_b200_create(type.c_str(), &a, freedofs ? freedofs->data() : nullptr, vertex_xyz ? vertex_xyz->data() : nullptr,
                             k.data(), v.data(), (int)k.size(), device, &h));
    n = A.nrows * A.bh;
  }
  BaseAMGPC(const BaseAMGPC &) = delete;
  BaseAMGPC &operator=(const BaseAMGPC &) = delete;
  ~BaseAMGPC() { ngsamg_b200_destroy(h); }

  // AMGMatrix(DOFMap([ProlMap...]), ...) analogue (src/base/solve/python_solve.cpp:57-76)
  void SetProlongations(const std::vector<SparseMat> &P)
  {
    std::vector<ngsamg_csr> a;
    for (auto &p : P) a.push_back(p.abi());
    check(ngsamg_b200_set_prolongations(h, (int)a.size(), a.data()));
  }
  void InitLevel() {}                                   // freedofs are bound at construction
  void FinalizeLevel() { if (!finalized) { check(ngsamg_b200_finalize(h)); finalized = true; } }

  // BaseMatrix quartet (amg_matrix.cpp:377-393); host or device pointers
  void Mult(const double *b, double *x) const { check(ngsamg_b200_apply(h, b, x)); }
  void MultAdd(double s, const double *b, double *x) const { check(ngsamg_b200_apply_add(h, s, b, x)); }
  void MultTrans(const double *b, double *x) const { Mult(b, x); }
  void MultTransAdd(double s, const double *b, double *x) const { MultAdd(s, b, x); }
  int64_t VHeight() const { return n; }
  int64_t VWidth() const { return n; }
  bool IsComplex() const { return false; }

  // introspection (python_amg.hpp:30-101)
  size_t GetNLevels(int = 0) const { return (size_t)ngsamg_b200_num_levels(h); }
  std::pair<int64_t, int> GetNDof(int level, int = 0) const
  {
    ngsamg_level_info i; check(ngsamg_b200_level_info(h, level, &i)); return {i.n, i.b};
  }
  double GetOC() const { return ngsamg_b200_operator_complexity(h); }
  // AMGMatrix::GetOC (amg_matrix.cpp:551-582): [OC, OC_l0, OC_l1, ...]
  std::vector<double> GetOCs() const {
    std::vector<double> occs((size_t)ngsamg_b200_operator_complexities(h, nullptr, 0));
    if (!occs.empty()) ngsamg_b200_operator_complexities(h, occs.data(), (int)occs.size());
    return occs;
  }
  BaseSmoother GetSmoother(int level) const
  {
    if (level + 1 >= (int)GetNLevels()) throw Exception("only have " + std::to_string(GetNLevels() - 1) + " smoothers");
    return BaseSmoother(h, level);
  }
  SparseMat GetLevelMatrix(int level) const
  {
    ngsamg_level_info i; check(ngsamg_b200_level_info(h, level, &i));
    SparseMat M; M.nrows = M.ncols = i.n; M.bh = M.bw = i.b;
    M.rowptr.resize(i.n + 1); M.col.resize(i.nnz); M.val.resize((size_t)i.nnz * i.b * i.b);
    check(ngsamg_b200_get_level_matrix(h, level, M.rowptr.data(), M.col.data(), M.val.data()));
    return M;
  }
  SparseMat GetProlongation(int level) const            // ProlMap::GetProl
  {
    ngsamg_level_info i; check(ngsamg_b200_level_info(h, level, &i));
    SparseMat P; P.nrows = i.n; P.ncols = i.ncoarse; P.bh = i.b; P.bw = i.bcoarse;
    P.rowptr.resize(i.n + 1); P.col.resize(i.nnz_prol); P.val.resize((size_t)i.nnz_prol * i.b * i.bcoarse);
    check(ngsamg_b200_get_prolongation(h, level, P.rowptr.data(), P.col.data(), P.val.data()));
    return P;
  }
  // ProlMap::TransferF2C / AddC2F (dof_map.cpp:633-709)
  void TransferF2C(int level, const double *x_fine, double *x_coarse) const { check(ngsamg_b200_restrict(h, level, x_fine, x_coarse)); }
  void AddC2F(int level, double fac, double *x_fine, const double *x_coarse) const { check(ngsamg_b200_prolong_add(h, level, fac, x_coarse, x_fine)); }
  ngsamg_b200_t *handle() const { return h; }
};

// ---- multi-rank (MPI) mirror ---------------------------------------------------------------------------------------
// ngla::ParallelDofs as the path uses it (dcc_map.cpp:494-543, hybrid_matrix.cpp:27-33): neighbour ranks + shared DOFs
class ParallelDofs {
public:
  int64_t ndof = 0;
  std::vector<int32_t> procs;                  // GetDistantProcs(), ascending
  std::vector<std::vector<int32_t>> exdofs;    // GetExchangeDofs(procs[k]), ascending, pairwise consistent
  ParallelDofs() = default;
  ParallelDofs(int64_t n, std::vector<int32_t> p, std::vector<std::vector<int32_t>> e) : ndof(n), procs(std::move(p)), exdofs(std::move(e))
  {
    if (procs.size() != exdofs.size()) throw Exception("ParallelDofs: one exchange list per distant proc");
  }
  const std::vector<int32_t> &GetDistantProcs() const { return procs; }
  const std::vector<int32_t> &GetExchangeDofs(int proc) const
  {
    for (size_t k = 0; k < procs.size(); k++) if (procs[k] == proc) return exdofs[k];
    throw Exception("ParallelDofs: not a distant proc");
  }
  int64_t GetNDofLocal() const { return ndof; }
};

// NgMPI_Comm stand-in: the two host callbacks of ngsamg_comm behind a C++ interface (an NGSolve adapter implements them with
// ISend/IRecv/WaitAll and AllReduce) + the optional NCCL communicator for the device data path
class Communicator {
public:
  virtual ~Communicator() = default;
  virtual int Rank() const = 0;
  virtual int Size() const = 0;
  virtual void Exchange(int npeers, const int32_t *peers, const void *const *sendbuf, const int64_t *sendbytes, void *const *recvbuf,
                        const int64_t *recvbytes) = 0;
  virtual void AllReduceSum(double *vals, int n) = 0;
  void *nccl = nullptr;
  ngsamg_comm abi()
  {
    ngsamg_comm c;
    c.rank = Rank(); c.size = Size(); c.ctx = this; c.nccl = nccl;
    c.exchange = [](void *ctx, int32_t np, const int32_t *pr, const void *const *sb, const int64_t *sn, void *const *rb, const int64_t *rn) -> int {
      try { static_cast<Communicator *>(ctx)->Exchange(np, pr, sb, sn, rb, rn); return 0; } catch (...) { return 1; }
    };
    c.allreduce_sum = [](void *ctx, double *v, int32_t n) -> int {
      try { static_cast<Communicator *>(ctx)->AllReduceSum(v, n); return 0; } catch (...) { return 1; }
    };
    return c;
  }
};

namespace detail {
struct HaloArrays {
  std::vector<int64_t> ptr{0};
  std::vector<int32_t> dofs;
  ngsamg_halo halo;
  explicit HaloArrays(const ParallelDofs &pd)
  {
    for (auto &l : pd.exdofs) { dofs.insert(dofs.end(), l.begin(), l.end()); ptr.push_back((int64_t)dofs.size()); }
    halo = ngsamg_halo{(int32_t)pd.procs.size(), pd.procs.data(), ptr.data(), dofs.data()};
  }
};
}  // namespace detail

// HybridMatrix (src/base/linalg/hybrid_matrix.hpp): A = M + G, plus what HybridGSSmoother::Finalize derives from it.
// Host only (DecomposeSparseMatrixHybrid hybrid_matrix.cpp:17-307, CalcHybridSmootherRDG hybrid_smoother_utils.hpp:146-176).
struct HybridMatrix {
  SparseMat M, G;
  std::vector<double> mod_diag;      // replacement diagonal of the hybrid smoother
  std::vector<int32_t> sweep_rank;   // position of every row in the stage order LOC_PART_1 | EX_PART | LOC_PART_2
  std::vector<uint8_t> master;       // DCCMap::GetMasterDOFs
  const SparseMat &GetM() const { return M; }
  const SparseMat &GetG() const { return G; }
};
inline HybridMatrix DecomposeHybrid(const SparseMat &A, const ParallelDofs &pd, Communicator &comm, const std::vector<uint8_t> *freedofs = nullptr)
{
  detail::HaloArrays ha(pd);
  ngsamg_comm c = comm.abi();
  ngsamg_csr a = A.abi();
  ngsamg_b200_hybrid_host *h = nullptr;
  int64_t nm = 0, ng = 0;
  check(ngsamg_b200_hybrid_host_begin(&a, freedofs ? freedofs->data() : nullptr, &ha.halo, &c, &h, &nm, &ng));
  HybridMatrix H;
  const size_t bs = (size_t)A.bh * A.bw;
  for (SparseMat *m : {&H.M, &H.G}) { m->nrows = m->ncols = A.nrows; m->bh = A.bh; m->bw = A.bw; m->rowptr.resize(A.nrows + 1); }
  H.M.col.resize(nm); H.M.val.resize(nm * bs); H.G.col.resize(ng); H.G.val.resize(ng * bs);
  H.mod_diag.resize(A.nrows * bs); H.sweep_rank.resize(A.nrows); H.master.resize(A.nrows);
  check(ngsamg_b200_hybrid_host_fetch(h, H.M.rowptr.data(), H.M.col.data(), H.M.val.data(), H.G.rowptr.data(), H.G.col.data(), H.G.val.data(),
                                      H.mod_diag.data(), H.sweep_rank.data(), H.master.data()));
  return H;
}

// BaseAMGPC on a ParallelMatrix: A = the rank's sub-assembled local matrix, pd = its ParallelDofs.  Mult / CGSolver::Solve are
// collective; b is DISTRIBUTED, x CUMULATED (amg_matrix.cpp:160-307).
class ParallelAMGPC : public BaseAMGPC {
public:
  ParallelAMGPC(const std::string &type, const SparseMat &A, const ParallelDofs &pd, Communicator &comm,
                const std::vector<uint8_t> *freedofs = nullptr, const Flags &flags = {}, const std::vector<double> *vertex_xyz = nullptr,
                int device = 0)
    : BaseAMGPC()
  {
    std::vector<const char *> k, v;
    for (auto &kv : flags) { k.push_back(kv.first.c_str()); v.push_back(kv.second.c_str()); }
    detail::HaloArrays ha(pd);
    cabi = comm.abi();      // must outlive the handle
    ngsamg_csr a = A.abi();
    ngsamg_b200_t *hh = nullptr;
    check(ngsamg_b200_create_parallel(type.c_str(), &a, freedofs ? freedofs->data() : nullptr, vertex_xyz ? vertex_xyz->data() : nullptr,
                                      &ha.halo, &cabi, k.data(), v.data(), (int)k.size(), device, &hh));
    adopt(hh, A.nrows * A.bh);
  }
  int GetNParallelLevels() const { return ngsamg_b200_num_parallel_levels(handle()); }
private:
  ngsamg_comm cabi;
};

// ngsolve.krylovspace.CGSolver(mat, pre, maxsteps, tol) as the reference's tests use it (tests/h1/amg_utils.py:346-362)
class CGSolver {
  const BaseAMGPC &pre; int maxsteps; double tol;
public:
  int iterations = 0;
  std::vector<double> errors;
  CGSolver(const SparseMat &, const BaseAMGPC &apre, int amaxsteps = 100, double atol = 1e-12) : pre(apre), maxsteps(amaxsteps), tol(atol) {}
  void Solve(const double *rhs, double *sol)
  {
    errors.assign(maxsteps + 2, 0.0);
    check(ngsamg_b200_pcg(pre.handle(), rhs, sol, tol, maxsteps, &iterations, errors.data()));
    errors.resize(iterations + 1);
  }
};

}  // namespace amg
