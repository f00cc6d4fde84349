// Multi-rank part of oracle/_ref/libngsamg_ref.so: class shells for the reference's hybrid (MPI-parallel) smoother, filled with the
// reference's own function bodies (oracle/_ref/frag/*.inc, cut out at build time) and compiled against the threaded MPI stand-in
// (ngs_standin_mpi.hpp).  TEST INFRASTRUCTURE ONLY.  Included by ref_harness.cpp after the single-rank shells (BaseSmoother, GSS3).
//
// Pinned through this file (tests/test_ref_pin_par.py compares oracle/oracle_par.py with it):
//   BasicDCCMap::CalcDOFMasters, DCCMap::AllocMPIStuff / StartDIS2CO .. ApplyG            dcc_map.cpp:17-302, 494-543
//   DecomposeSparseMatrixHybrid, HybridBaseMatrix::MultAdd / Mult                          hybrid_matrix.cpp:17-307, 393-453
//   MyAllReduceDofData                                                                     mpiwrap_extension.hpp:283-335
//   CalcHybridSmootherRDGItGeneric, CalcHybridSmootherRDG, HybridSmoother::CalcModDiag     hybrid_smoother_utils.hpp:11-178, hybrid_smoother.cpp:70-197
//   GSS3 range sweeps, GSS4, HybridGSSmoother::Finalize / SmoothStageRHS / SmoothStageRes  gssmoother.hpp:47-56,131-140, gssmoother.cpp:416-583, 620-861
//   HybridBaseSmoother::Start/Finish DIS2CO/CO2CU, Smooth, SmoothBack, SmoothImpl, SmoothImplRES, SmoothImplRHS, CallStageKernelsImpl
//                                                                                          hybrid_base_smoother.cpp:111-574
#pragma once

namespace amg {

template <class TM> using SparseMatrixTM = SparseMatrix<TM>;

// ---- utilities of the reference's own tree ----------------------------------------------------------------------------
#include "../_ref/frag/u_find_sorted.inc"
#include "../_ref/frag/u_merge3.inc"
#include "../_ref/frag/u_tabtrait.inc"
;
#include "../_ref/frag/u_merge2.inc"
#include "../_ref/frag/u_allreduce_dofdata.inc"

// ---- DCCMap -------------------------------------------------------------------------------------------------------------
template <class TSCAL> class DCCMap {
public:
  explicit DCCMap(shared_ptr<ParallelDofs> _pardofs) : pardofs(_pardofs), block_size(_pardofs->GetEntrySize()) {}
  virtual ~DCCMap() = default;
  shared_ptr<BitArray> GetMasterDOFs() const { return m_dofs; }
  shared_ptr<ParallelDofs> GetParallelDofs() const { return pardofs; }
  void StartDIS2CO(BaseVector &vec) const;
  void ApplyDIS2CO(BaseVector &vec) const;
  void FinishDIS2CO() const;
  void StartCO2CU(BaseVector &vec) const;
  void ApplyCO2CU(BaseVector &vec) const;
  void FinishCO2CU() const;
  FlatArray<int> GetMDOFs(int kp) const { return m_ex_dofs[kp]; }
  FlatArray<int> GetGDOFs(int kp) const { return g_ex_dofs[kp]; }
  void WaitD2C() const;
  void BufferG(BaseVector &vec) const;
  void ApplyM(BaseVector &vec) const;
  void BufferM(BaseVector &vec) const;
  void ApplyG(BaseVector &vec) const;

protected:
  void AllocMPIStuff();
  shared_ptr<ParallelDofs> pardofs;
  int block_size;
  shared_ptr<BitArray> m_dofs;
  // the request arrays are touched by the const exchange methods (MPI handles are values in the reference)
  mutable Array<NG_MPI_Request> m_reqs, m_send, m_recv, g_reqs, g_send, g_recv;
  Table<int> m_ex_dofs, g_ex_dofs;
  Table<TSCAL> m_buffer, g_buffer;
};
#include "../_ref/frag/dcc_alloc.inc"
#include "../_ref/frag/dcc_start_d2c.inc"
#include "../_ref/frag/dcc_apply_d2c.inc"
#include "../_ref/frag/dcc_finish_d2c.inc"
#include "../_ref/frag/dcc_start_c2c.inc"
#include "../_ref/frag/dcc_apply_c2c.inc"
#include "../_ref/frag/dcc_finish_c2c.inc"
#include "../_ref/frag/dcc_wait_d2c.inc"
#include "../_ref/frag/dcc_iterate.inc"
#include "../_ref/frag/dcc_buffer_g.inc"
#include "../_ref/frag/dcc_apply_m.inc"
#include "../_ref/frag/dcc_buffer_m.inc"
#include "../_ref/frag/dcc_apply_g.inc"

template <class TSCAL> class BasicDCCMap : public DCCMap<TSCAL> {
public:
  explicit BasicDCCMap(shared_ptr<ParallelDofs> _pardofs) : DCCMap<TSCAL>(_pardofs) {
    CalcDOFMasters();
    this->AllocMPIStuff();
  }

protected:
  using DCCMap<TSCAL>::pardofs;
  using DCCMap<TSCAL>::m_dofs;
  using DCCMap<TSCAL>::m_ex_dofs;
  using DCCMap<TSCAL>::g_ex_dofs;
  virtual void CalcDOFMasters();
};
#include "../_ref/frag/dcc_masters.inc"

// ---- hybrid matrix A = M + G ------------------------------------------------------------------------------------------------
#include "../_ref/frag/hyb_decompose.inc"

template <class TSCAL> class HybridBaseMatrix : public BaseMatrix {
public:
  HybridBaseMatrix(shared_ptr<ParallelDofs> parDOFs, shared_ptr<DCCMap<TSCAL>> dCCMap) : _parDOFs(parDOFs), _dCCMap(dCCMap) {}
  shared_ptr<ParallelDofs> GetParallelDofs() const { return _parDOFs; }
  DCCMap<TSCAL> &GetDCCMap() { return *_dCCMap; }
  DCCMap<TSCAL> const &GetDCCMap() const { return *_dCCMap; }
  shared_ptr<BaseMatrix> GetM() const { return _M; }
  shared_ptr<BaseMatrix> GetG() const { return _G; }
  bool HasGLocal() const { return _G != nullptr; }
  bool HasGGlobal() const { return !g_zero; }
  int VHeight() const override { return GetM()->VHeight(); }
  int VWidth() const override { return GetM()->VWidth(); }
  void MultAdd(double s, const BaseVector &x, BaseVector &y) const override;
  void Mult(const BaseVector &x, BaseVector &y) const override;
  virtual size_t EntrySize() const { return 1; }
  shared_ptr<BaseVector> CreateVector() const {
    auto v = make_shared<BaseVector>(Height(), EntrySize());
    v->parallel = true;
    return v;
  }

protected:
  void SetMG(shared_ptr<BaseMatrix> aM, shared_ptr<BaseMatrix> aG) {   // hybrid_matrix.cpp:334-357: is G zero on ALL ranks?
    _M = aM;
    _G = aG;
    int nzg = (_G != nullptr) ? 1 : 0;
    nzg = _parDOFs->GetCommunicator().AllReduce(nzg, NG_MPI_SUM);
    g_zero = (nzg == 0);
  }

private:
  shared_ptr<ParallelDofs> _parDOFs;
  shared_ptr<DCCMap<TSCAL>> _dCCMap;
  shared_ptr<BaseMatrix> _M, _G;
  bool g_zero = true;
};
#include "../_ref/frag/hyb_multadd.inc"
#include "../_ref/frag/hyb_mult.inc"

template <class TM> class HybridMatrix : public HybridBaseMatrix<double> {
public:
  using TSCAL = double;
  HybridMatrix(shared_ptr<SparseMatrix<TM>> A, shared_ptr<ParallelDofs> pds, shared_ptr<DCCMap<double>> dcc) : HybridBaseMatrix<double>(pds, dcc) {
    std::tie(spM, spG) = DecomposeSparseMatrixHybrid<TM>(A, pds, *dcc);
    this->SetMG(spM, spG);
  }
  shared_ptr<SparseMatrix<TM>> GetSpM() const { return spM; }
  shared_ptr<SparseMatrix<TM>> GetSpG() const { return spG; }
  size_t EntrySize() const override { return ngbla::Height<TM>(); }

protected:
  shared_ptr<SparseMatrix<TM>> spM, spG;
};

// ---- modified diagonal -------------------------------------------------------------------------------------------------
#include "../_ref/frag/rdg_generic.inc"
#include "../_ref/frag/rdg.inc"

// ---- hybrid smoothers --------------------------------------------------------------------------------------------------
enum SMOOTHING_DIRECTION : char { FORWARD = 0, BACKWARD = 1 };

class BackgroundMPIThread {   // comm-in-thread is not exercised (the harness always passes commInThread = false)
public:
  void StartInThread(function<void(void)>) { throw Exception("ref harness: BackgroundMPIThread is not part of the pin"); }
  void WaitForThread() { throw Exception("ref harness: BackgroundMPIThread is not part of the pin"); }
};

template <class TSCAL> class HybridBaseSmoother : public BaseSmoother {
public:
  HybridBaseSmoother(shared_ptr<HybridBaseMatrix<TSCAL>> A, int numLocSteps, bool commInThread, bool overlapComm)
      : BaseSmoother(A), _hybridA(A), _numLocSteps(numLocSteps), _commInThread(commInThread), _overlapComm(overlapComm) {
    _stashedGx = A->CreateVector();
  }
  void Smooth(BaseVector &x, BaseVector const &b, BaseVector &res, bool res_updated = false, bool update_res = false, bool x_zero = false) const override;
  void SmoothBack(BaseVector &x, BaseVector const &b, BaseVector &res, bool res_updated = false, bool update_res = false, bool x_zero = false) const override;

protected:
  HybridBaseMatrix<TSCAL> const &GetHybridA() const { return *_hybridA; }
  enum SMOOTH_STAGE : char { LOC_PART_1 = 0, EX_PART = 1, LOC_PART_2 = 2 };
  virtual void SmoothStageRHS(SMOOTH_STAGE const &stage, SMOOTHING_DIRECTION const &direction, BaseVector &x, BaseVector const &b, BaseVector &res, bool const &x_zero) const = 0;
  virtual void SmoothStageRes(SMOOTH_STAGE const &stage, SMOOTHING_DIRECTION const &direction, BaseVector &x, BaseVector const &b, BaseVector &res, bool const &x_zero) const = 0;

private:
  shared_ptr<HybridBaseMatrix<TSCAL>> _hybridA;
  int _numLocSteps;
  bool _commInThread, _overlapComm;
  shared_ptr<BaseVector> _stashedGx;
  shared_ptr<BackgroundMPIThread> NG_MPI_thread;
  template <SMOOTHING_DIRECTION DIR> void SmoothImpl(BaseVector &x, BaseVector const &b, BaseVector &res, bool res_updated, bool update_res, bool x_zero) const;
  template <SMOOTHING_DIRECTION DIR> void SmoothImplRHS(BaseVector &x, BaseVector const &b, BaseVector &res, bool x_zero) const;
  template <SMOOTHING_DIRECTION DIR> void SmoothImplRES(BaseVector &x, BaseVector const &b, BaseVector &res, bool x_zero) const;
  template <SMOOTHING_DIRECTION DIR, bool RES> void CallStageKernelsImpl(BaseVector &x, BaseVector const &b, BaseVector &res, bool const &x_zero, bool const &need_d2c) const;
  void StartDIS2CO(BaseVector &vec) const;
  void FinishDIS2CO(BaseVector &vec) const;
  void StartCO2CU(BaseVector &vec) const;
  void FinishCO2CU(BaseVector &vec) const;
};
#include "../_ref/frag/hbs_start_d2c.inc"
#include "../_ref/frag/hbs_finish_d2c.inc"
#include "../_ref/frag/hbs_start_c2c.inc"
#include "../_ref/frag/hbs_finish_c2c.inc"
#include "../_ref/frag/hbs_impl.inc"
#include "../_ref/frag/hbs_impl_res.inc"
#include "../_ref/frag/hbs_impl_rhs.inc"
#include "../_ref/frag/hbs_stages.inc"
#include "../_ref/frag/hbs_smooth.inc"
#include "../_ref/frag/hbs_smoothback.inc"

template <class TM> class HybridSmoother : public HybridBaseSmoother<double> {
public:
  using TSCAL = double;
  HybridSmoother(shared_ptr<HybridMatrix<TM>> _A, int _numLocSteps, bool _commInThread, bool _overlapComm)
      : HybridBaseSmoother<double>(_A, _numLocSteps, _commInThread, _overlapComm), _hybridSpA(_A) {}

protected:
  Array<TM> CalcModDiag(shared_ptr<BitArray> free);
  HybridMatrix<TM> &GetHybSparseA() { return *_hybridSpA; }
  HybridMatrix<TM> const &GetHybSparseA() const { return *_hybridSpA; }

private:
  shared_ptr<HybridMatrix<TM>> _hybridSpA;
};
#include "../_ref/frag/hyb_calcmoddiag.inc"

template <class TM> class GSS4 {
protected:
  Array<int> xdofs;
  shared_ptr<SparseMatrix<TM>> cA;
  Array<TM> dinv;
  bool pinv = false;

public:
  using TSCAL = double;
  static constexpr int BS() { return ngbla::Height<TM>(); }
  using TV = typename strip_vec<Vec<BS(), TSCAL>>::type;
  GSS4(shared_ptr<SparseMatrix<TM>> A, FlatArray<TM> repl_diag, shared_ptr<BitArray> subset = nullptr, bool _pinv = false);
  FlatArray<int> XDofs() const { return xdofs; }
  FlatArray<TM> DiagInverses() const { return dinv; }

protected:
  void SetUp(shared_ptr<SparseMatrix<TM>> A, shared_ptr<BitArray> subset);
  template <class TLAM> INLINE void iterate_rows(TLAM lam, bool bw) const;
  virtual void SmoothRESInternal(BaseVector &x, BaseVector &res, bool backwards) const;
  virtual void SmoothRHSInternal(BaseVector &x, const BaseVector &b, bool backwards) const;

public:
#include "../_ref/frag/gss4_smooth.inc"
#include "../_ref/frag/gss4_smoothback.inc"
#include "../_ref/frag/gss4_smoothres.inc"
#include "../_ref/frag/gss4_smoothbackres.inc"
};
#include "../_ref/frag/gss4_ctor_repl.inc"
#include "../_ref/frag/gss4_iterate.inc"
#include "../_ref/frag/gss4_setup.inc"
#include "../_ref/frag/gss4_res.inc"
#include "../_ref/frag/gss4_rhs.inc"

template <class TM> class HybridGSSmoother : public HybridSmoother<TM> {
public:
  using TSCAL = double;
  HybridGSSmoother(shared_ptr<HybridMatrix<TM>> _A, shared_ptr<BitArray> _subset, bool _pinv, bool _overlap, bool _in_thread, bool _symm_loc, int _nsteps_loc)
      // argument order as in the reference's ctor (gssmoother.cpp:603-617): (_overlap, _in_thread, _nsteps_loc) land in
      // (numLocSteps, commInThread, overlapComm) -- so there is one local pass and the exchange always overlaps
      : HybridSmoother<TM>(_A, _overlap, _in_thread, _nsteps_loc), subset(_subset), pinv(_pinv), symm_loc(_symm_loc) {}
  virtual void Finalize();
  size_t SplitInd() const { return split_ind; }
  shared_ptr<GSS3<TM>> Loc() const { return jac_loc; }
  shared_ptr<GSS4<TM>> Ex() const { return jac_ex; }

protected:
  using SMOOTH_STAGE = typename HybridBaseSmoother<TSCAL>::SMOOTH_STAGE;
  void SmoothStageRHS(SMOOTH_STAGE const &stage, SMOOTHING_DIRECTION const &direction, BaseVector &x, BaseVector const &b, BaseVector &res, bool const &x_zero) const override;
  void SmoothStageRes(SMOOTH_STAGE const &stage, SMOOTHING_DIRECTION const &direction, BaseVector &x, BaseVector const &b, BaseVector &res, bool const &x_zero) const override;
  shared_ptr<BitArray> subset;
  bool pinv = false;
  bool symm_loc = false;
  size_t split_ind;
  shared_ptr<GSS3<TM>> jac_loc, jac_exo;
  shared_ptr<GSS4<TM>> jac_ex;
};
#include "../_ref/frag/hgs_finalize.inc"
#include "../_ref/frag/hgs_stage_rhs.inc"
#include "../_ref/frag/hgs_stage_res.inc"

// ---- CtrMap: contraction of a distributed level onto its group master (dof_contract.cpp) -------------------------------------
template <class TV> constexpr int VecHeight() { return TV::HEIGHT; }
template <> constexpr int VecHeight<double>() { return 1; }
struct UDofsView {
  shared_ptr<ParallelDofs> pd;
  const NgMPI_Comm &GetCommunicator() const { return pd->GetCommunicator(); }
};
template <class TV> class CtrMap : public BaseDOFMapStep {
public:
  using TM = typename spm_entry<VecHeight<TV>(), VecHeight<TV>()>::type;
  using TSPM = SparseMatrix<TM>;
  using TSPM_TM = SparseMatrix<TM>;
  // group[0] is the master; dof_maps[k][j] = contracted dof of local dof j of group member k (master only)
  CtrMap(shared_ptr<ParallelDofs> originalDofs, shared_ptr<ParallelDofs> mappedDofs, Array<int> &&_group, Table<int> &&_dof_maps)
      : _orig(originalDofs), _mapped(mappedDofs), group(std::move(_group)), master(group[0]), dof_maps(std::move(_dof_maps)) {
    is_gm = (GetUDofs().GetCommunicator().Rank() == master);
  }
  void TransferF2C(const BaseVector *x_fine, BaseVector *x_coarse) const override;
  void AddF2C(double fac, const BaseVector *x_fine, BaseVector *x_coarse) const;
  void TransferC2F(BaseVector *x_fine, const BaseVector *x_coarse) const;
  void AddC2F(double fac, BaseVector *x_fine, const BaseVector *x_coarse) const override;
  bool IsMaster() const { return is_gm; }
  void SetUpMPIStuff();
  shared_ptr<TSPM> DoAssembleMatrix(shared_ptr<TSPM> mat) const;
  UDofsView GetUDofs() const { return UDofsView{_orig}; }
  shared_ptr<ParallelDofs> GetParallelDofs() const { return _orig; }
  shared_ptr<ParallelDofs> GetMappedParDofs() const { return _mapped; }

protected:
  shared_ptr<ParallelDofs> _orig, _mapped;
  Array<int> group;
  int master;
  bool is_gm;
  Table<int> dof_maps;
  mutable Array<NG_MPI_Request> reqs;
  Array<NG_MPI_Datatype> NG_MPI_types;
  mutable Table<TV> buffers;
};
#include "../_ref/frag/ctr_timer_f2c.inc"
#include "../_ref/frag/ctr_timer_c2f.inc"
#include "../_ref/frag/ctr_f2c.inc"
#include "../_ref/frag/ctr_addf2c.inc"
#include "../_ref/frag/ctr_c2f.inc"
#include "../_ref/frag/ctr_addc2f.inc"
#include "../_ref/frag/ctr_setup_mpi.inc"
#include "../_ref/frag/ctr_timer_mat.inc"
#include "../_ref/frag/ctr_assemble.inc"

}  // namespace amg
