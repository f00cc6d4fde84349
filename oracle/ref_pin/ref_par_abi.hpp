// C ABI of the multi-rank part (ref_par_*): R ranks = R host threads around the reference's hybrid smoother.  GLUE ONLY.
// Included at the end of ref_harness.cpp (uses its MatH helpers).
#pragma once
#include <thread>

namespace {
struct ParRank {
  int rank = 0, b = 1;
  i64 n = 0;
  MatH *A = nullptr, *M = nullptr, *G = nullptr;
  shared_ptr<BitArray> free;
  std::vector<int> peers;
  std::vector<std::vector<int>> ex;
  shared_ptr<ParallelDofs> pds;
  shared_ptr<BasicDCCMap<double>> dcc;
  shared_ptr<BaseMatrix> hyb;
  shared_ptr<BaseSmoother> sm;
};
struct ParH {
  int R;
  std::unique_ptr<World> world;
  std::vector<ParRank> rk;
};

// f(rank state) on R threads; the first failure aborts the world (nobody is left waiting for a message that never comes)
template <class H, class F> int run_ranks(H *h, F f) {
  std::vector<std::string> errs(h->R);
  std::vector<std::thread> th;
  for (int r = 0; r < h->R; r++)
    th.emplace_back([&, r] {
      try {
        f(h->rk[r]);
      } catch (const std::exception &e) {
        errs[r] = e.what();
        h->world->abort();
      }
    });
  for (auto &t : th) t.join();
  for (int pass = 0; pass < 2; pass++)     // report the rank that failed first-hand, not the ones that were woken up by the abort
    for (int r = 0; r < h->R; r++)
      if (!errs[r].empty() && (pass == 1 || errs[r].find("another rank failed") == std::string::npos)) { g_err = "rank " + std::to_string(r) + ": " + errs[r]; return 1; }
  return 0;
}

template <int B> void par_setup(ParH *h, ParRank &K, bool overlap, bool symm_loc, int nsteps_loc) {
  typedef typename spm_entry<B, B>::type TM;
  NgMPI_Comm comm(h->world.get(), K.rank);
  Array<int> cnt(K.peers.size()), peers(K.peers.size());
  for (size_t k = 0; k < K.peers.size(); k++) { cnt[k] = int(K.ex[k].size()); peers[k] = K.peers[k]; }
  Table<int> ext(cnt);
  for (size_t k = 0; k < K.peers.size(); k++) for (size_t j = 0; j < K.ex[k].size(); j++) ext[k][j] = K.ex[k][j];
  K.pds = make_shared<ParallelDofs>(comm, (size_t)K.n, B, peers, ext);
  K.dcc = make_shared<BasicDCCMap<double>>(K.pds);                                   // CalcDOFMasters + AllocMPIStuff
  auto hm = make_shared<HybridMatrix<TM>>(as<B, B>(K.A), K.pds, K.dcc);             // DecomposeSparseMatrixHybrid
  auto sm = make_shared<HybridGSSmoother<TM>>(hm, K.free, false, overlap, false, symm_loc, nsteps_loc);
  sm->Finalize();                                                                     // CalcModDiag, loc / ex split, GSS3 + GSS4
  K.hyb = hm;
  K.sm = sm;
  K.M = new MatH{B, B, hm->GetSpM()};
  K.G = hm->GetSpG() ? new MatH{B, B, hm->GetSpG()} : nullptr;
}

template <int B> void par_info(ParRank &K, i64 *split, uint8_t *master, double *dinv) {
  typedef typename spm_entry<B, B>::type TM;
  auto sm = dynamic_pointer_cast<HybridGSSmoother<TM>>(K.sm);
  *split = (i64)sm->SplitInd();
  auto md = K.dcc->GetMasterDOFs();
  for (i64 i = 0; i < K.n; i++) master[i] = md->Test(i) ? 1 : 0;
  auto dl = sm->Loc()->DiagInverses();
  for (size_t i = 0; i < dl.Size(); i++) std::memcpy(dinv + i * B * B, (const void *)&dl[i], sizeof(double) * B * B);
  if (sm->Ex()) {
    auto xd = sm->Ex()->XDofs();
    auto de = sm->Ex()->DiagInverses();
    for (size_t i = 0; i < xd.Size(); i++) std::memcpy(dinv + (size_t)xd[i] * B * B, (const void *)&de[i], sizeof(double) * B * B);
  }
}

inline PARALLEL_STATUS stat_of(int s) { return s == 0 ? DISTRIBUTED : CUMULATED; }
inline void load(BaseVector &v, const double *p, int status) {
  auto fv = v.FVDouble();
  std::memcpy(fv.Data(), p, sizeof(double) * fv.Size());
  v.parallel = true;
  v.SetParallelStatus(stat_of(status));
}
inline void store(const BaseVector &v, double *p) {
  auto fv = v.FVDouble();
  std::memcpy(p, fv.Data(), sizeof(double) * fv.Size());
}
}  // namespace

extern "C" {

void *ref_par_new(int R) {
  ParH *h = new ParH;
  h->R = R;
  h->world.reset(new World(R));
  h->rk.resize(R);
  for (int r = 0; r < R; r++) h->rk[r].rank = r;
  return h;
}

void ref_par_free(void *hv) {
  ParH *h = (ParH *)hv;
  for (auto &K : h->rk) { delete K.A; delete K.M; delete K.G; }
  delete h;
}

// the sub-assembled matrix of rank r and its sharing lists: peers ascending, ex_dofs[ex_ptr[k] .. ex_ptr[k+1]) = dofs shared with peers[k],
// ascending and in the same order on both sides
int ref_par_set_rank(void *hv, int r, i64 n, int b, const i64 *rp, const i32 *ci, const double *v, const uint8_t *freed, int npeers,
                     const i32 *peers, const i64 *ex_ptr, const i32 *ex_dofs) {
  ParH *h = (ParH *)hv;
  return guarded([&] {
    ParRank &K = h->rk[r];
    K.n = n;
    K.b = b;
    K.A = (MatH *)ref_mat_new(n, n, b, b, rp, ci, v);
    if (!K.A) throw Exception(g_err);
    K.free = nullptr;
    if (freed) {
      K.free = make_shared<BitArray>((size_t)n);
      for (i64 i = 0; i < n; i++) if (freed[i]) K.free->SetBit(i);
    }
    K.peers.assign(peers, peers + npeers);
    K.ex.resize(npeers);
    for (int k = 0; k < npeers; k++) K.ex[k].assign(ex_dofs + ex_ptr[k], ex_dofs + ex_ptr[k + 1]);
  });
}

// BasicDCCMap, HybridMatrix (M/G split), HybridGSSmoother::Finalize on every rank
int ref_par_setup(void *hv, int overlap, int symm_loc, int nsteps_loc) {
  ParH *h = (ParH *)hv;
  return run_ranks(h, [&](ParRank &K) {
    if (K.b == 1) par_setup<1>(h, K, overlap != 0, symm_loc != 0, nsteps_loc);
    else if (K.b == 3) par_setup<3>(h, K, overlap != 0, symm_loc != 0, nsteps_loc);
    else if (K.b == 6) par_setup<6>(h, K, overlap != 0, symm_loc != 0, nsteps_loc);
    else throw Exception("ref_par_setup: unsupported block size");
  });
}

const void *ref_par_M(const void *hv, int r) { return ((const ParH *)hv)->rk[r].M; }
const void *ref_par_G(const void *hv, int r) { return ((const ParH *)hv)->rk[r].G; }   // NULL: no G on this rank

// split_ind, master flags, and the inverted (modified) diagonal blocks the local smoothers hold (loc rows: GSS3, ex rows: GSS4, else 0)
int ref_par_info(void *hv, int r, i64 *split, uint8_t *master, double *dinv) {
  ParH *h = (ParH *)hv;
  return guarded([&] {
    ParRank &K = h->rk[r];
    if (K.b == 1) par_info<1>(K, split, master, dinv);
    else if (K.b == 3) par_info<3>(K, split, master, dinv);
    else par_info<6>(K, split, master, dinv);
  });
}

// the m_ex / g_ex lists of rank r towards its k-th neighbour (sizes first: pass NULL)
int ref_par_dcc_lists(void *hv, int r, int k, i64 *nm, i32 *m, i64 *ng, i32 *g) {
  ParH *h = (ParH *)hv;
  return guarded([&] {
    auto md = h->rk[r].dcc->GetMDOFs(k), gd = h->rk[r].dcc->GetGDOFs(k);
    *nm = (i64)md.Size();
    *ng = (i64)gd.Size();
    if (m) for (size_t i = 0; i < md.Size(); i++) m[i] = md[i];
    if (g) for (size_t i = 0; i < gd.Size(); i++) g[i] = gd[i];
  });
}

// one hybrid smoother call on all ranks.  x[r], b[r], res[r]: n_r * b doubles; status: 0 = DISTRIBUTED, 1 = CUMULATED (x must be 1)
int ref_par_smooth(void *hv, double **x, double **b, double **res, int status_b, int status_res, int res_updated, int update_res, int x_zero,
                   int backwards) {
  ParH *h = (ParH *)hv;
  return run_ranks(h, [&](ParRank &K) {
    BaseVector vx(K.n, K.b), vb(K.n, K.b), vr(K.n, K.b);
    load(vx, x[K.rank], 1);
    load(vb, b[K.rank], status_b);
    load(vr, res[K.rank], status_res);
    if (backwards) K.sm->SmoothBack(vx, vb, vr, res_updated != 0, update_res != 0, x_zero != 0);
    else K.sm->Smooth(vx, vb, vr, res_updated != 0, update_res != 0, x_zero != 0);
    if (vx.GetParallelStatus() != CUMULATED) throw Exception("ref_par_smooth: x is not CUMULATED on return");
    store(vx, x[K.rank]);
    store(vr, res[K.rank]);
  });
}

// y = (M + G) x  (HybridBaseMatrix::Mult; x CUMULATED, y DISTRIBUTED)
int ref_par_mult(void *hv, double **x, double **y) {
  ParH *h = (ParH *)hv;
  return run_ranks(h, [&](ParRank &K) {
    BaseVector vx(K.n, K.b), vy(K.n, K.b);
    load(vx, x[K.rank], 1);
    vy.parallel = true;
    K.hyb->Mult(vx, vy);
    store(vy, y[K.rank]);
  });
}

// which = 0: DISTRIBUTED -> CONCENTRATED (StartDIS2CO, ApplyDIS2CO, FinishDIS2CO); 1: CONCENTRATED -> CUMULATED (StartCO2CU, ApplyCO2CU, FinishCO2CU)
int ref_par_exchange(void *hv, int which, double **v) {
  ParH *h = (ParH *)hv;
  return run_ranks(h, [&](ParRank &K) {
    BaseVector vv(K.n, K.b);
    load(vv, v[K.rank], 0);
    if (which == 0) { K.dcc->StartDIS2CO(vv); K.dcc->ApplyDIS2CO(vv); K.dcc->FinishDIS2CO(); }
    else { K.dcc->StartCO2CU(vv); K.dcc->ApplyCO2CU(vv); K.dcc->FinishCO2CU(); }
    store(vv, v[K.rank]);
  });
}
}  // extern "C"

// =====================================================================================================================
// the whole multi-rank preconditioner: AMGMatrix::SmoothV (reference code) over distributed levels with the reference's hybrid
// smoothers and ProlMap transfers, one AMGMatrix per rank, R ranks = R host threads.  The step onto the coarsest, contracted
// level uses the reference's CtrMap (TransferF2C / TransferC2F / DoAssembleMatrix, dof_contract.cpp:49-228, 503-727) for ONE group
// whose master is rank 0; what is GLUE is only where it sits: CtrCoarse below plugs "contract, serial cycle of the nested
// hierarchy on the master, expand" in as the coarse solve, where the reference makes the CtrMap a DOFMap step and lets the other
// ranks drop out of the cycle.
// =====================================================================================================================
namespace {
struct ParLevel {
  int b = 1;
  MatH *A = nullptr, *P = nullptr, *PT = nullptr;
  shared_ptr<BitArray> free;
  std::vector<int> peers;
  std::vector<std::vector<int>> ex;
  shared_ptr<ParallelDofs> pds;
  shared_ptr<BasicDCCMap<double>> dcc;
  shared_ptr<BaseMatrix> hyb;
  shared_ptr<BaseSmoother> sm;
};
struct ParAmgH;
struct ParAmgRank {
  int rank = 0;
  std::vector<ParLevel> lev;
  AMGMatrix amg;
  std::vector<i64> ctr_map;   // local coarsest dof -> dof of the merged level
  shared_ptr<BaseDOFMapStep> ctr;                                   // CtrMap<TV> of the reference
  std::function<void(const BaseVector *, BaseVector *)> ctr_c2f;    // its TransferC2F (not part of BaseDOFMapStep here)
};
struct ParAmgH {
  int R = 0, nlev = 0;
  std::unique_ptr<World> world;
  std::vector<ParAmgRank> rk;
  AmgH *nested = nullptr;     // serial hierarchy on the merged level (owned by the caller), used by rank 0
  i64 n_merged = 0;
  MatH *merged = nullptr;     // CtrMap::DoAssembleMatrix on the master
};

// coarsest level: gather on rank 0 (members added in rank order), serial V-cycle of the nested hierarchy, scatter
class CtrCoarse : public BaseMatrix {
  ParAmgH *h;
  int rank, b;

public:
  CtrCoarse(ParAmgH *ah, int r, int ab) : h(ah), rank(r), b(ab) {}
  int VHeight() const override { return 0; }
  int VWidth() const override { return 0; }
  void MultAdd(double, const BaseVector &, BaseVector &) const override { throw Exception("CtrCoarse: MultAdd"); }
  void Mult(const BaseVector &rhs, BaseVector &x) const override {
    ParAmgRank &K = h->rk[rank];
    if (rank != 0) {
      K.ctr->TransferF2C(&rhs, nullptr);                   // members send their DISTRIBUTED values to the master ...
      K.ctr_c2f(&x, nullptr);                              // ... and receive the CUMULATED correction
    } else {
      const size_t N = (size_t)h->n_merged;
      BaseVector g(N, b), xg(N, b);
      K.ctr->TransferF2C(&rhs, &g);                        // master: zero, add own, add the members' through dof_maps
      h->nested->amg.SmoothV(xg, g);                       // the reference's serial cycle on the merged level
      K.ctr_c2f(&x, &xg);
    }
    if (x.GetParallelStatus() != CUMULATED) throw Exception("CtrCoarse: x is not CUMULATED after TransferC2F");
  }
};

template <int B> void paramg_level_setup(ParAmgH *h, ParAmgRank &K, int l, int sm_steps, bool sm_symm, bool overlap) {
  typedef typename spm_entry<B, B>::type TM;
  ParLevel &L = K.lev[l];
  NgMPI_Comm comm(h->world.get(), K.rank);
  Array<int> cnt(L.peers.size()), peers(L.peers.size());
  for (size_t k = 0; k < L.peers.size(); k++) { cnt[k] = int(L.ex[k].size()); peers[k] = L.peers[k]; }
  Table<int> ext(cnt);
  for (size_t k = 0; k < L.peers.size(); k++) for (size_t j = 0; j < L.ex[k].size(); j++) ext[k][j] = L.ex[k][j];
  L.pds = make_shared<ParallelDofs>(comm, L.A->m->Height(), B, peers, ext);
  L.dcc = make_shared<BasicDCCMap<double>>(L.pds);
  auto hm = make_shared<HybridMatrix<TM>>(as<B, B>(L.A), L.pds, L.dcc);
  auto sm = make_shared<HybridGSSmoother<TM>>(hm, L.free, false, overlap, false, false, 1);
  sm->Finalize();
  L.hyb = hm;
  L.sm = (sm_steps > 1 || sm_symm) ? shared_ptr<BaseSmoother>(make_shared<ProxySmoother>(sm, sm_steps, sm_symm)) : shared_ptr<BaseSmoother>(sm);
}

// CtrMap for the coarsest distributed level: group = all ranks, master = rank 0; DoAssembleMatrix gives the merged matrix there
template <int B> void paramg_ctr_setup(ParAmgH *h, ParAmgRank &K) {
  typedef typename std::conditional<B == 1, double, Vec<B>>::type TV;
  ParLevel &L = K.lev[h->nlev - 1];
  NgMPI_Comm comm(h->world.get(), K.rank);
  Array<int> cnt(L.peers.size()), peers(L.peers.size());
  for (size_t k = 0; k < L.peers.size(); k++) { cnt[k] = int(L.ex[k].size()); peers[k] = L.peers[k]; }
  Table<int> ext(cnt);
  for (size_t k = 0; k < L.peers.size(); k++) for (size_t j = 0; j < L.ex[k].size(); j++) ext[k][j] = L.ex[k][j];
  auto orig = make_shared<ParallelDofs>(comm, L.A->m->Height(), B, peers, ext);
  Array<int> group(h->R);
  for (int r = 0; r < h->R; r++) group[r] = r;
  shared_ptr<ParallelDofs> mapped;
  Table<int> maps;
  if (K.rank == 0) {
    Array<int> none(0);
    Table<int> noex(none);
    mapped = make_shared<ParallelDofs>(comm, (size_t)h->n_merged, B, none, noex);
    Array<int> per(h->R);
    for (int r = 0; r < h->R; r++) per[r] = int(h->rk[r].ctr_map.size());
    maps = Table<int>(per);
    for (int r = 0; r < h->R; r++) for (size_t j = 0; j < h->rk[r].ctr_map.size(); j++) maps[r][j] = int(h->rk[r].ctr_map[j]);
  }
  auto ctr = make_shared<CtrMap<TV>>(orig, mapped, std::move(group), std::move(maps));
  ctr->SetUpMPIStuff();
  auto merged = ctr->DoAssembleMatrix(as<B, B>(L.A));           // collective: members send, the master merges
  if (K.rank == 0) h->merged = new MatH{B, B, merged};
  K.ctr = ctr;
  K.ctr_c2f = [ctr](const BaseVector *xf, BaseVector *xc) { ctr->TransferC2F(const_cast<BaseVector *>(xf), xc); };
}

template <int BF, int BC> void paramg_rap(ParAmgRank &K, int l) {
  ParLevel &L = K.lev[l];
  auto pt = TransposeSPMImpl<BF, BC>(*as<BF, BC>(L.P));
  L.PT = new MatH{BC, BF, pt};
  K.lev[l + 1].A = new MatH{BC, BC, RestrictMatrix<BF, BC>(*pt, *as<BF, BF>(L.A), *as<BF, BC>(L.P))};   // local Galerkin product (DISTRIBUTED sum)
  K.lev[l + 1].b = BC;
  K.amg.map->AddStep(make_shared<ProlMap<typename spm_entry<BF, BC>::type>>(as<BF, BC>(L.P), as<BC, BF>(L.PT)));
}
}  // namespace

extern "C" {

// nlev levels per rank: 0 .. nlev-2 distributed (hybrid smoothers), nlev-1 the contracted coarsest level
void *ref_paramg_new(int R, int nlev) {
  ParAmgH *h = new ParAmgH;
  h->R = R;
  h->nlev = nlev;
  h->world.reset(new World(R));
  h->rk.resize(R);
  for (int r = 0; r < R; r++) {
    h->rk[r].rank = r;
    h->rk[r].lev.resize(nlev);
    h->rk[r].amg.map = make_shared<DOFMap>();
    h->rk[r].amg.n_levels = nlev;
  }
  return h;
}

void ref_paramg_free(void *hv) {
  ParAmgH *h = (ParAmgH *)hv;
  for (auto &K : h->rk) for (auto &L : K.lev) { delete L.A; delete L.P; delete L.PT; }
  delete h->merged;
  delete h;
}

int ref_paramg_set_matrix(void *hv, int r, i64 n, int b, const i64 *rp, const i32 *ci, const double *v, const uint8_t *freed) {
  ParAmgH *h = (ParAmgH *)hv;
  return guarded([&] {
    ParLevel &L = h->rk[r].lev[0];
    L.b = b;
    L.A = (MatH *)ref_mat_new(n, n, b, b, rp, ci, v);
    if (!L.A) throw Exception(g_err);
    if (freed) {
      L.free = make_shared<BitArray>((size_t)n);
      for (i64 i = 0; i < n; i++) if (freed[i]) L.free->SetBit(i);
    }
  });
}

int ref_paramg_set_halo(void *hv, int r, int l, int npeers, const i32 *peers, const i64 *ex_ptr, const i32 *ex_dofs) {
  ParAmgH *h = (ParAmgH *)hv;
  return guarded([&] {
    ParLevel &L = h->rk[r].lev[l];
    L.peers.assign(peers, peers + npeers);
    L.ex.assign(npeers, {});
    for (int k = 0; k < npeers; k++) L.ex[k].assign(ex_dofs + ex_ptr[k], ex_dofs + ex_ptr[k + 1]);
  });
}

// local prolongation of level l on rank r (levels in order); the local coarse matrix comes from the reference's RestrictMatrix
int ref_paramg_set_prol(void *hv, int r, int l, i64 nc, int bc, const i64 *rp, const i32 *ci, const double *v) {
  ParAmgH *h = (ParAmgH *)hv;
  return guarded([&] {
    ParAmgRank &K = h->rk[r];
    ParLevel &L = K.lev[l];
    if (!L.A || l + 1 >= h->nlev) throw Exception("ref_paramg_set_prol: levels must be added in order");
    L.P = (MatH *)ref_mat_new((i64)L.A->m->Height(), nc, L.b, bc, rp, ci, v);
    if (!L.P) throw Exception(g_err);
    bool done = false;
#define X(H, W) if (!done && L.b == H && bc == W) { paramg_rap<H, W>(K, l); done = true; }
    REF_FOR_SHAPES(X)
#undef X
    if (!done) throw Exception("ref_paramg_set_prol: unsupported block shapes");
  });
}

// contraction onto rank 0: map = local coarsest dof -> merged dof of rank r (call for every rank before ref_paramg_setup; the sharing
// lists of the coarsest level must have been given with ref_paramg_set_halo too)
int ref_paramg_set_contraction(void *hv, int r, i64 n, const i64 *map, i64 n_merged) {
  ParAmgH *h = (ParAmgH *)hv;
  return guarded([&] {
    h->rk[r].ctr_map.assign(map, map + n);
    h->n_merged = n_merged;
  });
}

int ref_paramg_setup(void *hv, int sm_steps, int sm_symm, int overlap) {
  ParAmgH *h = (ParAmgH *)hv;
  return run_ranks(h, [&](ParAmgRank &K) {
    AMGMatrix &M = K.amg;
    M.smoothers.SetSize(h->nlev - 1);
    M.x_level.SetSize(h->nlev); M.rhs_level.SetSize(h->nlev); M.res_level.SetSize(h->nlev);
    for (int l = 0; l < h->nlev; l++) {
      ParLevel &L = K.lev[l];
      if (!L.A) throw Exception("ref_paramg_setup: level matrix missing");
      const size_t n = L.A->m->Height();
      for (auto *arr : {&M.x_level, &M.rhs_level, &M.res_level}) {
        (*arr)[l] = make_shared<BaseVector>(n, L.b);
        (*arr)[l]->parallel = true;
      }
      if (l + 1 == h->nlev) break;
      if (L.b == 1) paramg_level_setup<1>(h, K, l, sm_steps, sm_symm != 0, overlap != 0);
      else if (L.b == 3) paramg_level_setup<3>(h, K, l, sm_steps, sm_symm != 0, overlap != 0);
      else if (L.b == 6) paramg_level_setup<6>(h, K, l, sm_steps, sm_symm != 0, overlap != 0);
      else throw Exception("ref_paramg_setup: unsupported block size");
      M.smoothers[l] = L.sm;
    }
    if (h->n_merged > 0) {
      const int bc = K.lev[h->nlev - 1].b;
      if (bc == 1) paramg_ctr_setup<1>(h, K);
      else if (bc == 3) paramg_ctr_setup<3>(h, K);
      else if (bc == 6) paramg_ctr_setup<6>(h, K);
      else throw Exception("ref_paramg_setup: unsupported block size on the contracted level");
    }
  });
}

// CtrMap::DoAssembleMatrix result (valid after ref_paramg_setup; lives on the master)
const void *ref_paramg_merged_matrix(void *hv) { return ((ParAmgH *)hv)->merged; }

// the serial hierarchy on the merged level (a finalized ref_amg_* handle, owned by the caller): "contract, V-cycle there, expand" becomes
// the coarse solve of every rank's AMGMatrix
int ref_paramg_set_nested(void *hv, void *nested) {
  ParAmgH *h = (ParAmgH *)hv;
  return guarded([&] {
    if (!h->merged) throw Exception("ref_paramg_set_nested: no contraction was set up");
    h->nested = (AmgH *)nested;
    for (auto &K : h->rk) {
      K.amg.crs_inv = make_shared<CtrCoarse>(h, K.rank, K.lev[h->nlev - 1].b);
      K.amg.has_crs_inv = true;
    }
  });
}

// x = C b: b[r] DISTRIBUTED local vectors, x[r] CUMULATED on return (AMGMatrix::SmoothV on every rank)
int ref_paramg_apply(void *hv, double **b, double **x) {
  ParAmgH *h = (ParAmgH *)hv;
  return run_ranks(h, [&](ParAmgRank &K) {
    ParLevel &L = K.lev[0];
    const size_t n = L.A->m->Height();
    BaseVector vx(n, L.b), vb(n, L.b);
    vx.parallel = true;
    load(vb, b[K.rank], 0);
    K.amg.SmoothV(vx, vb);
    if (vx.GetParallelStatus() != CUMULATED) throw Exception("ref_paramg_apply: x is not CUMULATED on return");
    store(vx, x[K.rank]);
  });
}

// y = (M + G) x on level 0 (HybridBaseMatrix::Mult), for the CG around the cycle
int ref_paramg_mult(void *hv, double **x, double **y) {
  ParAmgH *h = (ParAmgH *)hv;
  return run_ranks(h, [&](ParAmgRank &K) {
    ParLevel &L = K.lev[0];
    const size_t n = L.A->m->Height();
    BaseVector vx(n, L.b), vy(n, L.b);
    load(vx, x[K.rank], 1);
    vy.parallel = true;
    L.hyb->Mult(vx, vy);
    store(vy, y[K.rank]);
  });
}

// which: 0 = x_level, 1 = rhs_level, 2 = res_level of the last cycle
int ref_paramg_level_vec(void *hv, int r, int which, int l, double *out) {
  ParAmgH *h = (ParAmgH *)hv;
  return guarded([&] {
    const AMGMatrix &M = h->rk[r].amg;
    const auto &arr = which == 0 ? M.x_level : which == 1 ? M.rhs_level : M.res_level;
    auto fv = arr[l]->FVDouble();
    std::memcpy(out, fv.Data(), sizeof(double) * fv.Size());
  });
}

const void *ref_paramg_level_matrix(void *hv, int r, int l) { return ((ParAmgH *)hv)->rk[r].lev[l].A; }

i64 ref_paramg_level_size(void *hv, int r, int l) {
  ParAmgH *h = (ParAmgH *)hv;
  ParLevel &L = h->rk[r].lev[l];
  return L.A ? (i64)L.A->m->Height() * L.b : -1;
}
}  // extern "C"
