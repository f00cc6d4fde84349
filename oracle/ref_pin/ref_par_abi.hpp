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

// f(rank) on R threads; the first failure aborts the world (nobody is left waiting for a message that never comes)
template <class F> int run_ranks(ParH *h, F f) {
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
  for (int r = 0; r < h->R; r++)
    if (!errs[r].empty() && errs[r].find("another rank failed") == std::string::npos) { g_err = "rank " + std::to_string(r) + ": " + errs[r]; return 1; }
  for (int r = 0; r < h->R; r++)
    if (!errs[r].empty()) { g_err = errs[r]; return 1; }
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
