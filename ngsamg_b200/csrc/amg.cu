// amg.cu -- the hierarchy object behind the C ABI: setup (layouts, RAP, smoothers, coarse inverse),
// V-cycle (AMGMatrix::SmoothV, src/base/solve/amg_matrix.cpp:160-307) and PCG on one B200.
#include <cub/cub.cuh>

#include <atomic>
#include <functional>
#include <chrono>
#include <memory>
#include <mutex>

#include <dlfcn.h>
#include <unistd.h>
#include <nccl.h>   // types only: the library is dlopen'ed (libnccl.so.2), nothing links against it

#include "../../include/ngsamg_b200.h"
#include "device.hpp"
#include "kernels.cuh"
#include "kernels_tile.cuh"
#include "kernels_ctile.cuh"
#include "kernels_itile.cuh"
#include "kernels_rm.cuh"
#include "kernels_p2p.cuh"
#include "par.hpp"
#include "tiles.hpp"

namespace ngb {

namespace {

constexpr int TILE_ROWS = 256;  // rows per triangular-sweep tile == CTA size
constexpr int TB = 256;

inline unsigned nblk(i64 n, int tb = TB) { return (unsigned)((n + tb - 1) / tb); }
inline i64 round32(i64 n) { return (n + 31) / 32 * 32; }

struct Sell {
  i64 nrows_pad = 0, nslices = 0, total_slots = 0, nnz = 0;
  int bh = 1, bw = 1;
  i64 maxw = 0;          // slice width covering 97 % of the slices (0 = unknown): picks the row walk of the scalar SpMV (launch_spmv)
  i64 *slice_ptr = nullptr;
  i32 *col = nullptr;
  double *val = nullptr;
  SellView view() const { return SellView{slice_ptr, col, val}; }
  i64 bytes() const { return nnz * (8 * bh * bw + 4); }
  void release() { dev_free(slice_ptr); dev_free(col); dev_free(val); }
};

// row-major copy of a triangle (small levels, kernels_rm.cuh)
struct Rm {
  i64 *ptr = nullptr;
  i32 *col = nullptr;
  double *val = nullptr;
  i32 *gate = nullptr;
  RmView view() const { return RmView{ptr, col, val, gate}; }
  void release() { dev_free(ptr); dev_free(col); dev_free(val); dev_free(gate); ptr = nullptr; col = nullptr; val = nullptr; gate = nullptr; }
};

enum { SM_GS = 0, SM_JACOBI = 1, SM_BGS = 2 };

struct Level {
  i64 n = 0, npad = 0;
  int b = 1;
  i64 nnz = 0;
  HostBsr hA;  // plain matrix, original numbering (kept for introspection when small enough)
  bool keep_host = true;
  std::vector<uint8_t> free_mask;  // empty = all free
  std::vector<i32> perm;           // original -> level-scheduled (padded) row
  std::vector<i32> sweep_rank;     // position of each row in the Gauss-Seidel sweep; empty = the row number (reference order)
  i32 *d_perm = nullptr;
  uint8_t *d_freep = nullptr;
  Sell L, U, N;   // strictly lower / strictly upper / free-row couplings to non-free rows
  Rm rmL, rmU;    // row-major copies of L / U for the warp-per-row sweep of small levels (empty otherwise)
  // ---- block Gauss-Seidel (sm_type = bgs; BSmoother2, loc_block_gssmoother_impl.hpp): blocks = the aggregates of the next coarse map
  std::vector<i32> gs_block;        // vertex -> block (= coarse vertex), -1 = in no block (not smoothed)
  std::unique_ptr<Level> bgs;       // shadow level: layout of A~ = DB^-1 A (unit diagonal, no couplings inside a block) in THIS level's numbering
  Sell DBI;                         // DB^-1: the inverted dense diagonal blocks as a block-diagonal sparse matrix
  double *diag = nullptr, *dinv = nullptr;
  // Gauss-Seidel dependency structure
  int depth = 0;        // number of dependency levels (length of the critical path of a sweep)
  i64 nonfree_pad = 0;  // rows of dependency level 0 (non-free rows), padded
  int pre_l = 8, pre_u = 8;  // register slot cache of the triangular sweeps (scalar matrices)
  std::vector<i64> level_start;  // first (padded) row of every dependency level, plus npad
  i32 *d_bnd_fwd = nullptr, *d_bnd_bwd = nullptr;  // split-gate boundaries per slice (see k_gs_tri)
  // transfer to level+1
  HostBsr hP;
  Sell P, PT;
  i32 *d_pt_rowmap = nullptr;  // PT is stored with its rows sorted by length: storage row -> coarse (level-scheduled) row
  i64 nc = 0;
  int bc = 1;
  // work vectors, level-scheduled numbering, npad*b doubles
  double *x = nullptr, *y = nullptr, *rhs = nullptr, *res = nullptr, *tmp = nullptr, *wa = nullptr, *wb = nullptr;
  double *result = nullptr;  // where the level's solution lives after a V-cycle (x or y)
  int sm_type = SM_GS, sm_steps = 1;
  bool sm_symm = false, pinv = false;
  double omega = 1.0;
  std::vector<double> xyz;
  int *d_err = nullptr;
  // ---- multi-rank (hybrid) level: smoothing runs on M = master x master block, G is applied additively
  bool par = false;
  ParDofs pd;
  HostBsr hM, hG;                  // hybrid split in the level's local numbering (kept for introspection on small levels)
  std::vector<double> mod_diag;    // replacement diagonal of the hybrid smoother (AoS, local numbering)
  std::vector<uint8_t> gs_mask;    // rows smoothed on this rank: master & free
  Sell G;
  i64 nnz_m = 0, nnz_g = 0;
  // DCC halo lists (level-scheduled numbering), flattened per neighbour
  std::vector<i32> peers;
  std::vector<i64> m_off, g_off;   // npeers + 1 offsets (dofs) into the flattened lists / exchange buffers
  i32 *d_m_idx = nullptr, *d_g_idx = nullptr, *d_mu_dof = nullptr;
  i64 *d_mu_ptr = nullptr, *d_mu_pos = nullptr;
  i64 n_mu = 0;
  double *sendbuf = nullptr, *recvbuf = nullptr, *h_send = nullptr, *h_recv = nullptr;
  // peer-memory halo exchange (kernels_p2p.cuh): one IPC-exported arena per level [recv 2 x cap doubles | flag | ack | cnt | seq]
  bool p2p = false;
  unsigned char *p2p_arena = nullptr;
  std::vector<void *> p2p_opened;    // arenas of the neighbours mapped with cudaIpcOpenMemHandle
  P2PPeer *d_p2p_peer = nullptr;
  i64 *d_m_off = nullptr, *d_g_off = nullptr;
  P2PView p2p_view{};
  const std::vector<uint8_t> &mask() const { return par ? gs_mask : free_mask; }
  // ---- experimental two-level (tile) schedule of the triangular sweeps (flag b200_tile_sweep, tiles.hpp / kernels_tile.cuh)
  bool tiled = false;
  i64 ntiles = 0;
  int tile_maxs = 1;
  i32 *d_tile_slice = nullptr, *d_tile_nlev = nullptr, *d_tile_pred = nullptr, *d_tile_succ = nullptr;
  uint8_t *d_row_lvl = nullptr;
  i64 *d_tile_pred_ptr = nullptr, *d_tile_succ_ptr = nullptr;
  int *d_tile_done = nullptr;
  // CTA-per-tile sweep (kernels_ctile.cuh; tile capacity >= 128 rows): slab capacity, launch geometry
  std::vector<i32> h_tile_slice, h_tile_nlev, h_tile_nreal;   // nreal: rows of the tile that are not padding
  std::vector<i64> h_pred_ptr, h_succ_ptr;
  CTileMeta *d_meta_fwd = nullptr, *d_meta_bwd = nullptr;
  // prepared tile images (kernels_itile.cuh): the production path when no row has more than 7 entries per triangle
  bool itile = false;
  unsigned char *d_img[2] = {nullptr, nullptr};      // [forward (L, with diag), backward (U)]
  ITileMeta *d_imeta[2] = {nullptr, nullptr};
  int itile_cap = 0;
  size_t itile_smem = 0;
  int itile_grid[2] = {0, 0};
  i64 img_bytes_total = 0;
  int tile_nbuf = 1;
  int tile_cap_slots = 0;
  int ctile_grid[2] = {0, 0};     // [add_self]
  size_t ctile_smem = 0;
};

// libnccl.so.2 entry points, resolved at run time
struct NcclApi {
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
};
NcclApi &nccl_api()
{
  static NcclApi api;
  static bool loaded = false;
  static std::mutex mu;
  std::lock_guard<std::mutex> guard(mu);
  if (loaded) return api;
  void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);   // the copy the host application already loaded, if any
  if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!h) throw Error(std::string("cannot load libnccl.so.2: ") + dlerror());
  auto sym = [&](const char *n) { void *f = dlsym(h, n); if (!f) throw Error(std::string("libnccl: missing symbol ") + n); return f; };
  api.GetUniqueId = (decltype(api.GetUniqueId))sym("ncclGetUniqueId");
  api.CommInitRank = (decltype(api.CommInitRank))sym("ncclCommInitRank");
  api.CommDestroy = (decltype(api.CommDestroy))sym("ncclCommDestroy");
  api.Send = (decltype(api.Send))sym("ncclSend");
  api.Recv = (decltype(api.Recv))sym("ncclRecv");
  api.AllReduce = (decltype(api.AllReduce))sym("ncclAllReduce");
  api.GroupStart = (decltype(api.GroupStart))sym("ncclGroupStart");
  api.GroupEnd = (decltype(api.GroupEnd))sym("ncclGroupEnd");
  api.GetErrorString = (decltype(api.GetErrorString))sym("ncclGetErrorString");
  loaded = true;
  return api;
}
#define NGB_NCCL(call)                                                                                                   \
  do {                                                                                                                   \
    ncclResult_t r__ = (call);                                                                                           \
    if (r__ != ncclSuccess) throw ::ngb::Error(std::string("NCCL error: ") + nccl_api().GetErrorString(r__) + " (" #call ")"); \
  } while (0)

}  // namespace

// per-phase device timing of ONE eagerly executed V-cycle (ngsamg_b200_apply_phases): event pairs around every call of a category
enum { PH_TRI = 0, PH_PASS, PH_TRANSFER, PH_EXCHANGE, PH_GSPMV, PH_COARSE, PH_OTHER, PH_COUNT };
struct PhaseTimer {
  cudaStream_t st = nullptr;
  std::vector<cudaEvent_t> ev;
  std::vector<int> cat;
  size_t used = 0;
  int depth = 0;
  void begin(int c)
  {
    if (depth++ > 0) return;                       // nested scopes belong to the outermost category
    if (used + 2 > ev.size()) { ev.resize(used + 2); cudaEventCreate(&ev[used]); cudaEventCreate(&ev[used + 1]); }
    cat.push_back(c);
    cudaEventRecord(ev[used], st);
  }
  void end()
  {
    if (--depth > 0) return;
    cudaEventRecord(ev[used + 1], st);
    used += 2;
  }
  void collect(double *ms)
  {
    for (int c = 0; c < PH_COUNT; c++) ms[c] = 0.0;
    for (size_t k = 0; k * 2 < used; k++) {
      float t = 0;
      cudaEventElapsedTime(&t, ev[2 * k], ev[2 * k + 1]);
      ms[cat[k]] += t;
    }
  }
  ~PhaseTimer() { for (auto e : ev) cudaEventDestroy(e); }
};

struct Amg {
  std::string type;
  Flags flags;
  int device = 0;
  cudaStream_t st = nullptr;
  int num_sms = 148;
  std::vector<std::unique_ptr<Level>> lev;
  std::vector<HostBsr> injected;
  bool finalized = false;
  // coarsest exact solve
  double *d_cinv = nullptr;
  int cinv_n = 0;
  bool has_cinv = false;
  int regularize_coarse_dim = 0;   // 2 / 3: RegularizeMatrix on the coarsest diagonal blocks (elasticity + regularize_cmats), 0: off
  // V-cycle graph
  cudaGraphExec_t vgraph = nullptr;
  i64 vgraph_launches = 0;
  bool use_graph = true;
  // pcg vectors (level 0, permuted)
  double *cg_u = nullptr, *cg_s = nullptr, *cg_q = nullptr, *d_dot = nullptr, *d_partial = nullptr;
  // host staging (pinned)
  double *pin_a = nullptr, *pin_b = nullptr;
  i64 pin_n = 0;
  double *io_a = nullptr, *io_b = nullptr, *io_c = nullptr;  // device staging in original numbering
  static constexpr i64 IO_CHUNK = 1 << 23;                   // doubles per chunk of the pipelined host <-> device staging (64 MB)
  std::vector<cudaEvent_t> io_ev;
  i64 io_n = 0;
  i64 launches = 0;
  double ms_apply = 0, ms_pcg = 0, ms_setup = 0, ms_rap = 0, ms_host = 0, bytes_rap = 0;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  int tri_grid_cap[48] = {0};
  i64 tri_small_rows = 1000000;
  i64 rm_spmv_rows = 150000;      // block levels with at most this many rows run the parallel halves with a warp per row on the row-major copies
  int tri_block_warp_rows = 1;    // block matrices (3x3, 6x6): warp-per-row sweep on every level size
  int tri_rm_rows_per_warp = 8;   // grid of the row-major sweep: at least this many rows per warp
  i64 tri_rm_max_rows = 100000;   // larger small levels keep the SELL warp-per-row sweep (measured: 0.22 vs 0.29 ms at 519 k rows)
  i64 tri_rm_gate_rows = 4096;    // levels with more rows gate every row on its newest dependency before the per-lane polls
  int tri_rm = 1;                 // small levels sweep a row-major copy of the triangle (k_gs_tri_rm) instead of the SELL layout
  int tri_level_pdl = 1;          // colours of the per-colour launches overlap through programmatic dependent launch
  int tri_level_launch_depth = 24;
  i64 tri_level_launch_rows = 131072;
  double tri_gate_gap_levels = 0.0;
  unsigned tri_repoll_ns = 0;
  int tri_pollmode = 0;
  int itile_minb = 7;             // CTAs per SM the 256-row tile-image kernel is compiled for (6: no spills, 7: 72 registers)
  i64 spmv_small_rows = 200000;   // levels with fewer rows use the warp-per-row SpMV
  int tri_regate = 1;
  int tri_split = 0;
  PhaseTimer *pt = nullptr;          // non-null only inside ngsamg_b200_apply_phases
  unsigned long long *tri_trace = nullptr;  // debug tracing of the sync-free sweep (NGSAMG_B200_TRACE_FILE)
  int *d_err = nullptr;
  void check_watchdog();
  // ---- multi-rank
  bool par = false;
  bool owns_stream = true;
  Comm comm;
  ncclComm_t nccl = nullptr;
  int npar = 0;                       // levels 0..npar-1 are distributed (hybrid smoothers), level npar is contracted onto rank 0
  std::unique_ptr<Amg> nested;        // rank 0: the serial hierarchy below the contracted level
  Contraction ctr;                    // rank 0: dof maps of the contraction
  std::vector<i32 *> d_ctr_map;       // rank 0: device copies of ctr.dof_maps
  std::vector<i64> ctr_off;           // rank 0: offsets (doubles) of the ranks' segments in the gather buffers
  double *ctr_buf = nullptr, *h_ctr = nullptr;
  i64 exchanges = 0;
  void finalize_parallel();
  void build_halo(Level &L);
  void setup_p2p(Level &L);           // collective: exchange the IPC handles of the receive arenas of one distributed level
  bool halo_p2p = false;
  void dev_exchange(const std::vector<i32> &peers, const double *sendbuf, const std::vector<i64> &soff, double *recvbuf,
                    const std::vector<i64> &roff, double *h_send, double *h_recv);
  void dis2co(Level &L, double *v);   // DCCMap::StartDIS2CO + ApplyDIS2CO: ghost values travel to the master and are added there
  void co2cu(Level &L, double *v);    // DCCMap::StartCO2CU + ApplyCO2CU: master values overwrite the ghosts
  void contracted_solve(Level &L);
  void allreduce_scalars(double *h, int n);
  void hybrid_smooth_res(Level &L, bool backward, double *x, double *res, bool x_zero);
  void hybrid_smooth_rhs(Level &L, bool backward, double *x, const double *b, double *work, bool x_zero);
  void hybrid_smooth(Level &L, double *x, const double *b, double *res, bool ru, bool ur, bool xz, bool backward);
  void hybrid_level_smooth(Level &L, double *x, const double *b, double *res, bool ru, bool ur, bool xz, bool backward);

  ~Amg();
  void finalize();
  void build_level_layout(Level &L, const DevCsr &dA);
  void build_rm(const Sell &S, Rm &R, bool upper, i64 nonfree);
  void setup_bgs(Level &L, const DevCsr &dA);   // block Gauss-Seidel: DB^-1, A~ = DB^-1 A, shadow level; fixes the level's numbering
  void bgs_res(Level &L, bool backward, double *x, double *res);
  void bgs_rhs(Level &L, bool backward, double *x, const double *b);
  // shallow dependency DAG with many rows per level: one plain launch per level (k_gs_level)
  // (scalar levels only: a colour of a 6x6 level is a few thousand rows of ~25 blocks each -- too few threads with too long chains for a
  // thread-per-row launch; measured 9.6 ms per V-cycle on a 258 k-row level against ~3 ms with a warp per row on the row-major copy)
  bool level_launch(const Level &L) const
  {
    if (L.b > 1 && tri_block_warp_rows) return false;
    return L.depth <= tri_level_launch_depth && L.npad > tri_level_launch_rows && (int)L.level_start.size() == L.depth + 1;
  }
  void prepare_ctile(Level &L);
  bool prepare_itile(Level &L);
  bool setup_tiles(Level &L, int l, const HostBsr &A);
  void build_transfer_layout(Level &F, Level &C);
  void build_coarse_inverse(Level &L);
  void alloc_vectors(Level &L);
  // device primitives on level-scheduled vectors
  template <int B> void tri(Level &L, bool backward, bool add_self, bool write_r, const double *rin, const double *self, double *out, double *rout);
  void tri_dispatch(Level &L, bool backward, bool add_self, bool write_r, const double *rin, const double *self, double *out, double *rout);
  unsigned tri_sleep_ns = 100;
  int tri_prepoll = 1;
  int tri_gate_all = 1;
  int tri_ctas_per_sm = 0;
  void spmv_part(Level &L, int which /*0 L,1 U,2 L+D,3 U+D,4 full*/, const double *v, const double *y_in, double *y_out,
                 double alpha, double beta, double *xadd);
  void transfer(const Sell &S, const double *v, const double *y_in, double *y_out, double alpha, double beta, const i32 *rowmap = nullptr);
  void calc_residuum(Level &L, const double *x, const double *b, double *res, bool x_zero);
  void gs_res(Level &L, bool backward, double *x, double *res, bool x_zero);
  void gs_rhs(Level &L, bool backward, const double *x, const double *b, double *xout);
  void smooth_once(Level &L, double *x, const double *b, double *res, bool ru, bool ur, bool xz, bool backward);
  void level_smooth(Level &L, double *x, const double *b, double *res, bool ru, bool ur, bool xz, bool backward);
  void vcycle_record();
  void vcycle();  // rhs = lev[0].rhs -> x = lev[0].x
  enum { CYCLE_V = 0, CYCLE_W = 1, CYCLE_BS = 2 };
  int cycle = CYCLE_V;
  void coarse_solve_record();
  void restrict_record(int l, const double *res);
  void w_visit(int l);
  void v_from_level(int s, bool ru, bool ur, bool xz);
  void bs_record();
  double dot(i64 n, const double *a, const double *b);
  // io helpers
  void ensure_io(i64 n);
  const double *to_device(const double *p, i64 n, double *stage);
  bool is_device_ptr(const void *p);
  void from_device(double *dst, const double *src_dev, i64 n);
};

namespace {

template <class T>
T *upload_vec(const std::vector<T> &v, cudaStream_t st)
{
  T *d = dev_alloc<T>(v.size());
  if (!v.empty()) NGB_CUDA(cudaMemcpyAsync(d, v.data(), sizeof(T) * v.size(), cudaMemcpyHostToDevice, st));
  NGB_CUDA(cudaStreamSynchronize(st));
  return d;
}

// Gauss-Seidel dependency levels in the reference's row order (gssmoother.cpp:195-315): non-free rows form
// level 0 (they are never updated), a free row sits one level above its deepest free lower neighbour.
void level_schedule(const HostBsr &A, const std::vector<uint8_t> &free_mask, bool smoothed, Level &L, cudaStream_t st)
{
  const i64 n = A.nrows;
  std::vector<i32> lvl(n, 0);
  i32 depth = 1;
  const bool hf = !free_mask.empty();
  // sweep order: the reference's (row number) unless a multicolour order was requested for this level
  const bool natural = L.sweep_rank.empty();
  auto before = [&](i32 j, i32 i) { return natural ? (j < i) : (L.sweep_rank[j] < L.sweep_rank[i]); };
  if (smoothed) {
    std::vector<i32> inv;
    if (!natural) { inv.resize(n); for (i64 i = 0; i < n; i++) inv[L.sweep_rank[i]] = (i32)i; }
    for (i64 q = 0; q < n; q++) {
      const i64 i = natural ? q : inv[q];
      if (hf && !free_mask[i]) { lvl[i] = 0; continue; }
      i32 m = 0;
      for (i64 k = A.rowptr[i]; k < A.rowptr[i + 1]; k++) {
        const i32 j = A.col[k];
        if (natural && j >= i) break;  // columns ascending
        if (j == i || !before(j, (i32)i)) continue;
        if (hf && !free_mask[j]) continue;
        m = std::max(m, lvl[j]);
      }
      lvl[i] = m + 1;
      depth = std::max(depth, m + 2);
    }
  }
  std::vector<i64> cnt(depth, 0);
  for (i64 i = 0; i < n; i++) cnt[lvl[i]]++;
  // drop an empty level 0 (no Dirichlet rows)
  int shift = (smoothed && cnt[0] == 0 && depth > 1) ? 1 : 0;
  std::vector<i64> start(depth + 1, 0);
  for (int l = shift; l < depth; l++) start[l + 1] = start[l] + round32(cnt[l]);
  L.npad = std::max<i64>(start[depth], 32);
  L.depth = depth - shift;
  L.nonfree_pad = (shift == 0 && smoothed) ? round32(cnt[0]) : 0;
  L.level_start.assign(start.begin() + shift, start.end());
  L.perm.resize(n);
  if (!smoothed) {
    std::vector<i64> pos(start.begin(), start.end() - 1);
    for (i64 i = 0; i < n; i++) L.perm[i] = (i32)(pos[lvl[i]]++);
  } else {
    // Rows of one dependency level are mutually independent, so their order inside the level is free.  Group rows with the
    // same (lower, upper) entry counts -- slices (32 consecutive rows) then have uniform widths: less SELL padding and all lanes
    // of a warp reach their newest dependency in the same chunk -- and keep the original order inside a group (locality).
    // Counting sort on (level, lower count, upper count); stable in the row number.
    const int W = depth > 256 ? 32 : 128;
    const bool deep = depth > 64;
    std::vector<i32> sub(n);
    parallel_for(n, [&](i64 lo, i64 hi) {
      for (i64 i = lo; i < hi; i++) {
        int nl = 0, nu = 0;
        for (i64 k = A.rowptr[i]; k < A.rowptr[i + 1]; k++) {
          const i32 j = A.col[k];
          if (j == i) continue;
          if (hf && free_mask[i] && !free_mask[j]) continue;   // stored in the separate N part
          if (lvl[j] < lvl[i] || (lvl[j] == lvl[i] && before(j, (i32)i))) nl++; else nu++;
        }
        // deep DAGs (natural orderings): keep the row order (spatially compact slices keep the wavefront regions decoupled) and
        // only move rows that are wider than the register slot cache into slices of their own
        if (deep) { nl = nl <= 8 ? 0 : nl; nu = nu <= 8 ? 0 : nu; }
        sub[i] = std::min(nl, W - 1) * W + std::min(nu, W - 1);
      }
    });
    const i64 nb = (i64)depth * W * W;
    std::vector<i64> bstart(nb + 1, 0);
    for (i64 i = 0; i < n; i++) bstart[(i64)lvl[i] * W * W + sub[i] + 1]++;
    // bucket offsets: levels start at their padded positions, buckets inside a level are contiguous
    {
      i64 run = 0;
      for (int l = 0; l < depth; l++) {
        run = start[l];
        for (i64 b = (i64)l * W * W; b < (i64)(l + 1) * W * W; b++) { const i64 c = bstart[b + 1]; bstart[b + 1] = 0; bstart[b] = run; run += c; }
      }
    }
    for (i64 i = 0; i < n; i++) L.perm[i] = (i32)(bstart[(i64)lvl[i] * W * W + sub[i]]++);
  }
  if (smoothed && L.b == 1) {
    const i64 ns = L.npad / 32;
    const int nl = (int)L.level_start.size() - 1;
    std::vector<i32> bf(ns, 0), bb(ns, 0);
    for (int lv = 0; lv < nl; lv++) {
      const i32 f = (i32)(lv >= 1 ? L.level_start[lv - 1] : 0);
      const i32 e = (i32)(lv + 2 <= nl ? L.level_start[lv + 2] : L.npad);
      for (i64 s2 = L.level_start[lv] / 32; s2 < L.level_start[lv + 1] / 32; s2++) { bf[s2] = f; bb[s2] = e; }
    }
    L.d_bnd_fwd = upload_vec(bf, st);
    L.d_bnd_bwd = upload_vec(bb, st);
  }
  (void)st;
}

// depth of the dependency DAG of the sequential sweep (what level_schedule would report), without building anything
int sweep_depth(const HostBsr &A, const std::vector<uint8_t> &free_mask, const std::vector<i32> &sweep_rank)
{
  const i64 n = A.nrows;
  std::vector<i32> lvl(n, 0);
  const bool hf = !free_mask.empty(), natural = sweep_rank.empty();
  std::vector<i32> inv;
  if (!natural) { inv.resize(n); for (i64 i = 0; i < n; i++) inv[sweep_rank[i]] = (i32)i; }
  i32 depth = 0;
  for (i64 q = 0; q < n; q++) {
    const i64 i = natural ? q : inv[q];
    if (hf && !free_mask[i]) continue;
    i32 m = 0;
    for (i64 k = A.rowptr[i]; k < A.rowptr[i + 1]; k++) {
      const i32 j = A.col[k];
      if (natural && j >= i) break;
      if (j == i || (!natural && sweep_rank[j] >= sweep_rank[i]) || (hf && !free_mask[j])) continue;
      m = std::max(m, lvl[j]);
    }
    lvl[i] = m + 1;
    depth = std::max(depth, m + 1);
  }
  return depth;
}

void build_sell(i64 nrows_pad, int bh, int bw, const i32 *d_len, Sell &S, cudaStream_t st, i64 *launches)
{
  S.nrows_pad = nrows_pad; S.nslices = nrows_pad / 32; S.bh = bh; S.bw = bw;
  i64 *width = dev_alloc<i64>(S.nslices + 1);
  NGB_CUDA(cudaMemsetAsync(width, 0, sizeof(i64) * (S.nslices + 1), st));
  k_layout_width<<<nblk(S.nslices * 32), TB, 0, st>>>(S.nslices, d_len, width);
  S.slice_ptr = dev_alloc<i64>(S.nslices + 1);
  size_t tb = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, tb, width, S.slice_ptr, S.nslices + 1, st);
  void *tmp = dev_alloc<char>(tb);
  cub::DeviceScan::ExclusiveSum(tmp, tb, width, S.slice_ptr, S.nslices + 1, st);
  NGB_CUDA(cudaMemcpyAsync(&S.total_slots, S.slice_ptr + S.nslices, sizeof(i64), cudaMemcpyDeviceToHost, st));
  NGB_CUDA(cudaStreamSynchronize(st));
  cudaFree(tmp);
  dev_free(width);
  S.col = dev_alloc<i32>(S.total_slots * 32);
  S.val = dev_alloc<double>(S.total_slots * 32 * bh * bw);
  NGB_CUDA(cudaMemsetAsync(S.col, 0xFF, sizeof(i32) * std::max<i64>(S.total_slots * 32, 1), st));
  NGB_CUDA(cudaMemsetAsync(S.val, 0, sizeof(double) * std::max<i64>(S.total_slots * 32 * bh * bw, 1), st));
  if (launches) *launches += 2;
}

// width that covers 97 % of the slices (the masked walks handle wider slices correctly, just with more blocks)
i64 sell_max_width(const Sell &S, cudaStream_t st)
{
  std::vector<i64> sp(S.nslices + 1);
  NGB_CUDA(cudaMemcpyAsync(sp.data(), S.slice_ptr, sizeof(i64) * (S.nslices + 1), cudaMemcpyDeviceToHost, st));
  NGB_CUDA(cudaStreamSynchronize(st));
  std::vector<i64> hist(66, 0);
  for (i64 q = 0; q < S.nslices; q++) hist[std::min<i64>(sp[q + 1] - sp[q], 65)]++;
  i64 acc = 0;
  for (int w = 0; w < 66; w++) {
    acc += hist[w];
    if (acc >= 0.97 * (double)S.nslices) return std::max(w, 1);
  }
  return 65;
}

// dense inverse of an SPD-ish matrix by Gauss-Jordan with partial pivoting (host, setup only)
bool dense_invert(int n, std::vector<double> &a)
{
  std::vector<double> inv((size_t)n * n, 0.0);
  for (int i = 0; i < n; i++) inv[(size_t)i * n + i] = 1.0;
  for (int c = 0; c < n; c++) {
    int p = c;
    double best = std::fabs(a[(size_t)c * n + c]);
    for (int r = c + 1; r < n; r++)
      if (std::fabs(a[(size_t)r * n + c]) > best) { best = std::fabs(a[(size_t)r * n + c]); p = r; }
    if (best == 0.0) return false;
    if (p != c)
      for (int q = 0; q < n; q++) { std::swap(a[(size_t)c * n + q], a[(size_t)p * n + q]); std::swap(inv[(size_t)c * n + q], inv[(size_t)p * n + q]); }
    const double pv = 1.0 / a[(size_t)c * n + c];
    for (int q = 0; q < n; q++) { a[(size_t)c * n + q] *= pv; inv[(size_t)c * n + q] *= pv; }
    parallel_for(n, [&](i64 lo, i64 hi) {
      for (i64 r = lo; r < hi; r++) {
        if (r == c) continue;
        const double f = a[(size_t)r * n + c];
        if (f == 0.0) continue;
        for (int q = 0; q < n; q++) { a[(size_t)r * n + q] -= f * a[(size_t)c * n + q]; inv[(size_t)r * n + q] -= f * inv[(size_t)c * n + q]; }
      }
    }, 64);
  }
  a.swap(inv);
  return true;
}

// block_pinv (pseudo-inverse of a diagonal block): dense.cpp, declared in common.hpp

// ---- coarse-level numbering -------------------------------------------------------------------------
// The numbering of a COARSE level is an output of the hierarchy builder (the reference's comes from its agglomeration
// order).  We number coarse vertices colour-major from a greedy colouring of the Galerkin matrix graph, so that the
// sequential Gauss-Seidel sweep in that numbering has a dependency depth equal to the number of colours instead of
// O(n^(1/3)).  The sweep itself stays "rows in increasing number", exactly what GSS3 does.
// `fixed` (optional): rows that keep their number (the shared dofs of a distributed level: their relative order must stay the
// same on every sharer and the exchange lists ascending, the ParallelDofs contract of dcc_map.cpp:494-543); the other rows are dealt
// colour-major into the remaining numbers.
void greedy_coloring_perm(const HostBsr &A, std::vector<i32> &perm, int &ncolors, const std::vector<uint8_t> *fixed = nullptr,
                          std::vector<i32> *color_out = nullptr)
{
  const i64 n = A.nrows;
  std::vector<i32> color(n, -1);
  std::vector<i64> mark(1024, -1);
  ncolors = 0;
  for (i64 i = 0; i < n; i++) {
    for (i64 k = A.rowptr[i]; k < A.rowptr[i + 1]; k++) {
      const i32 j = A.col[k];
      if (j != i && color[j] >= 0) {
        if ((size_t)color[j] >= mark.size()) mark.resize(2 * color[j] + 2, -1);
        mark[color[j]] = i;
      }
    }
    i32 c = 0;
    while ((size_t)c < mark.size() && mark[c] == i) c++;
    if ((size_t)c >= mark.size()) mark.resize(2 * c + 2, -1);
    color[i] = c;
    ncolors = std::max(ncolors, c + 1);
  }
  // colour-major; inside a colour by the number of lower-coloured neighbours (uniform row lengths per slice), then by index
  std::vector<uint64_t> key(n);
  for (i64 i = 0; i < n; i++) {
    unsigned low = 0, up = 0;
    for (i64 k = A.rowptr[i]; k < A.rowptr[i + 1]; k++) { low += (color[A.col[k]] < color[i]); up += (color[A.col[k]] > color[i]); }
    key[i] = ((uint64_t)color[i] << 40) | ((uint64_t)std::min(low, 1048575u) << 20) | std::min(up, 1048575u);
  }
  std::vector<i32> order(n);
  for (i64 i = 0; i < n; i++) order[i] = (i32)i;
  std::stable_sort(order.begin(), order.end(), [&](i32 a, i32 b) { return key[a] < key[b]; });
  perm.resize(n);
  if (!fixed) {
    for (i64 q = 0; q < n; q++) perm[order[q]] = (i32)q;
    return;
  }
  std::vector<i32> slots;
  slots.reserve(n);
  for (i64 i = 0; i < n; i++) if (!(*fixed)[i]) slots.push_back((i32)i);
  size_t q = 0;
  for (i32 o : order) {
    if ((*fixed)[o]) perm[o] = o;
    else perm[o] = slots[q++];
  }
  if (color_out) color_out->swap(color);
}

// Numbering of a DISTRIBUTED coarse level.  Interior dofs: colour-major (as on one rank).  Shared dofs must keep the ParallelDofs contract
// (dcc_map.cpp:494-543: exchange lists ascending, k-th entry here == k-th entry on the neighbour), so they stay inside the block of their
// sharing class -- the blocks keep their places -- and are ordered INSIDE the block by the colour their MASTER gave them (sent to the
// other sharers once at setup), ties by the old, canonical order: the same order on every sharer, and the master's EX stage
// (hybrid_base_smoother.cpp:508-573) sweeps class after class, colour after colour instead of along the canonical numbering of a whole
// interface (dependency depth ~800 on a 3.7 M-row level).
void parallel_coloring_perm(const Comm &comm, const ParDofs &pd, const HostBsr &A, std::vector<i32> &perm, int &ncolors)
{
  const i64 n = A.nrows;
  std::vector<uint8_t> keep(n, 0);
  for (i64 i = 0; i < n; i++) keep[i] = pd.eqc[i] != 0;
  std::vector<i32> color;
  greedy_coloring_perm(A, perm, ncolors, &keep, &color);          // interior done; shared: identity so far
  std::vector<double> mc(n, 0.0);
  for (i64 i = 0; i < n; i++) if (keep[i] && pd.is_master(i)) mc[i] = (double)color[i];
  allreduce_dof_data(comm, pd, 1, mc);                            // every sharer learns the master's colour (the others contribute 0)
  // members of every class in old order = the slots of the class block
  const size_t ncls = pd.sharers.size();
  std::vector<std::vector<i32>> members(ncls);
  for (i64 i = 0; i < n; i++) if (keep[i]) members[pd.eqc[i]].push_back((i32)i);
  for (size_t c = 1; c < ncls; c++) {
    const std::vector<i32> &slots = members[c];
    std::vector<i32> ord(slots);
    std::stable_sort(ord.begin(), ord.end(), [&](i32 x, i32 y) { return mc[x] < mc[y]; });
    for (size_t k = 0; k < ord.size(); k++) perm[ord[k]] = slots[k];
  }
}

// B = Pi A Pi^T (rows and columns renumbered old -> perm[old]), columns re-sorted ascending
void permute_symmetric(HostBsr &A, const std::vector<i32> &perm)
{
  const i64 n = A.nrows;
  const int bs = A.bs();
  std::vector<i32> inv(n);
  for (i64 i = 0; i < n; i++) inv[perm[i]] = (i32)i;
  HostBsr B;
  B.nrows = A.nrows; B.ncols = A.ncols; B.bh = A.bh; B.bw = A.bw;
  B.rowptr.assign(n + 1, 0);
  for (i64 r = 0; r < n; r++) B.rowptr[r + 1] = B.rowptr[r] + (A.rowptr[inv[r] + 1] - A.rowptr[inv[r]]);
  B.col.resize(A.nnz());
  B.val.resize(A.nnz() * bs);
  parallel_for(n, [&](i64 lo, i64 hi) {
    std::vector<std::pair<i32, i64>> ent;
    for (i64 r = lo; r < hi; r++) {
      const i64 o = inv[r];
      ent.clear();
      for (i64 k = A.rowptr[o]; k < A.rowptr[o + 1]; k++) ent.emplace_back(perm[A.col[k]], k);
      std::sort(ent.begin(), ent.end());
      i64 p = B.rowptr[r];
      for (auto &e : ent) {
        B.col[p] = e.first;
        std::memcpy(&B.val[p * bs], &A.val[e.second * bs], sizeof(double) * bs);
        p++;
      }
    }
  }, 1024);
  A = std::move(B);
}

// P <- P Pi^T (columns renumbered), rows re-sorted
void renumber_columns(HostBsr &P, const std::vector<i32> &perm)
{
  const int bs = P.bs();
  parallel_for(P.nrows, [&](i64 lo, i64 hi) {
    std::vector<std::pair<i32, i64>> ent;
    std::vector<double> tmp;
    for (i64 r = lo; r < hi; r++) {
      const i64 b0 = P.rowptr[r], b1 = P.rowptr[r + 1];
      if (b1 == b0) continue;
      ent.clear();
      for (i64 k = b0; k < b1; k++) ent.emplace_back(perm[P.col[k]], k);
      std::sort(ent.begin(), ent.end());
      tmp.assign(P.val.begin() + b0 * bs, P.val.begin() + b1 * bs);
      for (size_t q = 0; q < ent.size(); q++) {
        P.col[b0 + q] = ent[q].first;
        std::memcpy(&P.val[(b0 + q) * bs], &tmp[(ent[q].second - b0) * bs], sizeof(double) * bs);
      }
    }
  }, 4096);
}

}  // namespace

struct PhaseScope {
  PhaseTimer *t;
  PhaseScope(Amg &a, int cat) : t(a.pt) { if (t) t->begin(cat); }
  ~PhaseScope() { if (t) t->end(); }
};
#define NGB_PHASE(cat) PhaseScope phase_scope_##cat(*this, cat)

Amg::~Amg()
{
  if (device >= 0) cudaSetDevice(device);
  std::function<void(Level &)> free_level = [&](Level &L) {
    L.G.release();
    dev_free(L.d_tile_slice); dev_free(L.d_tile_nlev); dev_free(L.d_tile_pred); dev_free(L.d_tile_succ); dev_free(L.d_row_lvl);
    dev_free(L.d_tile_pred_ptr); dev_free(L.d_tile_succ_ptr); dev_free(L.d_tile_done); dev_free(L.d_meta_fwd); dev_free(L.d_meta_bwd);
    for (int d = 0; d < 2; d++) { dev_free(L.d_img[d]); dev_free(L.d_imeta[d]); }
    dev_free(L.d_m_idx); dev_free(L.d_g_idx); dev_free(L.d_mu_dof); dev_free(L.d_mu_ptr); dev_free(L.d_mu_pos);
    dev_free(L.sendbuf); dev_free(L.recvbuf);
    for (void *q : L.p2p_opened) cudaIpcCloseMemHandle(q);
    dev_free(L.p2p_arena); dev_free(L.d_p2p_peer); dev_free(L.d_m_off); dev_free(L.d_g_off);
    if (L.h_send) cudaFreeHost(L.h_send);
    if (L.h_recv) cudaFreeHost(L.h_recv);
    dev_free(L.d_perm); dev_free(L.d_freep); dev_free(L.d_pt_rowmap); dev_free(L.d_bnd_fwd); dev_free(L.d_bnd_bwd);
    L.L.release(); L.U.release(); L.N.release(); L.P.release(); L.PT.release(); L.rmL.release(); L.rmU.release();
    dev_free(L.diag); dev_free(L.dinv);
    dev_free(L.x); dev_free(L.y); dev_free(L.rhs); dev_free(L.res); dev_free(L.tmp); dev_free(L.wa); dev_free(L.wb);
    L.DBI.release();
    if (L.bgs) free_level(*L.bgs);
  };
  for (auto &lp : lev) free_level(*lp);
  dev_free(d_cinv); dev_free(cg_u); dev_free(cg_s); dev_free(cg_q); dev_free(d_dot); dev_free(d_partial);
  dev_free(io_a); dev_free(io_b); dev_free(io_c); dev_free(d_err);
  if (pin_a) cudaFreeHost(pin_a);
  if (pin_b) cudaFreeHost(pin_b);
  if (vgraph) cudaGraphExecDestroy(vgraph);
  if (ev0) cudaEventDestroy(ev0);
  if (ev1) cudaEventDestroy(ev1);
  for (auto e : io_ev) cudaEventDestroy(e);
  for (auto &m : d_ctr_map) dev_free(m);
  dev_free(ctr_buf);
  if (h_ctr) cudaFreeHost(h_ctr);
  nested.reset();
  if (st && owns_stream) cudaStreamDestroy(st);
}

void Amg::alloc_vectors(Level &L)
{
  const size_t nb = (size_t)L.npad * L.b;
  L.result = nullptr;
  for (double **p : {&L.x, &L.y, &L.rhs, &L.res, &L.tmp}) {
    *p = dev_alloc<double>(nb);
    NGB_CUDA(cudaMemsetAsync(*p, 0, nb * sizeof(double), st));
  }
}

// plain CSR on the device (original numbering) -> level-scheduled split SELL layout + dinv
void Amg::build_level_layout(Level &L, const DevCsr &dA)
{
  const i64 n = L.n;
  const int b = L.b, bs = b * b;
  L.d_perm = upload_vec(L.perm, st);
  // permuted free flags (padding rows = not free)
  {
    std::vector<uint8_t> fp(L.npad, 0);
    for (i64 i = 0; i < n; i++) fp[L.perm[i]] = L.mask().empty() ? 1 : L.mask()[i];
    L.d_freep = upload_vec(fp, st);
  }
  i32 *len1 = dev_alloc<i32>(L.npad), *len2 = dev_alloc<i32>(L.npad);
  NGB_CUDA(cudaMemsetAsync(len1, 0, sizeof(i32) * L.npad, st));
  NGB_CUDA(cudaMemsetAsync(len2, 0, sizeof(i32) * L.npad, st));
  const bool has_n = L.nonfree_pad > 0;
  i32 *len3 = has_n ? dev_alloc<i32>(L.npad) : nullptr;
  if (has_n) NGB_CUDA(cudaMemsetAsync(len3, 0, sizeof(i32) * L.npad, st));
  k_layout_count<<<nblk(n), TB, 0, st>>>(n, dA.rowptr, dA.col, L.d_perm, L.d_perm, 1, len1, len2, (i32)L.nonfree_pad, len3);
  build_sell(L.npad, b, b, len1, L.L, st, &launches);
  build_sell(L.npad, b, b, len2, L.U, st, &launches);
  if (has_n) build_sell(L.npad, b, b, len3, L.N, st, &launches);
  L.diag = dev_alloc<double>((size_t)L.npad * bs);
  L.dinv = dev_alloc<double>((size_t)L.npad * bs);
  NGB_CUDA(cudaMemsetAsync(L.diag, 0, sizeof(double) * L.npad * bs, st));
  k_layout_fill<<<nblk(n), TB, 0, st>>>(n, bs, dA.rowptr, dA.col, dA.val, L.d_perm, L.d_perm, 1, L.L.slice_ptr, L.L.col, L.L.val,
                                        L.U.slice_ptr, L.U.col, L.U.val, L.diag, (i32)L.nonfree_pad, has_n ? L.N.slice_ptr : nullptr,
                                        L.N.col, L.N.val);
  dev_free(len3);
  {
    const int flip = (int)flags.num("b200_sort_flip", 0);
    if (flip < 2) {
      k_sell_sort_rows<<<nblk(L.npad), TB, 0, st>>>(L.npad, bs, L.L.slice_ptr, L.L.col, L.L.val, flip ? 1 : 0);
      k_sell_sort_rows<<<nblk(L.npad), TB, 0, st>>>(L.npad, bs, L.U.slice_ptr, L.U.col, L.U.val, flip ? 0 : 1);
    }
  }
  launches += 4;
  // true entry counts (for the byte model)
  {
    std::vector<i32> h1(L.npad), h2(L.npad);
    NGB_CUDA(cudaMemcpyAsync(h1.data(), len1, sizeof(i32) * L.npad, cudaMemcpyDeviceToHost, st));
    NGB_CUDA(cudaMemcpyAsync(h2.data(), len2, sizeof(i32) * L.npad, cudaMemcpyDeviceToHost, st));
    NGB_CUDA(cudaStreamSynchronize(st));
    i64 a = 0, c = 0;
    for (i64 i = 0; i < L.npad; i++) { a += h1[i]; c += h2[i]; }
    L.L.nnz = a; L.U.nnz = c;
  }
  dev_free(len1); dev_free(len2);
  {
    // slice width histogram -> register slot cache size of the sweeps: the smallest of 8 / 12 / 16 covering >= 97 % of the
    // slices (wider slices spill into the chunked path)
    auto pick = [&](const Sell &S) {
      std::vector<i64> sp(S.nslices + 1);
      NGB_CUDA(cudaMemcpyAsync(sp.data(), S.slice_ptr, sizeof(i64) * (S.nslices + 1), cudaMemcpyDeviceToHost, st));
      NGB_CUDA(cudaStreamSynchronize(st));
      i64 le8 = 0, le12 = 0, le16 = 0;
      for (i64 s = 0; s < S.nslices; s++) { const i64 w = sp[s + 1] - sp[s]; le8 += (w <= 8); le12 += (w <= 12); le16 += (w <= 16); }
      const double n = (double)std::max<i64>(S.nslices, 1);
      if (le8 >= 0.97 * n) return 8;
      if (le12 >= 0.97 * n) return 12;
      (void)le16;
      return 16;
    };
    const int force = (int)flags.num("b200_tri_pre", 0);
    L.pre_l = force ? force : pick(L.L);
    L.pre_u = force ? force : pick(L.U);
    L.L.maxw = sell_max_width(L.L, st);
    L.U.maxw = sell_max_width(L.U, st);
  }
  // dinv (GSS3::CalcDiags); the hybrid smoother passes a replacement diagonal (GSS3(A, repl_diag, ...), gssmoother.cpp:93-107)
  int *d_err = dev_alloc<int>(1);
  NGB_CUDA(cudaMemsetAsync(d_err, 0, sizeof(int), st));
  double *true_diag = L.diag, *md_planar = nullptr;
  if (L.par && !L.mod_diag.empty()) {
    double *aos = upload_vec(L.mod_diag, st);
    md_planar = dev_alloc<double>((size_t)L.npad * bs);
    NGB_CUDA(cudaMemsetAsync(md_planar, 0, sizeof(double) * L.npad * bs, st));
    k_aos_perm_to_planar<<<nblk(n), TB, 0, st>>>(n, bs, L.d_perm, aos, md_planar);
    NGB_CUDA(cudaStreamSynchronize(st));
    dev_free(aos);
    L.diag = md_planar;   // only while the inverses are computed
  }
  struct Restore { Level &L; double *t; double *&m; ~Restore() { L.diag = t; dev_free(m); } } restore{L, true_diag, md_planar};
  if (!L.pinv || L.sm_type == SM_JACOBI) {
    switch (b) {
      case 1: k_calc_dinv<1><<<nblk(L.npad), TB, 0, st>>>(L.npad, L.diag, L.d_freep, L.dinv, d_err); break;
      case 2: k_calc_dinv<2><<<nblk(L.npad), TB, 0, st>>>(L.npad, L.diag, L.d_freep, L.dinv, d_err); break;
      case 3: k_calc_dinv<3><<<nblk(L.npad), TB, 0, st>>>(L.npad, L.diag, L.d_freep, L.dinv, d_err); break;
      case 6: k_calc_dinv<6><<<nblk(L.npad, 64), 64, 0, st>>>(L.npad, L.diag, L.d_freep, L.dinv, d_err); break;
      default: throw Error("unsupported block size " + std::to_string(b));
    }
    launches++;
    int herr = 0;
    NGB_CUDA(cudaMemcpyAsync(&herr, d_err, sizeof(int), cudaMemcpyDeviceToHost, st));
    NGB_CUDA(cudaStreamSynchronize(st));
    if (herr) { dev_free(d_err); throw Error("singular diagonal block in smoother setup (use ngs_amg_regularize_cmats)"); }
  } else {
    // regularize_cmats => pinv smoothers (amg_pc.cpp:1124-1130): small host job, blocks round-trip through AoS
    double *aos = dev_alloc<double>((size_t)L.npad * bs);
    k_planar_to_aos<<<nblk(L.npad), TB, 0, st>>>(L.npad, bs, L.diag, aos);
    std::vector<double> h((size_t)L.npad * bs);
    std::vector<uint8_t> fp(L.npad);
    NGB_CUDA(cudaMemcpyAsync(h.data(), aos, sizeof(double) * h.size(), cudaMemcpyDeviceToHost, st));
    NGB_CUDA(cudaMemcpyAsync(fp.data(), L.d_freep, L.npad, cudaMemcpyDeviceToHost, st));
    NGB_CUDA(cudaStreamSynchronize(st));
    parallel_for(L.npad, [&](i64 lo, i64 hi) {
      for (i64 r = lo; r < hi; r++) {
        if (!fp[r]) { for (int e = 0; e < bs; e++) h[r * bs + e] = 0.0; continue; }
        block_pinv(b, &h[r * bs]);
      }
    }, 256);
    NGB_CUDA(cudaMemcpyAsync(aos, h.data(), sizeof(double) * h.size(), cudaMemcpyHostToDevice, st));
    k_aos_to_planar<<<nblk(L.npad), TB, 0, st>>>(L.npad, bs, aos, L.dinv);
    NGB_CUDA(cudaStreamSynchronize(st));
    launches += 2;
    dev_free(aos);
  }
  dev_free(d_err);
  // small levels: row-major copies of the triangles for the warp-per-row sweep (kernels_rm.cuh)
  // (block matrices at every size: the SELL walk of a warp-per-row sweep uses 8 of every 32 bytes it fetches -- a 386 k-row 6x6 level
  // moved 4x its 4.2 GB per sweep, 4.2 ms; the row-major copy is read in full sectors)
  const bool rm_size_ok = (L.b > 1 && tri_block_warp_rows) ? true : L.npad <= std::min(tri_small_rows, tri_rm_max_rows);
  if (L.sm_type == SM_GS && !L.tiled && rm_size_ok && tri_rm && !level_launch(L)) {
    build_rm(L.L, L.rmL, false, L.nonfree_pad);
    build_rm(L.U, L.rmU, true, L.nonfree_pad);
  }
}

// SELL triangle -> row-major copy (entries of a row contiguous, slot order kept)
void Amg::build_rm(const Sell &S, Rm &R, bool upper, i64 nonfree)
{
  const i64 np = S.nrows_pad;
  const int bs = S.bh * S.bw;
  i64 *cnt = dev_alloc<i64>(np + 1);
  R.ptr = dev_alloc<i64>(np + 1);
  k_rm_count<<<nblk(np + 1), TB, 0, st>>>(np, S.view(), cnt);
  size_t tb = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, tb, cnt, R.ptr, np + 1, st);
  void *tmp = dev_alloc<char>(tb);
  cub::DeviceScan::ExclusiveSum(tmp, tb, cnt, R.ptr, np + 1, st);
  i64 total = 0;
  NGB_CUDA(cudaMemcpyAsync(&total, R.ptr + np, sizeof(i64), cudaMemcpyDeviceToHost, st));
  NGB_CUDA(cudaStreamSynchronize(st));
  cudaFree(tmp);
  dev_free(cnt);
  R.col = dev_alloc<i32>(std::max<i64>(total, 1));
  R.val = dev_alloc<double>(std::max<i64>(total * bs, 1));
  R.gate = dev_alloc<i32>(std::max<i64>(np, 1));
  k_rm_fill<<<nblk(np), TB, 0, st>>>(np, bs, S.view(), R.ptr, R.col, R.val, R.gate, upper ? 1 : 0, (i32)nonfree);
  launches += 3;
}

// ---- block Gauss-Seidel (sm_type = bgs) --------------------------------------------------------------------------------------------
// Reference: BSmoother2 (src/base/smoothers/loc_block_gssmoother_impl.hpp): blocks = the aggregates of the next coarse map in coarse
// vertex order (GetGSBlocks, amg_pc_vertex_impl.hpp:1171-1269; vertices mapped to -1 are in no block and are not smoothed), per block
// the dense diagonal block DB and its inverse; RichardsonUpdate (:244-268): x_B += DB^-1 (b_B - A_{B,:} x); RichardsonUpdate_RES
// (:516-541): d = DB^-1 res_B, x_B += d, res -= A_{:,B} d; blocks ascending (forward) or descending (IterateBlocks :618-651), omega = 1.
// Here: with A~ = DB^-1 A (unit diagonal, NO couplings inside a block) the block update of B is the point update of its rows on A~, and
// the rows of a block are independent of each other -- a block sweep over A is an ordinary point sweep over A~ in block-major order.
// The shadow level L.bgs holds A~ in the split layout (all sweep kernels, schedules and row-major copies apply unchanged); the level itself
// keeps A in the same numbering for the residual updates and the SpMV.  A~ has the union pattern of a block's rows (more entries than
// A: the price of reusing the point kernels; the reference applies DB^-1 densely instead).
void Amg::setup_bgs(Level &L, const DevCsr &dA)
{
  const i64 n = L.n;
  const int b = L.b, bs = b * b;
  const HostBsr &A = L.hA;
  i64 nb = 0;
  for (i64 v = 0; v < n; v++) nb = std::max<i64>(nb, (i64)L.gs_block[v] + 1);
  // members of every block, ascending
  std::vector<i64> bptr(nb + 1, 0);
  for (i64 v = 0; v < n; v++)
    if (L.gs_block[v] >= 0 && (L.free_mask.empty() || L.free_mask[v])) bptr[L.gs_block[v] + 1]++;
  for (i64 k = 0; k < nb; k++) bptr[k + 1] += bptr[k];
  std::vector<i32> mem(bptr[nb]);
  {
    std::vector<i64> pos(bptr.begin(), bptr.end() - 1);
    for (i64 v = 0; v < n; v++)
      if (L.gs_block[v] >= 0 && (L.free_mask.empty() || L.free_mask[v])) mem[pos[L.gs_block[v]]++] = (i32)v;
  }
  std::vector<uint8_t> in_block(n, 0);
  for (i32 v : mem) in_block[v] = 1;
  // DB^-1 as a block-diagonal sparse matrix (original numbering); rows outside the blocks stay empty
  HostBsr D;
  D.nrows = D.ncols = n; D.bh = D.bw = b;
  D.rowptr.assign(n + 1, 0);
  for (i64 k = 0; k < nb; k++) for (i64 q = bptr[k]; q < bptr[k + 1]; q++) D.rowptr[mem[q] + 1] = bptr[k + 1] - bptr[k];
  for (i64 v = 0; v < n; v++) D.rowptr[v + 1] += D.rowptr[v];
  D.col.resize(D.rowptr[n]);
  D.val.assign((size_t)D.rowptr[n] * bs, 0.0);
  std::atomic<int> singular{0};
  parallel_for(nb, [&](i64 lo, i64 hi) {
    std::vector<double> dense;
    for (i64 k = lo; k < hi; k++) {
      const i64 m = bptr[k + 1] - bptr[k];
      if (m == 0) continue;
      const int N = (int)m * b;
      dense.assign((size_t)N * N, 0.0);
      for (i64 qi = 0; qi < m; qi++) {
        const i32 i = mem[bptr[k] + qi];
        for (i64 e = A.rowptr[i]; e < A.rowptr[i + 1]; e++) {
          const i32 j = A.col[e];
          if (L.gs_block[j] != (i32)k || !in_block[j]) continue;
          const i64 qj = std::lower_bound(mem.begin() + bptr[k], mem.begin() + bptr[k + 1], j) - (mem.begin() + bptr[k]);
          for (int p = 0; p < b; p++) for (int q = 0; q < b; q++) dense[(size_t)(qi * b + p) * N + qj * b + q] = A.val[e * bs + p * b + q];
        }
      }
      if (!dense_invert(N, dense)) { singular.store(1); continue; }
      for (i64 qi = 0; qi < m; qi++) {
        const i32 i = mem[bptr[k] + qi];
        for (i64 qj = 0; qj < m; qj++) {
          const i64 e = D.rowptr[i] + qj;
          D.col[e] = mem[bptr[k] + qj];
          for (int p = 0; p < b; p++) for (int q = 0; q < b; q++) D.val[e * bs + p * b + q] = dense[(size_t)(qi * b + p) * N + qj * b + q];
        }
      }
    }
  }, 64);
  if (singular.load()) throw Error("bgs: singular diagonal block");
  // A~ = DB^-1 A on the device; inside a block it is the identity by construction: set it exactly (rounding leaves ~1e-17 couplings
  // that would chain the rows of a block) and drop those entries from the pattern
  HostBsr At;
  {
    DevCsr dD, dAt;
    dev_csr_upload(D, dD, st);
    dev_spgemm(dD, dA, dAt, st, &launches);
    dev_csr_download(dAt, At, st, true);
    dev_csr_free(dD); dev_csr_free(dAt);
  }
  {
    HostBsr F;
    F.nrows = F.ncols = n; F.bh = F.bw = b;
    F.rowptr.assign(n + 1, 0);
    for (i64 i = 0; i < n; i++) {
      i64 c = in_block[i] ? 1 : 0;
      if (in_block[i])
        for (i64 e = At.rowptr[i]; e < At.rowptr[i + 1]; e++) c += (L.gs_block[At.col[e]] != L.gs_block[i] || !in_block[At.col[e]]);
      F.rowptr[i + 1] = F.rowptr[i] + c;
    }
    F.col.resize(F.rowptr[n]);
    F.val.assign((size_t)F.rowptr[n] * bs, 0.0);
    parallel_for(n, [&](i64 lo, i64 hi) {
      for (i64 i = lo; i < hi; i++) {
        if (!in_block[i]) continue;
        i64 o = F.rowptr[i];
        bool diag_done = false;
        auto put_diag = [&]() { F.col[o] = (i32)i; for (int p = 0; p < b; p++) F.val[o * bs + p * b + p] = 1.0; o++; diag_done = true; };
        for (i64 e = At.rowptr[i]; e < At.rowptr[i + 1]; e++) {
          const i32 j = At.col[e];
          if (L.gs_block[j] == L.gs_block[i] && in_block[j]) continue;
          if (!diag_done && j > i) put_diag();
          F.col[o] = j;
          for (int q = 0; q < bs; q++) F.val[o * bs + q] = At.val[e * bs + q];
          o++;
        }
        if (!diag_done) put_diag();
      }
    }, 1024);
    At = std::move(F);
  }
  // shadow level: A~ with the block-major sweep order
  L.bgs = std::make_unique<Level>();
  Level &S = *L.bgs;
  S.n = n; S.b = b; S.nnz = At.nnz();
  S.free_mask = in_block;
  S.sweep_rank.assign(n, 0);
  {
    i32 r = 0;
    for (i32 v : mem) S.sweep_rank[v] = r++;
    for (i64 v = 0; v < n; v++) if (!in_block[v]) S.sweep_rank[v] = r++;
  }
  S.hA = std::move(At);
  S.sm_type = SM_GS;
  S.pinv = false;
  level_schedule(S.hA, S.free_mask, true, S, st);
  S.d_err = d_err;
  {
    DevCsr dAt;
    dev_csr_upload(S.hA, dAt, st);
    build_level_layout(S, dAt);
    dev_csr_free(dAt);
  }
  alloc_vectors(S);
  HostBsr().rowptr.swap(S.hA.rowptr); std::vector<i32>().swap(S.hA.col); std::vector<double>().swap(S.hA.val);
  // the level itself lives in the same numbering (its own split is only used as a whole: residual, SpMV)
  L.perm = S.perm; L.npad = S.npad; L.nonfree_pad = S.nonfree_pad; L.depth = S.depth; L.level_start = S.level_start;
  L.hM = std::move(D);   // parked until the level's permutation is on the device (finish in finalize)
}

// SmoothRESSimple (loc_block_gssmoother_impl.hpp:691-706): d = (T~ + I)^-1 DB^-1 res over the blocks in sweep order, x += d, res -= A d
void Amg::bgs_res(Level &L, bool backward, double *x, double *res)
{
  Level &S = *L.bgs;
  const i64 np = L.npad * L.b;
  transfer(L.DBI, res, nullptr, L.tmp, 1.0, 0.0);
  tri_dispatch(S, backward, false, true, L.tmp, nullptr, L.y, S.res);
  k_axpby<<<nblk(np), TB, 0, st>>>(np, 1.0, L.y, 1.0, x);
  launches++;
  spmv_part(L, 4, L.y, res, res, -1.0, 1.0, nullptr);
}

// SmoothSimple (:672-688): x_B += DB^-1 (b_B - A_{B,:} x) over the blocks in sweep order == point sweep on A~ against DB^-1 b
void Amg::bgs_rhs(Level &L, bool backward, double *x, const double *b)
{
  Level &S = *L.bgs;
  transfer(L.DBI, b, nullptr, S.rhs, 1.0, 0.0);
  gs_rhs(S, backward, x, S.rhs, L.y);
  NGB_CUDA(cudaMemcpyAsync(x, L.y, sizeof(double) * L.npad * L.b, cudaMemcpyDeviceToDevice, st));
}

// the instantiations of the CTA-per-tile sweep: 256-row tiles run on 128-thread CTAs (6 per SM), 512-row tiles on 256-thread CTAs (3 per SM)
using CTileKernel = void (*)(SellView, const double *, const double *, const double *, const double *, double *, double *, CTileParams);
static int ctile_threads(int maxs) { return maxs <= 8 ? 128 : 256; }
static CTileKernel ctile_kernel(int maxs, int nbuf, bool add_self)
{
  if (maxs <= 8) {
    if (nbuf == 1) return add_self ? (CTileKernel)k_gs_ctile<128, 8, 1, true, false, 6> : (CTileKernel)k_gs_ctile<128, 8, 1, false, true, 6>;
    return add_self ? (CTileKernel)k_gs_ctile<128, 8, 2, true, false, 6> : (CTileKernel)k_gs_ctile<128, 8, 2, false, true, 6>;
  }
  if (nbuf == 1) return add_self ? (CTileKernel)k_gs_ctile<256, 16, 1, true, false, 3> : (CTileKernel)k_gs_ctile<256, 16, 1, false, true, 3>;
  return add_self ? (CTileKernel)k_gs_ctile<256, 16, 2, true, false, 3> : (CTileKernel)k_gs_ctile<256, 16, 2, false, true, 3>;
}
using ITileKernel = void (*)(const double *, const double *, double *, double *, ITileParams);
static ITileKernel itile_kernel(int maxs, bool add_self, int minb)
{
  if (maxs <= 8) {
    if (minb >= 7) return add_self ? (ITileKernel)k_gs_itile<128, 256, true, false, 7> : (ITileKernel)k_gs_itile<128, 256, false, true, 7>;
    return add_self ? (ITileKernel)k_gs_itile<128, 256, true, false, 6> : (ITileKernel)k_gs_itile<128, 256, false, true, 6>;
  }
  return add_self ? (ITileKernel)k_gs_itile<128, 512, true, false, 3> : (ITileKernel)k_gs_itile<128, 512, false, true, 3>;
}

// Two-level (tile) schedule of the triangular sweeps of a level (tiles.hpp): only for scalar levels that are big and whose sweep DAG is
// deep -- a colour-major coarse level has a shallower DAG than any tiling of it.  A = the matrix the sweep runs on (the level matrix, or
// the master-master block M of a distributed level with the hybrid stage order in L.sweep_rank).
bool Amg::setup_tiles(Level &L, int l, const HostBsr &A)
{
  const bool verbose = flags.str("log_level", "none") != "none";
  if (L.b != 1 || !flags.flag("b200_tile_sweep", true) || L.n < (i64)flags.num("b200_tile_min_rows", 200000)) return false;
  if (sweep_depth(A, L.mask(), L.sweep_rank) < (int)flags.num("b200_tile_min_depth", 150)) return false;
  TileSchedule ts;
  const int cap = (int)flags.num("b200_tile_rows", 256);
  if (cap != 32 && cap != 64 && cap != 256 && cap != 512)
    throw Error("ngs_amg_b200_tile_rows must be 32 or 64 (one warp per tile) or 256 or 512 (one CTA per tile)");
  int rounds = 5;
  while ((1 << rounds) < cap) rounds++;
  // matrices numbered like a structured grid get near-cubic boxes as cluster hints (ideal tile DAG); everything else the pairwise clustering
  std::vector<i32> hint;
  i64 gd[3] = {0, 0, 0};
  const bool grid = flags.flag("b200_tile_grid_hint", true) && grid_box_hint(A, L.mask(), cap, hint, gd);
  if (verbose && grid) std::fprintf(stderr, "[ngsamg_b200] level %d: numbered like a %lld x %lld x %lld grid: box-shaped tiles\n", l, (long long)gd[0], (long long)gd[1], (long long)gd[2]);
  build_tile_schedule(A, L.mask(), L.sweep_rank, (int)flags.num("b200_tile_rounds", rounds), cap, ts, grid ? &hint : nullptr);
  if (!ts.ok) return false;
  L.perm = ts.perm; L.npad = ts.npad; L.nonfree_pad = ts.nonfree_pad; L.depth = ts.tile_depth;
  L.level_start.clear();
  L.ntiles = ts.ntiles; L.tile_maxs = cap / 32;   // 1, 2: warp per tile; 8, 16: CTA per tile
  L.d_tile_slice = upload_vec(ts.tile_slice, st); L.d_tile_nlev = upload_vec(ts.tile_nlev, st);
  L.d_row_lvl = upload_vec(ts.row_lvl, st);
  L.d_tile_pred_ptr = upload_vec(ts.pred_ptr, st); L.d_tile_pred = upload_vec(ts.pred, st);
  L.d_tile_succ_ptr = upload_vec(ts.succ_ptr, st); L.d_tile_succ = upload_vec(ts.succ, st);
  L.d_tile_done = dev_alloc<int>((size_t)ts.ntiles);
  L.h_tile_slice = ts.tile_slice; L.h_tile_nlev = ts.tile_nlev; L.h_pred_ptr = ts.pred_ptr; L.h_succ_ptr = ts.succ_ptr;
  L.h_tile_nreal.assign(ts.ntiles, 0);
  parallel_for(ts.ntiles, [&](i64 lo, i64 hi) {
    for (i64 t = lo; t < hi; t++) {
      i32 c = 0;
      for (i64 r = (i64)ts.tile_slice[t] * 32; r < (i64)ts.tile_slice[t + 1] * 32; r++) c += ts.row_lvl[r] != 255;
      L.h_tile_nreal[t] = c;
    }
  }, 1024);
  L.tile_nbuf = (int)flags.num("b200_tile_nbuf", 1);
  if (L.tile_nbuf != 1 && L.tile_nbuf != 2) throw Error("ngs_amg_b200_tile_nbuf must be 1 or 2");
  L.tiled = true;
  if (verbose) std::fprintf(stderr, "[ngsamg_b200] level %d: tile schedule: %lld tiles, tile DAG depth %d, <= %d local levels\n", l, (long long)ts.ntiles, ts.tile_depth, ts.max_local_levels);
  return true;
}

// CTA-per-tile sweep: slab capacity (largest tile of L and U, in SELL slots), shared-memory opt-in, resident grid
void Amg::prepare_ctile(Level &L)
{
  i64 cap = 1;
  const i64 nt = L.ntiles;
  std::vector<CTileMeta> meta[2];
  for (int dir = 0; dir < 2; dir++) {
    const Sell *S = dir ? &L.U : &L.L;
    const std::vector<i64> &dp = dir ? L.h_succ_ptr : L.h_pred_ptr;
    std::vector<i64> sp(S->nslices + 1);
    NGB_CUDA(cudaMemcpyAsync(sp.data(), S->slice_ptr, sizeof(i64) * (S->nslices + 1), cudaMemcpyDeviceToHost, st));
    NGB_CUDA(cudaStreamSynchronize(st));
    meta[dir].resize(nt);
    for (i64 t = 0; t < nt; t++) {
      const i32 s0 = L.h_tile_slice[t], s1 = L.h_tile_slice[t + 1];
      const i64 slots = sp[s1] - sp[s0];
      cap = std::max(cap, slots);
      for (i32 s2 = s0; s2 < s1; s2++)
        if (sp[s2 + 1] - sp[s2] > 255) throw Error("tile sweep: a row with more than 255 entries in one triangle; disable ngs_amg_b200_tile_sweep for this matrix");
      if (dp[t] >= (i64)2147483647) throw Error("tile sweep: wait lists too long");
      meta[dir][t] = CTileMeta{sp[s0], s0, s1 - s0, (i32)slots, L.h_tile_nlev[t] | (L.h_tile_nreal[t] << 16), (i32)dp[t], (i32)(dp[t + 1] - dp[t])};
    }
  }
  L.d_meta_fwd = upload_vec(meta[0], st);
  L.d_meta_bwd = upload_vec(meta[1], st);
  L.tile_cap_slots = (int)cap;
  L.ctile_smem = ctile_smem_bytes(L.tile_maxs, L.tile_cap_slots, L.tile_nbuf);
  int dev_max = 0;
  NGB_CUDA(cudaDeviceGetAttribute(&dev_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
  if (L.ctile_smem > (size_t)dev_max)
    throw Error("tile sweep: a tile needs " + std::to_string(L.ctile_smem) + " bytes of shared memory (device limit " + std::to_string(dev_max) + "); use smaller tiles (ngs_amg_b200_tile_rows)");
  for (int as = 0; as < 2; as++) {
    CTileKernel k = ctile_kernel(L.tile_maxs, L.tile_nbuf, as == 1);
    {
      // the opt-in is per kernel, not per level: keep the largest request of any hierarchy of this process
      static std::mutex mu;
      static size_t granted[2][2][2] = {};
      std::lock_guard<std::mutex> guard(mu);
      size_t &g = granted[L.tile_maxs <= 8 ? 0 : 1][L.tile_nbuf - 1][as];
      g = std::max(g, L.ctile_smem);
      NGB_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g));
    }
    int occ = 0;
    NGB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k, ctile_threads(L.tile_maxs), L.ctile_smem));
    occ = std::max(1, occ);
    if (tri_ctas_per_sm > 0) occ = std::min(occ, tri_ctas_per_sm);
    L.ctile_grid[as] = (int)std::max<i64>(1, std::min<i64>(L.ntiles, (i64)occ * num_sms));
    if (flags.str("log_level", "none") != "none")
      std::fprintf(stderr, "[ngsamg_b200] tile sweep (%s): %lld tiles, %d slab(s) of %d slots, %zu B smem, %d CTAs/SM, grid %d\n", as ? "backward/rhs" : "forward/res",
                   (long long)L.ntiles, L.tile_nbuf, L.tile_cap_slots, L.ctile_smem, occ, L.ctile_grid[as]);
  }
}

// Prepared tile images (kernels_itile.cuh): built once per level and direction on the device; false = some row has more than 7 entries
// in a triangle (the general kernel k_gs_ctile stays in charge).
bool Amg::prepare_itile(Level &L)
{
  if (!flags.flag("b200_tile_image", true)) return false;
  const i64 nt = L.ntiles;
  const int maxrows = L.tile_maxs * 32;
  i32 *d_nxr = dev_alloc<i32>(nt), *d_nx = dev_alloc<i32>(nt), *d_nov = dev_alloc<i32>(nt), *d_max = dev_alloc<i32>(1);
  i32 *d_nreal = upload_vec(L.h_tile_nreal, st);
  std::vector<i32> nxr(nt), nx(nt), nov(nt);
  bool ok = true;
  for (int dir = 0; dir < 2 && ok; dir++) {
    const Sell &S = dir ? L.U : L.L;
    const std::vector<i64> &dp = dir ? L.h_succ_ptr : L.h_pred_ptr;
    NGB_CUDA(cudaMemsetAsync(d_max, 0, sizeof(i32), st));
    k_img_count<<<(unsigned)nt, 128, 0, st>>>((i32)nt, L.d_tile_slice, S.view(), d_nxr, d_nx, d_nov, d_max);
    i32 mx = 0;
    NGB_CUDA(cudaMemcpyAsync(&mx, d_max, sizeof(i32), cudaMemcpyDeviceToHost, st));
    NGB_CUDA(cudaMemcpyAsync(nxr.data(), d_nxr, sizeof(i32) * nt, cudaMemcpyDeviceToHost, st));
    NGB_CUDA(cudaMemcpyAsync(nx.data(), d_nx, sizeof(i32) * nt, cudaMemcpyDeviceToHost, st));
    NGB_CUDA(cudaMemcpyAsync(nov.data(), d_nov, sizeof(i32) * nt, cudaMemcpyDeviceToHost, st));
    NGB_CUDA(cudaStreamSynchronize(st));
    if (mx > IT_NV + IT_MAXOVF) { ok = false; break; }
    std::vector<i64> off(nt + 1, 0);
    std::vector<ITileMeta> meta(nt);
    i64 cap = 0;
    for (i64 t = 0; t < nt; t++) {
      const int nrow = (L.h_tile_slice[t + 1] - L.h_tile_slice[t]) * 32;
      const i64 b = itile_image_bytes(L.h_tile_nlev[t], nrow, nxr[t], nx[t], nov[t], dir == 0);
      if (nx[t] > 65535 || nov[t] > 4095 || b > (i64)200 * 1024) { ok = false; break; }
      off[t + 1] = off[t] + b;
      cap = std::max(cap, b);
      meta[t] = ITileMeta{off[t], L.h_tile_slice[t] * 32, nrow, (i32)b, 0, (i32)dp[t], (i32)(dp[t + 1] - dp[t])};
    }
    if (!ok) break;
    L.itile_cap = std::max(L.itile_cap, (int)cap);
    L.d_img[dir] = dev_alloc<unsigned char>((size_t)off[nt]);
    L.img_bytes_total += off[nt];
    i64 *d_off = upload_vec(off, st);
    if (maxrows <= 256)
      k_img_fill<256><<<(unsigned)nt, 128, 0, st>>>((i32)nt, L.d_tile_slice, L.d_tile_nlev, d_nreal, L.d_row_lvl, S.view(), L.dinv, L.diag, dir == 0 ? 1 : 0, d_off, d_nxr, d_nx, d_nov, L.d_img[dir]);
    else
      k_img_fill<512><<<(unsigned)nt, 128, 0, st>>>((i32)nt, L.d_tile_slice, L.d_tile_nlev, d_nreal, L.d_row_lvl, S.view(), L.dinv, L.diag, dir == 0 ? 1 : 0, d_off, d_nxr, d_nx, d_nov, L.d_img[dir]);
    launches += 2;
    L.d_imeta[dir] = upload_vec(meta, st);
    NGB_CUDA(cudaStreamSynchronize(st));
    NGB_CUDA(cudaGetLastError());
    dev_free(d_off);
  }
  dev_free(d_nxr); dev_free(d_nx); dev_free(d_nov); dev_free(d_max); dev_free(d_nreal);
  if (!ok) {
    for (int d = 0; d < 2; d++) { dev_free(L.d_img[d]); dev_free(L.d_imeta[d]); }
    return false;
  }
  L.itile_smem = itile_smem_bytes(maxrows, L.itile_cap);
  int dev_max = 0;
  NGB_CUDA(cudaDeviceGetAttribute(&dev_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
  if (L.itile_smem > (size_t)dev_max) { for (int d = 0; d < 2; d++) { dev_free(L.d_img[d]); dev_free(L.d_imeta[d]); } return false; }
  for (int as = 0; as < 2; as++) {
    ITileKernel k = itile_kernel(L.tile_maxs, as == 1, itile_minb);
    {
      static std::mutex mu;
      static size_t granted[3][2] = {};
      std::lock_guard<std::mutex> guard(mu);
      size_t &g = granted[L.tile_maxs <= 8 ? (itile_minb >= 7 ? 2 : 0) : 1][as];
      g = std::max(g, L.itile_smem);
      NGB_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g));
    }
    int occ = 0;
    NGB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k, 128, L.itile_smem));
    occ = std::max(1, occ);
    if (tri_ctas_per_sm > 0) occ = std::min(occ, tri_ctas_per_sm);
    L.itile_grid[as] = (int)std::max<i64>(1, std::min<i64>(L.ntiles, (i64)occ * num_sms));
    if (flags.str("log_level", "none") != "none")
      std::fprintf(stderr, "[ngsamg_b200] tile images (%s): %lld tiles, largest image %d B, %zu B smem, %d CTAs/SM, grid %d, %.2f GB of images\n", as ? "backward/rhs" : "forward/res",
                   (long long)L.ntiles, L.itile_cap, L.itile_smem, occ, L.itile_grid[as], L.img_bytes_total / 1e9);
  }
  L.itile = true;
  return true;
}

void Amg::build_transfer_layout(Level &F, Level &C)
{
  // P: rows = fine (level-scheduled), cols = coarse (level-scheduled); PT the other way round
  DevCsr dP, dPT;
  dev_csr_upload(F.hP, dP, st);
  dev_transpose(dP, dPT, st, &launches);
  HostBsr PT;                                  // pattern only: the storage order of the restriction rows is decided on the host
  dev_csr_download(dPT, PT, st, false);
  {
    i32 *len = dev_alloc<i32>(F.npad);
    NGB_CUDA(cudaMemsetAsync(len, 0, sizeof(i32) * F.npad, st));
    k_layout_count<<<nblk(F.n), TB, 0, st>>>(F.n, dP.rowptr, dP.col, F.d_perm, C.d_perm, 0, len, nullptr, 0, nullptr);
    build_sell(F.npad, F.b, F.bc, len, F.P, st, &launches);
    k_layout_fill<<<nblk(F.n), TB, 0, st>>>(F.n, F.b * F.bc, dP.rowptr, dP.col, dP.val, F.d_perm, C.d_perm, 0, F.P.slice_ptr, F.P.col,
                                            F.P.val, nullptr, nullptr, nullptr, nullptr, 0, nullptr, nullptr, nullptr);
    F.P.nnz = dP.nnz;
    dev_free(len);
    F.P.maxw = sell_max_width(F.P, st);
  }
  {
    // restriction matrix: rows stored sorted by length (uniform slices => little SELL padding); the kernel scatters its result
    // through rowmap.  Counting sort on the row length, stable in the coarse (level-scheduled) row number.
    const i64 npt = round32(C.n);
    std::vector<i32> order(C.n), inv(C.npad, -1);
    for (i64 c = 0; c < C.n; c++) inv[C.perm[c]] = (i32)c;
    i64 maxlen = 0;
    for (i64 c = 0; c < C.n; c++) maxlen = std::max<i64>(maxlen, PT.rowptr[c + 1] - PT.rowptr[c]);
    std::vector<i64> bucket(maxlen + 2, 0);
    for (i64 c = 0; c < C.n; c++) bucket[PT.rowptr[c + 1] - PT.rowptr[c] + 1]++;
    for (i64 l = 0; l <= maxlen; l++) bucket[l + 1] += bucket[l];
    std::vector<i32> spos(C.n), rowmap(npt, -1);
    if (flags.flag("b200_pt_locality", true)) {
      // inside a length class, order the rows by the first fine row they read: consecutive restriction rows then gather from the
      // same sectors of the fine residual (the coarse numbering is colour-major, i.e. spatially incoherent with the fine one)
      std::vector<i32> key(C.n, 0);
      parallel_for(C.n, [&](i64 lo, i64 hi) {
        for (i64 c = lo; c < hi; c++) {
          i32 m = std::numeric_limits<i32>::max();
          for (i64 k = PT.rowptr[c]; k < PT.rowptr[c + 1]; k++) m = std::min(m, F.perm[PT.col[k]]);
          key[c] = m;
        }
      });
      for (i64 c = 0; c < C.n; c++) order[c] = (i32)c;
      std::sort(order.begin(), order.end(), [&](i32 a, i32 b) {
        const i64 la = PT.rowptr[a + 1] - PT.rowptr[a], lb = PT.rowptr[b + 1] - PT.rowptr[b];
        return la < lb || (la == lb && (key[a] < key[b] || (key[a] == key[b] && a < b)));
      });
      for (i64 s2 = 0; s2 < C.n; s2++) { spos[order[s2]] = (i32)s2; rowmap[s2] = C.perm[order[s2]]; }
    } else
    for (i64 pc = 0; pc < C.npad; pc++) {
      const i32 c = inv[pc];
      if (c < 0) continue;
      const i64 s = bucket[PT.rowptr[c + 1] - PT.rowptr[c]]++;
      spos[c] = (i32)s;
      rowmap[s] = (i32)pc;
    }
    i32 *d_spos = upload_vec(spos, st);
    F.d_pt_rowmap = upload_vec(rowmap, st);
    i32 *len = dev_alloc<i32>(npt);
    NGB_CUDA(cudaMemsetAsync(len, 0, sizeof(i32) * npt, st));
    k_layout_count<<<nblk(C.n), TB, 0, st>>>(C.n, dPT.rowptr, dPT.col, d_spos, F.d_perm, 0, len, nullptr, 0, nullptr);
    build_sell(npt, F.bc, F.b, len, F.PT, st, &launches);
    k_layout_fill<<<nblk(C.n), TB, 0, st>>>(C.n, F.b * F.bc, dPT.rowptr, dPT.col, dPT.val, d_spos, F.d_perm, 0, F.PT.slice_ptr,
                                            F.PT.col, F.PT.val, nullptr, nullptr, nullptr, nullptr, 0, nullptr, nullptr, nullptr);
    F.PT.nnz = dPT.nnz;
    NGB_CUDA(cudaStreamSynchronize(st));
    dev_free(len);
    dev_free(d_spos);
  }
  launches += 4;
  NGB_CUDA(cudaStreamSynchronize(st));
  dev_csr_free(dP);
  dev_csr_free(dPT);
}

// CoarseLevelInv (amg_pc.cpp:843-928): exact inverse of the coarsest matrix on its free dofs; here an explicit
// dense inverse (the level has at most a few thousand scalar dofs) applied as a GEMV on the device.
void Amg::build_coarse_inverse(Level &L)
{
  const int b = L.b;
  const i64 N = L.n * b;
  if (N > 8192) throw Error("coarsest level too large for the dense inverse (" + std::to_string(N) + " scalar dofs); raise ngs_amg_max_levels");
  const HostBsr &A = L.hA;
  std::vector<i64> g2l(N, -1), l2g;
  for (i64 i = 0; i < L.n; i++)
    for (int p = 0; p < b; p++)
      if (L.free_mask.empty() || L.free_mask[i]) { g2l[i * b + p] = (i64)l2g.size(); l2g.push_back(i * b + p); }
  const int cn = (int)l2g.size();
  std::vector<double> M((size_t)cn * cn, 0.0);
  for (i64 i = 0; i < L.n; i++)
    for (i64 k = A.rowptr[i]; k < A.rowptr[i + 1]; k++)
      for (int p = 0; p < b; p++)
        for (int q = 0; q < b; q++) {
          const i64 r = g2l[i * b + p], c = g2l[(i64)A.col[k] * b + q];
          if (r >= 0 && c >= 0) M[(size_t)r * cn + c] = A.val[k * b * b + p * b + q];
        }
  if (regularize_coarse_dim) {
    // RegularizeMatrix on the diagonal blocks (amg_pc.cpp:861-862, elasticity_pc_impl.hpp:711-763) before the inverse
    std::vector<double> blk((size_t)b * b);
    for (i64 i = 0; i < L.n; i++)
      for (i64 k = A.rowptr[i]; k < A.rowptr[i + 1]; k++) {
        if (A.col[k] != i) continue;
        std::copy(A.val.begin() + k * b * b, A.val.begin() + (k + 1) * b * b, blk.begin());
        block_regularize(b, blk.data(), regularize_coarse_dim);
        for (int p = 0; p < b; p++)
          for (int q = 0; q < b; q++) {
            const i64 r = g2l[i * b + p], c = g2l[i * b + q];
            if (r >= 0 && c >= 0) M[(size_t)r * cn + c] = blk[p * b + q];
          }
      }
  }
  if (cn > 0 && !dense_invert(cn, M)) throw Error("coarsest level matrix is singular");
  // embed into the padded level numbering (identity permutation on the coarsest level)
  const i64 NP = L.npad * b;
  std::vector<double> full((size_t)NP * NP, 0.0);
  for (int r = 0; r < cn; r++)
    for (int c = 0; c < cn; c++) full[(size_t)L.perm[l2g[r] / b] * b * NP + (l2g[r] % b) * NP + (size_t)L.perm[l2g[c] / b] * b + (l2g[c] % b)] = M[(size_t)r * cn + c];
  d_cinv = upload_vec(full, st);
  cinv_n = (int)NP;
  has_cinv = true;
}

// ------------------------------------------------------------------------------------------------
// setup driver == BaseAMGPC::BuildAMGMat (amg_pc.cpp:565-736) + BaseAMGFactory::SetUpLevels (base_factory.cpp:219-353)
// ------------------------------------------------------------------------------------------------
namespace {
void build_plain_sell(const HostBsr &H, const i32 *d_rperm, const i32 *d_cperm, i64 nrows_pad, Sell &S, cudaStream_t st, i64 *launches);
}

void Amg::finalize()
{
  if (finalized) throw Error("finalize called twice");
  if (par) { finalize_parallel(); return; }
  auto t0 = std::chrono::steady_clock::now();
  double host_s = 0, rap_ms = 0, rap_bytes = 0;
  const bool verbose = flags.str("log_level", "none") != "none";   // factory log levels, base_factory.cpp:83-199
  auto tick = [](std::chrono::steady_clock::time_point t) { return std::chrono::duration<double>(std::chrono::steady_clock::now() - t).count(); };
  const int max_levels = (int)flags.num("max_levels", 10);            // base_factory.hpp:88-152
  const i64 max_coarse = (i64)flags.num("max_coarse_size", 50);
  {
    const std::string cyc = flags.str("mg_cycle", "V");       // AMGMatrix::SetCycle, amg_pc.cpp:578-586
    if (cyc == "V" || cyc == "v") cycle = CYCLE_V;
    else if (cyc == "W" || cyc == "w") cycle = CYCLE_W;
    else if (cyc == "BS" || cyc == "bs") cycle = CYCLE_BS;
    else throw Error("mg_cycle=" + cyc + " is not supported (V | W | BS)");
  }
  const std::string clev = flags.str("clev", "inv");
  const bool elast = type.find("elast") != std::string::npos;
  const int dim = (type.find("2d") != std::string::npos) ? 2 : 3;
  const bool regularize = flags.flag("regularize_cmats", elast);        // elasticity_pc_impl.hpp:139, h1_impl.hpp:275-279
  regularize_coarse_dim = (regularize && elast) ? (type.find("2d") != std::string::npos ? 2 : 3) : 0;
  CoarsenOptions copt;
  copt.max_per_row = (int)flags.num("sp_max_per_row", elast ? 1 + dim : 3);
  copt.min_frac = flags.num("sp_min_frac", dim == 3 ? 0.08 : 0.1);
  copt.omega = flags.num("sp_omega", 1.0);
  copt.smooth = flags.str("prol_type", "semi_aux_smoothed") != "piecewise";
  copt.rounds = (int)flags.num("spw_rounds", 3);

  d_err = dev_alloc<int>(1);
  NGB_CUDA(cudaMemsetAsync(d_err, 0, sizeof(int), st));
  DevCsr dA;
  dev_csr_upload(lev[0]->hA, dA, st);
  for (int l = 0;; l++) {
    Level &L = *lev[l];
    L.n = L.hA.nrows; L.b = L.hA.bh; L.nnz = L.hA.nnz();
    bool coarsest = injected.empty() ? (l + 1 >= max_levels || L.n <= max_coarse) : (l >= (int)injected.size());
    std::vector<i32> vmap;   // vertex -> coarse vertex of the built-in coarsening (the aggregates: blocks of the block smoother)
    if (!coarsest) {
      auto h0 = std::chrono::steady_clock::now();
      if (!injected.empty()) {
        L.hP = injected[l];
        if (L.hP.nrows != L.n || L.hP.bh != L.b) throw Error("injected prolongation " + std::to_string(l) + " does not match the level");
      } else {
        // elasticity: displacement-only fine level (b = dim) maps to displacement+rotation coarse levels
        int bc = L.b;
        if (elast && l == 0 && L.b == dim) bc = (dim == 3) ? 6 : 3;
        std::vector<double> cxyz;
        build_prolongation(L.hA, L.free_mask.empty() ? nullptr : L.free_mask.data(), bc, L.xyz, copt, L.hP, vmap, cxyz);
        if (L.hP.ncols == 0 || L.hP.ncols > 0.8 * L.n) { coarsest = true; vmap.clear(); }  // coarsening stalled
        else {
          auto nl = std::make_unique<Level>();
          nl->xyz = std::move(cxyz);
          lev.push_back(std::move(nl));
        }
      }
      host_s += tick(h0);
      if (verbose) std::fprintf(stderr, "[ngsamg_b200] level %d: n=%lld nnz=%lld  prolongation %.2f s (nc=%lld)\n", l, (long long)L.n, (long long)L.nnz, tick(h0), (long long)L.hP.ncols);
    }
    // smoother options for this level (SpecOpt semantics)
    {
      const std::string smt = flags.spec("sm_type", l, "gs");
      if (smt == "gs") L.sm_type = SM_GS;
      else if (smt == "jacobi") L.sm_type = SM_JACOBI;
      else if (smt == "bgs") L.sm_type = (coarsest || vmap.empty()) ? SM_GS : SM_BGS;   // no coarse map -> GS (SelectSmoother, amg_pc_vertex_impl.hpp:572-585)
      else throw Error("sm_type=" + smt + " is not supported by the B200 path (gs | jacobi | bgs)");
      L.sm_steps = std::max(1, std::atoi(flags.spec("sm_steps", l, "1").c_str()));
      const std::string sy = flags.spec("sm_symm", l, "0");
      L.sm_symm = (sy == "1" || sy == "True" || sy == "true");
      L.omega = (L.sm_type == SM_JACOBI) ? flags.num("sm_omega", 0.9) : 1.0;
      L.pinv = regularize;
    }
    DevCsr dAc;
    if (!coarsest) {
      // Galerkin product on the device: A_{l+1} = (P^T A_l) P
      if (injected.empty() == false && l + 1 >= (int)lev.size()) lev.push_back(std::make_unique<Level>());
      // P travels to the device once; transpose (K12) and both products (K10/K11) run there and are timed alone
      DevCsr dP, dPT, dPTA;
      dev_csr_upload(L.hP, dP, st);
      cudaEventRecord(ev0, st);
      dev_transpose(dP, dPT, st, &launches);
      dev_spgemm(dPT, dA, dPTA, st, &launches);
      dev_spgemm(dPTA, dP, dAc, st, &launches);
      cudaEventRecord(ev1, st);
      cudaEventSynchronize(ev1);
      float ms = 0;
      cudaEventElapsedTime(&ms, ev0, ev1);
      rap_ms += ms;
      // compulsory traffic of the triple product: read A_f, P and P^T once, write A_c once (M_f + 2 P + M_c)
      rap_bytes += (double)dA.nnz * (8.0 * dA.bs() + 4) + 8.0 * (dA.nrows + 1) + 2.0 * ((double)dP.nnz * (8.0 * dP.bs() + 4) + 8.0 * (dP.nrows + 1)) +
                   (double)dAc.nnz * (8.0 * dAc.bs() + 4) + 8.0 * (dAc.nrows + 1);
      dev_csr_free(dPTA); dev_csr_free(dP); dev_csr_free(dPT);
      L.nc = L.hP.ncols; L.bc = L.hP.bw;
      dev_csr_download(dAc, lev[l + 1]->hA, st, true);
      if (injected.empty() && flags.flag("b200_color_coarse", true)) {
        // colour-major coarse numbering (shallow Gauss-Seidel dependency DAG on the coarse level)
        auto h0 = std::chrono::steady_clock::now();
        std::vector<i32> cperm;
        int ncol = 0;
        greedy_coloring_perm(lev[l + 1]->hA, cperm, ncol);
        permute_symmetric(lev[l + 1]->hA, cperm);
        renumber_columns(L.hP, cperm);
        for (i32 &c : vmap) if (c >= 0) c = cperm[c];
        std::vector<double> &cx = lev[l + 1]->xyz;
        if (!cx.empty()) {
          std::vector<double> nx(cx.size());
          for (i64 i = 0; i < L.nc; i++) for (int k = 0; k < 3; k++) nx[(i64)cperm[i] * 3 + k] = cx[i * 3 + k];
          cx.swap(nx);
        }
        dev_csr_free(dAc);
        dev_csr_upload(lev[l + 1]->hA, dAc, st);
        host_s += tick(h0);
        if (verbose) std::fprintf(stderr, "[ngsamg_b200] level %d: colouring + renumbering of the coarse level %.2f s (%d colours)\n", l, tick(h0), ncol);
      }
    }
    {
      auto h0 = std::chrono::steady_clock::now();
      if (!coarsest && l == 0 && flags.str("b200_sm_order", "natural") == "multicolor") {
        // OPTIONAL multicolour Gauss-Seidel on the fine level: sweep in colour-major order instead of the caller's numbering
        // (a different smoother than the reference's natural-order sweep; reported separately)
        int ncol = 0;
        greedy_coloring_perm(L.hA, L.sweep_rank, ncol);
      }
      if (L.sm_type == SM_BGS) {
        // block Gauss-Seidel: the blocks are the aggregates; the shadow level of A~ = DB^-1 A fixes the numbering of this level
        L.gs_block = vmap;
        setup_bgs(L, dA);
      } else {
      // tile-major numbering + two-level schedule for big scalar levels with a deep sweep DAG (setup_tiles)
      if (!coarsest) setup_tiles(L, l, L.hA);
      if (!L.tiled) level_schedule(L.hA, L.mask(), !coarsest, L, st);
      }
      L.d_err = d_err;
      host_s += tick(h0);
      if (verbose) std::fprintf(stderr, "[ngsamg_b200] level %d: level schedule %.2f s (depth %d)\n", l, tick(h0), L.depth);
    }
    if (!coarsest) { build_level_layout(L, dA); if (L.tiled && L.tile_maxs > 2) { prepare_ctile(L); prepare_itile(L); } }
    else {
      L.d_perm = upload_vec(L.perm, st);
      if (clev == "inv") build_coarse_inverse(L);
    }
    if (L.sm_type == SM_BGS) {   // DB^-1 in the level's numbering (parked in hM by setup_bgs)
      build_plain_sell(L.hM, L.d_perm, L.d_perm, L.npad, L.DBI, st, &launches);
      L.hM = HostBsr();
    }
    alloc_vectors(L);
    dev_csr_free(dA);
    if (coarsest) { lev.resize(l + 1); break; }
    dA = dAc;
  }
  for (size_t l = 0; l + 1 < lev.size(); l++) build_transfer_layout(*lev[l], *lev[l + 1]);
  // host copies of big matrices are dropped (introspection then only reports sizes)
  for (auto &lp : lev)
    if ((double)lp->hA.nnz() * lp->hA.bs() > flags.num("keep_host_nnz", 4e8)) { lp->keep_host = false; HostBsr().rowptr.swap(lp->hA.rowptr); std::vector<i32>().swap(lp->hA.col); std::vector<double>().swap(lp->hA.val); }
  d_dot = dev_alloc<double>(4);
  d_partial = dev_alloc<double>(DOT_BLOCKS);
  NGB_CUDA(cudaStreamSynchronize(st));
  finalized = true;
  ms_setup = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
  ms_rap = rap_ms;
  bytes_rap = rap_bytes;
  ms_host = host_s * 1e3;
}


// ------------------------------------------------------------------------------------------------
// multi-rank setup.  Per distributed level: assembled local matrix -> class-respecting coarsening (purely local P, identical
// rows on all sharers) -> local Galerkin product on the device (the coarse matrix is again a distributed sum) -> hybrid
// split M/G + modified diagonal -> the single-rank layout machinery on M with the hybrid stage order as sweep order.
// When the global size falls below ngs_amg_b200_ctr_nv (or max_levels is hit) the level is contracted onto rank 0, which
// continues with a serial hierarchy (CtrMap, dof_contract.cpp; the reference's redistribution policy base_factory.cpp:573-682
// is replaced by this one threshold -- on one NVSwitch box 8 -> 1 is the only step worth taking).
// ------------------------------------------------------------------------------------------------
namespace {
void build_plain_sell(const HostBsr &H, const i32 *d_rperm, const i32 *d_cperm, i64 nrows_pad, Sell &S, cudaStream_t st, i64 *launches)
{
  DevCsr d;
  dev_csr_upload(H, d, st);
  i32 *len = dev_alloc<i32>(nrows_pad);
  NGB_CUDA(cudaMemsetAsync(len, 0, sizeof(i32) * nrows_pad, st));
  if (H.nrows) k_layout_count<<<nblk(H.nrows), TB, 0, st>>>(H.nrows, d.rowptr, d.col, d_rperm, d_cperm, 0, len, nullptr, 0, nullptr);
  build_sell(nrows_pad, H.bh, H.bw, len, S, st, launches);
  if (H.nrows) k_layout_fill<<<nblk(H.nrows), TB, 0, st>>>(H.nrows, H.bh * H.bw, d.rowptr, d.col, d.val, d_rperm, d_cperm, 0, S.slice_ptr, S.col, S.val,
                                                          nullptr, nullptr, nullptr, nullptr, 0, nullptr, nullptr, nullptr);
  S.nnz = H.nnz();
  NGB_CUDA(cudaStreamSynchronize(st));
  dev_free(len);
  dev_csr_free(d);
}
}  // namespace

void Amg::build_halo(Level &L)
{
  const ParDofs &pd = L.pd;
  const size_t np = pd.peers.size();
  L.peers = pd.peers;
  L.m_off.assign(np + 1, 0);
  L.g_off.assign(np + 1, 0);
  std::vector<i32> mi, gi;
  for (size_t kp = 0; kp < np; kp++) {
    for (i32 d : pd.m_ex[kp]) mi.push_back(L.perm[d]);
    for (i32 d : pd.g_ex[kp]) gi.push_back(L.perm[d]);
    L.m_off[kp + 1] = (i64)mi.size();
    L.g_off[kp + 1] = (i64)gi.size();
  }
  // master side of DIS2CO: per distinct master dof the buffer positions that are added to it, neighbours ascending
  std::vector<std::pair<i32, i64>> pr;
  for (i64 k = 0; k < (i64)mi.size(); k++) pr.emplace_back(mi[k], k);
  std::sort(pr.begin(), pr.end());
  std::vector<i32> mu;
  std::vector<i64> mptr{0}, mpos;
  for (size_t q = 0; q < pr.size(); q++) {
    if (q == 0 || pr[q].first != pr[q - 1].first) { if (q) mptr.push_back((i64)mpos.size()); mu.push_back(pr[q].first); }
    mpos.push_back(pr[q].second);
  }
  if (!pr.empty()) mptr.push_back((i64)mpos.size());
  L.n_mu = (i64)mu.size();
  L.d_m_idx = upload_vec(mi, st);
  L.d_g_idx = upload_vec(gi, st);
  L.d_mu_dof = upload_vec(mu, st);
  L.d_mu_ptr = upload_vec(mptr, st);
  L.d_mu_pos = upload_vec(mpos, st);
  const size_t cap = (size_t)std::max<i64>(std::max(L.m_off[np], L.g_off[np]), 1) * L.b;
  L.sendbuf = dev_alloc<double>(cap);
  L.recvbuf = dev_alloc<double>(cap);
  if (!nccl) {
    NGB_CUDA(cudaMallocHost((void **)&L.h_send, sizeof(double) * cap));
    NGB_CUDA(cudaMallocHost((void **)&L.h_recv, sizeof(double) * cap));
  }
}

// Peer-memory halo exchange of one distributed level (kernels_p2p.cuh).  Collective over the ranks; on any failure (no peer access, IPC
// not available) every rank falls back to the NCCL path.
void Amg::setup_p2p(Level &L)
{
  struct Msg {
    cudaIpcMemHandle_t handle;
    i64 cap, m_off, g_off;
    i32 slot, pid, np, pad;
    unsigned long long raw;
  };
  const size_t np = L.peers.size();
  const i64 cap = std::max<i64>(std::max(L.m_off[np], L.g_off[np]), 1) * L.b;
  const size_t ctl_off = sizeof(double) * 2 * (size_t)cap;
  const size_t bytes = ctl_off + sizeof(int) * (3 * np + 2);
  double ok = 1.0;
  std::vector<Msg> out(np), in(np);
  try {
    L.p2p_arena = dev_alloc<unsigned char>(bytes);
    NGB_CUDA(cudaMemsetAsync(L.p2p_arena, 0, bytes, st));
    NGB_CUDA(cudaStreamSynchronize(st));
    cudaIpcMemHandle_t h;
    NGB_CUDA(cudaIpcGetMemHandle(&h, L.p2p_arena));
    for (size_t k = 0; k < np; k++) out[k] = Msg{h, cap, L.m_off[k], L.g_off[k], (i32)k, (i32)getpid(), (i32)np, 0, (unsigned long long)(uintptr_t)L.p2p_arena};
  } catch (const Error &) { ok = 0.0; cudaGetLastError(); }
  {
    std::vector<const void *> sp;
    std::vector<void *> rp;
    std::vector<i64> sb(np, (i64)sizeof(Msg)), rb(np, (i64)sizeof(Msg));
    for (size_t k = 0; k < np; k++) { sp.push_back(&out[k]); rp.push_back(&in[k]); }
    comm.exchange_fixed(L.peers, sp, sb, rp, rb);
  }
  std::vector<P2PPeer> tab(np);
  if (ok > 0) {
    for (size_t k = 0; k < np; k++) {
      void *base = nullptr;
      if (in[k].pid == (i32)getpid()) base = (void *)(uintptr_t)in[k].raw;      // ranks that share a process
      else {
        if (cudaIpcOpenMemHandle(&base, in[k].handle, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { ok = 0.0; cudaGetLastError(); break; }
        L.p2p_opened.push_back(base);
      }
      // the neighbour's control words: flag[np'] | ack[np'] | cnt | seq -- this rank's slot in both arrays
      int *ctl = (int *)((unsigned char *)base + sizeof(double) * 2 * (size_t)in[k].cap);
      tab[k] = P2PPeer{(double *)base, ctl + in[k].slot, ctl + in[k].np + in[k].slot, in[k].cap, in[k].m_off, in[k].g_off};
    }
  }
  comm.allreduce_sum(&ok, 1);
  if (ok < comm.size()) {
    for (void *q : L.p2p_opened) cudaIpcCloseMemHandle(q);
    L.p2p_opened.clear();
    dev_free(L.p2p_arena);
    L.p2p = false;
    return;
  }
  L.d_p2p_peer = upload_vec(tab, st);
  L.d_m_off = upload_vec(L.m_off, st);
  L.d_g_off = upload_vec(L.g_off, st);
  int *ctl = (int *)(L.p2p_arena + ctl_off);
  L.p2p_view = P2PView{(int)np, L.d_p2p_peer, (double *)L.p2p_arena, cap, ctl, ctl + np, ctl + 2 * np, ctl + 3 * np + 1, L.d_m_off, L.d_g_off, d_err};
  L.p2p = true;
}

// One neighbour exchange of doubles on the library stream.  NCCL point-to-point (NVLink) when a communicator was given, else
// staged through pinned host memory and the caller's exchange callback.
void Amg::dev_exchange(const std::vector<i32> &peers, const double *sendbuf, const std::vector<i64> &soff, double *recvbuf,
                       const std::vector<i64> &roff, double *h_send, double *h_recv)
{
  const size_t np = peers.size();
  exchanges++;
  if (nccl) {
    if (np == 0) return;
    NcclApi &api = nccl_api();
    NGB_NCCL(api.GroupStart());
    for (size_t k = 0; k < np; k++) {
      const i64 sc = soff[k + 1] - soff[k], rc = roff[k + 1] - roff[k];
      if (sc > 0) NGB_NCCL(api.Send(sendbuf + soff[k], (size_t)sc, ncclDouble, peers[k], nccl, st));
      if (rc > 0) NGB_NCCL(api.Recv(recvbuf + roff[k], (size_t)rc, ncclDouble, peers[k], nccl, st));
    }
    NGB_NCCL(api.GroupEnd());
    return;
  }
  std::vector<i32> ap;
  std::vector<const void *> sp;
  std::vector<void *> rp;
  std::vector<i64> sb, rb;
  if (np && soff[np] > 0) NGB_CUDA(cudaMemcpyAsync(h_send, sendbuf, sizeof(double) * soff[np], cudaMemcpyDeviceToHost, st));
  NGB_CUDA(cudaStreamSynchronize(st));
  for (size_t k = 0; k < np; k++) {
    const i64 sc = soff[k + 1] - soff[k], rc = roff[k + 1] - roff[k];
    if (sc == 0 && rc == 0) continue;
    ap.push_back(peers[k]);
    sp.push_back(h_send + soff[k]); sb.push_back(sc * (i64)sizeof(double));
    rp.push_back(h_recv + roff[k]); rb.push_back(rc * (i64)sizeof(double));
  }
  comm.exchange_fixed(ap, sp, sb, rp, rb);
  if (np && roff[np] > 0) NGB_CUDA(cudaMemcpyAsync(recvbuf, h_recv, sizeof(double) * roff[np], cudaMemcpyHostToDevice, st));
}

static std::vector<i64> scaled(const std::vector<i64> &off, int b)
{
  std::vector<i64> r(off);
  for (auto &x : r) x *= b;
  return r;
}

void Amg::dis2co(Level &L, double *v)
{
  NGB_PHASE(PH_EXCHANGE);
  const size_t np = L.peers.size();
  if (np == 0) { exchanges++; if (!nccl) NGB_CUDA(cudaStreamSynchronize(st)); return; }
  const i64 ng = L.g_off[np], nm = L.m_off[np];
  if (L.p2p) {
    exchanges++;
    k_p2p_push<<<(unsigned)(np * P2P_CPB), P2P_THREADS, 0, st>>>(L.p2p_view, 0, L.b, L.d_g_idx, v);
    k_p2p_pull<<<std::max(1u, nblk(L.n_mu * L.b, P2P_THREADS)), P2P_THREADS, 0, st>>>(L.p2p_view, 0, L.b, L.n_mu, L.d_mu_dof, L.d_mu_ptr, L.d_mu_pos, v);
    launches += 2;
    return;
  }
  if (ng) { k_halo_pack<<<nblk(ng * L.b), TB, 0, st>>>(ng, L.b, L.d_g_idx, v, L.sendbuf, 1); launches++; }
  dev_exchange(L.peers, L.sendbuf, scaled(L.g_off, L.b), L.recvbuf, scaled(L.m_off, L.b), L.h_send, L.h_recv);
  if (nm) { k_halo_add<<<nblk(L.n_mu * L.b), TB, 0, st>>>(L.n_mu, L.b, L.d_mu_dof, L.d_mu_ptr, L.d_mu_pos, L.recvbuf, v); launches++; }
}

void Amg::co2cu(Level &L, double *v)
{
  NGB_PHASE(PH_EXCHANGE);
  const size_t np = L.peers.size();
  if (np == 0) { exchanges++; if (!nccl) NGB_CUDA(cudaStreamSynchronize(st)); return; }
  const i64 ng = L.g_off[np], nm = L.m_off[np];
  if (L.p2p) {
    exchanges++;
    k_p2p_push<<<(unsigned)(np * P2P_CPB), P2P_THREADS, 0, st>>>(L.p2p_view, 1, L.b, L.d_m_idx, v);
    k_p2p_pull<<<std::max(1u, nblk(ng * L.b, P2P_THREADS)), P2P_THREADS, 0, st>>>(L.p2p_view, 1, L.b, ng, L.d_g_idx, nullptr, nullptr, v);
    launches += 2;
    return;
  }
  if (nm) { k_halo_pack<<<nblk(nm * L.b), TB, 0, st>>>(nm, L.b, L.d_m_idx, v, L.sendbuf, 0); launches++; }
  dev_exchange(L.peers, L.sendbuf, scaled(L.m_off, L.b), L.recvbuf, scaled(L.g_off, L.b), L.h_send, L.h_recv);
  if (ng) { k_halo_set<<<nblk(ng * L.b), TB, 0, st>>>(ng, L.b, L.d_g_idx, L.recvbuf, v); launches++; }
}

void Amg::allreduce_scalars(double *h, int n)
{
  if (!par) return;
  comm.allreduce_sum(h, n);
}

// Level npar: CtrMap::TransferF2C (members send their local DISTRIBUTED vector to the group master, which adds them through
// the dof maps), the serial V-cycle of the contracted hierarchy on rank 0, CtrMap::TransferC2F (the master sends every member
// its CUMULATED values back)   -- dof_contract.cpp:49-228.
void Amg::contracted_solve(Level &L)
{
  NGB_PHASE(PH_COARSE);
  const int R = comm.size(), me = comm.rank();
  const i64 nl = L.n * L.b;
  if (me != 0) {
    std::vector<i32> p0{0};
    std::vector<i64> so{0, nl}, ro{0, 0};
    dev_exchange(p0, L.rhs, so, ctr_buf, ro, h_ctr, h_ctr);
    std::vector<i64> so2{0, 0}, ro2{0, nl};
    dev_exchange(p0, ctr_buf, so2, L.x, ro2, h_ctr, h_ctr);
    L.result = L.x;
    return;
  }
  std::vector<i32> others;
  for (int r = 1; r < R; r++) others.push_back(r);
  std::vector<i64> zero(R, 0), off(ctr_off.begin() + 1, ctr_off.end());   // segments of ranks 1..R-1 (rank 0 is local)
  for (auto &x : off) x -= ctr_off[1];
  dev_exchange(others, ctr_buf, zero, ctr_buf, off, h_ctr, h_ctr);
  Level &N0 = *nested->lev[0];
  NGB_CUDA(cudaMemsetAsync(N0.rhs, 0, sizeof(double) * N0.npad * N0.b, st));
  for (int r = 0; r < R; r++) {
    const i64 nr = ctr.n_local[r];
    if (!nr) continue;
    const double *src = (r == 0) ? L.rhs : ctr_buf + (ctr_off[r] - ctr_off[1]);
    k_ctr_scatter_add<<<nblk(nr * L.b), TB, 0, st>>>(nr, L.b, d_ctr_map[r], N0.d_perm, src, N0.rhs);
    launches++;
  }
  nested->vcycle();
  for (int r = 0; r < R; r++) {
    const i64 nr = ctr.n_local[r];
    if (!nr) continue;
    double *dst = (r == 0) ? L.x : ctr_buf + (ctr_off[r] - ctr_off[1]);
    k_ctr_gather<<<nblk(nr * L.b), TB, 0, st>>>(nr, L.b, d_ctr_map[r], N0.d_perm, N0.result, dst);
    launches++;
  }
  dev_exchange(others, ctr_buf, off, ctr_buf, zero, h_ctr, h_ctr);
  L.result = L.x;
}

void Amg::finalize_parallel()
{
  auto t0 = std::chrono::steady_clock::now();
  double host_s = 0, rap_ms = 0, rap_bytes = 0;
  auto tick = [](std::chrono::steady_clock::time_point t) { return std::chrono::duration<double>(std::chrono::steady_clock::now() - t).count(); };
  const bool verbose = flags.str("log_level", "none") != "none";
  const int me = comm.rank(), R = comm.size();
  const int max_levels = (int)flags.num("max_levels", 10);
  {
    const std::string cyc = flags.str("mg_cycle", "V");
    if (cyc != "V" && cyc != "v") throw Error("mg_cycle=" + cyc + ": the multi-rank path runs V-cycles only");
  }
  const bool elast = type.find("elast") != std::string::npos;
  const int dim = (type.find("2d") != std::string::npos) ? 2 : 3;
  const bool regularize = flags.flag("regularize_cmats", elast);
  regularize_coarse_dim = (regularize && elast) ? (type.find("2d") != std::string::npos ? 2 : 3) : 0;
  const i64 ctr_nv = (i64)flags.num("b200_ctr_nv", 1000000);   // contract onto rank 0 once the GLOBAL level has at most this many vertices
  CoarsenOptions copt;
  copt.max_per_row = (int)flags.num("sp_max_per_row", elast ? 1 + dim : 3);
  copt.min_frac = flags.num("sp_min_frac", dim == 3 ? 0.08 : 0.1);
  copt.omega = flags.num("sp_omega", 1.0);
  copt.smooth = flags.str("prol_type", "semi_aux_smoothed") != "piecewise";
  copt.rounds = (int)flags.num("spw_rounds", 3);
  use_graph = flags.flag("b200_cuda_graph_par", nccl != nullptr) && nccl;   // NCCL halo exchanges are captured into the V-cycle graph

  d_err = dev_alloc<int>(1);
  NGB_CUDA(cudaMemsetAsync(d_err, 0, sizeof(int), st));
  DevCsr dA;
  dev_csr_upload(lev[0]->hA, dA, st);
  bool force_contract = false;
  for (int l = 0;; l++) {
    Level &L = *lev[l];
    L.n = L.hA.nrows; L.b = L.hA.bh; L.nnz = L.hA.nnz();
    L.par = true;
    const ParDofs &pd = L.pd;
    i64 nmaster = 0;
    for (i64 d = 0; d < L.n; d++) nmaster += pd.master_of[d] < 0 ? 1 : 0;
    const i64 nglobal = comm.allreduce_sum(nmaster);
    const bool contract = force_contract || (l > 0 && nglobal <= ctr_nv) || (l + 1 >= max_levels) || R == 1;
    if (verbose) std::fprintf(stderr, "[ngsamg_b200 r%d] level %d: n_local=%lld n_global=%lld%s\n", me, l, (long long)L.n, (long long)nglobal, contract ? "  -> contracted onto rank 0" : "");
    if (contract) {
      // ---- contraction: rank 0 receives the level, everybody keeps a stub level holding its local vectors
      auto h0 = std::chrono::steady_clock::now();
      contract_to_root(comm, pd, L.hA, L.free_mask, L.xyz, ctr);
      if (me == 0 && l > 0 && flags.flag("b200_color_coarse", true) && ctr.A.nrows > 0) {
        // the numbering of the merged level is ours to choose (like every coarse numbering): colour-major, so that its sequential
        // sweep has a shallow dependency DAG; the dof maps of the contraction follow
        std::vector<i32> cperm;
        int ncol = 0;
        greedy_coloring_perm(ctr.A, cperm, ncol);
        permute_symmetric(ctr.A, cperm);
        for (auto &m : ctr.dof_maps) for (i32 &d : m) d = cperm[d];
        if (!ctr.free_mask.empty()) { std::vector<uint8_t> t(ctr.free_mask.size()); for (size_t i = 0; i < t.size(); i++) t[cperm[i]] = ctr.free_mask[i]; ctr.free_mask.swap(t); }
        if (!ctr.xyz.empty()) { std::vector<double> t(ctr.xyz.size()); for (size_t i = 0; i < t.size() / 3; i++) for (int k = 0; k < 3; k++) t[(size_t)cperm[i] * 3 + k] = ctr.xyz[i * 3 + k]; ctr.xyz.swap(t); }
      }
      host_s += tick(h0);
      dev_csr_free(dA);
      L.par = false;
      L.npad = std::max<i64>(round32(L.n), 32);
      L.perm.resize(L.n);
      for (i64 i = 0; i < L.n; i++) L.perm[i] = (i32)i;
      L.d_perm = upload_vec(L.perm, st);
      L.depth = 0;
      alloc_vectors(L);
      lev.resize(l + 1);
      npar = l;
      std::vector<double> sizes(R, 0.0);
      sizes[me] = (double)(L.n * L.b);
      comm.allreduce_sum(sizes.data(), R);
      ctr_off.assign(R + 1, 0);
      for (int r = 0; r < R; r++) ctr_off[r + 1] = ctr_off[r] + (i64)sizes[r];
      const size_t cap = (size_t)std::max<i64>(me == 0 ? ctr_off[R] - ctr_off[1] : L.n * L.b, 1);
      ctr_buf = dev_alloc<double>(cap);
      if (!nccl) NGB_CUDA(cudaMallocHost((void **)&h_ctr, sizeof(double) * cap));
      if (me == 0) {
        ctr.n_local.resize(R);
        for (int r = 0; r < R; r++) { d_ctr_map.push_back(upload_vec(ctr.dof_maps[r], st)); }
        nested = std::make_unique<Amg>();
        Amg &N = *nested;
        N.type = type; N.device = device; N.flags = flags;
        N.flags.set("max_levels", std::to_string(std::max(1, max_levels - l)));
        N.flags.set("mg_cycle", "V");
        N.st = st; N.owns_stream = false;
        NGB_CUDA(cudaEventCreate(&N.ev0));
        NGB_CUDA(cudaEventCreate(&N.ev1));
        N.num_sms = num_sms; N.use_graph = flags.flag("b200_cuda_graph", true) && !use_graph;
        N.tri_sleep_ns = tri_sleep_ns; N.tri_ctas_per_sm = tri_ctas_per_sm; N.tri_prepoll = tri_prepoll; N.tri_gate_all = tri_gate_all;
        N.tri_rm = tri_rm; N.tri_rm_rows_per_warp = tri_rm_rows_per_warp; N.tri_rm_gate_rows = tri_rm_gate_rows; N.tri_rm_max_rows = tri_rm_max_rows; N.tri_block_warp_rows = tri_block_warp_rows; N.rm_spmv_rows = rm_spmv_rows; N.tri_small_rows = tri_small_rows; N.tri_gate_gap_levels = tri_gate_gap_levels; N.tri_level_launch_depth = tri_level_launch_depth; N.tri_level_pdl = tri_level_pdl;
        N.tri_level_launch_rows = tri_level_launch_rows; N.tri_repoll_ns = tri_repoll_ns; N.tri_regate = tri_regate; N.tri_split = tri_split;
        auto NL = std::make_unique<Level>();
        NL->hA = std::move(ctr.A);
        NL->free_mask = ctr.free_mask;
        bool all = true;
        for (auto f : NL->free_mask) all &= (f != 0);
        if (all) NL->free_mask.clear();
        NL->xyz = ctr.xyz;
        N.lev.push_back(std::move(NL));
        N.finalize();
        launches += N.launches;
        rap_ms += N.ms_rap;
        rap_bytes += N.bytes_rap;
        host_s += N.ms_host * 1e-3;
      }
      break;
    }
    // ---- distributed level
    auto h0 = std::chrono::steady_clock::now();
    HostBsr Acum;
    cumulate_matrix(comm, pd, L.hA, Acum);
    std::vector<double> rowsum;
    assembled_row_sums(comm, pd, L.hA, rowsum);
    int bc = L.b;
    if (elast && l == 0 && L.b == dim) bc = (dim == 3) ? 6 : 3;
    std::vector<i32> vmap;
    std::vector<double> cxyz;
    ParCoarsen pc;
    pc.pd = &pd; pc.rowsum = &rowsum; pc.rank = me;
    build_prolongation(Acum, L.free_mask.empty() ? nullptr : L.free_mask.data(), bc, L.xyz, copt, L.hP, vmap, cxyz, &pc);
    {
      // coarsening stalled anywhere?  then contract this level instead (collective decision)
      i64 ncm = 0;
      std::vector<uint8_t> seen(L.hP.ncols, 0);
      for (i64 v = 0; v < L.n; v++) if (vmap[v] >= 0 && pd.master_of[v] < 0 && !seen[vmap[v]]) { seen[vmap[v]] = 1; ncm++; }
      const i64 ncg = comm.allreduce_sum(ncm);
      if (ncg == 0 || (double)ncg > 0.8 * (double)nglobal) { force_contract = true; l--; host_s += tick(h0); continue; }
    }
    auto nl = std::make_unique<Level>();
    nl->xyz = std::move(cxyz);
    coarse_pardofs(pd, vmap, L.hP.ncols, nl->pd, me);
    lev.push_back(std::move(nl));
    Level &C = *lev[l + 1];
    // hybrid split + modified diagonal + stage order
    hybrid_split(pd, L.hA, Acum, L.hM, L.hG);
    hybrid_mod_diag(comm, pd, Acum, L.hG, L.free_mask.empty() ? nullptr : L.free_mask.data(), L.mod_diag);
    hybrid_sweep_order(pd, L.free_mask.empty() ? nullptr : L.free_mask.data(), L.sweep_rank, L.gs_mask, nullptr);
    L.nnz_m = L.hM.nnz(); L.nnz_g = L.hG.nnz();
    HostBsr().rowptr.swap(Acum.rowptr); std::vector<i32>().swap(Acum.col); std::vector<double>().swap(Acum.val);
    host_s += tick(h0);
    {
      const std::string smt = flags.spec("sm_type", l, "gs");
      if (smt != "gs") throw Error("multi-rank levels support sm_type=gs only (BuildJacobiSmoother throws in parallel, amg_pc.cpp:1220-1223)");
      L.sm_type = SM_GS;
      L.sm_steps = std::max(1, std::atoi(flags.spec("sm_steps", l, "1").c_str()));
      const std::string sy = flags.spec("sm_symm", l, "0");
      L.sm_symm = (sy == "1" || sy == "True" || sy == "true");
      L.omega = 1.0;
      L.pinv = regularize;
    }
    // Galerkin product of the DISTRIBUTED local matrix: A_{l+1}^loc = P^T A_l^loc P (P rows are identical on all sharers)
    DevCsr dAc;
    {
      // P travels to the device once; transpose (K12) and both products (K10/K11) run there and are timed alone
      DevCsr dP, dPT, dPTA;
      dev_csr_upload(L.hP, dP, st);
      cudaEventRecord(ev0, st);
      dev_transpose(dP, dPT, st, &launches);
      dev_spgemm(dPT, dA, dPTA, st, &launches);
      dev_spgemm(dPTA, dP, dAc, st, &launches);
      cudaEventRecord(ev1, st);
      cudaEventSynchronize(ev1);
      float ms = 0;
      cudaEventElapsedTime(&ms, ev0, ev1);
      rap_ms += ms;
      // compulsory traffic of the triple product: read A_f, P and P^T once, write A_c once (M_f + 2 P + M_c)
      rap_bytes += (double)dA.nnz * (8.0 * dA.bs() + 4) + 8.0 * (dA.nrows + 1) + 2.0 * ((double)dP.nnz * (8.0 * dP.bs() + 4) + 8.0 * (dP.nrows + 1)) +
                   (double)dAc.nnz * (8.0 * dAc.bs() + 4) + 8.0 * (dAc.nrows + 1);
      dev_csr_free(dPTA); dev_csr_free(dP); dev_csr_free(dPT);
      L.nc = L.hP.ncols; L.bc = L.hP.bw;
      dev_csr_download(dAc, C.hA, st, true);
    }
    if (flags.flag("b200_color_coarse", true) && C.hA.nrows > 0) {
      // local renumbering of the coarse level (colour-major, see the single-rank path); the sharing lists follow
      auto h1 = std::chrono::steady_clock::now();
      std::vector<i32> cperm;
      int ncol = 0;
      // interior colour-major; shared dofs inside their class block by the master's colour (same order on every sharer)
      parallel_coloring_perm(comm, C.pd, C.hA, cperm, ncol);
      permute_symmetric(C.hA, cperm);
      renumber_columns(L.hP, cperm);
      permute_pardofs(C.pd, cperm);
      for (const auto &l : C.pd.ex)
        for (size_t k = 1; k < l.size(); k++)
          if (l[k] <= l[k - 1]) throw Error("coarse-level exchange dofs are not ascending after the renumbering");
      if (!C.xyz.empty()) {
        std::vector<double> nx(C.xyz.size());
        for (i64 i = 0; i < L.nc; i++) for (int k = 0; k < 3; k++) nx[(i64)cperm[i] * 3 + k] = C.xyz[i * 3 + k];
        C.xyz.swap(nx);
      }
      dev_csr_free(dAc);
      dev_csr_upload(C.hA, dAc, st);
      host_s += tick(h1);
    }
    {
      auto h1 = std::chrono::steady_clock::now();
      if (!setup_tiles(L, l, L.hM)) level_schedule(L.hM, L.gs_mask, true, L, st);
      L.d_err = d_err;
      host_s += tick(h1);
      if (verbose) std::fprintf(stderr, "[ngsamg_b200 r%d] level %d: hybrid level, nnz(M)=%lld nnz(G)=%lld, sweep depth %d\n", me, l, (long long)L.nnz_m, (long long)L.nnz_g, L.depth);
    }
    {
      DevCsr dM;
      dev_csr_upload(L.hM, dM, st);
      build_level_layout(L, dM);
      if (L.tiled && L.tile_maxs > 2) { prepare_ctile(L); prepare_itile(L); }
      dev_csr_free(dM);
    }
    build_plain_sell(L.hG, L.d_perm, L.d_perm, L.npad, L.G, st, &launches);
    build_halo(L);
    if (nccl && halo_p2p) setup_p2p(L);
    alloc_vectors(L);
    dev_csr_free(dA);
    dA = dAc;
  }
  for (int l = 0; l < npar; l++) build_transfer_layout(*lev[l], *lev[l + 1]);
  for (auto &lp : lev) {
    const bool big = (double)lp->hA.nnz() * lp->hA.bs() > flags.num("keep_host_nnz", 4e8);
    if (big) {
      lp->keep_host = false;
      for (HostBsr *m : {&lp->hA, &lp->hM, &lp->hG}) { HostBsr().rowptr.swap(m->rowptr); std::vector<i32>().swap(m->col); std::vector<double>().swap(m->val); }
    }
  }
  d_dot = dev_alloc<double>(4);
  d_partial = dev_alloc<double>(DOT_BLOCKS);
  NGB_CUDA(cudaStreamSynchronize(st));
  finalized = true;
  ms_setup = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
  ms_rap = rap_ms;
  bytes_rap = rap_bytes;
  ms_host = host_s * 1e3;
}

// ------------------------------------------------------------------------------------------------
// device primitives
// ------------------------------------------------------------------------------------------------
// The sync-free sweeps wait on rows/tiles computed by OTHER CTAs of the same launch: the whole grid has to be co-resident.  A
// cooperative launch makes the driver guarantee that (it refuses grids that cannot be resident and never runs such a grid
// half-resident next to another kernel) -- so two handles sharing a device, MPS neighbours or a host application's own kernels
// cannot dead-lock a sweep.  One process per GPU remains the intended deployment (include/ngsamg_b200.h).
template <class K, class... Args>
static void launch_resident_smem(K kern, int grid, int block, size_t smem, cudaStream_t st, Args... args)
{
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3((unsigned)block);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeCooperative;
  at[0].val.cooperative = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  NGB_CUDA(cudaLaunchKernelEx(&cfg, kern, args...));
}
template <class K, class... Args>
static void launch_resident(K kern, int grid, int block, cudaStream_t st, Args... args)
{
  launch_resident_smem(kern, grid, block, 0, st, args...);
}

template <int B>
void Amg::tri(Level &L, bool backward, bool add_self, bool write_r, const double *rin, const double *self, double *out, double *rout)
{
  const Sell &T = backward ? L.U : L.L;
  if constexpr (B == 1) {
    if (L.tiled && L.tile_maxs > 2) {
      // two-level sweep, one CTA per tile (kernels_ctile.cuh): sentinel-filled output, hint flags, slab fetched by bulk copies
      if (!add_self && !write_r) throw Error("tri: unsupported mode");
      NGB_CUDA(cudaMemsetAsync(out, 0xFF, sizeof(double) * L.npad, st));
      NGB_CUDA(cudaMemsetAsync(L.d_tile_done, 0, sizeof(int) * (size_t)L.ntiles, st));
      CTileParams prm{(i32)L.ntiles, backward ? 1 : 0, backward ? L.d_meta_bwd : L.d_meta_fwd, L.d_row_lvl, backward ? L.d_tile_succ : L.d_tile_pred,
                      L.d_tile_done, tri_sleep_ns, tri_repoll_ns, tri_pollmode, L.tile_cap_slots, d_err, tri_trace};
      if (L.nonfree_pad) {
        if (add_self) k_gs_tile_prefix<true, false><<<nblk(L.nonfree_pad), TB, 0, st>>>(L.nonfree_pad, rin, self, out, rout);
        else k_gs_tile_prefix<false, true><<<nblk(L.nonfree_pad), TB, 0, st>>>(L.nonfree_pad, rin, self, out, rout);
      }
      if (L.itile) {
        ITileParams ip{(i32)L.ntiles, backward ? 1 : 0, L.d_imeta[backward ? 1 : 0], L.d_img[backward ? 1 : 0], backward ? L.d_tile_succ : L.d_tile_pred,
                       L.d_tile_done, tri_sleep_ns, tri_repoll_ns, L.itile_cap, d_err, tri_trace};
        launch_resident_smem(itile_kernel(L.tile_maxs, add_self, itile_minb), L.itile_grid[add_self ? 1 : 0], 128, L.itile_smem, st, rin, self, out, rout, ip);
        launches += 2;
        return;
      }
      launch_resident_smem(ctile_kernel(L.tile_maxs, L.tile_nbuf, add_self), L.ctile_grid[add_self ? 1 : 0], ctile_threads(L.tile_maxs), L.ctile_smem, st, T.view(),
                           (const double *)L.diag, (const double *)L.dinv, rin, self, out, rout, prm);
      launches += 2;
      return;
    }
    if (L.tiled) {
      // EXPERIMENTAL two-level sweep: one warp per tile, flags between tiles (kernels_tile.cuh)
      if (!add_self && !write_r) throw Error("tri: unsupported mode");
      NGB_CUDA(cudaMemsetAsync(out, 0, sizeof(double) * L.npad, st));
      NGB_CUDA(cudaMemsetAsync(L.d_tile_done, 0, sizeof(int) * (size_t)L.ntiles, st));
      TileParams prm{(i32)L.ntiles, backward ? 1 : 0, L.d_tile_slice, L.d_tile_nlev, L.d_row_lvl,
                     backward ? L.d_tile_succ_ptr : L.d_tile_pred_ptr, backward ? L.d_tile_succ : L.d_tile_pred, L.d_tile_done, tri_sleep_ns, d_err};
      auto launch_tile = [&](auto kern, auto pre) {
        int occ = 0;
        NGB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 256, 0));
        const int cap = std::max(1, occ) * num_sms;          // the whole grid must be resident (tiles wait for each other)
        const int grid = (int)std::max<i64>(1, std::min<i64>((L.ntiles + 7) / 8, cap));
        if (L.nonfree_pad) pre<<<nblk(L.nonfree_pad), TB, 0, st>>>(L.nonfree_pad, rin, self, out, rout);
        launch_resident(kern, grid, 256, st, T.view(), (const double *)L.diag, (const double *)L.dinv, rin, self, out, rout, prm);
      };
      if (L.tile_maxs <= 1) {
        if (add_self) launch_tile(k_gs_tile<1, true, false>, k_gs_tile_prefix<true, false>);
        else launch_tile(k_gs_tile<1, false, true>, k_gs_tile_prefix<false, true>);
      } else {
        if (add_self) launch_tile(k_gs_tile<2, true, false>, k_gs_tile_prefix<true, false>);
        else launch_tile(k_gs_tile<2, false, true>, k_gs_tile_prefix<false, true>);
      }
      launches += 2;
      return;
    }
  }
  // sentinel-fill the output: a row is "published" once its entry is no longer the all-ones NaN
  NGB_CUDA(cudaMemsetAsync(out, 0xFF, sizeof(double) * L.npad * L.b, st));
  if (level_launch(L)) {
    // shallow dependency DAG with many rows per level: one plain launch per level (cached gathers, no polling)
    if (!add_self && !write_r) throw Error("tri: unsupported mode");
    if (!add_self && L.nonfree_pad) NGB_CUDA(cudaMemsetAsync(out, 0, sizeof(double) * L.nonfree_pad * L.b, st));
    for (int q = 0; q < L.depth; q++) {
      const int lv = backward ? L.depth - 1 - q : q;
      const i64 r0 = L.level_start[lv], r1 = L.level_start[lv + 1];
      if (r1 <= r0) continue;
      // every colour but the first is launched with programmatic stream serialization behind the previous colour (see k_gs_level)
      if (tri_level_pdl && q > 0) {
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(nblk(r1 - r0));
        cfg.blockDim = dim3(TB);
        cfg.stream = st;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = at;
        cfg.numAttrs = 1;
        const SellView tv = T.view();
        const double *dg = L.diag, *di = L.dinv;
        const i64 nf = L.nonfree_pad;
        if (add_self) NGB_CUDA(cudaLaunchKernelEx(&cfg, k_gs_level<B, true, false, true>, tv, dg, di, rin, self, out, rout, r0, r1, nf));
        else NGB_CUDA(cudaLaunchKernelEx(&cfg, k_gs_level<B, false, true, true>, tv, dg, di, rin, self, out, rout, r0, r1, nf));
      } else if (add_self) k_gs_level<B, true, false, false><<<nblk(r1 - r0), TB, 0, st>>>(T.view(), L.diag, L.dinv, rin, self, out, rout, r0, r1, L.nonfree_pad);
      else k_gs_level<B, false, true, false><<<nblk(r1 - r0), TB, 0, st>>>(T.view(), L.diag, L.dinv, rin, self, out, rout, r0, r1, L.nonfree_pad);
      launches++;
    }
    return;
  }
  if (L.npad <= tri_small_rows || (B > 1 && tri_block_warp_rows)) {
    // small level: warp-per-row variant (chain cost independent of the row width).  Block matrices use it at EVERY size: the
    // thread-per-block-row kernel walks a 3x3 row in chunks of two entries, one gate + one poll round trip each (measured at 3.07 M
    // P2 nodes: 9.5 / 14.6 ms per sweep against 1.7 ms for 0.9 M rows with a warp per row)
    const int sidx = 24 + (B == 1 ? 0 : B == 2 ? 1 : B == 3 ? 2 : 3) * 2 + (add_self ? 1 : 0);
    auto launch_small = [&](auto kern) {
      if (!tri_grid_cap[sidx]) {
        int occ = 0;
        NGB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 256, 0));
        tri_grid_cap[sidx] = std::max(1, occ) * num_sms;
      }
      const i64 want = (L.npad + 8 * 16 - 1) / (8 * 16);   // >= 16 rows per warp
      const int grid = (int)std::max<i64>(1, std::min<i64>(want, tri_grid_cap[sidx]));
      TriParams prm{L.npad / 32, backward ? 1 : 0, tri_sleep_ns, tri_prepoll, tri_gate_all, 0, 0u, 0, L.nonfree_pad, d_err, tri_pollmode, nullptr, nullptr};
      launch_resident(kern, grid, 256, st, T.view(), (const double *)L.diag, (const double *)L.dinv, rin, self, out, rout, prm);
    };
    if (!add_self && !write_r) throw Error("tri: unsupported mode");
    const Rm &R = backward ? L.rmU : L.rmL;
    if (R.ptr && tri_rm) {
      const int ridx = 32 + (B == 1 ? 0 : B == 2 ? 1 : B == 3 ? 2 : 3) * 2 + (add_self ? 1 : 0);
      auto launch_rm = [&](auto kern) {
        if (!tri_grid_cap[ridx]) {
          int occ = 0;
          NGB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, RM_THREADS, 0));
          tri_grid_cap[ridx] = std::max(1, occ) * num_sms;
        }
        constexpr int WPC = RM_THREADS / 32;
        const i64 want = (L.npad + WPC * tri_rm_rows_per_warp - 1) / (WPC * tri_rm_rows_per_warp);
        const int grid = (int)std::max<i64>(1, std::min<i64>(want, tri_grid_cap[ridx]));
        // tiny levels (a handful of warps): every lane polls its own entries, one L2 round trip per dependency level; larger levels gate on
        // the newest dependency first (two round trips on the critical row, but a spinning warp costs one sector per poll instead of ~50)
        const int gate = (tri_prepoll && L.npad > tri_rm_gate_rows) ? 1 : 0;
        const int pf = L.npad <= tri_rm_max_rows ? 1 : 0;   // block values: L2 prefetch one row ahead on the latency-bound (small) levels only
        TriParams prm{L.npad / 32, backward ? 1 : 0, tri_sleep_ns, gate, pf, 0, 0u, 0, L.nonfree_pad, d_err, tri_pollmode, nullptr, tri_trace};
        launch_resident(kern, grid, RM_THREADS, st, R.view(), (const double *)L.diag, (const double *)L.dinv, rin, self, out, rout, prm);
      };
      if (add_self) launch_rm(k_gs_tri_rm<B, true, false>);
      else launch_rm(k_gs_tri_rm<B, false, true>);
      launches += 2;
      return;
    }
    if (add_self) launch_small(k_gs_tri_small<B, true, false>);
    else launch_small(k_gs_tri_small<B, false, true>);
    launches += 2;
    return;
  }
  // register slot cache: smallest of 8/12/16 that covers (almost) all slices of this part
  const int pre = (B == 1) ? (backward ? L.pre_u : L.pre_l) : 0;
  const int idx = ((B == 1 ? 0 : B == 2 ? 1 : B == 3 ? 2 : 3) * 2 + (add_self ? 1 : 0)) * 3 + (pre == 16 ? 2 : pre == 12 ? 1 : 0);
  auto launch = [&](auto kern) {
    if (!tri_grid_cap[idx]) {
      int occ = 0;
      NGB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 256, 0));
      occ = std::max(1, occ);
      if (tri_ctas_per_sm > 0) occ = std::min(occ, tri_ctas_per_sm);
      tri_grid_cap[idx] = occ * num_sms;   // the whole grid must be resident (see k_gs_tri)
    }
    const i64 nslices = L.npad / 32;
    const int grid = (int)std::min<i64>((nslices + 7) / 8, tri_grid_cap[idx]);
    const i64 gap = tri_gate_gap_levels > 0 ? (i64)(tri_gate_gap_levels * (double)L.npad / std::max(1, L.depth)) : 0;
    TriParams prm{nslices, backward ? 1 : 0, tri_sleep_ns, tri_prepoll, tri_gate_all, gap, tri_repoll_ns, tri_regate, L.nonfree_pad, d_err, tri_pollmode, tri_split ? (backward ? L.d_bnd_bwd : L.d_bnd_fwd) : nullptr, tri_trace};
    launch_resident(kern, grid, 256, st, T.view(), (const double *)L.diag, (const double *)L.dinv, rin, self, out, rout, prm);
  };
  if (!add_self && !write_r) throw Error("tri: unsupported mode");
  if constexpr (B == 1) {
    if (add_self) { if (pre == 16) launch(k_gs_tri<1, true, false, 16>); else if (pre == 12) launch(k_gs_tri<1, true, false, 12>); else launch(k_gs_tri<1, true, false, 8>); }
    else { if (pre == 16) launch(k_gs_tri<1, false, true, 16>); else if (pre == 12) launch(k_gs_tri<1, false, true, 12>); else launch(k_gs_tri<1, false, true, 8>); }
  } else {
    if (add_self) launch(k_gs_tri<B, true, false, 0>);
    else launch(k_gs_tri<B, false, true, 0>);
  }
  launches += 2;
}

void Amg::tri_dispatch(Level &L, bool backward, bool add_self, bool write_r, const double *rin, const double *self, double *out, double *rout)
{
  NGB_PHASE(PH_TRI);
  switch (L.b) {
    case 1: tri<1>(L, backward, add_self, write_r, rin, self, out, rout); break;
    case 2: tri<2>(L, backward, add_self, write_r, rin, self, out, rout); break;
    case 3: tri<3>(L, backward, add_self, write_r, rin, self, out, rout); break;
    case 6: tri<6>(L, backward, add_self, write_r, rin, self, out, rout); break;
    default: throw Error("unsupported block size");
  }
}

template <int BH, int BW, bool S2, bool D>
static void launch_spmv(cudaStream_t st, i64 small_rows, i64 npad, const Sell &a, const Sell *b, const double *diag, const double *v, const double *y_in,
                        double *y_out, double alpha, double beta, double *xadd, const Sell *nfp = nullptr, const i32 *rowmap = nullptr)
{
  const SellView none{nullptr, nullptr, nullptr};
  const SellView s3 = (nfp && nfp->slice_ptr) ? nfp->view() : none;
  // warp-per-row SpMV on the SELL layout only where rows are scarce: the lanes of a row read strided (8 useful bytes per 32-byte sector
  // for blocks), so block levels switch to it later (threshold in scalar entries per block row)
  if (npad * (BH * BW > 1 ? (i64)BH * BW / 2 : 1) <= small_rows) {
    const unsigned grid = (unsigned)std::max<i64>(1, std::min<i64>((npad + 7) / 8, 148 * 8));
    k_sell_spmv_small<BH, BW, S2, D><<<grid, TB, 0, st>>>(npad, a.view(), b ? b->view() : a.view(), diag, v, y_in, y_out, alpha, beta,
                                                         xadd, s3, rowmap);
    return;
  }
  // scalar matrices: rows of at most 4 / 8 slots are walked as ONE masked block of batched loads (prolongation; level-0 U-pass);
  // everything else keeps the legacy walk (measured at 311^3: prolong 0.49 -> 0.34 ms, U-pass 0.69 -> 0.62 ms; the (L+D)-pass, the
  // full SpMV and the level-1 passes lose occupancy / pipelining with the batched walk and stay on the legacy one)
  if (BH == 1 && BW == 1 && !S2 && !D && !b && a.maxw > 0 && a.maxw <= 4)
    k_sell_spmv<BH, BW, S2, D, 4><<<nblk(npad), TB, 0, st>>>(npad, a.view(), a.view(), diag, v, y_in, y_out, alpha, beta, xadd, s3, rowmap);
  else if (BH == 1 && BW == 1 && !S2 && !D && !b && !rowmap && a.maxw > 0 && a.maxw <= 8)
    k_sell_spmv<BH, BW, S2, D, 8><<<nblk(npad), TB, 0, st>>>(npad, a.view(), a.view(), diag, v, y_in, y_out, alpha, beta, xadd, s3, rowmap);
  else
    k_sell_spmv<BH, BW, S2, D, 0><<<nblk(npad), TB, 0, st>>>(npad, a.view(), b ? b->view() : a.view(), diag, v, y_in, y_out, alpha, beta, xadd,
                                                            s3, rowmap);
}

template <int B>
static void launch_rm_spmv(cudaStream_t st, Level &L, int which, const double *v, const double *y_in, double *y_out, double alpha, double beta)
{
  const Rm &A1 = (which == 1 || which == 3) ? L.rmU : L.rmL;
  const unsigned grid = (unsigned)std::max<i64>(1, std::min<i64>((L.npad + 7) / 8, 148 * 8));
  if (which == 4) k_rm_spmv<B, true, true><<<grid, TB, 0, st>>>(L.npad, L.rmL.view(), L.rmU.view(), L.diag, v, y_in, y_out, alpha, beta);
  else if (which >= 2) k_rm_spmv<B, false, true><<<grid, TB, 0, st>>>(L.npad, A1.view(), A1.view(), L.diag, v, y_in, y_out, alpha, beta);
  else k_rm_spmv<B, false, false><<<grid, TB, 0, st>>>(L.npad, A1.view(), A1.view(), L.diag, v, y_in, y_out, alpha, beta);
}

void Amg::spmv_part(Level &L, int which, const double *v, const double *y_in, double *y_out, double alpha, double beta, double *xadd)
{
  NGB_PHASE(PH_PASS);
  // block levels with few rows (too few threads for the thread-per-row kernel, and its warp-per-row variant wastes 3/4 of every sector
  // on blocks): warp per row on the row-major copies
  if (L.b > 1 && L.rmL.ptr && L.rmU.ptr && !xadd && !L.N.slice_ptr && L.npad <= rm_spmv_rows) {
    switch (L.b) {
      case 2: launch_rm_spmv<2>(st, L, which, v, y_in, y_out, alpha, beta); break;
      case 3: launch_rm_spmv<3>(st, L, which, v, y_in, y_out, alpha, beta); break;
      case 6: launch_rm_spmv<6>(st, L, which, v, y_in, y_out, alpha, beta); break;
      default: throw Error("unsupported block size");
    }
    launches++;
    return;
  }
  const Sell &A1 = (which == 1 || which == 3) ? L.U : L.L;
  const bool s2 = (which == 4), d = (which >= 2);
#define NGB_SPMV(B)                                                                                                     \
  if (s2) launch_spmv<B, B, true, true>(st, spmv_small_rows, L.npad, L.L, &L.U, L.diag, v, y_in, y_out, alpha, beta, xadd, &L.N);       \
  else if (d) launch_spmv<B, B, false, true>(st, spmv_small_rows, L.npad, A1, nullptr, L.diag, v, y_in, y_out, alpha, beta, xadd, &L.N); \
  else launch_spmv<B, B, false, false>(st, spmv_small_rows, L.npad, A1, nullptr, L.diag, v, y_in, y_out, alpha, beta, xadd);
  switch (L.b) {
    case 1: NGB_SPMV(1) break;
    case 2: NGB_SPMV(2) break;
    case 3: NGB_SPMV(3) break;
    case 6: NGB_SPMV(6) break;
    default: throw Error("unsupported block size");
  }
#undef NGB_SPMV
  launches++;
}

void Amg::transfer(const Sell &S, const double *v, const double *y_in, double *y_out, double alpha, double beta, const i32 *rowmap)
{
  NGB_PHASE(PH_TRANSFER);
  const int key = S.bh * 10 + S.bw;
#define NGB_TR(H, W) launch_spmv<H, W, false, false>(st, spmv_small_rows, S.nrows_pad, S, nullptr, nullptr, v, y_in, y_out, alpha, beta, nullptr, nullptr, rowmap)
  switch (key) {
    case 11: NGB_TR(1, 1); break;
    case 22: NGB_TR(2, 2); break;
    case 33: NGB_TR(3, 3); break;
    case 66: NGB_TR(6, 6); break;
    case 36: NGB_TR(3, 6); break;
    case 63: NGB_TR(6, 3); break;
    case 23: NGB_TR(2, 3); break;
    case 32: NGB_TR(3, 2); break;
    default: throw Error("unsupported transfer block shape");
  }
#undef NGB_TR
  launches++;
}

// BaseSmoother::CalcResiduum (base_smoother.hpp:132-142)
void Amg::calc_residuum(Level &L, const double *x, const double *b, double *res, bool x_zero)
{
  if (x_zero) { NGB_CUDA(cudaMemcpyAsync(res, b, sizeof(double) * L.npad * L.b, cudaMemcpyDeviceToDevice, st)); return; }
  spmv_part(L, 4, x, b, res, -1.0, 1.0, nullptr);
}

// GSS3::SmoothRESInternal (gssmoother.cpp:260-315) in gather form: with delta = (T + D~)^-1 res (T = L forward, U backward),
// x += delta and res -= A delta.  The triangular part runs level-scheduled, the other half is a plain SpMV.
void Amg::gs_res(Level &L, bool backward, double *x, double *res, bool x_zero)
{
  double *d = x_zero ? x : L.tmp;
  tri_dispatch(L, backward, false, true, res, nullptr, d, res);
  spmv_part(L, backward ? 0 : 1, d, res, res, -1.0, 1.0, x_zero ? nullptr : x);
}

// GSS3::SmoothRHSInternal (gssmoother.cpp:195-257): x_i += dinv_i (b_i - A_i x); the not-yet-updated half of the row
// (and the diagonal) is applied first as a plain SpMV, the updated half by the sync-free triangular sweep, which
// writes the new iterate to a second buffer (xout != x) because the output doubles as the dependency flags.
void Amg::gs_rhs(Level &L, bool backward, const double *x, const double *b, double *xout)
{
  spmv_part(L, backward ? 2 : 3, x, b, L.tmp, -1.0, 1.0, nullptr);
  tri_dispatch(L, backward, true, false, L.tmp, x, xout, nullptr);
}

// GSS3::Smooth / SmoothBack (gssmoother.cpp:349-398); RichardsonSmoother::Smooth for Jacobi (base_smoother.cpp:61-83)
void Amg::smooth_once(Level &L, double *x, const double *b, double *res, bool ru, bool ur, bool xz, bool backward)
{
  if (L.sm_type == SM_GS) {
    if (ru) {
      if (ur) gs_res(L, backward, x, res, xz);
      else { if (xz) NGB_CUDA(cudaMemsetAsync(x, 0, sizeof(double) * L.npad * L.b, st)); gs_rhs(L, backward, x, b, L.y); NGB_CUDA(cudaMemcpyAsync(x, L.y, sizeof(double) * L.npad * L.b, cudaMemcpyDeviceToDevice, st)); }
    } else {
      if (ur) { calc_residuum(L, x, b, res, xz); gs_res(L, backward, x, res, xz); }
      else { if (xz) NGB_CUDA(cudaMemsetAsync(x, 0, sizeof(double) * L.npad * L.b, st)); gs_rhs(L, backward, x, b, L.y); NGB_CUDA(cudaMemcpyAsync(x, L.y, sizeof(double) * L.npad * L.b, cudaMemcpyDeviceToDevice, st)); }
    }
  } else if (L.sm_type == SM_BGS) {
    // BSmoother2::SmoothWO (loc_block_gssmoother_impl.hpp:656-669)
    if (xz) NGB_CUDA(cudaMemsetAsync(x, 0, sizeof(double) * L.npad * L.b, st));
    if (ru && ur) bgs_res(L, backward, x, res);
    else {
      bgs_rhs(L, backward, x, b);
      if (ur) calc_residuum(L, x, b, res, false);
    }
  } else {
    if (xz) NGB_CUDA(cudaMemsetAsync(x, 0, sizeof(double) * L.npad * L.b, st));
    const double *src;
    if (!ru && xz) src = b;
    else { if (!ru) calc_residuum(L, x, b, res, false); src = res; }
    switch (L.b) {
      case 1: k_jacobi_update<1><<<nblk(L.npad), TB, 0, st>>>(L.npad, L.omega, L.dinv, src, x); break;
      case 2: k_jacobi_update<2><<<nblk(L.npad), TB, 0, st>>>(L.npad, L.omega, L.dinv, src, x); break;
      case 3: k_jacobi_update<3><<<nblk(L.npad), TB, 0, st>>>(L.npad, L.omega, L.dinv, src, x); break;
      case 6: k_jacobi_update<6><<<nblk(L.npad), TB, 0, st>>>(L.npad, L.omega, L.dinv, src, x); break;
      default: throw Error("unsupported block size");
    }
    launches++;
    if (ur) calc_residuum(L, x, b, res, false);
  }
}

// HybridBaseSmoother::SmoothImplRES (hybrid_base_smoother.cpp:294-404): x CUMULATED, res DISTRIBUTED = b - A x_old.
// The local sweep only sees M, so G x_old is stashed and the residual is corrected with res += G x_old - G x_new.
void Amg::hybrid_smooth_res(Level &L, bool backward, double *x, double *res, bool x_zero)
{
  const bool stash = !x_zero && L.G.nnz;
  double *gx = L.y;
  if (stash) transfer(L.G, x, nullptr, gx, 1.0, 0.0);
  dis2co(L, res);
  gs_res(L, backward, x, res, x_zero);
  co2cu(L, x);
  if (L.G.nnz) {
    if (stash) { k_axpby<<<nblk(L.npad * L.b), TB, 0, st>>>(L.npad * L.b, 1.0, gx, 1.0, res); launches++; }
    transfer(L.G, x, res, res, -1.0, 1.0);
  }
}

// HybridBaseSmoother::SmoothImplRHS (hybrid_base_smoother.cpp:407-446): the local sweep runs against b - G x (DISTRIBUTED -> DIS2CO);
// `work` holds that right-hand side (the reference uses its stashed vector / res as work space)
void Amg::hybrid_smooth_rhs(Level &L, bool backward, double *x, const double *b, double *work, bool x_zero)
{
  const size_t bytes = sizeof(double) * L.npad * L.b;
  if (x_zero) NGB_CUDA(cudaMemsetAsync(x, 0, bytes, st));
  if (!x_zero && L.G.nnz) transfer(L.G, x, b, work, -1.0, 1.0);
  else NGB_CUDA(cudaMemcpyAsync(work, b, bytes, cudaMemcpyDeviceToDevice, st));
  dis2co(L, work);
  gs_rhs(L, backward, x, work, L.y);
  NGB_CUDA(cudaMemcpyAsync(x, L.y, bytes, cudaMemcpyDeviceToDevice, st));
  co2cu(L, x);
}

// HybridBaseSmoother::SmoothImpl (hybrid_base_smoother.cpp:242-290): the cost heuristic that picks the RES or the RHS form
void Amg::hybrid_smooth(Level &L, double *x, const double *b, double *res, bool ru, bool ur, bool xz, bool backward)
{
  if (ur) {
    if (!ru) {
      if (!xz) {
        hybrid_smooth_rhs(L, backward, x, b, res, false);
        spmv_part(L, 4, x, b, res, -1.0, 1.0, nullptr);                 // res = b - (M + G) x   (HybridBaseMatrix, DISTRIBUTED)
        if (L.G.nnz) transfer(L.G, x, res, res, -1.0, 1.0);
      } else {
        NGB_CUDA(cudaMemcpyAsync(res, b, sizeof(double) * L.npad * L.b, cudaMemcpyDeviceToDevice, st));
        hybrid_smooth_res(L, backward, x, res, true);
      }
    } else hybrid_smooth_res(L, backward, x, res, xz);
  } else hybrid_smooth_rhs(L, backward, x, b, res, xz);
}

// ProxySmoother around the hybrid smoother (sm_steps > 1 or sm_symm, amg_pc.cpp:1079-1082)
void Amg::hybrid_level_smooth(Level &L, double *x, const double *b, double *res, bool ru, bool ur, bool xz, bool backward)
{
  const int k = std::max(1, L.sm_steps);
  if (L.sm_symm) {
    hybrid_smooth(L, x, b, res, ru, ur, xz, false);
    hybrid_smooth(L, x, b, res, ur, ur, false, true);
    for (int j = 0; j < k - 1; j++) {
      hybrid_smooth(L, x, b, res, ur, ur, false, false);
      hybrid_smooth(L, x, b, res, ur, ur, false, true);
    }
  } else {
    hybrid_smooth(L, x, b, res, ru, ur, xz, backward);
    for (int j = 0; j < k - 1; j++) hybrid_smooth(L, x, b, res, ur, ur, false, backward);
  }
}

// ProxySmoother (base_smoother.hpp:169-229) + SmoothK / SmoothBackK / SmoothSymmK (:79-112)
void Amg::level_smooth(Level &L, double *x, const double *b, double *res, bool ru, bool ur, bool xz, bool backward)
{
  const int k = std::max(1, L.sm_steps);
  if (L.sm_symm) {
    smooth_once(L, x, b, res, ru, ur, xz, false);
    smooth_once(L, x, b, res, ur, ur, false, true);
    for (int j = 0; j < k - 1; j++) {
      smooth_once(L, x, b, res, ur, ur, false, false);
      smooth_once(L, x, b, res, ur, ur, false, true);
    }
  } else {
    smooth_once(L, x, b, res, ru, ur, xz, backward);
    for (int j = 0; j < k - 1; j++) smooth_once(L, x, b, res, ur, ur, false, backward);
  }
}

// AMGMatrix::SmoothV (amg_matrix.cpp:160-307), single rank: Distribute/Cumulate are no-ops.
void Amg::vcycle_record()
{
  if (par) {
    // AMGMatrix::SmoothV on a distributed hierarchy; the smoother calls are HybridBaseSmoother::SmoothImplRES / SmoothImplRHS
    // (hybrid_base_smoother.cpp:294-446) with CallStageKernelsImpl's stage order folded into the sweep order of M:
    //   pre : res = b (DISTRIBUTED) ; DIS2CO(res) ; forward sweep on M (x = 0) ; CO2CU(x) ; res -= G x ; restrict
    //   post: x += P x_c ; t = b - G x ; DIS2CO(t) ; backward sweep on M against t ; CO2CU(x)
    for (int l = 0; l < npar; l++) {
      Level &L = *lev[l];
      Level &C = *lev[l + 1];
      if (L.sm_symm || L.sm_steps != 1) {
        // general protocol (ProxySmoother): x = 0, res = b, Smooth(x, b, res, true, true, true)
        NGB_CUDA(cudaMemsetAsync(L.x, 0, sizeof(double) * L.npad * L.b, st));
        NGB_CUDA(cudaMemcpyAsync(L.res, L.rhs, sizeof(double) * L.npad * L.b, cudaMemcpyDeviceToDevice, st));
        hybrid_level_smooth(L, L.x, L.rhs, L.res, true, true, true, false);
        transfer(L.PT, L.res, nullptr, C.rhs, 1.0, 0.0, L.d_pt_rowmap);
        continue;
      }
      { NGB_PHASE(PH_OTHER); NGB_CUDA(cudaMemcpyAsync(L.res, L.rhs, sizeof(double) * L.npad * L.b, cudaMemcpyDeviceToDevice, st)); }
      dis2co(L, L.res);
      tri_dispatch(L, false, false, true, L.res, nullptr, L.x, L.res);
      spmv_part(L, 1, L.x, L.res, L.res, -1.0, 1.0, nullptr);
      co2cu(L, L.x);
      if (L.G.nnz) { NGB_PHASE(PH_GSPMV); transfer(L.G, L.x, L.res, L.res, -1.0, 1.0); }
      transfer(L.PT, L.res, nullptr, C.rhs, 1.0, 0.0, L.d_pt_rowmap);
    }
    contracted_solve(*lev[npar]);
    for (int l = npar - 1; l >= 0; l--) {
      Level &L = *lev[l];
      Level &C = *lev[l + 1];
      transfer(L.P, C.result, L.x, L.x, 1.0, 1.0);
      if (L.sm_symm || L.sm_steps != 1) {
        hybrid_level_smooth(L, L.x, L.rhs, L.res, false, false, false, true);
        L.result = L.x;
        continue;
      }
      if (L.G.nnz) { NGB_PHASE(PH_GSPMV); transfer(L.G, L.x, L.rhs, L.res, -1.0, 1.0); }
      else { NGB_PHASE(PH_OTHER); NGB_CUDA(cudaMemcpyAsync(L.res, L.rhs, sizeof(double) * L.npad * L.b, cudaMemcpyDeviceToDevice, st)); }
      dis2co(L, L.res);
      gs_rhs(L, true, L.x, L.res, L.y);
      co2cu(L, L.y);
      L.result = L.y;
    }
    return;
  }
  if (cycle == CYCLE_W) { w_visit(0); return; }
  if (cycle == CYCLE_BS) { bs_record(); return; }
  const int NL = (int)lev.size();
  for (int l = 0; l + 1 < NL; l++) {
    Level &L = *lev[l];
    Level &C = *lev[l + 1];
    // x_l = 0 ; res_l = b_l ; Smooth(x, b, res, true, true, true)     (:193-206)
    // (res_l = b_l is folded into the sweep: the triangular kernel reads rhs and writes res)
    if (L.sm_type == SM_GS && !L.sm_symm && L.sm_steps == 1) {
      tri_dispatch(L, false, false, true, L.rhs, nullptr, L.x, L.res);
      spmv_part(L, 1, L.x, L.res, L.res, -1.0, 1.0, nullptr);
    } else {
      NGB_CUDA(cudaMemcpyAsync(L.res, L.rhs, sizeof(double) * L.npad * L.b, cudaMemcpyDeviceToDevice, st));
      level_smooth(L, L.x, L.rhs, L.res, true, true, true, false);
    }
    // TransferF2C: rhs_{l+1} = P^T res_l     (:212, dof_map.cpp:633-654)
    transfer(L.PT, L.res, nullptr, C.rhs, 1.0, 0.0, L.d_pt_rowmap);
  }
  {
    Level &L = *lev[NL - 1];
    NGB_PHASE(PH_COARSE);
    if (has_cinv) {  // x_L = A_L^-1 rhs_L   (:217-247)
      k_dense_gemv<<<nblk((i64)cinv_n * 32), TB, 0, st>>>(cinv_n, d_cinv, L.rhs, L.x);
      launches++;
    } else NGB_CUDA(cudaMemsetAsync(L.x, 0, sizeof(double) * L.npad * L.b, st));
    L.result = L.x;
  }
  for (int l = NL - 2; l >= 0; l--) {
    Level &L = *lev[l];
    Level &C = *lev[l + 1];
    // AddC2F: x_l += P x_{l+1}   (:263, dof_map.cpp:694-709)
    transfer(L.P, C.result, L.x, L.x, 1.0, 1.0);
    // SmoothBack(x, b, res, false, false, false)   (:302)
    if (L.sm_type == SM_GS && !L.sm_symm && L.sm_steps == 1) {
      gs_rhs(L, true, L.x, L.rhs, L.y);   // the new iterate is produced in the second buffer
      L.result = L.y;
    } else {
      level_smooth(L, L.x, L.rhs, L.res, false, false, false, true);
      L.result = L.x;
    }
  }
}

// ---- W and BS cycles (AMGMatrix::SmoothW / SmoothBS / SmoothVFromLevel, amg_matrix.cpp:37-157, 310-374), single rank.
// Everything goes through the general smoother protocol (level_smooth); iterates live in L.x, residuals in L.res.
void Amg::coarse_solve_record()
{
  Level &L = *lev.back();
  if (has_cinv) { k_dense_gemv<<<nblk((i64)cinv_n * 32), TB, 0, st>>>(cinv_n, d_cinv, L.rhs, L.x); launches++; }
  else NGB_CUDA(cudaMemsetAsync(L.x, 0, sizeof(double) * L.npad * L.b, st));
  L.result = L.x;
}

void Amg::restrict_record(int l, const double *res)
{
  Level &L = *lev[l];
  transfer(L.PT, res, nullptr, lev[l + 1]->rhs, 1.0, 0.0, L.d_pt_rowmap);
}

// the recursion of SmoothW (:46-104).  On level 0 the reference first runs a V-type visit and then the W-type visit below, which
// restarts from x = 0, res = b: the first visit leaves no trace in the result and is not run.
void Amg::w_visit(int l)
{
  const int NL = (int)lev.size();
  if (l + 1 >= NL) { coarse_solve_record(); return; }
  Level &L = *lev[l];
  Level &C = *lev[l + 1];
  const size_t bytes = sizeof(double) * L.npad * L.b;
  NGB_CUDA(cudaMemsetAsync(L.x, 0, bytes, st));
  NGB_CUDA(cudaMemcpyAsync(L.res, L.rhs, bytes, cudaMemcpyDeviceToDevice, st));
  level_smooth(L, L.x, L.rhs, L.res, true, true, true, false);
  restrict_record(l, L.res);
  w_visit(l + 1);
  transfer(L.P, C.x, L.x, L.x, 1.0, 1.0);
  level_smooth(L, L.x, L.rhs, L.res, false, true, false, true);    // SmoothBack(x, b, res, false, true, false)   :82
  level_smooth(L, L.x, L.rhs, L.res, true, true, false, false);    // Smooth(x, b, res, true, true, false)        :83
  restrict_record(l, L.res);
  w_visit(l + 1);
  transfer(L.P, C.x, L.x, L.x, 1.0, 1.0);
  level_smooth(L, L.x, L.rhs, L.res, false, false, false, true);   // :89
  L.result = L.x;
}

// SmoothVFromLevel (:310-374) on the work vectors of level s
void Amg::v_from_level(int s, bool ru, bool ur, bool xz)
{
  const int NL = (int)lev.size();
  Level &S = *lev[s];
  level_smooth(S, S.x, S.rhs, S.res, ru, true, xz, false);
  restrict_record(s, S.res);
  for (int l = s + 1; l + 1 < NL; l++) {
    Level &L = *lev[l];
    const size_t bytes = sizeof(double) * L.npad * L.b;
    NGB_CUDA(cudaMemsetAsync(L.x, 0, bytes, st));
    NGB_CUDA(cudaMemcpyAsync(L.res, L.rhs, bytes, cudaMemcpyDeviceToDevice, st));
    level_smooth(L, L.x, L.rhs, L.res, true, true, true, false);
    restrict_record(l, L.res);
  }
  coarse_solve_record();
  for (int l = NL - 2; l > s; l--) {
    Level &L = *lev[l];
    transfer(L.P, lev[l + 1]->x, L.x, L.x, 1.0, 1.0);
    level_smooth(L, L.x, L.rhs, L.res, false, false, false, true);
  }
  transfer(S.P, lev[s + 1]->x, S.x, S.x, 1.0, 1.0);
  level_smooth(S, S.x, S.rhs, S.res, false, ur, false, true);
}

// SmoothBS (:107-157): every level is smoothed by a V-cycle that starts there
void Amg::bs_record()
{
  const int NL = (int)lev.size();
  for (int l = 0; l + 1 < NL; l++) {
    Level &L = *lev[l];
    const size_t bytes = sizeof(double) * L.npad * L.b;
    NGB_CUDA(cudaMemsetAsync(L.x, 0, bytes, st));
    NGB_CUDA(cudaMemcpyAsync(L.res, L.rhs, bytes, cudaMemcpyDeviceToDevice, st));
    v_from_level(l, true, true, true);
    restrict_record(l, L.res);
  }
  coarse_solve_record();
  for (int l = NL - 2; l >= 0; l--) {
    Level &L = *lev[l];
    transfer(L.P, lev[l + 1]->x, L.x, L.x, 1.0, 1.0);
    v_from_level(l, false, false, false);
    L.result = L.x;
  }
}

void Amg::vcycle()
{
  if (!use_graph) { vcycle_record(); return; }
  if (!vgraph) {
    cudaGraph_t g = nullptr;
    const i64 l0 = launches;
    NGB_CUDA(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
    try { vcycle_record(); }
    catch (...) { cudaStreamEndCapture(st, &g); if (g) cudaGraphDestroy(g); throw; }
    NGB_CUDA(cudaStreamEndCapture(st, &g));
    NGB_CUDA(cudaGraphInstantiate(&vgraph, g, 0));
    cudaGraphDestroy(g);
    vgraph_launches = launches - l0;
    launches = l0;
  }
  NGB_CUDA(cudaGraphLaunch(vgraph, st));
  launches += vgraph_launches;
}

void Amg::check_watchdog()
{
  int e = 0;
  NGB_CUDA(cudaMemcpyAsync(&e, d_err, sizeof(int), cudaMemcpyDeviceToHost, st));
  NGB_CUDA(cudaStreamSynchronize(st));
  if (e) {
    NGB_CUDA(cudaMemsetAsync(d_err, 0, sizeof(int), st));
    throw Error("Gauss-Seidel sweep: dependency wait timed out (watchdog) -- results invalid");
  }
  if (nested) nested->check_watchdog();   // rank 0: the serial hierarchy below the contracted level has its own flag
}

double Amg::dot(i64 n, const double *a, const double *b)
{
  k_dot_partial<<<DOT_BLOCKS, DOT_THREADS, 0, st>>>(n, a, b, d_partial);
  k_dot_final<<<1, DOT_THREADS, 0, st>>>(DOT_BLOCKS, d_partial, d_dot);
  launches += 2;
  double h = 0;
  if (par && nccl) {   // InnerProduct(cumulated, distributed): local dot over all local dofs, summed over the ranks
    NGB_NCCL(nccl_api().AllReduce(d_dot, d_dot, 1, ncclDouble, ncclSum, nccl, st));
    exchanges++;
  }
  NGB_CUDA(cudaMemcpyAsync(&h, d_dot, sizeof(double), cudaMemcpyDeviceToHost, st));
  NGB_CUDA(cudaStreamSynchronize(st));
  if (par && !nccl) allreduce_scalars(&h, 1);
  return h;
}

bool Amg::is_device_ptr(const void *p)
{
  cudaPointerAttributes at;
  cudaError_t e = cudaPointerGetAttributes(&at, p);
  if (e != cudaSuccess) { cudaGetLastError(); return false; }
  return at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged;
}

void Amg::ensure_io(i64 n)
{
  if (n <= io_n) return;
  dev_free(io_a); dev_free(io_b); dev_free(io_c);
  if (pin_a) cudaFreeHost(pin_a);
  if (pin_b) cudaFreeHost(pin_b);
  io_a = dev_alloc<double>(n); io_b = dev_alloc<double>(n); io_c = dev_alloc<double>(n);
  NGB_CUDA(cudaMallocHost((void **)&pin_a, sizeof(double) * n));
  NGB_CUDA(cudaMallocHost((void **)&pin_b, sizeof(double) * n));
  io_n = n; pin_n = n;
}

// returns a device pointer holding p[0..n): p itself if it is device memory, else a staged copy
const double *Amg::to_device(const double *p, i64 n, double *stage)
{
  if (is_device_ptr(p)) return p;
  // pageable -> pinned (host threads) -> device, pipelined in chunks: the DMA of chunk c runs while the threads copy chunk c + 1
  for (i64 off = 0; off < n; off += IO_CHUNK) {
    const i64 len = std::min<i64>(IO_CHUNK, n - off);
    parallel_for(len, [&](i64 lo, i64 hi) { std::memcpy(pin_a + off + lo, p + off + lo, sizeof(double) * (hi - lo)); }, 1 << 17);
    NGB_CUDA(cudaMemcpyAsync(stage + off, pin_a + off, sizeof(double) * len, cudaMemcpyHostToDevice, st));
  }
  NGB_CUDA(cudaStreamSynchronize(st));
  return stage;
}

void Amg::from_device(double *dst, const double *src_dev, i64 n)
{
  if (is_device_ptr(dst)) {
    if (dst != src_dev) NGB_CUDA(cudaMemcpyAsync(dst, src_dev, sizeof(double) * n, cudaMemcpyDeviceToDevice, st));
    NGB_CUDA(cudaStreamSynchronize(st));
    return;
  }
  // device -> pinned -> pageable, pipelined: the threads copy chunk c out of the pinned buffer while the DMA of the later chunks runs
  const i64 nchunks = (n + IO_CHUNK - 1) / IO_CHUNK;
  while ((i64)io_ev.size() < nchunks) { cudaEvent_t e; NGB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming)); io_ev.push_back(e); }
  for (i64 c = 0; c < nchunks; c++) {
    const i64 off = c * IO_CHUNK, len = std::min<i64>(IO_CHUNK, n - off);
    NGB_CUDA(cudaMemcpyAsync(pin_b + off, src_dev + off, sizeof(double) * len, cudaMemcpyDeviceToHost, st));
    NGB_CUDA(cudaEventRecord(io_ev[c], st));
  }
  for (i64 c = 0; c < nchunks; c++) {
    const i64 off = c * IO_CHUNK, len = std::min<i64>(IO_CHUNK, n - off);
    NGB_CUDA(cudaEventSynchronize(io_ev[c]));
    parallel_for(len, [&](i64 lo, i64 hi) { std::memcpy(dst + off + lo, pin_b + off + lo, sizeof(double) * (hi - lo)); }, 1 << 17);
  }
}

}  // namespace ngb

// =================================================================================================
// C ABI
// =================================================================================================
using namespace ngb;

static thread_local std::string g_err;
struct ngsamg_b200 { Amg amg; };
struct ngsamg_b200_spm { DevCsr d; cudaStream_t st; };

#define NGB_TRY try {
#define NGB_CATCH                                                     \
  }                                                                   \
  catch (const std::exception &e) { g_err = e.what(); return 1; }     \
  catch (...) { g_err = "unknown error"; return 1; }                  \
  return 0;

static void require_device(int device)
{
  int cnt = 0;
  cudaError_t e = cudaGetDeviceCount(&cnt);
  if (e != cudaSuccess || cnt == 0)
    throw Error(std::string("no CUDA device available (") + (e != cudaSuccess ? cudaGetErrorString(e) : "device count 0") +
                "): ngsamg_b200 has no CPU fallback");
  if (device < 0 || device >= cnt) throw Error("invalid CUDA device index " + std::to_string(device));
  NGB_CUDA(cudaSetDevice(device));
}

static void check_csr(const ngsamg_csr *A, const char *what)
{
  if (!A || !A->rowptr || (A->rowptr[A->nrows] > 0 && (!A->col || !A->val))) throw Error(std::string(what) + ": null matrix arrays");
  if (A->bh < 1 || A->bw < 1 || A->bh > 6 || A->bw > 6) throw Error(std::string(what) + ": unsupported block shape");
  if (A->nrows >= (i64)2147483647 - 64 || A->ncols >= (i64)2147483647 - 64) throw Error(std::string(what) + ": more than 2^31 block rows per GPU");
  if (A->nrows < 0 || A->ncols < 0 || A->rowptr[0] != 0) throw Error(std::string(what) + ": malformed matrix (negative size or rowptr[0] != 0)");
  // NGSolve SparseMatrix invariants every consumer relies on (level schedule, L/D/U split, binary searches of the SpGEMM, device
  // gathers): row pointers monotone, column numbers in range and strictly ascending inside a row.  One O(nnz) pass.
  std::atomic<int> bad{0};
  parallel_for(A->nrows, [&](i64 lo, i64 hi) {
    for (i64 i = lo; i < hi && !bad.load(std::memory_order_relaxed); i++) {
      const i64 b0 = A->rowptr[i], b1 = A->rowptr[i + 1];
      if (b1 < b0) { bad = 1; return; }
      i64 prev = -1;
      for (i64 k = b0; k < b1; k++) {
        const i64 c = A->col[k];
        if (c < 0 || c >= A->ncols) { bad = 2; return; }
        if (c <= prev) { bad = 3; return; }
        prev = c;
      }
    }
  }, 1 << 16);
  if (bad == 1) throw Error(std::string(what) + ": row pointers are not monotone");
  if (bad == 2) throw Error(std::string(what) + ": column index out of range");
  if (bad == 3) throw Error(std::string(what) + ": column indices of a row must be strictly ascending (NGSolve SparseMatrix layout)");
}

static void copy_csr(const ngsamg_csr *A, HostBsr &h)
{
  h.nrows = A->nrows; h.ncols = A->ncols; h.bh = A->bh; h.bw = A->bw;
  h.rowptr.assign(A->rowptr, A->rowptr + A->nrows + 1);
  const i64 nnz = h.rowptr.back();
  h.col.resize(nnz);
  h.val.resize(nnz * h.bs());
  parallel_for(nnz, [&](i64 lo, i64 hi) {
    std::memcpy(h.col.data() + lo, A->col + lo, sizeof(i32) * (hi - lo));
    std::memcpy(h.val.data() + lo * h.bs(), A->val + lo * h.bs(), sizeof(double) * (hi - lo) * h.bs());
  }, 1 << 20);
}

extern "C" {

const char *ngsamg_b200_last_error(void) { return g_err.c_str(); }

static void create_impl(const char *type, const ngsamg_csr *A, const uint8_t *free_mask, const double *vertex_xyz,
                        const char *const *flag_keys, const char *const *flag_vals, int nflags, int device, ngsamg_b200_t **out)
{
  if (!out) throw Error("create: out is null");
  *out = nullptr;
  require_device(device);
  check_csr(A, "create");
  if (A->nrows != A->ncols || A->bh != A->bw) throw Error("create: the system matrix must be square with square blocks");
  std::string t = type ? type : "";
  for (const char *pre : {"NgsAMG.", "ngs_amg."})
    if (t.rfind(pre, 0) == 0) t = t.substr(std::strlen(pre));
  static const char *known[] = {"h1_scal", "h1_2d", "h1_3d", "elast_2d", "elast_3d"};
  bool ok = false;
  for (auto k : known) ok |= (t == k);
  if (!ok) throw Error("unknown preconditioner type '" + t + "' (h1_scal, h1_2d, h1_3d, elast_2d, elast_3d)");
  const int dim = t.find("2d") != std::string::npos ? 2 : 3;
  if (t == "h1_scal" && A->bh != 1) throw Error("h1_scal expects a scalar (1x1 block) matrix");
  if ((t == "h1_2d" && A->bh != 2) || (t == "h1_3d" && A->bh != 3)) throw Error(t + ": block size does not match the dimension");
  if (t == "elast_3d" && A->bh != 3 && A->bh != 6) throw Error("elast_3d expects 3x3 (displacement) or 6x6 (displacement+rotation) blocks");
  if (t == "elast_2d" && A->bh != 2 && A->bh != 3) throw Error("elast_2d expects 2x2 or 3x3 blocks");
  if (t.find("elast") != std::string::npos && !vertex_xyz) throw Error(t + ": vertex coordinates are required (rigid body modes)");
  (void)dim;
  auto h = std::make_unique<ngsamg_b200>();
  Amg &a = h->amg;
  a.type = t;
  a.device = device;
  for (int i = 0; i < nflags; i++)
    if (flag_keys[i]) a.flags.set(flag_keys[i], flag_vals && flag_vals[i] ? flag_vals[i] : "1");
  NGB_CUDA(cudaStreamCreateWithFlags(&a.st, cudaStreamNonBlocking));
  NGB_CUDA(cudaEventCreate(&a.ev0));
  NGB_CUDA(cudaEventCreate(&a.ev1));
  cudaDeviceProp prop;
  NGB_CUDA(cudaGetDeviceProperties(&prop, device));
  a.num_sms = prop.multiProcessorCount;
  a.use_graph = a.flags.flag("b200_cuda_graph", true);
  a.tri_sleep_ns = (unsigned)a.flags.num("b200_tri_sleep_ns", 100);
  a.tri_ctas_per_sm = (int)a.flags.num("b200_tri_ctas_per_sm", 0);
  a.tri_prepoll = (int)a.flags.num("b200_tri_prepoll", 1);
  a.tri_gate_all = (int)a.flags.num("b200_tri_gate_all", 1);
  a.tri_small_rows = (i64)a.flags.num("b200_tri_small_rows", 1000000);
  a.tri_rm = (int)a.flags.num("b200_tri_rm", 1);
  // opt-in: measured on 2 GPUs a DIS2CO + CO2CU pair costs 41 / 37 us (levels 0 / 1) over ncclSend/ncclRecv and 54 / 34 us through the
  // peer-memory kernels (profiles/r02_bench_2gpu_halo_ab.txt) -- no gain yet, NCCL stays the default
  a.halo_p2p = a.flags.flag("b200_halo_p2p", false);
  a.tri_rm_rows_per_warp = std::max(1, (int)a.flags.num("b200_tri_rm_rows_per_warp", 8));
  a.tri_rm_gate_rows = (i64)a.flags.num("b200_tri_rm_gate_rows", 4096);
  a.tri_block_warp_rows = (int)a.flags.num("b200_tri_block_warp_rows", 1);
  a.rm_spmv_rows = (i64)a.flags.num("b200_rm_spmv_rows", 150000);
  a.tri_rm_max_rows = (i64)a.flags.num("b200_tri_rm_max_rows", 100000);
  a.tri_gate_gap_levels = a.flags.num("b200_tri_gate_gap", 0.0);
  a.tri_level_launch_depth = (int)a.flags.num("b200_tri_level_launch_depth", 24);
  a.tri_level_pdl = (int)a.flags.num("b200_tri_level_pdl", 1);
  a.tri_level_launch_rows = (i64)a.flags.num("b200_tri_level_launch_rows", 131072);
  a.tri_repoll_ns = (unsigned)a.flags.num("b200_tri_repoll_ns", 0);
  a.tri_regate = (int)a.flags.num("b200_tri_regate", 1);
  a.tri_split = (int)a.flags.num("b200_tri_split", 0);
  a.spmv_small_rows = (i64)a.flags.num("b200_spmv_small_rows", 200000);
  a.tri_pollmode = (int)a.flags.num("b200_tri_pollmode", 0);
  a.itile_minb = (int)a.flags.num("b200_tile_minb", 7);
  auto L = std::make_unique<Level>();
  copy_csr(A, L->hA);
  if (free_mask) {
    bool all = true;
    for (i64 i = 0; i < A->nrows; i++) all &= (free_mask[i] != 0);
    if (!all) L->free_mask.assign(free_mask, free_mask + A->nrows);
  }
  if (vertex_xyz) L->xyz.assign(vertex_xyz, vertex_xyz + 3 * A->nrows);
  a.lev.push_back(std::move(L));
  *out = h.release();
}

int ngsamg_b200_create(const char *type, const ngsamg_csr *A, const uint8_t *free_mask, const double *vertex_xyz,
                       const char *const *flag_keys, const char *const *flag_vals, int nflags, int device, ngsamg_b200_t **out)
{
  NGB_TRY
  create_impl(type, A, free_mask, vertex_xyz, flag_keys, flag_vals, nflags, device, out);
  NGB_CATCH
}

static void halo_to_pardofs(const ngsamg_halo *halo, i64 n, int rank, ParDofs &pd)
{
  pd = ParDofs();
  pd.n = n;
  if (halo) {
    if (halo->npeers < 0 || (halo->npeers > 0 && (!halo->peers || !halo->ex_ptr))) throw Error("halo: null arrays");
    for (int k = 0; k < halo->npeers; k++) {
      pd.peers.push_back(halo->peers[k]);
      pd.ex.emplace_back(halo->ex_dofs + halo->ex_ptr[k], halo->ex_dofs + halo->ex_ptr[k + 1]);
      for (size_t q = 1; q < pd.ex.back().size(); q++)
        if (pd.ex.back()[q] <= pd.ex.back()[q - 1]) throw Error("halo: the shared dof lists must be ascending");
    }
  }
  pd.derive(rank);
}

int ngsamg_b200_create_parallel(const char *type, const ngsamg_csr *A, const uint8_t *free_mask, const double *vertex_xyz,
                                const ngsamg_halo *halo, const ngsamg_comm *comm, const char *const *flag_keys,
                                const char *const *flag_vals, int nflags, int device, ngsamg_b200_t **out)
{
  NGB_TRY
  if (!comm) throw Error("create_parallel: comm is null");
  if (comm->size < 1 || comm->rank < 0 || comm->rank >= comm->size) throw Error("create_parallel: invalid rank / size");
  if (comm->size > 1 && (!comm->exchange || !comm->allreduce_sum)) throw Error("create_parallel: the communicator needs both callbacks");
  create_impl(type, A, free_mask, vertex_xyz, flag_keys, flag_vals, nflags, device, out);
  Amg &a = (*out)->amg;
  try {
    a.par = true;
    a.comm.c = *comm;
    a.nccl = (ncclComm_t)comm->nccl;
    if (a.nccl) nccl_api();
    Level &L = *a.lev[0];
    // the reference keeps the freedofs mask even when every dof is free; the stage split depends on it (gssmoother.cpp:664-678)
    if (free_mask && L.free_mask.empty()) L.free_mask.assign(free_mask, free_mask + A->nrows);
    halo_to_pardofs(halo, A->nrows, comm->rank, L.pd);
  } catch (...) { delete *out; *out = nullptr; throw; }
  NGB_CATCH
}

int ngsamg_b200_nccl_unique_id(char id[128])
{
  NGB_TRY
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is expected to be 128 bytes");
  ncclUniqueId u;
  NGB_NCCL(nccl_api().GetUniqueId(&u));
  std::memcpy(id, &u, 128);
  NGB_CATCH
}

int ngsamg_b200_nccl_comm_init(const char id[128], int rank, int size, int device, void **nccl_comm)
{
  NGB_TRY
  if (!nccl_comm) throw Error("null output");
  require_device(device);
  ncclUniqueId u;
  std::memcpy(&u, id, 128);
  ncclComm_t c = nullptr;
  NGB_NCCL(nccl_api().CommInitRank(&c, size, u, rank));
  *nccl_comm = (void *)c;
  NGB_CATCH
}

int ngsamg_b200_nccl_comm_destroy(void *nccl_comm)
{
  NGB_TRY
  if (nccl_comm) NGB_NCCL(nccl_api().CommDestroy((ncclComm_t)nccl_comm));
  NGB_CATCH
}

int ngsamg_b200_get_halo(ngsamg_b200_t *h, int level, int32_t *npeers, int32_t *peers, int64_t *ex_ptr, int32_t *ex_dofs)
{
  NGB_TRY
  if (!h) throw Error("null handle");
  Amg &a = h->amg;
  if (!a.par) throw Error("get_halo: not a multi-rank hierarchy");
  if (level < 0 || level >= (int)a.lev.size()) throw Error("level out of range");
  const ParDofs &pd = a.lev[level]->pd;
  if (npeers) *npeers = (i32)pd.peers.size();
  i64 off = 0;
  for (size_t k = 0; k < pd.peers.size(); k++) {
    if (peers) peers[k] = pd.peers[k];
    if (ex_ptr) ex_ptr[k] = off;
    if (ex_dofs) std::memcpy(ex_dofs + off, pd.ex[k].data(), sizeof(i32) * pd.ex[k].size());
    off += (i64)pd.ex[k].size();
  }
  if (ex_ptr) ex_ptr[pd.peers.size()] = off;
  NGB_CATCH
}

int ngsamg_b200_get_hybrid(ngsamg_b200_t *h, int level, int which, int64_t *nnz, int64_t *rowptr, int32_t *col, double *val,
                           double *mod_diag)
{
  NGB_TRY
  if (!h) throw Error("null handle");
  Amg &a = h->amg;
  if (!a.par || level < 0 || level >= a.npar) throw Error("get_hybrid: not a distributed level");
  Level &L = *a.lev[level];
  if (which == 2) { if (nnz) { nnz[0] = L.nnz_m; nnz[1] = L.nnz_g; } }
  else {
    if (!L.keep_host) throw Error("hybrid matrices were not kept on the host (raise ngs_amg_keep_host_nnz)");
    const HostBsr &M = which == 0 ? L.hM : L.hG;
    if (nnz) *nnz = M.nnz();
    if (rowptr) std::memcpy(rowptr, M.rowptr.data(), sizeof(i64) * (M.nrows + 1));
    if (col) std::memcpy(col, M.col.data(), sizeof(i32) * M.nnz());
    if (val) std::memcpy(val, M.val.data(), sizeof(double) * M.nnz() * M.bs());
  }
  if (mod_diag) std::memcpy(mod_diag, L.mod_diag.data(), sizeof(double) * L.mod_diag.size());
  NGB_CATCH
}

int ngsamg_b200_num_parallel_levels(ngsamg_b200_t *h) { return (h && h->amg.finalized && h->amg.par) ? h->amg.npar : 0; }

// transport of the per-sweep halo exchange of `level`: 0 host-staged callbacks, 1 NCCL send/recv, 2 NVLink peer memory; -1 = not a distributed level
int ngsamg_b200_halo_transport(ngsamg_b200_t *h, int level)
{
  if (!h || !h->amg.finalized || !h->amg.par || level < 0 || level >= h->amg.npar) return -1;
  const Level &L = *h->amg.lev[level];
  return L.p2p ? 2 : (h->amg.nccl ? 1 : 0);
}

ngsamg_b200_t *ngsamg_b200_get_contracted(ngsamg_b200_t *h)
{
  // the nested hierarchy lives inside the parent; hand out a view that shares its storage
  if (!h || !h->amg.par || !h->amg.nested) return nullptr;
  return reinterpret_cast<ngsamg_b200_t *>(h->amg.nested.get());
}

int ngsamg_b200_get_contraction_map(ngsamg_b200_t *h, int rank, int64_t *n, int32_t *map)
{
  NGB_TRY
  if (!h) throw Error("null handle");
  Amg &a = h->amg;
  if (!a.par || a.comm.rank() != 0) throw Error("get_contraction_map: only on rank 0 of a multi-rank hierarchy");
  if (rank < 0 || rank >= (int)a.ctr.dof_maps.size()) throw Error("rank out of range");
  if (n) *n = (i64)a.ctr.dof_maps[rank].size();
  if (map) std::memcpy(map, a.ctr.dof_maps[rank].data(), sizeof(i32) * a.ctr.dof_maps[rank].size());
  NGB_CATCH
}

// host-only: one step of the class-respecting (multi-rank) coarsening, exactly what finalize() runs per distributed level
struct ngsamg_b200_parcoarsen { HostBsr P; std::vector<i32> vmap; std::vector<double> cxyz; ParDofs cpd; };

int ngsamg_b200_coarsen_parallel_begin(const ngsamg_csr *A, const uint8_t *free_mask, const double *vertex_xyz, const ngsamg_halo *halo,
                                       const ngsamg_comm *comm, int bcoarse, int max_per_row, double min_frac, double omega, int smooth,
                                       int rounds, ngsamg_b200_parcoarsen **out, int64_t *ncoarse, int64_t *nnz, int32_t *npeers_coarse,
                                       int64_t *nshared_coarse)
{
  NGB_TRY
  if (!comm || !out) throw Error("null argument");
  check_csr(A, "coarsen_parallel");
  HostBsr hA, Acum;
  copy_csr(A, hA);
  ParDofs pd;
  halo_to_pardofs(halo, A->nrows, comm->rank, pd);
  Comm c;
  c.c = *comm;
  cumulate_matrix(c, pd, hA, Acum);
  std::vector<double> rowsum, xyz;
  assembled_row_sums(c, pd, hA, rowsum);
  if (vertex_xyz) xyz.assign(vertex_xyz, vertex_xyz + 3 * A->nrows);
  CoarsenOptions o;
  o.max_per_row = max_per_row; o.min_frac = min_frac; o.omega = omega; o.smooth = smooth != 0; o.rounds = rounds;
  ParCoarsen pc;
  pc.pd = &pd; pc.rowsum = &rowsum; pc.rank = comm->rank;
  auto r = std::make_unique<ngsamg_b200_parcoarsen>();
  build_prolongation(Acum, free_mask, bcoarse, xyz, o, r->P, r->vmap, r->cxyz, &pc);
  coarse_pardofs(pd, r->vmap, r->P.ncols, r->cpd, comm->rank);
  if (ncoarse) *ncoarse = r->P.ncols;
  if (nnz) *nnz = r->P.nnz();
  if (npeers_coarse) *npeers_coarse = (i32)r->cpd.peers.size();
  if (nshared_coarse) { i64 t = 0; for (auto &l : r->cpd.ex) t += (i64)l.size(); *nshared_coarse = t; }
  *out = r.release();
  NGB_CATCH
}

int ngsamg_b200_coarsen_parallel_fetch(ngsamg_b200_parcoarsen *m, int64_t *rowptr, int32_t *col, double *val, int32_t *vmap, double *cxyz,
                                       int32_t *peers, int64_t *ex_ptr, int32_t *ex_dofs)
{
  NGB_TRY
  if (!m) throw Error("null handle");
  if (rowptr) std::memcpy(rowptr, m->P.rowptr.data(), sizeof(i64) * (m->P.nrows + 1));
  if (col) std::memcpy(col, m->P.col.data(), sizeof(i32) * m->P.nnz());
  if (val) std::memcpy(val, m->P.val.data(), sizeof(double) * m->P.nnz() * m->P.bs());
  if (vmap) std::memcpy(vmap, m->vmap.data(), sizeof(i32) * m->vmap.size());
  if (cxyz && !m->cxyz.empty()) std::memcpy(cxyz, m->cxyz.data(), sizeof(double) * m->cxyz.size());
  i64 off = 0;
  for (size_t k = 0; k < m->cpd.peers.size(); k++) {
    if (peers) peers[k] = m->cpd.peers[k];
    if (ex_ptr) ex_ptr[k] = off;
    if (ex_dofs) std::memcpy(ex_dofs + off, m->cpd.ex[k].data(), sizeof(i32) * m->cpd.ex[k].size());
    off += (i64)m->cpd.ex[k].size();
  }
  if (ex_ptr) ex_ptr[m->cpd.peers.size()] = off;
  delete m;
  NGB_CATCH
}

struct ngsamg_b200_hybrid_host { HostBsr M, G; std::vector<double> md; std::vector<i32> sweep; std::vector<uint8_t> master; };

int ngsamg_b200_hybrid_host_begin(const ngsamg_csr *A, const uint8_t *free_mask, const ngsamg_halo *halo, const ngsamg_comm *comm,
                                  ngsamg_b200_hybrid_host **out, int64_t *nnz_m, int64_t *nnz_g)
{
  NGB_TRY
  if (!comm || !out) throw Error("null argument");
  check_csr(A, "hybrid_host");
  HostBsr hA, Acum;
  copy_csr(A, hA);
  ParDofs pd;
  halo_to_pardofs(halo, A->nrows, comm->rank, pd);
  Comm c;
  c.c = *comm;
  auto r = std::make_unique<ngsamg_b200_hybrid_host>();
  cumulate_matrix(c, pd, hA, Acum);
  hybrid_split(pd, hA, Acum, r->M, r->G);
  hybrid_mod_diag(c, pd, Acum, r->G, free_mask, r->md);
  std::vector<uint8_t> sm;
  hybrid_sweep_order(pd, free_mask, r->sweep, sm, nullptr);
  r->master.resize(A->nrows);
  for (i64 d = 0; d < A->nrows; d++) r->master[d] = pd.master_of[d] < 0 ? 1 : 0;
  if (nnz_m) *nnz_m = r->M.nnz();
  if (nnz_g) *nnz_g = r->G.nnz();
  *out = r.release();
  NGB_CATCH
}

int ngsamg_b200_hybrid_host_fetch(ngsamg_b200_hybrid_host *m, int64_t *m_rowptr, int32_t *m_col, double *m_val, int64_t *g_rowptr,
                                  int32_t *g_col, double *g_val, double *mod_diag, int32_t *sweep_rank, uint8_t *master)
{
  NGB_TRY
  if (!m) throw Error("null handle");
  auto cp = [](const HostBsr &H, int64_t *rp, int32_t *ci, double *v) {
    if (rp) std::memcpy(rp, H.rowptr.data(), sizeof(i64) * (H.nrows + 1));
    if (ci) std::memcpy(ci, H.col.data(), sizeof(i32) * H.nnz());
    if (v) std::memcpy(v, H.val.data(), sizeof(double) * H.nnz() * H.bs());
  };
  cp(m->M, m_rowptr, m_col, m_val);
  cp(m->G, g_rowptr, g_col, g_val);
  if (mod_diag) std::memcpy(mod_diag, m->md.data(), sizeof(double) * m->md.size());
  if (sweep_rank) std::memcpy(sweep_rank, m->sweep.data(), sizeof(i32) * m->sweep.size());
  if (master) std::memcpy(master, m->master.data(), m->master.size());
  delete m;
  NGB_CATCH
}

// host-only: contraction of one distributed level onto rank 0 (par.cpp: contract_to_root = CtrMap::DoAssembleMatrix + the dof maps).  Collective;
// only rank 0 gets data: the merged matrix and, per rank, the map local dof -> merged dof.
struct ngsamg_b200_contract_host { Contraction c; };

int ngsamg_b200_contract_host_begin(const ngsamg_csr *A, const uint8_t *free_mask, const ngsamg_halo *halo, const ngsamg_comm *comm,
                                    ngsamg_b200_contract_host **out, int64_t *n_merged, int64_t *nnz_merged, int64_t *map_total)
{
  NGB_TRY
  if (!comm || !out) throw Error("null argument");
  check_csr(A, "contract_host");
  HostBsr hA;
  copy_csr(A, hA);
  ParDofs pd;
  halo_to_pardofs(halo, A->nrows, comm->rank, pd);
  Comm c;
  c.c = *comm;
  std::vector<uint8_t> fm;
  if (free_mask) fm.assign(free_mask, free_mask + A->nrows);
  auto r = std::make_unique<ngsamg_b200_contract_host>();
  contract_to_root(c, pd, hA, fm, std::vector<double>(), r->c);
  if (n_merged) *n_merged = r->c.A.nrows;
  if (nnz_merged) *nnz_merged = r->c.A.nnz();
  if (map_total) { *map_total = 0; for (auto &m : r->c.dof_maps) *map_total += (i64)m.size(); }
  *out = r.release();
  NGB_CATCH
}

// rank 0: rowptr[n_merged + 1], col, val of the merged matrix; dof_map: concatenation over the ranks of local -> merged (map_ptr[R + 1] offsets).
// Other ranks: nothing is written.  Frees the handle.
int ngsamg_b200_contract_host_fetch(ngsamg_b200_contract_host *m, int64_t *rowptr, int32_t *col, double *val, int64_t *map_ptr, int32_t *dof_map)
{
  NGB_TRY
  if (!m) throw Error("null handle");
  const HostBsr &H = m->c.A;
  if (H.nrows > 0 || !m->c.dof_maps.empty()) {
    if (rowptr) std::memcpy(rowptr, H.rowptr.data(), sizeof(i64) * (H.nrows + 1));
    if (col) std::memcpy(col, H.col.data(), sizeof(i32) * H.nnz());
    if (val) std::memcpy(val, H.val.data(), sizeof(double) * H.nnz() * H.bs());
    i64 off = 0;
    for (size_t r = 0; r < m->c.dof_maps.size(); r++) {
      if (map_ptr) map_ptr[r] = off;
      if (dof_map) std::memcpy(dof_map + off, m->c.dof_maps[r].data(), sizeof(i32) * m->c.dof_maps[r].size());
      off += (i64)m->c.dof_maps[r].size();
    }
    if (map_ptr) map_ptr[m->c.dof_maps.size()] = off;
  }
  delete m;
  NGB_CATCH
}

int ngsamg_b200_set_prolongations(ngsamg_b200_t *h, int nprol, const ngsamg_csr *P)
{
  NGB_TRY
  if (!h) throw Error("null handle");
  if (h->amg.finalized) throw Error("set_prolongations must be called before finalize");
  h->amg.injected.clear();
  for (int l = 0; l < nprol; l++) {
    check_csr(&P[l], "set_prolongations");
    h->amg.injected.emplace_back();
    copy_csr(&P[l], h->amg.injected.back());
  }
  if (nprol == 0) { h->amg.flags.set("max_levels", "1"); }
  NGB_CATCH
}

int ngsamg_b200_finalize(ngsamg_b200_t *h)
{
  NGB_TRY
  if (!h) throw Error("null handle");
  NGB_CUDA(cudaSetDevice(h->amg.device));
  h->amg.finalize();
  NGB_CATCH
}

void ngsamg_b200_destroy(ngsamg_b200_t *h) { delete h; }

static Amg &ready(ngsamg_b200_t *h)
{
  if (!h) throw Error("null handle");
  if (!h->amg.finalized) throw Error("hierarchy not finalized");
  NGB_CUDA(cudaSetDevice(h->amg.device));
  return h->amg;
}
static Level &get_level(Amg &a, int level)
{
  if (level < 0 || level >= (int)a.lev.size()) throw Error("level out of range");
  return *a.lev[level];
}

// blocks of the block Gauss-Seidel smoother of `level` (GetGSBlocks, amg_pc_vertex_impl.hpp:1171-1269): block_of[v] = block (coarse vertex) of
// vertex v, -1 = in no block; returns 1 with an error if the level is not smoothed by bgs
int ngsamg_b200_get_gs_blocks(ngsamg_b200_t *h, int level, int32_t *block_of)
{
  NGB_TRY
  Amg &a = ready(h);
  Level &L = get_level(a, level);
  if (L.sm_type != SM_BGS) throw Error("get_gs_blocks: level " + std::to_string(level) + " is not smoothed by block Gauss-Seidel");
  if (block_of)
    for (i64 v = 0; v < L.n; v++) block_of[v] = (L.free_mask.empty() || L.free_mask[v]) ? L.gs_block[v] : -1;
  NGB_CATCH
}


static void apply_impl(Amg &a, double s, const double *b, double *x, bool add)
{
  Level &L = *a.lev[0];
  const i64 n = L.n * L.b;
  a.ensure_io(n);
  cudaEventRecord(a.ev0, a.st);
  const double *bd = a.to_device(b, n, a.io_a);
  k_permute_in<<<nblk(L.n), TB, 0, a.st>>>(L.n, L.b, L.d_perm, bd, L.rhs);
  a.vcycle();
  const bool xdev = a.is_device_ptr(x);
  double *xd = xdev ? x : a.io_b;
  if (add && !xdev) { a.to_device(x, n, a.io_b); }
  k_permute_out<<<nblk(L.n), TB, 0, a.st>>>(L.n, L.b, L.d_perm, L.result, xd, s, add ? 1 : 0);
  a.launches += 2;
  if (!xdev) a.from_device(x, xd, n);
  cudaEventRecord(a.ev1, a.st);
  NGB_CUDA(cudaEventSynchronize(a.ev1));
  float ms = 0;
  cudaEventElapsedTime(&ms, a.ev0, a.ev1);
  a.ms_apply = ms;
  NGB_CUDA(cudaGetLastError());
  a.check_watchdog();
}

int ngsamg_b200_apply(ngsamg_b200_t *h, const double *b, double *x)
{
  NGB_TRY
  apply_impl(ready(h), 1.0, b, x, false);
  NGB_CATCH
}

int ngsamg_b200_apply_add(ngsamg_b200_t *h, double s, const double *b, double *x)
{
  NGB_TRY
  apply_impl(ready(h), s, b, x, true);
  NGB_CATCH
}

// one V-cycle executed EAGERLY (no CUDA graph) with an event pair around every phase: where the cycle spends its device time.
// ms[7] = { triangular sweeps, parallel halves of the sweeps (U / L+D passes), restriction + prolongation, halo exchanges (DIS2CO / CO2CU:
// pack + NCCL or host staging + unpack), G products of the hybrid smoother, coarse part (contracted hierarchy incl. gather/scatter, or the
// coarsest solve), other (vector copies) }.  Collective on a distributed hierarchy.  The sum is larger than a graph-launched cycle
// (launch gaps are exposed); the SHARES are what it is for.
int ngsamg_b200_apply_phases(ngsamg_b200_t *h, const double *b, double *x, double *ms)
{
  NGB_TRY
  Amg &a = ready(h);
  if (!ms) throw Error("apply_phases: ms is null");
  PhaseTimer pt;
  pt.st = a.st;
  const bool ug = a.use_graph;
  a.use_graph = false;
  a.pt = &pt;
  struct Restore { Amg &a; bool ug; ~Restore() { a.pt = nullptr; a.use_graph = ug; } } restore{a, ug};
  apply_impl(a, 1.0, b, x, false);
  NGB_CUDA(cudaStreamSynchronize(a.st));
  pt.collect(ms);
  NGB_CATCH
}

static void ensure_scratch(Amg &a, Level &L)
{
  if (L.wa) return;
  const size_t nb = (size_t)L.npad * L.b;
  L.wa = dev_alloc<double>(nb);
  L.wb = dev_alloc<double>(nb);
  NGB_CUDA(cudaMemsetAsync(L.wa, 0, nb * sizeof(double), a.st));
  NGB_CUDA(cudaMemsetAsync(L.wb, 0, nb * sizeof(double), a.st));
}

int ngsamg_b200_spmv_add(ngsamg_b200_t *h, int level, double s, const double *x, double *y)
{
  NGB_TRY
  Amg &a = ready(h);
  Level &L = get_level(a, level);
  if (level == (int)a.lev.size() - 1 && a.lev.size() > 1) throw Error("spmv_add: the coarsest level keeps no sparse matrix on the device");
  if (!L.L.slice_ptr) throw Error("spmv_add: level has no device matrix");
  const i64 n = L.n * L.b;
  a.ensure_io(n);
  ensure_scratch(a, L);
  const double *xd = a.to_device(x, n, a.io_a);
  k_permute_in<<<nblk(L.n), TB, 0, a.st>>>(L.n, L.b, L.d_perm, xd, L.wa);
  a.spmv_part(L, 4, L.wa, nullptr, L.wb, 1.0, 0.0, nullptr);
  if (a.par && L.par && L.G.nnz) a.transfer(L.G, L.wa, L.wb, L.wb, 1.0, 1.0);   // HybridBaseMatrix::MultAdd: (M + G) x, x CUMULATED -> y DISTRIBUTED
  const bool ydev = a.is_device_ptr(y);
  double *yd = ydev ? y : a.io_b;
  if (!ydev) a.to_device(y, n, a.io_b);
  k_permute_out<<<nblk(L.n), TB, 0, a.st>>>(L.n, L.b, L.d_perm, L.wb, yd, s, 1);
  a.launches += 2;
  a.from_device(y, yd, n);
  NGB_CUDA(cudaGetLastError());
  NGB_CATCH
}

int ngsamg_b200_smooth(ngsamg_b200_t *h, int level, double *x, const double *b, double *res, int res_updated, int update_res,
                       int x_zero, int backwards)
{
  NGB_TRY
  Amg &a = ready(h);
  Level &L = get_level(a, level);
  if (!L.L.slice_ptr) throw Error("smooth: level has no smoother (coarsest level)");
  const i64 n = L.n * L.b;
  a.ensure_io(n);
  const double *xd = a.to_device(x, n, a.io_a);
  k_permute_in<<<nblk(L.n), TB, 0, a.st>>>(L.n, L.b, L.d_perm, xd, L.x);
  const double *bd = a.to_device(b, n, a.io_a);
  k_permute_in<<<nblk(L.n), TB, 0, a.st>>>(L.n, L.b, L.d_perm, bd, L.rhs);
  if (res) {
    const double *rd = a.to_device(res, n, a.io_a);
    k_permute_in<<<nblk(L.n), TB, 0, a.st>>>(L.n, L.b, L.d_perm, rd, L.res);
  }
  // multi-rank level: HybridBaseSmoother::Smooth / SmoothBack -- collective; x CUMULATED, b and res DISTRIBUTED
  if (a.par && L.par) a.hybrid_level_smooth(L, L.x, L.rhs, L.res, res_updated != 0, update_res != 0, x_zero != 0, backwards != 0);
  else a.level_smooth(L, L.x, L.rhs, L.res, res_updated != 0, update_res != 0, x_zero != 0, backwards != 0);
  k_permute_out<<<nblk(L.n), TB, 0, a.st>>>(L.n, L.b, L.d_perm, L.x, a.io_b, 1.0, 0);
  a.from_device(x, a.io_b, n);
  if (res) {
    k_permute_out<<<nblk(L.n), TB, 0, a.st>>>(L.n, L.b, L.d_perm, L.res, a.io_b, 1.0, 0);
    a.from_device(res, a.io_b, n);
  }
  a.launches += 5;
  NGB_CUDA(cudaGetLastError());
  a.check_watchdog();
  NGB_CATCH
}

int ngsamg_b200_restrict(ngsamg_b200_t *h, int level, const double *xf, double *xc)
{
  NGB_TRY
  Amg &a = ready(h);
  Level &F = get_level(a, level);
  Level &C = get_level(a, level + 1);
  const i64 nf = F.n * F.b, nc = C.n * C.b;
  a.ensure_io(std::max(nf, nc));
  const double *xd = a.to_device(xf, nf, a.io_a);
  k_permute_in<<<nblk(F.n), TB, 0, a.st>>>(F.n, F.b, F.d_perm, xd, F.res);
  a.transfer(F.PT, F.res, nullptr, C.rhs, 1.0, 0.0, F.d_pt_rowmap);
  k_permute_out<<<nblk(C.n), TB, 0, a.st>>>(C.n, C.b, C.d_perm, C.rhs, a.io_b, 1.0, 0);
  a.from_device(xc, a.io_b, nc);
  a.launches += 2;
  NGB_CUDA(cudaGetLastError());
  NGB_CATCH
}

// exact solve on the coarsest level (crs_inv->Mult, amg_matrix.cpp:228-233); x = 0 when the hierarchy was built with clev=none
int ngsamg_b200_coarse_solve(ngsamg_b200_t *h, const double *rhs, double *x)
{
  NGB_TRY
  Amg &a = ready(h);
  Level &C = get_level(a, (int)a.lev.size() - 1);
  const i64 nc = C.n * C.b;
  a.ensure_io(nc);
  const double *rd = a.to_device(rhs, nc, a.io_a);
  k_permute_in<<<nblk(C.n), TB, 0, a.st>>>(C.n, C.b, C.d_perm, rd, C.rhs);
  if (a.has_cinv) k_dense_gemv<<<nblk((i64)a.cinv_n * 32), TB, 0, a.st>>>(a.cinv_n, a.d_cinv, C.rhs, C.x);
  else NGB_CUDA(cudaMemsetAsync(C.x, 0, sizeof(double) * C.npad * C.b, a.st));
  k_permute_out<<<nblk(C.n), TB, 0, a.st>>>(C.n, C.b, C.d_perm, C.x, a.io_b, 1.0, 0);
  a.from_device(x, a.io_b, nc);
  a.launches += 3;
  NGB_CUDA(cudaGetLastError());
  NGB_CATCH
}

int ngsamg_b200_prolong_add(ngsamg_b200_t *h, int level, double fac, const double *xc, double *xf)
{
  NGB_TRY
  Amg &a = ready(h);
  Level &F = get_level(a, level);
  Level &C = get_level(a, level + 1);
  const i64 nf = F.n * F.b, nc = C.n * C.b;
  a.ensure_io(std::max(nf, nc));
  const double *cd = a.to_device(xc, nc, a.io_a);
  k_permute_in<<<nblk(C.n), TB, 0, a.st>>>(C.n, C.b, C.d_perm, cd, C.x);
  const double *fd = a.to_device(xf, nf, a.io_a);
  k_permute_in<<<nblk(F.n), TB, 0, a.st>>>(F.n, F.b, F.d_perm, fd, F.x);
  a.transfer(F.P, C.x, F.x, F.x, fac, 1.0);
  k_permute_out<<<nblk(F.n), TB, 0, a.st>>>(F.n, F.b, F.d_perm, F.x, a.io_b, 1.0, 0);
  a.from_device(xf, a.io_b, nf);
  a.launches += 3;
  NGB_CUDA(cudaGetLastError());
  NGB_CATCH
}

// ngsolve.krylovspace.CGSolver restated (see oracle/ngsamg_oracle.c:orc_amg_pcg for the recurrence), all vectors stay in
// the level-scheduled numbering on the device; per iteration: 1 SpMV, 1 V-cycle (CUDA graph), 2 dots, 2 fused updates.
int ngsamg_b200_pcg(ngsamg_b200_t *h, const double *rhs, double *x, double tol, int maxsteps, int *iters, double *errors)
{
  NGB_TRY
  Amg &a = ready(h);
  Level &L = *a.lev[0];
  if (a.lev.size() < 2 && !(a.par && a.npar == 0)) throw Error("pcg needs at least two levels");
  if (a.par && a.npar == 0) throw Error("pcg: the hierarchy was contracted at level 0 (single rank?) -- use the single-rank constructor");
  const i64 n = L.n * L.b, np = L.npad * L.b;
  a.ensure_io(n);
  if (!a.cg_u) { a.cg_u = dev_alloc<double>(np); a.cg_s = dev_alloc<double>(np); a.cg_q = dev_alloc<double>(np); }
  cudaEventRecord(a.ev0, a.st);
  const double *rd = a.to_device(rhs, n, a.io_a);
  double *d = L.rhs, *u = a.cg_u, *s = a.cg_s, *q = a.cg_q;
  NGB_CUDA(cudaMemsetAsync(u, 0, sizeof(double) * np, a.st));
  k_permute_in<<<nblk(L.n), TB, 0, a.st>>>(L.n, L.b, L.d_perm, rd, d);
  a.vcycle();
  const double *w = L.result;   // fixed by the (captured) cycle: x or y buffer of level 0
  NGB_CUDA(cudaMemcpyAsync(s, w, sizeof(double) * np, cudaMemcpyDeviceToDevice, a.st));
  double wdn = a.dot(np, w, d);
  const double err0 = std::sqrt(std::fabs(wdn));
  if (errors) errors[0] = err0;
  int it = 0;
  if (wdn != 0.0)
    for (it = 1; it <= maxsteps; it++) {
      a.spmv_part(L, 4, s, nullptr, q, 1.0, 0.0, nullptr);
      if (a.par && L.par && L.G.nnz) a.transfer(L.G, s, q, q, 1.0, 1.0);   // HybridBaseMatrix::Mult: (M + G) x, x CUMULATED -> DISTRIBUTED
      const double wd = wdn;
      const double as_s = a.dot(np, s, q);
      const double alpha = wd / as_s;
      k_cg_update<<<nblk(np), TB, 0, a.st>>>(np, alpha, s, q, u, d);
      a.vcycle();
      wdn = a.dot(np, w, d);
      const double beta = wdn / wd;
      k_axpby<<<nblk(np), TB, 0, a.st>>>(np, 1.0, w, beta, s);
      a.launches += 2;
      const double err = std::sqrt(std::fabs(wd));
      if (errors) errors[it] = err;
      if (err < tol * err0) break;
    }
  if (it > maxsteps) it = maxsteps;
  if (iters) *iters = it;
  const bool xdev = a.is_device_ptr(x);
  double *xd = xdev ? x : a.io_b;
  k_permute_out<<<nblk(L.n), TB, 0, a.st>>>(L.n, L.b, L.d_perm, u, xd, 1.0, 0);
  a.launches += 2;
  if (!xdev) a.from_device(x, xd, n);
  cudaEventRecord(a.ev1, a.st);
  NGB_CUDA(cudaEventSynchronize(a.ev1));
  float ms = 0;
  cudaEventElapsedTime(&ms, a.ev0, a.ev1);
  a.ms_pcg = ms;
  NGB_CUDA(cudaGetLastError());
  a.check_watchdog();
  NGB_CATCH
}

int ngsamg_b200_num_levels(ngsamg_b200_t *h) { return (h && h->amg.finalized) ? (int)h->amg.lev.size() : 0; }

// which kernel sweeps the level: 0 row-level (k_gs_tri / k_gs_level / k_gs_tri_small), 1 warp per tile, 2 CTA per tile, 3 CTA per tile on
// prepared tile images, 4 warp per row on the row-major copy of a small level (k_gs_tri_rm); -1 = no such level
int ngsamg_b200_level_sweep_kind(ngsamg_b200_t *h, int level)
{
  if (!h || !h->amg.finalized || level < 0 || level >= (int)h->amg.lev.size()) return -1;
  const Level &L = *h->amg.lev[level];
  if (!L.tiled) return (L.rmL.ptr && h->amg.tri_rm && !h->amg.level_launch(L)) ? 4 : 0;
  if (L.tile_maxs <= 2) return 1;
  return L.itile ? 3 : 2;
}

static void level_bytes(const Level &L, i64 &m, i64 &p, i64 &v)
{
  m = (L.par ? L.nnz_m + L.nnz_g : L.nnz) * (8 * (i64)L.b * L.b + 4) + 4 * (L.n + 1);   // hybrid level: M and G are each read once per sweep
  p = L.hP.nnz() ? L.hP.nnz() * (8 * (i64)L.b * L.bc + 4) + 4 * (L.n + 1) : 0;
  v = 8 * L.n * L.b;
}

int ngsamg_b200_level_info(ngsamg_b200_t *h, int level, ngsamg_level_info *info)
{
  NGB_TRY
  Amg &a = ready(h);
  Level &L = get_level(a, level);
  if (!info) throw Error("null info");
  info->n = L.n; info->b = L.b; info->nnz = L.nnz;
  info->nnz_prol = L.P.nnz; info->ncoarse = L.nc; info->bcoarse = L.bc;
  info->gs_depth = L.depth;
  level_bytes(L, info->bytes_matrix, info->bytes_prol, info->bytes_vec);
  if (!L.P.nnz) info->bytes_prol = 0;
  NGB_CATCH
}

int ngsamg_b200_get_level_matrix(ngsamg_b200_t *h, int level, int64_t *rowptr, int32_t *col, double *val)
{
  NGB_TRY
  Amg &a = ready(h);
  Level &L = get_level(a, level);
  if (!L.keep_host) throw Error("level matrix was not kept on the host (raise ngs_amg_keep_host_nnz)");
  if (rowptr) std::memcpy(rowptr, L.hA.rowptr.data(), sizeof(i64) * (L.n + 1));
  if (col) std::memcpy(col, L.hA.col.data(), sizeof(i32) * L.hA.nnz());
  if (val) std::memcpy(val, L.hA.val.data(), sizeof(double) * L.hA.nnz() * L.hA.bs());
  NGB_CATCH
}

int ngsamg_b200_get_prolongation(ngsamg_b200_t *h, int level, int64_t *rowptr, int32_t *col, double *val)
{
  NGB_TRY
  Amg &a = ready(h);
  Level &L = get_level(a, level);
  if (level + 1 >= (int)a.lev.size()) throw Error("the coarsest level has no prolongation");
  if (rowptr) std::memcpy(rowptr, L.hP.rowptr.data(), sizeof(i64) * (L.n + 1));
  if (col) std::memcpy(col, L.hP.col.data(), sizeof(i32) * L.hP.nnz());
  if (val) std::memcpy(val, L.hP.val.data(), sizeof(double) * L.hP.nnz() * L.hP.bs());
  NGB_CATCH
}

int ngsamg_b200_get_level_vector(ngsamg_b200_t *h, int level, int which, double *out)
{
  NGB_TRY
  Amg &a = ready(h);
  Level &L = get_level(a, level);
  const double *src = which == 0 ? (L.result ? L.result : L.x) : which == 1 ? L.rhs : which == 2 ? L.res : nullptr;
  if (!src) throw Error("get_level_vector: which must be 0 (x), 1 (rhs) or 2 (res)");
  const i64 n = L.n * L.b;
  a.ensure_io(n);
  k_permute_out<<<nblk(L.n), TB, 0, a.st>>>(L.n, L.b, L.d_perm, src, a.io_b, 1.0, 0);
  a.from_device(out, a.io_b, n);
  NGB_CATCH
}

int ngsamg_b200_get_sweep_order(ngsamg_b200_t *h, int level, int32_t *rank)
{
  NGB_TRY
  Amg &a = ready(h);
  Level &L = get_level(a, level);
  if (!rank) throw Error("null output");
  for (i64 i = 0; i < L.n; i++) rank[i] = L.sweep_rank.empty() ? (i32)i : L.sweep_rank[i];
  NGB_CATCH
}

// AMGMatrix::GetOC (amg_matrix.cpp:551-582): occs[1 + l] = cycle factor * nops_l / nze_0 for the levels that carry a smoother, 0 for the
// coarsest (exactly solved) level; occs[0] = their sum.  nze = scalar non-zeros (GetScalNZE: NZE * entry size, utils_sparseLA.cpp:324-342),
// nops_l = nze_l, times nsteps * (symm ? 2 : 1) under a ProxySmoother (base_smoother.cpp:38-41); cycle factor 1 (V), 2^l (W), 2 (1 + l) (BS).
int ngsamg_b200_operator_complexities(ngsamg_b200_t *h, double *occs, int cap)
{
  if (!h || !h->amg.finalized) return 0;
  auto &lev = h->amg.lev;
  const int nl = (int)lev.size();
  const int nsm = std::max(0, nl - 1);
  const int nout = 1 + nsm + (h->amg.has_cinv ? 1 : 0);
  if (!occs || cap < nout) return nout;
  const double nze0 = (double)lev[0]->nnz * lev[0]->b * lev[0]->b;
  double sum = 0.0;
  for (int l = 0; l < nout - 1; l++) {
    double v = 0.0;
    if (l < nsm) {
      const Level &L = *lev[l];
      double nops = (double)L.nnz * L.b * L.b;
      if (L.sm_steps > 1 || L.sm_symm) nops *= (double)L.sm_steps * (L.sm_symm ? 2 : 1);
      const double fac = h->amg.cycle == Amg::CYCLE_W ? std::pow(2.0, l) : h->amg.cycle == Amg::CYCLE_BS ? 2.0 * (1 + l) : 1.0;
      v = fac * nops / nze0;
    }
    occs[1 + l] = v;
    sum += v;
  }
  occs[0] = sum;
  return nout;
}

double ngsamg_b200_operator_complexity(ngsamg_b200_t *h)
{
  double occs[64];
  const int n = ngsamg_b200_operator_complexities(h, occs, 64);
  return (n > 0 && n <= 64) ? occs[0] : 0.0;
}

double ngsamg_b200_vcycle_bytes(ngsamg_b200_t *h)
{
  if (!h || !h->amg.finalized) return 0.0;
  // B_V = sum_{l<L} [ 2(M_l + D_l) + 2 P_l + 9 v_l + 2 v_{l+1} ] + B_coarse        
  auto &lev = h->amg.lev;
  double B = 0;
  if (h->amg.par) {
    // this rank's share: its distributed levels (+ the contracted serial hierarchy on rank 0)
    for (int l = 0; l < h->amg.npar; l++) {
      i64 m, p, v, m2, p2, v2;
      level_bytes(*lev[l], m, p, v);
      level_bytes(*lev[l + 1], m2, p2, v2);
      const double D = 8.0 * lev[l]->n * lev[l]->b * lev[l]->b;
      B += 2.0 * (m + D) + 2.0 * p + 9.0 * v + 2.0 * v2;
    }
    if (h->amg.nested) B += ngsamg_b200_vcycle_bytes(reinterpret_cast<ngsamg_b200_t *>(h->amg.nested.get()));
    return B;
  }
  for (size_t l = 0; l + 1 < lev.size(); l++) {
    i64 m, p, v, m2, p2, v2;
    level_bytes(*lev[l], m, p, v);
    level_bytes(*lev[l + 1], m2, p2, v2);
    const double D = 8.0 * lev[l]->n * lev[l]->b * lev[l]->b;
    B += 2.0 * (m + D) + 2.0 * p + 9.0 * v + 2.0 * v2;
  }
  const Level &C = *lev.back();
  const double N = (double)C.n * C.b;
  B += 8.0 * N * N + 16.0 * N;
  return B;
}

double ngsamg_b200_last_ms(ngsamg_b200_t *h, int what)
{
  if (!h) return 0.0;
  switch (what) {
    case 0: return h->amg.ms_apply;
    case 1: return h->amg.ms_pcg;
    case 2: return h->amg.ms_setup;
    case 3: return h->amg.ms_rap;
    case 4: return h->amg.ms_host;
    case 5: return h->amg.bytes_rap;   // compulsory bytes of all Galerkin products (M_f + 2 P + M_c per level)
  }
  return 0.0;
}

int64_t ngsamg_b200_launch_count(ngsamg_b200_t *h) { return h ? h->amg.launches : 0; }

// ---- standalone sparse kernels ---------------------------------------------------------------------
static void upload_abi(const ngsamg_csr *A, DevCsr &d, cudaStream_t st)
{
  dev_csr_upload_raw(A->nrows, A->ncols, A->bh, A->bw, A->rowptr, A->col, A->val, d, st);
}

int ngsamg_b200_matmul_begin(const ngsamg_csr *A, const ngsamg_csr *B, int device, ngsamg_b200_spm **out, int64_t *nrows, int64_t *nnz)
{
  NGB_TRY
  require_device(device);
  check_csr(A, "matmul"); check_csr(B, "matmul");
  DevCsr dA, dB;
  auto r = std::make_unique<ngsamg_b200_spm>();
  r->st = nullptr;
  upload_abi(A, dA, nullptr);
  upload_abi(B, dB, nullptr);
  dev_spgemm(dA, dB, r->d, nullptr, nullptr);
  dev_csr_free(dA); dev_csr_free(dB);
  if (nrows) *nrows = r->d.nrows;
  if (nnz) *nnz = r->d.nnz;
  *out = r.release();
  NGB_CATCH
}

int ngsamg_b200_rap_begin(const ngsamg_csr *A, const ngsamg_csr *P, int device, ngsamg_b200_spm **out, int64_t *nrows, int64_t *nnz)
{
  NGB_TRY
  require_device(device);
  check_csr(A, "rap"); check_csr(P, "rap");
  DevCsr dA, dP, dPT, dPTA;
  upload_abi(A, dA, nullptr);
  upload_abi(P, dP, nullptr);
  dev_transpose(dP, dPT, nullptr, nullptr);
  auto r = std::make_unique<ngsamg_b200_spm>();
  r->st = nullptr;
  dev_spgemm(dPT, dA, dPTA, nullptr, nullptr);
  dev_spgemm(dPTA, dP, r->d, nullptr, nullptr);
  dev_csr_free(dA); dev_csr_free(dP); dev_csr_free(dPT); dev_csr_free(dPTA);
  if (nrows) *nrows = r->d.nrows;
  if (nnz) *nnz = r->d.nnz;
  *out = r.release();
  NGB_CATCH
}

int ngsamg_b200_transpose_begin(const ngsamg_csr *A, int device, ngsamg_b200_spm **out, int64_t *nrows, int64_t *nnz)
{
  NGB_TRY
  require_device(device);
  check_csr(A, "transpose");
  DevCsr dA;
  upload_abi(A, dA, nullptr);
  auto r = std::make_unique<ngsamg_b200_spm>();
  r->st = nullptr;
  dev_transpose(dA, r->d, nullptr, nullptr);
  dev_csr_free(dA);
  if (nrows) *nrows = r->d.nrows;
  if (nnz) *nnz = r->d.nnz;
  *out = r.release();
  NGB_CATCH
}

int ngsamg_b200_spm_fetch(ngsamg_b200_spm *m, int64_t *rowptr, int32_t *col, double *val)
{
  NGB_TRY
  if (!m) throw Error("null matrix handle");
  HostBsr h;
  dev_csr_download(m->d, h, m->st, val != nullptr);
  if (rowptr) std::memcpy(rowptr, h.rowptr.data(), sizeof(i64) * (h.nrows + 1));
  if (col) std::memcpy(col, h.col.data(), sizeof(i32) * h.nnz());
  if (val) std::memcpy(val, h.val.data(), sizeof(double) * h.nnz() * h.bs());
  dev_csr_free(m->d);
  delete m;
  NGB_CATCH
}

// ---- per-kernel timing for the roofline (bench.py) ---------------------------------------------------
// Launches ONE kernel of the V-cycle `reps` times on the library stream between two CUDA events and returns the average
// duration plus the algorithmic bytes of one launch (matrix part + dense diagonal arrays + each streamed vector once).
int ngsamg_b200_profile_kernel(ngsamg_b200_t *h, int level, int which, int reps, double *ms_avg, double *bytes)
{
  NGB_TRY
  Amg &a = ready(h);
  Level &L = get_level(a, level);
  if (level + 1 >= (int)a.lev.size()) throw Error("profile_kernel: level has no smoother / transfer");
  Level &C = *a.lev[level + 1];
  const double bb = 8.0 * L.b * L.b + 4.0, D = 8.0 * L.n * L.b * L.b, v = 8.0 * L.n * L.b, vc = 8.0 * C.n * C.b;
  const double pb = L.P.nnz * (8.0 * L.b * L.bc + 4.0);
  double B = 0;
  auto run = [&]() {
    switch (which) {
      case 0: a.tri_dispatch(L, false, false, true, L.rhs, nullptr, L.x, L.res); B = L.L.nnz * bb + 2 * D + 3 * v; break;
      case 1: a.spmv_part(L, 1, L.x, L.res, L.res, -1.0, 1.0, nullptr); B = L.U.nnz * bb + 3 * v; break;
      case 2: a.spmv_part(L, 2, L.x, L.rhs, L.tmp, -1.0, 1.0, nullptr); B = L.L.nnz * bb + D + 3 * v; break;
      case 3: a.tri_dispatch(L, true, true, false, L.tmp, L.x, L.y, nullptr); B = L.U.nnz * bb + D + 3 * v; break;
      case 4: ensure_scratch(a, L); a.spmv_part(L, 4, L.x, nullptr, L.wa, 1.0, 0.0, nullptr); B = (L.L.nnz + L.U.nnz) * bb + D + 2 * v; break;
      case 5: a.transfer(L.PT, L.res, nullptr, C.rhs, 1.0, 0.0, L.d_pt_rowmap); B = pb + v + vc; break;
      case 6: a.transfer(L.P, C.x, L.x, L.x, 1.0, 1.0); B = pb + 2 * v + vc; break;
      case 7: a.tri_dispatch(L, false, true, false, L.tmp, L.x, L.y, nullptr); B = L.L.nnz * bb + D + 3 * v; break;   // forward, RHS form
      case 8: a.tri_dispatch(L, true, false, true, L.rhs, nullptr, L.x, L.res); B = L.U.nnz * bb + 2 * D + 3 * v; break;  // backward, RES form
      case 9: a.dis2co(L, L.tmp); a.co2cu(L, L.tmp); B = 0; break;   // one DIS2CO + one CO2CU halo exchange (collective over the ranks)
      default: throw Error("profile_kernel: unknown kernel id");
    }
  };
  run();  // warm-up (also resolves lazy allocations)
  NGB_CUDA(cudaStreamSynchronize(a.st));
  cudaEventRecord(a.ev0, a.st);
  for (int r = 0; r < reps; r++) run();
  cudaEventRecord(a.ev1, a.st);
  NGB_CUDA(cudaEventSynchronize(a.ev1));
  float ms = 0;
  cudaEventElapsedTime(&ms, a.ev0, a.ev1);
  if (ms_avg) *ms_avg = ms / std::max(reps, 1);
  if (bytes) *bytes = B;
  if (const char *tf = std::getenv("NGSAMG_B200_TRACE_FILE")) {
    if ((which == 0 || which == 3) && L.tiled && L.tile_maxs > 2) {
      // CTA-per-tile sweep: 8 words per tile + the tile DAG (predecessor lists) for the offline analysis (scripts/analyze_ctile_trace.py)
      const i64 nt = L.ntiles;
      a.tri_trace = dev_alloc<unsigned long long>(nt * 16);
      NGB_CUDA(cudaMemsetAsync(a.tri_trace, 0, sizeof(unsigned long long) * nt * 16, a.st));
      run();
      NGB_CUDA(cudaStreamSynchronize(a.st));
      std::vector<unsigned long long> ht(nt * 16);
      NGB_CUDA(cudaMemcpy(ht.data(), a.tri_trace, sizeof(unsigned long long) * nt * 16, cudaMemcpyDeviceToHost));
      dev_free(a.tri_trace);
      a.tri_trace = nullptr;
      std::vector<i64> pp(nt + 1);
      NGB_CUDA(cudaMemcpy(pp.data(), which == 0 ? L.d_tile_pred_ptr : L.d_tile_succ_ptr, sizeof(i64) * (nt + 1), cudaMemcpyDeviceToHost));
      std::vector<i32> pl(std::max<i64>(pp[nt], 1));
      NGB_CUDA(cudaMemcpy(pl.data(), which == 0 ? L.d_tile_pred : L.d_tile_succ, sizeof(i32) * pp[nt], cudaMemcpyDeviceToHost));
      std::string fn = std::string(tf) + (which == 0 ? ".ctile.fwd" : ".ctile.bwd");
      if (FILE *f = std::fopen(fn.c_str(), "wb")) {
        const i64 np = pp[nt];
        std::fwrite(&nt, sizeof(i64), 1, f);
        std::fwrite(&np, sizeof(i64), 1, f);
        std::fwrite(pp.data(), sizeof(i64), nt + 1, f);
        std::fwrite(pl.data(), sizeof(i32), np, f);
        std::fwrite(ht.data(), sizeof(unsigned long long), ht.size(), f);
        std::fclose(f);
      }
    } else if ((which == 0 || which == 3) && L.rmL.ptr && a.tri_rm && !a.level_launch(L)) {
      // row-major warp-per-row sweep: 12 stamps per row (see k_gs_tri_rm)
      const i64 nr = L.npad;
      a.tri_trace = dev_alloc<unsigned long long>(nr * 12);
      NGB_CUDA(cudaMemsetAsync(a.tri_trace, 0, sizeof(unsigned long long) * nr * 12, a.st));
      run();
      NGB_CUDA(cudaStreamSynchronize(a.st));
      std::vector<unsigned long long> ht(nr * 12);
      NGB_CUDA(cudaMemcpy(ht.data(), a.tri_trace, sizeof(unsigned long long) * nr * 12, cudaMemcpyDeviceToHost));
      dev_free(a.tri_trace);
      a.tri_trace = nullptr;
      std::string fn = std::string(tf) + ".l" + std::to_string(level) + (which == 0 ? ".rm.fwd" : ".rm.bwd");
      if (FILE *f = std::fopen(fn.c_str(), "wb")) {
        const i64 nl = (i64)L.level_start.size();
        std::fwrite(&nr, sizeof(i64), 1, f);
        std::fwrite(&nl, sizeof(i64), 1, f);
        std::fwrite(L.level_start.data(), sizeof(i64), nl, f);
        std::fwrite(ht.data(), sizeof(unsigned long long), ht.size(), f);
        std::fclose(f);
      }
    } else if (which == 0 || which == 3) {
      const i64 ns = L.npad / 32;
      a.tri_trace = dev_alloc<unsigned long long>(ns * 3);
      NGB_CUDA(cudaMemsetAsync(a.tri_trace, 0, sizeof(unsigned long long) * ns * 3, a.st));
      run();
      NGB_CUDA(cudaStreamSynchronize(a.st));
      std::vector<unsigned long long> ht(ns * 3);
      NGB_CUDA(cudaMemcpy(ht.data(), a.tri_trace, sizeof(unsigned long long) * ns * 3, cudaMemcpyDeviceToHost));
      dev_free(a.tri_trace);
      a.tri_trace = nullptr;
      std::string fn = std::string(tf) + (which == 0 ? ".fwd" : ".bwd");
      if (FILE *f = std::fopen(fn.c_str(), "wb")) {
        const i64 nl = (i64)L.level_start.size();
        std::fwrite(&ns, sizeof(i64), 1, f);
        std::fwrite(&nl, sizeof(i64), 1, f);
        std::fwrite(L.level_start.data(), sizeof(i64), nl, f);
        std::fwrite(ht.data(), sizeof(unsigned long long), ht.size(), f);
        std::fclose(f);
      }
    }
  }
  NGB_CUDA(cudaGetLastError());
  a.check_watchdog();
  NGB_CATCH
}

// measurement aid: change a run-time tunable of the sweeps on a finalized hierarchy (the captured V-cycle graph is dropped)
int ngsamg_b200_set_tunable(ngsamg_b200_t *h, const char *name, double value)
{
  NGB_TRY
  Amg &a = ready(h);
  const std::string k = name ? name : "";
  if (k == "tri_sleep_ns") a.tri_sleep_ns = (unsigned)value;
  else if (k == "tri_repoll_ns") a.tri_repoll_ns = (unsigned)value;
  else if (k == "tri_prepoll") a.tri_prepoll = (int)value;
  else if (k == "tri_pollmode") a.tri_pollmode = (int)value;
  else if (k == "tri_gate_all") a.tri_gate_all = (int)value;
  else if (k == "tri_block_warp_rows") a.tri_block_warp_rows = (int)value;
  else if (k == "rm_spmv_rows") a.rm_spmv_rows = (i64)value;
  else if (k == "tri_regate") a.tri_regate = (int)value;
  else if (k == "tri_ctas_per_sm") { a.tri_ctas_per_sm = (int)value; for (int &c : a.tri_grid_cap) c = 0; }
  else if (k == "tri_rm") a.tri_rm = (int)value;
  else if (k == "tri_rm_rows_per_warp") a.tri_rm_rows_per_warp = std::max(1, (int)value);
  else if (k == "tri_rm_gate_rows") a.tri_rm_gate_rows = (i64)value;
  else if (k == "tri_small_rows") a.tri_small_rows = (i64)value;
  else if (k == "tri_level_launch_depth") a.tri_level_launch_depth = (int)value;
  else if (k == "tri_level_pdl") a.tri_level_pdl = (int)value;
  else if (k == "tri_level_launch_rows") a.tri_level_launch_rows = (i64)value;
  else if (k == "spmv_small_rows") a.spmv_small_rows = (i64)value;
  else if (k == "use_graph") a.use_graph = value != 0;
  else throw Error("set_tunable: unknown tunable '" + k + "'");
  if (a.vgraph) { NGB_CUDA(cudaStreamSynchronize(a.st)); cudaGraphExecDestroy(a.vgraph); a.vgraph = nullptr; }
  NGB_CATCH
}

// ---- host-side DOF-map construction -------------------------------------------------------------
struct ngsamg_b200_hostspm { HostBsr P; std::vector<i32> vmap; std::vector<double> cxyz; };

int ngsamg_b200_coarsen_begin(const ngsamg_csr *A, const uint8_t *free_mask, const double *vertex_xyz, int bcoarse, int max_per_row,
                              double min_frac, double omega, int smooth, int rounds, ngsamg_b200_hostspm **out, int64_t *ncoarse,
                              int64_t *nnz)
{
  NGB_TRY
  check_csr(A, "coarsen");
  HostBsr hA;
  copy_csr(A, hA);
  std::vector<double> xyz;
  if (vertex_xyz) xyz.assign(vertex_xyz, vertex_xyz + 3 * A->nrows);
  CoarsenOptions o;
  o.max_per_row = max_per_row; o.min_frac = min_frac; o.omega = omega; o.smooth = smooth != 0; o.rounds = rounds;
  auto r = std::make_unique<ngsamg_b200_hostspm>();
  build_prolongation(hA, free_mask, bcoarse, xyz, o, r->P, r->vmap, r->cxyz);
  if (ncoarse) *ncoarse = r->P.ncols;
  if (nnz) *nnz = r->P.nnz();
  *out = r.release();
  NGB_CATCH
}

int ngsamg_b200_coarsen_fetch(ngsamg_b200_hostspm *m, int64_t *rowptr, int32_t *col, double *val, int32_t *vmap, double *cxyz)
{
  NGB_TRY
  if (!m) throw Error("null handle");
  if (rowptr) std::memcpy(rowptr, m->P.rowptr.data(), sizeof(i64) * (m->P.nrows + 1));
  if (col) std::memcpy(col, m->P.col.data(), sizeof(i32) * m->P.nnz());
  if (val) std::memcpy(val, m->P.val.data(), sizeof(double) * m->P.nnz() * m->P.bs());
  if (vmap) std::memcpy(vmap, m->vmap.data(), sizeof(i32) * m->vmap.size());
  if (cxyz && !m->cxyz.empty()) std::memcpy(cxyz, m->cxyz.data(), sizeof(double) * m->cxyz.size());
  delete m;
  NGB_CATCH
}

}  // extern "C"
