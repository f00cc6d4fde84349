/*
 * ngsamg_b200.h -- C ABI of the B200-native NgsAMG preconditioner-apply hot path.
 *
 * The reference (LukasKogler/NgsAMG) has NO C ABI: it is a C++ add-on that registers
 * `ngcomp::Preconditioner` subclasses with NGSolve (src/base/utils/amg_register.hpp:79-98) and is
 * driven through NGSolve's BaseMatrix interface.  Each entry point below names the reference
 * interface it replaces (paths relative to the reference root); INTEGRATION.md shows the NGSolve-side
 * adapter a maintainer would add on top of these calls.
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on error; ngsamg_b200_last_error() gives the text
 *     (the reference throws ngcore::Exception, e.g. src/base/precond/amg_pc.cpp:430,446).
 *   - matrices are block-CSR in NGSolve SparseMatrix<Mat<bh,bw,double>> layout: rowptr[nrows+1] (int64),
 *     col[nnz] (int32, ascending per row), val[nnz*bh*bw] (row-major blocks).  Vectors are AoS doubles.
 *   - one handle == one GPU == one caller thread (the reference's AMGMatrix is not re-entrant either:
 *     shared work vectors mutated inside const Mult, src/base/solve/amg_matrix.cpp:187-189).
 *   - vector arguments may be HOST or DEVICE pointers; the library detects which
 *     (cudaPointerGetAttributes) and stages host buffers through pinned memory.
 *   - there is no CPU fallback: without a CUDA device every call fails with an error.
 *   - the Gauss-Seidel sweeps are sync-free kernels whose CTAs wait for each other's results: they are launched cooperatively
 *     (cudaLaunchAttributeCooperative), so the driver rejects a grid that cannot be co-resident instead of letting it dead-lock, and
 *     every wait loop shares one watchdog flag (a timed-out sweep makes apply / pcg return an error).  Run ONE process per GPU and
 *     do not share the device with other long-running kernels while a V-cycle is in flight (MPS time-slicing is fine, it is only slow).
 */
#ifndef NGSAMG_B200_H
#define NGSAMG_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ngsamg_b200 ngsamg_b200_t; /* opaque: owns all device memory of one hierarchy */

typedef struct ngsamg_csr {
  int64_t nrows, ncols; /* block rows / block cols */
  int32_t bh, bw;       /* block height / width (1x1 H1; 3x3, 6x6, 3x6, 6x3 elasticity) */
  const int64_t *rowptr;
  const int32_t *col;
  const double *val;
} ngsamg_csr;

typedef struct ngsamg_level_info {
  int64_t n;          /* block rows of the level matrix          (AMGMatrix::GetNDof, amg_matrix.cpp:396-)  */
  int32_t b;          /* block size                                                                  */
  int64_t nnz;        /* stored blocks of A_l                                                        */
  int64_t nnz_prol;   /* stored blocks of P_l (0 on the last level)                                  */
  int64_t ncoarse;    /* columns of P_l                                                              */
  int32_t bcoarse;    /* block width of P_l                                                          */
  int32_t gs_depth;   /* number of dependency levels of the level-scheduled Gauss-Seidel sweep       */
  int64_t bytes_matrix, bytes_prol, bytes_vec; /* algorithmic bytes M_l, P_l, v_l of SURVEY.md §8d   */
} ngsamg_level_info;

/* ---- construction ------------------------------------------------------------------------------
 * ngsamg_b200_create replaces the strict-algebraic constructor + InitLevel + (deferred) FinalizeLevel:
 *   NgsAMG.h1_scal(mat, freedofs, **kwargs)            src/h1/python_h1.cpp:24-33
 *   BaseAMGPC(A, flags, name); InitLevel(freedofs)     src/base/precond/amg_pc.cpp:346-353, 398-410
 * `type`      : registered preconditioner name, "h1_scal" | "h1_2d" | "h1_3d" | "elast_2d" | "elast_3d",
 *               with or without the "NgsAMG." / "ngs_amg." prefix (amg_register.hpp:85-97, elasticity.hpp:104-140)
 * `A`         : the assembled fine matrix (host arrays, copied)
 * `free_mask` : freedofs BitArray as bytes, one per block row (NULL = all free)
 * `vertex_xyz`: nrows x 3 vertex coordinates (elasticity: rigid-body modes; NULL for H1)
 * `flags`     : nflags (key, value) string pairs, keys as in the reference with the "ngs_amg_" prefix -- max_levels, max_coarse_size,
 *               mg_cycle (V | W | BS, Options::MG_CYCLE amg_pc.cpp:293), clev, sm_type, sm_steps, sm_symm (+ "_spec" per-level lists),
 *               regularize_cmats, sp_max_per_row, sp_min_frac, sp_omega, prol_type, log_level; B200-specific ones carry a "b200_" infix
 *               (amg_pc.hpp:168-172, Options::SetFromFlags amg_pc.cpp:270-339); unknown keys are ignored like
 *               NGSolve Flags does.  Lists (`*_spec`) are comma separated.
 * The hierarchy is not built yet; call ngsamg_b200_set_prolongations (optional) then ngsamg_b200_finalize. */
int ngsamg_b200_create(const char *type, const ngsamg_csr *A, const uint8_t *free_mask, const double *vertex_xyz,
                       const char *const *flag_keys, const char *const *flag_vals, int nflags, int device,
                       ngsamg_b200_t **out);

/* Inject the DOF maps instead of running the built-in coarsening: P[l] maps level l+1 -> level l.
 * Replaces AMGMatrix(DOFMap([ProlMap...]), smoothers, ...), src/base/solve/python_solve.cpp:57-76. */
int ngsamg_b200_set_prolongations(ngsamg_b200_t *h, int nprol, const ngsamg_csr *P);

/* BaseAMGPC::FinalizeLevel -> BuildAMGMat (amg_pc.cpp:413-434, 565-736): coarsening + prolongations
 * (unless injected), Galerkin RAP on the device (ProlMap::AssembleMatrix -> RestrictMatrix, dof_map.cpp:817-834,
 * utils_sparseMM.hpp:93-109), smoother setup (BuildGSSmoother amg_pc.cpp:1096-1138, GSS3::CalcDiags
 * gssmoother.cpp:142-170), coarsest inverse (CoarseLevelInv amg_pc.cpp:843-928). */
int ngsamg_b200_finalize(ngsamg_b200_t *h);

void ngsamg_b200_destroy(ngsamg_b200_t *h);
const char *ngsamg_b200_last_error(void);

/* ---- preconditioner apply (the hot path) -------------------------------------------------------
 * x = C b     BaseAMGPC::Mult -> AMGMatrix::Mult -> SmoothV   amg_pc.cpp:467-470, amg_matrix.cpp:160-307, 377-378
 * x += s C b  AMGMatrix::MultAdd                               amg_matrix.cpp:385-389
 * MultTrans == Mult and MultTransAdd == MultAdd (amg_matrix.cpp:381-393): call the same entry points. */
int ngsamg_b200_apply(ngsamg_b200_t *h, const double *b, double *x);
int ngsamg_b200_apply_add(ngsamg_b200_t *h, double s, const double *b, double *x);

/* y += s * A_level * x   (SparseMatrix::MultAdd on a level matrix; CG's A*s; BaseSmoother::CalcResiduum
 * base_smoother.hpp:132-142).  level 0 == the matrix given to create(). */
int ngsamg_b200_spmv_add(ngsamg_b200_t *h, int level, double s, const double *x, double *y);

/* BaseSmoother::Smooth / SmoothBack of the smoother of `level` with the reference flag protocol
 * (base_smoother.hpp:68-112; GSS3::Smooth/SmoothBack gssmoother.cpp:349-398; ProxySmoother :181-196).
 * Smoother-only entry == NgsAMG.CreateHybridGSS(...).Smooth, src/base/smoothers/python_smoothers.cpp:144-194.
 * On a distributed level of a multi-rank handle this is HybridBaseSmoother::Smooth / SmoothBack (hybrid_base_smoother.cpp:214-574):
 * collective, x CUMULATED, b and res DISTRIBUTED, RES or RHS form picked by SmoothImpl's cost heuristic (:242-290); spmv_add is then
 * HybridBaseMatrix::MultAdd, y(DISTRIBUTED) += s (M + G) x(CUMULATED) (hybrid_matrix.cpp:393-411). */
int ngsamg_b200_smooth(ngsamg_b200_t *h, int level, double *x, const double *b, double *res, int res_updated,
                       int update_res, int x_zero, int backwards);

/* ProlMap::TransferF2C / AddC2F on level `level` (dof_map.cpp:633-654, 694-709):
 *   restrict: xc = P_l^T xf ;  prolong_add: xf += fac * P_l xc */
int ngsamg_b200_restrict(ngsamg_b200_t *h, int level, const double *xf, double *xc);
int ngsamg_b200_prolong_add(ngsamg_b200_t *h, int level, double fac, const double *xc, double *xf);
/* the exact coarsest-level solve alone: x = A_L^-1 rhs on the free dofs (crs_inv->Mult, amg_matrix.cpp:228-233; CoarseLevelInv,
 * amg_pc.cpp:843-928); x = 0 if the hierarchy has no coarse inverse (clev=none).  Building block of AMGMatrix::CINV (amg_matrix.cpp:407-435). */
int ngsamg_b200_coarse_solve(ngsamg_b200_t *h, const double *rhs, double *x);

/* PCG with the V-cycle as preconditioner == ngsolve.krylovspace.CGSolver(mat, pre, maxsteps, tol) as the reference
 * tests call it (tests/h1/amg_utils.py:346-349).  errors has room for maxsteps+1 doubles (errors[0] = err0);
 * *iters = CGSolver.iterations.  rhs/x host or device. */
int ngsamg_b200_pcg(ngsamg_b200_t *h, const double *rhs, double *x, double tol, int maxsteps, int *iters,
                    double *errors);

/* ---- introspection (AMGMatrix::GetNLevels/GetNDof amg_matrix.cpp:396-, GetOC :551-582) -------------- */
int ngsamg_b200_num_levels(ngsamg_b200_t *h);
/* which kernel sweeps the level (measurement / test aid): 0 row-level sync-free or per-colour launches, 1 warp per tile, 2 CTA per tile,
 * 3 CTA per tile on tile images prepared at setup, 4 warp per row on the row-major copy of a small level; -1 = no such level */
int ngsamg_b200_level_sweep_kind(ngsamg_b200_t *h, int level);
/* sm_type = bgs (block Gauss-Seidel, BSmoother2 loc_block_gssmoother_impl.hpp:244-268, 516-541, 656-706): the blocks of `level` as the
 * reference builds them from the coarse map (GetGSBlocks, amg_pc_vertex_impl.hpp:1171-1269): block_of[v] = coarse vertex of v, -1 = in no
 * block (not smoothed).  n entries.  Fails if the level is not smoothed by bgs (levels without a coarse map fall back to gs). */
int ngsamg_b200_get_gs_blocks(ngsamg_b200_t *h, int level, int32_t *block_of);
/* measurement aid: change a run-time tunable of the sweep kernels ("tri_sleep_ns", "tri_prepoll", "tri_rm", "tri_rm_rows_per_warp",
 * "tri_small_rows", "tri_level_launch_depth", "tri_level_launch_rows", "spmv_small_rows", "use_graph", ...) on a finalized hierarchy */
int ngsamg_b200_set_tunable(ngsamg_b200_t *h, const char *name, double value);
int ngsamg_b200_level_info(ngsamg_b200_t *h, int level, ngsamg_level_info *info);
/* copy the level matrix A_l / the prolongation P_l (original DOF numbering) into caller arrays sized from
 * level_info: rowptr[n+1], col[nnz], val[nnz*b*b] resp. val[nnz_prol*b*bcoarse].  Any pointer may be NULL.
 * These are what the bit-exact pattern / DOF-map parity tests read. */
int ngsamg_b200_get_level_matrix(ngsamg_b200_t *h, int level, int64_t *rowptr, int32_t *col, double *val);
int ngsamg_b200_get_prolongation(ngsamg_b200_t *h, int level, int64_t *rowptr, int32_t *col, double *val);
/* level work vectors after the last apply: which = 0 x_level, 1 rhs_level, 2 res_level (amg_matrix.cpp:19-26) */
int ngsamg_b200_get_level_vector(ngsamg_b200_t *h, int level, int which, double *out);
/* position of every row of `level` in the Gauss-Seidel sweep: rank[i] == i (the reference's order, gssmoother.cpp:195-315)
 * unless the optional multicolour smoother was requested for the fine level (flag ngs_amg_b200_sm_order=multicolor). */
int ngsamg_b200_get_sweep_order(ngsamg_b200_t *h, int level, int32_t *rank);
/* AMGMatrix::GetOC (src/base/solve/amg_matrix.cpp:551-582): occs = [OC, OC_0, OC_1, ...], OC_l = cycle factor * nops_l / nze_0 for the
 * levels that carry a smoother (nze = scalar non-zeros, nops = nze times nsteps * (symm ? 2 : 1) under a ProxySmoother; factor 1 / 2^l / 2(1+l)
 * for V / W / BS), 0 for the exactly solved coarsest level, OC = their sum.  Returns the number of entries (pass occs = NULL to query it). */
int ngsamg_b200_operator_complexities(ngsamg_b200_t *h, double *occs, int cap);
/* occs[0] of the above */
double ngsamg_b200_operator_complexity(ngsamg_b200_t *h);
/* one V-cycle run eagerly with CUDA events around every phase (measurement aid; reference timers: "AMGMatrix::Mult", "GSSmoother", the
 * hybrid smoother's comm timers hybrid_base_smoother.cpp:498-574).  ms[7] = { triangular sweeps, parallel halves (U / L+D passes), transfers,
 * halo exchanges, G products, coarse part (contracted hierarchy or coarsest solve), other }.  Collective on a distributed hierarchy. */
int ngsamg_b200_apply_phases(ngsamg_b200_t *h, const double *b, double *x, double *ms);
/* algorithmic bytes of one V(1,1)-cycle, SURVEY.md §8d formula B_V, from the actual level sizes */
double ngsamg_b200_vcycle_bytes(ngsamg_b200_t *h);
/* device milliseconds (CUDA events on the library stream) of the last apply / pcg call, setup phases */
double ngsamg_b200_last_ms(ngsamg_b200_t *h, int what); /* 0 apply, 1 pcg, 2 setup total, 3 setup RAP (device transpose + both SpGEMMs of every level), 4 setup host, 5 = compulsory BYTES of those products (M_f + 2 P + M_c) */
/* number of kernel launches issued by the library since create (bench.py's gpu_launches) */
int64_t ngsamg_b200_launch_count(ngsamg_b200_t *h);

/* ---- multi-rank (one rank == one GPU) ---------------------------------------------------------------
 * The reference runs one MPI rank per subdomain: ParallelDofs lists the DOFs shared with every neighbour rank, the fine
 * matrix is the rank's sub-assembled (DISTRIBUTED) contribution, and vectors are exchanged through the DCCMap
 * (src/base/linalg/dcc_map.cpp:76-302, 494-543).  The B200 path keeps that decomposition: one process and one handle per GPU.
 *
 * ngsamg_halo  == ParallelDofs::GetDistantProcs() + GetExchangeDofs(p): for neighbour k the shared local DOFs
 *   ex_dofs[ex_ptr[k] .. ex_ptr[k+1]), ascending, and the k-th DOF shared with rank p here is the k-th DOF shared with this
 *   rank on p (NGSolve's convention).  peers ascending.
 * ngsamg_comm  == the communicator.  Setup-phase (host) traffic goes through two caller-supplied callbacks -- an NGSolve
 *   adapter implements them with its NgMPI_Comm, the test harness with a process group (gloo) or plain threads:
 *     exchange      : post sendbuf[k] (sendbytes[k] bytes) to peers[k] and receive recvbytes[k] bytes from it into recvbuf[k],
 *                     for all k, then return (MPI_Isend/Irecv + Waitall); sizes are known to both sides.
 *     allreduce_sum : in-place sum over all ranks of n doubles (every rank receives the result).
 *   The per-sweep device traffic (DIS2CO / CO2CU halo exchange, CG dot products, coarse-level gather) uses NCCL point-to-point
 *   over NVLink when `nccl` holds an ncclComm_t (see ngsamg_b200_nccl_*); with nccl == NULL it is staged through host memory
 *   and the two callbacks (any MPI; also how several ranks share ONE GPU in the parity tests).
 *   Callbacks return 0 on success. */
typedef struct ngsamg_halo {
  int32_t npeers;
  const int32_t *peers;
  const int64_t *ex_ptr;   /* npeers + 1 */
  const int32_t *ex_dofs;
} ngsamg_halo;

typedef struct ngsamg_comm {
  int32_t rank, size;
  void *ctx;
  int (*exchange)(void *ctx, int32_t npeers, const int32_t *peers, const void *const *sendbuf, const int64_t *sendbytes,
                  void *const *recvbuf, const int64_t *recvbytes);
  int (*allreduce_sum)(void *ctx, double *vals, int32_t n);
  void *nccl; /* ncclComm_t of the same ranks, or NULL */
} ngsamg_comm;

/* Multi-rank constructor: like ngsamg_b200_create, with A the rank's LOCAL sub-assembled matrix over its local DOFs (shared
 * interface DOFs are duplicated on every sharer; the global matrix is the sum over the ranks), free_mask / vertex_xyz over the
 * local DOFs (consistent on shared DOFs).  Replaces BaseAMGPC on a ParallelMatrix (amg_pc.cpp:398-434) + the hybrid smoother
 * setup (HybridMatrix / HybridGSSmoother, hybrid_matrix.cpp:17-307, gssmoother.cpp:603-700) + CtrMap coarse-level
 * redistribution (dof_contract.cpp).  finalize / apply / pcg are then COLLECTIVE over the ranks of `comm`:
 *   apply: b is the rank's DISTRIBUTED rhs, x the CUMULATED result (AMGMatrix::SmoothV amg_matrix.cpp:160-307)
 *   pcg  : rhs DISTRIBUTED, x CUMULATED; dot products are all-reduced.
 * `comm` (and the callback context) must stay valid for the life of the handle. */
int ngsamg_b200_create_parallel(const char *type, const ngsamg_csr *A, const uint8_t *free_mask, const double *vertex_xyz,
                                const ngsamg_halo *halo, const ngsamg_comm *comm, const char *const *flag_keys,
                                const char *const *flag_vals, int nflags, int device, ngsamg_b200_t **out);

/* NCCL plumbing for the device data path (the library dlopens libnccl.so.2; plain pointers only): rank 0 creates the 128-byte
 * unique id, the host application broadcasts it, every rank calls comm_init with its device.  comm_destroy frees it. */
int ngsamg_b200_nccl_unique_id(char id[128]);
int ngsamg_b200_nccl_comm_init(const char id[128], int rank, int size, int device, void **nccl_comm);
int ngsamg_b200_nccl_comm_destroy(void *nccl_comm);

/* multi-rank introspection (parity tests): sharing information of `level` on this rank (two-call: NULL arrays -> sizes). */
int ngsamg_b200_get_halo(ngsamg_b200_t *h, int level, int32_t *npeers, int32_t *peers, int64_t *ex_ptr, int32_t *ex_dofs);
/* which = 0: M, 1: G of the hybrid split of `level` (local numbering of the level), 2: sizes only; mod_diag: n*b*b doubles */
int ngsamg_b200_get_hybrid(ngsamg_b200_t *h, int level, int which, int64_t *nnz, int64_t *rowptr, int32_t *col, double *val,
                           double *mod_diag);
/* number of distributed levels (levels 0 .. npar-1 are smoothed on every rank; level npar is contracted onto rank 0) */
int ngsamg_b200_num_parallel_levels(ngsamg_b200_t *h);
/* transport of the per-sweep halo exchange of a distributed level: 0 host-staged through the callbacks, 1 NCCL send/recv, 2 NVLink peer
 * memory (IPC-mapped receive buffers, push / pull kernels -- the default with an NCCL communicator); -1 = not a distributed level */
int ngsamg_b200_halo_transport(ngsamg_b200_t *h, int level);
/* rank 0 only: the serial hierarchy below the contracted level (borrowed handle; all single-rank entry points work on it) and,
 * for every rank r, the map local DOF of the contracted level -> DOF of the merged level (CtrMap dof_maps). */
ngsamg_b200_t *ngsamg_b200_get_contracted(ngsamg_b200_t *h);
int ngsamg_b200_get_contraction_map(ngsamg_b200_t *h, int rank, int64_t *n, int32_t *map);

/* host-only pieces of the multi-rank setup (no device needed; collective over comm) -- the CPU tests drive them over gloo:
 * hybrid split + modified diagonal of one level.  Results are fetched with ngsamg_b200_hybrid_host_fetch. */
/* one step of the multi-rank DOF-map construction (what finalize() runs per distributed level): assembled local matrix ->
 * coarsening that respects the sharing classes (vertices only merge with vertices shared by the same ranks, rows of shared DOFs are
 * bit-identical on every sharer; cf. the EQC-wise agglomeration + hierarchic prolongation, vertex_factory_impl.hpp:1845-1848).
 * fetch copies P, vmap, coarse coordinates and the sharing lists of the COARSE level (peers[npeers_coarse], ex_ptr[npeers_coarse+1],
 * ex_dofs[nshared_coarse]) and frees the handle.  Host only, collective over comm. */
typedef struct ngsamg_b200_parcoarsen ngsamg_b200_parcoarsen;
int ngsamg_b200_coarsen_parallel_begin(const ngsamg_csr *A, const uint8_t *free_mask, const double *vertex_xyz, const ngsamg_halo *halo,
                                       const ngsamg_comm *comm, int bcoarse, int max_per_row, double min_frac, double omega, int smooth,
                                       int rounds, ngsamg_b200_parcoarsen **out, int64_t *ncoarse, int64_t *nnz, int32_t *npeers_coarse,
                                       int64_t *nshared_coarse);
int ngsamg_b200_coarsen_parallel_fetch(ngsamg_b200_parcoarsen *m, int64_t *rowptr, int32_t *col, double *val, int32_t *vmap, double *cxyz,
                                       int32_t *peers, int64_t *ex_ptr, int32_t *ex_dofs);

typedef struct ngsamg_b200_hybrid_host ngsamg_b200_hybrid_host;
int ngsamg_b200_hybrid_host_begin(const ngsamg_csr *A, const uint8_t *free_mask, const ngsamg_halo *halo, const ngsamg_comm *comm,
                                  ngsamg_b200_hybrid_host **out, int64_t *nnz_m, int64_t *nnz_g);
int ngsamg_b200_hybrid_host_fetch(ngsamg_b200_hybrid_host *m, int64_t *m_rowptr, int32_t *m_col, double *m_val, int64_t *g_rowptr,
                                  int32_t *g_col, double *g_val, double *mod_diag, int32_t *sweep_rank, uint8_t *master);

/* host-only: contraction of one distributed level onto rank 0 -- CtrMap::DoAssembleMatrix (src/base/coarsening/dof_contract.cpp:557-727) for ONE
 * group with master 0: the members' local (DISTRIBUTED) matrices are remapped through the dof maps, the column lists merged (structural union:
 * entries that cancel stay) and coinciding entries summed in rank order.  Collective over the ranks of `comm`; rank 0 receives the merged matrix and
 * the maps local dof -> merged dof of every rank (master dofs numbered rank by rank, a ghost takes its master's number).  Lets the CPU tests check
 * the host contraction the multi-GPU path uses without a device.  map_total = length of the concatenated dof maps (rank 0). */
typedef struct ngsamg_b200_contract_host ngsamg_b200_contract_host;
int ngsamg_b200_contract_host_begin(const ngsamg_csr *A, const uint8_t *free_mask, const ngsamg_halo *halo, const ngsamg_comm *comm,
                                    ngsamg_b200_contract_host **out, int64_t *n_merged, int64_t *nnz_merged, int64_t *map_total);
int ngsamg_b200_contract_host_fetch(ngsamg_b200_contract_host *m, int64_t *rowptr, int32_t *col, double *val, int64_t *map_ptr, int32_t *dof_map);

/* ---- standalone sparse kernels (setup path) ----------------------------------------------------
 * Galerkin product on the device: Ac = (P^T A) P.   RestrictMatrix<H,W>, utils_sparseMM.hpp:93-109;
 * MatMultABImpl utils_sparseMM.cpp:107-238; TransposeSPMImpl :54-93.
 * Two-call protocol: rap_begin computes the product and returns an opaque result + its sizes; rap_fetch copies it
 * out and frees it. */
typedef struct ngsamg_b200_spm ngsamg_b200_spm;
int ngsamg_b200_rap_begin(const ngsamg_csr *A, const ngsamg_csr *P, int device, ngsamg_b200_spm **out, int64_t *nrows,
                          int64_t *nnz);
int ngsamg_b200_matmul_begin(const ngsamg_csr *A, const ngsamg_csr *B, int device, ngsamg_b200_spm **out,
                             int64_t *nrows, int64_t *nnz);
int ngsamg_b200_transpose_begin(const ngsamg_csr *A, int device, ngsamg_b200_spm **out, int64_t *nrows, int64_t *nnz);
int ngsamg_b200_spm_fetch(ngsamg_b200_spm *m, int64_t *rowptr, int32_t *col, double *val);

/* ---- measurement hook (bench.py roofline) ---------------------------------------------------------
 * Launch ONE kernel of the V-cycle on `level` `reps` times between two CUDA events on the library stream.
 * which: 0 forward triangular sweep (RES form), 1 U-pass, 2 (L+D)-pass, 3 backward triangular sweep (RHS form),
 *        4 plain SpMV, 5 restriction, 6 prolongation-add.
 * ms_avg = average device time of one launch; bytes = algorithmic bytes of one launch. */
int ngsamg_b200_profile_kernel(ngsamg_b200_t *h, int level, int which, int reps, double *ms_avg, double *bytes);

/* ---- DOF-map construction (host side, no device needed) ------------------------------------------
 * The built-in coarsening that finalize() runs per level when no prolongations were injected: pairwise agglomeration +
 * smoothed prolongation in the spirit of BuildCoarseMap / BuildCoarseDOFMap (src/base/factory/vertex_factory_impl.hpp:503-548,
 * 796-865); see ngsamg_b200/csrc/coarsen.cpp.  bcoarse = coarse block size (== A->bh except elasticity level 0: 3 -> 6).
 * coarsen_fetch copies P (rowptr[n+1], col, val), the vertex map vmap[n] (-1 = Dirichlet/dropped) and the coarse vertex
 * coordinates cxyz[ncoarse*3] (only if vertex_xyz was given) and frees the handle. */
typedef struct ngsamg_b200_hostspm ngsamg_b200_hostspm;
int ngsamg_b200_coarsen_begin(const ngsamg_csr *A, const uint8_t *free_mask, const double *vertex_xyz, int bcoarse,
                              int max_per_row, double min_frac, double omega, int smooth, int rounds,
                              ngsamg_b200_hostspm **out, int64_t *ncoarse, int64_t *nnz);
int ngsamg_b200_coarsen_fetch(ngsamg_b200_hostspm *m, int64_t *rowptr, int32_t *col, double *val, int32_t *vmap,
                              double *cxyz);

/* ---- pseudo-inverse of one diagonal block (host only; ngsamg_b200/csrc/dense.cpp) ---------------------------------------------
 * What the smoother set-up applies to every diagonal block when ngs_amg_regularize_cmats asks for pinv smoothers: the reference's
 * CalcPseudoInverseTryNormal(Mat<N,N>&) (src/base/utils/utils_denseLA.hpp:1549-1562) with its zero-row detection (:1237-1405), the direct
 * inverse attempt TryDirectInverse_simple (utils_denseLA.cpp:458-555) and the eigenvalue fall-back (:1474-1519).  m: n x n, row-major, in
 * place.  Exposed so that the CPU tests can compare it with the reference's own code without a device.  Returns non-zero on bad arguments. */
int ngsamg_b200_block_pinv(int n, double *m);
/* RegularizeMatrix of the elasticity preconditioners (src/elasticity/elasticity_pc_impl.hpp:711-763, local branch), one diagonal block of the
 * coarsest matrix: dim 3 / n 6 = RegTM<0,6,6> (utils_denseLA.hpp:1198-1234), dim 2 / n 3 = unit rotational entry when |m(2,2)| < 1e-8; other
 * shapes are left alone.  Applied by finalize() before the coarsest matrix is inverted when ngs_amg_regularize_cmats is set (amg_pc.cpp:861). */
int ngsamg_b200_block_regularize(int n, double *m, int dim);

/* ---- two-level (tile) schedule of the sequential Gauss-Seidel sweep (host only; ngsamg_b200/csrc/tiles.hpp) ---------------
 * Groups the smoothed rows of a level matrix into compact tiles of <= max_rows graph-neighbouring rows, orders the tiles by the levels of
 * the tile dependency DAG and the rows of a tile by their tile-local dependency levels.  Executing tile after tile, level after level, is
 * a topological order of the row DAG of GSS3's sweep (gssmoother.cpp:195-315), i.e. it reproduces the reference's result while the number
 * of cross-SM hops on the critical path falls from the row-DAG depth to the tile-DAG depth.  The device kernel consuming the schedule is
 * experimental (flag ngs_amg_b200_tile_sweep); these entry points let tests validate the schedule without a device.
 * rounds < 0: the setup default (box-shaped clusters when the matrix is numbered like a structured grid, else pairwise clustering).
 * info[9] = { ok, ntiles, npad, nonfree_pad, tile_depth, max_local_levels, merged_tiles, violations (self-check), npred }.
 * fetch copies perm[n], tile_slice[ntiles+1], tile_nlev[ntiles], row_lvl[npad], pred_ptr[ntiles+1], pred[npred] and frees the handle. */
typedef struct ngsamg_b200_tiles ngsamg_b200_tiles;
int ngsamg_b200_tile_schedule_begin(const ngsamg_csr *A, const uint8_t *smoothed_mask, const int32_t *sweep_rank, int rounds, int max_rows,
                                    ngsamg_b200_tiles **out, int64_t *info);
/* the same with caller-supplied clusters (cluster[i] >= 0 for every smoothed row) instead of the built-in pairwise clustering, e.g. the boxes
 * of a structured grid: the clusters are a hint, the scheduler still splits whatever is not executable as one piece */
int ngsamg_b200_tile_schedule_hinted(const ngsamg_csr *A, const uint8_t *smoothed_mask, const int32_t *sweep_rank, const int32_t *cluster, int max_rows,
                                     ngsamg_b200_tiles **out, int64_t *info);
int ngsamg_b200_tile_schedule_fetch(ngsamg_b200_tiles *m, int32_t *perm, int32_t *tile_slice, int32_t *tile_nlev, uint8_t *row_lvl,
                                    int64_t *pred_ptr, int32_t *pred);
const char *ngsamg_b200_tiles_last_error(void);

#ifdef __cplusplus
}
#endif
#endif
