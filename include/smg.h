/*
 * smg.h -- C ABI of the B200-native surface-multigrid V-cycle library (libsmg.so).
 *
 * This is the drop-in boundary for ONE hot path of HTDerekLiu/surface_multigrid_code:
 * the inner loop of mg_VCycle / min_quad_with_fixed_mg_solve plus the Galerkin
 * products of min_quad_with_fixed_mg_precompute.  Every entry point names the
 * reference interface it replaces (file:line under the reference tree).
 *
 * Conventions (identical to what the reference's Eigen objects hold):
 *   - sparse matrices are CSC with int32 indices: colptr[cols+1], rowidx[nnz],
 *     val[nnz]  (Eigen::SparseMatrix<double>::outerIndexPtr/innerIndexPtr/valuePtr
 *     of a compressed col-major matrix); explicit zeros are meaningful and kept;
 *   - dense blocks are column-major double, n x k with leading dimension n
 *     (Eigen::VectorXd / MatrixXd);
 *   - all pointers are HOST pointers owned by the caller unless the function name
 *     ends in _device; the library copies in/out;
 *   - every function returns an smg_status (0 = ok); smg_last_error() explains.
 *   - a handle is driven by one host thread at a time (the reference path is
 *     single-threaded and non-reentrant through `mg`).
 *
 * There is NO CPU fallback behind this ABI: without a CUDA device smg_create
 * fails with SMG_E_CUDA (except in plan-only mode, device == SMG_DEVICE_NONE,
 * which runs the host-side index/topology planning only and refuses all compute).
 */
#ifndef SMG_H
#define SMG_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SMG_VERSION 100 /* 0.1.0 */

typedef struct smg_handle smg_handle;

typedef enum smg_status {
  SMG_OK = 0,
  SMG_E_INVALID = 1,       /* bad argument / size mismatch */
  SMG_E_CUDA = 2,          /* CUDA runtime error or no device */
  SMG_E_NLEVELS = 3,       /* fewer than 2 levels (mg_precompute.cpp:39 TODO is not replicated) */
  SMG_E_NONFINITE = 4,     /* residual became NaN/Inf */
  SMG_E_STATE = 5,         /* call order violated (e.g. solve before precompute) */
  SMG_E_CUSOLVER = 6,      /* coarse factorisation failed (matrix not SPD?) */
  SMG_E_NCCL = 7,
  SMG_E_NOT_SYMMETRIC = 8, /* sparsity pattern of A is not symmetric */
  SMG_E_UNSUPPORTED = 9,
  SMG_E_INTERNAL = 10      /* a device-side wait timed out (halo exchange: a peer rank died) or an internal check
                              failed; results are invalid */
} smg_status;

typedef enum smg_smoother {
  /* exact reference order: lexicographic Gauss-Seidel (mg_VCycle.cpp:146-177),
   * parallelised by level-scheduling its dependency DAG; bit-identical results */
  SMG_SMOOTHER_WAVEFRONT = 0,
  /* multicolour Gauss-Seidel (fast mode): same fixed points, different sweep order */
  SMG_SMOOTHER_MULTICOLOUR = 1
} smg_smoother;

/* limits: right-hand-side columns per call (k), rows of the coarsest level (its inverse is
 * kept dense: 8 n^2 bytes during precompute) */
#define SMG_MAX_RHS 32
#define SMG_MAX_COARSE_ROWS 16384

#define SMG_DEVICE_CURRENT (-1)
#define SMG_DEVICE_NONE (-2) /* plan-only handle: host index planning, no CUDA */

typedef struct smg_options {
  int pre_relax;  /* default 2  (min_quad_with_fixed_mg.cpp:102,324) */
  int post_relax; /* default 2  (min_quad_with_fixed_mg.cpp:103,325) */
  int smoother;   /* smg_smoother, default SMG_SMOOTHER_MULTICOLOUR */
  int device;     /* CUDA ordinal, SMG_DEVICE_CURRENT or SMG_DEVICE_NONE */
  int use_graph;  /* 1: replay each V-cycle as one CUDA graph (default) */
  int verbose;    /* 1: print the residual per iteration like the reference (cpp:334,349) */
  int locality_reorder; /* 1: order rows inside a phase by a BFS (Cuthill-McKee) rank (default) */
  int sigma;      /* SELL sort window in rows (default 256; 1 = no length sorting) */
  int patch_rows; /* target rows per patch of the communication-avoiding smoother that runs a
                     whole relax call of a small level in one launch;
                     0 = automatic, < 0 = off (one kernel per colour phase on every level) */
  int reserved0;
  int reserved[6];
} smg_options;

void smg_default_options(smg_options *opt);
int smg_version(void);
const char *smg_status_string(int status);
/* message of the last failure on this handle (never NULL) */
const char *smg_last_error(const smg_handle *h);

/* ---- lifetime ------------------------------------------------------------ */
int smg_create(smg_handle **out, const smg_options *opt /* NULL = defaults */);
void smg_destroy(smg_handle *h);

/* ---- hierarchy ------------------------------------------------------------
 * Replaces reading mg[lv].P_full / .P / .PT as left by mg_precompute
 * (src/mg_precompute.cpp:57-78; struct src/mg_data.h:11-44).
 * n_levels >= 2.  For lv = 1..n_levels-1, P_*[lv-1] is the CSC prolongation from
 * level lv to level lv-1: n_rows[lv-1] x n_rows[lv].  (The hierarchy BUILD stays
 * on the CPU in the caller, exactly as in the reference.) */
int smg_set_hierarchy(smg_handle *h, int n_levels, const int *n_rows,
                      const int *const *P_colptr, const int *const *P_rowidx,
                      const double *const *P_val);

/* ---- min_quad_with_fixed_mg_precompute ------------------------------------
 * Replaces src/min_quad_with_fixed_mg.cpp:137-257 (n_known >= 0: variant with
 * fixed values; `known` in caller order, duplicates allowed) and :3-51
 * (n_known < 0: variant without fixed values; `known` ignored).
 * A is n x n CSC and must have a symmetric pattern (reference asserts symmetry,
 * cpp:14,149).  Does: unknown = setdiff; LHS/Auk slices; P row slicing and
 * column pruning (> 1e-15) level by level; Galerkin A_l = PT*A_{l-1}*P on the
 * GPU; +1e-12 on the coarsest diagonal; A_diag; coarse factorisation. */
int smg_precompute(smg_handle *h, int n, const int *A_colptr, const int *A_rowidx,
                   const double *A_val, const int *known, int n_known);

/* Numeric-only refresh with the sparsity pattern (and known set) of the last
 * smg_precompute: new values of A, same work as precompute minus the index
 * planning.  This is what 05_example_mean_curvature_flow/main.cpp:74 needs every
 * time step. */
int smg_update_values(smg_handle *h, const double *A_val);

/* ---- min_quad_with_fixed_mg_solve -----------------------------------------
 * Replaces src/min_quad_with_fixed_mg.cpp:288-361 (fixed) / :80-135 (free), all
 * overloads (tol default 1e-3, max_iter default 20 live in the adapter).
 * RHS, z0, z: n x k col-major; known_val: n_known x k (may be NULL when there
 * are no fixed values).  r_his must hold max_iter doubles; *n_his receives the
 * number of residual measurements; *converged = !(last measured residual > tol)
 * (the reference's return value, including its stale-residual quirk). */
int smg_solve(smg_handle *h, const double *RHS, const double *known_val, const double *z0, int k,
              double tol, int max_iter, double *z, double *r_his, int *n_his, int *converged);

/* Same, but RHS / known_val / z0 / z are DEVICE pointers on the handle's device
 * (r_his, n_his, converged stay host).  No host<->device traffic except the
 * per-iteration residual scalar. */
int smg_solve_device(smg_handle *h, const double *d_RHS, const double *d_known_val,
                     const double *d_z0, int k, double tol, int max_iter, double *d_z,
                     double *r_his, int *n_his, int *converged);

/* ---- mean-curvature-flow step with device-side assembly -------------------------------
 * Replaces the per-step body of 05_example_mean_curvature_flow/main.cpp:66-76:
 *   M = massmatrix(U, F, BARYCENTRIC); LHS = M - delta * L; RHS = M * U;
 *   min_quad_with_fixed_mg_precompute(LHS, ...); min_quad_with_fixed_mg_solve(RHS, Upre, ..., tol, ...)
 * Only U crosses the bus per step (24 bytes per vertex each way instead of the whole LHS).
 * Precondition: smg_precompute in the variant without fixed values (n_known < 0) with any
 * SPD matrix that has the sparsity pattern of L (M - delta * L has it: L stores its diagonal;
 * e.g. the first step's system, or I - L).
 * smg_mcf_setup: F is nF x 3 column-major (Eigen::MatrixXi), L_val are the values of the
 * cotangent matrix in the CSC order of that pattern (igl::cotmatrix(V, F, L), computed once).
 * smg_mcf_step: U, U_out are nV x 3 column-major; the rest as smg_solve (z0 = U).
 * Tests: tests/test_gpu_zz_mcf.py (GPU, against the host path), tests/test_mcf_core.py (CPU). */
int smg_mcf_setup(smg_handle *h, int nV, int nF, const int *F, const double *L_val, double delta);
int smg_mcf_step(smg_handle *h, const double *U, double tol, int max_iter, double *U_out,
                 double *r_his, int *n_his, int *converged);

/* ---- mg_VCycle.h operators (host buffers, level sizes per smg_level_rows) ---
 * Vectors are in the caller's (reference) row numbering of that level. */
/* mg_VCycle (src/mg_VCycle.cpp:3-59): one V(pre,post)-cycle from level lv down;
 * u is in/out. */
int smg_vcycle(smg_handle *h, int lv, int pre, int post, const double *B, double *u, int k);
/* relax (src/mg_VCycle.cpp:113-178): `iters` Gauss-Seidel sweeps, u in/out */
int smg_relax(smg_handle *h, int lv, int iters, const double *B, double *u, int k);
/* A (src/mg_VCycle.cpp:62-70): Au = mg[lv].A * u */
int smg_apply_A(smg_handle *h, int lv, const double *u, double *Au, int k);
/* residual r = B - A u (src/mg_VCycle.cpp:41-42), fused */
int smg_residual(smg_handle *h, int lv, const double *B, const double *u, double *r, int k);
/* ||B - A u||_F (src/min_quad_with_fixed_mg.cpp:110,332), fused, deterministic */
int smg_residual_norm(smg_handle *h, int lv, const double *B, const double *u, int k,
                      double *norm);
/* restrict (src/mg_VCycle.cpp:72-81): Rx = mg[lv+1].PT * x */
int smg_restrict(smg_handle *h, int lv, const double *x, double *Rx, int k);
/* prolong (src/mg_VCycle.cpp:83-92): Px = mg[lv+1].P * x */
int smg_prolong(smg_handle *h, int lv, const double *x, double *Px, int k);
/* coarseSolve (src/mg_VCycle.cpp:181-201): u = u + A_coarsest^-1 B */
int smg_coarse_solve(smg_handle *h, const double *B, double *u, int k);

/* ---- multi-GPU: row-range partition of the fine levels over the ranks of one node ----
 * The reference is a single-threaded CPU code (no MPI/NCCL anywhere): this block has no
 * reference counterpart, it is the row-range partition BASELINE.json's north_star asks for.
 * One process (or thread) per GPU, one handle per rank.  Levels 0 .. dist_levels-1 are
 * partitioned by rows (strips of a breadth-first order of the level's graph); every rank
 * runs the smoother / residual / restriction / prolongation of its own rows and exchanges
 * halo values with the other ranks through peer-mapped device memory (stores over NVLink +
 * epoch flags, inside the same CUDA graph as the compute kernels; no host round trip, no
 * NCCL call on the data path).  Coarser levels and the coarse direct solve are replicated.
 * Call order: smg_create, smg_dist_init, smg_dist_get_handle, [the host all-gathers the
 * blobs, e.g. torch.distributed.all_gather / MPI_Allgather / a file], smg_dist_connect,
 * then smg_set_hierarchy / smg_precompute / smg_solve as usual.  Every compute entry point
 * becomes COLLECTIVE: all ranks make the same calls with the same (replicated) arguments and
 * all ranks receive the complete result.  Synchronise the ranks (host barrier) before
 * smg_destroy. */
/* comm_bytes: size of the peer-mapped staging buffer (0 = 256 MiB).  Plan-only handles
 * accept the call and partition the plan without allocating anything. */
int smg_dist_init(smg_handle *h, int rank, int world, size_t comm_bytes);
int smg_dist_handle_bytes(void);
/* this rank's export blob (smg_dist_handle_bytes() bytes): CUDA IPC handle + pid */
int smg_dist_get_handle(smg_handle *h, void *blob);
/* all_blobs: world blobs in rank order.  Buffers of ranks living in the same process are
 * used directly (peer access), others are opened with cudaIpcOpenMemHandle. */
int smg_dist_connect(smg_handle *h, const void *all_blobs);
/* Host-side helper for callers without MPI / torch.distributed: all-gather `bytes` bytes per
 * rank through files in `dir`, a directory every rank of the node can see (e.g.
 * /dev/shm/<job>); `tag` distinguishes rendezvous rounds inside one directory.  Rank r writes
 * dir/tag.r (temp file + rename), then polls for the other world-1 files.  Returns
 * SMG_E_INTERNAL after timeout_ms.  No CUDA involved.  A file carries its writer's pid and a
 * per-process sequence number: only files of live processes are accepted (the ranks must share
 * a pid namespace: one node, one container), every rank acknowledges what it read, and a rank
 * removes its file once all peers have acknowledged it -- leftovers of an earlier run with the
 * same dir / tag are never taken for a peer's blob.  Use a different tag per rendezvous round. */
int smg_rendezvous_files(const char *dir, const char *tag, int rank, int world, const void *mine,
                         size_t bytes, void *all, int timeout_ms);
/* smg_dist_get_handle + smg_rendezvous_files + smg_dist_connect in one call */
int smg_dist_connect_files(smg_handle *h, const char *dir, const char *tag, int timeout_ms);
/* exact = 1: halo exchange after every colour (the N-rank smoother is then the same
 * multicolour Gauss-Seidel as on one GPU); 0 (default): one exchange per sweep (Gauss-Seidel
 * inside a rank, Jacobi coupling across ranks); 2: one exchange per relax call (the sweeps of
 * a call see the other ranks' rows as of the start of the call).  dist_levels < 0: automatic (always level 0;
 * below it, dist_min_rows > 0: levels with at least that many rows per rank; dist_min_rows = 0: levels of at
 * least 500 000 rows with at least 50 000 rows per rank -- smaller levels are latency-bound and their halo
 * exchanges cost more than the split saves).  Before smg_precompute. */
int smg_dist_set_options(smg_handle *h, int exact, int dist_levels, int dist_min_rows);
/* out[8]: [0] rank [1] world [2] partitioned levels [3] halo exchanges issued so far
 * [4] connected [5] staging doubles per (peer, parity) slot */
int smg_dist_info(const smg_handle *h, int64_t *out);
/* out[8] for level lv: [0] layout (0 plain, 1 partitioned, 2 split) [1] parts
 * [2] rows owned by this rank [3] halo rows of u this rank receives per exchange
 * [4] halo rows it sends [5] halo rows of r received [6] first owned row [7] end */
int smg_dist_level_info(const smg_handle *h, int lv, int64_t *out);
/* owner rank of every row of level lv (reference numbering) */
int smg_dist_get_part(const smg_handle *h, int lv, int *part_of_row);
typedef enum smg_exchange_id { SMG_X_HALO_U = 0, SMG_X_HALO_R = 1, SMG_X_HALO_PU = 2,
                               SMG_X_GATHER = 3 } smg_exchange_id;
/* rows (reference numbering of level lv) whose values travel src -> dst in exchange
 * `which`; idx may be NULL to query *n */
int smg_dist_get_exchange(const smg_handle *h, int lv, int which, int src, int dst, int *idx,
                          int *n);

/* ---- index / topology outputs (bit-exact parity surface) ------------------ */
int smg_num_levels(const smg_handle *h);
int smg_level_rows(const smg_handle *h, int lv); /* rows of mg[lv].A, <0 on error */
int smg_num_unknown(const smg_handle *h);
/* data.unknown (src/min_quad_with_fixed_mg.cpp:156-158,178), ascending */
int smg_get_unknown(const smg_handle *h, int *unknown);
/* kept columns of mg[lv].P after pruning (cpp:190-204); *n_keep = -1 when level
 * lv was not pruned (loop broke earlier, cpp:216-219) */
int smg_get_keep(const smg_handle *h, int lv, int *keep /* may be NULL */, int *n_keep);
typedef enum smg_matrix_id { SMG_MAT_A = 0, SMG_MAT_P = 1, SMG_MAT_PT = 2, SMG_MAT_LHS = 3,
                             SMG_MAT_AUK = 4 } smg_matrix_id;
int smg_matrix_dims(const smg_handle *h, int lv, int which, int *rows, int *cols, int *nnz);
/* copies mg[lv].A / .P / .PT / data.LHS / data.Auk back to the host as CSC
 * (values are read back from the device for A) */
int smg_matrix_copy(smg_handle *h, int lv, int which, int *colptr, int *rowidx, double *val);
/* mg[lv].A_diag (src/min_quad_with_fixed_mg.cpp:244-246) */
int smg_get_diag(smg_handle *h, int lv, double *diag);
/* smoother schedule: number of phases (colours or wavefront levels) on level lv
 * and, if phase_of_row != NULL, the phase of every row (reference numbering). */
int smg_get_phases(const smg_handle *h, int lv, int *n_phases, int *phase_of_row);
/* the library's row numbering of level lv: perm[r] = reference row stored at position r.
 * Rows are grouped (DESIGN.md sections 3 and 6): by smoother phase; on a row-partitioned level
 * by (part, phase); on the replicated level below one by (phase, part).  *n_groups receives the
 * number of groups and, if group_ptr != NULL, group_ptr[0..*n_groups] their row offsets. */
int smg_get_row_order(const smg_handle *h, int lv, int *perm, int *n_groups, int *group_ptr);
/* storage statistics of level lv's SELL-32 matrix: stored (padded) entries */
int smg_level_padded_nnz(const smg_handle *h, int lv, int64_t *padded);

/* level statistics for roofline arithmetic, out[8]:
 *  [0] rows n_l  [1] stored entries of mg[lv].A (reference pattern, zeros included)
 *  [2] entries of the compute pattern (what the kernels must read, unpadded)
 *  [3] SELL-32 padded entries of A  [4] non-zero entries of mg[lv].P (lv >= 1)
 *  [5] padded entries of P by fine row  [6] padded entries of PT by coarse row
 *  [7] smoother phases */
int smg_level_stats(const smg_handle *h, int lv, int64_t *out);

/* ---- patch smoother (csrc/patch.hpp) -------------------------------
 * Host-side check of the communication-avoiding schedule of level lv (works on plan-only
 * handles): lays out one relax call of `iters` sweeps in patches of about target_rows rows
 * (kind 0: + residual and restriction, 1: prolongation first) and, if verify != 0, proves
 * symbolically that every row update reads its neighbours at exactly the version the
 * phase-by-phase schedule would.  out[10]: [0] patches [1] owned rows [2] local rows (owned +
 * halo, summed over patches) [3] rows whose right-hand side is read [4] row updates
 * [5] largest patch blob in bytes [6] largest shared-memory vector count [7] bytes of all blobs
 * [8] most passes of 256 rows over all phases of any patch [9] stored sweep entries */
int smg_patch_plan(const smg_handle *h, int lv, int kind, int iters, int target_rows, int smem_limit,
                   int verify, int64_t *out);
/* 1 if the last solve ran its loop on the device (one graph launch: conditional WHILE node around the
 * V-cycle, residual test in solve_decide_kernel), 0 if the host drove the iterations (graphs off, ranks
 * sharing a device, max_iter > 4096, or a driver without conditional graph nodes) */
int smg_solve_on_device(const smg_handle *h);
/* number of patches level lv is smoothed with on this handle (0: one kernel per colour phase) */
int smg_level_patched(const smg_handle *h, int lv);

/* ---- measurement ----------------------------------------------------------
 * Times `reps` back-to-back launches of one hot-path kernel on level lv with k
 * right-hand sides using CUDA events on the handle's stream; returns the mean
 * milliseconds per launch (for a smoother: per full sweep = all phases) and the
 * number of kernel launches per rep.  Operands are the level's resident work
 * vectors (contents are clobbered; call before a solve, not inside one). */
typedef enum smg_kernel_id {
  SMG_K_RESIDUAL = 0,   /* r = b - A u */
  SMG_K_RELAX_SWEEP = 1,/* one Gauss-Seidel sweep (all phases) */
  SMG_K_RESTRICT = 2,   /* b_{lv+1} = PT r_lv */
  SMG_K_PROLONG_ADD = 3,/* u_lv += P u_{lv+1} */
  SMG_K_RESIDUAL_NORM = 4,
  SMG_K_COARSE_SOLVE = 5,
  SMG_K_VCYCLE = 6,     /* one full V-cycle from level 0 (graph when enabled) */
  /* one iteration of the solve loop (min_quad_with_fixed_mg.cpp:330-347): residual
   * norm with its host read-back, then one V-cycle */
  SMG_K_MG_ITERATION = 7,
  /* the pre-smoothing of a V-cycle: opt.pre_relax Gauss-Seidel sweeps back to back */
  SMG_K_RELAX_PRE = 8
} smg_kernel_id;
int smg_time_kernel(smg_handle *h, int which, int lv, int k, int reps, int flush_l2,
                    float *ms_per_rep, int *launches_per_rep);
/* Device timeline of one solve-loop iteration (residual norm + V-cycle, replayed as a
 * CUDA graph with PDL edges): for every kernel launch the earliest CTA start and the
 * latest CTA end (%globaltimer), in microseconds from the first start.  `names`
 * receives one '\n'-terminated label per event.  Profiling aid, not on the hot path. */
int smg_trace_iteration(smg_handle *h, int k, int max_events, char *names, int names_cap,
                        double *t0_us, double *t1_us, int *n_events);
/* kernel launches issued by this handle since creation (your-kernels only;
 * graph replays count their kernel nodes) */
int64_t smg_launch_count(const smg_handle *h);
/* milliseconds of the last smg_solve / smg_precompute phases: [0] H2D, [1] device
 * solve loop, [2] D2H, [3] precompute host planning, [4] precompute device */
int smg_get_timings(const smg_handle *h, double *ms, int n);
/* cudaStream_t of the handle (as void*), so callers can order their own work */
void *smg_get_stream(const smg_handle *h);

#ifdef __cplusplus
}
#endif
#endif /* SMG_H */
