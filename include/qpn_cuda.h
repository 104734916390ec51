/*
 * libqpn_cuda -- C ABI of the B200 (sm_100a) equilibrium engine for Quadratic
 * Program Networks.  This is the drop-in boundary for the numeric hot path of
 * forrestlaine/QuadraticProgramNetworks.jl v0.4.0: each entry point replaces one
 * reference call site (cited per function).  The Julia side reaches these with
 * `ccall`; nothing here depends on torch or on C++ types.
 *
 * Conventions
 *  - fp64 everywhere; matrices are column-major (Julia native); batched vectors
 *    are  n x batch  column-major, i.e. instance b starts at  ptr + b*n.
 *  - +-Inf are passed as IEEE infinities.
 *  - Return value: 0 ok, <0 error (message via qpn_last_error).  Per-instance
 *    outcomes are only reported in status arrays (StatusCode of avi.jl:1-6):
 *    1 SUCCESS, 2 RAY_TERM, 3 MAX_ITERS, 4 FAILURE.  No exceptions, no callbacks.
 *  - Host entry points copy in/out on the handle's stream and do not retain host
 *    pointers past return.  `_dev` entry points take device pointers of the
 *    handle's device and run asynchronously on the stream given (0 = handle's).
 *    A handle keeps scratch that is not ordered between streams (plan buffers, the
 *    tableau workspace of the global-memory path, the staging arena, the cycle-check
 *    history of a resident level): every `_dev` call on one handle must use the SAME
 *    stream, and a host-pointer call on that handle may only follow once that stream
 *    has been synchronised.  Concurrent streams need one handle each (they are cheap:
 *    the network path creates one per host thread).
 *  - One handle per GPU; calls on one handle must be serialised by the caller.
 *  - There is no CPU fallback: without a usable CUDA device qpn_create fails.
 */
#ifndef QPN_CUDA_H
#define QPN_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct qpn_handle qpn_handle;

#define QPN_SUCCESS 1
#define QPN_RAY_TERM 2
#define QPN_MAX_ITERS 3
#define QPN_FAILURE 4

/* basis / active-set codes written to basis_out (same coding as
 * comp_indices' primary sets, avi_solutions.jl:568-585) */
#define QPN_AT_LOWER 1
#define QPN_BASIC 2
#define QPN_AT_UPPER 3
#define QPN_FIXED 4

/* Matrix operand of an AVI: either dense column-major (dense != NULL) or CSC
 * (colptr/rowval/nzval, PATH's Cint indices as in avi.jl:11-12).  index_base is 1
 * for Julia's SparseMatrixCSC, 0 for C.  When is_shared == 0 the dense array is
 * n x n x batch and nzval is nnz x batch (the pattern is always shared). */
typedef struct {
    const double *dense;
    const int32_t *colptr;
    const int32_t *rowval;
    const double *nzval;
    int32_t nnz;
    int32_t index_base;
    int32_t is_shared;
} qpn_matrix;

int qpn_create(int device, qpn_handle **out);
int qpn_destroy(qpn_handle *h);
const char *qpn_last_error(qpn_handle *h);       /* h may be NULL: last create error */
int qpn_device(qpn_handle *h);
/* Number of kernel launches this handle has issued (bench.py's gpu_launches). */
int64_t qpn_launch_count(qpn_handle *h);
int qpn_synchronize(qpn_handle *h);
/* Of those, launches that took the global-memory tableau path (sizes beyond the shared-memory tableau). */
int64_t qpn_big_launch_count(qpn_handle *h);

/*
 * Replaces PATHSolver.solve_mcp at /root/reference/src/avi.jl:64-70 plus the check at
 * avi.jl:71-75 (`solve_avi`).  Finds z with (M z + q) complementary to l <= z <= u,
 * started from z0, by bounded-variable complementary pivoting (one CTA per instance).
 *   q, z0, z_out:  n x batch.   l, u:  n (lu_is_shared) or n x batch.
 *   status_out, pivots_out: batch.   basis_out: n x batch int8 (may be NULL).
 *   max_pivots <= 0 selects 50 n + 100.
 */
int qpn_avi_solve_batched(qpn_handle *h, int n, int batch, const qpn_matrix *M, const double *q,
                          const double *l, const double *u, int lu_is_shared, const double *z0,
                          int max_pivots, double *z_out, int32_t *status_out, int32_t *pivots_out,
                          int8_t *basis_out);
int qpn_avi_solve_batched_dev(qpn_handle *h, int n, int batch, const qpn_matrix *M, const double *q,
                              const double *l, const double *u, int lu_is_shared, const double *z0,
                              int max_pivots, double *z_out, int32_t *status_out,
                              int32_t *pivots_out, int8_t *basis_out, void *stream);

/*
 * Replaces check_avi_solution, /root/reference/src/avi.jl:148-156.
 *   bad_out[b] = number of violated conditions (0 = solution ok); r_out (n x batch, may be
 *   NULL) = M z + q.
 */
int qpn_check_avi_batched(qpn_handle *h, int n, int batch, const qpn_matrix *M, const double *q,
                          const double *l, const double *u, int lu_is_shared, const double *z,
                          double tol, int32_t *bad_out, double *r_out);

/* A generalized AVI (struct GAVI, avi.jl:29-39), dense blocks shared by the batch:
 *   (M z + N w + o) comp. l1 <= z1 <= u1 ;  z2 comp. l2 <= A z + B w <= u2,  z = [z1; z2]. */
typedef struct {
    int32_t d1, d2, np;
    const double *M;   /* d1 x (d1+d2) */
    const double *N;   /* d1 x np      */
    const double *o;   /* d1           */
    const double *l1, *u1;
    const double *A;   /* d2 x (d1+d2) */
    const double *B;   /* d2 x np      */
    const double *l2, *u2;
} qpn_gavi;

/*
 * Replaces solve_gavi, /root/reference/src/avi.jl:101-111: the presolve projection
 * find_closest_feasible! (avi.jl:79-99, OSQP in the reference), the GAVI->AVI lift
 * `convert` (avi.jl:113-128) and solve_avi, fused in one kernel.
 *   w: np x batch, z0: (d1+d2) x batch, z_out: (d1+d2) x batch,
 *   basis_out: (d1+2 d2) x batch (may be NULL), zfull_out: (d1+2 d2) x batch (may be NULL).
 */
int qpn_gavi_solve_batched(qpn_handle *h, const qpn_gavi *g, int batch, const double *w,
                           const double *z0, int presolve, int max_pivots, double *z_out,
                           double *zfull_out, int32_t *status_out, int32_t *pivots_out,
                           int8_t *basis_out);
int qpn_gavi_solve_batched_dev(qpn_handle *h, const qpn_gavi *g, int batch, const double *w,
                               const double *z0, int presolve, int max_pivots, double *z_out,
                               double *zfull_out, int32_t *status_out, int32_t *pivots_out,
                               int8_t *basis_out, void *stream);

/*
 * Replaces comp_indices(gavi, z, w), /root/reference/src/avi_solutions.jl:511-612 (the
 * request branch is inert).  mask_out: (d1+d2) x batch, one 4-bit mask per
 * index: bit0 -> set 1, bit1 -> set 2, bit2 -> set 3, bit3 -> set 4 (block 2: 5..8).
 * A mask of 0 marks an index for which the reference's @assert would fire.
 */
int qpn_comp_indices_batched(qpn_handle *h, const qpn_gavi *g, int batch, const double *z,
                             const double *w, double tol, int8_t *mask_out);

/*
 * Replaces Base.in(x, ::Poly; tol), /root/reference/src/sets.jl:820-825,850-853, for
 * npoly polyhedra over the same embedded dimension d, stacked row-wise:
 *   A: mtot x d, l/u: mtot, rl/ru: mtot (1 = strict '<', 0 = '<='; NULL = all closed),
 *   poly_ptr: npoly+1 row offsets.  x: d x npts.
 *   in_out: npoly x npts (uint8), in_out[p + npoly*j] = (x_j in poly p).
 */
int qpn_halfspace_in_batched(qpn_handle *h, int npoly, int d, int mtot, const int32_t *poly_ptr,
                             const double *A, const double *l, const double *u, const uint8_t *rl,
                             const uint8_t *ru, int npts, const double *x, double tol,
                             uint8_t *in_out);

/* One node's view of verify_solution: Qd = Q[dec,:] (nd x nv), qd = q[dec], stacked
 * constraint rows A (m x nv), l, u, dec (nd 0-based indices into x). */
typedef struct {
    int32_t nd, nv, m;
    const double *Qd, *qd, *A, *l, *u;
    const int32_t *dec;
} qpn_node;

/*
 * Replaces verify_solution, /root/reference/src/qp_processing.jl:57-149 (without the
 * check_convexity / debug branches): feasibility at 1e-3, active sets at 1e-2, dual
 * recovery  lam = Abar \ qt  (Householder QR), acceptance at tol, and the
 * sign-constrained least-squares fallback (PATH in the reference) through the same
 * pivoting solve.  x: nv x batch.
 *   solution_out: batch (1/0);  lam_out: m x batch;  how_out: batch (0 infeasible,
 *   1 unconstrained, 2 least squares, 3 fallback accepted, 4 fallback rejected,
 *   5 fallback failed);  active_out: m x batch (0 inactive, 1 lower, 2 upper, 3 both),
 *   either may be NULL.
 */
int qpn_verify_solution_batched(qpn_handle *h, const qpn_node *node, int batch, const double *x,
                                double tol, uint8_t *solution_out, double *lam_out,
                                int32_t *how_out, int8_t *active_out);

/* A level of a network with no child solution pieces (bottom level, or a flat Nash
 * game): its players and the level GAVI assembled by solve_qep (avi.jl:382-404). */
typedef struct {
    int32_t nv, nplayers;
    const qpn_node *players;
    qpn_gavi gavi;
    const int32_t *dec;     /* gavi decision indices into x (nd_level of them) */
    int32_t nd_level;
    const int32_t *par;     /* parameter indices into x (gavi.np of them) */
    int32_t max_iters;      /* QPNetOptions.max_iters, programs.jl:63 */
    int32_t num_projections;/* cycle check, algorithm.jl:14-30; 0 disables */
    const double *proj;     /* nv x num_projections */
} qpn_level;

/*
 * Replaces the iterate-until-equilibrium loop of solve_base! for one level without
 * children (/root/reference/src/algorithm.jl:13-118: process_qp -> verify_solution per
 * player, solve_qep when not an equilibrium, the 1e-4 disagreement test and the cycle
 * check), fused in one kernel, one CTA per instance.
 *   x_init, x_out: nv x batch.  solved_out: batch (1 solved, 0 not).
 *   iters_out, pivots_out: batch.  lam_out (may be NULL): sum(m_p) x batch.
 */
int qpn_level_equilibrium_batched(qpn_handle *h, const qpn_level *lv, int batch,
                                  const double *x_init, double *x_out, uint8_t *solved_out,
                                  int32_t *iters_out, int32_t *pivots_out, double *lam_out);
int qpn_level_equilibrium_batched_dev(qpn_handle *h, const qpn_level *lv, int batch,
                                      const double *x_init, double *x_out, uint8_t *solved_out,
                                      int32_t *iters_out, int32_t *pivots_out, double *lam_out,
                                      void *stream);

/*
 * Resident form: upload a level's problem data (players, GAVI blocks, index maps,
 * projection vectors) once -- SURVEY.md 8e: "problem matrices broadcast once" -- and run
 * batches against it.  `qpn_level_equilibrium_resident` takes HOST x_init / outputs and
 * copies only those; `_resident_dev` takes DEVICE x_init / outputs and is asynchronous on
 * `stream` (0 = the handle's stream).  lam_out may be NULL.
 */
typedef struct qpn_level_dev qpn_level_dev;
int qpn_level_upload(qpn_handle *h, const qpn_level *lv, qpn_level_dev **out);
int qpn_level_release(qpn_handle *h, qpn_level_dev *lvd);
int qpn_level_equilibrium_resident(qpn_handle *h, qpn_level_dev *lvd, int batch, const double *x_init,
                                   double *x_out, uint8_t *solved_out, int32_t *iters_out,
                                   int32_t *pivots_out, double *lam_out);
int qpn_level_equilibrium_resident_dev(qpn_handle *h, qpn_level_dev *lvd, int batch,
                                       const double *x_init, double *x_out, uint8_t *solved_out,
                                       int32_t *iters_out, int32_t *pivots_out, double *lam_out,
                                       void *stream);

/*
 * Engine options (no reference counterpart).  "force_big" = 1 routes every pivoting solve through
 * the global-memory tableau path that otherwise serves only sizes beyond the shared-memory
 * tableau (lifted n > ~166; up to n = 1,536); "big_ctas_per_sm" caps that path's resident CTAs; "big_slot_in_smem" = 0 keeps
 * that path's tableau slot in global memory even when a resident level's compact slot (swept rows only) would fit
 * shared memory (default 1); "big_smem_threads" > 0 fixes the threads per CTA of that shared-memory-slot form
 * (default 0: by the number of swept rows).
 */
int qpn_set_option(qpn_handle *h, const char *name, int64_t value);

/* Shape of a resident level's precomputed plans: out[0..7] = {lifted AVI size n, its live columns after the
 * plan, pivots the plan ran once for the whole batch, presolve AVI size, its live columns, its plan pivots,
 * 1 if the level runs on the global-memory tableau path, 0}. */
int qpn_level_info(qpn_handle *h, qpn_level_dev *lvd, int32_t *out);

/* Page-lock a host buffer the caller owns (a Julia `Matrix{Float64}` is pageable: `GC.@preserve` keeps it alive but does
 * not pin it) so that the host-pointer entry points reach it without staging: cudaHostRegister with the portable and
 * mapped flags.  Unregister before the buffer is freed.  Registration costs ~ 0.1 ms per MB: do it once per buffer that
 * is reused, not per call. */
int qpn_host_register(qpn_handle *h, void *ptr, size_t bytes);
int qpn_host_unregister(qpn_handle *h, void *ptr);
/* sizeof of every struct that crosses this ABI, so that a binding (ctypes, Julia) can check its own mirror at load
 * time: out[0..5] = {qpn_matrix, qpn_gavi, qpn_node, qpn_level, qpn_net_desc, 0}. */
int qpn_abi_struct_sizes(int32_t *out);

/* Device buffers owned by the handle (for callers without their own allocator). */
int qpn_malloc(qpn_handle *h, size_t bytes, void **dptr);
int qpn_free(qpn_handle *h, void *dptr);
int qpn_memcpy_h2d(qpn_handle *h, void *dst, const void *src, size_t bytes);
int qpn_memcpy_d2h(qpn_handle *h, void *dst, const void *src, size_t bytes);


/* =====================================================================================================
 * Networks with children: the batched solve(qpn, inits::Matrix) of BASELINE.json's north_star.
 *
 * Replaces, for a whole batch of independent instances, the recursion of solve_base!
 * (/root/reference/src/algorithm.jl:1-127) with process_qp / combine (src/qp_processing.jl:151-291), the
 * IntersectionRoot walk (src/intersection.jl:55-151) and collect(LocalGAVISolutions)
 * (src/avi_solutions.jl:200-215,241-321,400-496).  The host side is a native state machine (csrc/net/): every
 * instance advances until it needs numbers, the pending requests of all instances are regrouped by (kind, node /
 * level GAVI) and run as one kernel launch per group over a list of instance slots whose x stays resident on the
 * GPU; local pieces, projections and the LP predicates of the set algebra are memoised by exact problem data and
 * shared by every instance, host thread and batch of the net.
 *
 * The network is passed as flat arrays (what setup(:name) produces; players, polys, levels 0-based here):
 */
typedef struct {
    int32_t nv, nplayers, nlevels, npolys;
    const double *Q;                       /* nplayers x nv x nv, row-major per player (symmetric)      programs.jl:172-201 */
    const double *q;                       /* nplayers x nv */
    const int32_t *var_ptr, *var_idx;      /* CSR over players: the player's own variables (QP.var_indices) */
    const int32_t *con_ptr, *con_idx;      /* CSR over players: its constraint polys (constraint_indices order) */
    const int32_t *child_ptr, *child_idx;  /* CSR over players: network_edges after add_edges! (programs.jl:274-285) */
    const int32_t *level_of;               /* per player: 0-based level (network_depth_map) */
    const int32_t *poly_ptr;               /* npolys + 1 row offsets into poly_A / poly_l / poly_u */
    const double *poly_A;                  /* rows x nv, ROW-major: one normalised Slice per row (sets.jl:68-92) */
    const double *poly_l, *poly_u;
    int32_t max_iters, num_projections, exploration_vertices, gen_solution_map, check_for_cycling;  /* QPNetOptions */
    const uint8_t *remove_subsets_at;      /* per level: 1 = remove_subsets (levels_to_remove_subsets); NULL = every level */
    const double *proj;                    /* num_projections vectors of nv entries (algorithm.jl:10-12) */
} qpn_net_desc;

typedef struct qpn_net qpn_net;
int qpn_net_create(qpn_handle *h, const qpn_net_desc *desc, qpn_net **out);
int qpn_net_destroy(qpn_net *net);
const char *qpn_net_last_error(qpn_net *net);
/* "threads": host threads that drive the batch (each with its own stream), default 2; "profile": see qpn_net_profile. */
int qpn_net_set_option(qpn_net *net, const char *name, int64_t value);
/*
 * solve(qpn, inits::Matrix): inits nv x batch (host).  x_out: nv x batch -- x_opt where solved_out[b] = 1, the
 * reference's x_fail otherwise (algorithm.jl:116,125).  level_iters_out (may be NULL): nlevels x batch loop passes of
 * solve_base! per level; error_out (may be NULL): batch, 0 = none, 1 cycling detected, 2 AVI solve error,
 * 3 disagreement between verify and solve_qep, 4 max_iters, 5 empty solution graph, 6 comp_indices assertion,
 * 7 too many solutions to combine, 8 solution graphs not populated, 9 cycle check without projections (low byte;
 * for 2 the next byte holds the StatusCode solve_qep returned and the third byte the 0-based level).
 */
int qpn_net_solve_batched(qpn_net *net, int batch, const double *inits, double *x_out, uint8_t *solved_out,
                          int32_t *level_iters_out, int32_t *error_out);
/* The same with inits / x_out in DEVICE memory of the net's GPU (the flags and counters are produced on the host). */
int qpn_net_solve_batched_dev(qpn_net *net, int batch, const double *inits_dev, double *x_out_dev, uint8_t *solved_out,
                              int32_t *level_iters_out, int32_t *error_out);
/* Kernel accounting since the net was created.  out[0..23]: for k in {verify, solve_qep, membership, grouping (the
 * segmented sort of every cohort's members by their answers, the boundary scan and the gather of one representative's
 * answers per part), cycle check}: out[4k] launches, out[4k+1] units (instances; pairs for membership), out[4k+2] summed
 * duration in ms -- durations only accumulate while option "profile" = 1 (every launch is then bracketed by CUDA events
 * on its stream); out[20] / out[21]: bytes copied host-to-device / device-to-host. */
int qpn_net_profile(qpn_net *net, double *out);
/* The solution graphs of the last batch (ret.Sol of algorithm.jl:116): number of pieces of player `player` for
 * instance b (-1: none), the id of piece k, and a piece's rows (A: m x nv row-major; rl / ru: 1 = strict). */
int qpn_net_sol_count(qpn_net *net, int b, int player);
int qpn_net_sol_piece(qpn_net *net, int b, int player, int k);
int qpn_net_piece_rows(qpn_net *net, int piece);
int qpn_net_piece_get(qpn_net *net, int piece, double *A, double *l, double *u, uint8_t *rl, uint8_t *ru);
/* out[0..15] = {kernel launches, rounds, requests, batched calls, LPs solved, pieces, nodes, level GAVIs,
 * collect misses, combine misses, host ns in the instance logic, host ns in / waiting for the numeric backend (both
 * summed over the host threads), cohort splits, host ns applying the answers of a round, LPs of emptiness tests,
 * batched LP calls} since the net was created. */
int qpn_net_stats(qpn_net *net, int64_t *out);

#ifdef __cplusplus
}
/* layout of the structs above on the LP64 targets this library is built for (checked again by the bindings at load time) */
static_assert(sizeof(qpn_matrix) == 48, "qpn_matrix layout");
static_assert(sizeof(qpn_gavi) == 88, "qpn_gavi layout");
static_assert(sizeof(qpn_node) == 64, "qpn_node layout");
static_assert(sizeof(qpn_level) == 144, "qpn_level layout");
static_assert(sizeof(qpn_net_desc) == 160, "qpn_net_desc layout");
#endif
#endif /* QPN_CUDA_H */
