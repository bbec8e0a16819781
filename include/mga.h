/*
 * mga.h — C ABI of the B200-native Mixed-Graph-ADMM solver hot path (libmga.so).
 *
 * The reference (JiQi-da/Mixed-Graph-ADMM) has no FFI: its boundary is the Python class
 * ADMM_algorithm (ADMM.py:11-648).  Every entry point below replaces one group of methods of
 * that class and cites them; mixed_graph_admm_b200/ADMM.py is the host-side mirror that binds
 * these symbols through ctypes and keeps the reference's names and signatures.
 *
 * Conventions
 *  - plain pointers and sizes only; no torch / C++ types.
 *  - signals are (B, T, N, 1) row-major, node index fastest (ADMM.py:140, 333-335).
 *  - "device pointer" = memory of the plan's CUDA device, owned by the caller.
 *  - `stream` is a cudaStream_t passed as void*; every call is asynchronous on it unless
 *    it says otherwise.  NULL = the legacy default stream.
 *  - return value: 0 = MGA_OK, otherwise a negative mga_status; mga_last_error() gives the text
 *    (thread-local).  No exception crosses the ABI.
 *  - there is NO CPU fallback: every compute entry point needs a CUDA device and fails with
 *    MGA_ERR_CUDA when none is usable.
 */
#ifndef MGA_H_
#define MGA_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MGA_VERSION 110 /* 0.1.1: mga_admm_solve_host returns the CG coefficients */

typedef enum {
  MGA_OK = 0,
  MGA_ERR_INVALID = -1,      /* bad argument (shape, enum, NULL) */
  MGA_ERR_INDEX = -2,        /* neighbour index outside [-1, N): ADMM.py:204-206 "Index out of bounds" */
  MGA_ERR_CUDA = -3,         /* CUDA runtime error; text in mga_last_error() */
  MGA_ERR_UNSUPPORTED = -4,  /* valid in the reference but not built here */
  MGA_ERR_NONFINITE = -5     /* NaN/Inf met where the reference asserts (ADMM.py:560-606) */
} mga_status;

typedef enum { MGA_F32 = 0, MGA_F64 = 1 } mga_dtype;

/* how apply_op_Ldr_T forms the "father" sum (ADMM.py:196-215) */
typedef enum {
  MGA_LDRT_SCATTER = 0, /* use_kNN=True: scatter_add over the kNN table == gather over its transpose */
  MGA_LDRT_GATHER = 1   /* use_kNN=False: gather with the forward table and row-i weights */
} mga_ldrt_mode;

/* the temporal (directed) graph variant (ADMM.py:37-52, 150-194) */
typedef enum {
  MGA_TEMPORAL_GRAPH = 0, /* kNN / physical table with weights d_ew */
  MGA_TEMPORAL_LINE = 1,  /* use_line_graph=True, skip_connection == 1: first difference in time */
  MGA_TEMPORAL_BAND = 2   /* use_line_graph=True, skip_connection > 1: (T, skip, N) banded stencil */
} mga_temporal_kind;

typedef enum { MGA_ABL_NONE = 0, MGA_ABL_DGTV = 1, MGA_ABL_DGLR = 2, MGA_ABL_UT = 3 } mga_ablation;

/* operators of mga_apply — ADMM.py:138-228, 371-399 */
typedef enum {
  MGA_OP_LU = 0,     /* apply_op_Lu      ADMM.py:138-148 */
  MGA_OP_LDR = 1,    /* apply_op_Ldr     ADMM.py:150-177 */
  MGA_OP_LDRT = 2,   /* apply_op_Ldr_T   ADMM.py:179-223 (incl. quirk Q1 on row t=0) */
  MGA_OP_CLDR = 3,   /* apply_op_cLdr    ADMM.py:225-228 */
  MGA_OP_LHS_X = 4,  /* LHS_x            ADMM.py:371-387 */
  MGA_OP_LHS_ZU = 5, /* LHS_zu           ADMM.py:389-390 */
  MGA_OP_LHS_ZD = 6  /* LHS_zd           ADMM.py:392-399 */
} mga_op;

typedef enum { MGA_SYS_X = 0, MGA_SYS_ZU = 1, MGA_SYS_ZD = 2 } mga_system;

/* which implementation mga_admm_solve uses */
typedef enum {
  MGA_MODE_AUTO = 0,     /* resident when eligible (fp32, fixed iteration counts, ablation None, time-invariant weights,
                            window fits one CTA; forecasting, mask mode and every temporal-graph variant), else streaming */
  MGA_MODE_STREAMING = 1,/* state in HBM/L2, one fused kernel per CG phase */
  MGA_MODE_RESIDENT = 2, /* one CTA per window, CG vectors in registers, gathered vectors in SMEM */
  MGA_MODE_STREAMING_POINT = 3 /* streaming, always the general one-thread-per-lattice-point kernels (every dtype,
                                  mask, ablation, temporal variant); MGA_MODE_STREAMING picks the node-major
                                  kernels (time-tiled shared-memory CG kernels for graphs or node tiles that fit a
                                  CTA, L1-gather kernels otherwise) when the call is fp32, forecasting,
                                  fixed-iteration, ablation None */
} mga_mode;

/* The graph tensors ADMM_algorithm.__init__ leaves behind (ADMM.py:25-52).  HOST pointers; the
 * plan copies what it needs.  nbr_* are the reference's connect_list slices: int64, -1 = no
 * neighbour (quirk Q6).  Weight tables may be time-expanded or not (quirk Q8):
 *   u_w: (u_w_T, N, ku) with u_w_T in {1, T};  d_w: (d_w_T, N, kd) with d_w_T in {1, T-1}.
 * For MGA_TEMPORAL_LINE nbr_d / d_w are ignored.  For MGA_TEMPORAL_BAND d_w is the reference's
 * (T, skip, N) table and kd = skip. */
typedef struct {
  int32_t n_nodes, T, t_in;
  int32_t ku;           const int64_t* nbr_u; const float* u_w; int32_t u_w_T;
  int32_t kd;           const int64_t* nbr_d; const float* d_w; int32_t d_w_T;
  int32_t ldrt_mode;    /* mga_ldrt_mode */
  int32_t temporal;     /* mga_temporal_kind */
} mga_graph_desc;

/* ADMM_info + ablation (ADMM.py:59-64, 31) */
typedef struct {
  double rho, rho_u, rho_d, mu_u, mu_d1, mu_d2;
  int32_t ablation;     /* mga_ablation */
  int32_t reserved;
} mga_params;

typedef struct mga_plan mga_plan;

/* Diagnostics, one row of MGA_DIAG_COLS doubles per outer iteration: batch-wide SUMS (the
 * host mirror takes sqrt / divides by B).  ADMM.py:609-637. */
enum {
  MGA_DIAG_DX2 = 0,      /* sum (x - x_old)^2            -> x_shift_list      :612 */
  MGA_DIAG_X_ZU2 = 1,    /* sum (x - zu)^2               -> primal[zu]        :616 */
  MGA_DIAG_DZU2 = 2,     /* sum (zu - zu_old)^2          -> dual[zu]          :618 */
  MGA_DIAG_GLR = 3,      /* sum x . Lu x                 -> GLR_list (/B)     :619 */
  MGA_DIAG_RECOVER2 = 4, /* sum (Hx - y)^2               -> recover_list      :625 */
  MGA_DIAG_PHI_LDX2 = 5, /* sum (phi - Ldr x)^2          -> primal[phi]       :628 */
  MGA_DIAG_DPHI2 = 6,    /* sum (phi - phi_old)^2        -> dual[phi]         :630 */
  MGA_DIAG_DGTV = 7,     /* sum |Ldr x|                  -> DGTV_list (/B)    :631 */
  MGA_DIAG_X_ZD2 = 8,    /* sum (x - zd)^2               -> primal[zd]        :634 */
  MGA_DIAG_DZD2 = 9,     /* sum (zd - zd_old)^2          -> dual[zd]          :636 */
  MGA_DIAG_DGLR = 10,    /* sum (Ldr x)^2                -> DGLR_list (/B)    :637 */
  MGA_DIAG_NONFINITE = 11,/* count of non-finite entries met in x, zu, zd, phi, gamma (asserts :575-606) */
  MGA_DIAG_COLS = 12
};

/* Optional outputs of mga_admm_solve; any pointer may be NULL.  Device pointers except where
 * noted.  Shapes: signals (B,T,N); diag (n_outer, MGA_DIAG_COLS) double; dx_sum (n_outer,T,N)
 * double = sum over the batch of (x - x_old) (the host forms mean -> norm, ADMM.py:614);
 * alpha/beta (n_outer, 3, max_cg_iter, B) in the signal dtype, system order x, zu, zd
 * (ADMM.py:572-591); cg_iters HOST int32 (n_outer, 3): iterations used, -1 = not converged
 * (ADMM.py:362, 368); outer_done HOST int32: outer iterations executed (ADMM.py:645-646). */
typedef struct {
  void *zu, *zd, *phi, *gamma, *gamma_u, *gamma_d;
  double* diag;
  double* dx_sum;
  void *alpha, *beta;
  int32_t* cg_iters;
  int32_t* outer_done;
} mga_admm_outputs;

/* ---- plan: replaces the graph state of ADMM_algorithm.__init__ (ADMM.py:15-98) and the index
 * validation of ADMM.py:204-206.  Builds int32 ELL tables, the transposed CSR for L_d^T (entries
 * in scatter order), detects time-invariant weights, uploads to `device`.  Synchronous. */
int mga_plan_create(const mga_graph_desc* desc, int device, mga_plan** out);
void mga_plan_destroy(mga_plan* plan);
/* 1 if mga_admm_solve(MODE_AUTO) would run the resident kernel for this dtype / batch. */
int mga_plan_resident_eligible(const mga_plan* plan, int dtype);
/* multiprocessor count / bytes of dynamic SMEM the resident kernel uses (0 if not eligible) */
int mga_plan_info(const mga_plan* plan, int32_t* sm_count, int32_t* resident_smem_bytes,
                  int32_t* resident_threads, int32_t* max_in_degree);

/* ---- operators: y = op(x).  apply_op_* / LHS_* (ADMM.py:138-228, 371-399).  `mask` (device,
 * same shape/dtype as x, may be NULL) is LHS_x's H when given (ADMM.py:375-376).  x != y. */
int mga_apply(mga_plan* plan, int op, const mga_params* prm, const void* x, void* y,
              const void* mask, int64_t B, int dtype, void* stream);

/* ---- CG_solver (ADMM.py:329-368) for one of the three systems.  x_inout holds x0 on entry.
 * tol <= 0 runs exactly max_iter iterations with no host sync ("unrolled"); tol > 0 applies
 * the batch-global test max_b sqrt(r.r) < tol after every iteration (quirk Q3; one host sync
 * per iteration) and returns the count in *iters_out (HOST; -1 = not converged).
 * alpha_out / beta_out: device (max_iter, B) or NULL.  `mask_first` (device or NULL) goes to
 * the initial residual only (quirk Q4). */
int mga_cg_solve(mga_plan* plan, int system, const mga_params* prm, const void* rhs, void* x_inout,
                 const void* mask_first, int64_t B, int dtype, int max_iter, double tol,
                 int32_t* iters_out, void* alpha_out, void* beta_out, void* stream);

/* Which implementation mga_cg_solve uses (mga_mode; default MGA_MODE_AUTO): with a fixed iteration count
 * (tol <= 0), no mask and a resident-eligible plan the whole solve runs in ONE launch, one window per CTA,
 * iterates in registers / shared memory (rhs and x0 read once, x written once); otherwise one fused kernel
 * per CG phase over vectors in HBM.  MGA_MODE_STREAMING forces the latter (benchmarks of the HBM path). */
int mga_plan_set_cg_mode(mga_plan* plan, int mode);

/* ---- initial_guess (ADMM.py:766-781): y (B,t_in,N) -> x (B,T,N).  t_mean / t_var are the
 * float32-rounded mean(t) and mean(t^2)-mean(t)^2 the reference computes on the host. */
int mga_initial_guess(mga_plan* plan, const void* y, void* x, int64_t B, int dtype,
                      double t_mean, double t_var, void* stream);

/* ---- fused elementwise steps of combined_loop */
/* RHS_x (ADMM.py:552-559): rhs = Ldr_T(gamma + rho phi)/2 + (rho_u zu + rho_d zd)/2
 *                               - (gamma_u + gamma_d)/2 + H^T y.  y has y_rows time rows. */
int mga_rhs_x(mga_plan* plan, const mga_params* prm, const void* gamma, const void* phi,
              const void* zu, const void* zd, const void* gamma_u, const void* gamma_d,
              const void* y, int y_rows, void* rhs, int64_t B, int dtype, void* stream);
/* gamma_z += rho_z (x - z)  (ADMM.py:595-597) */
int mga_dual_ascent(mga_plan* plan, double rho_z, const void* x, const void* z, void* gamma_z,
                    int64_t B, int dtype, void* stream);
/* phi_direct + gamma ascent (ADMM.py:401-408, 603-605): phi = soft_{mu_d1/rho}(Ldr x - gamma/rho);
 * gamma += rho (phi - Ldr x).  phi_out may alias nothing else; gamma updated in place. */
int mga_prox_phi_dual(mga_plan* plan, const mga_params* prm, const void* x, void* gamma_inout,
                      void* phi_out, int64_t B, int dtype, void* stream);
/* phi_direct alone (ADMM.py:401-408); gamma is read-only */
int mga_phi_direct(mga_plan* plan, const mga_params* prm, const void* x, const void* gamma,
                   void* phi_out, int64_t B, int dtype, void* stream);

/* ---- the whole of combined_loop after argument checking (ADMM.py:528-648): initial guess,
 * n_outer x (RHS, 3 CG solves, dual ascent, phi prox, diagnostics, stop test).
 * y: (B, y_rows, N) with y_rows = t_in (forecast) or T (mask mode, `mask` non-NULL);
 * x_out: (B, T, N).  cg_tol / admm_tol <= 0 disable the stop tests (fixed iteration counts, no
 * host sync inside).  t_mean / t_var as in mga_initial_guess.  mode: mga_mode. */
int mga_admm_solve(mga_plan* plan, const mga_params* prm, const void* y, int y_rows, const void* mask,
                   void* x_out, int64_t B, int dtype, int n_outer, int max_cg_iter, double cg_tol,
                   double admm_tol, double t_mean, double t_var, int want_diag,
                   const mga_admm_outputs* outs, int mode, void* stream);

/* ---- the same loop in "cluster mode": one thread-block cluster per window (time slabs dealt to the CTAs, halo rows
 * exchanged through distributed shared memory), in the signal's own precision, with the reference's stop tests
 * (ADMM.py:360, 645) decided on the device - one launch for a whole tolerance-driven solve.  mga_admm_solve(MODE_AUTO)
 * picks it for tolerance mode at B = 1 (the reference's own call pattern: float64, CG_tol 1e-8, ADMM_tol 1e-6,
 * ADMM.py:76-80), for float64 batches with fixed counts and for plans with per-time-step weight tables (T,N,k) /
 * (T-1,N,K); this entry point runs it on request.  Forecasting or mask mode (pass the mask through mga_admm_solve),
 * ablation None, T <= 32, N <= 1024; cg_tol / admm_tol > 0 need B = 1 (for B > 1 the reference's tests are
 * batch-global).  Arguments as mga_admm_solve. */
int mga_cluster_solve(mga_plan* plan, const mga_params* prm, const void* y, void* x_out, int64_t B, int dtype,
                      int n_outer, int max_cg_iter, double cg_tol, double admm_tol, double t_mean, double t_var,
                      int want_diag, const mga_admm_outputs* outs, void* stream);

/* ---- same, with HOST buffers (the end-to-end call every caller of the reference makes: CPU tensors in, CPU tensor
 * out).  Forecasting mode, fixed iteration counts.  When the plan runs in resident mode the whole batch is ONE
 * persistent launch: y is uploaded chunk by chunk while the kernel already solves the first chunks, finished chunks
 * are downloaded while it solves the rest.  Otherwise the batch is cut into chunks that are uploaded / solved /
 * downloaded on three streams.  diag_host (n_outer, MGA_DIAG_COLS) and dx_sum_host (n_outer, T, N) doubles may be
 * NULL.  alpha_host / beta_host: HOST (n_outer, 3, max_cg_iter, B) in the signal dtype (the lists of ADMM.py:572-591),
 * both or neither.  Synchronous: returns when every host buffer is complete.  Buffers need not be pinned (pinned
 * ones are copied asynchronously).  chunk <= 0: automatic. */
int mga_admm_solve_host(mga_plan* plan, const mga_params* prm, const void* y_host, int y_rows,
                        void* x_host, int64_t B, int dtype, int n_outer, int max_cg_iter,
                        double t_mean, double t_var, int want_diag, double* diag_host,
                        double* dx_sum_host, void* alpha_host, void* beta_host, int mode, int64_t chunk);

/* ---- kNN tables by shortest-path distance (utils.py:183-204), host only, bit-identical to the
 * reference's networkx + heapq result.  edges (E,2) int64, dists (E,) float64.
 * out_nodes (N, k+1) int32 (-1 pad), out_dists (N, k+1) float32 (inf pad). */
int mga_knn_build(int32_t n_nodes, int64_t n_edges, const int64_t* edges, const double* dists,
                  int32_t k, int32_t* out_nodes, float* out_dists);

/* ---- host-only self-check of the resident kernel's plan-time schedule (no GPU needed): verifies
 * that the internal node order, the per-row visit order and the per-warp in-list ELL are a pure
 * re-ordering of `desc` (self links of the temporal table are held apart as one weight per node).
 * stats (6 doubles, may be NULL): average shared-memory wavefronts per quarter-warp gather phase
 * {forward table before, after, in-list before, after} scheduling, then the in-list steps summed
 * over the warps {before, after}. */
int mga_schedule_selfcheck(const mga_graph_desc* desc, double* stats);

/* number of kernels this library has launched in this process (bench.py's gpu_launches) */
int64_t mga_launch_count(void);
const char* mga_last_error(void);
int mga_version(void);

#ifdef __cplusplus
}
#endif
#endif /* MGA_H_ */
