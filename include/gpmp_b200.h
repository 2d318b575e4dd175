/* gpmp_b200 C-ABI: the drop-in boundary for GPmp's exact-GP inner loop on B200 (sm_100a).
 *
 * The reference (GPmp 0.9.37) has no FFI: its seam is the Python-level `gpmp.num` backend namespace
 * plus `gpmp.kernel` / `gpmp.core.Model`.  These entry points are what a CUDA backend
 * module for that seam binds (ctypes stub in INTEGRATION.md); each one names the reference interface it
 * replaces (paths relative to the reference root).
 *
 * Conventions
 *  - fp64 only (the reference rejects float32, gpmp/config.py:59-78).  All matrices row-major with an
 *    explicit leading dimension (elements).  "lower" matrices define entries j <= i only.
 *  - Pointers named *_dev are DEVICE pointers owned by the caller (torch tensors); covariance
 *    parameters travel BY VALUE in a host struct (they arrive from SciPy as a host vector,
 *    SURVEY.md B.5), so one evaluation needs no host->device copy at all.
 *  - No allocation, no exceptions, no implicit synchronisation: every call only enqueues work on the
 *    given stream (a cudaStream_t passed as void*).  Workspace is caller-provided; sizes come from the
 *    *_workspace_bytes queries.  Numerical failure (non-positive pivot k, 1-based) is written to a
 *    device `info` word, LAPACK style; the host wrapper maps it to torch.linalg.LinAlgError / +inf
 *    (reference convention: gpmp/core/likelihood.py:45-48,121-124).
 *  - Return value: 0 = enqueued; negative = rejected (GPMP_ERR_*).
 *  - Re-entrant per (stream, workspace): the library-owned look-ahead streams / events exist once per
 *    (device, caller stream), and the only host-side state shared between calls -- their registry and the
 *    record of which likelihood workspaces already hold a leading block of T = L^-1 (see gpmp_lik_value) --
 *    is mutex-protected.  (All reference callers are single-threaded loops:
 *    kernel/parameter_selection.py:253, mcmc/param_posterior.py:752.)
 */
#ifndef GPMP_B200_H
#define GPMP_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GPMP_OK 0
#define GPMP_ERR_ARG (-1)
#define GPMP_ERR_DIM (-2)        /* d > GPMP_MAX_DIM or p > GPMP_MAX_P */
#define GPMP_ERR_WORKSPACE (-3)  /* workspace too small */
#define GPMP_ERR_CUDA (-100)
#define GPMP_ERR_ALIGN (-101)    /* pointer / leading dimension not 16-byte compatible */

#define GPMP_MAX_DIM 32
#define GPMP_MAX_P 16
#define GPMP_MAX_Q 31 /* columns of the linear-predictor basis P */

/* Covariance specification: sigma2 * Matern_{p+1/2}(|| diag(exp(loginvrho)) (x - y) ||) plus, on the
 * same-set diagonal, either the nugget 10*sigma2*eps (gpmp/kernel/matern.py:88-94) or a noise variance
 * exp(log_tau2) (examples/gpmp_example07_nd_regression.py:95-111).  Parameter order of the gradient:
 * [d/d log_sigma2, (d/d log_tau2 if noise), d/d loginvrho_0..d-1]. */
typedef struct gpmp_cov_spec {
    int p;     /* Matern regularity nu = p + 1/2, 0 <= p <= GPMP_MAX_P */
    int d;     /* input dimension, 1 <= d <= GPMP_MAX_DIM */
    int noise; /* 0: nugget convention of maternp_covariance; 1: + exp(log_tau2) * I */
    int reserved;
    double log_sigma2;
    double log_tau2;
    double loginvrho[GPMP_MAX_DIM];
} gpmp_cov_spec;

/* ---- library / measurement ------------------------------------------------------------------ */
int gpmp_abi_version(void);
/* Kernels launched by this library since load (all classes). */
unsigned long long gpmp_launch_count(void);
/* Per-class CUDA-event profiling (classes: 0 covariance build, 1 DMMA GEMM, 2 panel factor,
 * 3 dK contraction, 4 small/reduction, 5 batched criterion).  enable!=0 brackets each launch with
 * events on its stream; gpmp_prof_read synchronises those events and returns accumulated
 * milliseconds, launches and algorithmic work (flops or bytes), then clears the class.  enable == 2
 * additionally issues the look-ahead streams' launches on the caller's stream (same kernels, serialised), so
 * the class times are exclusive kernel times. */
int gpmp_prof_enable(int enable);
int gpmp_prof_read(int kernel_class, double* ms, unsigned long long* launches, double* work);
/* FP64 tensor-pipe probe for the roofline denominator (MEASURED_PEAKS.json carries no
 * FP64 entry): `ctas` CTAs of 16 warps issue `iters` rounds of 8 independent DMMA.8x8x4 from registers;
 * *flops_out receives the flops enqueued (2*8*8*4 per DMMA).  The caller times the launch with CUDA events. */
int gpmp_measure_dmma_peak(int ctas, int iters, double* sink_dev, double* flops_out, void* stream);

/* ---- L0/L1: distances and Matern covariance --------------------------------------------------
 * replaces gnp.scaled_distance (gpmp/num/torch_backend.py:810-820, numpy_backend.py:432-436):
 * D[n x m] = || exp(loginvrho) * (x_i - y_k) ||, direct differences (the NumPy/SciPy-cdist form).
 * y_dev == NULL or y_dev == x_dev means the same set (exact zero diagonal). */
int gpmp_scaled_distance(const double* loginvrho_host, int d, const double* x_dev, int n,
                         const double* y_dev, int m, double* D_dev, long long ldd, void* stream);
/* replaces gnp.scaled_distance_elementwise (torch_backend.py:823-829): out[n]. */
int gpmp_scaled_distance_elementwise(const double* loginvrho_host, int d, const double* x_dev,
                                     const double* y_dev, int n, double* out_dev, void* stream);
/* replaces gp.kernel.maternp_kernel (gpmp/kernel/matern.py:32-64), elementwise on count values;
 * dk_dev (optional) receives dk/dh for the autograd backward. */
int gpmp_maternp_kernel(int p, const double* h_dev, double* k_dev, double* dk_dev, long long count,
                        void* stream);

#define GPMP_COV_FULL 0  /* all n x m entries; same set: lower tiles computed, mirrored, diag added */
#define GPMP_COV_LOWER 1 /* same set only: entries j <= i written, strict upper left untouched */
/* replaces gp.kernel.maternp_covariance, pairwise=False (gpmp/kernel/matern.py:67-141).
 * y_dev == NULL: same set (diagonal term added); otherwise cross-covariance (no diagonal term). */
int gpmp_matern_cov(const gpmp_cov_spec* spec, const double* x_dev, int n, const double* y_dev, int m,
                    double* K_dev, long long ldk, int mode, void* stream);
/* pairwise=True branch (matern.py:91-92,116-117): out[n]; y_dev == NULL gives sigma2 * ones. */
int gpmp_matern_cov_pairwise(const gpmp_cov_spec* spec, const double* x_dev, const double* y_dev, int n,
                             double* out_dev, void* stream);
/* Backward of the covariance op for a dense upstream gradient G (n x m):
 * grad[j] = sum_ik G_ik dK_ik/dtheta_j, dK regenerated tile by tile (never materialised).
 * grad_dev has 1 + noise + d entries; partial_dev is scratch from gpmp_contract_workspace_bytes. */
size_t gpmp_contract_workspace_bytes(int n, int m, int d);
int gpmp_matern_cov_backward(const gpmp_cov_spec* spec, const double* x_dev, int n, const double* y_dev,
                             int m, const double* G_dev, long long ldg, double* grad_dev,
                             void* partial_dev, size_t partial_bytes, void* stream);

/* Backward of gnp.scaled_distance for a dense upstream gradient G (n x m): grad_dev[1 + j] =
 * sum_ik G_ik dD_ik/dloginvrho_j (grad_dev[0] is not used); zero contribution where D_ik == 0, like the
 * reference's custom_sqrt (torch_backend.py:783-788). */
int gpmp_scaled_distance_backward(const double* loginvrho_host, int d, const double* x_dev, int n,
                                  const double* y_dev, int m, const double* G_dev, long long ldg,
                                  double* grad_dev, void* partial_dev, size_t partial_bytes, void* stream);

/* ---- L0: Cholesky family (replaces gnp.cholesky / cholesky_solve / cholesky_inv /
 * solve_triangular / cho_factor / cho_solve: gpmp/num/torch_backend.py:111,841-890;
 * gpmp/num/numpy_backend.py:458-469) -------------------------------------------------------------
 * gpmp_potrf: blocked right-looking Cholesky of the leading n x n LOWER matrix of A (row-major, in
 * place).  On exit the lower tiles hold L, the diagonal 128-tiles have an explicitly zero upper part
 * and the strictly-upper 128-tiles hold the mirrored L^T tiles (so both L and L^T can be streamed
 * row-wise).  Rows n..nrows-1 of A (nrows >= n, first n columns) are carried along in every panel
 * solve, so on exit they hold  B L^-T : each extra row is a right-hand side whitened by L (this is how
 * z and the mean basis P are whitened for free).  work_dev (gpmp_potrf_workspace_bytes) receives the
 * inverses of the diagonal blocks, which gpmp_potri / gpmp_trsm_rows start from.
 * info_dev: 0, or the 1-based index of the first non-positive pivot (must be zeroed by the caller). */
size_t gpmp_potrf_workspace_bytes(int n, int nrows);
int gpmp_potrf(double* A_dev, int n, int nrows, long long lda, void* work_dev, size_t work_bytes,
               int* info_dev, void* stream);
/* gpmp_potri: from the factor and the gpmp_potrf workspace, T = L^-1 by block doubling
 * (Tlo_dev: T, lower; Tup_dev: T^T, upper) and Kinv_dev = lower triangle of K^-1 = T^T T.
 * All three are n x ld, caller-provided; Kinv doubles as scratch while T is being assembled. */
int gpmp_potri(const double* L_dev, int n, long long ldl, const void* potrf_work_dev, double* Tlo_dev,
               double* Tup_dev, double* Kinv_dev, long long ld, void* stream);
/* Triangular solves with many right-hand sides stored as ROWS of Bt (m x n):
 * trans == 0:  Bt <- Bt L^-T  (each row b becomes L^-1 b);  trans == 1:  Bt <- Bt L^-1  (L^-T b).
 * scratch_dev: m x 512 doubles. */
int gpmp_trsm_rows(const double* L_dev, int n, long long ldl, const void* potrf_work_dev, double* Bt_dev,
                   int m, long long ldb, int trans, void* scratch_dev, void* stream);
/* C[M x N] = alpha * A[M x K] * B[N x K]^T + beta * C on the FP64 tensor pipe (DMMA); row-major, both
 * operands K-contiguous.  tri: 0 none; 1/2: A is upper/lower triangular on the tile grid (k >= row /
 * k <= row); 3/4: same for B.  lower != 0 computes only the 128-tiles on or below the diagonal.
 * All leading dimensions even, pointers 16-byte aligned. */
int gpmp_gemm_nt(const double* A_dev, long long lda, const double* B_dev, long long ldb, double* C_dev,
                 long long ldc, int M, int N, int K, double alpha, double beta, int tri, int lower,
                 void* stream);
/* out[c][r] = in[r][c] */
int gpmp_transpose(const double* in_dev, long long ldi, double* out_dev, long long ldo, int rows,
                   int cols, void* stream);

/* ---- L2: Gaussian likelihoods (replaces Model.negative_log_likelihood_zero_mean /
 * negative_log_likelihood / negative_log_restricted_likelihood and their covparam gradients:
 * gpmp/core/likelihood.py:18-129, gpmp/core/linalg.py:49-88, autograd of
 * gpmp/num/torch_backend.py:547-604) -----------------------------------------------------------
 * One caller-owned workspace holds everything an evaluation produces (factor L, whitened rows, the
 * small R factor and, for the gradient, T and K^-1), so value() and grad() are separate enqueues and
 * grad() may be skipped (SLSQP line searches call the value more often than the gradient).
 *
 * gpmp_lik_value:  REML (q >= 1, P_dev = n x q row-major basis) or zero-mean ML (q == 0):
 *   out_dev[0] = criterion value  0.5((n-q) log 2pi + logdet + quad)   (+inf if not positive definite)
 *   out_dev[1] = logdet,  out_dev[2] = quadratic form,  out_dev[3] = 2 sum log diag(L),
 *   out_dev[4] = 2 sum log diag(R~),  out_dev[5] = 2 sum log diag(R0),  out_dev[6] = (double) info
 *   (out_dev holds 8 doubles)
 *   spec != NULL: K is built from (spec, x_dev) by the fused Matern kernel (lower triangle only);
 *   spec == NULL: K_dev (n x n, lower triangle read) is a user-composed covariance.
 *   With a workspace sized for the gradient (want_grad = 1) and n a multiple of the column-group width with
 *   8..32 groups, the call also computes the leading block of T = L^-1 (the first half of the columns,
 *   rounded down to a power of two of groups) under the latency-bound tail of the factorisation, on a
 *   low-priority library stream joined before the call's last kernel; gpmp_lik_grad / gpmp_lik_loo on the same
 *   workspace skip that block.  Values are unaffected.
 * gpmp_lik_grad (after gpmp_lik_value on the same workspace, which must have been sized with
 * want_grad = 1):
 *   grad_dev[0 .. 1+noise+d)  = d value / d covparam            (spec != NULL)
 *   dz_dev[n]    (optional)   = d value / d z  (= alpha; feeds the meanparam gradient of
 *                               negative_log_likelihood, likelihood.py:87-89)
 *   dK_dev       (optional)   = d value / d K, symmetric n x n  (the composable path) */
size_t gpmp_lik_workspace_bytes(int n, int q, int d, int want_grad);
int gpmp_lik_value(const gpmp_cov_spec* spec, const double* K_dev, long long ldk, const double* x_dev,
                   int n, const double* z_dev, const double* P_dev, int q, void* work_dev,
                   size_t work_bytes, double* out_dev, int* info_dev, void* stream);
int gpmp_lik_grad(const gpmp_cov_spec* spec, const double* x_dev, int n, int q, void* work_dev,
                  size_t work_bytes, double* grad_dev, double* dz_dev, double* dK_dev, long long lddk,
                  void* stream);

/* The same evaluation with the factorisation partitioned over the GPUs of one node (one process per GPU,
 * SURVEY.md 8e "one large Cholesky"): column groups of gpmp_lik_dist_block(n) columns, group g owned by rank
 * g mod G.  Every rank holds a full workspace (gpmp_lik_workspace_bytes) but keeps only its own groups up to
 * date.  Per group, in order:  owner: gpmp_lik_dist_group (factor the group, fill panel_dev with the group's
 * block column of L: (n+q+1-k0) x block doubles, ld = block);  caller: broadcast panel_dev (ncclBroadcast /
 * torch.distributed.broadcast);  non-owners: gpmp_lik_dist_store (file the panel into their L);  every rank:
 * gpmp_lik_dist_update for each later group it owns (columns [col0, col1)).  After the last group
 * gpmp_lik_dist_finish leaves on every rank exactly the state gpmp_lik_value leaves (out_dev as there), so
 * gpmp_predict_chunk / gpmp_lik_grad / gpmp_lik_loo can follow. */
int gpmp_lik_dist_block(int n);
int gpmp_lik_dist_prepare(const gpmp_cov_spec* spec, const double* K_dev, long long ldk, const double* x_dev,
                          int n, const double* z_dev, const double* P_dev, int q, void* work_dev,
                          size_t work_bytes, int* info_dev, void* stream);
int gpmp_lik_dist_group(int n, int q, void* work_dev, size_t work_bytes, int k0, double* panel_dev,
                        int* info_dev, void* stream);
int gpmp_lik_dist_store(int n, int q, void* work_dev, size_t work_bytes, int k0, const double* panel_dev,
                        void* stream);
int gpmp_lik_dist_update(int n, int q, void* work_dev, size_t work_bytes, int k0, const double* panel_dev,
                         int col0, int col1, void* stream);
int gpmp_lik_dist_finish(int n, int q, void* work_dev, size_t work_bytes, double* out_dev, int* info_dev,
                         void* stream);

/* Gradient of the partitioned evaluation ("gradient contraction K4 -- M rows partitioned,
 * all-reduce of d+1 doubles").  After gpmp_lik_dist_finish on a workspace sized with want_grad = 1 every rank
 * holds the full factor; the 2n^3/3 of the gradient is split by blocks of gpmp_lik_dist_block(n) rows:
 *   gpmp_lik_dist_tup_rows      rows [row0, row0+rows) of T^T = L^-T (unit right-hand sides, forward solve
 *                               started at their block);  the caller broadcasts those rows to every rank;
 *   gpmp_lik_dist_kinv_rows     the same rows of K^-1 (columns 0 .. row0+rows) = T^T[rows] (T^T)^T;
 *   gpmp_lik_dist_u_cols        columns [row0, row0+rows) of U = [Q~^T; r^T] T;  the caller sums U over ranks;
 *   gpmp_lik_dist_contract_rows 0.5 sum over the lower tiles of those rows of M dK/dtheta -> grad_dev[1+noise+d];
 *                               the caller sums over its blocks and all-reduces.
 * gpmp_lik_ws_offset / gpmp_lik_ws_ld locate the matrices inside the workspace (0 A, 1 T^T, 2 K^-1, 3 U, 4 T)
 * so the caller can hand row blocks to its collective library without copies. */
size_t gpmp_lik_ws_offset(int n, int q, int which);
long long gpmp_lik_ws_ld(int n);
int gpmp_lik_dist_tup_rows(int n, int q, void* work_dev, size_t work_bytes, int row0, int rows, void* stream);
int gpmp_lik_dist_kinv_rows(int n, int q, void* work_dev, size_t work_bytes, int row0, int rows, void* stream);
int gpmp_lik_dist_u_cols(int n, int q, void* work_dev, size_t work_bytes, int row0, int rows, void* stream);
int gpmp_lik_dist_contract_rows(const gpmp_cov_spec* spec, const double* x_dev, int n, int q, void* work_dev,
                                size_t work_bytes, int row0, int rows, double* grad_dev, void* stream);

/* Leave-one-out predictions by virtual cross-validation (replaces gpmp/core/loo.py:65-130 behind Model.loo,
 * gpmp/core/model.py:309-343) after gpmp_lik_value on a workspace sized with want_grad = 1 and the (centred)
 * data z_dev that was whitened:  eloo_i = (Pi z)_i / Pi_ii,  s2loo_i = 1 / Pi_ii,  zloo_i = z_i - eloo_i, with
 * Pi = K^-1 - K^-1 P (P^T K^-1 P)^-1 P^T K^-1 (= K^-1 when q == 0). */
int gpmp_lik_loo(int n, int q, void* work_dev, size_t work_bytes, const double* z_dev, double* zloo_dev,
                 double* s2loo_dev, double* eloo_dev, void* stream);

/* ---- L2: kriging prediction (replaces gpmp/core/kriging.py:35-116,170-199 behind Model.predict,
 * gpmp/core/model.py:227-307) for one chunk of m test points, after gpmp_lik_value on work_dev (the
 * fitted state: L, whitened data, Q~, R~).
 *   Vt_dev (m x ldv, ldv >= n): spec != NULL -> filled here with K(xt, xi); spec == NULL -> the caller
 *     has put a user-composed K(xt, xi) there.  On exit: rows lambda_t^T if want_lambda, else scratch.
 *   Pt_dev (m x q) mean basis at xt (NULL when q == 0);  ktt_dev[m] prior variances (NULL: sigma2).
 *   mean_dev[m], var_dev[m] outputs (variance not clamped: Model.predict clamps and warns).
 *   want_lambda: 0 Vt holds the whitened cross-covariances v_t = L^-1 k(xi, x_t) on exit; 1 rows lambda_t^T;
 *   2 rows w_t = v_t - Q~ e_t, the kriging weights in whitened coordinates (lambda_t = L^-T w_t): conditioning
 *   of sample paths needs only W (L^-1 delta)^T, so the n x m weight matrix of
 *   gpmp/core/sample_paths.py:66-182 is never formed.
 *   scratch_dev: m x NB doubles of solve scratch (NB = gpmp_lik_dist_block(n)), then, 256-byte aligned, the
 *   per-point record (m x (q + 2)): e_t = Q~^T v_t - R~^-T p_t (q entries), v_t . r, |v_t|^2.  The full posterior
 *   covariance (return_type = 1 of gpmp/core/kriging.py:170-199) is K_tt - V V^T + E E^T from Vt (mode 0) and
 *   the e_t of that record. */
size_t gpmp_predict_scratch_bytes(int n, int q, int m);
int gpmp_predict_chunk(const gpmp_cov_spec* spec, const double* x_dev, int n, int q, void* work_dev,
                       size_t work_bytes, const double* xt_dev, int m, const double* Pt_dev,
                       const double* ktt_dev, double* Vt_dev, long long ldv, void* scratch_dev,
                       size_t scratch_bytes, double* mean_dev, double* var_dev, int want_lambda,
                       void* stream);

/* Row solves against the factor held in a likelihood workspace (after gpmp_lik_value / gpmp_lik_dist_finish):
 * rows of Bt (m x n): trans == 0  b -> L^-1 b,  trans == 1  b -> L^-T b.  scratch_dev: m x 512 doubles. */
int gpmp_lik_trsm_rows(int n, int q, void* work_dev, size_t work_bytes, double* Bt_dev, int m, long long ldb,
                       int trans, void* scratch_dev, void* stream);

/* ---- batched criterion (replaces the serial loop of gpmp/mcmc/param_posterior.py:739-759): N
 * independent REML / ML values at theta_1..theta_N on fixed (x, z, P).  theta_dev: N x (1+noise+d)
 * rows [log sigma2, (log tau2,) loginvrho...].  values_dev[N] (+inf where not positive definite),
 * info_dev[N].  The workspace decides how many particles are in flight at once:
 * work_bytes >= gpmp_criterion_batched_bytes(n, q, 1), ideally (n, q, N). */
size_t gpmp_criterion_batched_bytes(int n, int q, int nbatch);
int gpmp_criterion_batched(const gpmp_cov_spec* spec, const double* theta_dev, int N, const double* x_dev,
                           int n, const double* z_dev, const double* P_dev, int q, void* work_dev,
                           size_t work_bytes, double* values_dev, int* info_dev, void* stream);

/* ---- batched criterion WITH gradients: N independent values and d value / d theta rows in one call.
 * Replaces the per-particle value-and-gradient loop of gpmp/mcmc/svgd.py:310-313 (gnp.value_and_grad per
 * particle) and the per-batch loop of BatchDifferentiableSelectionCriterion.evaluate_pre_grad
 * (gpmp/num/torch_backend.py:686-712).  Entry b uses theta_b (row b of theta_dev), the points
 * x_dev + b * x_stride and the observations z_dev + b * z_stride (strides in doubles; 0 = all entries share
 * the data: particles; n*d / n = consecutive mini-batches of n points); the mean basis P_dev (n x q, may be
 * NULL for q = 0) is shared.  values_dev[N], grads_dev[N x (1+noise+d)], info_dev[N].  Entries whose matrix is
 * not positive definite get value +inf and an undefined gradient row (the caller zeroes it, like
 * gpmp/num/torch_backend.py:528-529).  work_bytes >= gpmp_criterion_batched_grad_bytes(n, q, d, 1). */
size_t gpmp_criterion_batched_grad_bytes(int n, int q, int d, int nbatch);
int gpmp_criterion_batched_grad(const gpmp_cov_spec* spec, const double* theta_dev, int N, const double* x_dev,
                                long long x_stride, int n, const double* z_dev, long long z_stride,
                                const double* P_dev, int q, void* work_dev, size_t work_bytes,
                                double* values_dev, double* grads_dev, int* info_dev, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GPMP_B200_H */
