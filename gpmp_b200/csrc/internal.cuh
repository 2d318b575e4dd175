// Internal declarations shared by the translation units of libgpmp_b200 (not part of the C-ABI).
#pragma once
#include "common.cuh"

namespace gpmp {

struct MaternDev {
    int p, d;
    double sigma2, diag_add, c;
    double dscale;               // -c^2/(2p-1) (p>=1) or -c (p==0): k'(h)/h = dscale*exp(-t)*q_{p-1}(t)
    double coef[GPMP_MAX_P];     // a_i, i=0..p-1:   q_p(t)     = 1 + sum a_i (2t)^(p-i)
    double coefm1[GPMP_MAX_P];   // a'_i, i=0..p-2:  q_{p-1}(t) = 1 + sum a'_i (2t)^(p-1-i)
    double bq[GPMP_MAX_P + 1];   // the same polynomials in powers of t = c h (Horner by FMA): q_p(2t) = sum_k bq[k] t^k
    double bqm1[GPMP_MAX_P + 1]; // q_{p-1}(2t) = sum_k bqm1[k] t^k  (k <= p-1)
    double invrho[GPMP_MAX_DIM];
};

enum CovModeInternal { COV_RECT = 0, COV_SYM_FULL = 1, COV_SYM_LOWER = 2 };

// ---- matern.cu
int make_matern_dev(const gpmp_cov_spec* s, MaternDev* m, bool same_set);
int launch_prep_theta(const gpmp_cov_spec* spec, const double* theta_dev, int N, int same_set, MaternDev* out,
                      cudaStream_t stream);
int launch_matern_cov(const gpmp_cov_spec* spec, const MaternDev* mdev, int batch, long long strideK,
                      const double* x, int n, const double* y, int mcols, double* K, long long ldk, int mode,
                      int dist_only, cudaStream_t stream, long long strideX = 0);
size_t contract_workspace_bytes(int n, int mcols, int d);
// batched contraction: entry b uses mdev[b], G + b strideG, Ut + b strideU, x + b strideX (elements) and writes
// grad + b (1 + noise + d); partial must hold batch x tiles x (2 + d) doubles
struct ContractBatch { int batch; const MaternDev* mdev; long long strideG, strideU, strideX; };
int launch_contract(const gpmp_cov_spec* spec, const double* x, int n, const double* y, int mcols, const double* G,
                    long long ldg, const double* Ut, long long ldu, int r, int sym, int dist_only, double half,
                    double* grad, void* partial, size_t partial_bytes, cudaStream_t stream, int tile_row0 = 0,
                    int tile_row1 = -1, const ContractBatch* cb = nullptr);
int launch_pairwise(const gpmp_cov_spec* spec, const double* x, const double* y, int n, double* out, int dist_only,
                    cudaStream_t stream);
int launch_maternp_elementwise(int p, const double* h, double* k, double* dk, long long count, cudaStream_t stream);

// ---- potrf.cu
int potrf_block_size(int n);
// Buffers of T = L^-1 (Tlo lower, Tup = T^T upper, both n x ld) and an n x ld scratch: when given, the factorisation
// also leaves the leading early_prefix(n, NB) x early_prefix(n, NB) block of T complete (computed under its tail).
struct EarlyInverse { double* Tlo; double* Tup; double* X; long long ld; };
int early_prefix(int n, int NB);
int potrf_core(double* A, long long lda, long long strideA, int n, int nrows, int NB, double* Tlo, double* Tup,
               long long strideT, double* W, long long strideW, int* info, long long strideInfo, int batch,
               cudaStream_t stream, double* Tsub = nullptr, long long strideTsub = 0, int tsub_tiles = 0,
               const EarlyInverse* early = nullptr);
// Tsub: block-diagonal (32x32) tile inverses for the substitution solve; tsub_tiles = 128x128 tiles it holds per
// matrix (ceil(n / 128) for the look-ahead path of a single matrix, 1 for the batched value-only path)
int potri_core(const double* L, int n, long long ldl, int NB, const double* Tlo_c, const double* Tup_c,
               double* Tlo, double* Tup, double* Kinv, long long ldk, cudaStream_t stream, int batch = 1,
               long long strideL = 0, long long strideTc = 0, long long strideT = 0, int prefix = 0,
               cudaEvent_t t_done = nullptr);
// t_done: recorded on the stream once T is complete (before the K^-1 product)
// prefix: the leading prefix x prefix block of T is already complete (EarlyInverse)
int trsm_rows_core(const double* A, int n, long long lda, int NB, const double* Tlo_c, const double* Tup_c,
                   double* Bt, int m, long long ldb, int trans, double* W, cudaStream_t stream,
                   int first_block = 0);

int dist_group(double* A, long long lda, int n, int nrows, int NB, double* Tlo, double* Tup, int k0, double* panel,
               int* info, cudaStream_t stream, double* Tsub = nullptr);
int dist_store(double* A, long long lda, int n, int nrows, int NB, int k0, const double* panel, cudaStream_t stream);
int dist_update(double* A, long long lda, int n, int nrows, int NB, int k0, const double* panel, int col0, int col1,
                cudaStream_t stream);
int dist_finish(double* A, long long lda, int n, int nrows, int NB, double* Tlo, double* Tup, double* scratch,
                int* info, cudaStream_t stream);
int debug_chain_stamps(long long* out);
int debug_potf2(double* A, long long lda, int nb, double* Tlo, double* Tup, int* info, long long* dbg,
                cudaStream_t stream);

// ---- lik.cu
struct LoadRowsArgs {
    const double* P; const double* z; int n, q;
    double* rows; long long ld; long long stride;  // destination rows (q+1) x ld
    double* p0rows; long long ld0;                  // optional second copy of the P rows (batch 0 only)
    long long strideZ = 0;                          // batch stride of z (elements; 0 = shared observations)
};

struct FinalizeArgs {
    double* rows; long long ld; long long strideRows;  // (q+1) x n whitened rows: P~^T then z~ (in/out)
    const double* p0rows; long long ld0;                // q x n raw P^T (shared), orthogonalised on a copy
    double* p0work; long long strideP0;                 // q x ld0 scratch per batch entry (may alias for batch 1)
    const double* Ldiag; long long ldl; long long strideL;  // L (diag read at i*(ldl+1))
    int n, q;
    double* out; long long strideOut;       // [value, logdet, quad, 2 sum log L_ii, logdetR~*2, logdetR0*2, info]: 7 slots when
                                            // strideOut is 0 or >= 7, the value alone otherwise
    double* Rt; long long strideRt;         // optional (q+1) x (q+1): upper triangle = R~ with the z column
    const int* info; long long strideInfo;  // non-zero -> value = +inf
    const double* ldr0_in;                  // optional precomputed sum log R0_ii (skips the raw-basis pass)
};

struct URowsArgs {
    const double* R; long long ldr; int r;
    const double* Tup; long long ldt;
    double* U; long long ldu;
    int n;
    int j0, j1;  // columns of U to compute: [j0, j1)  (j1 <= 0: all)
    int batch = 1; long long strideR = 0, strideT = 0, strideU = 0;  // batched form (blockIdx.y = entry)
};

struct DenseGradArgs {
    const double* Kinv; long long ldk; const double* U; long long ldu; int r; int n;
    double* dK; long long lddk; double half;
};

int launch_load_rows(const LoadRowsArgs& a, int batch, cudaStream_t stream);
int launch_copy_lower(const double* K, long long ldk, double* A, long long lda, int n, cudaStream_t stream);
int launch_finalize(const FinalizeArgs& a, int batch, cudaStream_t stream);
int launch_logdet_r0(const double* p0rows, double* p0work, long long ld, int n, int q, double* out,
                     cudaStream_t stream);
int launch_urows(const URowsArgs& a, cudaStream_t stream);
int launch_dense_grad(const DenseGradArgs& a, cudaStream_t stream);
int launch_unit_rows(double* B, long long ld, int m, int n, int col0, cudaStream_t stream);
int launch_loo(const double* Kinv, long long ldk, const double* U, long long ldu, int q, int n, const double* z,
               double* zloo, double* s2loo, double* eloo, cudaStream_t stream);

// ---- predict.cu
struct RowDotsArgs {
    double* V; long long ldv; int m, n, q;          // V rows v_t (m x n)
    const double* R; long long ldr;                  // (q+1) rows: Q~ rows then the residual r
    const double* Rt;                                // (q+1) x (q+1): R~ and, in the last column, Q~^T z~
    const double* Pt;                                // m x q basis at the test points (NULL if q == 0)
    const double* ktt; double ktt_scalar;            // prior variances
    double* dots;                                    // m x (q+2) scratch: e_t (q), v.r, |v|^2
    double* mean; double* var;
};
int launch_rowdots(const RowDotsArgs& a, cudaStream_t stream);
int launch_wrows(const RowDotsArgs& a, cudaStream_t stream);
int launch_transpose(const double* in, long long ldi, double* out, long long ldo, int rows, int cols,
                     cudaStream_t stream);

}  // namespace gpmp
