// The C-ABI of libgpmp_b200 (include/gpmp_b200.h): argument checks, workspace layouts and the launch
// sequences.  Nothing here allocates, synchronises or throws; every call only enqueues on `stream`.
#include <mutex>
#include <vector>
#include "internal.cuh"

namespace gpmp {

// ---- launch accounting / CUDA-event profiling -----------------------------------------------------
static unsigned long long g_launches = 0;
static ProfState g_prof = {};
struct EvPair { cudaEvent_t a, b; };
static std::vector<EvPair> g_events[KC_COUNT];
static std::vector<cudaEvent_t> g_pool;
struct TimelineEntry { int cls; size_t idx; cudaStream_t stream; };
static std::vector<TimelineEntry> g_timeline;  // launch order across classes (development timeline dump)

ProfState& prof() { return g_prof; }
static cudaEvent_t get_event() {
    if (!g_pool.empty()) {
        cudaEvent_t e = g_pool.back();
        g_pool.pop_back();
        return e;
    }
    cudaEvent_t e;
    cudaEventCreate(&e);
    return e;
}
void prof_begin(int cls, cudaStream_t s) {
    ++g_launches;
    if (!g_prof.enabled) return;
    EvPair p;
    p.a = get_event();
    p.b = get_event();
    cudaEventRecord(p.a, s);
    g_timeline.push_back({cls, g_events[cls].size(), s});
    g_events[cls].push_back(p);
}
void prof_end(int cls, double work, cudaStream_t s) {
    if (!g_prof.enabled) return;
    cudaEventRecord(g_events[cls].back().b, s);
    g_prof.launches[cls] += 1;
    g_prof.work[cls] += work;
}

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
static inline long long round_ld(int n) { return ((long long)n + 15) / 16 * 16; }

// ---- workspace layouts ---------------------------------------------------------------------------
struct PotrfWs {
    int NB, nblk;
    size_t off_tlo, off_tup, off_w, off_tsub, total;
};
static PotrfWs potrf_ws(int n, int nrows) {
    PotrfWs w;
    w.NB = potrf_block_size(n);
    w.nblk = ceil_div(n > 0 ? n : 1, w.NB);
    size_t tb = (size_t)w.nblk * w.NB * w.NB * 8;
    w.off_tlo = 0;
    w.off_tup = align_up(tb, 256);
    w.off_w = w.off_tup + align_up(tb, 256);
    size_t wrows = (size_t)(nrows > w.NB ? nrows : w.NB);
    w.off_tsub = w.off_w + align_up(2 * wrows * w.NB * 8, 256);  // two panel buffers (look-ahead)
    // block-diagonal tile inverses for the chain's substitution solve: one 128x128 tile per 128 columns
    w.total = w.off_tsub + align_up((size_t)ceil_div(n > 0 ? n : 1, 128) * 128 * 128 * 8, 256);
    return w;
}

struct LikWs {
    int n, q, r, nrows;
    long long lda;
    PotrfWs pw;
    size_t off_A, off_potrf, off_p0rows, off_p0work, off_small, off_U, off_Tlo, off_Tup, off_Kinv, off_partial;
    size_t total_value, total_grad;
};
static LikWs lik_ws(int n, int q, int d) {
    LikWs w;
    w.n = n; w.q = q; w.r = q + 1; w.nrows = n + w.r;
    w.lda = round_ld(n);
    w.pw = potrf_ws(n, w.nrows);
    size_t o = 0;
    w.off_A = o; o += align_up((size_t)(w.nrows + 7) * w.lda * 8, 256);
    w.off_potrf = o; o += w.pw.total;
    w.off_p0rows = o; o += align_up((size_t)(q > 0 ? q : 1) * w.lda * 8, 256);
    w.off_p0work = o; o += align_up((size_t)(q > 0 ? q : 1) * w.lda * 8, 256);
    w.off_small = o; o += align_up((size_t)(w.r * w.r + 16) * 8, 256);
    w.off_U = o; o += align_up((size_t)w.r * w.lda * 8, 256);
    w.total_value = o;
    w.off_Tlo = o; o += align_up((size_t)n * w.lda * 8, 256);
    w.off_Tup = o; o += align_up((size_t)n * w.lda * 8, 256);
    w.off_Kinv = o; o += align_up((size_t)n * w.lda * 8, 256);
    w.off_partial = o; o += align_up(contract_workspace_bytes(n, n, d > 0 ? d : 1), 256);
    w.total_grad = o;
    return w;
}

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// Which likelihood workspaces hold a leading block of T = L^-1 already (left by gpmp_lik_value under the tail of its
// factorisation, potrf.cu early_inverse): workspace address -> size of the block.  Every call that rebuilds the fit
// state of a workspace clears its entry first; gpmp_lik_grad / gpmp_lik_loo consume it.
static std::mutex g_early_mutex;
static std::vector<std::pair<const void*, int>> g_early;
static void early_set(const void* work, int prefix) {
    std::lock_guard<std::mutex> lock(g_early_mutex);
    for (auto it = g_early.begin(); it != g_early.end(); ++it)
        if (it->first == work) { g_early.erase(it); break; }
    if (prefix > 0) g_early.emplace_back(work, prefix);
}
static int early_get(const void* work) {
    std::lock_guard<std::mutex> lock(g_early_mutex);
    for (auto& e : g_early)
        if (e.first == work) return e.second;
    return 0;
}

// Library-owned streams for the chunk pipeline of the batched criterion: one set per (device, caller stream).
constexpr int BATCH_SLOTS = 4;
struct BatchStreams {
    cudaStream_t caller = nullptr;
    int dev = -1;
    cudaStream_t q[BATCH_SLOTS] = {};
    cudaEvent_t fork = nullptr, join[BATCH_SLOTS] = {};
    bool ok = false;
};
static std::mutex g_bs_mutex;
static std::vector<BatchStreams*> g_bs_sets;
static BatchStreams* batch_streams(cudaStream_t caller) {
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lock(g_bs_mutex);
    for (BatchStreams* b : g_bs_sets)
        if (b->dev == dev && b->caller == caller) return b;
    BatchStreams* b = new BatchStreams();
    b->caller = caller;
    b->dev = dev;
    b->ok = cudaEventCreateWithFlags(&b->fork, cudaEventDisableTiming) == cudaSuccess;
    for (int k = 0; k < BATCH_SLOTS && b->ok; ++k)
        b->ok = cudaStreamCreateWithFlags(&b->q[k], cudaStreamNonBlocking) == cudaSuccess &&
                cudaEventCreateWithFlags(&b->join[k], cudaEventDisableTiming) == cudaSuccess;
    g_bs_sets.push_back(b);
    return b;
}

// FP64 tensor pipe issue-rate probe: every warp keeps 8 independent DMMA.8x8x4 chains in registers.
__global__ void __launch_bounds__(512, 1) dmma_peak_kernel(double* sink, int iters, double a, double b) {
    double c0[8], c1[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) c0[i] = c1[i] = (double)(threadIdx.x + i);
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int i = 0; i < 8; ++i) dmma884(c0[i], c1[i], a, b);
    double acc = 0.0;
#pragma unroll
    for (int i = 0; i < 8; ++i) acc += c0[i] + c1[i];
    if (acc == 123.456) sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;  // keeps the chains alive
}

}  // namespace gpmp

using namespace gpmp;

extern "C" {

int gpmp_abi_version(void) { return 1; }
unsigned long long gpmp_launch_count(void) { return g_launches; }

int gpmp_measure_dmma_peak(int ctas, int iters, double* sink_dev, double* flops_out, void* stream) {
    if (ctas <= 0 || iters <= 0 || !sink_dev) return GPMP_ERR_ARG;
    dmma_peak_kernel<<<ctas, 512, 0, (cudaStream_t)stream>>>(sink_dev, iters, 1.0000001, 1e-9);
    GPMP_CHECK_LAUNCH();
    // 16 warps x 8 DMMAs per round, 8 x 8 x 4 multiply-adds each
    if (flops_out) *flops_out = (double)ctas * 16.0 * 8.0 * (double)iters * 2.0 * 8.0 * 8.0 * 4.0;
    return GPMP_OK;
}

int gpmp_prof_enable(int enable) {
    g_prof.enabled = enable == 2 ? 2 : (enable ? 1 : 0);  // 2: also serialise the look-ahead streams
    return GPMP_OK;
}
// development hook (not part of the header): rows [class, stream id, start ms, duration ms] of every launch
// recorded since profiling was enabled, in launch order; call before gpmp_prof_read (which recycles the events)
int gpmp_debug_timeline(double* out, int max_rows) {
    if (g_timeline.empty()) return 0;
    std::vector<cudaStream_t> streams;
    const EvPair& first = g_events[g_timeline[0].cls][g_timeline[0].idx];
    int n = 0;
    for (auto& t : g_timeline) {
        if (n >= max_rows) break;
        const EvPair& p = g_events[t.cls][t.idx];
        if (cudaEventSynchronize(p.b) != cudaSuccess) return -1;
        float st = 0.f, du = 0.f;
        cudaEventElapsedTime(&st, first.a, p.a);
        cudaEventElapsedTime(&du, p.a, p.b);
        size_t sid = 0;
        for (; sid < streams.size(); ++sid)
            if (streams[sid] == t.stream) break;
        if (sid == streams.size()) streams.push_back(t.stream);
        out[4 * n + 0] = t.cls; out[4 * n + 1] = (double)sid; out[4 * n + 2] = st; out[4 * n + 3] = du;
        ++n;
    }
    return n;
}

int gpmp_prof_read(int cls, double* ms, unsigned long long* launches, double* work) {
    if (cls < 0 || cls >= KC_COUNT) return GPMP_ERR_ARG;
    g_timeline.clear();
    double total = 0.0;
    for (auto& p : g_events[cls]) {
        if (cudaEventSynchronize(p.b) != cudaSuccess) return GPMP_ERR_CUDA;
        float t = 0.f;
        cudaEventElapsedTime(&t, p.a, p.b);
        total += t;
        g_pool.push_back(p.a);
        g_pool.push_back(p.b);
    }
    g_events[cls].clear();
    if (ms) *ms = total;
    if (launches) *launches = g_prof.launches[cls];
    if (work) *work = g_prof.work[cls];
    g_prof.launches[cls] = 0;
    g_prof.work[cls] = 0.0;
    return GPMP_OK;
}

// ---- distances / covariance ---------------------------------------------------------------------
static int dist_spec(const double* loginvrho, int d, gpmp_cov_spec* s) {
    if (!loginvrho || d < 1 || d > GPMP_MAX_DIM) return GPMP_ERR_DIM;
    s->p = 0; s->d = d; s->noise = 0; s->reserved = 0; s->log_sigma2 = 0.0; s->log_tau2 = 0.0;
    for (int j = 0; j < GPMP_MAX_DIM; ++j) s->loginvrho[j] = j < d ? loginvrho[j] : 0.0;
    return GPMP_OK;
}

int gpmp_scaled_distance(const double* loginvrho_host, int d, const double* x_dev, int n, const double* y_dev,
                         int m, double* D_dev, long long ldd, void* stream) {
    gpmp_cov_spec s;
    int rc = dist_spec(loginvrho_host, d, &s);
    if (rc) return rc;
    if (!x_dev || !D_dev || n < 0 || m < 0) return GPMP_ERR_ARG;
    const bool same = (y_dev == nullptr || y_dev == x_dev);
    return launch_matern_cov(&s, nullptr, 1, 0, x_dev, n, same ? nullptr : y_dev, same ? n : m, D_dev, ldd,
                             same ? COV_SYM_FULL : COV_RECT, 1, (cudaStream_t)stream);
}

int gpmp_scaled_distance_elementwise(const double* loginvrho_host, int d, const double* x_dev, const double* y_dev,
                                     int n, double* out_dev, void* stream) {
    gpmp_cov_spec s;
    int rc = dist_spec(loginvrho_host, d, &s);
    if (rc) return rc;
    if (!x_dev || !out_dev || n < 0) return GPMP_ERR_ARG;
    return launch_pairwise(&s, x_dev, y_dev, n, out_dev, 1, (cudaStream_t)stream);
}

int gpmp_maternp_kernel(int p, const double* h_dev, double* k_dev, double* dk_dev, long long count, void* stream) {
    if (p < 0 || p > GPMP_MAX_P) return GPMP_ERR_DIM;
    if (!h_dev || !k_dev || count < 0) return GPMP_ERR_ARG;
    return launch_maternp_elementwise(p, h_dev, k_dev, dk_dev, count, (cudaStream_t)stream);
}

int gpmp_matern_cov(const gpmp_cov_spec* spec, const double* x_dev, int n, const double* y_dev, int m,
                    double* K_dev, long long ldk, int mode, void* stream) {
    if (!spec || !x_dev || !K_dev || n < 0 || m < 0) return GPMP_ERR_ARG;
    const bool same = (y_dev == nullptr);
    int cm;
    if (same) cm = mode == GPMP_COV_LOWER ? COV_SYM_LOWER : COV_SYM_FULL;
    else {
        if (mode != GPMP_COV_FULL) return GPMP_ERR_ARG;
        cm = COV_RECT;
    }
    return launch_matern_cov(spec, nullptr, 1, 0, x_dev, n, y_dev, same ? n : m, K_dev, ldk, cm, 0,
                             (cudaStream_t)stream);
}

int gpmp_matern_cov_pairwise(const gpmp_cov_spec* spec, const double* x_dev, const double* y_dev, int n,
                             double* out_dev, void* stream) {
    if (!spec || !x_dev || !out_dev || n < 0) return GPMP_ERR_ARG;
    return launch_pairwise(spec, x_dev, y_dev, n, out_dev, 0, (cudaStream_t)stream);
}

size_t gpmp_contract_workspace_bytes(int n, int m, int d) { return contract_workspace_bytes(n, m, d); }

int gpmp_matern_cov_backward(const gpmp_cov_spec* spec, const double* x_dev, int n, const double* y_dev, int m,
                             const double* G_dev, long long ldg, double* grad_dev, void* partial_dev,
                             size_t partial_bytes, void* stream) {
    if (!spec || !x_dev || !G_dev || !grad_dev || !partial_dev) return GPMP_ERR_ARG;
    return launch_contract(spec, x_dev, n, y_dev, y_dev ? m : n, G_dev, ldg, nullptr, 0, 0, 0, 0, 1.0, grad_dev,
                           partial_dev, partial_bytes, (cudaStream_t)stream);
}

int gpmp_scaled_distance_backward(const double* loginvrho_host, int d, const double* x_dev, int n,
                                  const double* y_dev, int m, const double* G_dev, long long ldg, double* grad_dev,
                                  void* partial_dev, size_t partial_bytes, void* stream) {
    gpmp_cov_spec s;
    int rc = dist_spec(loginvrho_host, d, &s);
    if (rc) return rc;
    if (!x_dev || !G_dev || !grad_dev || !partial_dev) return GPMP_ERR_ARG;
    // grad_dev[0] is unused (no sigma2); grad_dev[1..d] = d/d loginvrho_j
    return launch_contract(&s, x_dev, n, y_dev, y_dev ? m : n, G_dev, ldg, nullptr, 0, 0, 0, 1, 1.0, grad_dev,
                           partial_dev, partial_bytes, (cudaStream_t)stream);
}

// ---- Cholesky family ---------------------------------------------------------------------------
size_t gpmp_potrf_workspace_bytes(int n, int nrows) { return potrf_ws(n, nrows > n ? nrows : n).total; }

int gpmp_potrf(double* A_dev, int n, int nrows, long long lda, void* work_dev, size_t work_bytes, int* info_dev,
               void* stream) {
    if (!A_dev || !work_dev || n < 0 || nrows < n || lda < n) return GPMP_ERR_ARG;
    if ((lda & 1) || !aligned16(A_dev) || !aligned16(work_dev)) return GPMP_ERR_ALIGN;
    PotrfWs w = potrf_ws(n, nrows);
    if (work_bytes < w.total) return GPMP_ERR_WORKSPACE;
    char* base = static_cast<char*>(work_dev);
    return potrf_core(A_dev, lda, 0, n, nrows, w.NB, (double*)(base + w.off_tlo), (double*)(base + w.off_tup), 0,
                      (double*)(base + w.off_w), 0, info_dev, 0, 1, (cudaStream_t)stream,
                      (double*)(base + w.off_tsub), 0, ceil_div(n > 0 ? n : 1, 128));
}

int gpmp_potri(const double* L_dev, int n, long long ldl, const void* potrf_work_dev, double* Tlo_dev,
               double* Tup_dev, double* Kinv_dev, long long ld, void* stream) {
    if (!L_dev || !potrf_work_dev || !Tlo_dev || !Tup_dev || !Kinv_dev || n < 0 || ld < n) return GPMP_ERR_ARG;
    if ((ld & 1) || (ldl & 1)) return GPMP_ERR_ALIGN;
    PotrfWs w = potrf_ws(n, n);
    const char* base = static_cast<const char*>(potrf_work_dev);
    return potri_core(L_dev, n, ldl, w.NB, (const double*)(base + w.off_tlo), (const double*)(base + w.off_tup),
                      Tlo_dev, Tup_dev, Kinv_dev, ld, (cudaStream_t)stream);
}

int gpmp_trsm_rows(const double* L_dev, int n, long long ldl, const void* potrf_work_dev, double* Bt_dev, int m,
                   long long ldb, int trans, void* scratch_dev, void* stream) {
    if (!L_dev || !potrf_work_dev || !Bt_dev || !scratch_dev || n < 0 || m < 0) return GPMP_ERR_ARG;
    if ((ldl & 1) || (ldb & 1)) return GPMP_ERR_ALIGN;
    PotrfWs w = potrf_ws(n, n);
    const char* base = static_cast<const char*>(potrf_work_dev);
    return trsm_rows_core(L_dev, n, ldl, w.NB, (const double*)(base + w.off_tlo), (const double*)(base + w.off_tup),
                          Bt_dev, m, ldb, trans, (double*)scratch_dev, (cudaStream_t)stream);
}

int gpmp_gemm_nt(const double* A_dev, long long lda, const double* B_dev, long long ldb, double* C_dev,
                 long long ldc, int M, int N, int K, double alpha, double beta, int tri, int lower, void* stream) {
    if (!A_dev || !B_dev || !C_dev || M < 0 || N < 0 || K < 0 || tri < 0 || tri > 4) return GPMP_ERR_ARG;
    GemmDesc g = gemm_desc();
    g.A = A_dev; g.lda = lda; g.B = B_dev; g.ldb = ldb; g.C = C_dev; g.ldc = ldc;
    g.M = M; g.N = N; g.K = K; g.alpha = alpha; g.beta = beta; g.krange = tri; g.lower = lower ? 1 : 0;
    g.reverse = (tri == KR_TO_ROW);
    return launch_gemm_nt(g, (cudaStream_t)stream);
}

int gpmp_transpose(const double* in_dev, long long ldi, double* out_dev, long long ldo, int rows, int cols,
                   void* stream) {
    if (!in_dev || !out_dev || rows < 0 || cols < 0) return GPMP_ERR_ARG;
    return launch_transpose(in_dev, ldi, out_dev, ldo, rows, cols, (cudaStream_t)stream);
}

// development / test hook (not declared in the header): the persistent form of gpmp_gemm_nt (gemm.cu
// gemm_nt_persist_kernel) with its SM filter exposed, so that tests can force every placement case -- all CTAs
// admitted (sm_first = 0), a single SM admitted, none admitted (the last CTA to leave then does all the work).
// batch2 products: pair b uses A + b stride2A, B + b stride2B, C + b stride2C.
int gpmp_debug_gemm_nt_persist(const double* A_dev, long long lda, const double* B_dev, long long ldb, double* C_dev,
                               long long ldc, int M, int N, int K, double alpha, double beta, int tri, int lower,
                               int batch2, long long stride2A, long long stride2B, long long stride2C, int sm_first,
                               void* stream) {
    if (!A_dev || !B_dev || !C_dev || M < 0 || N < 0 || K < 0 || tri < 0 || tri > 4 || batch2 < 1) return GPMP_ERR_ARG;
    GemmDesc g = gemm_desc();
    g.A = A_dev; g.lda = lda; g.B = B_dev; g.ldb = ldb; g.C = C_dev; g.ldc = ldc;
    g.M = M; g.N = N; g.K = K; g.alpha = alpha; g.beta = beta; g.krange = tri; g.lower = lower ? 1 : 0;
    g.reverse = (tri == KR_TO_ROW);
    g.batch2 = batch2; g.stride2A = stride2A; g.stride2B = stride2B; g.stride2C = stride2C;
    return launch_gemm_nt_persist(g, (cudaStream_t)stream, sm_first);
}

// development hook: phase timestamps (clock64) of the diagonal-tile kernel; not declared in the header
// development hook: phase clocks of the last stamped chain-step launch (GPMP_DEV_STAMPS=1), 64 values
int gpmp_debug_chain_stamps(long long* out_host) { return debug_chain_stamps(out_host); }
int gpmp_debug_potf2(double* A_dev, long long lda, int nb, double* Tlo_dev, double* Tup_dev, int* info_dev,
                     long long* clocks_dev, void* stream) {
    return debug_potf2(A_dev, lda, nb, Tlo_dev, Tup_dev, info_dev, clocks_dev, (cudaStream_t)stream);
}

// ---- likelihoods ---------------------------------------------------------------------------------
size_t gpmp_lik_workspace_bytes(int n, int q, int d, int want_grad) {
    if (n < 0 || q < 0 || q > GPMP_MAX_Q) return 0;
    LikWs w = lik_ws(n, q, d);
    return want_grad ? w.total_grad : w.total_value;
}

static int lik_prepare(const gpmp_cov_spec* spec, const double* K_dev, long long ldk, const double* x_dev, int n,
                       const double* z_dev, const double* P_dev, int q, const LikWs& w, char* base, int* info_dev,
                       cudaStream_t s) {
    double* A = (double*)(base + w.off_A);
    int rc;
    early_set(base, 0);  // the workspace's fit state is being rebuilt
    if (cudaMemsetAsync(info_dev, 0, sizeof(int), s) != cudaSuccess) return GPMP_ERR_CUDA;
    if (spec) rc = launch_matern_cov(spec, nullptr, 1, 0, x_dev, n, nullptr, n, A, w.lda, COV_SYM_LOWER, 0, s);
    else rc = launch_copy_lower(K_dev, ldk, A, w.lda, n, s);
    if (rc) return rc;
    LoadRowsArgs lr;
    lr.P = P_dev; lr.z = z_dev; lr.n = n; lr.q = q;
    lr.rows = A + (long long)n * w.lda; lr.ld = w.lda; lr.stride = 0;
    lr.p0rows = q > 0 ? (double*)(base + w.off_p0rows) : nullptr; lr.ld0 = w.lda;
    return launch_load_rows(lr, 1, s);
}

static int lik_finalize(int n, int q, const LikWs& w, char* base, double* out_dev, int* info_dev, cudaStream_t s) {
    double* A = (double*)(base + w.off_A);
    FinalizeArgs f;
    f.rows = A + (long long)n * w.lda; f.ld = w.lda; f.strideRows = 0;
    f.p0rows = (double*)(base + w.off_p0rows); f.ld0 = w.lda;
    f.p0work = (double*)(base + w.off_p0work); f.strideP0 = 0;
    f.Ldiag = A; f.ldl = w.lda; f.strideL = 0;
    f.n = n; f.q = q;
    f.out = out_dev; f.strideOut = 0;
    f.Rt = (double*)(base + w.off_small); f.strideRt = 0;
    f.info = info_dev; f.strideInfo = 0;
    f.ldr0_in = nullptr;
    return launch_finalize(f, 1, s);
}

static int lik_check(const gpmp_cov_spec* spec, const double* K_dev, const double* x_dev, int n, const double* z_dev,
                     const double* P_dev, int q, const void* work_dev) {
    if (n <= 0 || q < 0 || q > GPMP_MAX_Q || !z_dev || !work_dev) return GPMP_ERR_ARG;
    if (q > 0 && !P_dev) return GPMP_ERR_ARG;
    if (!spec && !K_dev) return GPMP_ERR_ARG;
    if (spec && !x_dev) return GPMP_ERR_ARG;
    if (!aligned16(work_dev)) return GPMP_ERR_ALIGN;
    return GPMP_OK;
}

int gpmp_lik_value(const gpmp_cov_spec* spec, const double* K_dev, long long ldk, const double* x_dev, int n,
                   const double* z_dev, const double* P_dev, int q, void* work_dev, size_t work_bytes,
                   double* out_dev, int* info_dev, void* stream) {
    int rc = lik_check(spec, K_dev, x_dev, n, z_dev, P_dev, q, work_dev);
    if (rc) return rc;
    if (!out_dev || !info_dev) return GPMP_ERR_ARG;
    cudaStream_t s = (cudaStream_t)stream;
    LikWs w = lik_ws(n, q, spec ? spec->d : 1);
    if (work_bytes < w.total_value) return GPMP_ERR_WORKSPACE;
    char* base = static_cast<char*>(work_dev);
    rc = lik_prepare(spec, K_dev, ldk, x_dev, n, z_dev, P_dev, q, w, base, info_dev, s);
    if (rc) return rc;
    char* pb = base + w.off_potrf;
    // a workspace that also holds the buffers of the gradient: the leading block of T = L^-1 is computed under the
    // tail of the factorisation (the K^-1 buffer is its scratch)
    EarlyInverse early{(double*)(base + w.off_Tlo), (double*)(base + w.off_Tup), (double*)(base + w.off_Kinv), w.lda};
    const bool with_t = work_bytes >= w.off_partial;
    rc = potrf_core((double*)(base + w.off_A), w.lda, 0, n, w.nrows, w.pw.NB, (double*)(pb + w.pw.off_tlo),
                    (double*)(pb + w.pw.off_tup), 0, (double*)(pb + w.pw.off_w), 0, info_dev, 0, 1, s,
                    (double*)(pb + w.pw.off_tsub), 0, ceil_div(n, 128), with_t ? &early : nullptr);
    if (rc) return rc;
    if (with_t) early_set(base, early_prefix(n, w.pw.NB));
    return lik_finalize(n, q, w, base, out_dev, info_dev, s);
}

// ---- the same evaluation with the factorisation partitioned over GPUs (panel exchange by the caller) -----
int gpmp_lik_dist_block(int n) { return potrf_block_size(n); }

int gpmp_lik_dist_prepare(const gpmp_cov_spec* spec, const double* K_dev, long long ldk, const double* x_dev, int n,
                          const double* z_dev, const double* P_dev, int q, void* work_dev, size_t work_bytes,
                          int* info_dev, void* stream) {
    int rc = lik_check(spec, K_dev, x_dev, n, z_dev, P_dev, q, work_dev);
    if (rc) return rc;
    if (!info_dev) return GPMP_ERR_ARG;
    LikWs w = lik_ws(n, q, spec ? spec->d : 1);
    if (work_bytes < w.total_value) return GPMP_ERR_WORKSPACE;
    return lik_prepare(spec, K_dev, ldk, x_dev, n, z_dev, P_dev, q, w, static_cast<char*>(work_dev), info_dev,
                       (cudaStream_t)stream);
}

int gpmp_lik_dist_group(int n, int q, void* work_dev, size_t work_bytes, int k0, double* panel_dev, int* info_dev,
                        void* stream) {
    if (n <= 0 || q < 0 || q > GPMP_MAX_Q || !work_dev || !panel_dev || !info_dev || k0 < 0 || k0 >= n)
        return GPMP_ERR_ARG;
    LikWs w = lik_ws(n, q, 1);
    if (work_bytes < w.total_value || (k0 % w.pw.NB) != 0) return GPMP_ERR_WORKSPACE;
    char* base = static_cast<char*>(work_dev);
    char* pb = base + w.off_potrf;
    return dist_group((double*)(base + w.off_A), w.lda, n, w.nrows, w.pw.NB, (double*)(pb + w.pw.off_tlo),
                      (double*)(pb + w.pw.off_tup), k0, panel_dev, info_dev, (cudaStream_t)stream,
                      (double*)(pb + w.pw.off_tsub));
}

int gpmp_lik_dist_store(int n, int q, void* work_dev, size_t work_bytes, int k0, const double* panel_dev,
                        void* stream) {
    if (n <= 0 || q < 0 || q > GPMP_MAX_Q || !work_dev || !panel_dev || k0 < 0 || k0 >= n) return GPMP_ERR_ARG;
    LikWs w = lik_ws(n, q, 1);
    if (work_bytes < w.total_value || (k0 % w.pw.NB) != 0) return GPMP_ERR_WORKSPACE;
    char* base = static_cast<char*>(work_dev);
    return dist_store((double*)(base + w.off_A), w.lda, n, w.nrows, w.pw.NB, k0, panel_dev, (cudaStream_t)stream);
}

int gpmp_lik_dist_update(int n, int q, void* work_dev, size_t work_bytes, int k0, const double* panel_dev, int col0,
                         int col1, void* stream) {
    if (n <= 0 || q < 0 || q > GPMP_MAX_Q || !work_dev || !panel_dev || k0 < 0 || k0 >= n) return GPMP_ERR_ARG;
    LikWs w = lik_ws(n, q, 1);
    if (work_bytes < w.total_value || (k0 % w.pw.NB) != 0) return GPMP_ERR_WORKSPACE;
    if (col1 > n) col1 = n;
    char* base = static_cast<char*>(work_dev);
    return dist_update((double*)(base + w.off_A), w.lda, n, w.nrows, w.pw.NB, k0, panel_dev, col0, col1,
                       (cudaStream_t)stream);
}

int gpmp_lik_dist_finish(int n, int q, void* work_dev, size_t work_bytes, double* out_dev, int* info_dev,
                         void* stream) {
    if (n <= 0 || q < 0 || q > GPMP_MAX_Q || !work_dev || !out_dev || !info_dev) return GPMP_ERR_ARG;
    LikWs w = lik_ws(n, q, 1);
    if (work_bytes < w.total_value) return GPMP_ERR_WORKSPACE;
    cudaStream_t s = (cudaStream_t)stream;
    char* base = static_cast<char*>(work_dev);
    char* pb = base + w.off_potrf;
    int rc = dist_finish((double*)(base + w.off_A), w.lda, n, w.nrows, w.pw.NB, (double*)(pb + w.pw.off_tlo),
                         (double*)(pb + w.pw.off_tup), (double*)(pb + w.pw.off_w), info_dev, s);
    if (rc) return rc;
    return lik_finalize(n, q, w, base, out_dev, info_dev, s);
}

int gpmp_lik_grad(const gpmp_cov_spec* spec, const double* x_dev, int n, int q, void* work_dev, size_t work_bytes,
                  double* grad_dev, double* dz_dev, double* dK_dev, long long lddk, void* stream) {
    if (n <= 0 || q < 0 || q > GPMP_MAX_Q || !work_dev) return GPMP_ERR_ARG;
    if (spec && (!x_dev || !grad_dev)) return GPMP_ERR_ARG;
    cudaStream_t s = (cudaStream_t)stream;
    LikWs w = lik_ws(n, q, spec ? spec->d : 1);
    if (work_bytes < w.total_grad) return GPMP_ERR_WORKSPACE;
    char* base = static_cast<char*>(work_dev);
    double* A = (double*)(base + w.off_A);
    char* pb = base + w.off_potrf;
    double* Tlo = (double*)(base + w.off_Tlo);
    double* Tup = (double*)(base + w.off_Tup);
    double* Kinv = (double*)(base + w.off_Kinv);
    double* U = (double*)(base + w.off_U);
    // U = [Q~; r] T needs only T: it runs on a library-owned side stream under the K^-1 = T^T T product
    BatchStreams* bs = (n >= 2048 && prof().enabled != 2) ? batch_streams(s) : nullptr;
    const bool side = bs && bs->ok;
    int rc = potri_core(A, n, w.lda, w.pw.NB, (const double*)(pb + w.pw.off_tlo), (const double*)(pb + w.pw.off_tup),
                        Tlo, Tup, Kinv, w.lda, s, 1, 0, 0, 0, early_get(base), side ? bs->fork : nullptr);
    if (rc) return rc;
    cudaStream_t su = side ? bs->q[0] : s;
    if (side && cudaStreamWaitEvent(su, bs->fork, 0) != cudaSuccess) return GPMP_ERR_CUDA;
    URowsArgs u;
    u.R = A + (long long)n * w.lda; u.ldr = w.lda; u.r = w.r; u.Tup = Tup; u.ldt = w.lda; u.U = U; u.ldu = w.lda;
    u.n = n; u.j0 = 0; u.j1 = 0;
    rc = launch_urows(u, su);
    if (rc) return rc;
    if (side && (cudaEventRecord(bs->join[0], su) != cudaSuccess || cudaStreamWaitEvent(s, bs->join[0], 0) != cudaSuccess))
        return GPMP_ERR_CUDA;
    if (spec) {
        rc = launch_contract(spec, x_dev, n, nullptr, n, Kinv, w.lda, U, w.lda, w.r, 1, 0, 0.5, grad_dev,
                             base + w.off_partial, w.total_grad - w.off_partial, s);
        if (rc) return rc;
    }
    if (dz_dev) {
        if (cudaMemcpyAsync(dz_dev, U + (long long)q * w.lda, (size_t)n * 8, cudaMemcpyDeviceToDevice, s) !=
            cudaSuccess)
            return GPMP_ERR_CUDA;
    }
    if (dK_dev) {
        DenseGradArgs dg;
        dg.Kinv = Kinv; dg.ldk = w.lda; dg.U = U; dg.ldu = w.lda; dg.r = w.r; dg.n = n;
        dg.dK = dK_dev; dg.lddk = lddk; dg.half = 0.5;
        rc = launch_dense_grad(dg, s);
        if (rc) return rc;
    }
    return GPMP_OK;
}

// ---- gradient of the partitioned evaluation: every piece works on one block of rows --------------------
size_t gpmp_lik_ws_offset(int n, int q, int which) {
    if (n <= 0 || q < 0 || q > GPMP_MAX_Q) return (size_t)-1;
    LikWs w = lik_ws(n, q, 1);
    switch (which) {
        case 0: return w.off_A;
        case 1: return w.off_Tup;
        case 2: return w.off_Kinv;
        case 3: return w.off_U;
        case 4: return w.off_Tlo;
        default: return (size_t)-1;
    }
}
long long gpmp_lik_ws_ld(int n) { return round_ld(n); }

struct DistGradCtx {
    LikWs w; char* base; double *A, *Tup, *Kinv, *U; const double *Tlo_c, *Tup_c; double* Wsc;
};
static int dist_grad_ctx(int n, int q, void* work_dev, size_t work_bytes, int row0, int rows, DistGradCtx* c) {
    if (n <= 0 || q < 0 || q > GPMP_MAX_Q || !work_dev || row0 < 0 || rows <= 0 || row0 + rows > n)
        return GPMP_ERR_ARG;
    c->w = lik_ws(n, q, 1);
    if (work_bytes < c->w.off_partial) return GPMP_ERR_WORKSPACE;
    if (row0 % c->w.pw.NB) return GPMP_ERR_ARG;
    c->base = static_cast<char*>(work_dev);
    char* pb = c->base + c->w.off_potrf;
    c->A = (double*)(c->base + c->w.off_A);
    c->Tup = (double*)(c->base + c->w.off_Tup);
    c->Kinv = (double*)(c->base + c->w.off_Kinv);
    c->U = (double*)(c->base + c->w.off_U);
    c->Tlo_c = (const double*)(pb + c->w.pw.off_tlo);
    c->Tup_c = (const double*)(pb + c->w.pw.off_tup);
    c->Wsc = (double*)(pb + c->w.pw.off_w);
    return GPMP_OK;
}

int gpmp_lik_dist_tup_rows(int n, int q, void* work_dev, size_t work_bytes, int row0, int rows, void* stream) {
    DistGradCtx c;
    int rc = dist_grad_ctx(n, q, work_dev, work_bytes, row0, rows, &c);
    if (rc) return rc;
    cudaStream_t s = (cudaStream_t)stream;
    double* Bt = c.Tup + (long long)row0 * c.w.lda;
    rc = launch_unit_rows(Bt, c.w.lda, rows, n, row0, s);
    if (rc) return rc;
    return trsm_rows_core(c.A, n, c.w.lda, c.w.pw.NB, c.Tlo_c, c.Tup_c, Bt, rows, c.w.lda, 0, c.Wsc, s,
                          row0 / c.w.pw.NB);
}

int gpmp_lik_dist_kinv_rows(int n, int q, void* work_dev, size_t work_bytes, int row0, int rows, void* stream) {
    DistGradCtx c;
    int rc = dist_grad_ctx(n, q, work_dev, work_bytes, row0, rows, &c);
    if (rc) return rc;
    GemmDesc g = gemm_desc();
    g.A = c.Tup + (long long)row0 * c.w.lda; g.lda = c.w.lda;
    g.B = c.Tup; g.ldb = c.w.lda;
    g.C = c.Kinv + (long long)row0 * c.w.lda; g.ldc = c.w.lda;
    g.M = rows; g.N = row0 + rows; g.K = n; g.krange = KR_FROM_ROW; g.ktrim_off = row0;
    return launch_gemm_nt(g, (cudaStream_t)stream);
}

int gpmp_lik_dist_u_cols(int n, int q, void* work_dev, size_t work_bytes, int row0, int rows, void* stream) {
    DistGradCtx c;
    int rc = dist_grad_ctx(n, q, work_dev, work_bytes, row0, rows, &c);
    if (rc) return rc;
    URowsArgs u;
    u.R = c.A + (long long)n * c.w.lda; u.ldr = c.w.lda; u.r = c.w.r; u.Tup = c.Tup; u.ldt = c.w.lda;
    u.U = c.U; u.ldu = c.w.lda; u.n = n; u.j0 = row0; u.j1 = row0 + rows;
    return launch_urows(u, (cudaStream_t)stream);
}

int gpmp_lik_dist_contract_rows(const gpmp_cov_spec* spec, const double* x_dev, int n, int q, void* work_dev,
                                size_t work_bytes, int row0, int rows, double* grad_dev, void* stream) {
    if (!spec || !x_dev || !grad_dev) return GPMP_ERR_ARG;
    DistGradCtx c;
    int rc = dist_grad_ctx(n, q, work_dev, work_bytes, row0, rows, &c);
    if (rc) return rc;
    LikWs w = lik_ws(n, q, spec->d);
    if (work_bytes < w.total_grad) return GPMP_ERR_WORKSPACE;
    return launch_contract(spec, x_dev, n, nullptr, n, c.Kinv, w.lda, c.U, w.lda, w.r, 1, 0, 0.5, grad_dev,
                           c.base + w.off_partial, w.total_grad - w.off_partial, (cudaStream_t)stream, row0 / 64,
                           (row0 + rows + 63) / 64);
}

int gpmp_lik_loo(int n, int q, void* work_dev, size_t work_bytes, const double* z_dev, double* zloo_dev,
                 double* s2loo_dev, double* eloo_dev, void* stream) {
    if (n <= 0 || q < 0 || q > GPMP_MAX_Q || !work_dev || !z_dev || !zloo_dev || !s2loo_dev || !eloo_dev)
        return GPMP_ERR_ARG;
    cudaStream_t s = (cudaStream_t)stream;
    LikWs w = lik_ws(n, q, 1);
    if (work_bytes < w.total_grad - (w.total_grad - w.off_partial)) return GPMP_ERR_WORKSPACE;
    char* base = static_cast<char*>(work_dev);
    double* A = (double*)(base + w.off_A);
    char* pb = base + w.off_potrf;
    double* Tlo = (double*)(base + w.off_Tlo);
    double* Tup = (double*)(base + w.off_Tup);
    double* Kinv = (double*)(base + w.off_Kinv);
    double* U = (double*)(base + w.off_U);
    int rc = potri_core(A, n, w.lda, w.pw.NB, (const double*)(pb + w.pw.off_tlo), (const double*)(pb + w.pw.off_tup),
                        Tlo, Tup, Kinv, w.lda, s, 1, 0, 0, 0, early_get(base));
    if (rc) return rc;
    URowsArgs u;
    u.R = A + (long long)n * w.lda; u.ldr = w.lda; u.r = w.r; u.Tup = Tup; u.ldt = w.lda; u.U = U; u.ldu = w.lda;
    u.n = n; u.j0 = 0; u.j1 = 0;
    rc = launch_urows(u, s);
    if (rc) return rc;
    return launch_loo(Kinv, w.lda, U, w.lda, q, n, z_dev, zloo_dev, s2loo_dev, eloo_dev, s);
}

// ---- prediction -----------------------------------------------------------------------------------
size_t gpmp_predict_scratch_bytes(int n, int q, int m) {
    const int NB = potrf_block_size(n);
    return align_up((size_t)m * NB * 8, 256) + align_up((size_t)m * (q + 2) * 8, 256);
}

int gpmp_predict_chunk(const gpmp_cov_spec* spec, const double* x_dev, int n, int q, void* work_dev,
                       size_t work_bytes, const double* xt_dev, int m, const double* Pt_dev, const double* ktt_dev,
                       double* Vt_dev, long long ldv, void* scratch_dev, size_t scratch_bytes, double* mean_dev,
                       double* var_dev, int want_lambda, void* stream) {
    if (n <= 0 || m < 0 || q < 0 || q > GPMP_MAX_Q || !work_dev || !Vt_dev || !scratch_dev || !mean_dev || !var_dev)
        return GPMP_ERR_ARG;
    if (spec && (!x_dev || !xt_dev)) return GPMP_ERR_ARG;
    if (q > 0 && !Pt_dev) return GPMP_ERR_ARG;
    if (!spec && !ktt_dev) return GPMP_ERR_ARG;
    if ((ldv & 1) || ldv < n || m > 65535) return GPMP_ERR_ARG;
    if (scratch_bytes < gpmp_predict_scratch_bytes(n, q, m)) return GPMP_ERR_WORKSPACE;
    if (m == 0) return GPMP_OK;
    cudaStream_t s = (cudaStream_t)stream;
    LikWs w = lik_ws(n, q, spec ? spec->d : 1);
    if (work_bytes < w.total_value) return GPMP_ERR_WORKSPACE;
    char* base = static_cast<char*>(work_dev);
    double* A = (double*)(base + w.off_A);
    char* pb = base + w.off_potrf;
    const double* Tlo_c = (const double*)(pb + w.pw.off_tlo);
    const double* Tup_c = (const double*)(pb + w.pw.off_tup);
    double* Wsc = (double*)scratch_dev;
    double* dots = (double*)((char*)scratch_dev + align_up((size_t)m * w.pw.NB * 8, 256));
    int rc;
    if (spec) {
        rc = launch_matern_cov(spec, nullptr, 1, 0, xt_dev, m, x_dev, n, Vt_dev, ldv, COV_RECT, 0, s);
        if (rc) return rc;
    }
    rc = trsm_rows_core(A, n, w.lda, w.pw.NB, Tlo_c, Tup_c, Vt_dev, m, ldv, 0, Wsc, s);
    if (rc) return rc;
    RowDotsArgs rd;
    rd.V = Vt_dev; rd.ldv = ldv; rd.m = m; rd.n = n; rd.q = q;
    rd.R = A + (long long)n * w.lda; rd.ldr = w.lda;
    rd.Rt = (const double*)(base + w.off_small);
    rd.Pt = Pt_dev; rd.ktt = ktt_dev; rd.ktt_scalar = spec ? exp(spec->log_sigma2) : 0.0;
    rd.dots = dots; rd.mean = mean_dev; rd.var = var_dev;
    rc = launch_rowdots(rd, s);
    if (rc) return rc;
    if (want_lambda) {
        rc = launch_wrows(rd, s);
        if (rc) return rc;
        if (want_lambda == 1) {
            rc = trsm_rows_core(A, n, w.lda, w.pw.NB, Tlo_c, Tup_c, Vt_dev, m, ldv, 1, Wsc, s);
            if (rc) return rc;
        }
    }
    return GPMP_OK;
}

int gpmp_lik_trsm_rows(int n, int q, void* work_dev, size_t work_bytes, double* Bt_dev, int m, long long ldb,
                       int trans, void* scratch_dev, void* stream) {
    if (n <= 0 || q < 0 || q > GPMP_MAX_Q || !work_dev || !Bt_dev || !scratch_dev || m < 0) return GPMP_ERR_ARG;
    if (ldb & 1) return GPMP_ERR_ALIGN;
    LikWs w = lik_ws(n, q, 1);
    if (work_bytes < w.total_value) return GPMP_ERR_WORKSPACE;
    char* base = static_cast<char*>(work_dev);
    char* pb = base + w.off_potrf;
    return trsm_rows_core((double*)(base + w.off_A), n, w.lda, w.pw.NB, (const double*)(pb + w.pw.off_tlo),
                          (const double*)(pb + w.pw.off_tup), Bt_dev, m, ldb, trans, (double*)scratch_dev,
                          (cudaStream_t)stream);
}

// ---- batched criterion -----------------------------------------------------------------------------
struct BatchWs {
    int n, q, r, nrows, NB, nblk;
    long long lda;
    size_t per_A, per_T, per_W, per_mdev, per_tsub, per_particle;
    size_t shared;  // p0rows + p0work + ldr0 + dummy out
};
static BatchWs batch_ws(int n, int q) {
    BatchWs w;
    w.n = n; w.q = q; w.r = q + 1; w.nrows = n + w.r;
    w.lda = round_ld(n);
    w.NB = potrf_block_size(n);
    w.nblk = ceil_div(n, w.NB);
    w.per_A = align_up((size_t)(w.nrows + 1) * w.lda * 8, 256);
    w.per_T = align_up((size_t)w.nblk * w.NB * w.NB * 8, 256);
    w.per_W = align_up((size_t)2 * (w.nrows > w.NB ? w.nrows : w.NB) * w.NB * 8, 256);
    w.per_mdev = align_up(sizeof(MaternDev), 256);
    w.per_tsub = 128 * 128 * 8;
    w.per_particle = w.per_A + 2 * w.per_T + w.per_W + w.per_mdev + w.per_tsub;
    w.shared = 2 * align_up((size_t)(q > 0 ? q : 1) * w.lda * 8, 256) + 256 + 256;
    return w;
}

size_t gpmp_criterion_batched_bytes(int n, int q, int nbatch) {
    if (n <= 0 || q < 0 || nbatch <= 0) return 0;
    BatchWs w = batch_ws(n, q);
    return w.shared + (size_t)nbatch * w.per_particle;
}

int gpmp_criterion_batched(const gpmp_cov_spec* spec, const double* theta_dev, int N, const double* x_dev, int n,
                           const double* z_dev, const double* P_dev, int q, void* work_dev, size_t work_bytes,
                           double* values_dev, int* info_dev, void* stream) {
    if (!spec || !theta_dev || !x_dev || !z_dev || !work_dev || !values_dev || !info_dev) return GPMP_ERR_ARG;
    if (n <= 0 || N < 0 || q < 0 || q > GPMP_MAX_Q || (q > 0 && !P_dev)) return GPMP_ERR_ARG;
    if (N == 0) return GPMP_OK;
    cudaStream_t s = (cudaStream_t)stream;
    BatchWs w = batch_ws(n, q);
    if (work_bytes < w.shared + w.per_particle) return GPMP_ERR_WORKSPACE;
    long long cap = (long long)((work_bytes - w.shared) / w.per_particle);
    if (cap > 32768) cap = 32768;
    if (cap > N) cap = N;
    char* base = static_cast<char*>(work_dev);
    size_t rowsb = align_up((size_t)(q > 0 ? q : 1) * w.lda * 8, 256);
    double* p0rows = (double*)base;
    double* p0work = (double*)(base + rowsb);
    double* ldr0 = (double*)(base + 2 * rowsb);
    char* pbase = base + w.shared;
    double* A = (double*)pbase;
    double* Tlo = (double*)(pbase + (size_t)cap * w.per_A);
    double* Tup = (double*)(pbase + (size_t)cap * (w.per_A + w.per_T));
    double* W = (double*)(pbase + (size_t)cap * (w.per_A + 2 * w.per_T));
    MaternDev* mdev = (MaternDev*)(pbase + (size_t)cap * (w.per_A + 2 * w.per_T + w.per_W));
    double* Tsub = (double*)(pbase + (size_t)cap * (w.per_A + 2 * w.per_T + w.per_W + w.per_mdev));
    const long long sA = (long long)(w.per_A / 8), sT = (long long)(w.per_T / 8), sW = (long long)(w.per_W / 8);
    const long long sTsub = (long long)(w.per_tsub / 8);
    int rc;
    if (cudaMemsetAsync(info_dev, 0, sizeof(int) * (size_t)N, s) != cudaSuccess) return GPMP_ERR_CUDA;
    const int w_theta = 1 + spec->noise + spec->d;
    // The particles in flight are split over up to BATCH_SLOTS workspace slots, each served by its own
    // library-owned stream: the phases of one chunk are bound by different things (K build: instruction issue;
    // tile kernel: the pivot chain's latency; panel solve / SYRK: the DMMA pipe), so chunks at different phases
    // share the SMs instead of taking turns.  A slot is always used on the same stream, which orders its reuse.
    cudaStream_t slot_stream[BATCH_SLOTS];
    int nslots = (int)(cap / 256);
    nslots = nslots < 1 ? 1 : (nslots > BATCH_SLOTS ? BATCH_SLOTS : nslots);
    if (prof().enabled == 2) nslots = 1;  // serialised profiling pass: exclusive kernel times
    BatchStreams* bs = nslots > 1 ? batch_streams(s) : nullptr;
    if (!bs || !bs->ok) nslots = 1;
    const long long chunk = (cap + nslots - 1) / nslots;
    for (int k = 0; k < nslots; ++k) slot_stream[k] = nslots > 1 ? bs->q[k] : s;
    if (q > 0) {
        // the raw basis rows and sum log R0_ii are shared by every chunk: written once, on the caller's stream,
        // before the fork
        LoadRowsArgs l0;
        l0.P = P_dev; l0.z = z_dev; l0.n = n; l0.q = q;
        l0.rows = A + (long long)n * w.lda; l0.ld = w.lda; l0.stride = sA;
        l0.p0rows = p0rows; l0.ld0 = w.lda;
        rc = launch_load_rows(l0, 1, s);
        if (rc) return rc;
        rc = launch_logdet_r0(p0rows, p0work, w.lda, n, q, ldr0, s);
        if (rc) return rc;
    }
    if (nslots > 1) {
        if (cudaEventRecord(bs->fork, s) != cudaSuccess) return GPMP_ERR_CUDA;
        for (int k = 0; k < nslots; ++k)
            if (cudaStreamWaitEvent(slot_stream[k], bs->fork, 0) != cudaSuccess) return GPMP_ERR_CUDA;
    }
    int slot = 0;
    for (long long c0 = 0; c0 < N; c0 += chunk, slot = (slot + 1) % nslots) {
        const int nb = (int)((N - c0) < chunk ? (N - c0) : chunk);
        cudaStream_t cs = slot_stream[slot];
        const long long e0 = (long long)slot * chunk;  // first workspace entry of this slot
        double* As = A + e0 * sA;
        MaternDev* ms = mdev + e0;  // MaternDev entries are packed (sizeof), not per_mdev-strided
        rc = launch_prep_theta(spec, theta_dev + c0 * w_theta, nb, 1, ms, cs);
        if (rc) return rc;
        rc = launch_matern_cov(spec, ms, nb, sA, x_dev, n, nullptr, n, As, w.lda, COV_SYM_LOWER, 0, cs);
        if (rc) return rc;
        LoadRowsArgs lr;
        lr.P = P_dev; lr.z = z_dev; lr.n = n; lr.q = q;
        lr.rows = As + (long long)n * w.lda; lr.ld = w.lda; lr.stride = sA;
        lr.p0rows = nullptr; lr.ld0 = w.lda;
        rc = launch_load_rows(lr, nb, cs);
        if (rc) return rc;
        rc = potrf_core(As, w.lda, sA, n, w.nrows, w.NB, Tlo + e0 * sT, Tup + e0 * sT, sT, W + e0 * sW, sW,
                        info_dev + c0, 1, nb, cs, Tsub + e0 * sTsub, sTsub, 1);
        if (rc) return rc;
        FinalizeArgs f;
        f.rows = lr.rows; f.ld = w.lda; f.strideRows = sA;
        f.p0rows = p0rows; f.ld0 = w.lda; f.p0work = p0work; f.strideP0 = 0;
        f.Ldiag = As; f.ldl = w.lda; f.strideL = sA;
        f.n = n; f.q = q;
        f.Rt = nullptr; f.strideRt = 0;
        f.info = info_dev + c0; f.strideInfo = 1;
        f.out = values_dev + c0; f.strideOut = 1;
        f.ldr0_in = q > 0 ? ldr0 : nullptr;
        rc = launch_finalize(f, nb, cs);
        if (rc) return rc;
    }
    if (nslots > 1) {
        for (int k = 0; k < nslots; ++k)
            if (cudaEventRecord(bs->join[k], slot_stream[k]) != cudaSuccess ||
                cudaStreamWaitEvent(s, bs->join[k], 0) != cudaSuccess)
                return GPMP_ERR_CUDA;
    }
    return GPMP_OK;
}


// ---- batched criterion with gradients (SVGD particles, mini-batches) -----------------------------------------
// Entry b: theta_b, points x + b x_stride, observations z + b z_stride (strides in elements, 0 = shared), shared
// mean basis P.  Same pipeline as gpmp_lik_value + gpmp_lik_grad with a batch dimension: K build, factorisation
// with the tile inverses kept, whitening, T by block doubling, K^-1 = T^T T, U = [Q~; r] T, contraction against
// the regenerated dK tiles.  Workspace per entry: the matrix, three n x n matrices (Tlo, Tup, K^-1) and panels.
struct BatchGradWs {
    BatchWs v;
    size_t per_full, per_U, per_partial, per_entry;
};
static BatchGradWs batch_grad_ws(int n, int q, int d) {
    BatchGradWs g;
    g.v = batch_ws(n, q);
    g.per_full = align_up((size_t)n * g.v.lda * 8, 256);
    g.per_U = align_up((size_t)g.v.r * g.v.lda * 8, 256);
    g.per_partial = align_up(contract_workspace_bytes(n, n, d), 256);
    g.per_entry = g.v.per_A + 2 * g.v.per_T + g.v.per_W + g.v.per_mdev + 3 * g.per_full + g.per_U + g.per_partial;
    return g;
}

size_t gpmp_criterion_batched_grad_bytes(int n, int q, int d, int nbatch) {
    if (n <= 0 || q < 0 || d <= 0 || nbatch <= 0) return 0;
    BatchGradWs g = batch_grad_ws(n, q, d);
    return g.v.shared + (size_t)nbatch * g.per_entry;
}

int gpmp_criterion_batched_grad(const gpmp_cov_spec* spec, const double* theta_dev, int N, const double* x_dev,
                                long long x_stride, int n, const double* z_dev, long long z_stride,
                                const double* P_dev, int q, void* work_dev, size_t work_bytes, double* values_dev,
                                double* grads_dev, int* info_dev, void* stream) {
    if (!spec || !theta_dev || !x_dev || !z_dev || !work_dev || !values_dev || !grads_dev || !info_dev)
        return GPMP_ERR_ARG;
    if (n <= 0 || N < 0 || q < 0 || q > GPMP_MAX_Q || (q > 0 && !P_dev) || x_stride < 0 || z_stride < 0)
        return GPMP_ERR_ARG;
    if (N == 0) return GPMP_OK;
    cudaStream_t s = (cudaStream_t)stream;
    BatchGradWs gw = batch_grad_ws(n, q, spec->d);
    const BatchWs& w = gw.v;
    if (work_bytes < w.shared + gw.per_entry) return GPMP_ERR_WORKSPACE;
    long long cap = (long long)((work_bytes - w.shared) / gw.per_entry);
    if (cap > 32768) cap = 32768;
    if (cap > N) cap = N;
    char* base = static_cast<char*>(work_dev);
    size_t rowsb = align_up((size_t)(q > 0 ? q : 1) * w.lda * 8, 256);
    double* p0rows = (double*)base;
    double* p0work = (double*)(base + rowsb);
    double* ldr0 = (double*)(base + 2 * rowsb);
    char* pbase = base + w.shared;
    size_t off = 0;
    double* A = (double*)(pbase + off); off += (size_t)cap * w.per_A;
    double* Tlo_c = (double*)(pbase + off); off += (size_t)cap * w.per_T;
    double* Tup_c = (double*)(pbase + off); off += (size_t)cap * w.per_T;
    double* W = (double*)(pbase + off); off += (size_t)cap * w.per_W;
    MaternDev* mdev = (MaternDev*)(pbase + off); off += (size_t)cap * w.per_mdev;
    double* Tlo = (double*)(pbase + off); off += (size_t)cap * gw.per_full;
    double* Tup = (double*)(pbase + off); off += (size_t)cap * gw.per_full;
    double* Kinv = (double*)(pbase + off); off += (size_t)cap * gw.per_full;
    double* U = (double*)(pbase + off); off += (size_t)cap * gw.per_U;
    double* partial = (double*)(pbase + off);
    const long long sA = (long long)(w.per_A / 8), sT = (long long)(w.per_T / 8), sW = (long long)(w.per_W / 8);
    const long long sF = (long long)(gw.per_full / 8), sU = (long long)(gw.per_U / 8);
    int rc;
    if (cudaMemsetAsync(info_dev, 0, sizeof(int) * (size_t)N, s) != cudaSuccess) return GPMP_ERR_CUDA;
    const int w_theta = 1 + spec->noise + spec->d;
    for (long long c0 = 0; c0 < N; c0 += cap) {
        const int nb = (int)((N - c0) < cap ? (N - c0) : cap);
        const double* xb = x_dev + c0 * x_stride;
        rc = launch_prep_theta(spec, theta_dev + c0 * w_theta, nb, 1, mdev, s);
        if (rc) return rc;
        rc = launch_matern_cov(spec, mdev, nb, sA, xb, n, nullptr, n, A, w.lda, COV_SYM_LOWER, 0, s, x_stride);
        if (rc) return rc;
        LoadRowsArgs lr;
        lr.P = P_dev; lr.z = z_dev + c0 * z_stride; lr.strideZ = z_stride; lr.n = n; lr.q = q;
        lr.rows = A + (long long)n * w.lda; lr.ld = w.lda; lr.stride = sA;
        lr.p0rows = (q > 0 && c0 == 0) ? p0rows : nullptr; lr.ld0 = w.lda;
        rc = launch_load_rows(lr, nb, s);
        if (rc) return rc;
        // factorisation with the 128-wide tile inverses and the NB-wide block inverses (no Tsub: full path)
        rc = potrf_core(A, w.lda, sA, n, w.nrows, w.NB, Tlo_c, Tup_c, sT, W, sW, info_dev + c0, 1, nb, s, nullptr, 0);
        if (rc) return rc;
        FinalizeArgs f;
        f.rows = lr.rows; f.ld = w.lda; f.strideRows = sA;
        f.p0rows = p0rows; f.ld0 = w.lda; f.p0work = p0work; f.strideP0 = 0;
        f.Ldiag = A; f.ldl = w.lda; f.strideL = sA;
        f.n = n; f.q = q;
        f.Rt = nullptr; f.strideRt = 0;
        f.info = info_dev + c0; f.strideInfo = 1;
        f.out = values_dev + c0; f.strideOut = 1;
        f.ldr0_in = q > 0 ? ldr0 : nullptr;
        if (c0 == 0 && q > 0) {
            rc = launch_logdet_r0(p0rows, p0work, w.lda, n, q, ldr0, s);
            if (rc) return rc;
        }
        rc = launch_finalize(f, nb, s);
        if (rc) return rc;
        rc = potri_core(A, n, w.lda, w.NB, Tlo_c, Tup_c, Tlo, Tup, Kinv, w.lda, s, nb, sA, sT, sF);
        if (rc) return rc;
        URowsArgs u;
        u.R = lr.rows; u.ldr = w.lda; u.r = w.r; u.Tup = Tup; u.ldt = w.lda; u.U = U; u.ldu = w.lda;
        u.n = n; u.j0 = 0; u.j1 = 0; u.batch = nb; u.strideR = sA; u.strideT = sF; u.strideU = sU;
        rc = launch_urows(u, s);
        if (rc) return rc;
        ContractBatch cb{nb, mdev, sF, sU, x_stride};
        rc = launch_contract(spec, xb, n, nullptr, n, Kinv, w.lda, U, w.lda, w.r, 1, 0, 0.5,
                             grads_dev + c0 * w_theta, partial, (size_t)cap * gw.per_partial, s, 0, -1, &cb);
        if (rc) return rc;
    }
    return GPMP_OK;
}

}  // extern "C"
