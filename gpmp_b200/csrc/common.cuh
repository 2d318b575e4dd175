// Shared device/host helpers for the gpmp_b200 sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/gpmp_b200.h"

namespace gpmp {

// ---------------------------------------------------------------------------------------------
// Launch accounting + optional per-class CUDA-event profiling (bench.py reads these through the
// C-ABI: gpmp_launch_count / gpmp_prof_*).  Profiling brackets each launch of a class with events
// on the launching stream; it is off by default and never synchronises by itself.
// ---------------------------------------------------------------------------------------------
enum KernelClass {
    KC_MATERN = 0,    // K1 covariance build (bytes)
    KC_GEMM = 1,      // DMMA NT GEMM / SYRK / TRSM-as-GEMM (flops)
    KC_POTF2 = 2,     // diagonal-block factor+invert (flops)
    KC_CONTRACT = 3,  // K4 dK contraction (bytes)
    KC_SMALL = 4,     // reductions, QR of whitened rows, trmv, misc
    KC_BATCHED = 5,   // batched small-n criterion
    KC_COUNT = 6
};

struct ProfState {
    int enabled;
    unsigned long long launches[KC_COUNT];
    double work[KC_COUNT];  // algorithmic flops or bytes accumulated per class
};
ProfState& prof();
void prof_begin(int cls, cudaStream_t s);
void prof_end(int cls, double work, cudaStream_t s);

struct LaunchScope {
    int cls;
    double work;
    cudaStream_t s;
    LaunchScope(int c, double w, cudaStream_t st) : cls(c), work(w), s(st) { prof_begin(c, st); }
    ~LaunchScope() { prof_end(cls, work, s); }
};

#define GPMP_CHECK_LAUNCH()                                   \
    do {                                                      \
        cudaError_t e__ = cudaGetLastError();                 \
        if (e__ != cudaSuccess) return GPMP_ERR_CUDA;         \
    } while (0)

// ---------------------------------------------------------------------------------------------
// Device primitives
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    // native FP64 tensor instruction on sm_100a: SASS DMMA.8x8x4
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// 16-byte async copy global->shared with zero-fill: copies src_bytes (0..16), zero-fills the rest.
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, int src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N));
}

// 1/sqrt(a) in double: the hardware seed rsqrt.approx.ftz.f64 (MUFU.RSQ64H, ~2^-23 relative, full double
// range) and two Newton steps -- branch-free, ~10 instructions (the library rsqrt()/sqrt() are long
// software sequences with slow paths).
__device__ __forceinline__ double fast_rsqrt(double a) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
    const double h = 0.5 * a;
    y = y * fma(-h * y, y, 1.5);
    y = y * fma(-h * y, y, 1.5);
    return y;
}
// sqrt(a) for a >= 0 to ~2 ulp (a * rsqrt(a)); exact zero preserved, NaN / negative give NaN
__device__ __forceinline__ double fast_sqrt(double a) {
    const double r = a * fast_rsqrt(a);
    return a == 0.0 ? 0.0 : r;
}

// exp(-t) for t >= 0 to ~2 ulp with a short straight-line sequence: k = round(t log2 e), r = k ln2 - t in
// [-0.347, 0.347], degree-12 Taylor polynomial of e^r (remainder 1.7e-16 relative), scaling by 2^-k through
// the exponent field.  Results below 2^-1021 flush to zero; NaN propagates.  Branch-free (selects only).
__device__ __forceinline__ double exp_neg(double t) {
    const double tc = fmin(t, 707.0);
    const double kf = rint(tc * 1.4426950408889634);
    double r = fma(kf, 6.93147180369123816490e-01, -tc);
    r = fma(kf, 1.90821492927058770002e-10, r);
    double p = 2.08767569878681e-09;
    p = fma(p, r, 2.505210838544172e-08);
    p = fma(p, r, 2.755731922398589e-07);
    p = fma(p, r, 2.7557319223985893e-06);
    p = fma(p, r, 2.48015873015873e-05);
    p = fma(p, r, 0.0001984126984126984);
    p = fma(p, r, 0.001388888888888889);
    p = fma(p, r, 0.008333333333333333);
    p = fma(p, r, 0.041666666666666664);
    p = fma(p, r, 0.16666666666666666);
    p = fma(p, r, 0.5);
    p = fma(p, r, 1.0);
    p = fma(p, r, 1.0);
    const int k = (int)kf;
    double res = p * __longlong_as_double((long long)(1023 - k) << 52);
    res = t < 707.0 ? res : 0.0;
    return t != t ? t : res;
}

// ---- short forms for the covariance tile kernels (K1 / K4): fewer FP64 operations and no conversion or
// select chains (those kernels are bound by instruction issue, not by HBM) ------------------------------------
// sqrt(a), a >= 0: one third-order step from the hardware seed y ~ a^-1/2 (relative error 2^-23):
//   s = a y,  e = 1 - s y,  sqrt(a) = s (1 - e)^-1/2 = s (1 + e/2 + 3 e^2/8 + O(e^3)),  O(e^3) < 1e-20.
// The seed is taken at a + 1e-300 so that a = 0 gives s = 0 and an exact 0 without a test; NaN propagates.
__device__ __forceinline__ double sqrt_seeded(double a) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a + 1e-300));
    const double s = a * y;
    const double e = fma(-s, y, 1.0);
    const double t = fma(e, 0.375, 0.5) * e;
    return fma(s, t, s);
}

// exp(-t) for 0 <= t <= 707 (the caller clamps), ~2 ulp:  K = round(32 t / ln 2) by the add-magic-number trick
// (the integer lands in the low word of the sum: no FRND / F2I), r = K ln2/32 - t in [-ln2/64, ln2/64],
// exp(-t) = 2^-(K>>5) * tab[K & 31] * exp(r) with tab[j] = 2^(-j/32) (32 doubles, passed in shared memory) and a
// degree-6 Taylor polynomial (remainder 3.5e-18); the power of two goes straight into the exponent field.
// A NaN argument gives an arbitrary finite result here: callers that can see NaN multiply by a NaN polynomial.
constexpr int EXP_TAB = 32;
__device__ __forceinline__ double exp_neg_tab(double t, const double* __restrict__ tab) {
    const double d = fma(t, 46.16624130844683, 6755399441055744.0);  // 32 / ln 2 ; 1.5 * 2^52
    const int K = __double2loint(d);
    const double kf = d - 6755399441055744.0;
    double r = fma(kf, 0.021660849392446835, -t);  // ln2/32, high 37 bits (K * hi is exact)
    r = fma(kf, 5.145609244655338e-14, r);
    double p = 1.3888888888888889e-03;
    p = fma(p, r, 8.3333333333333332e-03);
    p = fma(p, r, 4.1666666666666664e-02);
    p = fma(p, r, 1.6666666666666666e-01);
    p = fma(p, r, 0.5);
    p = fma(p, r, 1.0);
    p = fma(p, r, 1.0);
    const double e = tab[K & (EXP_TAB - 1)] * p;  // in (0.49, 1.02]
    return __hiloint2double(__double2hiint(e) - ((K >> 5) << 20), __double2loint(e));
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Block-wide sum; result valid in thread 0 (and broadcast through smem to all). blockDim.x <= 1024.
__device__ __forceinline__ double block_sum(double v, double* red /* >= 33 doubles of smem */) {
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) red[w] = v;
    __syncthreads();
    if (w == 0) {
        double t = lane < nw ? red[lane] : 0.0;
        t = warp_sum(t);
        if (lane == 0) red[32] = t;
    }
    __syncthreads();
    return red[32];
}

__host__ __device__ static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
static inline long long ceil_div_ll(long long a, long long b) { return (a + b - 1) / b; }

// ---------------------------------------------------------------------------------------------
// Internal kernel-launch entry points shared across translation units
// ---------------------------------------------------------------------------------------------
enum KRange { KR_FULL = 0, KR_FROM_ROW = 1, KR_TO_ROW = 2, KR_FROM_COL = 3, KR_TO_COL = 4 };

struct GemmDesc {
    // C[M x N] = alpha * A[M x K] * B[N x K]^T + beta * C   (all row-major, "NT": K contiguous)
    const double* A; long long lda; long long strideA;
    const double* B; long long ldb; long long strideB;
    double* C;       long long ldc; long long strideC;
    double* Ct;      long long ldct; long long strideCt;  // optional mirror: Ct[j][i] = C[i][j]
    int M, N, K;
    double alpha, beta;
    int lower;   // 1: only tiles with tile_col <= tile_row (tile indices taken on the same origin)
    int krange;  // KRange: trims the K loop at tile granularity for triangular operands
    int batch;   // blockIdx.z: strideA/B/C/Ct (elements)
    // second batch level (blockIdx.y), e.g. the block pairs of one level of the triangular inverse
    int batch2;
    long long stride2A, stride2B, stride2C, stride2Ct;
    int reverse;  // visit tiles in reverse order (longest-K tiles first for KR_TO_ROW)
    int ktrim_off;  // added to the tile origin before trimming: the operand is a row / column block of a larger
                    // triangular matrix that starts ktrim_off rows (columns) further down
};
inline GemmDesc gemm_desc() {
    GemmDesc g{};
    g.alpha = 1.0;
    g.batch = 1;
    g.batch2 = 1;
    return g;
}
int launch_gemm_nt(const GemmDesc& g, cudaStream_t stream);
// gemm.cu: the same product by persistent CTAs that keep off the SMs below sm_first (batch == 1 only)
int launch_gemm_nt_persist(const GemmDesc& g, cudaStream_t stream, int sm_first);
// gemm_tma.cu: 0 launched, 1 not a shape / environment it serves (take the cp.async kernel), < 0 error
int launch_gemm_nt_tma(const GemmDesc& g, cudaStream_t stream);

}  // namespace gpmp
