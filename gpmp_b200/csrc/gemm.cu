// FP64 tensor-core (DMMA.8x8x4) "NT" GEMM:  C[MxN] = alpha * A[MxK] * B[NxK]^T + beta * C
//
// This one kernel is the flop engine of the whole path: SYRK trailing updates and TRSM-by-inverse
// of the right-looking Cholesky, the level GEMMs of the triangular inverse, LAUUM (K^-1 = T^T T),
// and the whitening/conditioning GEMMs of predict.  Every product on the path is arranged so that
// both operands are read with K contiguous (row-major "K-major" tiles), see DESIGN.md.
//
// Two tile shapes of the same template (warp tile 32x32 = 4x4 DMMA tiles, 64 accumulator registers):
//   default CTA 64x64, 4 warps as 2 x 2, 3-stage cp.async pipeline of 16-k chunks (48 KB), 4 CTAs per SM:
//           independent CTAs cover each other's barrier / fragment-load bubbles, and small products
//           (a handful of tiles inside a diagonal block) still spread over many SMs;
//   wide    CTA 128x128, 16 warps as 4 x 4, 3 stages x 2 chunks (192 KB), 1 CTA per SM: only for in-place
//           products, which need a single column tile per row block.
// Shared tiles are dense 128-byte rows with the 16-byte chunk index XOR-swizzled by (row & 7) -- the
// layout a TMA SWIZZLE_128B box produces.  Inside a K chunk the k index is permuted (lane kk owns
// k = 4*kk + s at MMA step s) so each lane fetches its 4 steps with two conflict-free LDS.128.
#include <atomic>
#include "common.cuh"

namespace gpmp {

constexpr int BK = 16;
constexpr int KGRAN = 128;  // granularity of the triangular K trimming (the tile grid of the operands)

struct GemmKArgs {
    GemmDesc g;
    int tiles_m, tiles_n;
};

__device__ __forceinline__ void decode_tile(const GemmKArgs& a, int t, int& ti, int& tj) {
    if (!a.g.lower) {
        ti = t / a.tiles_n;
        tj = t - ti * a.tiles_n;
        return;
    }
    // lower tiles, row-major over (ti, tj<=min(ti, tiles_n-1))
    long long tri = (long long)a.tiles_n * (a.tiles_n + 1) / 2;
    if (t < tri) {
        int r = (int)((sqrt(8.0 * (double)t + 1.0) - 1.0) * 0.5);
        while ((long long)(r + 1) * (r + 2) / 2 <= t) ++r;
        while ((long long)r * (r + 1) / 2 > t) --r;
        ti = r;
        tj = t - (int)((long long)r * (r + 1) / 2);
    } else {
        int u = t - (int)tri;
        ti = a.tiles_n + u / a.tiles_n;
        tj = u % a.tiles_n;
    }
}

// ROWS x 16 doubles -> swizzled smem tile; thread -> (row = tid/8 + (THREADS/8)*i, chunk = tid%8)
template <int ROWS, int THREADS>
__device__ __forceinline__ void load_tile(uint32_t sdst, const double* __restrict__ src, long long ld,
                                          int row0, int nrows, int k, int k1, int tid) {
    const int c = tid & 7;
    const int kk = k + 2 * c;
    const int rem = k1 - kk;
    const int nb = rem >= 2 ? 16 : (rem == 1 ? 8 : 0);
#pragma unroll
    for (int i = 0; i < ROWS * 8 / THREADS; ++i) {
        const int r = (tid >> 3) + (THREADS / 8) * i;
        const int grow = row0 + r;
        const int bytes = (grow < nrows) ? nb : 0;
        const double* p = bytes ? (src + (long long)grow * ld + kk) : src;
        const uint32_t d = sdst + r * 128 + ((c ^ (r & 7)) << 4);
        cp_async16(d, p, bytes);
    }
}

template <int WM, int WN, int MI, int NI, int STAGES, int SUBK>
struct GemmCfg {
    static constexpr int BM = WM * MI * 8, BN = WN * NI * 8, THREADS = WM * WN * 32;
    static constexpr int A_BYTES = BM * BK * 8, B_BYTES = BN * BK * 8;
    static constexpr int SUB_BYTES = A_BYTES + B_BYTES;
    static constexpr int STAGE_BYTES = SUBK * SUB_BYTES;
    static constexpr int SMEM = STAGES * STAGE_BYTES;
};

// One output tile: tile index t of the launch's tile list (already reversed if the launch asks for it), second /
// first batch level yb / zb.  Every thread of the CTA calls it; shared memory may be reused by the next call after
// a CTA-wide barrier.
template <int WM, int WN, int MI, int NI, int STAGES, int SUBK>
__device__ __forceinline__ void gemm_tile(const GemmKArgs& a, int t, long long yb, long long zb, unsigned char* smem) {
    using Cfg = GemmCfg<WM, WN, MI, NI, STAGES, SUBK>;
    constexpr int BM = Cfg::BM, BN = Cfg::BN, THREADS = Cfg::THREADS;
    const GemmDesc& g = a.g;
    int ti, tj;
    decode_tile(a, t, ti, tj);
    const int m0 = ti * BM, n0 = tj * BN;
    const double* __restrict__ A = g.A + zb * g.strideA + yb * g.stride2A;
    const double* __restrict__ B = g.B + zb * g.strideB + yb * g.stride2B;
    double* __restrict__ C = g.C + zb * g.strideC + yb * g.stride2C;

    // K trimming for triangular operands (granularity KGRAN = the operands' 128-tile grid; the zero part
    // of a triangular operand may hold anything outside its diagonal tiles -- e.g. the mirrored other
    // triangle -- so trimming is also what makes the product correct).
    int k0 = 0, k1 = g.K;
    const int mt = m0 + g.ktrim_off, nt = n0 + g.ktrim_off;
    if (g.krange == KR_FROM_ROW) k0 = min(mt / KGRAN * KGRAN, g.K);
    else if (g.krange == KR_TO_ROW) k1 = min(g.K, (mt / KGRAN + 1) * KGRAN);
    else if (g.krange == KR_FROM_COL) k0 = min(nt / KGRAN * KGRAN, g.K);
    else if (g.krange == KR_TO_COL) k1 = min(g.K, (nt / KGRAN + 1) * KGRAN);
    const int nk = k1 > k0 ? (k1 - k0 + SUBK * BK - 1) / (SUBK * BK) : 0;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp / WN, wn = warp % WN;
    const int gq = lane >> 2, kk = lane & 3;  // fragment row/col group, k lane
    const uint32_t sbase = smem_u32(smem);

    double acc[MI][NI][2];
#pragma unroll
    for (int i = 0; i < MI; ++i)
#pragma unroll
        for (int j = 0; j < NI; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    auto load_stage = [&](int slot, int chunk) {
        const uint32_t st = sbase + slot * Cfg::STAGE_BYTES;
#pragma unroll
        for (int u = 0; u < SUBK; ++u) {
            const int kc = k0 + (chunk * SUBK + u) * BK;
            load_tile<BM, THREADS>(st + u * Cfg::SUB_BYTES, A, g.lda, m0, g.M, kc, k1, tid);
            load_tile<BN, THREADS>(st + u * Cfg::SUB_BYTES + Cfg::A_BYTES, B, g.ldb, n0, g.N, kc, k1, tid);
        }
    };

    // prologue
#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) {
        if (s < nk) load_stage(s, s);
        cp_async_commit();
    }

    // per-thread swizzled fragment offsets (row & 7 == gq for every fragment row of this lane)
    const uint32_t aoff = (wm * MI * 8 + gq) * 128;
    const uint32_t boff = Cfg::A_BYTES + (wn * NI * 8 + gq) * 128;
    const uint32_t c0 = ((2 * kk) ^ gq) << 4, c1 = ((2 * kk + 1) ^ gq) << 4;

    for (int kt = 0; kt < nk; ++kt) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        {
            const int nx = kt + STAGES - 1;
            if (nx < nk) load_stage(nx % STAGES, nx);
            cp_async_commit();
        }
#pragma unroll
        for (int u = 0; u < SUBK; ++u) {
            const unsigned char* st = smem + (kt % STAGES) * Cfg::STAGE_BYTES + u * Cfg::SUB_BYTES;
            double af[MI][4];
#pragma unroll
            for (int mi = 0; mi < MI; ++mi) {
                const double2 v0 = *reinterpret_cast<const double2*>(st + aoff + mi * 1024 + c0);
                const double2 v1 = *reinterpret_cast<const double2*>(st + aoff + mi * 1024 + c1);
                af[mi][0] = v0.x; af[mi][1] = v0.y; af[mi][2] = v1.x; af[mi][3] = v1.y;
            }
            double bf[NI][4];
#pragma unroll
            for (int ni = 0; ni < NI; ++ni) {
                const double2 w0 = *reinterpret_cast<const double2*>(st + boff + ni * 1024 + c0);
                const double2 w1 = *reinterpret_cast<const double2*>(st + boff + ni * 1024 + c1);
                bf[ni][0] = w0.x; bf[ni][1] = w0.y; bf[ni][2] = w1.x; bf[ni][3] = w1.y;
            }
            // k-step outermost: an accumulator is revisited only after MI*NI other DMMAs, so the pipe never
            // waits on its own result
#pragma unroll
            for (int s = 0; s < 4; ++s)
#pragma unroll
                for (int ni = 0; ni < NI; ++ni)
#pragma unroll
                    for (int mi = 0; mi < MI; ++mi)
                        dmma884(acc[mi][ni][0], acc[mi][ni][1], af[mi][s], bf[ni][s]);
        }
    }
    cp_async_wait<0>();

    // epilogue
    const double alpha = g.alpha, beta = g.beta;
    double* __restrict__ Ct = g.Ct ? g.Ct + zb * g.strideCt + yb * g.stride2Ct : nullptr;
#pragma unroll
    for (int mi = 0; mi < MI; ++mi) {
        const int row = m0 + wm * MI * 8 + mi * 8 + gq;
        if (row >= g.M) continue;
#pragma unroll
        for (int ni = 0; ni < NI; ++ni) {
            const int col = n0 + wn * NI * 8 + ni * 8 + 2 * kk;
            if (col >= g.N) continue;
            double* cp = C + (long long)row * g.ldc + col;
            double v0 = alpha * acc[mi][ni][0], v1 = alpha * acc[mi][ni][1];
            if (col + 1 < g.N) {
                if (beta != 0.0) {
                    const double2 old = *reinterpret_cast<const double2*>(cp);
                    v0 += beta * old.x;
                    v1 += beta * old.y;
                }
                *reinterpret_cast<double2*>(cp) = make_double2(v0, v1);
                if (Ct) {
                    Ct[(long long)col * g.ldct + row] = v0;
                    Ct[(long long)(col + 1) * g.ldct + row] = v1;
                }
            } else {
                if (beta != 0.0) v0 += beta * cp[0];
                cp[0] = v0;
                if (Ct) Ct[(long long)col * g.ldct + row] = v0;
            }
        }
    }
}

template <int WM, int WN, int MI, int NI, int MINB, int STAGES, int SUBK>
__global__ void __launch_bounds__(WM* WN * 32, MINB) gemm_nt_kernel(const GemmKArgs a) {
    extern __shared__ __align__(1024) unsigned char smem[];
    const int t = a.g.reverse ? (int)(gridDim.x - 1 - blockIdx.x) : (int)blockIdx.x;
    gemm_tile<WM, WN, MI, NI, STAGES, SUBK>(a, t, blockIdx.y, blockIdx.z, smem);
}

// Persistent form for products that must leave part of the machine to other work (the triangular-inverse levels
// that run under the tail of a factorisation, potrf.cu): CTAs that find themselves on an SM below `sm_first` exit at
// once; the others draw (tile, pair) indices from a counter in global memory until the list is exhausted.  If half
// of the grid has arrived and still nobody has drawn a tile (other kernels keep the admitted SMs full while the
// avoided ones are free, so the scheduler sends the whole grid there to exit), the CTAs that arrive from then on
// work wherever they land, and the last CTA to leave finishes whatever is left and resets the counters for their
// next user: progress and correctness never depend on placement.  (Measured: letting the first 16 CTAs work
// unconditionally instead costs the headline 0.1-0.25 ms -- they land exactly on the SMs kept free for the chain.)
__device__ unsigned int g_gemm_counters[3 * 1024];
template <int WM, int WN, int MI, int NI, int MINB, int STAGES, int SUBK>
__global__ void __launch_bounds__(WM* WN * 32, MINB)
gemm_nt_persist_kernel(const GemmKArgs a, int ntiles, int slot, int sm_first) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ unsigned int next;
    unsigned int* ctr = g_gemm_counters + 3 * slot;
    unsigned int smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    const unsigned int total = (unsigned int)ntiles * (unsigned int)a.g.batch2;
    __shared__ unsigned int admitted;  // decided by ONE thread: the whole CTA must take the same branch
    if (threadIdx.x == 0) {
        const unsigned int order = atomicAdd(&ctr[2], 1u);  // arrival order of this CTA
        const bool stranded = order >= gridDim.x / 2 && *reinterpret_cast<volatile unsigned int*>(&ctr[0]) == 0u;
        admitted = ((int)smid >= sm_first || stranded) ? 1u : 0u;
    }
    __syncthreads();
    if (admitted) {
        for (;;) {
            __syncthreads();  // the previous tile's shared memory and `next` are no longer read
            if (threadIdx.x == 0) next = atomicAdd(&ctr[0], 1u);
            __syncthreads();
            const unsigned int w = next;
            if (w >= total) break;
            const int yb = (int)(w / (unsigned int)ntiles), t0 = (int)(w - (unsigned int)yb * (unsigned int)ntiles);
            gemm_tile<WM, WN, MI, NI, STAGES, SUBK>(a, a.g.reverse ? ntiles - 1 - t0 : t0, yb, 0, smem);
        }
    }
    // Leaving.  The last CTA to leave finishes whatever is left of the list itself, wherever it runs (only if every
    // other CTA was turned away from its SM -- correctness must not depend on where the scheduler places CTAs),
    // and resets the counter pair.
    __shared__ unsigned int last;
    if (threadIdx.x == 0) {
        __threadfence();
        last = atomicAdd(&ctr[1], 1u) == gridDim.x - 1 ? 1u : 0u;
    }
    __syncthreads();
    if (!last) return;
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) next = atomicAdd(&ctr[0], 1u);
        __syncthreads();
        const unsigned int w = next;
        if (w >= total) break;
        const int yb = (int)(w / (unsigned int)ntiles), t0 = (int)(w - (unsigned int)yb * (unsigned int)ntiles);
        gemm_tile<WM, WN, MI, NI, STAGES, SUBK>(a, a.g.reverse ? ntiles - 1 - t0 : t0, yb, 0, smem);
    }
    if (threadIdx.x == 0) {
        ctr[0] = 0u;
        ctr[1] = 0u;
        ctr[2] = 0u;
        __threadfence();
    }
}

template <int WM, int WN, int MI, int NI, int MINB, int STAGES = 3, int SUBK = 2>
static int launch_cfg(const GemmDesc& g, cudaStream_t stream) {
    using Cfg = GemmCfg<WM, WN, MI, NI, STAGES, SUBK>;
    static unsigned long long configured = 0;  // one bit per device: the attribute is per context
    auto kern = gemm_nt_kernel<WM, WN, MI, NI, MINB, STAGES, SUBK>;
    int dev = 0;
    cudaGetDevice(&dev);
    if (!((configured >> (dev & 63)) & 1ull)) {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM) != cudaSuccess)
            return GPMP_ERR_CUDA;
        configured |= 1ull << (dev & 63);
    }
    GemmKArgs a;
    a.g = g;
    a.tiles_m = ceil_div(g.M, Cfg::BM);
    a.tiles_n = ceil_div(g.N, Cfg::BN);
    long long ntiles;
    if (g.lower) {
        if (a.tiles_n > a.tiles_m) a.tiles_n = a.tiles_m;
        ntiles = (long long)a.tiles_n * (a.tiles_n + 1) / 2 + (long long)(a.tiles_m - a.tiles_n) * a.tiles_n;
    } else {
        ntiles = (long long)a.tiles_m * a.tiles_n;
    }
    const double kavg = g.krange == KR_FULL ? (double)g.K : 0.5 * (double)g.K;
    const double work = 2.0 * (double)ntiles * Cfg::BM * Cfg::BN * kavg * g.batch * g.batch2;
    LaunchScope scope(KC_GEMM, work, stream);
    dim3 grid((unsigned)ntiles, (unsigned)g.batch2, (unsigned)g.batch);
    kern<<<grid, Cfg::THREADS, Cfg::SMEM, stream>>>(a);
    GPMP_CHECK_LAUNCH();
    return GPMP_OK;
}

int launch_gemm_nt_persist(const GemmDesc& g, cudaStream_t stream, int sm_first) {
    if (g.M <= 0 || g.N <= 0 || g.batch2 <= 0) return GPMP_OK;
    if (g.batch != 1) return GPMP_ERR_ARG;
    if ((g.lda & 1) || (g.ldb & 1) || (g.ldc & 1) || (g.stride2A & 1) || (g.stride2B & 1) || (g.stride2C & 1))
        return GPMP_ERR_ALIGN;
    using Cfg = GemmCfg<2, 2, 4, 4, 3, 1>;
    auto kern = gemm_nt_persist_kernel<2, 2, 4, 4, 4, 3, 1>;
    static unsigned long long configured = 0;
    static std::atomic<unsigned int> next_slot{0};  // (two host threads must never draw the same counter pair)
    int dev = 0;
    cudaGetDevice(&dev);
    if (!((configured >> (dev & 63)) & 1ull)) {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM) != cudaSuccess)
            return GPMP_ERR_CUDA;
        configured |= 1ull << (dev & 63);
    }
    GemmKArgs a;
    a.g = g;
    a.tiles_m = ceil_div(g.M, Cfg::BM);
    a.tiles_n = ceil_div(g.N, Cfg::BN);
    long long ntiles;
    if (g.lower) {
        if (a.tiles_n > a.tiles_m) a.tiles_n = a.tiles_m;
        ntiles = (long long)a.tiles_n * (a.tiles_n + 1) / 2 + (long long)(a.tiles_m - a.tiles_n) * a.tiles_n;
    } else {
        ntiles = (long long)a.tiles_m * a.tiles_n;
    }
    const long long total = ntiles * g.batch2;
    const double kavg = g.krange == KR_FROM_ROW || g.krange == KR_TO_ROW || g.krange == KR_FROM_COL ||
                                g.krange == KR_TO_COL
                            ? 0.5 * (double)g.K
                            : (double)g.K;
    LaunchScope scope(KC_GEMM, 2.0 * (double)total * Cfg::BM * Cfg::BN * kavg, stream);
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const unsigned int grid = (unsigned int)(total < 4LL * sms ? total : 4LL * sms);
    const int slot = (int)(next_slot.fetch_add(1u) & 1023u);  // (a slot is reused 1024 persistent launches later)
    kern<<<grid, Cfg::THREADS, Cfg::SMEM, stream>>>(a, (int)ntiles, slot, sm_first);
    GPMP_CHECK_LAUNCH();
    return GPMP_OK;
}

int launch_gemm_nt(const GemmDesc& g, cudaStream_t stream) {
    if (g.M <= 0 || g.N <= 0 || g.batch <= 0 || g.batch2 <= 0) return GPMP_OK;
    // 16-byte cp.async / vector epilogue requirements
    if ((g.lda & 1) || (g.ldb & 1) || (g.ldc & 1)) return GPMP_ERR_ALIGN;
    if ((reinterpret_cast<uintptr_t>(g.A) & 15) || (reinterpret_cast<uintptr_t>(g.B) & 15) ||
        (reinterpret_cast<uintptr_t>(g.C) & 15))
        return GPMP_ERR_ALIGN;
    if ((g.strideA & 1) || (g.strideB & 1) || (g.strideC & 1)) return GPMP_ERR_ALIGN;
    if ((g.stride2A & 1) || (g.stride2B & 1) || (g.stride2C & 1)) return GPMP_ERR_ALIGN;
    // Shape choice (measured on B200, TFLOP/s at K = 128 / 512 / 8192):
    //   64x64 tiles, 4 warps, 3 stages x 16 k (48 KB), 4 CTAs/SM      24.1 / 30.9 / 32.8   <- default
    //   64x64 tiles, 4 warps, 4 stages x 16 k (64 KB), 3 CTAs/SM      21.3 / 30.5 / 32.9
    //   128x128 tiles, 16 warps, 3 stages x 32 k (192 KB), 1 CTA/SM   18.5 / 25.8 / 32.0
    //   32x64 / 64x32 / 32x32 tiles, 4 warps, 6-8 CTAs/SM (8192^3)     29.0 / 30.3 / 25.8
    // Several independent CTAs per SM hide each other's barrier and fragment-load bubbles, which a single
    // big CTA cannot (the DMMA pipe sat at 85 % with 1 CTA/SM).  An in-place product (the single-column-tile
    // panel solves, A == C) must keep one column tile per row block, so N > 64 takes the 128-wide shape.
    const bool in_place = (g.A == g.C || g.B == g.C);
    if (in_place && g.N > 64) return launch_cfg<4, 4, 4, 4, 1>(g, stream);
    // products that fill the machine with 128 x 128 tiles: the TMA / mbarrier kernel (gemm_tma.cu)
    {
        const int rc = launch_gemm_nt_tma(g, stream);
        if (rc <= 0) return rc;
    }
    // Latency shapes for launches that cannot fill the machine (the chain's K=128 updates in the tail of a
    // factorisation): a lone 64x64 tile costs 5 us + 0.8 us per 16-wide k step because one warp per scheduler
    // partition issues all its DMMAs; 32-row / 32x32 tiles spread the same work over 2x / 4x the SMs.
    if (!in_place && g.batch == 1 && g.batch2 == 1) {
        const long long tm = ceil_div(g.M, 64), tn = ceil_div(g.N, 64);
        const long long t64 = g.lower ? (tn < tm ? tn * (tn + 1) / 2 + (tm - tn) * tn : tm * (tm + 1) / 2) : tm * tn;
        if (t64 <= (g.lower ? 296 : 148)) return launch_cfg<2, 2, 2, 2, 8, 3, 1>(g, stream);
        if (t64 <= 296 && !g.lower) return launch_cfg<2, 2, 2, 4, 6, 3, 1>(g, stream);  // (lower lists need square tiles)
    }
    return launch_cfg<2, 2, 4, 4, 4, 3, 1>(g, stream);
}

}  // namespace gpmp
