// TMA-fed variant of the FP64 DMMA "NT" GEMM (gemm.cu) for the products that fill the machine: the K = NB trailing
// updates of the factorisation, the level products of the triangular inverse and K^-1 = T^T T.
//
//   CTA tile 128 x 128, 16 warps as 4 x 4 (warp tile 32 x 32 = 4 x 4 DMMA.8x8x4, the same fragment code as
//   gemm.cu), one CTA per SM.  Operand chunks of 16 k (128-byte rows) are fetched by cp.async.bulk.tensor.2d
//   (TMA, CU_TENSOR_MAP_SWIZZLE_128B: the hardware writes exactly the XOR-swizzled 128-byte-row layout the
//   fragment loads of gemm.cu expect) into a ring of TSTAGES stages, each with a "full" mbarrier (armed with the
//   stage's byte count, completed by the TMA) and an "empty" mbarrier (one arrival per warp).  There is no
//   CTA-wide barrier in the main loop: a warp waits only for the bytes of its next chunk, so the 16 warps drift
//   apart by up to the pipeline depth and cover each other's fragment-load bubbles the way the four independent
//   64 x 64 CTAs of gemm.cu do -- with half the operand traffic per flop (128-wide tiles) and no per-thread
//   address arithmetic or cp.async issue in the loop.  One thread re-arms a stage and issues its two TMA loads
//   one iteration after the stage was consumed, when its release has normally already happened.
//   Out-of-range rows / k are zero-filled by the TMA unit, so ragged M, N, K need no predicates on the load side.
#include <cuda.h>
#include <cstdlib>
#include "common.cuh"

namespace gpmp {

constexpr int TBM = 128, TBN = 128, TBK = 16;
constexpr int TSTAGES = 6;
constexpr int T_A_BYTES = TBM * TBK * 8, T_B_BYTES = TBN * TBK * 8;
constexpr int T_STAGE_BYTES = T_A_BYTES + T_B_BYTES;
constexpr int T_THREADS = 512;
constexpr int T_SMEM = TSTAGES * T_STAGE_BYTES + 2 * TSTAGES * 8 + 1024;  // + barriers + alignment slack
constexpr int T_KGRAN = 128;

struct GemmTmaArgs {
    GemmDesc g;
    int tiles_m, tiles_n;
};

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arm(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(bar), "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
        "l"(map), "r"(c0), "r"(c1), "r"(bar)
        : "memory");
}

__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(dst),
        "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(bar)
        : "memory");
}

__device__ __forceinline__ void decode_tile_sq(const GemmTmaArgs& a, int t, int& ti, int& tj) {
    if (!a.g.lower) {
        ti = t / a.tiles_n;
        tj = t - ti * a.tiles_n;
        return;
    }
    const long long tri = (long long)a.tiles_n * (a.tiles_n + 1) / 2;
    if (t < tri) {
        int r = (int)((sqrt(8.0 * (double)t + 1.0) - 1.0) * 0.5);
        while ((long long)(r + 1) * (r + 2) / 2 <= t) ++r;
        while ((long long)r * (r + 1) / 2 > t) --r;
        ti = r;
        tj = t - (int)((long long)r * (r + 1) / 2);
    } else {
        const int u = t - (int)tri;
        ti = a.tiles_n + u / a.tiles_n;
        tj = u % a.tiles_n;
    }
}

__global__ void __launch_bounds__(T_THREADS, 1)
gemm_nt_tma_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                   const GemmTmaArgs a) {
    extern __shared__ unsigned char smem_raw[];
    // SWIZZLE_128B needs 1024-byte aligned tiles
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const GemmDesc& g = a.g;
    int ti, tj;
    decode_tile_sq(a, g.reverse ? (int)(gridDim.x - 1 - blockIdx.x) : (int)blockIdx.x, ti, tj);
    const int m0 = ti * TBM, n0 = tj * TBN;

    int k0 = 0, k1 = g.K;
    const int mt = m0 + g.ktrim_off, nt = n0 + g.ktrim_off;
    if (g.krange == KR_FROM_ROW) k0 = min(mt / T_KGRAN * T_KGRAN, g.K);
    else if (g.krange == KR_TO_ROW) k1 = min(g.K, (mt / T_KGRAN + 1) * T_KGRAN);
    else if (g.krange == KR_FROM_COL) k0 = min(nt / T_KGRAN * T_KGRAN, g.K);
    else if (g.krange == KR_TO_COL) k1 = min(g.K, (nt / T_KGRAN + 1) * T_KGRAN);
    const int nk = k1 > k0 ? (k1 - k0 + TBK - 1) / TBK : 0;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp >> 2, wn = warp & 3;
    const int gq = lane >> 2, kk = lane & 3;
    const uint32_t sbase = smem_u32(smem);
    const uint32_t bar_full = sbase + TSTAGES * T_STAGE_BYTES;
    const uint32_t bar_empty = bar_full + TSTAGES * 8;

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < TSTAGES; ++s) {
            mbar_init(bar_full + 8 * s, 1);
            mbar_init(bar_empty + 8 * s, T_THREADS / 32);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();

    auto issue = [&](int chunk) {
        const int s = chunk % TSTAGES;
        const uint32_t st = sbase + s * T_STAGE_BYTES;
        const uint32_t fb = bar_full + 8 * s;
        mbar_arm(fb, T_STAGE_BYTES);
        const int kc = k0 + chunk * TBK;
        if (g.batch2 > 1) {
            // second batch level (the block pairs of one level of the triangular inverse): third map coordinate
            tma_load_3d(st, &mapA, kc, m0, (int)blockIdx.y, fb);
            tma_load_3d(st + T_A_BYTES, &mapB, kc, n0, (int)blockIdx.y, fb);
        } else {
            tma_load_2d(st, &mapA, kc, m0, fb);
            tma_load_2d(st + T_A_BYTES, &mapB, kc, n0, fb);
        }
    };
    if (tid == 0) {
        const int pre = nk < TSTAGES ? nk : TSTAGES;
        for (int c = 0; c < pre; ++c) issue(c);
    }

    double acc[4][4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    const uint32_t aoff = (wm * 32 + gq) * 128;
    const uint32_t boff = T_A_BYTES + (wn * 32 + gq) * 128;
    const uint32_t c0 = ((2 * kk) ^ gq) << 4, c1 = ((2 * kk + 1) ^ gq) << 4;

    for (int kt = 0; kt < nk; ++kt) {
        const int s = kt % TSTAGES;
        mbar_wait(bar_full + 8 * s, (kt / TSTAGES) & 1);
        const unsigned char* st = smem + s * T_STAGE_BYTES;
        // two half-chunks of 8 k: 16 fragment registers each for A and B beside the 64 accumulators
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const uint32_t ch = h == 0 ? c0 : c1;
            double af[4][2], bf[4][2];
#pragma unroll
            for (int mi = 0; mi < 4; ++mi) {
                const double2 v = *reinterpret_cast<const double2*>(st + aoff + mi * 1024 + ch);
                af[mi][0] = v.x; af[mi][1] = v.y;
            }
#pragma unroll
            for (int ni = 0; ni < 4; ++ni) {
                const double2 w = *reinterpret_cast<const double2*>(st + boff + ni * 1024 + ch);
                bf[ni][0] = w.x; bf[ni][1] = w.y;
            }
            if (h == 1) {
                // the stage is free for this warp as soon as its last fragments are in registers
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_empty + 8 * s);
            }
#pragma unroll
            for (int q = 0; q < 2; ++q)
#pragma unroll
                for (int ni = 0; ni < 4; ++ni)
#pragma unroll
                    for (int mi = 0; mi < 4; ++mi) dmma884(acc[mi][ni][0], acc[mi][ni][1], af[mi][q], bf[ni][q]);
        }
        // refill, one iteration late: the stage of chunk kt - 1 has normally been released by every warp by now
        if (tid == 0 && kt >= 1 && kt - 1 + TSTAGES < nk) {
            const int sp = (kt - 1) % TSTAGES;
            mbar_wait(bar_empty + 8 * sp, ((kt - 1) / TSTAGES) & 1);
            issue(kt - 1 + TSTAGES);
        }
    }

    // epilogue (same as gemm.cu)
    const double alpha = g.alpha, beta = g.beta;
    double* __restrict__ C = g.C + (long long)blockIdx.y * g.stride2C;
    double* __restrict__ Ct = g.Ct ? g.Ct + (long long)blockIdx.y * g.stride2Ct : nullptr;
#pragma unroll
    for (int mi = 0; mi < 4; ++mi) {
        const int row = m0 + wm * 32 + mi * 8 + gq;
        if (row >= g.M) continue;
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) {
            const int col = n0 + wn * 32 + ni * 8 + 2 * kk;
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

// ---- host side: tensor maps through the driver entry point (no link-time dependency on libcuda) ----------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
        else
            cudaGetLastError();  // not an error of the caller: the cp.async kernel serves instead
    }
    return fn;
}

static bool make_map(CUtensorMap* map, const double* base, long long ld, int rows, int K, int box_rows,
                     int batch2 = 1, long long stride2 = 0) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return false;
    if (batch2 > 1) {
        const cuuint64_t dims[3] = {(cuuint64_t)K, (cuuint64_t)rows, (cuuint64_t)batch2};
        const cuuint64_t strides[2] = {(cuuint64_t)ld * 8, (cuuint64_t)stride2 * 8};
        const cuuint32_t box[3] = {(cuuint32_t)TBK, (cuuint32_t)box_rows, 1};
        const cuuint32_t estr[3] = {1, 1, 1};
        return fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, const_cast<double*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
    }
    const cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)ld * 8};
    const cuuint32_t box[2] = {(cuuint32_t)TBK, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    return fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double*>(base), dims, strides, box, estr,
              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// Returns GPMP_OK when the product was launched on the TMA kernel, 1 when the shape / environment is not one it
// serves (the caller then takes the cp.async kernel), a negative GPMP_ERR_* on a launch failure.
int launch_gemm_nt_tma(const GemmDesc& g, cudaStream_t stream) {
    static const bool disabled = getenv("GPMP_DEV_NO_TMA") != nullptr;
    if (disabled) return 1;
    if (g.batch != 1) return 1;
    if (g.batch2 > 1 && ((g.stride2A & 1) || (g.stride2B & 1) || g.stride2A <= 0 || g.stride2B <= 0)) return 1;
    if (g.A == g.C || g.B == g.C) return 1;  // in-place products keep the single-column-tile kernel
    if (g.K < TBK || (g.lda & 1) || (g.ldb & 1)) return 1;
    // Short k loops stay on the 64 x 64 kernel: with one 128 x 128 CTA per SM nothing covers a tile's prologue and
    // epilogue, and at K = 512 they are ~15 % of its life (measured, n = 8192 SYRK, TFLOP/s: K = 128 19.5 vs 24.0,
    // K = 256 24.5 vs 29.3, K = 512 28.3 vs 30.9; K = 8192 35.4 vs 32.8).  Development knob: GPMP_DEV_TMA_MINK.
    static const int min_k = getenv("GPMP_DEV_TMA_MINK") ? atoi(getenv("GPMP_DEV_TMA_MINK")) : 640;
    if ((g.krange == KR_FULL ? g.K : g.K / 2) < min_k) return 1;
    GemmTmaArgs a;
    a.g = g;
    a.tiles_m = ceil_div(g.M, TBM);
    a.tiles_n = ceil_div(g.N, TBN);
    long long ntiles;
    if (g.lower) {
        if (a.tiles_n > a.tiles_m) a.tiles_n = a.tiles_m;
        ntiles = (long long)a.tiles_n * (a.tiles_n + 1) / 2 + (long long)(a.tiles_m - a.tiles_n) * a.tiles_n;
    } else {
        ntiles = (long long)a.tiles_m * a.tiles_n;
    }
    if (ntiles * g.batch2 < 148) return 1;  // cannot fill the machine with one 128 x 128 tile per SM: latency shapes
    CUtensorMap mapA, mapB;
    if (!make_map(&mapA, g.A, g.lda, g.M, g.K, TBM, g.batch2, g.stride2A) ||
        !make_map(&mapB, g.B, g.ldb, g.N, g.K, TBN, g.batch2, g.stride2B))
        return 1;
    static unsigned long long configured = 0;
    int dev = 0;
    cudaGetDevice(&dev);
    if (!((configured >> (dev & 63)) & 1ull)) {
        if (cudaFuncSetAttribute(gemm_nt_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, T_SMEM) != cudaSuccess)
            return GPMP_ERR_CUDA;
        configured |= 1ull << (dev & 63);
    }
    // algorithmic flops of the launch: triangular K ranges average to K/2 over a square operand
    const double kavg = g.krange == KR_FULL ? (double)g.K : 0.5 * (double)g.K;
    const double work = 2.0 * (double)ntiles * TBM * TBN * kavg * g.batch2;
    LaunchScope scope(KC_GEMM, work, stream);
    gemm_nt_tma_kernel<<<dim3((unsigned)ntiles, (unsigned)g.batch2), T_THREADS, T_SMEM, stream>>>(mapA, mapB, a);
    GPMP_CHECK_LAUNCH();
    return GPMP_OK;
}

}  // namespace gpmp
