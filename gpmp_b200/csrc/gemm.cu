// FP64 tensor-core (DMMA.8x8x4) "NT" GEMM:  C[MxN] = alpha * A[MxK] * B[NxK]^T + beta * C
//
// This one kernel is the flop engine of the whole path: SYRK trailing updates and TRSM-by-inverse
// of the right-looking Cholesky, the level GEMMs of the triangular inverse, LAUUM (K^-1 = T^T T),
// and the whitening/conditioning GEMMs of predict.  Every product on the path is arranged so that
// both operands are read with K contiguous (row-major "K-major" tiles), see DESIGN.md.
//
// Tiling: CTA 128x128, K chunk 16 (one 128-byte row per tile row), 4-stage cp.async pipeline,
// 8 warps as 2(M) x 4(N), warp tile 64x32 = 8x4 DMMA tiles, 128 accumulator registers/thread.
// Shared tiles are dense 128-byte rows with the 16-byte chunk index XOR-swizzled by (row & 7) -- the
// layout a TMA SWIZZLE_128B box produces.  Inside a K chunk the k index is permuted (lane kk owns
// k = 4*kk + s at MMA step s) so each lane fetches its 4 steps with two conflict-free LDS.128.
#include "common.cuh"

namespace gpmp {

constexpr int BM = 128, BN = 128, BK = 16, STAGES = 4, GEMM_THREADS = 256;
constexpr int TILE_BYTES = BM * BK * 8;       // 16 KB
constexpr int STAGE_BYTES = 2 * TILE_BYTES;   // A tile + B tile
constexpr int GEMM_SMEM = STAGES * STAGE_BYTES;  // 128 KB

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

__device__ __forceinline__ void load_tile(uint32_t sdst, const double* __restrict__ src, long long ld,
                                          int row0, int nrows, int k, int k1, int tid) {
    // 128 rows x 8 chunks of 16 B; thread -> (row = tid/8 + 32*i, chunk = tid%8)
    int c = tid & 7;
    int kk = k + 2 * c;
    int rem = k1 - kk;
    int nb = rem >= 2 ? 16 : (rem == 1 ? 8 : 0);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int r = (tid >> 3) + 32 * i;
        int grow = row0 + r;
        int bytes = (grow < nrows) ? nb : 0;
        const double* p = bytes ? (src + (long long)grow * ld + kk) : src;
        uint32_t d = sdst + r * 128 + ((c ^ (r & 7)) << 4);
        cp_async16(d, p, bytes);
    }
}

__global__ void __launch_bounds__(GEMM_THREADS, 1) gemm_nt_kernel(const GemmKArgs a) {
    extern __shared__ __align__(1024) unsigned char smem[];
    const GemmDesc& g = a.g;
    int ti, tj;
    decode_tile(a, g.reverse ? (int)(gridDim.x - 1 - blockIdx.x) : (int)blockIdx.x, ti, tj);
    const int m0 = ti * BM, n0 = tj * BN;
    const long long zb = blockIdx.z, yb = blockIdx.y;
    const double* __restrict__ A = g.A + zb * g.strideA + yb * g.stride2A;
    const double* __restrict__ B = g.B + zb * g.strideB + yb * g.stride2B;
    double* __restrict__ C = g.C + zb * g.strideC + yb * g.stride2C;

    // K trimming for triangular operands (tile granularity; the operand's zero part may hold
    // anything -- e.g. the other triangle of a symmetrised store -- so trimming is also what makes
    // the product correct, and callers keep triangular operands 128-aligned with the tile grid).
    int k0 = 0, k1 = g.K;
    if (g.krange == KR_FROM_ROW) k0 = min(m0, g.K);
    else if (g.krange == KR_TO_ROW) k1 = min(g.K, m0 + BM);
    else if (g.krange == KR_FROM_COL) k0 = min(n0, g.K);
    else if (g.krange == KR_TO_COL) k1 = min(g.K, n0 + BN);
    const int nk = k1 > k0 ? (k1 - k0 + BK - 1) / BK : 0;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp >> 2, wn = warp & 3;      // 2 x 4 warps
    const int gq = lane >> 2, kk = lane & 3;       // fragment row/col group, k lane
    const uint32_t sbase = smem_u32(smem);

    double acc[8][4][2];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    // prologue
#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) {
        if (s < nk) {
            uint32_t st = sbase + s * STAGE_BYTES;
            load_tile(st, A, g.lda, m0, g.M, k0 + s * BK, k1, tid);
            load_tile(st + TILE_BYTES, B, g.ldb, n0, g.N, k0 + s * BK, k1, tid);
        }
        cp_async_commit();
    }

    // per-thread swizzled fragment offsets (row & 7 == gq for every fragment row of this lane)
    const uint32_t aoff = (wm * 64 + gq) * 128;
    const uint32_t boff = TILE_BYTES + (wn * 32 + gq) * 128;
    const uint32_t c0 = ((2 * kk) ^ gq) << 4, c1 = ((2 * kk + 1) ^ gq) << 4;

    for (int kt = 0; kt < nk; ++kt) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        {
            int nx = kt + STAGES - 1;
            if (nx < nk) {
                uint32_t st = sbase + (nx % STAGES) * STAGE_BYTES;
                load_tile(st, A, g.lda, m0, g.M, k0 + nx * BK, k1, tid);
                load_tile(st + TILE_BYTES, B, g.ldb, n0, g.N, k0 + nx * BK, k1, tid);
            }
            cp_async_commit();
        }
        const unsigned char* st = smem + (kt % STAGES) * STAGE_BYTES;
        double af[8][4];
#pragma unroll
        for (int mi = 0; mi < 8; ++mi) {
            const double2 v0 = *reinterpret_cast<const double2*>(st + aoff + mi * 1024 + c0);
            const double2 v1 = *reinterpret_cast<const double2*>(st + aoff + mi * 1024 + c1);
            af[mi][0] = v0.x; af[mi][1] = v0.y; af[mi][2] = v1.x; af[mi][3] = v1.y;
        }
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) {
            const double2 w0 = *reinterpret_cast<const double2*>(st + boff + ni * 1024 + c0);
            const double2 w1 = *reinterpret_cast<const double2*>(st + boff + ni * 1024 + c1);
            const double bf[4] = {w0.x, w0.y, w1.x, w1.y};
#pragma unroll
            for (int s = 0; s < 4; ++s)
#pragma unroll
                for (int mi = 0; mi < 8; ++mi) dmma884(acc[mi][ni][0], acc[mi][ni][1], af[mi][s], bf[s]);
        }
    }
    cp_async_wait<0>();

    // epilogue
    const double alpha = g.alpha, beta = g.beta;
    double* __restrict__ Ct = g.Ct ? g.Ct + zb * g.strideCt + yb * g.stride2Ct : nullptr;
#pragma unroll
    for (int mi = 0; mi < 8; ++mi) {
        const int row = m0 + wm * 64 + mi * 8 + gq;
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

int launch_gemm_nt(const GemmDesc& g, cudaStream_t stream) {
    if (g.M <= 0 || g.N <= 0 || g.batch <= 0 || g.batch2 <= 0) return GPMP_OK;
    // 16-byte cp.async / vector epilogue requirements
    if ((g.lda & 1) || (g.ldb & 1) || (g.ldc & 1)) return GPMP_ERR_ALIGN;
    if ((reinterpret_cast<uintptr_t>(g.A) & 15) || (reinterpret_cast<uintptr_t>(g.B) & 15) ||
        (reinterpret_cast<uintptr_t>(g.C) & 15))
        return GPMP_ERR_ALIGN;
    if ((g.strideA & 1) || (g.strideB & 1) || (g.strideC & 1)) return GPMP_ERR_ALIGN;
    if ((g.stride2A & 1) || (g.stride2B & 1) || (g.stride2C & 1)) return GPMP_ERR_ALIGN;
    static bool configured = false;
    if (!configured) {
        if (cudaFuncSetAttribute(gemm_nt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM) !=
            cudaSuccess)
            return GPMP_ERR_CUDA;
        configured = true;
    }
    GemmKArgs a;
    a.g = g;
    a.tiles_m = ceil_div(g.M, BM);
    a.tiles_n = ceil_div(g.N, BN);
    long long ntiles;
    if (g.lower) {
        if (a.tiles_n > a.tiles_m) a.tiles_n = a.tiles_m;
        ntiles = (long long)a.tiles_n * (a.tiles_n + 1) / 2 + (long long)(a.tiles_m - a.tiles_n) * a.tiles_n;
    } else {
        ntiles = (long long)a.tiles_m * a.tiles_n;
    }
    double kavg = g.krange == KR_FULL ? (double)g.K : 0.5 * (double)g.K;
    double work = 2.0 * (double)ntiles * BM * BN * kavg * g.batch * g.batch2;
    LaunchScope scope(KC_GEMM, work, stream);
    dim3 grid((unsigned)ntiles, (unsigned)g.batch2, (unsigned)g.batch);
    gemm_nt_kernel<<<grid, GEMM_THREADS, GEMM_SMEM, stream>>>(a);
    GPMP_CHECK_LAUNCH();
    return GPMP_OK;
}

}  // namespace gpmp
