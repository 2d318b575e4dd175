// K2: blocked right-looking fp64 Cholesky, triangular inverse and K^-1, all on the DMMA GEMM.
//
// Storage convention ("both-ways" triangles): every triangular matrix on the path is kept so that
// both its rows and its columns can be read K-contiguous by the NT GEMM:
//   A      n x n  : L in the lower tiles; diagonal 128-tiles hold L_aa with an explicitly ZERO upper
//                   part; strictly-upper tiles hold the mirrored L^T tiles.
//   Tlo/Tup        : T = L^-1 (lower, zero upper part in diagonal tiles) and T^T (upper, zero lower
//                   part in diagonal tiles) as two separate matrices.
// K-range trimming in the GEMM is at 128-tile granularity, which is why diagonal tiles carry real
// zeros and why every block boundary used below is a multiple of 128.
//
// gpmp_potrf: two-level right-looking factorisation.  Outer block NB (128/256/512): the NB diagonal
// block is factored by 128-wide sub-steps (potf2 tile kernel -> in-place TRSM by the tile inverse ->
// SYRK), its inverse is assembled by block doubling (inv [[A,0],[B,C]] = [[A^-1,0],[-C^-1 B A^-1,C^-1]]),
// the panel below is solved with ONE GEMM against that inverse and the trailing matrix gets ONE
// K=NB SYRK.  Extra rows n..nrows-1 ride along in every panel solve (they leave as B L^-T).
#include "internal.cuh"

namespace gpmp {

constexpr int PT = 128;          // base tile
constexpr int PLD = 129;         // smem leading dimension of the tile
constexpr int XLD = 65;          // smem leading dimension of the doubling scratch
constexpr int POTF2_THREADS = 512;
constexpr int POTF2_SMEM = (PT * PLD + 64 * XLD + PT) * 8;

struct Potf2Args {
    double* A; long long lda; long long strideA;      // tile origin (diagonal position), in/out
    double* Tlo; double* Tup; long long ldt; long long strideT;  // inverse tile out (may be null)
    int nb;            // live size of the tile (<= 128); the rest is padded with identity
    int* info; long long strideInfo;
    int row0;          // global index of the tile's first row (for info)
};

// ---- tiny warp-level DMMA GEMM over shared memory -------------------------------------------
// For every 8x8 output tile (i8, j8) accepted by `pick`, computes sum_k a(i,k) * b(j,k) over
// k in [0,K) (K multiple of 4) and hands the two accumulators of each lane to `out(i, j, c0, c1)`
// (element (i, j) and (i, j+1)).
template <class FA, class FB, class FP, class FO>
__device__ __forceinline__ void smem_mma(int M8, int N8, int K, FA a, FB b, FP pick, FO out) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const int gq = lane >> 2, kk = lane & 3;
    for (int t = warp; t < M8 * N8; t += nw) {
        const int i8 = t / N8, j8 = t - i8 * N8;
        if (!pick(i8, j8)) continue;
        double c0 = 0.0, c1 = 0.0;
        const int i = i8 * 8 + gq, j = j8 * 8 + gq;
        for (int k = 0; k < K; k += 4) dmma884(c0, c1, a(i, k + kk), b(j, k + kk));
        out(i, j8 * 8 + 2 * kk, c0, c1);
    }
}

// Factor one 128x128 diagonal tile (lower) and invert the factor.  One CTA per tile.
// smem S: lower = L, strict upper = T^T (T_ij stored at S[j][i]), tdiag = diag(T).
__global__ void __launch_bounds__(POTF2_THREADS, 1) potf2_kernel(const Potf2Args a) {
    extern __shared__ __align__(16) double sm[];
    double* S = sm;
    double* X = sm + PT * PLD;
    double* tdiag = X + 64 * XLD;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long zb = blockIdx.x;
    double* __restrict__ A = a.A + zb * a.strideA;
    const int nb = a.nb;

    // load the lower triangle, identity padding outside the live block
    for (int e = tid; e < PT * PT; e += POTF2_THREADS) {
        const int r = e >> 7, c = e & 127;
        double v = 0.0;
        if (r < nb && c <= r) v = A[(long long)r * a.lda + c];
        else if (r == c) v = 1.0;
        S[r * PLD + c] = v;
    }
    __syncthreads();

    int bad = 0;  // 1-based local index of the first non-positive pivot (warp 0 only)
    for (int jb = 0; jb < 4; ++jb) {
        const int c0 = jb * 32;
        if (warp == 0) {
            // (a) left-looking Cholesky of the 32x32 diagonal block, lane = row
            double* D = S + c0 * PLD + c0;
            for (int j = 0; j < 32; ++j) {
                double s0 = 0.0, s1 = 0.0;
                if (lane >= j) {
                    int k = 0;
                    for (; k + 1 < j; k += 2) {
                        s0 = fma(D[lane * PLD + k], D[j * PLD + k], s0);
                        s1 = fma(D[lane * PLD + k + 1], D[j * PLD + k + 1], s1);
                    }
                    if (k < j) s0 = fma(D[lane * PLD + k], D[j * PLD + k], s0);
                }
                const double v = (lane >= j ? D[lane * PLD + j] : 0.0) - (s0 + s1);
                const double piv = __shfl_sync(0xffffffffu, v, j);
                if (!(piv > 0.0) && bad == 0) bad = c0 + j + 1;
                const double dj = sqrt(piv);
                if (lane == j) D[j * PLD + j] = dj;
                else if (lane > j) D[lane * PLD + j] = v / dj;
                __syncwarp();
            }
            // (b) invert it: lane = column j of T; T_ij kept at the transposed (upper) position
            {
                const int j = lane;
                const double tjj = 1.0 / D[j * PLD + j];
                tdiag[c0 + j] = tjj;
                for (int i = j + 1; i < 32; ++i) {
                    double s0 = D[i * PLD + j] * tjj, s1 = 0.0;
                    int k = j + 1;
                    for (; k + 1 < i; k += 2) {
                        s0 = fma(D[i * PLD + k], D[j * PLD + k], s0);
                        s1 = fma(D[i * PLD + k + 1], D[j * PLD + k + 1], s1);
                    }
                    if (k < i) s0 = fma(D[i * PLD + k], D[j * PLD + k], s0);
                    D[j * PLD + i] = -(s0 + s1) / D[i * PLD + i];
                }
            }
        }
        __syncthreads();
        const int rb = c0 + 32;       // first row below the diagonal block
        const int mrows = PT - rb;    // rows below
        if (mrows > 0) {
            // (c) X = S[rb.., c0..c0+32) * Td^T   (Td = inverse of the diagonal block), into scratch
            //     thread -> (row, group of 8 columns)
            for (int e = tid; e < mrows * 4; e += POTF2_THREADS) {
                const int r = e % mrows, cg = e / mrows;
                const double* src = S + (rb + r) * PLD + c0;
                double o[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) o[u] = 0.0;
                for (int k = 0; k < 8 * cg + 8; ++k) {
                    const double v = src[k];
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const int j = 8 * cg + u;
                        // T[j][k] : k<j at S[c0+k][c0+j], k==j tdiag, k>j zero
                        const double t = k < j ? S[(c0 + k) * PLD + c0 + j] : (k == j ? tdiag[c0 + j] : 0.0);
                        o[u] = fma(v, t, o[u]);
                    }
                }
                // scratch is 64 x XLD: rows r (< 96) do not fit -> use two halves of 48 rows x 32
                double* dst = X + (r % 48) * XLD + (r / 48) * 32 + 0;
#pragma unroll
                for (int u = 0; u < 8; ++u) dst[8 * cg + u] = o[u];
            }
            __syncthreads();
            for (int e = tid; e < mrows * 32; e += POTF2_THREADS) {
                const int r = e >> 5, c = e & 31;
                S[(rb + r) * PLD + c0 + c] = X[(r % 48) * XLD + (r / 48) * 32 + c];
            }
            __syncthreads();
            // (d) trailing update of the lower 8x8 tiles: S[rb.., rb..] -= P P^T, P = S[rb.., c0..c0+32)
            const double* P = S + rb * PLD + c0;
            double* C = S + rb * PLD + rb;
            smem_mma(
                mrows / 8, mrows / 8, 32, [&](int i, int k) { return P[i * PLD + k]; },
                [&](int j, int k) { return P[j * PLD + k]; }, [&](int i8, int j8) { return j8 <= i8; },
                [&](int i, int j, double c0v, double c1v) {
                    // diagonal 8x8 tiles: touch the lower part only (the upper part is reserved for T^T)
                    if (j <= i) C[i * PLD + j] -= c0v;
                    if (j + 1 <= i) C[i * PLD + j + 1] -= c1v;
                });
            __syncthreads();
        }
    }
    if (warp == 0 && lane == 0 && bad && a.info) {
        int* ip = a.info + zb * a.strideInfo;
        atomicCAS(ip, 0, a.row0 + bad);
    }

    // ---- inverse by doubling: 32 -> 64 -> 128 ------------------------------------------------
    // T(i,k) for the already inverted diagonal blocks
    auto Tget = [&](int i, int k) -> double {
        return i > k ? S[k * PLD + i] : (i == k ? tdiag[i] : 0.0);
    };
    for (int s = 32; s < PT; s *= 2) {
        for (int pr = 0; pr < PT / (2 * s); ++pr) {
            const int a0 = 2 * pr * s, b0 = a0 + s;
            // X (s x s) = L_ba * T_aa
            smem_mma(
                s / 8, s / 8, s, [&](int i, int k) { return S[(b0 + i) * PLD + a0 + k]; },
                [&](int j, int k) { return Tget(a0 + k, a0 + j); }, [&](int, int) { return true; },
                [&](int i, int j, double c0v, double c1v) {
                    X[i * XLD + j] = c0v;
                    X[i * XLD + j + 1] = c1v;
                });
            __syncthreads();
            // T_ba = -T_bb * X, stored transposed in the upper part: S[a0 + j][b0 + i]
            smem_mma(
                s / 8, s / 8, s, [&](int i, int k) { return Tget(b0 + i, b0 + k); },
                [&](int j, int k) { return X[k * XLD + j]; }, [&](int, int) { return true; },
                [&](int i, int j, double c0v, double c1v) {
                    S[(a0 + j) * PLD + b0 + i] = -c0v;
                    S[(a0 + j + 1) * PLD + b0 + i] = -c1v;
                });
            __syncthreads();
        }
    }

    // ---- write back: L (lower, zero upper) into A; T into Tlo (lower) / Tup (upper) -------------
    for (int e = tid; e < PT * PT; e += POTF2_THREADS) {
        const int r = e >> 7, c = e & 127;
        if (r < nb && c < nb) A[(long long)r * a.lda + c] = c <= r ? S[r * PLD + c] : 0.0;
    }
    if (a.Tlo) {
        double* __restrict__ Tlo = a.Tlo + zb * a.strideT;
        double* __restrict__ Tup = a.Tup + zb * a.strideT;
        for (int e = tid; e < PT * PT; e += POTF2_THREADS) {
            const int r = e >> 7, c = e & 127;
            if (r < nb && c < nb) {
                const double t = r > c ? S[c * PLD + r] : (r == c ? tdiag[r] : 0.0);  // T[r][c]
                const double tt = c > r ? S[r * PLD + c] : (r == c ? tdiag[r] : 0.0); // T^T[r][c] = T[c][r]
                Tlo[(long long)r * a.ldt + c] = t;
                Tup[(long long)r * a.ldt + c] = tt;
            }
        }
    }
}

static int launch_potf2(const Potf2Args& a, int batch, cudaStream_t stream) {
    static bool configured = false;
    if (!configured) {
        if (cudaFuncSetAttribute(potf2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, POTF2_SMEM) !=
            cudaSuccess)
            return GPMP_ERR_CUDA;
        configured = true;
    }
    LaunchScope scope(KC_POTF2, (double)batch * (PT * (double)PT * PT), stream);  // n^3/3 + 2n^3/3
    potf2_kernel<<<batch, POTF2_THREADS, POTF2_SMEM, stream>>>(a);
    GPMP_CHECK_LAUNCH();
    return GPMP_OK;
}

// ---- panel copy-back: W (rows x nb, ld ldw) -> A panel (lower) and its mirror in the upper tiles --
struct CopyPanelArgs {
    const double* W; long long ldw; long long strideW;
    double* Alo; double* Aup; long long lda; long long strideA;  // Alo: panel origin; Aup: mirrored origin
    int rows, cols, mirror_rows;  // rows of the panel; only the first mirror_rows are mirrored
};
__global__ void __launch_bounds__(256) copy_panel_kernel(const CopyPanelArgs a) {
    __shared__ double t[32][33];
    const long long zb = blockIdx.z;
    const double* __restrict__ W = a.W + zb * a.strideW;
    double* __restrict__ Alo = a.Alo + zb * a.strideA;
    double* __restrict__ Aup = a.Aup ? a.Aup + zb * a.strideA : nullptr;
    const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int i = ty; i < 32; i += 8) {
        const int r = r0 + i, c = c0 + tx;
        double v = 0.0;
        if (r < a.rows && c < a.cols) {
            v = W[(long long)r * a.ldw + c];
            Alo[(long long)r * a.lda + c] = v;
        }
        t[i][tx] = v;
    }
    __syncthreads();
    if (Aup) {
        for (int i = ty; i < 32; i += 8) {
            const int c = c0 + i, r = r0 + tx;  // writes Aup[c][r] = W[r][c]
            if (r < a.mirror_rows && c < a.cols) Aup[(long long)c * a.lda + r] = t[tx][i];
        }
    }
}

static int launch_copy_panel(const CopyPanelArgs& a, int batch, cudaStream_t stream) {
    if (a.rows <= 0 || a.cols <= 0) return GPMP_OK;
    LaunchScope scope(KC_SMALL, 0.0, stream);
    dim3 grid(ceil_div(a.cols, 32), ceil_div(a.rows, 32), batch);
    copy_panel_kernel<<<grid, 256, 0, stream>>>(a);
    GPMP_CHECK_LAUNCH();
    return GPMP_OK;
}

int potrf_block_size(int n) {
    if (n <= 1024) return 128;
    if (n <= 4096) return 256;
    return 512;
}

// Doubling pass over the diagonal blocks of size s inside [0, len): for each pair (first block full,
// second block of size s2 <= s) computes T21 = -T22 * L21 * T11 into Tlo (and its mirror into Tup).
// L, Tlo, Tup, Xt are addressed from the origin of the range; Xt is scratch with the same indexing.
static int doubling_level(const double* L, long long ldl, long long strideL, double* Tlo, double* Tup,
                          long long ldt, long long strideT, double* Xt, long long ldx, long long strideX,
                          int len, int s, int batch, cudaStream_t stream) {
    const int npairs_full = len / (2 * s);
    const int rem = len - npairs_full * 2 * s;  // leftover: a ragged pair if rem > s
    for (int pass = 0; pass < 2; ++pass) {
        int np, s2;
        long long o;  // origin (block index offset) of the first pair of this pass
        if (pass == 0) { np = npairs_full; s2 = s; o = 0; }
        else { np = rem > s ? 1 : 0; s2 = rem - s; o = (long long)npairs_full * 2 * s; }
        if (np <= 0) continue;
        // Xt (s x s2) = Tup11 (rows j, k >= j) x L21^T   [NT: A = Tup11, B = L21]
        GemmDesc g = gemm_desc();
        g.A = Tup + o * (ldt + 1); g.lda = ldt; g.strideA = strideT; g.stride2A = 2LL * s * (ldt + 1);
        g.B = L + (o + s) * ldl + o; g.ldb = ldl; g.strideB = strideL; g.stride2B = 2LL * s * (ldl + 1);
        g.C = Xt + o * ldx + (o + s); g.ldc = ldx; g.strideC = strideX; g.stride2C = 2LL * s * (ldx + 1);
        g.M = s; g.N = s2; g.K = s; g.krange = KR_FROM_ROW; g.batch = batch; g.batch2 = np;
        int rc = launch_gemm_nt(g, stream);
        if (rc) return rc;
        // T21 (s2 x s) = -Tlo22 (rows i, k <= i) x Xt^T  [NT: A = Tlo22, B = Xt]
        GemmDesc h = gemm_desc();
        h.A = Tlo + (o + s) * (ldt + 1); h.lda = ldt; h.strideA = strideT; h.stride2A = 2LL * s * (ldt + 1);
        h.B = Xt + o * ldx + (o + s); h.ldb = ldx; h.strideB = strideX; h.stride2B = 2LL * s * (ldx + 1);
        h.C = Tlo + (o + s) * ldt + o; h.ldc = ldt; h.strideC = strideT; h.stride2C = 2LL * s * (ldt + 1);
        h.Ct = Tup + o * ldt + (o + s); h.ldct = ldt; h.strideCt = strideT; h.stride2Ct = 2LL * s * (ldt + 1);
        h.M = s2; h.N = s; h.K = s2; h.alpha = -1.0; h.krange = KR_TO_ROW; h.reverse = 1;
        h.batch = batch; h.batch2 = np;
        rc = launch_gemm_nt(h, stream);
        if (rc) return rc;
    }
    return GPMP_OK;
}

// Core factorisation (optionally batched over blockIdx.z with element strides).
//   A      (nrows x lda)            in: lower of K (+ extra rows); out: L both-ways (+ whitened rows)
//   Tlo/Tup compact diagonal-block inverses: nblk blocks of NB x NB (ld NB), block b at b*NB*NB
//   W      panel scratch (nrows x NB, ld NB), needed when NB > 128
int potrf_core(double* A, long long lda, long long strideA, int n, int nrows, int NB, double* Tlo, double* Tup,
               long long strideT, double* W, long long strideW, int* info, long long strideInfo, int batch,
               cudaStream_t stream) {
    if (n <= 0) return GPMP_OK;
    int rc;
    for (int k = 0; k < n; k += NB) {
        const int nbk = min(NB, n - k);
        double* Akk = A + (long long)k * (lda + 1);
        double* Tlo_k = Tlo + (long long)(k / NB) * NB * NB;
        double* Tup_k = Tup + (long long)(k / NB) * NB * NB;
        // ---- diagonal block: 128-wide sub-steps
        for (int j0 = 0; j0 < nbk; j0 += PT) {
            const int jb = min(PT, nbk - j0);
            Potf2Args pa;
            pa.A = Akk + (long long)j0 * (lda + 1); pa.lda = lda; pa.strideA = strideA;
            pa.Tlo = Tlo_k + (long long)j0 * (NB + 1); pa.Tup = Tup_k + (long long)j0 * (NB + 1);
            pa.ldt = NB; pa.strideT = strideT; pa.nb = jb; pa.info = info; pa.strideInfo = strideInfo;
            pa.row0 = k + j0;
            rc = launch_potf2(pa, batch, stream);
            if (rc) return rc;
            const int m = nbk - (j0 + jb);
            if (m > 0) {
                // in-place TRSM of the rows below inside the diagonal block (single column tile)
                GemmDesc g = gemm_desc();
                g.A = Akk + (long long)(j0 + jb) * lda + j0; g.lda = lda; g.strideA = strideA;
                g.B = pa.Tlo; g.ldb = NB; g.strideB = strideT;
                g.C = Akk + (long long)(j0 + jb) * lda + j0; g.ldc = lda; g.strideC = strideA;
                g.Ct = Akk + (long long)j0 * lda + (j0 + jb); g.ldct = lda; g.strideCt = strideA;
                g.M = m; g.N = jb; g.K = jb; g.batch = batch;
                rc = launch_gemm_nt(g, stream);
                if (rc) return rc;
                GemmDesc h = gemm_desc();
                h.A = g.C; h.lda = lda; h.strideA = strideA;
                h.B = g.C; h.ldb = lda; h.strideB = strideA;
                h.C = Akk + (long long)(j0 + jb) * (lda + 1); h.ldc = lda; h.strideC = strideA;
                h.M = m; h.N = m; h.K = jb; h.alpha = -1.0; h.beta = 1.0; h.lower = 1; h.batch = batch;
                rc = launch_gemm_nt(h, stream);
                if (rc) return rc;
            }
        }
        // ---- inverse of the NB diagonal block by doubling (also for the last block: the compact
        //      inverses are what the row solves and the triangular inverse start from)
        const int r0 = k + nbk;
        const int M = nrows - r0;
        for (int s = PT; s < nbk; s *= 2) {
            // scratch: the strictly-upper part of W is free here (W is consumed only by the panel solve)
            rc = doubling_level(Akk, lda, strideA, Tlo_k, Tup_k, NB, strideT, W, NB, strideW, nbk, s, batch,
                                stream);
            if (rc) return rc;
        }
        if (M <= 0) break;
        // ---- panel solve: rows r0..nrows-1 of columns k..k+nbk
        double* P = A + (long long)r0 * lda + k;
        const int mirror_rows = max(0, n - r0);
        const double* Pnl;  // solved panel as GEMM operand
        long long ldp, strideP;
        if (nbk <= PT) {
            GemmDesc g = gemm_desc();
            g.A = P; g.lda = lda; g.strideA = strideA;
            g.B = Tlo_k; g.ldb = NB; g.strideB = strideT;
            g.C = P; g.ldc = lda; g.strideC = strideA;
            g.M = M; g.N = nbk; g.K = nbk; g.batch = batch;
            rc = launch_gemm_nt(g, stream);
            if (rc) return rc;
            if (mirror_rows > 0) {
                CopyPanelArgs c;
                c.W = P; c.ldw = lda; c.strideW = strideA; c.Alo = P; c.Aup = A + (long long)k * lda + r0;
                c.lda = lda; c.strideA = strideA; c.rows = mirror_rows; c.cols = nbk; c.mirror_rows = mirror_rows;
                rc = launch_copy_panel(c, batch, stream);
                if (rc) return rc;
            }
            Pnl = P; ldp = lda; strideP = strideA;
        } else {
            GemmDesc g = gemm_desc();
            g.A = P; g.lda = lda; g.strideA = strideA;
            g.B = Tlo_k; g.ldb = NB; g.strideB = strideT;
            g.C = W; g.ldc = NB; g.strideC = strideW;
            g.M = M; g.N = nbk; g.K = nbk; g.krange = KR_TO_COL; g.batch = batch;
            rc = launch_gemm_nt(g, stream);
            if (rc) return rc;
            CopyPanelArgs c;
            c.W = W; c.ldw = NB; c.strideW = strideW; c.Alo = P; c.Aup = A + (long long)k * lda + r0;
            c.lda = lda; c.strideA = strideA; c.rows = M; c.cols = nbk; c.mirror_rows = mirror_rows;
            rc = launch_copy_panel(c, batch, stream);
            if (rc) return rc;
            Pnl = W; ldp = NB; strideP = strideW;
        }
        // ---- trailing update (lower tiles; the extra rows are the rectangular tail of the tile list)
        const int Nn = n - r0;
        if (Nn > 0) {
            GemmDesc h = gemm_desc();
            h.A = Pnl; h.lda = ldp; h.strideA = strideP;
            h.B = Pnl; h.ldb = ldp; h.strideB = strideP;
            h.C = A + (long long)r0 * (lda + 1); h.ldc = lda; h.strideC = strideA;
            h.M = M; h.N = Nn; h.K = nbk; h.alpha = -1.0; h.beta = 1.0; h.lower = 1; h.batch = batch;
            rc = launch_gemm_nt(h, stream);
            if (rc) return rc;
        }
    }
    return GPMP_OK;
}

// ---- potri: Tlo/Tup (n x n) from L and the compact NB-block inverses, then Kinv = T^T T (lower) ----
struct CopyDiagArgs {
    const double* slo; const double* sup; int NB;  // compact blocks (ld NB)
    double* dlo; double* dup; long long ld;        // full matrices
    int n;
};
__global__ void copy_diag_blocks_kernel(const CopyDiagArgs a) {
    const int b = blockIdx.y;
    const int nbk = min(a.NB, a.n - b * a.NB);
    const long long src0 = (long long)b * a.NB * a.NB;
    const long long dst0 = (long long)b * a.NB * (a.ld + 1);
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < (long long)nbk * nbk;
         e += (long long)gridDim.x * blockDim.x) {
        const int r = (int)(e / nbk), c = (int)(e - (long long)r * nbk);
        a.dlo[dst0 + (long long)r * a.ld + c] = a.slo[src0 + (long long)r * a.NB + c];
        a.dup[dst0 + (long long)r * a.ld + c] = a.sup[src0 + (long long)r * a.NB + c];
    }
}

int potri_core(const double* L, int n, long long ldl, int NB, const double* Tlo_c, const double* Tup_c,
               double* Tlo, double* Tup, double* Kinv, long long ldk, cudaStream_t stream) {
    if (n <= 0) return GPMP_OK;
    int rc;
    const int nblk = ceil_div(n, NB);
    {
        CopyDiagArgs c;
        c.slo = Tlo_c; c.sup = Tup_c; c.NB = NB; c.dlo = Tlo; c.dup = Tup; c.ld = ldk; c.n = n;
        LaunchScope scope(KC_SMALL, 0.0, stream);
        dim3 grid(min(64, ceil_div(NB * NB, 256)), nblk);
        copy_diag_blocks_kernel<<<grid, 256, 0, stream>>>(c);
        GPMP_CHECK_LAUNCH();
    }
    // doubling NB -> n; scratch for X^T: the Kinv buffer (written only by the final product)
    for (long long s = NB; s < n; s *= 2) {
        rc = doubling_level(L, ldl, 0, Tlo, Tup, ldk, 0, Kinv, ldk, 0, n, (int)s, 1, stream);
        if (rc) return rc;
    }
    // Kinv (lower) = T^T T :  Kinv_ij = sum_{k >= i} Tup[i][k] Tup[j][k]
    GemmDesc g = gemm_desc();
    g.A = Tup; g.lda = ldk; g.B = Tup; g.ldb = ldk; g.C = Kinv; g.ldc = ldk;
    g.M = n; g.N = n; g.K = n; g.lower = 1; g.krange = KR_FROM_ROW;
    return launch_gemm_nt(g, stream);
}

// ---- row-wise triangular solves with many right-hand sides stored as rows ----------------------
// trans == 0:  Bt <- Bt L^-T   (forward; uses Tlo diagonal-block inverses and L's lower tiles)
// trans == 1:  Bt <- Bt L^-1   (backward; uses Tup diagonal-block inverses and the mirrored upper tiles)
// Block size NB as in the factorisation (complete compact inverses required).  W: m x NB scratch.
int trsm_rows_core(const double* A, int n, long long lda, int NB, const double* Tlo_c, const double* Tup_c,
                   double* Bt, int m, long long ldb, int trans, double* W, cudaStream_t stream) {
    if (n <= 0 || m <= 0) return GPMP_OK;
    int rc;
    const int nblk = ceil_div(n, NB);
    for (int bi = 0; bi < nblk; ++bi) {
        const int b = trans ? nblk - 1 - bi : bi;
        const int k = b * NB, nbk = min(NB, n - k);
        const double* T = (trans ? Tup_c : Tlo_c) + (long long)b * NB * NB;
        double* Bk = Bt + k;
        // diagonal solve Bk <- Bk * T^T (trans 0: T lower, k <= col) or Bk * Tup^T (k >= col)
        GemmDesc g = gemm_desc();
        g.A = Bk; g.lda = ldb; g.B = T; g.ldb = NB; g.M = m; g.N = nbk; g.K = nbk;
        g.krange = trans ? KR_FROM_COL : KR_TO_COL;
        if (nbk <= PT) {
            g.C = Bk; g.ldc = ldb; g.krange = KR_FULL;
            rc = launch_gemm_nt(g, stream);
            if (rc) return rc;
        } else {
            g.C = W; g.ldc = NB;
            rc = launch_gemm_nt(g, stream);
            if (rc) return rc;
            CopyPanelArgs c;
            c.W = W; c.ldw = NB; c.strideW = 0; c.Alo = Bk; c.Aup = nullptr; c.lda = ldb; c.strideA = 0;
            c.rows = m; c.cols = nbk; c.mirror_rows = 0;
            rc = launch_copy_panel(c, 1, stream);
            if (rc) return rc;
        }
        // update of the remaining columns
        GemmDesc h = gemm_desc();
        h.A = Bk; h.lda = ldb; h.M = m; h.K = nbk; h.alpha = -1.0; h.beta = 1.0;
        if (!trans) {
            const int rest = n - (k + nbk);
            if (rest <= 0) continue;
            // B[:, k+nbk:] -= Bk * L[k+nbk:, k:k+nbk]^T
            h.B = A + (long long)(k + nbk) * lda + k; h.ldb = lda; h.N = rest;
            h.C = Bt + k + nbk; h.ldc = ldb;
        } else {
            if (k <= 0) continue;
            // B[:, :k] -= Bk * L[k:k+nbk, :k]  = Bk * (mirror tiles A[:k, k:k+nbk])^T
            h.B = A + k; h.ldb = lda; h.N = k;
            h.C = Bt; h.ldc = ldb;
        }
        rc = launch_gemm_nt(h, stream);
        if (rc) return rc;
    }
    return GPMP_OK;
}

}  // namespace gpmp
