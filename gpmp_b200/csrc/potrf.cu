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
// gpmp_potrf: right-looking factorisation in column groups of NB (128/256/512) columns.  Inside a group,
// 128-wide steps: tile kernel (factor + inverse of the 128x128 diagonal tile), solve of ALL rows below by
// one GEMM against the tile inverse (into the group's panel buffer), copy-back + mirror, K=128 update of
// the group's remaining columns.  After the group's last step the buffer holds its complete solved panel
// and the trailing matrix gets ONE K=NB SYRK (lower tiles).  Extra rows n..nrows-1 ride along in every
// solve (they leave as B L^-T).  The NB-wide inverses of the diagonal blocks, which the triangular inverse
// and the row solves start from, are assembled afterwards by block doubling
// (inv [[A,0],[B,C]] = [[A^-1,0],[-C^-1 B A^-1,C^-1]]).  Single large matrices run a multi-stream
// look-ahead (potrf_core, depth 4) whose chain is ONE fused launch per 128-column step (chain_step_kernel: solve
// of the rows below, in-group updates and factorisation of the next tile, with flags in global memory between
// its CTAs) for groups of full tiles, and the factor-only tile kernel + substitution solve (trsm_tile_kernel) +
// K = 128 GEMM otherwise; with the gradient's buffers at hand the leading block of T = L^-1 is computed under the
// latency-bound tail of the factorisation (early_inverse: persistent GEMMs that keep off the chain's SMs).
// Batched value-only sweeps use the packed factor-only tile kernel (two tiles per SM) and one solve CTA per
// matrix; the same pieces serve the panel-partitioned multi-GPU loop (dist_*).
#include <cstdlib>
#include <mutex>
#include <vector>
#include "internal.cuh"

namespace gpmp {

constexpr int PT = 128;          // base tile
constexpr int PLD = 132;         // smem leading dimension of the tile (4 mod 16: conflict-free DMMA fragment loads)
constexpr int XLD = 68;          // smem leading dimension of the scratch (4 mod 16)
constexpr int POTF2_THREADS = 512;
constexpr int POTF2_SMEM = (PT * PLD + 96 * XLD + PT) * 8;
// Packed lower tile: block row b (32 rows) keeps its 32 (b + 1) leading columns with leading dimension
// 32 b + 36 (= 4 mod 16: conflict-free DMMA fragment loads); 10752 doubles instead of 128 x 132.
constexpr int TS_LP = 10752;
__device__ __forceinline__ int ts_ld(int b) { return 32 * b + 36; }
__device__ __forceinline__ int ts_base(int b) { return 512 * b * (b - 1) + 1152 * b; }
__device__ __forceinline__ int pk(int r, int c) {
    const int b = r >> 5;
    return ts_base(b) + (r - 32 * b) * ts_ld(b) + c;
}

struct Potf2Args {
    double* A; long long lda; long long strideA;      // tile origin (diagonal position), in/out
    double* Tlo; double* Tup; long long ldt; long long strideT;  // inverse tile out (may be null)
    int nb;            // live size of the tile (<= 128); the rest is padded with identity
    int* info; long long strideInfo;
    int row0;          // global index of the tile's first row (for info)
    long long* dbg;    // optional: clock64() at the phase boundaries (development only)
    int mode;          // POTF2_FULL: factor + inverse; POTF2_INVERT: the tile already holds L, invert only;
                       // POTF2_FACTOR: factor + the four 32x32 diagonal-block inverses only (block-diagonal T
                       // into Tlo, Tup not written) -- what the substitution solve of the chain needs
};
enum { POTF2_FULL = 0, POTF2_INVERT = 1, POTF2_FACTOR = 2 };
#define POTF2_STAMP(i)                                              \
    do {                                                            \
        if (a.dbg && threadIdx.x == 0) a.dbg[i] = clock64();        \
    } while (0)

// ---- tiny warp-level DMMA GEMM over shared memory -------------------------------------------
// For every 8x8 output tile (i8, j8) accepted by `pick` (tiles dealt round-robin to warps wid, wid+nw, ...),
// computes sum_k a(i,k) * b(j,k) over k in [0,K) and hands the two accumulators of each lane to
// `out(i, j, c0, c1)` (elements (i, j) and (i, j+1)).  K is a template parameter: all fragments of a tile
// are fetched before the first DMMA so the shared-memory latency is paid once per tile.
template <int K, class FA, class FB, class FP, class FO>
__device__ __forceinline__ void smem_mma(int M8, int N8, int wid, int nw, FA a, FB b, FP pick, FO out) {
    const int lane = threadIdx.x & 31;
    const int gq = lane >> 2, kk = lane & 3;
    for (int t = wid; t < M8 * N8; t += nw) {
        const int i8 = t / N8, j8 = t - i8 * N8;
        if (!pick(i8, j8)) continue;
        const int i = i8 * 8 + gq, j = j8 * 8 + gq;
        double av[K / 4], bv[K / 4];
#pragma unroll
        for (int s = 0; s < K / 4; ++s) {
            av[s] = a(i, 4 * s + kk);
            bv[s] = b(j, 4 * s + kk);
        }
        // four independent accumulator pairs: the DMMA dependency chain is K/16 deep instead of K/4
        double c0 = 0.0, c1 = 0.0, d0 = 0.0, d1 = 0.0, e0 = 0.0, e1 = 0.0, f0 = 0.0, f1 = 0.0;
#pragma unroll
        for (int s = 0; s < K / 4; s += 4) {
            dmma884(c0, c1, av[s], bv[s]);
            if (s + 1 < K / 4) dmma884(d0, d1, av[s + 1], bv[s + 1]);
            if (s + 2 < K / 4) dmma884(e0, e1, av[s + 2], bv[s + 2]);
            if (s + 3 < K / 4) dmma884(f0, f1, av[s + 3], bv[s + 3]);
        }
        c0 += e0; c1 += e1; d0 += f0; d1 += f1;
        out(i, j8 * 8 + 2 * kk, c0 + d0, c1 + d1);
    }
}

// Rank-32 update of lower 8x8 tiles:  C[i][j] -= sum_{k<32} P[i][k] P[j][k]  for the tile rows i8 in
// [i8_lo, i8_hi) and j8 in [0, i8].  prow(r) / crow(r) return the address of row r of the 32-column panel / of
// the target (column 0 = tile column 0).  The tiles are dealt to the warps in contiguous, equally long runs
// (row-major), so a run mostly stays inside one tile row and keeps its A fragments; every tile is two
// independent chains of four DMMAs.
template <class RP, class RC>
__device__ __forceinline__ void syrk32_rows(int i8_lo, int i8_hi, int wid, int nw, RP prow, RC crow) {
    const int lane = threadIdx.x & 31;
    const int gq = lane >> 2, kk = lane & 3;
    const int total = (i8_hi * (i8_hi + 1) - i8_lo * (i8_lo + 1)) / 2;
    const int run = (total + nw - 1) / nw;
    int t = wid * run;
    const int t_end = min(total, t + run);
    if (t >= t_end) return;
    int i8 = i8_lo, j8 = t;
    while (j8 > i8) { j8 -= i8 + 1; ++i8; }
    int loaded = -1;
    double av[8];
    for (; t < t_end; ++t) {
        const int i = i8 * 8 + gq;
        if (loaded != i8) {
            const double* pa = prow(i) + kk;
#pragma unroll
            for (int s = 0; s < 8; ++s) av[s] = pa[4 * s];
            loaded = i8;
        }
        const double* pb = prow(j8 * 8 + gq) + kk;
        double bv[8];
#pragma unroll
        for (int s = 0; s < 8; ++s) bv[s] = pb[4 * s];
        double c0 = 0.0, c1 = 0.0, d0 = 0.0, d1 = 0.0;
#pragma unroll
        for (int s = 0; s < 8; s += 2) {
            dmma884(c0, c1, av[s], bv[s]);
            dmma884(d0, d1, av[s + 1], bv[s + 1]);
        }
        double* cr = crow(i);
        const int j = j8 * 8 + 2 * kk;
        // diagonal 8x8 tiles: touch the lower part only (the upper part is reserved for T^T)
        if (j <= i) cr[j] -= c0 + d0;
        if (j + 1 <= i) cr[j + 1] -= c1 + d1;
        if (++j8 > i8) { j8 = 0; ++i8; }
    }
}

// ---- factorisation of one 32x32 diagonal block of the tile: two cooperating warps ----------------------------
// The block (origin D, leading dimension ld, in shared memory) is walked in four 8-column sub-blocks.  The PIVOT
// warp owns the serial chain: one lane factors the 8x8 block in registers, then the warp solves only the next
// eight rows and updates only the next 8x8 diagonal block, and goes on.  The HELPER warp (another scheduler
// partition) solves the remaining rows of the block and applies the rest of the rank-8 update one sub-block
// behind, under the pivot lane's next chain.  Two named-barrier rendezvous per sub-block order the hand-overs.
__device__ __forceinline__ void pair_sync(int id) { asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory"); }

__device__ __forceinline__ void blk8_factor(double* D, int ld, int o, double* rinv_blk, int& bad, int base) {
    // 8x8 diagonal block: the pivot chain, entirely in registers (one lane)
    double m[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j <= i; ++j) m[i][j] = D[(o + i) * ld + o + j];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const double piv = m[j][j];
        if (!(piv > 0.0) && bad == 0) bad = base + o + j + 1;
        const double ri = fast_rsqrt(piv);
        m[j][j] = piv * ri;
        rinv_blk[o + j] = ri;
#pragma unroll
        for (int i = j + 1; i < 8; ++i) m[i][j] *= ri;
#pragma unroll
        for (int i = j + 1; i < 8; ++i)
#pragma unroll
            for (int k = j + 1; k <= i; ++k) m[i][k] = fma(-m[i][j], m[k][j], m[i][k]);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j <= i; ++j) D[(o + i) * ld + o + j] = m[i][j];
}

// one row of the block below the 8x8: x = p L8^-T
__device__ __forceinline__ void blk8_solve_row(double* D, int ld, int o, const double* rinv_blk, int row) {
    double xr[8];
    double* myrow = D + row * ld;
#pragma unroll
    for (int t = 0; t < 8; ++t) xr[t] = myrow[o + t];
#pragma unroll
    for (int t = 0; t < 8; ++t) {
        xr[t] *= rinv_blk[o + t];
#pragma unroll
        for (int u = t + 1; u < 8; ++u) xr[u] = fma(-xr[t], D[(o + u) * ld + o + t], xr[u]);
    }
#pragma unroll
    for (int t = 0; t < 8; ++t) myrow[o + t] = xr[t];
}

// rank-8 update of the lower 8x8 tiles of the remainder (origin o + 8): FIRST = only tile (0,0) (the next
// diagonal block), otherwise every tile of the tile rows >= 1.  All fragments first, then the independent DMMAs,
// then the read-modify-writes.
template <bool FIRST>
__device__ __forceinline__ void blk8_update(double* D, int ld, int o) {
    const int lane = threadIdx.x & 31;
    const double* Px = D + (o + 8) * ld + o;
    double* Cx = D + (o + 8) * ld + o + 8;
    const int m8 = (24 - o) / 8;
    const int gq = lane >> 2, kk = lane & 3;
    double f0[3], f1[3];
#pragma unroll
    for (int t = 0; t < 3; ++t) {
        const bool on = FIRST ? t == 0 : t < m8;
        f0[t] = on ? Px[(t * 8 + gq) * ld + kk] : 0.0;
        f1[t] = on ? Px[(t * 8 + gq) * ld + 4 + kk] : 0.0;
    }
    double acc[6][2];
#pragma unroll
    for (int t = 0; t < 6; ++t) acc[t][0] = acc[t][1] = 0.0;
#pragma unroll
    for (int i8 = 0; i8 < 3; ++i8)
#pragma unroll
        for (int j8 = 0; j8 <= i8; ++j8) {
            const int t = i8 * (i8 + 1) / 2 + j8;
            const bool on = FIRST ? i8 == 0 : (i8 >= 1 && i8 < m8);  // warp-uniform
            if (on) {
                dmma884(acc[t][0], acc[t][1], f0[i8], f0[j8]);
                dmma884(acc[t][0], acc[t][1], f1[i8], f1[j8]);
            }
        }
#pragma unroll
    for (int i8 = 0; i8 < 3; ++i8)
#pragma unroll
        for (int j8 = 0; j8 <= i8; ++j8) {
            const int t = i8 * (i8 + 1) / 2 + j8;
            const bool on = FIRST ? i8 == 0 : (i8 >= 1 && i8 < m8);
            if (on) {
                const int i = i8 * 8 + gq, j = j8 * 8 + 2 * kk;
                if (j <= i) Cx[i * ld + j] -= acc[t][0];
                if (j + 1 <= i) Cx[i * ld + j + 1] -= acc[t][1];
            }
        }
}

// PIVOT = true: the chain warp; false: the helper warp.  Both must be called by all 32 lanes of their warp.
template <bool PIVOT>
__device__ __forceinline__ void diag32_factor(double* D, int ld, double* rinv_blk, int c0, int& bad) {
    const int lane = threadIdx.x & 31;
    if (PIVOT) {
        for (int sb = 0; sb < 4; ++sb) {
            const int o = sb * 8;
            if (lane == 0) blk8_factor(D, ld, o, rinv_blk, bad, c0);
            __syncwarp();
            if (sb < 3) {
                pair_sync(1);  // the helper has finished the previous sub-block's rows and tiles
                if (lane >= o + 8 && lane < o + 16) blk8_solve_row(D, ld, o, rinv_blk, lane);
                __syncwarp();
                blk8_update<true>(D, ld, o);
                __syncwarp();
                pair_sync(2);  // rows o+8..o+15 are solved: the helper may take the rest
            }
        }
    } else {
        for (int sb = 0; sb < 3; ++sb) {
            const int o = sb * 8;
            pair_sync(1);
            pair_sync(2);
            if (lane >= o + 16) blk8_solve_row(D, ld, o, rinv_blk, lane);
            __syncwarp();
            if (o < 16) blk8_update<false>(D, ld, o);
            __syncwarp();
        }
    }
}

// Factor one 128x128 diagonal tile (lower) and invert the factor.  One CTA per tile.
//
// The tile lives in shared memory for the whole kernel:  S lower = L,  S strict upper = T^T
// (T_ij kept at S[j][i]),  rinv = 1 / diag(L) = diag(T).  Everything is hierarchical 8 -> 32 -> 128 so that
// the only serial code is the pivot chain of an 8x8 block in the registers of one lane:
//   factor   4 panels of 32 columns, right-looking.  Inside a panel warp 0 walks four 8-column sub-panels:
//            one lane factors the 8x8 diagonal block, the lanes below solve their row against it, the warp
//            applies the rank-8 update with DMMA.  Then every thread below the panel solves one row by
//            forward substitution and all warps apply the rank-32 update with DMMA.
//   invert   8x8 diagonal blocks in registers, then block doubling T_ba = -T_bb (L_ba T_aa) with DMMA:
//            8 -> 16 -> 32 inside one warp per 32-block, 32 -> 64 -> 128 with all warps.
__global__ void __launch_bounds__(POTF2_THREADS, 1) potf2_kernel(const Potf2Args a) {
    extern __shared__ __align__(16) double sm[];
    double* S = sm;
    double* X = sm + PT * PLD;
    double* rinv = X + 96 * XLD;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = POTF2_THREADS / 32;
    const long long zb = blockIdx.x;
    double* __restrict__ A = a.A + zb * a.strideA;
    const int nb = a.nb;
    POTF2_STAMP(0);

    // load the lower triangle with 16-byte async copies (all in flight at once), then patch the strict
    // upper part to zero and pad the dead rows/columns with identity
    {
        const uint32_t sb = smem_u32(S);
        for (int e = tid; e < PT * (PT / 2); e += POTF2_THREADS) {
            const int r = e >> 6, c = (e & 63) * 2;  // chunk of two columns
            if (c > r) continue;
            int bytes = 0;
            if (r < nb) bytes = (c + 1 < nb) ? 16 : (c < nb ? 8 : 0);
            const double* src = bytes ? A + (long long)r * a.lda + c : A;
            cp_async16(sb + (uint32_t)(r * PLD + c) * 8u, src, bytes);
        }
        cp_async_commit();
        cp_async_wait<0>();
        __syncthreads();
        for (int e = tid; e < PT * (PT / 2); e += POTF2_THREADS) {
            // strict upper part, two columns at a time (the chunk holding the diagonal keeps its first entry)
            const int r = e >> 6, c = (e & 63) * 2;
            if (c + 1 <= r) continue;
            if (c > r) S[r * PLD + c] = 0.0;
            S[r * PLD + c + 1] = 0.0;
        }
        if (tid >= nb && tid < PT) S[tid * PLD + tid] = 1.0;
    }
    __syncthreads();

    POTF2_STAMP(1);
    const bool vec = ((a.lda & 1) == 0) && ((a.ldt & 1) == 0);
    if (a.mode == POTF2_FACTOR) {
        // factor-only (the chain): the explicit zeros of the strict upper part do not depend on the factor,
        // so they leave now, under the pivot chain, instead of lengthening the write-back at the end
        for (int e = tid; e < PT * (PT / 2); e += POTF2_THREADS) {
            const int r = e >> 6, c = (e & 63) * 2;
            if (r >= nb || c >= nb || c <= r) continue;  // chunks wholly above the diagonal
            double* dst = A + (long long)r * a.lda + c;
            if (vec && c + 1 < nb) *reinterpret_cast<double2*>(dst) = make_double2(0.0, 0.0);
            else {
                dst[0] = 0.0;
                if (c + 1 < nb) dst[1] = 0.0;
            }
        }
    }
    int bad = 0;  // 1-based local index of the first non-positive pivot (warp 0, lane 0 only)
    if (a.mode == POTF2_INVERT) {
        if (tid < PT) rinv[tid] = 1.0 / S[tid * PLD + tid];
        __syncthreads();
    }
    for (int jb = 0; jb < (a.mode == POTF2_INVERT ? 0 : 4); ++jb) {
        const int c0 = jb * 32;
        POTF2_STAMP(2 + 3 * jb);
        if (warp == 0) {
            diag32_factor<true>(S + c0 * PLD + c0, PLD, rinv + c0, c0, bad);
        } else if (warp == 1) {
            diag32_factor<false>(S + c0 * PLD + c0, PLD, rinv + c0, c0, bad);
        } else if (jb > 0 && (warp & 3)) {
            // look-ahead: while warps 0 and 1 factor this diagonal block, the other warps finish the previous
            // panel's trailing update (everything but this diagonal block, which was updated first).  Warps of
            // the pivot warp's scheduler partition stay out: its DFMA chain would queue behind their DMMAs in
            // the shared FP64 pipe.
            const double* P = S + c0 * PLD + c0 - 32;
            double* C = S + c0 * PLD + c0;
            syrk32_rows(4, (PT - c0) / 8, warp - 2 - (warp >> 2), nwarps - nwarps / 4 - 1,
                        [&](int r) { return P + r * PLD; }, [&](int r) { return C + r * PLD; });
        }
        __syncthreads();
        POTF2_STAMP(3 + 3 * jb);
        const int rb = c0 + 32;       // first row below the diagonal block
        const int mrows = PT - rb;    // rows below
        if (mrows > 0) {
            // rows below: x = p Ld^-T by forward substitution, one thread per row (Ld broadcast from smem)
            if (tid < mrows) {
                double* prow = S + (rb + tid) * PLD + c0;
                const double* D = S + c0 * PLD + c0;
                double x[32];
#pragma unroll
                for (int k = 0; k < 32; ++k) x[k] = prow[k];
#pragma unroll
                for (int k = 0; k < 32; ++k) {
                    x[k] *= rinv[c0 + k];
#pragma unroll
                    for (int j = k + 1; j < 32; ++j) x[j] = fma(-x[k], D[j * PLD + k], x[j]);
                }
#pragma unroll
                for (int k = 0; k < 32; ++k) prow[k] = x[k];
            }
            __syncthreads();
            POTF2_STAMP(4 + 3 * jb);
            // trailing update, first part: the next diagonal block only (lower 8x8 tiles of S[rb..rb+32, rb..rb+32)
            // -= P P^T, P = S[rb.., c0..c0+32)); the rest follows under the next block's pivot chain
            const double* P = S + rb * PLD + c0;
            double* C = S + rb * PLD + rb;
            syrk32_rows(0, 4, warp, nwarps, [&](int r) { return P + r * PLD; }, [&](int r) { return C + r * PLD; });
            __syncthreads();
        }
    }
    if (warp == 0 && lane == 0 && bad && a.info) {
        int* ip = a.info + zb * a.strideInfo;
        atomicCAS(ip, 0, a.row0 + bad);
    }

    POTF2_STAMP(14);
    // T(i,k) for the already inverted diagonal blocks (tile coordinates)
    auto Tget = [&](int i, int k) -> double {
        return i > k ? S[k * PLD + i] : (i == k ? rinv[i] : 0.0);
    };
    // ---- inverse of the four 32x32 diagonal blocks, one warp each: 8x8 in registers, then 8 -> 16 -> 32 ---
    if (warp < 4) {
        const int c0 = warp * 32;
        double* Xw = X + warp * 16 * XLD;  // per-warp scratch (16 x 16)
        if (lane < 4) {
            const int q0 = c0 + 8 * lane;
            double l[8][8], t[8][8];
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < i; ++j) l[i][j] = S[(q0 + i) * PLD + q0 + j];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                t[j][j] = rinv[q0 + j];
#pragma unroll
                for (int i = j + 1; i < 8; ++i) {
                    double acc = 0.0;
#pragma unroll
                    for (int k = j; k < i; ++k) acc = fma(l[i][k], t[k][j], acc);
                    t[i][j] = -acc * rinv[q0 + i];
                }
            }
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < i; ++j) S[(q0 + j) * PLD + q0 + i] = t[i][j];
        }
        __syncwarp();
        // s = 8: pairs (0,1) and (2,3) of 8-blocks
        for (int pr = 0; pr < 2; ++pr) {
            const int a0 = c0 + 16 * pr, b0 = a0 + 8;
            smem_mma<8>(
                1, 1, 0, 1, [&](int i, int k) { return S[(b0 + i) * PLD + a0 + k]; },
                [&](int j, int k) { return Tget(a0 + k, a0 + j); }, [&](int, int) { return true; },
                [&](int i, int j, double c0v, double c1v) {
                    Xw[(8 * pr + i) * XLD + j] = c0v;
                    Xw[(8 * pr + i) * XLD + j + 1] = c1v;
                });
        }
        __syncwarp();
        for (int pr = 0; pr < 2; ++pr) {
            const int a0 = c0 + 16 * pr, b0 = a0 + 8;
            smem_mma<8>(
                1, 1, 0, 1, [&](int i, int k) { return Tget(b0 + i, b0 + k); },
                [&](int j, int k) { return Xw[(8 * pr + k) * XLD + j]; }, [&](int, int) { return true; },
                [&](int i, int j, double c0v, double c1v) {
                    S[(a0 + j) * PLD + b0 + i] = -c0v;
                    S[(a0 + j + 1) * PLD + b0 + i] = -c1v;
                });
        }
        __syncwarp();
        // s = 16
        {
            const int a0 = c0, b0 = c0 + 16;
            smem_mma<16>(
                2, 2, 0, 1, [&](int i, int k) { return S[(b0 + i) * PLD + a0 + k]; },
                [&](int j, int k) { return Tget(a0 + k, a0 + j); }, [&](int, int) { return true; },
                [&](int i, int j, double c0v, double c1v) {
                    Xw[i * XLD + j] = c0v;
                    Xw[i * XLD + j + 1] = c1v;
                });
            __syncwarp();
            smem_mma<16>(
                2, 2, 0, 1, [&](int i, int k) { return Tget(b0 + i, b0 + k); },
                [&](int j, int k) { return Xw[k * XLD + j]; }, [&](int, int) { return true; },
                [&](int i, int j, double c0v, double c1v) {
                    S[(a0 + j) * PLD + b0 + i] = -c0v;
                    S[(a0 + j + 1) * PLD + b0 + i] = -c1v;
                });
        }
    }
    __syncthreads();
    POTF2_STAMP(15);

    // ---- inverse by doubling: 32 -> 64 (two pairs) -> 128, all warps -----------------------------------
    if (a.mode != POTF2_FACTOR) {
    for (int pr = 0; pr < 2; ++pr) {
        const int a0 = 64 * pr, b0 = a0 + 32;
        double* Xp = X + pr * 32 * XLD;
        smem_mma<32>(
            4, 4, warp, nwarps, [&](int i, int k) { return S[(b0 + i) * PLD + a0 + k]; },
            [&](int j, int k) { return Tget(a0 + k, a0 + j); }, [&](int, int) { return true; },
            [&](int i, int j, double c0v, double c1v) {
                Xp[i * XLD + j] = c0v;
                Xp[i * XLD + j + 1] = c1v;
            });
    }
    __syncthreads();
    for (int pr = 0; pr < 2; ++pr) {
        const int a0 = 64 * pr, b0 = a0 + 32;
        const double* Xp = X + pr * 32 * XLD;
        smem_mma<32>(
            4, 4, warp, nwarps, [&](int i, int k) { return Tget(b0 + i, b0 + k); },
            [&](int j, int k) { return Xp[k * XLD + j]; }, [&](int, int) { return true; },
            [&](int i, int j, double c0v, double c1v) {
                S[(a0 + j) * PLD + b0 + i] = -c0v;
                S[(a0 + j + 1) * PLD + b0 + i] = -c1v;
            });
    }
    __syncthreads();
    POTF2_STAMP(16);
    smem_mma<64>(
        8, 8, warp, nwarps, [&](int i, int k) { return S[(64 + i) * PLD + k]; },
        [&](int j, int k) { return Tget(k, j); }, [&](int, int) { return true; },
        [&](int i, int j, double c0v, double c1v) {
            X[i * XLD + j] = c0v;
            X[i * XLD + j + 1] = c1v;
        });
    __syncthreads();
    smem_mma<64>(
        8, 8, warp, nwarps, [&](int i, int k) { return Tget(64 + i, 64 + k); },
        [&](int j, int k) { return X[k * XLD + j]; }, [&](int, int) { return true; },
        [&](int i, int j, double c0v, double c1v) {
            S[j * PLD + 64 + i] = -c0v;
            S[(j + 1) * PLD + 64 + i] = -c1v;
        });
    __syncthreads();
    }
    POTF2_STAMP(17);

    // ---- write back: L (lower, zero upper) into A; T into Tlo (lower) / Tup (upper) -------------
    // two columns per thread and step: 16-byte stores when the row is even-aligned and fully live
    const bool factor_only = a.mode == POTF2_FACTOR;
    for (int e = tid; e < PT * (PT / 2); e += POTF2_THREADS) {
        const int r = e >> 6, c = (e & 63) * 2;
        if (r >= nb || c >= nb) continue;
        // factor-only: the zeros are already written and T is needed on its diagonal 32-blocks only (those
        // include the zeros above their diagonals, which the substitution solve reads)
        if (factor_only && c > r && (r >> 5) != (c >> 5)) continue;
        const double l0 = c <= r ? S[r * PLD + c] : 0.0;
        const double l1 = c + 1 <= r ? S[r * PLD + c + 1] : 0.0;
        double* dst = A + (long long)r * a.lda + c;
        if (vec && c + 1 < nb) *reinterpret_cast<double2*>(dst) = make_double2(l0, l1);
        else {
            dst[0] = l0;
            if (c + 1 < nb) dst[1] = l1;
        }
        if (a.Tlo && (!factor_only || (r >> 5) == (c >> 5))) {
            double* __restrict__ Tlo = a.Tlo + zb * a.strideT + (long long)r * a.ldt + c;
            double* __restrict__ Tup = a.Tup ? a.Tup + zb * a.strideT + (long long)r * a.ldt + c : nullptr;
            // T[r][c] lives at S[c][r] (r > c); T^T[r][c] = T[c][r] lives at S[r][c] (c > r)
            const double t0 = r > c ? S[c * PLD + r] : (r == c ? rinv[r] : 0.0);
            const double t1 = r > c + 1 ? S[(c + 1) * PLD + r] : (r == c + 1 ? rinv[r] : 0.0);
            const double u0 = c > r ? S[r * PLD + c] : (r == c ? rinv[r] : 0.0);
            const double u1 = c + 1 > r ? S[r * PLD + c + 1] : (r == c + 1 ? rinv[r] : 0.0);
            if (vec && c + 1 < nb) {
                *reinterpret_cast<double2*>(Tlo) = make_double2(t0, t1);
                if (Tup) *reinterpret_cast<double2*>(Tup) = make_double2(u0, u1);
            } else {
                Tlo[0] = t0;
                if (Tup) Tup[0] = u0;
                if (c + 1 < nb) {
                    Tlo[1] = t1;
                    if (Tup) Tup[1] = u1;
                }
            }
        }
    }
    __syncthreads();
    POTF2_STAMP(18);
}


constexpr int PF_THREADS = 256;
constexpr int PF_XLD = 20;
constexpr int PF_SMEM = (TS_LP + 4 * 16 * PF_XLD + PT) * 8;

// Factorisation of a packed 128x128 tile resident in shared memory (S: packed layout pk(r, c), lower part = the
// matrix, strict upper part of the diagonal 32-blocks zero, dead rows padded with identity): on return the lower
// part holds L, the strict upper part of the diagonal blocks T_bb^T (the 32x32 block inverses) and rinv 1 / diag(L).
// X: 4 x 16 x PF_XLD doubles of scratch.  All NT threads of the CTA must call it; bad_out (written by thread 0
// only) is the 1-based local index of the first non-positive pivot, 0 if none.
// `pre(jb, w, nw)` is run by the nw look-ahead warps (index w) at the start of the diagonal phase of panel jb, before
// their share of the previous panel's trailing update: the fused chain step finishes the tile's own K = 128 update
// there, one column block ahead of the factorisation.
struct NoPre {
    __device__ __forceinline__ void operator()(int, int, int) const {}
};
template <int NT, class Pre = NoPre>
__device__ __forceinline__ void factor_packed(double* S, double* X, double* rinv, int& bad_out, Pre pre = Pre()) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = NT / 32;
    int bad = 0;  // 1-based local index of the first non-positive pivot (warp 0, lane 0 only)
    for (int jb = 0; jb < 4; ++jb) {
        const int c0 = jb * 32, ldd = ts_ld(jb);
        double* D = S + pk(c0, c0);  // diagonal block: rows share the leading dimension ldd
        if (warp == 0) {
            diag32_factor<true>(D, ldd, rinv + c0, c0, bad);
        } else if (warp == 1) {
            diag32_factor<false>(D, ldd, rinv + c0, c0, bad);
        } else if (warp & 3) {
            // look-ahead: the other warps finish the previous panel's trailing update (everything but this
            // diagonal block, updated first) while warps 0 and 1 walk the pivot chain (warp 4 shares warp 0's
            // scheduler partition and FP64 pipe: it stays out)
            pre(jb, warp - 2 - (warp >> 2), nwarps - nwarps / 4 - 1);
            if (jb > 0)
                syrk32_rows(4, (PT - c0) / 8, warp - 2 - (warp >> 2), nwarps - nwarps / 4 - 1,
                            [&](int r) { return S + pk(c0 + r, c0 - 32); }, [&](int r) { return S + pk(c0 + r, c0); });
        }
        __syncthreads();
        const int rb = c0 + 32;       // first row below the diagonal block
        const int mrows = PT - rb;    // rows below
        if (mrows > 0) {
            // rows below: x = p Ld^-T by forward substitution, one thread per row (Ld broadcast from smem)
            if (tid < mrows) {
                double* prow = S + pk(rb + tid, c0);
                double x[32];
#pragma unroll
                for (int k = 0; k < 32; ++k) x[k] = prow[k];
#pragma unroll
                for (int k = 0; k < 32; ++k) {
                    x[k] *= rinv[c0 + k];
#pragma unroll
                    for (int j = k + 1; j < 32; ++j) x[j] = fma(-x[k], D[j * ldd + k], x[j]);
                }
#pragma unroll
                for (int k = 0; k < 32; ++k) prow[k] = x[k];
            }
            __syncthreads();
            // trailing update, first part: the next diagonal block only; the rest follows under the next chain
            syrk32_rows(0, 4, warp, nwarps, [&](int r) { return S + pk(rb + r, c0); },
                        [&](int r) { return S + pk(rb + r, rb); });
            __syncthreads();
        }
    }
    if (warp == 0 && lane == 0) bad_out = bad;

    // ---- inverse of the four 32x32 diagonal blocks, one warp each: 8x8 in registers, then 8 -> 16 -> 32 ---
    if (warp < 4) {
        const int c0 = warp * 32, ldd = ts_ld(warp);
        double* D = S + pk(c0, c0);
        const double* rv = rinv + c0;
        // T(i,k) of this block (block coordinates); T_ik (i > k) is kept at D[k][i]
        auto Tg = [&](int i, int k) -> double { return i > k ? D[k * ldd + i] : (i == k ? rv[i] : 0.0); };
        double* Xw = X + warp * 16 * PF_XLD;  // per-warp scratch (16 x 16)
        if (lane < 4) {
            const int q0 = 8 * lane;
            double l[8][8], t[8][8];
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < i; ++j) l[i][j] = D[(q0 + i) * ldd + q0 + j];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                t[j][j] = rv[q0 + j];
#pragma unroll
                for (int i = j + 1; i < 8; ++i) {
                    double acc = 0.0;
#pragma unroll
                    for (int k = j; k < i; ++k) acc = fma(l[i][k], t[k][j], acc);
                    t[i][j] = -acc * rv[q0 + i];
                }
            }
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < i; ++j) D[(q0 + j) * ldd + q0 + i] = t[i][j];
        }
        __syncwarp();
        // s = 8: pairs (0,1) and (2,3) of 8-blocks
        for (int pr = 0; pr < 2; ++pr) {
            const int a0 = 16 * pr, b0 = a0 + 8;
            smem_mma<8>(
                1, 1, 0, 1, [&](int i, int k) { return D[(b0 + i) * ldd + a0 + k]; },
                [&](int j, int k) { return Tg(a0 + k, a0 + j); }, [&](int, int) { return true; },
                [&](int i, int j, double c0v, double c1v) {
                    Xw[(8 * pr + i) * PF_XLD + j] = c0v;
                    Xw[(8 * pr + i) * PF_XLD + j + 1] = c1v;
                });
        }
        __syncwarp();
        for (int pr = 0; pr < 2; ++pr) {
            const int a0 = 16 * pr, b0 = a0 + 8;
            smem_mma<8>(
                1, 1, 0, 1, [&](int i, int k) { return Tg(b0 + i, b0 + k); },
                [&](int j, int k) { return Xw[(8 * pr + k) * PF_XLD + j]; }, [&](int, int) { return true; },
                [&](int i, int j, double c0v, double c1v) {
                    D[(a0 + j) * ldd + b0 + i] = -c0v;
                    D[(a0 + j + 1) * ldd + b0 + i] = -c1v;
                });
        }
        __syncwarp();
        // s = 16
        smem_mma<16>(
            2, 2, 0, 1, [&](int i, int k) { return D[(16 + i) * ldd + k]; },
            [&](int j, int k) { return Tg(k, j); }, [&](int, int) { return true; },
            [&](int i, int j, double c0v, double c1v) {
                Xw[i * PF_XLD + j] = c0v;
                Xw[i * PF_XLD + j + 1] = c1v;
            });
        __syncwarp();
        smem_mma<16>(
            2, 2, 0, 1, [&](int i, int k) { return Tg(16 + i, 16 + k); },
            [&](int j, int k) { return Xw[k * PF_XLD + j]; }, [&](int, int) { return true; },
            [&](int i, int j, double c0v, double c1v) {
                D[j * ldd + 16 + i] = -c0v;
                D[(j + 1) * ldd + 16 + i] = -c1v;
            });
    }
    __syncthreads();
}

// ---- factor-only tile kernel (the chain's and the batched sweeps' version) ----------------------------------
// Same arithmetic as potf2_kernel in POTF2_FACTOR mode (factor + the four 32x32 diagonal-block inverses), but
// the tile is kept packed (lower block rows only, 84 KB) and the CTA has 8 warps, so two tiles share an SM:
// a batched sweep overlaps the serial pivot chain of one tile with the tensor-pipe phases of the other.

__global__ void __launch_bounds__(PF_THREADS, 2) potf2_factor_kernel(const Potf2Args a) {
    extern __shared__ __align__(16) double sm[];
    double* S = sm;                       // packed tile: lower = L, strict upper of the diagonal blocks = T_bb^T
    double* X = sm + TS_LP;               // per-warp 16 x 16 scratch of the block inversion
    double* rinv = X + 4 * 16 * PF_XLD;
    const int tid = threadIdx.x;
    const long long zb = blockIdx.x;
    double* __restrict__ A = a.A + zb * a.strideA;
    const int nb = a.nb;
    {
        const uint32_t sb = smem_u32(S);
        for (int e = tid; e < PT * (PT / 2); e += PF_THREADS) {
            const int r = e >> 6, c = (e & 63) * 2;  // chunk of two columns
            if (c > r) continue;
            int bytes = 0;
            if (r < nb) bytes = (c + 1 < nb) ? 16 : (c < nb ? 8 : 0);
            const double* src = bytes ? A + (long long)r * a.lda + c : A;
            cp_async16(sb + (uint32_t)pk(r, c) * 8u, src, bytes);
        }
        cp_async_commit();
        cp_async_wait<0>();
        __syncthreads();
        // strict upper part of the diagonal blocks (the chunk holding the diagonal keeps its first entry)
        for (int e = tid; e < PT * 16; e += PF_THREADS) {
            const int r = e >> 4, c = (r & ~31) + (e & 15) * 2;
            if (c + 1 <= r) continue;
            if (c > r) S[pk(r, c)] = 0.0;
            S[pk(r, c + 1)] = 0.0;
        }
        if (tid >= nb && tid < PT) S[pk(tid, tid)] = 1.0;
    }
    __syncthreads();

    int bad = 0;  // 1-based local index of the first non-positive pivot (thread 0 only)
    factor_packed<PF_THREADS>(S, X, rinv, bad);
    if (tid == 0 && bad && a.info) {
        int* ip = a.info + zb * a.strideInfo;
        atomicCAS(ip, 0, a.row0 + bad);
    }

    // ---- write back: L (lower, zero upper) into A; the diagonal blocks of T into Tlo ------------------------
    const bool vec = ((a.lda & 1) == 0) && ((a.ldt & 1) == 0);
    for (int e = tid; e < PT * (PT / 2); e += PF_THREADS) {
        const int r = e >> 6, c = (e & 63) * 2;
        if (r >= nb || c >= nb) continue;
        const double l0 = c <= r ? S[pk(r, c)] : 0.0;
        const double l1 = c + 1 <= r ? S[pk(r, c + 1)] : 0.0;
        double* dst = A + (long long)r * a.lda + c;
        if (vec && c + 1 < nb) *reinterpret_cast<double2*>(dst) = make_double2(l0, l1);
        else {
            dst[0] = l0;
            if (c + 1 < nb) dst[1] = l1;
        }
        if (a.Tlo && (r >> 5) == (c >> 5)) {
            // T[r][c] (r > c, same diagonal block) lives at S[c][r]
            double* __restrict__ Tlo = a.Tlo + zb * a.strideT + (long long)r * a.ldt + c;
            const double t0 = r > c ? S[pk(c, r)] : (r == c ? rinv[r] : 0.0);
            const double t1 = r > c + 1 ? S[pk(c + 1, r)] : (r == c + 1 ? rinv[r] : 0.0);
            if (vec && c + 1 < nb) *reinterpret_cast<double2*>(Tlo) = make_double2(t0, t1);
            else {
                Tlo[0] = t0;
                if (c + 1 < nb) Tlo[1] = t1;
            }
        }
    }
}

static int launch_potf2(const Potf2Args& a, int batch, cudaStream_t stream) {
    static unsigned long long configured = 0;  // one bit per device: the attribute is per context
    int dev = 0;
    cudaGetDevice(&dev);
    if (!((configured >> (dev & 63)) & 1ull)) {
        if (cudaFuncSetAttribute(potf2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, POTF2_SMEM) !=
            cudaSuccess)
            return GPMP_ERR_CUDA;
        configured |= 1ull << (dev & 63);
    }
    LaunchScope scope(KC_POTF2, (double)batch * (PT * (double)PT * PT), stream);  // n^3/3 + 2n^3/3
    // more tiles than SMs: the packed 8-warp kernel (two tiles per SM); otherwise the 16-warp kernel (latency)
    if (a.mode == POTF2_FACTOR && batch > 148) {
        static unsigned long long configured_f = 0;
        if (!((configured_f >> (dev & 63)) & 1ull)) {
            if (cudaFuncSetAttribute(potf2_factor_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PF_SMEM) !=
                cudaSuccess)
                return GPMP_ERR_CUDA;
            configured_f |= 1ull << (dev & 63);
        }
        potf2_factor_kernel<<<batch, PF_THREADS, PF_SMEM, stream>>>(a);
        GPMP_CHECK_LAUNCH();
        return GPMP_OK;
    }
    potf2_kernel<<<batch, POTF2_THREADS, POTF2_SMEM, stream>>>(a);
    GPMP_CHECK_LAUNCH();
    return GPMP_OK;
}

// development hook (not part of the C-ABI header): one tile with phase timestamps
int debug_potf2(double* A, long long lda, int nb, double* Tlo, double* Tup, int* info, long long* dbg,
                cudaStream_t stream) {
    Potf2Args pa;
    pa.A = A; pa.lda = lda; pa.strideA = 0; pa.Tlo = Tlo; pa.Tup = Tup; pa.ldt = PT; pa.strideT = 0; pa.nb = nb;
    pa.info = info; pa.strideInfo = 0; pa.row0 = 0; pa.dbg = dbg; pa.mode = POTF2_FULL;
    return launch_potf2(pa, 1, stream);
}

// ---- panel solve by blocked substitution (the chain's solve: needs only the 32x32 block inverses) ---------
// X = P L^-T for the rows below a 128-wide tile, in blocks of 64 rows, one 8-row strip per warp.  Rows are
// independent, so after the operands are staged every warp walks its strip through the four 32-column block
// steps on its own:  R_b = P_b - sum_{c<b} X_c L_bc^T  (DMMA, K = 32 b),  X_b = R_b T_bb^T  (DMMA, K = 32).
// A CTA stages the operand tile once (lower block rows only, packed) and then walks `blocks_per_cta`
// consecutive row blocks with its two halves (8 warps, one buffer and one named barrier each) taking alternate
// blocks independently, so the load / store phases of one half overlap the tensor-pipe phase of the other:
// the chain launches one block per CTA (latency), batched evaluations one CTA per matrix (throughput).
// Output: X into the group panel buffer, into A in place, and mirrored into the upper tiles (one pass).
// Batched value-only runs also pass the few extra (whitening) rows of the matrix separately: a seventeenth warp
// solves them as one more strip and every block then applies  E[:, rows] -= X_E X_rows^T  itself, so neither this
// solve nor the trailing update carries a ragged 2-row tile per matrix.
constexpr int TS_ROWS = 64, TS_HALF_THREADS = 256, TS_THREADS = 2 * TS_HALF_THREADS + 32, TS_LD = 132;  // 132 = 4 mod 16: conflict-free DMMA fragments
constexpr int TS_EXTRA = 8;   // extra rows the kernel can carry (one strip)
constexpr int TS_SMEM = (TS_LP + 2 * TS_ROWS * TS_LD + TS_EXTRA * TS_LD) * 8;

struct TrsmTileArgs {
    double* P; long long lda; long long strideA;      // rows below the tile in A (in/out)
    const double* Ltile;                              // the tile's factor in A (lower), same lda / strideA
    const double* Tsub; long long ldt; long long strideT;  // block-diagonal T: 32x32 inverses of the diagonal blocks
    double* W; long long ldw; long long strideW;      // panel buffer rows (out)
    double* Aup;                                      // mirror origin: Aup[j * lda + r] = X[r][j]
    int M, nb, mirror_rows;
    int blocks_per_cta;   // consecutive 64-row blocks walked by one CTA
    int inplace_from;     // rows below this index are not written back into A (value-only batched runs need
                          // only the panel buffer for them)
    int block_rows;       // rows per block: 64, or 32 for launches with few blocks (half the work per CTA, twice
                          // the CTAs: the chain's solves are latency-bound)
    double* E;            // optional extra rows at the tile's columns (E[i * lda + c], same strideA): solved in
    int nextra;           // place; their columns right of the tile (E + nb + r) receive -X_E X_r^T for every row r
};

__device__ __forceinline__ void half_barrier(int half) {
    asm volatile("bar.sync %0, %1;" ::"r"(1 + half), "r"(TS_HALF_THREADS) : "memory");
}

__global__ void __launch_bounds__(TS_THREADS, 1) trsm_tile_kernel(const TrsmTileArgs a) {
    extern __shared__ __align__(16) double sm[];
    double* Lp = sm;                     // packed operand: L_bc below the diagonal blocks, T_bb on them
    double* Xbuf = sm + TS_LP;           // one buffer of [64][TS_LD] per half
    double* Es = Xbuf + 2 * TS_ROWS * TS_LD;  // [8][TS_LD]: the extra rows' strip
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int half = warp >> 3, hw = warp & 7, htid = tid & (TS_HALF_THREADS - 1);  // warp 16: half == 2
    const long long zb = blockIdx.z;
    double* __restrict__ P = a.P + zb * a.strideA;
    const double* __restrict__ Lt = a.Ltile + zb * a.strideA;
    const double* __restrict__ Ts = a.Tsub + zb * a.strideT;
    double* __restrict__ W = a.W + zb * a.strideW;
    const int nb = a.nb;
    const int blk0 = blockIdx.x * a.blocks_per_cta;
    const int nextra = a.E ? a.nextra : 0;
    double* __restrict__ E = a.E ? a.E + zb * a.strideA : nullptr;
    const int R = a.block_rows, rshift = R == 64 ? 6 : 5;
    int nblocks = min(a.blocks_per_cta, ceil_div(a.M, R) - blk0);
    if (nblocks <= 0) {
        if (nextra == 0 || blockIdx.x != 0) return;
        nblocks = 1;  // no rows below the tile: one pass for the extra rows only
    }
    double* Xs = Xbuf + (half & 1) * TS_ROWS * TS_LD;
    // stage the rows of one block into this half's buffer (16-byte async copies, zero fill beyond M / nb)
    auto stage_rows = [&](int blk) {
        const uint32_t xb = smem_u32(Xs);
        const int r0 = blk * R;
        for (int e = htid; e < R * (PT / 2); e += TS_HALF_THREADS) {
            const int r = e >> 6, c = (e & 63) * 2;
            int bytes = 0;
            if (r0 + r < a.M) bytes = (c + 1 < nb) ? 16 : (c < nb ? 8 : 0);
            const double* src = bytes ? P + (long long)(r0 + r) * a.lda + c : P;
            cp_async16(xb + (uint32_t)(r * TS_LD + c) * 8u, src, bytes);
        }
    };
    {
        const uint32_t lb = smem_u32(Lp);
        for (int e = tid; e < PT * (PT / 2); e += TS_THREADS) {
            const int i = e >> 6, c = (e & 63) * 2;
            const int b = i >> 5;
            if (c >= 32 * (b + 1)) continue;  // blocks above the diagonal blocks are never read
            const bool diag_blk = (c >> 5) == b;
            int bytes = 0;
            if (i < nb) bytes = (c + 1 < nb) ? 16 : (c < nb ? 8 : 0);
            const double* src = diag_blk ? Ts + (long long)i * a.ldt + c : Lt + (long long)i * a.lda + c;
            cp_async16(lb + (uint32_t)(ts_base(b) + (i - 32 * b) * ts_ld(b) + c) * 8u, bytes ? src : Lt, bytes);
        }
    }
    if (nextra) {
        const uint32_t eb = smem_u32(Es);
        for (int e = tid; e < TS_EXTRA * (PT / 2); e += TS_THREADS) {
            const int r = e >> 6, c = (e & 63) * 2;
            int bytes = 0;
            if (r < nextra) bytes = (c + 1 < nb) ? 16 : (c < nb ? 8 : 0);
            cp_async16(eb + (uint32_t)(r * TS_LD + c) * 8u, bytes ? E + (long long)r * a.lda + c : E, bytes);
        }
    }
    if (half < 2 && half < nblocks) stage_rows(blk0 + half);
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();
    if (nb < PT) {
        // dead rows / columns of a ragged tile behave like an identity block
        for (int i = nb + tid; i < PT; i += TS_THREADS) {
            const int b = i >> 5;
            Lp[ts_base(b) + (i - 32 * b) * ts_ld(b) + i] = 1.0;
        }
        __syncthreads();
    }
    // the chunk that holds a diagonal entry of T also carries the entry right of it (zero in Tsub): fine.
    const int gq = lane >> 2, kk = lane & 3;
    // X = R L^-T for the 8-row strip whose lane row is xrow (A-fragment row and C row), in place
    auto solve_strip = [&](double* xrow) {
        for (int b = 0; b < 4; ++b) {
            const int cb = 32 * b, ldb = ts_ld(b);
            const double* Lb = Lp + ts_base(b) + gq * ldb + kk;  // + 8 j8 ldb + k
            double acc[4][2];
#pragma unroll
            for (int j8 = 0; j8 < 4; ++j8) acc[j8][0] = acc[j8][1] = 0.0;
            // sum_{k < 32 b} X[i][k] L[cb + j][k]
            for (int k0 = 0; k0 < cb; k0 += 32) {
                // 32 columns per round: all fragments first, then the 32 DMMAs (four independent chains)
                double av[8], bv[8][4];
#pragma unroll
                for (int s = 0; s < 8; ++s) {
                    av[s] = xrow[k0 + 4 * s + kk];
#pragma unroll
                    for (int j8 = 0; j8 < 4; ++j8) bv[s][j8] = Lb[8 * j8 * ldb + k0 + 4 * s];
                }
#pragma unroll
                for (int s = 0; s < 8; ++s)
#pragma unroll
                    for (int j8 = 0; j8 < 4; ++j8) dmma884(acc[j8][0], acc[j8][1], av[s], bv[s][j8]);
            }
            // R = P_b - acc, written back in place (C layout: row gq, columns 2kk, 2kk+1 of each 8-column tile)
#pragma unroll
            for (int j8 = 0; j8 < 4; ++j8) {
                xrow[cb + 8 * j8 + 2 * kk] -= acc[j8][0];
                xrow[cb + 8 * j8 + 2 * kk + 1] -= acc[j8][1];
                acc[j8][0] = acc[j8][1] = 0.0;
            }
            __syncwarp();
            // X_b = R T_bb^T : sum_k R[i][k] T_bb[j][k]; T_bb is lower, so the 8-column tile j8 stops at k = 8 (j8 + 1)
#pragma unroll
            for (int s = 0; s < 8; ++s) {
                const double av = xrow[cb + 4 * s + kk];
#pragma unroll
                for (int j8 = 0; j8 < 4; ++j8)
                    if (s < 2 * (j8 + 1)) dmma884(acc[j8][0], acc[j8][1], av, Lb[8 * j8 * ldb + cb + 4 * s]);
            }
            __syncwarp();
#pragma unroll
            for (int j8 = 0; j8 < 4; ++j8) {
                xrow[cb + 8 * j8 + 2 * kk] = acc[j8][0];
                xrow[cb + 8 * j8 + 2 * kk + 1] = acc[j8][1];
            }
            __syncwarp();
        }
    };
    if (half == 2) {
        // seventeenth warp: the extra rows' strip, once
        if (nextra) {
            solve_strip(Es + gq * TS_LD);
            __syncwarp();
        }
        __syncthreads();
        if (nextra && blockIdx.x == 0) {
            // the solved extra rows themselves, in place
            for (int e = lane; e < nextra * PT; e += 32) {
                const int i = e >> 7, c = e & (PT - 1);
                if (c < nb) E[(long long)i * a.lda + c] = Es[i * TS_LD + c];
            }
        }
        return;
    }
    // the two halves (8 warps each, own buffer, own barrier) walk alternate row blocks independently: the
    // load / store phases of one overlap the tensor-pipe phase of the other
    bool first = true;
    for (int it = half; it < nblocks || first; it += 2) {
        const bool live = it < nblocks;
        const int r0 = (blk0 + it) * R;
        if (live) {
            if (!first) {
                stage_rows(blk0 + it);
                cp_async_commit();
                cp_async_wait<0>();
                half_barrier(half);
            }
            if (hw * 8 < R) solve_strip(Xs + (hw * 8 + gq) * TS_LD);
        }
        if (first) __syncthreads();  // the extra strip is solved (every warp passes here exactly once)
        else half_barrier(half);
        first = false;
        if (!live) break;
        // write out: panel buffer + A in place (rows of this block), then the mirrored upper tiles
        for (int e = htid; e < R * (PT / 2); e += TS_HALF_THREADS) {
            const int r = e >> 6, c = (e & 63) * 2;
            if (r0 + r >= a.M || c >= nb) continue;
            const double v0 = Xs[r * TS_LD + c], v1 = Xs[r * TS_LD + c + 1];
            double* wp = W + (long long)(r0 + r) * a.ldw + c;
            double* ap = P + (long long)(r0 + r) * a.lda + c;
            const bool inplace = r0 + r >= a.inplace_from;
            if (c + 1 < nb) {
                *reinterpret_cast<double2*>(wp) = make_double2(v0, v1);
                if (inplace) *reinterpret_cast<double2*>(ap) = make_double2(v0, v1);
            } else {
                wp[0] = v0;
                if (inplace) ap[0] = v0;
            }
        }
        if (a.Aup && r0 < a.mirror_rows) {
            double* __restrict__ Aup = a.Aup + zb * a.strideA;
            for (int e = htid; e < PT * R; e += TS_HALF_THREADS) {
                const int j = e >> rshift, r = e & (R - 1);  // consecutive threads -> consecutive rows r: contiguous in Aup
                if (j < nb && r0 + r < a.mirror_rows) Aup[(long long)j * a.lda + r0 + r] = Xs[r * TS_LD + j];
            }
        }
        if (nextra && hw * 8 < R && r0 + hw * 8 < a.M) {
            // extra rows: their part right of the tile, columns of this block:  E[i][nb + r0 + r] -= X_E[i] . X[r]
            // one 8x8 output tile per warp (rows = extra rows, columns = the warp's own strip), K = 128 on DMMA
            const double* er = Es + gq * TS_LD + kk;
            const double* xr = Xs + (hw * 8 + gq) * TS_LD + kk;
            double c[4][2];
#pragma unroll
            for (int q = 0; q < 4; ++q) c[q][0] = c[q][1] = 0.0;
#pragma unroll
            for (int k0 = 0; k0 < PT; k0 += 32) {
                double av[8], bv[8];
#pragma unroll
                for (int q = 0; q < 8; ++q) { av[q] = er[k0 + 4 * q]; bv[q] = xr[k0 + 4 * q]; }
#pragma unroll
                for (int q = 0; q < 8; ++q) dmma884(c[q & 3][0], c[q & 3][1], av[q], bv[q]);
            }
            const double s0 = (c[0][0] + c[1][0]) + (c[2][0] + c[3][0]);
            const double s1 = (c[0][1] + c[1][1]) + (c[2][1] + c[3][1]);
            const int r = hw * 8 + 2 * kk;
            if (gq < nextra) {
                double* ep = E + (long long)gq * a.lda + nb + r0 + r;
                if (r0 + r < a.M) ep[0] -= s0;
                if (r0 + r + 1 < a.M) ep[1] -= s1;
            }
        }
        half_barrier(half);  // the buffer is refilled by the next iteration of this half
    }
}

static int launch_trsm_tile(TrsmTileArgs a, int batch, cudaStream_t stream) {
    if (a.M < 0) a.M = 0;
    if (a.E == nullptr) a.nextra = 0;
    if (a.M <= 0 && a.nextra <= 0) return GPMP_OK;
    static unsigned long long configured = 0;
    int dev = 0;
    cudaGetDevice(&dev);
    if (!((configured >> (dev & 63)) & 1ull)) {
        if (cudaFuncSetAttribute(trsm_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TS_SMEM) !=
            cudaSuccess)
            return GPMP_ERR_CUDA;
        configured |= 1ull << (dev & 63);
    }
    LaunchScope scope(KC_GEMM, 2.0 * (double)a.M * PT * PT * batch, stream);
    // enough CTAs for two per SM (148 SMs); beyond that every CTA walks several row blocks of its matrix
    a.block_rows = (batch == 1 && a.nextra == 0 && ceil_div(a.M, TS_ROWS) <= 148) ? 32 : TS_ROWS;
    const int nblocks = max(1, ceil_div(a.M, a.block_rows));
    // (one CTA per matrix when it carries extra rows: it rewrites them in place)
    const int splits = a.nextra > 0 ? 1 : min(nblocks, max(1, ceil_div(2 * 148, batch)));
    a.blocks_per_cta = ceil_div(nblocks, splits);
    dim3 grid(ceil_div(nblocks, a.blocks_per_cta), 1, batch);
    trsm_tile_kernel<<<grid, TS_THREADS, TS_SMEM, stream>>>(a);
    GPMP_CHECK_LAUNCH();
    return GPMP_OK;
}

// ---- fused chain step: one launch per 128-column step of a column group -------------------------------------------
// For the factored tile k (columns [col, col + 128), factor and 32x32 block inverses already in place) one launch
// does the whole right-looking step inside the group:
//   solve     X = P L_k^-T for every row below the tile (blocked substitution, as trsm_tile_kernel), written into A
//             in place and mirrored into the upper tiles;
//   update    the `nrem` 128-column blocks that remain in the group receive -X X_j^T (K = 128, lower 8x8 tiles
//             only), X_j = the solved rows of tile j's own row block;
//   factor    tile k + 1 is factored (factor + 32x32 block inverses, packed in shared memory) as soon as its own 128
//             rows are solved and its diagonal tile is updated.
// CTA 0 (the "narrow" CTA) owns the rows of tile k + 1: it solves them, publishes them, applies the K = 128 update to
// the diagonal tile from shared memory and factors it -- the only serial dependency between two steps.  The other
// ("wide") CTAs take 64 rows each from the rows below; the ones whose rows are row blocks of the group's later
// tiles publish their solved rows too.  A consumer waits for a producer through a flag in global memory
// (st.release / ld.acquire, gpu scope); producers are always CTAs with a lower block index than any CTA that can
// be left waiting for an SM (indices 0 .. 2 nrem - 2), so the waits cannot deadlock.
// Replaces, per step, the tile kernel + the panel solve + the in-group K = 128 GEMMs (3-4 dependent launches).
struct ChainStepArgs {
    double* A; long long lda;
    int n, nrows;            // order of the matrix; rows including the extra (whitening) rows below it
    int col;                 // first column of tile k
    int nrem;                // full 128-column tiles of the group after tile k (0: solve only)
    const double* Tsub_k;    // 128 x 128 (ld 128): the 32x32 diagonal-block inverses of tile k
    double* Tsub_next;       // out: the same for tile k + 1 (nrem >= 1)
    int* info;
    unsigned int* flags;     // ring of CS_FLAGS words
    unsigned int flag_id;    // producer p publishes flags[(flag_id + p) & (CS_FLAGS - 1)] = flag_id + p
    int stamps;              // development: the narrow CTA and wide CTA 0 record phase clocks
};
constexpr int CS_THREADS = 512;
constexpr int CS_FLAGS = 1 << 16;
constexpr int CS_SMEM = (TS_LP + PT * TS_LD + PT) * 8;
__device__ unsigned int g_chain_flags[CS_FLAGS];
__device__ long long g_cs_stamps[64];  // development: clock64() at the phase boundaries of the last stamped launch
#define CS_STAMP(i)                                                 \
    do {                                                            \
        if (stamp_base >= 0 && threadIdx.x == 0) g_cs_stamps[stamp_base + (i)] = clock64(); \
    } while (0)

__device__ __forceinline__ unsigned int ld_acquire_u32(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_u32(unsigned int* p, unsigned int v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// -x through the integer pipe (the FP64 pipe is the one the DMMAs need)
__device__ __forceinline__ double flip_sign(double x) {
    return __hiloint2double(__double2hiint(x) ^ (int)0x80000000, __double2loint(x));
}

// X = R L^-T for one 8-row strip (lane row xrow: A-fragment row and C row), in place; Lp = the packed operand
// (L_bc below the diagonal blocks, T_bb on them) -- the substitution of trsm_tile_kernel
__device__ __forceinline__ void solve_strip_packed(const double* Lp, double* xrow) {
    const int lane = threadIdx.x & 31, gq = lane >> 2, kk = lane & 3;
    for (int b = 0; b < 4; ++b) {
        const int cb = 32 * b, ldb = ts_ld(b);
        const double* Lb = Lp + ts_base(b) + gq * ldb + kk;
        double acc[4][2];
#pragma unroll
        for (int j8 = 0; j8 < 4; ++j8) acc[j8][0] = acc[j8][1] = 0.0;
        for (int k0 = 0; k0 < cb; k0 += 32) {
            double av[8], bv[8][4];
#pragma unroll
            for (int s = 0; s < 8; ++s) {
                av[s] = xrow[k0 + 4 * s + kk];
#pragma unroll
                for (int j8 = 0; j8 < 4; ++j8) bv[s][j8] = Lb[8 * j8 * ldb + k0 + 4 * s];
            }
#pragma unroll
            for (int s = 0; s < 8; ++s)
#pragma unroll
                for (int j8 = 0; j8 < 4; ++j8) dmma884(acc[j8][0], acc[j8][1], av[s], bv[s][j8]);
        }
#pragma unroll
        for (int j8 = 0; j8 < 4; ++j8) {
            xrow[cb + 8 * j8 + 2 * kk] -= acc[j8][0];
            xrow[cb + 8 * j8 + 2 * kk + 1] -= acc[j8][1];
            acc[j8][0] = acc[j8][1] = 0.0;
        }
        __syncwarp();
#pragma unroll
        for (int s = 0; s < 8; ++s) {
            const double av = xrow[cb + 4 * s + kk];
#pragma unroll
            for (int j8 = 0; j8 < 4; ++j8)
                if (s < 2 * (j8 + 1)) dmma884(acc[j8][0], acc[j8][1], av, Lb[8 * j8 * ldb + cb + 4 * s]);
        }
        __syncwarp();
#pragma unroll
        for (int j8 = 0; j8 < 4; ++j8) {
            xrow[cb + 8 * j8 + 2 * kk] = acc[j8][0];
            xrow[cb + 8 * j8 + 2 * kk + 1] = acc[j8][1];
        }
        __syncwarp();
    }
}

__global__ void __launch_bounds__(CS_THREADS, 1) chain_step_kernel(const ChainStepArgs a) {
    extern __shared__ __align__(16) double sm[];
    double* Lp = sm;                 // packed operand of the solve; later: B rows (wide) / the packed next tile (narrow)
    double* Xs = sm + TS_LP;         // [128][TS_LD]: the solved rows; wide CTAs use rows 0..63, rows 64.. hold B rows
    double* rinv = Xs + PT * TS_LD;  // [128]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int gq = lane >> 2, kk = lane & 3;
    const long long lda = a.lda;
    double* A = a.A;
    const int col = a.col, rb = col + PT;
    const bool has_narrow = a.nrem >= 1;
    const bool narrow = has_narrow && blockIdx.x == 0;
    if (has_narrow && blockIdx.x == gridDim.x - 1) {
        // last CTA: the mirrored upper tile of the narrow CTA's rows (off the chain: nothing in this launch reads it)
        if (tid == 0) {
            const unsigned int* f = a.flags + (a.flag_id & (CS_FLAGS - 1));
            while (ld_acquire_u32(f) != a.flag_id) __nanosleep(256);
        }
        __syncthreads();
        const uint32_t xb = smem_u32(Xs);
        const double* P = A + (long long)rb * lda + col;
        for (int e = tid; e < PT * (PT / 2); e += CS_THREADS) {
            const int r = e >> 6, c = (e & 63) * 2;
            cp_async16(xb + (uint32_t)(r * TS_LD + c) * 8u, P + (long long)r * lda + c, 16);
        }
        cp_async_commit();
        cp_async_wait<0>();
        __syncthreads();
        for (int e = tid; e < PT * PT; e += CS_THREADS) {
            const int j = e >> 7, r = e & (PT - 1);
            A[(long long)(col + j) * lda + rb + r] = Xs[r * TS_LD + j];
        }
        return;
    }
    const int wb = (int)blockIdx.x - (has_narrow ? 1 : 0);               // wide block index
    const int r0 = narrow ? rb : (has_narrow ? rb + PT : rb) + 64 * wb;  // first global row of this CTA
    const int R = narrow ? PT : 64;
    const int live = min(R, a.nrows - r0);                               // rows that exist
    if (live <= 0) return;
    const int stamp_base = a.stamps ? (narrow ? 0 : (wb == 0 ? 16 : -1)) : -1;
    CS_STAMP(0);

    // ---- stage the operand tile (packed) and this CTA's rows ---------------------------------------------------
    {
        const double* Lt = A + (long long)col * (lda + 1);
        const uint32_t lb = smem_u32(Lp);
        for (int e = tid; e < PT * (PT / 2); e += CS_THREADS) {
            const int i = e >> 6, c = (e & 63) * 2;
            const int b = i >> 5;
            if (c >= 32 * (b + 1)) continue;
            const bool diag_blk = (c >> 5) == b;
            const double* src = diag_blk ? a.Tsub_k + (long long)i * PT + c : Lt + (long long)i * lda + c;
            cp_async16(lb + (uint32_t)(ts_base(b) + (i - 32 * b) * ts_ld(b) + c) * 8u, src, 16);
        }
        const uint32_t xb = smem_u32(Xs);
        const double* P = A + (long long)r0 * lda + col;
        for (int e = tid; e < R * (PT / 2); e += CS_THREADS) {
            const int r = e >> 6, c = (e & 63) * 2;
            const bool on = r < live;
            cp_async16(xb + (uint32_t)(r * TS_LD + c) * 8u, on ? P + (long long)r * lda + c : P, on ? 16 : 0);
        }
        cp_async_commit();
        cp_async_wait<0>();
        __syncthreads();
    }
    CS_STAMP(1);
    // ---- solve: one 8-row strip per warp -------------------------------------------------------------------------
    if (warp * 8 < R) solve_strip_packed(Lp, Xs + (warp * 8 + gq) * TS_LD);

    if (narrow) {
        // The diagonal tile's own update, C - X X^T, in units of 2 x 2 8x8 tiles (16 x 16 entries, one warp each; the
        // 36 lower units of the 8 x 8 unit grid).  Only the first 32 columns (15 units) are needed before the
        // factorisation starts; the units of columns 32..63 (11) and 64..127 (10) are computed by the look-ahead warps
        // of the tile factorisation under the pivot chains of its first and second panel.
        const double* Cd = A + (long long)rb * (lda + 1);
        double* S = Lp;  // the packed tile takes the place of the solve's operand
        auto unit_load = [&](int ui, int uj, double (&acc)[2][2][2]) {
#pragma unroll
            for (int i = 0; i < 2; ++i)
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const double2 v = *reinterpret_cast<const double2*>(Cd + (long long)((2 * ui + i) * 8 + gq) * lda +
                                                                        (2 * uj + j) * 8 + 2 * kk);
                    acc[i][j][0] = v.x; acc[i][j][1] = v.y;
                }
        };
        auto unit_run = [&](int ui, int uj, double (&acc)[2][2][2]) {
            const double* ap = Xs + (16 * ui + gq) * TS_LD + kk;
            const double* bp = Xs + (16 * uj + gq) * TS_LD + kk;
            for (int k0 = 0; k0 < PT; k0 += 32) {
                double fa[2][8], fb[2][8];
#pragma unroll
                for (int s = 0; s < 8; ++s) {
#pragma unroll
                    for (int i = 0; i < 2; ++i) {
                        fa[i][s] = flip_sign(ap[8 * i * TS_LD + k0 + 4 * s]);
                        fb[i][s] = bp[8 * i * TS_LD + k0 + 4 * s];
                    }
                }
#pragma unroll
                for (int s = 0; s < 8; ++s)
#pragma unroll
                    for (int i = 0; i < 2; ++i)
#pragma unroll
                        for (int j = 0; j < 2; ++j) dmma884(acc[i][j][0], acc[i][j][1], fa[i][s], fb[j][s]);
            }
#pragma unroll
            for (int i = 0; i < 2; ++i)
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const int r = (2 * ui + i) * 8 + gq, c = (2 * uj + j) * 8 + 2 * kk;
                    if (c <= r) S[pk(r, c)] = acc[i][j][0];
                    if (c + 1 <= r) S[pk(r, c + 1)] = acc[i][j][1];
                }
        };
        // first 32 columns: units (ui, 0), ui = 0..7 -> warps 0..7; (ui, 1), ui = 1..7 -> warps 8..14
        const bool u_on = warp < 15;
        const int u_i = warp < 8 ? warp : warp - 7, u_j = warp < 8 ? 0 : 1;
        double acc0[2][2][2];
        if (u_on) unit_load(u_i, u_j, acc0);  // (the latency hides under the write-out)
        __syncthreads();  // every strip is solved
        CS_STAMP(2);
        // publish the solved rows: A in place (the mirrored upper tile is written by the last CTA of the grid)
        for (int e = tid; e < PT * (PT / 2); e += CS_THREADS) {
            const int r = e >> 6, c = (e & 63) * 2;
            *reinterpret_cast<double2*>(A + (long long)(rb + r) * lda + col + c) =
                make_double2(Xs[r * TS_LD + c], Xs[r * TS_LD + c + 1]);
        }
        __syncthreads();  // the solve's operand in Lp is dead from here on: the units write the packed tile there
        if (u_on) unit_run(u_i, u_j, acc0);
        // the stores above have long left by now: the fence is cheap here, and the wide CTAs have slack
        __threadfence();
        __syncthreads();
        if (tid == 0) st_release_u32(a.flags + (a.flag_id & (CS_FLAGS - 1)), a.flag_id);
        CS_STAMP(3);
        // strict upper part of the diagonal 32-blocks: zero (as the packed tile kernel expects)
        for (int e = tid; e < PT * 32; e += CS_THREADS) {
            const int r = e >> 5, c = (r & ~31) + (e & 31);
            if (c > r) S[pk(r, c)] = 0.0;
        }
        __syncthreads();
        CS_STAMP(4);
        int bad = 0;
        factor_packed<CS_THREADS>(S, Xs + PT * TS_LD - 4 * 16 * PF_XLD, rinv, bad, [&](int jb, int w, int nw) {
            // columns 32..63 under the first pivot chain: units (ui, 2), ui = 2..7, (ui, 3), ui = 3..7 (11 = nw);
            // columns 64..127 under the second: (ui, uj), 4 <= uj <= ui <= 7 (10)
            if (jb > 1) return;
            int ui = -1, uj = 0;
            if (jb == 0) {
                if (w < 6) { ui = 2 + w; uj = 2; }
                else if (w < 11) { ui = w - 3; uj = 3; }
            } else if (w < 10) {
                uj = w < 4 ? 4 : (w < 7 ? 5 : (w < 9 ? 6 : 7));
                ui = uj + w - (uj == 4 ? 0 : (uj == 5 ? 4 : (uj == 6 ? 7 : 9)));
            }
            if (ui >= 0) {
                double acc[2][2][2];
                unit_load(ui, uj, acc);
                unit_run(ui, uj, acc);
            }
            // the previous panel's trailing update (next in this phase) reads what the units above wrote
            if (jb == 1) asm volatile("bar.sync 3, %0;" ::"r"(32 * nw) : "memory");
        });
        CS_STAMP(6);
        if (tid == 0 && bad && a.info) atomicCAS(a.info, 0, rb + bad);
        double* An = A + (long long)rb * (lda + 1);
        for (int e = tid; e < PT * (PT / 2); e += CS_THREADS) {
            const int r = e >> 6, c = (e & 63) * 2;
            const double l0 = c <= r ? S[pk(r, c)] : 0.0;
            const double l1 = c + 1 <= r ? S[pk(r, c + 1)] : 0.0;
            *reinterpret_cast<double2*>(An + (long long)r * lda + c) = make_double2(l0, l1);
            if ((r >> 5) == (c >> 5)) {
                const double t0 = r > c ? S[pk(c, r)] : (r == c ? rinv[r] : 0.0);
                const double t1 = r > c + 1 ? S[pk(c + 1, r)] : (r == c + 1 ? rinv[r] : 0.0);
                *reinterpret_cast<double2*>(a.Tsub_next + (long long)r * PT + c) = make_double2(t0, t1);
            }
        }
        CS_STAMP(7);
        return;
    }

    // ---- wide CTA ------------------------------------------------------------------------------------------------
    __syncthreads();
    CS_STAMP(2);
    for (int e = tid; e < 64 * (PT / 2); e += CS_THREADS) {
        const int r = e >> 6, c = (e & 63) * 2;
        if (r < live)
            *reinterpret_cast<double2*>(A + (long long)(r0 + r) * lda + col + c) =
                make_double2(Xs[r * TS_LD + c], Xs[r * TS_LD + c + 1]);
    }
    if (r0 < a.n) {
        const int mlive = min(live, a.n - r0);  // only rows of the matrix itself have a mirrored tile
        for (int e = tid; e < PT * 64; e += CS_THREADS) {
            const int j = e >> 6, r = e & 63;
            if (r < mlive) A[(long long)(col + j) * lda + r0 + r] = Xs[r * TS_LD + j];
        }
    }
    const int gend = rb + PT * a.nrem;  // first row below the group's own row blocks
    if (r0 < gend) {
        // this CTA's rows belong to a later tile of the group: others read them as the B operand of that tile's block
        __threadfence();
        __syncthreads();
        if (tid == 0) {
            const unsigned int id = a.flag_id + 1u + (unsigned int)wb;
            st_release_u32(a.flags + (id & (CS_FLAGS - 1)), id);
        }
    }
    // updates of the group's remaining column blocks: block t (columns cj = col + 128 t) needs the solved rows
    // [cj, cj + 128) as B, in halves of 64 rows: half 0 -> the Lp region, half 1 -> rows 64.. of Xs
    CS_STAMP(3);
    double* Bh[2] = {Lp, Xs + 64 * TS_LD};
    const int wi = warp & 3, wj = warp >> 2;  // warp tile: rows 16 wi .., columns 32 wj .. of the 64 x 128 block
    for (int t = 1; t <= a.nrem; ++t) {
        const int cj = col + PT * t;
        if (cj > r0 + live - 1) break;  // this block lies right of the CTA's rows entirely
        const int nh = (cj + 64 <= r0 + live - 1) ? 2 : 1;  // halves with columns <= the last row
        __syncthreads();  // previous users of the B buffers (the solve / the previous block) are done
        if (tid == 0) {
            for (int h = 0; h < nh; ++h) {
                const unsigned int id = a.flag_id + (t == 1 ? 0u : 1u + 2u * (unsigned int)(t - 2) + (unsigned int)h);
                const unsigned int* f = a.flags + (id & (CS_FLAGS - 1));
                while (ld_acquire_u32(f) != id) __nanosleep(64);
                if (t == 1) break;  // tile k + 1's rows are one producer
            }
        }
        __syncthreads();
        for (int h = 0; h < nh; ++h) {
            const uint32_t bb = smem_u32(Bh[h]);
            const double* src = A + (long long)(cj + 64 * h) * lda + col;
            for (int e = tid; e < 64 * (PT / 2); e += CS_THREADS) {
                const int r = e >> 6, c = (e & 63) * 2;
                cp_async16(bb + (uint32_t)(r * TS_LD + c) * 8u, src + (long long)r * lda + c, 16);
            }
        }
        cp_async_commit();
        CS_STAMP(1 + 3 * t);
        // C values of the warp's 2 x 4 tiles while the B rows arrive
        const int h = wj >> 1;  // half the warp's columns fall in
        const bool active = h < nh;
        double acc[2][4][2];
        double* Cb = A + (long long)r0 * lda + cj;
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int r = 16 * wi + 8 * i + gq, c = 32 * wj + 8 * j + 2 * kk;
                acc[i][j][0] = acc[i][j][1] = 0.0;
                if (active && r < live && cj + c <= r0 + r) {
                    const double2 v = *reinterpret_cast<const double2*>(Cb + (long long)r * lda + c);
                    acc[i][j][0] = v.x; acc[i][j][1] = v.y;
                }
            }
        cp_async_wait<0>();
        __syncthreads();
        CS_STAMP(2 + 3 * t);
        if (active) {
            const double* Bp = Bh[h] + (32 * (wj & 1) + gq) * TS_LD + kk;
            const double* Ap = Xs + (16 * wi + gq) * TS_LD + kk;
            for (int k0 = 0; k0 < PT; k0 += 32) {
                double fa[2][8], fb[4][8];
#pragma unroll
                for (int s = 0; s < 8; ++s) {
#pragma unroll
                    for (int i = 0; i < 2; ++i) fa[i][s] = flip_sign(Ap[8 * i * TS_LD + k0 + 4 * s]);
#pragma unroll
                    for (int j = 0; j < 4; ++j) fb[j][s] = Bp[8 * j * TS_LD + k0 + 4 * s];
                }
#pragma unroll
                for (int s = 0; s < 8; ++s)
#pragma unroll
                    for (int i = 0; i < 2; ++i)
#pragma unroll
                        for (int j = 0; j < 4; ++j) dmma884(acc[i][j][0], acc[i][j][1], fa[i][s], fb[j][s]);
            }
#pragma unroll
            for (int i = 0; i < 2; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int r = 16 * wi + 8 * i + gq, c = 32 * wj + 8 * j + 2 * kk;
                    // 8x8 tiles that reach above the diagonal keep their upper entries untouched
                    if (r < live && cj + c <= r0 + r) {
                        if (cj + c + 1 <= r0 + r)
                            *reinterpret_cast<double2*>(Cb + (long long)r * lda + c) = make_double2(acc[i][j][0], acc[i][j][1]);
                        else
                            Cb[(long long)r * lda + c] = acc[i][j][0];
                    }
                }
        }
        CS_STAMP(3 + 3 * t);
    }
}

static unsigned int chain_flag_base(unsigned int count) {
    static std::mutex m;
    static unsigned int next = 1;
    std::lock_guard<std::mutex> lock(m);
    const unsigned int b = next;
    next += count;
    if (next < b) next = 1 + count;  // (wrap-around: ids only need to differ from what a slot last held)
    return b;
}

static int dev_env(const char* name);
// development hook (not part of the C-ABI header): the phase clocks of the last stamped chain-step launch
int debug_chain_stamps(long long* out) {
    return cudaMemcpyFromSymbol(out, g_cs_stamps, sizeof(long long) * 64) == cudaSuccess ? GPMP_OK : GPMP_ERR_CUDA;
}

static int launch_chain_step(ChainStepArgs a, cudaStream_t stream) {
    static unsigned long long configured = 0;
    static unsigned int* flags_dev[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (!((configured >> (dev & 63)) & 1ull)) {
        if (cudaFuncSetAttribute(chain_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, CS_SMEM) != cudaSuccess)
            return GPMP_ERR_CUDA;
        void* p = nullptr;
        if (cudaGetSymbolAddress(&p, g_chain_flags) != cudaSuccess) return GPMP_ERR_CUDA;
        flags_dev[dev & 63] = static_cast<unsigned int*>(p);
        configured |= 1ull << (dev & 63);
    }
    a.flags = flags_dev[dev & 63];
    static const bool stamps = dev_env("GPMP_DEV_STAMPS") != 0;
    a.stamps = stamps ? 1 : 0;
    const bool has_narrow = a.nrem >= 1;
    const int rb = a.col + PT;
    const int wide_rows = a.nrows - (has_narrow ? rb + PT : rb);
    const int nwide = wide_rows > 0 ? ceil_div(wide_rows, 64) : 0;
    const int grid = (has_narrow ? 2 : 0) + nwide;  // narrow CTA first, the mirror CTA of its rows last
    if (grid <= 0) return GPMP_OK;
    const double M = (double)(a.nrows - rb);
    LaunchScope scope(KC_GEMM, 2.0 * M * PT * PT * (1.0 + a.nrem), stream);
    chain_step_kernel<<<grid, CS_THREADS, CS_SMEM, stream>>>(a);
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

// development knobs (unset in production): GPMP_DEV_NB overrides the column-group width for n > 1024,
// GPMP_DEV_LA the look-ahead depth (1 .. LA_DEPTH)
static int dev_env(const char* name) {
    const char* v = getenv(name);
    return v ? atoi(v) : 0;
}

int potrf_block_size(int n) {
    if (n <= 1024) return 128;
    static const int forced = dev_env("GPMP_DEV_NB");
    if (forced == 256 || forced == 512) return forced;
    if (n <= 4096) return 256;
    return 512;
}

// Doubling pass over the diagonal blocks of size s inside [0, len): for each pair (first block full,
// second block of size s2 <= s) computes T21 = -T22 * L21 * T11 into Tlo (and its mirror into Tup).
// L, Tlo, Tup, Xt are addressed from the origin of the range; Xt is scratch with the same indexing.
// pair0: the first pair0 full pairs are already done (the leading block inverted under the factorisation's tail);
// sm_first >= 0: persistent launches that keep off the SMs below sm_first (batch == 1 only).
static int doubling_level(const double* L, long long ldl, long long strideL, double* Tlo, double* Tup,
                          long long ldt, long long strideT, double* Xt, long long ldx, long long strideX,
                          int len, int s, int batch, cudaStream_t stream, int pair0 = 0, int sm_first = -1) {
    const int npairs_full = len / (2 * s);
    const int rem = len - npairs_full * 2 * s;  // leftover: a ragged pair if rem > s
    auto launch = [&](const GemmDesc& d) {
        return sm_first >= 0 && batch == 1 ? launch_gemm_nt_persist(d, stream, sm_first) : launch_gemm_nt(d, stream);
    };
    for (int pass = 0; pass < 2; ++pass) {
        int np, s2;
        long long o;  // origin (block index offset) of the first pair of this pass
        if (pass == 0) { np = npairs_full - min(pair0, npairs_full); s2 = s; o = (long long)min(pair0, npairs_full) * 2 * s; }
        else { np = rem > s ? 1 : 0; s2 = rem - s; o = (long long)npairs_full * 2 * s; }
        if (np <= 0) continue;
        // Xt (s x s2) = Tup11 (rows j, k >= j) x L21^T   [NT: A = Tup11, B = L21]
        GemmDesc g = gemm_desc();
        g.A = Tup + o * (ldt + 1); g.lda = ldt; g.strideA = strideT; g.stride2A = 2LL * s * (ldt + 1);
        g.B = L + (o + s) * ldl + o; g.ldb = ldl; g.strideB = strideL; g.stride2B = 2LL * s * (ldl + 1);
        g.C = Xt + o * ldx + (o + s); g.ldc = ldx; g.strideC = strideX; g.stride2C = 2LL * s * (ldx + 1);
        g.M = s; g.N = s2; g.K = s; g.krange = KR_FROM_ROW; g.batch = batch; g.batch2 = np;
        int rc = launch(g);
        if (rc) return rc;
        // T21 (s2 x s) = -Tlo22 (rows i, k <= i) x Xt^T  [NT: A = Tlo22, B = Xt]
        GemmDesc h = gemm_desc();
        h.A = Tlo + (o + s) * (ldt + 1); h.lda = ldt; h.strideA = strideT; h.stride2A = 2LL * s * (ldt + 1);
        h.B = Xt + o * ldx + (o + s); h.ldb = ldx; h.strideB = strideX; h.stride2B = 2LL * s * (ldx + 1);
        h.C = Tlo + (o + s) * ldt + o; h.ldc = ldt; h.strideC = strideT; h.stride2C = 2LL * s * (ldt + 1);
        h.Ct = Tup + o * ldt + (o + s); h.ldct = ldt; h.strideCt = strideT; h.stride2Ct = 2LL * s * (ldt + 1);
        h.M = s2; h.N = s; h.K = s2; h.alpha = -1.0; h.krange = KR_TO_ROW; h.reverse = 1;
        h.batch = batch; h.batch2 = np;
        rc = launch(h);
        if (rc) return rc;
    }
    return GPMP_OK;
}

// ---- pieces of one outer step (all optionally batched over blockIdx.z with element strides) ------
struct PotrfCtx {
    double* A; long long lda, strideA;
    int n, nrows, NB;
    double* Tlo; double* Tup; long long strideT;
    int* info; long long strideInfo;
    int batch;
    double* Tsub;  // optional: 128x128 tile per 128 columns receiving the block-diagonal (32x32) inverses
    long long strideTsub;  // batched: ONE such tile per batch entry, reused by every step (value-only path)
    int nextra;    // > 0: rows n .. n + nextra - 1 of A ride through the panel solves as the solve kernel's extra
                   // strip (batched value-only path); nrows == n then
};

// Column group [k0, k0+gw) (gw <= NB): 128-wide steps, each = tile factor+inverse, solve of ALL rows below
// the tile against the tile inverse (out of place into the group panel buffer, 64-tiles when the product
// is small), copy-back + mirror, and the K=128 update of the columns that remain inside the group.  When
// the loop ends the buffer holds the complete solved panel of the group: W[(r - k0)][j] = L[r][k0 + j] for
// every row r below the tile of column j (rows n..nrows-1 included), so no NB-wide triangular solve and no
// NB-wide block inverse sit on the critical path (the inverses are doubled up after the factorisation).
// One 128-wide step of the group starting at k0: tile factor+inverse, solve of all rows below (into the
// group panel buffer), copy-back + mirror.
static int tile_step(const PotrfCtx& c, int k0, int j0, double* Wg, long long strideW, cudaStream_t stream) {
    const int NB = c.NB, gw = min(NB, c.n - k0);
    const long long lda = c.lda;
    double* Tlo_k = c.Tlo + (long long)(k0 / NB) * NB * NB;
    double* Tup_k = c.Tup + (long long)(k0 / NB) * NB * NB;
    const int jb = min(PT, gw - j0), col = k0 + j0, rb = col + jb;
    Potf2Args pa;
    pa.A = c.A + (long long)col * (lda + 1); pa.lda = lda; pa.strideA = c.strideA;
    pa.Tlo = Tlo_k + (long long)j0 * (NB + 1); pa.Tup = Tup_k + (long long)j0 * (NB + 1);
    pa.ldt = NB; pa.strideT = c.strideT; pa.nb = jb; pa.info = c.info; pa.strideInfo = c.strideInfo;
    pa.row0 = col; pa.dbg = nullptr; pa.mode = POTF2_FULL;
    int rc = launch_potf2(pa, c.batch, stream);
    if (rc) return rc;
    const int M = c.nrows - rb;
    if (M <= 0) return GPMP_OK;
    double* Pa = c.A + (long long)rb * lda + col;         // rows below the tile, in A
    double* Pw = Wg + (long long)(rb - k0) * NB + j0;      // the same rows in the group panel buffer
    GemmDesc g = gemm_desc();
    g.A = Pa; g.lda = lda; g.strideA = c.strideA;
    g.B = pa.Tlo; g.ldb = NB; g.strideB = c.strideT;
    g.C = Pw; g.ldc = NB; g.strideC = strideW;
    g.M = M; g.N = jb; g.K = jb; g.batch = c.batch;
    rc = launch_gemm_nt(g, stream);
    if (rc) return rc;
    CopyPanelArgs cp;
    cp.W = Pw; cp.ldw = NB; cp.strideW = strideW; cp.Alo = Pa; cp.Aup = c.A + (long long)col * lda + rb;
    cp.lda = lda; cp.strideA = c.strideA; cp.rows = M; cp.cols = jb; cp.mirror_rows = max(0, c.n - rb);
    return launch_copy_panel(cp, c.batch, stream);
}

// The chain's version of tile_step: factor-only tile kernel (no 128-wide inverse) and the substitution solve,
// which also writes A in place and the mirrored tiles (no copy kernel).  The 128-wide inverse of the tile is
// computed off the chain by tile_inverse().
static int tile_step_chain(const PotrfCtx& c, int k0, int j0, double* Wg, long long strideW, cudaStream_t stream) {
    const int NB = c.NB, gw = min(NB, c.n - k0);
    const long long lda = c.lda;
    const int jb = min(PT, gw - j0), col = k0 + j0, rb = col + jb;
    double* Tsub = c.batch > 1 ? c.Tsub : c.Tsub + (long long)(col / PT) * PT * PT;
    Potf2Args pa;
    pa.A = c.A + (long long)col * (lda + 1); pa.lda = lda; pa.strideA = c.strideA;
    pa.Tlo = Tsub; pa.Tup = nullptr; pa.ldt = PT; pa.strideT = c.strideTsub; pa.nb = jb;
    pa.info = c.info; pa.strideInfo = c.strideInfo; pa.row0 = col; pa.dbg = nullptr; pa.mode = POTF2_FACTOR;
    int rc = launch_potf2(pa, c.batch, stream);
    if (rc) return rc;
    const int M = c.nrows - rb;
    if (M <= 0 && c.nextra == 0) return GPMP_OK;
    TrsmTileArgs t;
    t.E = c.nextra ? c.A + (long long)c.n * lda + col : nullptr; t.nextra = c.nextra;
    t.P = c.A + (long long)rb * lda + col; t.lda = lda; t.strideA = c.strideA;
    t.Ltile = pa.A; t.Tsub = Tsub; t.ldt = PT; t.strideT = c.strideTsub;
    t.W = Wg + (long long)(rb - k0) * NB + j0; t.ldw = NB; t.strideW = strideW;
    t.Aup = c.batch > 1 ? nullptr : c.A + (long long)col * lda + rb;  // batched values need no mirrored tiles
    t.M = M; t.nb = jb; t.mirror_rows = max(0, c.n - rb);
    t.blocks_per_cta = 1;
    t.inplace_from = c.batch > 1 ? max(0, c.n - rb) : 0;  // batched values: only the extra rows are read back from A
    return launch_trsm_tile(t, c.batch, stream);
}

// Factor-only tile kernel for the first tile of the group starting at k0 (the fused chain steps do the rest).
static int tile_factor_chain(const PotrfCtx& c, int k0, cudaStream_t stream) {
    Potf2Args pa;
    pa.A = c.A + (long long)k0 * (c.lda + 1); pa.lda = c.lda; pa.strideA = c.strideA;
    pa.Tlo = c.Tsub + (long long)(k0 / PT) * PT * PT; pa.Tup = nullptr; pa.ldt = PT; pa.strideT = c.strideTsub;
    pa.nb = min(PT, c.n - k0); pa.info = c.info; pa.strideInfo = c.strideInfo; pa.row0 = k0; pa.dbg = nullptr;
    pa.mode = POTF2_FACTOR;
    return launch_potf2(pa, 1, stream);
}

// 128-wide inverse of the tile at column `col` from its factor (off the chain).
static int tile_inverse(const PotrfCtx& c, int k0, int j0, cudaStream_t stream) {
    const int NB = c.NB, gw = min(NB, c.n - k0);
    const int jb = min(PT, gw - j0), col = k0 + j0;
    Potf2Args pa;
    pa.A = c.A + (long long)col * (c.lda + 1); pa.lda = c.lda; pa.strideA = c.strideA;
    pa.Tlo = c.Tlo + (long long)(k0 / NB) * NB * NB + (long long)j0 * (NB + 1);
    pa.Tup = c.Tup + (long long)(k0 / NB) * NB * NB + (long long)j0 * (NB + 1);
    pa.ldt = NB; pa.strideT = c.strideT; pa.nb = jb; pa.info = nullptr; pa.strideInfo = 0; pa.row0 = col;
    pa.dbg = nullptr; pa.mode = POTF2_INVERT;
    return launch_potf2(pa, c.batch, stream);
}

// K=128 update inside the group starting at k0: the solved step at j0 updates the group's columns
// [c0, c1) (offsets inside the group, multiples of 128, c0 > j0), all rows from c0 down.
static int in_group_update(const PotrfCtx& c, int k0, int j0, int c0, int c1, double* Wg, long long strideW,
                           cudaStream_t stream) {
    const int NB = c.NB, gw = min(NB, c.n - k0);
    c1 = min(c1, gw);
    if (c0 >= c1) return GPMP_OK;
    const int jb = min(PT, gw - j0), R = k0 + c0, M = c.nrows - R;
    if (M <= 0) return GPMP_OK;
    const double* Pw = Wg + (long long)c0 * NB + j0;
    GemmDesc h = gemm_desc();
    h.A = Pw; h.lda = NB; h.strideA = strideW;
    h.B = Pw; h.ldb = NB; h.strideB = strideW;
    h.C = c.A + (long long)R * (c.lda + 1); h.ldc = c.lda; h.strideC = c.strideA;
    h.M = M; h.N = c1 - c0; h.K = jb; h.alpha = -1.0; h.beta = 1.0; h.lower = 1; h.batch = c.batch;
    return launch_gemm_nt(h, stream);
}

// Whole group on one stream (the non-pipelined path).
static int group_panel(const PotrfCtx& c, int k0, double* Wg, long long strideW, cudaStream_t stream,
                       bool force_light = false) {
    const int gw = min(c.NB, c.n - k0);
    // batched value-only evaluations (Tsub given, batch > 1) never need the 128-wide inverses: factor-only
    // tile kernel + substitution solve; the owner of a group in the partitioned factorisation takes the same
    // steps (force_light: every rank rebuilds the tile inverses from the finished factor anyway)
    const bool light = c.Tsub != nullptr && (c.batch > 1 || force_light);
    for (int j0 = 0; j0 < gw; j0 += PT) {
        int rc = light ? tile_step_chain(c, k0, j0, Wg, strideW, stream) : tile_step(c, k0, j0, Wg, strideW, stream);
        if (rc) return rc;
        rc = in_group_update(c, k0, j0, j0 + PT, gw, Wg, strideW, stream);
        if (rc) return rc;
    }
    return GPMP_OK;
}

// NB-wide inverses of the diagonal blocks from the 128-tile inverses, by doubling; all full blocks of one
// matrix go in one batched launch per level.  Xscr: scratch of >= nblk * NB * NB doubles per batch entry.
// g0 / g1 (single matrices only): restrict to the full groups [g0, g1) (g1 < 0: to the end, ragged last group
// included).
static int block_inverses(const PotrfCtx& c, double* Xscr, long long strideX, cudaStream_t stream, int g0 = 0,
                          int g1 = -1) {
    const int NB = c.NB, n = c.n;
    if (NB <= PT) return GPMP_OK;
    if (c.batch > 1 && c.Tsub != nullptr) return GPMP_OK;  // value-only batched path: no tile inverses exist
    const int nfull = n / NB, rem = g1 >= 0 ? 0 : n - nfull * NB;
    const int gl = g1 >= 0 ? min(g1, nfull) : nfull;
    int rc;
    for (int s = PT; s < NB; s *= 2) {
        if (c.batch == 1) {
            if (gl > g0) {
                rc = doubling_level(c.A + (long long)g0 * NB * (c.lda + 1), c.lda, (long long)NB * (c.lda + 1),
                                    c.Tlo + (long long)g0 * NB * NB, c.Tup + (long long)g0 * NB * NB, NB,
                                    (long long)NB * NB, Xscr + (long long)g0 * NB * NB, NB, (long long)NB * NB, NB, s,
                                    gl - g0, stream);
                if (rc) return rc;
            }
        } else {
            for (int b = 0; b < nfull; ++b) {
                rc = doubling_level(c.A + (long long)b * NB * (c.lda + 1), c.lda, c.strideA,
                                    c.Tlo + (long long)b * NB * NB, c.Tup + (long long)b * NB * NB, NB, c.strideT,
                                    Xscr, NB, strideX, NB, s, c.batch, stream);
                if (rc) return rc;
            }
        }
        if (rem > s) {
            rc = doubling_level(c.A + (long long)nfull * NB * (c.lda + 1), c.lda, c.strideA,
                                c.Tlo + (long long)nfull * NB * NB, c.Tup + (long long)nfull * NB * NB, NB,
                                c.strideT, Xscr, NB, strideX, rem, s, c.batch, stream);
            if (rc) return rc;
        }
    }
    return GPMP_OK;
}

// Trailing update with the solved panel of step k, restricted to the block columns [col0, col1) of the
// trailing matrix (offsets relative to r0; col1 < 0: to the end).  Lower tiles; the extra rows are the
// rectangular tail of the tile list.
static int trailing_update(const PotrfCtx& c, int k, const double* Pnl, long long ldp, long long strideP,
                           int col0, int col1, cudaStream_t stream) {
    const int NB = c.NB, nbk = min(NB, c.n - k), r0 = k + nbk;
    const int Nn = c.n - r0;
    if (col1 < 0 || col1 > Nn) col1 = Nn;
    if (col0 >= col1) return GPMP_OK;
    const int M = c.nrows - r0 - col0;
    GemmDesc h = gemm_desc();
    h.A = Pnl + (long long)col0 * ldp; h.lda = ldp; h.strideA = strideP;
    h.B = Pnl + (long long)col0 * ldp; h.ldb = ldp; h.strideB = strideP;
    h.C = c.A + (long long)(r0 + col0) * (c.lda + 1); h.ldc = c.lda; h.strideC = c.strideA;
    h.M = M; h.N = col1 - col0; h.K = nbk; h.alpha = -1.0; h.beta = 1.0; h.lower = 1; h.batch = c.batch;
    return launch_gemm_nt(h, stream);
}

// Library-owned streams and events of the look-ahead pipeline: one set per (device, caller stream), so two
// caller streams that factor concurrently never share chain streams or reuse each other's events (the ABI is
// re-entrant per (stream, workspace)).  The sets live for the life of the process.
constexpr int LA_DEPTH = 4;    // deepest look-ahead the stream / event sets are sized for
constexpr int LA_DEFAULT = 4;  // depth used: measured at n = 8192 with the fused chain step (value, ms): depth 1 8.82,
                               // 2 8.81, 3 8.54, 4 8.47 (with three chain launches per step the chain was the slower
                               // side under a running bulk update at every depth: 8.96 / 9.10 / - / 9.17)
struct LookAhead {
    cudaStream_t caller = nullptr;
    int dev = -1;
    cudaStream_t side = nullptr;    // the chain: what the next tile factorisation is waiting for
    cudaStream_t helper = nullptr;  // updates inside the next column group that are not on the chain
    cudaStream_t ahead[LA_DEPTH] = {};  // look-ahead updates: group g receives panels g-D .. g-2 on ahead[g % D]
    cudaStream_t early = nullptr;   // the leading block of T = L^-1, inverted under the factorisation's tail
    std::vector<cudaEvent_t> ev;
    bool ok = false;
};
static std::mutex g_la_mutex;
static std::vector<LookAhead*> g_la_sets;

static LookAhead* lookahead(cudaStream_t caller, int nevents) {
    int dev = 0;
    cudaGetDevice(&dev);
    LookAhead* la = nullptr;
    {
        std::lock_guard<std::mutex> lock(g_la_mutex);
        for (LookAhead* c : g_la_sets)
            if (c->dev == dev && c->caller == caller) { la = c; break; }
        if (!la) {
            la = new LookAhead();
            la->caller = caller;
            la->dev = dev;
            int lo = 0, hi = 0;
            cudaDeviceGetStreamPriorityRange(&lo, &hi);
            const int mid = hi < lo ? hi + 1 : hi;  // one step below the chain when the device has the range
            la->ok = cudaStreamCreateWithPriority(&la->side, cudaStreamNonBlocking, hi) == cudaSuccess &&
                     cudaStreamCreateWithPriority(&la->helper, cudaStreamNonBlocking, hi) == cudaSuccess;
            for (int i = 0; i < LA_DEPTH && la->ok; ++i)
                la->ok = cudaStreamCreateWithPriority(&la->ahead[i], cudaStreamNonBlocking, mid) == cudaSuccess;
            la->ok = la->ok && cudaStreamCreateWithPriority(&la->early, cudaStreamNonBlocking, lo) == cudaSuccess;
            g_la_sets.push_back(la);
        }
    }
    // a set is only ever used from its caller stream's thread of control, so growing it needs no lock
    while (la->ok && (int)la->ev.size() < nevents) {
        cudaEvent_t e;
        if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) { la->ok = false; break; }
        la->ev.push_back(e);
    }
    return la;
}

struct CopyDiagArgs {
    const double* slo; const double* sup; int NB;  // compact blocks (ld NB)
    double* dlo; double* dup; long long ld;        // full matrices
    int n;
    long long strideS, strideD;                    // batch strides (blockIdx.z) of the compact / full matrices
};
__global__ void copy_diag_blocks_kernel(const CopyDiagArgs a) {
    const int b = blockIdx.y;
    const int nbk = min(a.NB, a.n - b * a.NB);
    const long long src0 = (long long)blockIdx.z * a.strideS + (long long)b * a.NB * a.NB;
    const long long dst0 = (long long)blockIdx.z * a.strideD + (long long)b * a.NB * (a.ld + 1);
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < (long long)nbk * nbk;
         e += (long long)gridDim.x * blockDim.x) {
        const int r = (int)(e / nbk), c = (int)(e - (long long)r * nbk);
        a.dlo[dst0 + (long long)r * a.ld + c] = a.slo[src0 + (long long)r * a.NB + c];
        a.dup[dst0 + (long long)r * a.ld + c] = a.sup[src0 + (long long)r * a.NB + c];
    }
}

// ---- the leading block of T = L^-1 under the tail of the factorisation ---------------------------------------
// At n = 8192 the last third of the factorisation is bound by the chain's latency and leaves most SMs idle, while
// the gradient that follows starts with n^3/3 of GEMM work (T = L^-1 by block doubling) whose levels inside the
// leading P x P block only need the first P columns of L -- final long before the factorisation ends.  When the
// caller hands over the buffers of T (gpmp_lik_value with a gradient-sized workspace) those levels are enqueued on a
// low-priority stream once the bulk updates have run dry, as persistent GEMMs that keep off the first EARLY_SM_FIRST
// SMs (where the chain's whole-SM CTAs then always find room).  early_prefix() is the contract between the two
// calls: a pure function of (n, NB), so that gpmp_lik_grad knows which pairs are done without any state.
constexpr int EARLY_SM_FIRST = 40;
static int early_sm_first() {
    static const int v = dev_env("GPMP_DEV_EARLYSM");
    return v > 0 ? v : EARLY_SM_FIRST;
}
int early_prefix(int n, int NB) {
    static const bool off = dev_env("GPMP_DEV_NOEARLY") != 0;
    if (off || NB < 256 || n % NB != 0) return 0;
    const int nblk = n / NB;
    if (nblk < 8 || n > 16384) return 0;  // (beyond that the bulk updates bound the whole factorisation)
    int p = 1;
    while (2 * p <= nblk / 2) p *= 2;
    return p * NB;
}

static int early_inverse(const PotrfCtx& c, const EarlyInverse& e, int P, cudaStream_t stream, int sm_first) {
    const int NB = c.NB;
    // NB-wide inverses of the first P / NB diagonal blocks from their tile inverses (scratch: the K^-1 buffer)
    int rc = block_inverses(c, e.X, 0, stream, 0, P / NB);
    if (rc) return rc;
    {
        CopyDiagArgs cd;
        cd.slo = c.Tlo; cd.sup = c.Tup; cd.NB = NB; cd.dlo = e.Tlo; cd.dup = e.Tup; cd.ld = e.ld; cd.n = c.n;
        cd.strideS = 0; cd.strideD = 0;
        LaunchScope scope(KC_SMALL, 0.0, stream);
        dim3 grid(min(64, ceil_div(NB * NB, 256)), P / NB, 1);
        copy_diag_blocks_kernel<<<grid, 256, 0, stream>>>(cd);
        GPMP_CHECK_LAUNCH();
    }
    for (long long s = NB; s < P; s *= 2) {
        rc = doubling_level(c.A, c.lda, 0, e.Tlo, e.Tup, e.ld, 0, e.X, e.ld, 0, P, (int)s, 1, stream, 0, sm_first);
        if (rc) return rc;
    }
    return GPMP_OK;
}

// Core factorisation.
//   A      (nrows x lda)            in: lower of K (+ extra rows); out: L both-ways (+ whitened rows)
//   Tlo/Tup compact diagonal-block inverses: nblk blocks of NB x NB (ld NB), block b at b*NB*NB
//   W      panel scratch: TWO buffers of max(nrows, NB) x NB (ld NB) per batch entry, the second at
//          W + Wrows*NB (the in-group updates of one group read its buffer while the next group fills the other)
// Single matrices with several column groups run a depth-D look-ahead (D = LA_DEFAULT, up to LA_DEPTH): the
// caller's stream does the bulk of every K=NB trailing update -- the column groups more than D ahead -- while
// library-owned priority streams bring the next D groups up to date one (panel, group) product at a time and run
// the 128-wide steps of the next group, so the chain may run up to D groups ahead of the bulk.  Trailing updates
// read the solved panel straight from A (its final place; nothing writes it again), so a lagging bulk never holds
// a panel buffer.
int potrf_core(double* A, long long lda, long long strideA, int n, int nrows, int NB, double* Tlo, double* Tup,
               long long strideT, double* W, long long strideW, int* info, long long strideInfo, int batch,
               cudaStream_t stream, double* Tsub, long long strideTsub, int tsub_tiles, const EarlyInverse* early) {
    if (n <= 0) return GPMP_OK;
    // batched value-only runs: the few whitening rows below the matrix travel as the solve kernel's extra strip
    const int nextra = (batch > 1 && Tsub != nullptr && nrows > n && nrows - n <= TS_EXTRA) ? nrows - n : 0;
    if (nextra) nrows = n;
    PotrfCtx c{A, lda, strideA, n, nrows, NB, Tlo, Tup, strideT, info, strideInfo, batch, Tsub, strideTsub, nextra};
    const int nblk = ceil_div(n, NB);
    const long long wrows = nrows > NB ? nrows : NB;
    double* Wb[2] = {W, W + wrows * NB};
    int rc;
    // the pipelined path keeps one block-diagonal inverse tile per 128 columns: the caller must have sized Tsub
    // for that (tsub_tiles); the batched value-only callers pass ONE tile per matrix and never come here
    const bool pipelined = batch == 1 && nblk >= 3 && NB > PT && Tsub != nullptr && tsub_tiles >= ceil_div(n, PT);
    const int EV = 11;
    LookAhead* la = pipelined ? lookahead(stream, 4 + 2 * LA_DEPTH + EV * (nblk + LA_DEPTH + 2)) : nullptr;
    if (!pipelined || !la->ok) {
        if (batch == 1 && Tsub != nullptr && tsub_tiles < ceil_div(n, PT)) c.Tsub = nullptr;  // full tile inverses
        for (int k = 0; k < n; k += NB) {
            const int gw = min(NB, n - k), r0 = k + gw;
            rc = group_panel(c, k, Wb[0], strideW, stream);
            if (rc) return rc;
            if (nrows - r0 <= 0) break;
            rc = trailing_update(c, k, Wb[0] + (long long)gw * NB, NB, strideW, 0, -1, stream);
            if (rc) return rc;
        }
        rc = block_inverses(c, Wb[0], strideW, stream);
        if (rc) return rc;
        // the contract of early_prefix() holds on this path too (in stream order, nothing to overlap with)
        const int P0 = early && batch == 1 ? early_prefix(n, NB) : 0;
        return P0 > 0 ? early_inverse(c, *early, P0, stream, -1) : GPMP_OK;
    }
    // Streams.  s (caller's): the bulk K=NB updates.  B (chain): only what the next tile factorisation waits
    // for -- the update of the next 128 columns, the tile kernel, the solve below it.  H (helper): the other
    // updates inside the next group (its remaining columns from the previous panel and from its own earlier
    // steps) and the tile inverses.  Q[g % D]: the updates of group g by the panels g-D .. g-2, in that order
    // (stream order is the read-modify-write order of the group's columns).  Events order every other
    // read-modify-write of a column block.
    static const int forced_depth = dev_env("GPMP_DEV_LA");
    const int D = min(forced_depth > 0 ? min(forced_depth, LA_DEPTH) : LA_DEFAULT, nblk - 1);
    // profiling mode 2 (gpmp_prof_enable(2)): the same launches, all on the caller's stream, so the per-class
    // CUDA-event times are exclusive kernel times
    const bool serial = prof().enabled == 2;
    cudaStream_t s = stream, B = serial ? stream : la->side, H = serial ? stream : la->helper;
    cudaStream_t Q[LA_DEPTH];
    for (int i = 0; i < LA_DEPTH; ++i) Q[i] = serial ? stream : la->ahead[i];
    auto ev = [&](int g, int i) { return la->ev[4 + 2 * LA_DEPTH + g * EV + i]; };  // g = 0 .. nblk + D
    enum { E_PANEL = 0, E_REST = 1, E_HEADREST = 2, E_TRSM = 3 /* +c, c<4 */, E_IN = 7 /* +c, c<2 */, E_G2 = 9 };
    // steps of the group starting at k0 (group index g); head_rest: wait for the helper's head update first
    static const bool no_fuse = dev_env("GPMP_DEV_NOFUSE") != 0;
    auto run_group = [&](int g, int k0, double* Wn, bool head_rest) -> int {
        const int gw = min(NB, n - k0), nblocks = ceil_div(gw, PT);
        if (!no_fuse && gw % PT == 0) {
            // groups of full tiles: the first tile on its own (its columns were completed by the head update on
            // this stream), then ONE fused launch per step -- solve below tile cb, every in-group update of the step,
            // factorisation of tile cb + 1 (chain_step_kernel).  The in-group updates write the group's remaining
            // columns, so the first step follows the helper's part of the head update.
            int rc2 = tile_factor_chain(c, k0, B);
            if (rc2) return rc2;
            const unsigned int fid = chain_flag_base(8u * (unsigned int)nblocks);
            for (int cb = 0; cb < nblocks; ++cb) {
                const int col = k0 + cb * PT;
                if (cb == 0 && head_rest) cudaStreamWaitEvent(B, ev(g, E_HEADREST), 0);
                ChainStepArgs cs;
                cs.A = c.A; cs.lda = c.lda; cs.n = c.n; cs.nrows = c.nrows; cs.col = col;
                cs.nrem = nblocks - 1 - cb;
                cs.Tsub_k = c.Tsub + (long long)(col / PT) * PT * PT;
                cs.Tsub_next = c.Tsub + (long long)(col / PT + 1) * PT * PT;
                cs.info = c.info; cs.flags = nullptr; cs.flag_id = fid + 8u * (unsigned int)cb;
                rc2 = launch_chain_step(cs, B);
                if (rc2) return rc2;
                cudaEventRecord(ev(g, E_TRSM + cb), B);
                // the 128-wide inverse of the tile is needed only after the factorisation: helper stream
                cudaStreamWaitEvent(H, ev(g, E_TRSM + cb), 0);
                rc2 = tile_inverse(c, k0, cb * PT, H);
                if (rc2) return rc2;
            }
            cudaEventRecord(ev(g, E_PANEL), B);
            return GPMP_OK;
        }
        for (int cb = 0; cb < nblocks; ++cb) {
            const int j0 = cb * PT;
            int rc2 = tile_step_chain(c, k0, j0, Wn, strideW, B);
            if (rc2) return rc2;
            cudaEventRecord(ev(g, E_TRSM + cb), B);
            if (cb + 1 < nblocks) {
                if (cb == 0) {
                    if (head_rest) cudaStreamWaitEvent(B, ev(g, E_HEADREST), 0);
                } else {
                    cudaStreamWaitEvent(B, ev(g, E_IN + cb - 1), 0);
                }
                rc2 = in_group_update(c, k0, j0, j0 + PT, j0 + 2 * PT, Wn, strideW, B);
                if (rc2) return rc2;
            }
            if (cb + 2 < nblocks) {
                cudaStreamWaitEvent(H, ev(g, E_TRSM + cb), 0);
                rc2 = in_group_update(c, k0, j0, j0 + 2 * PT, gw, Wn, strideW, H);
                if (rc2) return rc2;
                cudaEventRecord(ev(g, E_IN + cb), H);
            } else {
                cudaStreamWaitEvent(H, ev(g, E_TRSM + cb), 0);
            }
            // the 128-wide inverse of this tile is needed only after the factorisation: helper stream,
            // behind the update the chain may be waiting for
            rc2 = tile_inverse(c, k0, j0, H);
            if (rc2) return rc2;
        }
        cudaEventRecord(ev(g, E_PANEL), B);
        return GPMP_OK;
    };
    const int P_early = early ? early_prefix(n, NB) : 0;
    static const int forced_b = dev_env("GPMP_DEV_EARLYB");
    // (as soon as the leading block's columns are final: measured at n = 8192, evals/s with the trigger after group
    // 7 / 8 / 9 / 10 / 11 / 12: 50.4 / 50.3 / 50.2 / 49.9 / 49.6 / 48.9; without the early block 48.6)
    const int b_early = forced_b > 0 ? forced_b : P_early / NB - 1;
    bool early_started = false;
    // fork
    if (cudaEventRecord(la->ev[0], s) != cudaSuccess || cudaStreamWaitEvent(B, la->ev[0], 0) != cudaSuccess ||
        cudaStreamWaitEvent(H, la->ev[0], 0) != cudaSuccess)
        return GPMP_ERR_CUDA;
    for (int i = 0; i < D; ++i)
        if (cudaStreamWaitEvent(Q[i], la->ev[0], 0) != cudaSuccess) return GPMP_ERR_CUDA;
    rc = run_group(0, 0, Wb[0], false);
    if (rc) return rc;
    for (int b = 0; b < nblk; ++b) {
        const int k = b * NB, gw = min(NB, n - k), r0 = k + gw;
        if (nrows - r0 <= 0) break;
        const double* Pk = A + (long long)r0 * lda + k;  // solved panel rows r0.. of group b, in place
        const int Nn = n - r0;                // trailing columns
        const int nb_next = min(NB, Nn);      // width of the next group (0 when only extra rows remain)
        const int head0 = min(PT, nb_next);
        // group b+1 by panel b: first 128 columns on the chain, the rest on the helper; both behind the update
        // of that group by panel b-1 (a look-ahead product when D >= 2, part of the bulk otherwise)
        if (b > 0) {
            cudaEvent_t dep = D >= 2 ? ev(b + 1, E_G2) : ev(b - 1, E_REST);
            cudaStreamWaitEvent(B, dep, 0);
            cudaStreamWaitEvent(H, dep, 0);
        }
        if (nb_next > 0) {
            rc = trailing_update(c, k, Pk, lda, 0, 0, head0, B);
            if (rc) return rc;
            if (nb_next > head0) {
                cudaStreamWaitEvent(H, ev(b, E_PANEL), 0);
                rc = trailing_update(c, k, Pk, lda, 0, head0, nb_next, H);
                if (rc) return rc;
                cudaEventRecord(ev(b + 1, E_HEADREST), H);
            }
        }
        // groups b+2 .. b+D by panel b, one product each on the group's own stream; the farthest one follows
        // the bulk update by panel b-1, which still covered that group
        for (int d = 2; d <= D; ++d) {
            const int c0 = (d - 1) * NB;
            if (c0 >= Nn) break;
            const int c1 = min(d * NB, Nn), g = b + d;
            cudaStream_t q = Q[g % D];
            cudaStreamWaitEvent(q, ev(b, E_PANEL), 0);
            if (d == D && b > 0) cudaStreamWaitEvent(q, ev(b - 1, E_REST), 0);
            rc = trailing_update(c, k, Pk, lda, 0, c0, c1, q);
            if (rc) return rc;
            if (d == 2) cudaEventRecord(ev(g, E_G2), q);
        }
        // bulk of the trailing update (the groups beyond the look-ahead window) on the caller's stream
        cudaStreamWaitEvent(s, ev(b, E_PANEL), 0);
        rc = trailing_update(c, k, Pk, lda, 0, min(Nn, D * NB), -1, s);
        if (rc) return rc;
        cudaEventRecord(ev(b, E_REST), s);
        if (P_early > 0 && !early_started && b >= b_early) {
            // the bulk updates have (nearly) run dry and the first P_early columns of L are final: the levels of
            // T = L^-1 inside the leading block, behind the tile inverses the helper stream has been given so far
            cudaStream_t E = serial ? stream : la->early;
            cudaEventRecord(la->ev[4 + LA_DEPTH], H);
            cudaStreamWaitEvent(E, la->ev[4 + LA_DEPTH], 0);
            cudaStreamWaitEvent(E, ev(b, E_REST), 0);
            rc = early_inverse(c, *early, P_early, E, early_sm_first());
            if (rc) return rc;
            early_started = true;
        }
        // next group's steps
        if (nb_next > 0) {
            rc = run_group(b + 1, r0, Wb[(b + 1) & 1], nb_next > head0);
            if (rc) return rc;
        }
    }
    // join
    if (cudaEventRecord(la->ev[1], B) != cudaSuccess || cudaStreamWaitEvent(s, la->ev[1], 0) != cudaSuccess)
        return GPMP_ERR_CUDA;
    if (cudaEventRecord(la->ev[2], H) != cudaSuccess || cudaStreamWaitEvent(s, la->ev[2], 0) != cudaSuccess)
        return GPMP_ERR_CUDA;
    for (int i = 0; i < D; ++i)
        if (cudaEventRecord(la->ev[4 + i], Q[i]) != cudaSuccess ||
            cudaStreamWaitEvent(s, la->ev[4 + i], 0) != cudaSuccess)
            return GPMP_ERR_CUDA;
    if (P_early > 0) {
        if (!early_started) {
            rc = early_inverse(c, *early, P_early, s, -1);
            if (rc) return rc;
        } else if (!serial) {
            if (cudaEventRecord(la->ev[3], la->early) != cudaSuccess || cudaStreamWaitEvent(s, la->ev[3], 0) != cudaSuccess)
                return GPMP_ERR_CUDA;
        }
    }
    return block_inverses(c, Wb[0], strideW, s, P_early / NB, -1);
}

// ---- panel-partitioned factorisation across GPUs ------------------------------------------------------
// One process per GPU holds the whole work matrix but only keeps its own column groups (group g belongs to
// rank g mod G) up to date.  The owner factors a group into the panel buffer, the caller broadcasts the
// buffer (NCCL), every rank applies the panel to the column groups it owns, non-owners also file the panel
// into their copy of L.  The pieces below are what the C-ABI exposes for that loop.
struct Copy2DArgs { const double* src; long long lds; double* dst; long long ldd; int rows, cols; };
__global__ void copy2d_kernel(const Copy2DArgs a) {
    const int r = blockIdx.y;
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < a.cols; c += gridDim.x * blockDim.x)
        a.dst[(long long)r * a.ldd + c] = a.src[(long long)r * a.lds + c];
}
static int launch_copy2d(const double* src, long long lds, double* dst, long long ldd, int rows, int cols,
                         cudaStream_t stream) {
    if (rows <= 0 || cols <= 0) return GPMP_OK;
    Copy2DArgs a{src, lds, dst, ldd, rows, cols};
    LaunchScope scope(KC_SMALL, 0.0, stream);
    dim3 grid(ceil_div(cols, 128), rows);
    copy2d_kernel<<<grid, 128, 0, stream>>>(a);
    GPMP_CHECK_LAUNCH();
    return GPMP_OK;
}

// Owner: factor the group at k0 (its columns carry every earlier update) and leave in `panel` (ld NB) the
// group's block column of L: row (r - k0) = L[r][k0 .. k0+gw) for r = k0 .. nrows-1 (diagonal tiles included).
int dist_group(double* A, long long lda, int n, int nrows, int NB, double* Tlo, double* Tup, int k0, double* panel,
               int* info, cudaStream_t stream, double* Tsub) {
    // Tsub: one block-diagonal inverse tile per 128 columns (the potrf workspace has them): the group is factored
    // with the chain's steps -- factor-only tile kernel (34 us instead of 62) and the substitution solve, which
    // writes the panel buffer, A and the mirrored tiles in one pass (no GEMM against the tile inverse, no copy)
    PotrfCtx c{A, lda, 0, n, nrows, NB, Tlo, Tup, 0, info, 0, 1, Tsub, 0, 0};
    int rc = group_panel(c, k0, panel, 0, stream, Tsub != nullptr);
    if (rc) return rc;
    const int gw = min(NB, n - k0);
    for (int j0 = 0; j0 < gw; j0 += PT) {
        const int jb = min(PT, gw - j0), col = k0 + j0;
        rc = launch_copy2d(A + (long long)col * (lda + 1), lda, panel + (long long)j0 * NB + j0, NB, jb, jb, stream);
        if (rc) return rc;
    }
    return GPMP_OK;
}

// Non-owner: file a received panel into A (diagonal tiles, rows below each tile, mirrored upper tiles).
int dist_store(double* A, long long lda, int n, int nrows, int NB, int k0, const double* panel, cudaStream_t stream) {
    const int gw = min(NB, n - k0);
    for (int j0 = 0; j0 < gw; j0 += PT) {
        const int jb = min(PT, gw - j0), col = k0 + j0, rb = col + jb;
        int rc = launch_copy2d(panel + (long long)j0 * NB + j0, NB, A + (long long)col * (lda + 1), lda, jb, jb, stream);
        if (rc) return rc;
        const int M = nrows - rb;
        if (M <= 0) continue;
        CopyPanelArgs cp;
        cp.W = panel + (long long)(rb - k0) * NB + j0; cp.ldw = NB; cp.strideW = 0;
        cp.Alo = A + (long long)rb * lda + col; cp.Aup = A + (long long)col * lda + rb;
        cp.lda = lda; cp.strideA = 0; cp.rows = M; cp.cols = jb; cp.mirror_rows = max(0, n - rb);
        rc = launch_copy_panel(cp, 1, stream);
        if (rc) return rc;
    }
    return GPMP_OK;
}

// Every rank: trailing update of the absolute columns [col0, col1) (beyond the group at k0) with its panel.
int dist_update(double* A, long long lda, int n, int nrows, int NB, int k0, const double* panel, int col0, int col1,
                cudaStream_t stream) {
    PotrfCtx c{A, lda, 0, n, nrows, NB, nullptr, nullptr, 0, nullptr, 0, 1, nullptr, 0};
    const int gw = min(NB, n - k0), r0 = k0 + gw;
    if (col0 < r0 || col1 <= col0) return GPMP_ERR_ARG;
    return trailing_update(c, k0, panel + (long long)gw * NB, NB, 0, col0 - r0, col1 - r0, stream);
}

// Every rank, after the last panel: tile inverses from the (received) factor, then the NB-wide inverses.
int dist_finish(double* A, long long lda, int n, int nrows, int NB, double* Tlo, double* Tup, double* scratch,
                int* info, cudaStream_t stream) {
    for (int k0 = 0; k0 < n; k0 += NB) {
        const int gw = min(NB, n - k0);
        double* Tlo_k = Tlo + (long long)(k0 / NB) * NB * NB;
        double* Tup_k = Tup + (long long)(k0 / NB) * NB * NB;
        const int full = gw / PT, rem = gw - full * PT;
        Potf2Args pa;
        pa.A = A + (long long)k0 * (lda + 1); pa.lda = lda; pa.strideA = (long long)PT * (lda + 1);
        pa.Tlo = Tlo_k; pa.Tup = Tup_k; pa.ldt = NB; pa.strideT = (long long)PT * (NB + 1);
        pa.nb = PT; pa.info = nullptr; pa.strideInfo = 0; pa.row0 = k0; pa.dbg = nullptr; pa.mode = POTF2_INVERT;
        int rc;
        if (full > 0) {
            rc = launch_potf2(pa, full, stream);
            if (rc) return rc;
        }
        if (rem > 0) {
            pa.A += (long long)full * pa.strideA; pa.Tlo += (long long)full * pa.strideT;
            pa.Tup += (long long)full * pa.strideT; pa.nb = rem;
            rc = launch_potf2(pa, 1, stream);
            if (rc) return rc;
        }
    }
    PotrfCtx c{A, lda, 0, n, nrows, NB, Tlo, Tup, 0, info, 0, 1, nullptr, 0};
    return block_inverses(c, scratch, 0, stream);
}

// ---- potri: Tlo/Tup (n x n) from L and the compact NB-block inverses, then Kinv = T^T T (lower) ----
int potri_core(const double* L, int n, long long ldl, int NB, const double* Tlo_c, const double* Tup_c,
               double* Tlo, double* Tup, double* Kinv, long long ldk, cudaStream_t stream, int batch,
               long long strideL, long long strideTc, long long strideT, int prefix, cudaEvent_t t_done) {
    // batched form: entry b reads L + b strideL and the compact inverses + b strideTc, and writes Tlo / Tup / Kinv
    // + b strideT (all three share one stride)
    if (n <= 0 || batch <= 0) return GPMP_OK;
    int rc;
    const int nblk = ceil_div(n, NB);
    {
        CopyDiagArgs c;
        c.slo = Tlo_c; c.sup = Tup_c; c.NB = NB; c.dlo = Tlo; c.dup = Tup; c.ld = ldk; c.n = n;
        c.strideS = strideTc; c.strideD = strideT;
        LaunchScope scope(KC_SMALL, 0.0, stream);
        dim3 grid(min(64, ceil_div(NB * NB, 256)), nblk, batch);
        copy_diag_blocks_kernel<<<grid, 256, 0, stream>>>(c);
        GPMP_CHECK_LAUNCH();
    }
    // doubling NB -> n; scratch for X^T: the Kinv buffer (written only by the final product)
    // (prefix: the leading prefix x prefix block of T was completed under the tail of the factorisation)
    for (long long s = NB; s < n; s *= 2) {
        const int pair0 = (batch == 1 && s < prefix) ? (int)(prefix / (2 * s)) : 0;
        rc = doubling_level(L, ldl, strideL, Tlo, Tup, ldk, strideT, Kinv, ldk, strideT, n, (int)s, batch, stream,
                            pair0);
        if (rc) return rc;
    }
    // T is complete: work that needs only T (the U rows of the gradient) may start on another stream now
    if (t_done && cudaEventRecord(t_done, stream) != cudaSuccess) return GPMP_ERR_CUDA;
    // Kinv (lower) = T^T T :  Kinv_ij = sum_{k >= i} Tup[i][k] Tup[j][k]
    GemmDesc g = gemm_desc();
    g.A = Tup; g.lda = ldk; g.strideA = strideT; g.B = Tup; g.ldb = ldk; g.strideB = strideT;
    g.C = Kinv; g.ldc = ldk; g.strideC = strideT;
    g.M = n; g.N = n; g.K = n; g.lower = 1; g.krange = KR_FROM_ROW; g.batch = batch;
    return launch_gemm_nt(g, stream);
}

// ---- row-wise triangular solves with many right-hand sides stored as rows ----------------------
// trans == 0:  Bt <- Bt L^-T   (forward; uses Tlo diagonal-block inverses and L's lower tiles)
// trans == 1:  Bt <- Bt L^-1   (backward; uses Tup diagonal-block inverses and the mirrored upper tiles)
// Block size NB as in the factorisation (complete compact inverses required).  W: m x NB scratch.
// first_block (trans == 0 only): the right-hand sides are zero left of block first_block, so the forward
// substitution starts there (unit-vector right-hand sides of the distributed triangular inverse).
int trsm_rows_core(const double* A, int n, long long lda, int NB, const double* Tlo_c, const double* Tup_c,
                   double* Bt, int m, long long ldb, int trans, double* W, cudaStream_t stream, int first_block) {
    if (n <= 0 || m <= 0) return GPMP_OK;
    int rc;
    const int nblk = ceil_div(n, NB);
    for (int bi = (trans ? 0 : first_block); bi < nblk; ++bi) {
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
