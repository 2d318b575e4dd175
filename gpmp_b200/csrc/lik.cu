// K3: Gaussian likelihoods by Cholesky whitening, and the assembly of the gradient operator.
//
// REML without the contrast matrix W (reference: core/likelihood.py:92-129 forms
// W = Q[:, q:] of a complete QR and G = W^T K W with two n^3 GEMMs): with K = L L^T the mean basis P and
// the data z are whitened as extra ROWS of the factorisation (P~ = L^-1 P, z~ = L^-1 z), then one CTA
// orthogonalises the q+1 rows [P~^T ; z~^T] by twice-iterated classical Gram-Schmidt:
//   log det(W^T K W) = 2 sum log L_ii + 2 sum log R~_ii - 2 sum log R0_ii    (P = Q0 R0, P~ = Q~ R~)
//   (W^T z)^T (W^T K W)^-1 W^T z = || z~ - Q~ Q~^T z~ ||^2                   (the residual row itself)
// Gradient (A.4): M = K^-1 - U^T U with U = [Q~^T ; r^T] T  (T = L^-1, r the residual row): rows 0..q-1
// are B^T = (L^-T Q~)^T and row q is alpha^T = (Pi z)^T, so dvalue/dz = alpha and
// dvalue/dtheta_j = 0.5 sum_ik M_ik dK_ik/dtheta_j, contracted tile by tile in matern.cu.
#include <math.h>
#include "internal.cuh"

namespace gpmp {

// ---- row loader: rows 0..q-1 <- P^T, row q <- z (batch-shared inputs, per-batch destination) --------
__global__ void load_rows_kernel(const LoadRowsArgs a) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n) return;
    double* rows = a.rows + (long long)blockIdx.z * a.stride;
    for (int j = 0; j < a.q; ++j) {
        const double v = a.P[(long long)i * a.q + j];
        rows[(long long)j * a.ld + i] = v;
        if (a.p0rows && blockIdx.z == 0) a.p0rows[(long long)j * a.ld0 + i] = v;
    }
    rows[(long long)a.q * a.ld + i] = a.z[(long long)blockIdx.z * a.strideZ + i];
}
int launch_load_rows(const LoadRowsArgs& a, int batch, cudaStream_t stream) {
    LaunchScope scope(KC_SMALL, 0.0, stream);
    dim3 grid(ceil_div(a.n, 256), 1, batch);
    load_rows_kernel<<<grid, 256, 0, stream>>>(a);
    GPMP_CHECK_LAUNCH();
    return GPMP_OK;
}

// ---- copy the lower triangle of a user-supplied covariance into the work matrix -------------------
struct CopyLowerArgs { const double* K; long long ldk; double* A; long long lda; int n; };
__global__ void copy_lower_kernel(const CopyLowerArgs a) {
    const int r = blockIdx.y;
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c <= r; c += gridDim.x * blockDim.x)
        a.A[(long long)r * a.lda + c] = a.K[(long long)r * a.ldk + c];
}
int launch_copy_lower(const double* K, long long ldk, double* A, long long lda, int n, cudaStream_t stream) {
    if (n <= 0) return GPMP_OK;
    CopyLowerArgs a{K, ldk, A, lda, n};
    LaunchScope scope(KC_SMALL, 0.0, stream);
    dim3 grid(min(ceil_div(n, 256), 64), n);
    copy_lower_kernel<<<grid, 256, 0, stream>>>(a);
    GPMP_CHECK_LAUNCH();
    return GPMP_OK;
}

// ---- finalize: CGS2 of the whitened rows + log-determinants -> criterion value ---------------------
constexpr int FIN_THREADS = 1024;


__device__ __forceinline__ void block_reduce_vec(double* v, int cnt, double* red /* [32][8] + [8] */) {
    // reduces cnt (<= 8) per-thread values across the block; results broadcast into v[]
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int c = 0; c < cnt; ++c) v[c] = warp_sum(v[c]);
    __syncthreads();
    if (lane == 0)
        for (int c = 0; c < cnt; ++c) red[w * 8 + c] = v[c];
    __syncthreads();
    if (w == 0) {
        for (int c = 0; c < cnt; ++c) {
            double t = lane < nw ? red[lane * 8 + c] : 0.0;
            t = warp_sum(t);
            if (lane == 0) red[32 * 8 + c] = t;
        }
    }
    __syncthreads();
    for (int c = 0; c < cnt; ++c) v[c] = red[32 * 8 + c];
}

// Orthogonalise row j of V against rows 0..j-1 (already orthonormal), twice.  Returns ||v_j||^2 after.
// hcol (optional, stride hs): accumulates the projection coefficients <v_j, v_k>, k < j (column j of R).
__device__ double cgs2_row(double* V, long long ld, int n, int j, double* red, double* hcol, int hs) {
    double* vj = V + (long long)j * ld;
    for (int pass = 0; pass < 2; ++pass) {
        for (int kb = 0; kb < j; kb += 8) {
            const int cnt = min(8, j - kb);
            double h[8];
#pragma unroll
            for (int c = 0; c < 8; ++c) h[c] = 0.0;
            for (int i = threadIdx.x; i < n; i += blockDim.x) {
                const double x = vj[i];
#pragma unroll
                for (int c = 0; c < 8; ++c)
                    if (c < cnt) h[c] = fma(x, V[(long long)(kb + c) * ld + i], h[c]);
            }
            block_reduce_vec(h, cnt, red);
            if (hcol && threadIdx.x == 0)
                for (int c = 0; c < cnt; ++c) hcol[(kb + c) * hs] = (pass ? hcol[(kb + c) * hs] : 0.0) + h[c];
            for (int i = threadIdx.x; i < n; i += blockDim.x) {
                double x = vj[i];
#pragma unroll
                for (int c = 0; c < 8; ++c)
                    if (c < cnt) x = fma(-h[c], V[(long long)(kb + c) * ld + i], x);
                vj[i] = x;
            }
            __syncthreads();
        }
    }
    double s[1] = {0.0};
    for (int i = threadIdx.x; i < n; i += blockDim.x) s[0] = fma(vj[i], vj[i], s[0]);
    block_reduce_vec(s, 1, red);
    return s[0];
}

__global__ void __launch_bounds__(FIN_THREADS, 1) finalize_kernel(const FinalizeArgs a) {
    __shared__ double red[33 * 8];
    const long long b = blockIdx.x;
    double* V = a.rows + b * a.strideRows;
    const int n = a.n, q = a.q;
    double ld_rt = 0.0, ld_r0 = 0.0;
    // whitened rows: P~ rows normalised in place (-> Q~^T), z~ row left as the residual r
    double quad = 0.0;
    double* Rt = a.Rt ? a.Rt + b * a.strideRt : nullptr;
    for (int j = 0; j <= q; ++j) {
        const double nrm2 = cgs2_row(V, a.ld, n, j, red, Rt ? Rt + j : nullptr, q + 1);
        if (Rt && threadIdx.x == 0) Rt[j * (q + 1) + j] = sqrt(nrm2);
        if (j < q) {
            const double nrm = sqrt(nrm2);
            ld_rt += log(nrm);
            const double inv = 1.0 / nrm;
            double* vj = V + (long long)j * a.ld;
            for (int i = threadIdx.x; i < n; i += blockDim.x) vj[i] *= inv;
            __syncthreads();
        } else {
            quad = nrm2;
        }
    }
    // raw basis: log det(P^T P) = 2 sum log R0_ii
    if (a.ldr0_in) {
        ld_r0 = a.ldr0_in[0];
    } else if (q > 0) {
        double* V0 = a.p0work + b * a.strideP0;
        for (int j = 0; j < q; ++j)
            for (int i = threadIdx.x; i < n; i += blockDim.x)
                V0[(long long)j * a.ld0 + i] = a.p0rows[(long long)j * a.ld0 + i];
        __syncthreads();
        for (int j = 0; j < q; ++j) {
            const double nrm2 = cgs2_row(V0, a.ld0, n, j, red, nullptr, 0);
            const double nrm = sqrt(nrm2);
            ld_r0 += log(nrm);
            const double inv = 1.0 / nrm;
            double* vj = V0 + (long long)j * a.ld0;
            for (int i = threadIdx.x; i < n; i += blockDim.x) vj[i] *= inv;
            __syncthreads();
        }
    }
    const double* L = a.Ldiag + b * a.strideL;
    double s[1] = {0.0};
    for (int i = threadIdx.x; i < n; i += blockDim.x) s[0] += log(L[(long long)i * (a.ldl + 1)]);
    block_reduce_vec(s, 1, red);
    if (threadIdx.x == 0) {
        const double ldl2 = 2.0 * s[0];
        const double logdet = ldl2 + 2.0 * ld_rt - 2.0 * ld_r0;
        double val = 0.5 * ((double)(n - q) * 1.8378770664093453 + logdet + quad);  // log(2 pi)
        const int bad = a.info ? a.info[b * a.strideInfo] : 0;
        if (bad || !(val == val)) val = INFINITY;
        double* o = a.out + b * a.strideOut;
        o[0] = val;
        if (a.strideOut >= 7 || a.strideOut == 0) {  // the 7-slot record (value first)
            o[1] = logdet; o[2] = quad; o[3] = ldl2; o[4] = 2.0 * ld_rt; o[5] = 2.0 * ld_r0;
            o[6] = (double)bad;  // info mirrored as a double: one 64-byte readback serves the host wrapper
        }
    }
}
// sum log R0_ii of the raw basis P (q rows p0rows, orthogonalised on the scratch copy p0work)
struct LogdetR0Args { const double* p0rows; double* p0work; long long ld; int n, q; double* out; };
__global__ void __launch_bounds__(FIN_THREADS, 1) logdet_r0_kernel(const LogdetR0Args a) {
    __shared__ double red[33 * 8];
    for (int j = 0; j < a.q; ++j)
        for (int i = threadIdx.x; i < a.n; i += blockDim.x)
            a.p0work[(long long)j * a.ld + i] = a.p0rows[(long long)j * a.ld + i];
    __syncthreads();
    double acc = 0.0;
    for (int j = 0; j < a.q; ++j) {
        const double nrm = sqrt(cgs2_row(a.p0work, a.ld, a.n, j, red, nullptr, 0));
        acc += log(nrm);
        const double inv = 1.0 / nrm;
        double* vj = a.p0work + (long long)j * a.ld;
        for (int i = threadIdx.x; i < a.n; i += blockDim.x) vj[i] *= inv;
        __syncthreads();
    }
    if (threadIdx.x == 0) a.out[0] = acc;
}
int launch_logdet_r0(const double* p0rows, double* p0work, long long ld, int n, int q, double* out,
                     cudaStream_t stream) {
    LogdetR0Args a{p0rows, p0work, ld, n, q, out};
    LaunchScope scope(KC_SMALL, 0.0, stream);
    logdet_r0_kernel<<<1, FIN_THREADS, 0, stream>>>(a);
    GPMP_CHECK_LAUNCH();
    return GPMP_OK;
}

int launch_finalize(const FinalizeArgs& a, int batch, cudaStream_t stream) {
    LaunchScope scope(KC_SMALL, 0.0, stream);
    finalize_kernel<<<batch, FIN_THREADS, 0, stream>>>(a);
    GPMP_CHECK_LAUNCH();
    return GPMP_OK;
}

// ---- U = R * T :  U[a][j] = sum_{k >= j} R[a][k] * Tup[j][k]   (R = [Q~^T ; r^T], r rows) ---------------
constexpr int UR_WARPS = 8;
__global__ void __launch_bounds__(UR_WARPS * 32) urows_kernel(const URowsArgs a) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int j = a.j0 + blockIdx.x * UR_WARPS + warp;
    if (j >= a.j1) return;
    const long long zb = blockIdx.y;
    const double* __restrict__ R = a.R + zb * a.strideR;
    double* __restrict__ U = a.U + zb * a.strideU;
    const double* __restrict__ t = a.Tup + zb * a.strideT + (long long)j * a.ldt;
    for (int a0 = 0; a0 < a.r; a0 += 4) {
        const int cnt = min(4, a.r - a0);
        double s[4] = {0.0, 0.0, 0.0, 0.0};
        for (int k = j + lane; k < a.n; k += 32) {
            const double tv = t[k];
#pragma unroll
            for (int c = 0; c < 4; ++c)
                if (c < cnt) s[c] = fma(tv, R[(long long)(a0 + c) * a.ldr + k], s[c]);
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            if (c < cnt) {
                const double v = warp_sum(s[c]);
                if (lane == 0) U[(long long)(a0 + c) * a.ldu + j] = v;
            }
        }
    }
}
int launch_urows(const URowsArgs& a0, cudaStream_t stream) {
    if (a0.n <= 0 || a0.r <= 0) return GPMP_OK;
    URowsArgs a = a0;
    if (a.j1 <= 0 || a.j1 > a.n) a.j1 = a.n;
    if (a.j0 < 0) a.j0 = 0;
    if (a.j0 >= a.j1) return GPMP_OK;
    LaunchScope scope(KC_SMALL, 8.0 * a.n * (a.n + 1.0) / 2.0 * a.batch, stream);
    dim3 grid(ceil_div(a.j1 - a.j0, UR_WARPS), a.batch > 0 ? a.batch : 1);
    urows_kernel<<<grid, UR_WARPS * 32, 0, stream>>>(a);
    GPMP_CHECK_LAUNCH();
    return GPMP_OK;
}

// ---- leave-one-out by virtual cross-validation (core/loo.py:65-130): diag(Pi) and Pi z from K^-1 and U ----
struct LooArgs {
    const double* Kinv; long long ldk; const double* U; long long ldu; int q, n;
    const double* z; double* zloo; double* s2loo; double* eloo;
};
__global__ void loo_kernel(const LooArgs a) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n) return;
    double dg = a.Kinv[(long long)i * (a.ldk + 1)];
    for (int c = 0; c < a.q; ++c) {
        const double u = a.U[(long long)c * a.ldu + i];
        dg = fma(-u, u, dg);
    }
    const double e = a.U[(long long)a.q * a.ldu + i] / dg;  // (Pi z)_i / Pi_ii
    a.eloo[i] = e;
    a.s2loo[i] = 1.0 / dg;
    a.zloo[i] = a.z[i] - e;
}
int launch_loo(const double* Kinv, long long ldk, const double* U, long long ldu, int q, int n, const double* z,
               double* zloo, double* s2loo, double* eloo, cudaStream_t stream) {
    LooArgs a{Kinv, ldk, U, ldu, q, n, z, zloo, s2loo, eloo};
    LaunchScope scope(KC_SMALL, 0.0, stream);
    loo_kernel<<<ceil_div(n, 256), 256, 0, stream>>>(a);
    GPMP_CHECK_LAUNCH();
    return GPMP_OK;
}

// rows [0, m) of B (ld) <- unit vectors e_{col0 + i}  (right-hand sides of the distributed triangular inverse)
struct UnitRowsArgs { double* B; long long ld; int m, n, col0; };
__global__ void unit_rows_kernel(const UnitRowsArgs a) {
    const int i = blockIdx.y;
    double* row = a.B + (long long)i * a.ld;
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < a.n; c += gridDim.x * blockDim.x)
        row[c] = (c == a.col0 + i) ? 1.0 : 0.0;
}
int launch_unit_rows(double* B, long long ld, int m, int n, int col0, cudaStream_t stream) {
    if (m <= 0 || n <= 0) return GPMP_OK;
    UnitRowsArgs a{B, ld, m, n, col0};
    LaunchScope scope(KC_SMALL, 0.0, stream);
    dim3 grid(min(ceil_div(n, 256), 64), m);
    unit_rows_kernel<<<grid, 256, 0, stream>>>(a);
    GPMP_CHECK_LAUNCH();
    return GPMP_OK;
}

// ---- dense dvalue/dK = 0.5 (K^-1 - U^T U) for the composable path (user-built covariance) ------------
__global__ void __launch_bounds__(256) dense_grad_kernel(const DenseGradArgs a) {
    __shared__ double ui[GPMP_MAX_Q + 1][32], uj[GPMP_MAX_Q + 1][32];
    const int i0 = blockIdx.y * 32, j0 = blockIdx.x * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int e = threadIdx.x; e < a.r * 32; e += 256) {
        const int c = e >> 5, l = e & 31;
        ui[c][l] = i0 + l < a.n ? a.U[(long long)c * a.ldu + i0 + l] : 0.0;
        uj[c][l] = j0 + l < a.n ? a.U[(long long)c * a.ldu + j0 + l] : 0.0;
    }
    __syncthreads();
    for (int ii = ty; ii < 32; ii += 8) {
        const int i = i0 + ii, j = j0 + tx;
        if (i >= a.n || j >= a.n) continue;
        double v = i >= j ? a.Kinv[(long long)i * a.ldk + j] : a.Kinv[(long long)j * a.ldk + i];
        for (int c = 0; c < a.r; ++c) v = fma(-ui[c][ii], uj[c][tx], v);
        a.dK[(long long)i * a.lddk + j] = a.half * v;
    }
}
int launch_dense_grad(const DenseGradArgs& a, cudaStream_t stream) {
    if (a.n <= 0) return GPMP_OK;
    LaunchScope scope(KC_SMALL, 0.0, stream);
    dim3 grid(ceil_div(a.n, 32), ceil_div(a.n, 32));
    dense_grad_kernel<<<grid, 256, 0, stream>>>(a);
    GPMP_CHECK_LAUNCH();
    return GPMP_OK;
}

}  // namespace gpmp
