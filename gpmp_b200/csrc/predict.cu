// K5: kriging prediction in whitened coordinates (reference: core/kriging.py:35-116,
// 170-199 solves the (n+q) saddle system and returns lambda explicitly).
//
// With K = L L^T, P~ = L^-1 P = Q~ R~ and the residual r = z~ - Q~ Q~^T z~ left in the work matrix by the
// likelihood pipeline, a chunk of test points is one row-wise pass over V = K(xt, xi) L^-T (rows v_t):
//   c_t = Q~^T v_t,  g_t = R~^-T p_t,  e_t = c_t - g_t
//   mean_t = v_t . r + g_t . (Q~^T z~)        var_t = k_tt - |v_t|^2 + |e_t|^2
//   lambda_t = L^-T (v_t - Q~ e_t)            (only when the caller wants the weights)
// Zero / parameterised mean is the q = 0 case (r = z~).
#include "internal.cuh"

namespace gpmp {

constexpr int RD_WARPS = 8;

__global__ void __launch_bounds__(RD_WARPS * 32) rowdots_kernel(const RowDotsArgs a) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int t = blockIdx.x * RD_WARPS + warp;
    if (t >= a.m) return;
    const double* __restrict__ v = a.V + (long long)t * a.ldv;
    const int q = a.q, r = q + 1;
    double* dots = a.dots + (long long)t * (r + 1);  // [c_0..c_{q-1}, v.r, |v|^2]
    double vv = 0.0;
    for (int a0 = 0; a0 < r; a0 += 4) {
        const int cnt = min(4, r - a0);
        double s[4] = {0.0, 0.0, 0.0, 0.0};
        double sv = 0.0;
        for (int k = lane; k < a.n; k += 32) {
            const double x = v[k];
            sv = fma(x, x, sv);
#pragma unroll
            for (int c = 0; c < 4; ++c)
                if (c < cnt) s[c] = fma(x, a.R[(long long)(a0 + c) * a.ldr + k], s[c]);
        }
#pragma unroll
        for (int c = 0; c < 4; ++c)
            if (c < cnt) {
                const double w = warp_sum(s[c]);
                if (lane == 0) dots[a0 + c] = w;
            }
        if (a0 == 0) vv = warp_sum(sv);
    }
    __syncwarp();
    if (lane == 0) {
        dots[r] = vv;
        // g = R~^-T p_t by forward substitution; e = c - g (stored over c)
        const double* Rt = a.Rt;
        double mean = dots[q], e2 = 0.0;
        double g[GPMP_MAX_Q];
        for (int j = 0; j < q; ++j) {
            double sacc = a.Pt ? a.Pt[(long long)t * q + j] : 0.0;
            for (int k = 0; k < j; ++k) sacc -= Rt[k * r + j] * g[k];
            g[j] = sacc / Rt[j * r + j];
            mean = fma(g[j], Rt[j * r + q], mean);
            const double e = dots[j] - g[j];
            dots[j] = e;
            e2 = fma(e, e, e2);
        }
        const double ktt = a.ktt ? a.ktt[t] : a.ktt_scalar;
        a.mean[t] = mean;
        a.var[t] = ktt - vv + e2;
    }
}
int launch_rowdots(const RowDotsArgs& a, cudaStream_t stream) {
    if (a.m <= 0) return GPMP_OK;
    LaunchScope scope(KC_SMALL, 8.0 * (double)a.m * a.n, stream);
    rowdots_kernel<<<ceil_div(a.m, RD_WARPS), RD_WARPS * 32, 0, stream>>>(a);
    GPMP_CHECK_LAUNCH();
    return GPMP_OK;
}

// V[t][:] -= sum_a e[t][a] * Q~[a][:]
__global__ void __launch_bounds__(256) wrows_kernel(const RowDotsArgs a) {
    const int t = blockIdx.y;
    const int q = a.q;
    const double* e = a.dots + (long long)t * (q + 2);
    double* v = a.V + (long long)t * a.ldv;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += gridDim.x * blockDim.x) {
        double x = v[i];
        for (int c = 0; c < q; ++c) x = fma(-e[c], a.R[(long long)c * a.ldr + i], x);
        v[i] = x;
    }
}
int launch_wrows(const RowDotsArgs& a, cudaStream_t stream) {
    if (a.m <= 0 || a.q <= 0) return GPMP_OK;
    LaunchScope scope(KC_SMALL, 16.0 * (double)a.m * a.n, stream);
    dim3 grid(min(ceil_div(a.n, 256), 32), a.m);
    wrows_kernel<<<grid, 256, 0, stream>>>(a);
    GPMP_CHECK_LAUNCH();
    return GPMP_OK;
}

// out[c][r] = in[r][c]
struct TransposeArgs { const double* in; long long ldi; double* out; long long ldo; int rows, cols; };
__global__ void __launch_bounds__(256) transpose_kernel(const TransposeArgs a) {
    __shared__ double t[32][33];
    const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int i = ty; i < 32; i += 8) {
        const int r = r0 + i, c = c0 + tx;
        t[i][tx] = (r < a.rows && c < a.cols) ? a.in[(long long)r * a.ldi + c] : 0.0;
    }
    __syncthreads();
    for (int i = ty; i < 32; i += 8) {
        const int c = c0 + i, r = r0 + tx;
        if (r < a.rows && c < a.cols) a.out[(long long)c * a.ldo + r] = t[tx][i];
    }
}
int launch_transpose(const double* in, long long ldi, double* out, long long ldo, int rows, int cols,
                     cudaStream_t stream) {
    if (rows <= 0 || cols <= 0) return GPMP_OK;
    TransposeArgs a{in, ldi, out, ldo, rows, cols};
    LaunchScope scope(KC_SMALL, 16.0 * (double)rows * cols, stream);
    dim3 grid(ceil_div(cols, 32), ceil_div(rows, 32));
    transpose_kernel<<<grid, 256, 0, stream>>>(a);
    GPMP_CHECK_LAUNCH();
    return GPMP_OK;
}

}  // namespace gpmp
