// K1: fused anisotropic scaled-distance + half-integer Matern tile kernel, and K4: the dK-regenerating
// gradient contraction.  Both use the same 64x64 tile / 4x4 register block: the x and y tiles are
// pre-scaled by exp(loginvrho) while being staged in shared memory ([dim][point] layout: broadcast /
// conflict-free reads), distances are direct differences (the SciPy-cdist form the NumPy reference uses,
// numpy_backend.py:432-436), stores are 16-byte vectors, 256 B contiguous per half-warp.
#include <math.h>
#include "internal.cuh"

namespace gpmp {

static void matern_coef(int p, double* out) {
    // exp of gammaln differences, as kernel/matern.py:59-63 (table: num/shared.py:21-41)
    for (int i = 0; i < p; ++i)
        out[i] = exp(lgamma(p + 1.0) - lgamma(2.0 * p + 1.0) + lgamma(p + i + 1.0) - lgamma(i + 1.0) -
                     lgamma(p - i + 1.0));
}

int make_matern_dev(const gpmp_cov_spec* s, MaternDev* m, bool same_set) {
    if (!s) return GPMP_ERR_ARG;
    if (s->d < 1 || s->d > GPMP_MAX_DIM || s->p < 0 || s->p > GPMP_MAX_P) return GPMP_ERR_DIM;
    m->p = s->p;
    m->d = s->d;
    m->sigma2 = exp(s->log_sigma2);
    m->diag_add = 0.0;
    if (same_set) m->diag_add = s->noise ? exp(s->log_tau2) : 10.0 * m->sigma2 * 2.220446049250313e-16;
    m->c = 2.0 * sqrt(s->p + 0.5);
    m->dscale = s->p >= 1 ? -(m->c * m->c) / (2.0 * s->p - 1.0) : -m->c;
    for (int i = 0; i < GPMP_MAX_P; ++i) m->coef[i] = m->coefm1[i] = 0.0;
    matern_coef(s->p, m->coef);
    if (s->p >= 1) matern_coef(s->p - 1, m->coefm1);
    // powers of t: (2t)^(p-i) = 2^(p-i) t^(p-i)
    for (int k = 0; k <= GPMP_MAX_P; ++k) m->bq[k] = m->bqm1[k] = 0.0;
    m->bq[0] = m->bqm1[0] = 1.0;
    for (int i = 0; i < s->p; ++i) m->bq[s->p - i] = m->coef[i] * ldexp(1.0, s->p - i);
    for (int i = 0; i + 1 < s->p; ++i) m->bqm1[s->p - 1 - i] = m->coefm1[i] * ldexp(1.0, s->p - 1 - i);
    for (int j = 0; j < GPMP_MAX_DIM; ++j) m->invrho[j] = j < s->d ? exp(s->loginvrho[j]) : 0.0;
    return GPMP_OK;
}

// Batched parameters: Theta[N][1 + noise + d] (device) -> MaternDev[N].  tmpl carries p, d, c, the
// polynomial coefficients; only sigma2 / diagonal term / inverse length-scales differ per row.
struct PrepThetaArgs {
    MaternDev tmpl; const double* theta; int N, noise, same_set; MaternDev* out;
};
__global__ void prep_theta_kernel(const PrepThetaArgs a) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.N) return;
    const int d = a.tmpl.d, w = 1 + a.noise + d;
    const double* th = a.theta + (long long)i * w;
    MaternDev m = a.tmpl;
    m.sigma2 = exp(th[0]);
    m.diag_add = 0.0;
    if (a.same_set) m.diag_add = a.noise ? exp(th[1]) : 10.0 * m.sigma2 * 2.220446049250313e-16;
    for (int j = 0; j < d; ++j) m.invrho[j] = exp(th[1 + a.noise + j]);
    a.out[i] = m;
}
int launch_prep_theta(const gpmp_cov_spec* spec, const double* theta_dev, int N, int same_set, MaternDev* out,
                      cudaStream_t stream) {
    PrepThetaArgs a;
    int rc = make_matern_dev(spec, &a.tmpl, false);
    if (rc) return rc;
    a.theta = theta_dev; a.N = N; a.noise = spec->noise; a.same_set = same_set; a.out = out;
    LaunchScope scope(KC_SMALL, 0.0, stream);
    prep_theta_kernel<<<ceil_div(N, 128), 128, 0, stream>>>(a);
    GPMP_CHECK_LAUNCH();
    return GPMP_OK;
}

#define GPMP_BIGF (1.7976931348623157e308 / 1000.0)  /* inftobigf, torch_backend.py:500-502 */

// q(t) = 1 + sum_{i<p} a_i u^(p-i), u = 2t, by Horner in u
__device__ __forceinline__ double matern_poly(const double* __restrict__ a, int p, double u) {
    double acc = 0.0;
    for (int i = 0; i < p; ++i) acc = (acc + a[i]) * u;
    return 1.0 + acc;
}
__device__ __forceinline__ double matern_k(const MaternDev& m, double h) {
    if (isinf(h)) h = GPMP_BIGF;
    const double t = m.c * h;
    return exp(-t) * matern_poly(m.coef, m.p, 2.0 * t);
}
// k'(h)/h for p >= 1 (finite at 0); for p == 0 returns k'(h) = -c exp(-t) (caller divides by h)
__device__ __forceinline__ double matern_dk_over_h(const MaternDev& m, double h) {
    if (isinf(h)) h = GPMP_BIGF;
    const double t = m.c * h;
    const double e = exp(-t);
    if (m.p == 0) return m.dscale * e;
    return m.dscale * e * matern_poly(m.coefm1, m.p - 1, 2.0 * t);
}

constexpr int CT = 64;          // tile edge
constexpr int CTP = CT + 2;     // row stride of the staged point tiles [dim][point]: 2 doubles of padding spread the
                                // dimension rows over the banks while the tile is filled (the fill walks the row-major
                                // (point, dim) array, so consecutive threads write different dimension rows)
constexpr int COV_THREADS = 256;
constexpr int COV_TILES_PER_CTA = 4;

// 2^(-j/32), j = 0..31, correctly rounded: the table of exp_neg_tab (staged into shared memory by the tile kernels)
__device__ const double EXP2M_TAB[EXP_TAB] = {
    1.0, 0.9785720620877001, 0.9576032806985737, 0.93708381705515,
    0.9170040432046712, 0.8973545375015536, 0.8781260801866497, 0.859309649061239,
    0.8408964152537145, 0.8228777390769825, 0.8052451659746271, 0.7879904225539432,
    0.7711054127039704, 0.7545822137967114, 0.7384130729697497, 0.7225904034885233,
    0.7071067811865476, 0.691954940981916, 0.6771277734684463, 0.6626183215798707,
    0.6484197773255048, 0.6345254785958666, 0.620928906036742, 0.6076236799902345,
    0.5946035575013605, 0.5818624293887887, 0.5693943173783458, 0.5571933712979462,
    0.5452538663326288, 0.5335702003384118, 0.5221368912137069, 0.5109485743270583};

constexpr double T_CLAMP = 707.0;  // exp(-707) = 8e-308: beyond it the kernel value is 0 (the reference underflows at 745)

// Kernel constants held in registers.  P >= 0: regularity known at compile time, Horner coefficients (in powers
// of t = c h, sigma2 folded in) in registers and fully unrolled (p = 0..4 cover every example of the reference but
// one); P == -1: any p up to GPMP_MAX_P, coefficients read from the shared copy of MaternDev.
// The tile kernels stage their points pre-scaled by c / rho_j, so the distance they accumulate IS t = c h.
template <int P>
struct MaternRegs {
    static constexpr int NA = P > 0 ? P + 1 : 1, NB_ = P > 1 ? P : 1;
    double sigma2, sdscale;    // sdscale = sigma2 * dscale
    double bq[NA], bqm1[NB_];  // sigma2 q_p(2t) and sigma2 dscale q_{p-1}(2t) as polynomials in t
    const MaternDev* sm;
    const double* tab;         // shared-memory copy of EXP2M_TAB
    int p;
    __device__ __forceinline__ void init(const MaternDev* m, const double* exp_tab) {
        sm = m; p = m->p; sigma2 = m->sigma2; sdscale = m->sigma2 * m->dscale; tab = exp_tab;
        if (P > 0) {
#pragma unroll
            for (int k = 0; k < NA; ++k) bq[k] = m->sigma2 * m->bq[k];
        } else {
            bq[0] = m->sigma2;
        }
        if (P > 1) {
#pragma unroll
            for (int k = 0; k < NB_; ++k) bqm1[k] = sdscale * m->bqm1[k];
        } else {
            bqm1[0] = sdscale;
        }
    }
    // sigma2 q_p(2t)
    __device__ __forceinline__ double poly(double t) const {
        if (P >= 0) {
            double acc = bq[NA - 1];
#pragma unroll
            for (int k = NA - 2; k >= 0; --k) acc = fma(acc, t, bq[k]);
            return acc;
        }
        double acc = sm->bq[p];
        for (int k = p - 1; k >= 0; --k) acc = fma(acc, t, sm->bq[k]);
        return sigma2 * acc;
    }
    // sigma2 dscale q_{p-1}(2t)   (p >= 1)
    __device__ __forceinline__ double polym1(double t) const {
        if (P >= 0) {
            double acc = bqm1[NB_ - 1];
#pragma unroll
            for (int k = NB_ - 2; k >= 0; --k) acc = fma(acc, t, bqm1[k]);
            return acc;
        }
        double acc = sm->bqm1[p - 1];
        for (int k = p - 2; k >= 0; --k) acc = fma(acc, t, sm->bqm1[k]);
        return sdscale * acc;
    }
    // sigma2 k_p at t = c h (t >= 0, +inf or NaN).  t is clamped before the polynomial, so an infinite distance
    // gives exactly 0 (the reference replaces inf by fmax/1000 first, torch_backend.py:500-502) and a NaN
    // distance gives NaN through the polynomial (through a select for p = 0, which has none).
    __device__ __forceinline__ double cov_t(double t) const {
        const bool big = t > T_CLAMP;
        const double tc = big ? T_CLAMP : t;
        const double e = exp_neg_tab(tc, tab);
        double k = e * poly(tc);
        const bool p0 = P >= 0 ? (P == 0) : (p == 0);
        if (p0) k = tc != tc ? tc : k;
        return big ? 0.0 : k;
    }
    // sigma2 k_p(h) and sigma2 k_p'(h)/h (p >= 1) or sigma2 k_0'(h) (p == 0), at t = c h
    __device__ __forceinline__ void cov_and_dk_t(double t, double& kc, double& dkh) const {
        const bool big = t > T_CLAMP;
        const double tc = big ? T_CLAMP : t;
        const double e = exp_neg_tab(tc, tab);
        const bool p0 = P >= 0 ? (P == 0) : (p == 0);
        double k = e * poly(tc);
        double dk = p0 ? sdscale * e : e * polym1(tc);
        if (p0) { k = tc != tc ? tc : k; dk = tc != tc ? tc : dk; }
        kc = big ? 0.0 : k;
        dkh = big ? 0.0 : dk;
    }
};

__device__ __forceinline__ void stage_exp_tab(double* dst) {
    if (threadIdx.x < EXP_TAB) dst[threadIdx.x] = EXP2M_TAB[threadIdx.x];
}

// Stage the parameter record in shared memory (from the per-batch device array or from the kernel
// parameters) so that every later access is a shared-memory broadcast, not a generic load.
__device__ __forceinline__ void stage_matern(MaternDev* dst, const MaternDev* src_dev, const MaternDev& src_param) {
    double* d = reinterpret_cast<double*>(dst);
    if (src_dev) {
        const double* s = reinterpret_cast<const double*>(src_dev);
        for (int e = threadIdx.x; e < (int)(sizeof(MaternDev) / 8); e += blockDim.x) d[e] = s[e];
    } else if (threadIdx.x == 0) {
        *dst = src_param;
    }
    __syncthreads();
}

enum CovMode { CM_RECT = COV_RECT, CM_SYM_FULL = COV_SYM_FULL, CM_SYM_LOWER = COV_SYM_LOWER };

struct CovArgs {
    MaternDev m;
    const MaternDev* mdev;  // optional per-batch parameters (device array indexed by blockIdx.z)
    long long strideK;      // batch stride of K (elements)
    long long strideX;      // batch stride of the point set (elements; 0 = every entry uses the same points)
    const double* x; const double* y;
    double* K; long long ldk;
    int n, mcols;
    int mode;       // CovMode
    int dist_only;  // write the distance instead of the covariance
    int vec_ok;     // 16-byte stores allowed
    int tiles_n;
    long long ntiles;  // tiles of one matrix
};

__device__ __forceinline__ void tri_decode(int t, int& ti, int& tj) {
    int r = (int)((sqrt(8.0 * (double)t + 1.0) - 1.0) * 0.5);
    while ((long long)(r + 1) * (r + 2) / 2 <= t) ++r;
    while ((long long)r * (r + 1) / 2 > t) --r;
    ti = r;
    tj = t - (int)((long long)r * (r + 1) / 2);
}

__device__ __forceinline__ void stage_points(double* __restrict__ s, const double* __restrict__ pts, int base, int npts,
                                             const MaternDev& m, int tid, double scale) {
    // s[j][r] = scale * invrho[j] * pts[base + r][j]; coalesced over the row-major (point, dim) array.
    // scale = c for covariance values (the accumulated distance is then t = c h), 1 for plain distances.
    const int d = m.d;
    for (int e = tid; e < CT * d; e += COV_THREADS) {
        int r = e / d, j = e - r * d;
        int gr = base + r;
        s[j * CTP + r] = gr < npts ? (scale * m.invrho[j]) * pts[(long long)gr * d + j] : 0.0;
    }
}

// Shared memory (dynamic): x tile [d][CTP], y tile [d][CTP].  The parameter record is read straight from the kernel
// parameters (constant bank) unless the launch is batched over per-entry records in device memory.
template <int P>
__global__ void __launch_bounds__(COV_THREADS, 4) matern_cov_kernel(const __grid_constant__ CovArgs a) {
    extern __shared__ __align__(16) double cov_sm[];
    __shared__ MaternDev msh;
    __shared__ double etab[EXP_TAB];
    stage_exp_tab(etab);
    if (a.mdev) stage_matern(&msh, a.mdev + blockIdx.z, a.m);
    const MaternDev& m = a.mdev ? msh : a.m;
    MaternRegs<P> mr;
    mr.init(&m, etab);
    double* xs = cov_sm;
    double* ys = cov_sm + (size_t)m.d * CTP;
    double* __restrict__ Kb = a.K + (long long)blockIdx.z * a.strideK;
    // A CTA walks COV_TILES_PER_CTA consecutive tiles of the (row-major) tile list: the launch prologue and the
    // x tile (same tile row) are paid once per run, and the tile index is decoded once, then stepped.
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const long long t_first = (long long)blockIdx.x * COV_TILES_PER_CTA;
    int ti, tj;
    if (a.mode == CM_RECT) {
        ti = (int)(t_first / a.tiles_n);
        tj = (int)(t_first - (long long)ti * a.tiles_n);
    } else {
        tri_decode((int)t_first, ti, tj);
    }
    const double pscale = a.dist_only ? 1.0 : m.c;
    int staged_ti = -1;
    for (int it = 0; it < COV_TILES_PER_CTA && t_first + it < a.ntiles; ++it) {
    const int r0 = ti * CT, c0 = tj * CT;
    if (it > 0) __syncthreads();  // every warp is done with the previous y (and x) tile
    if (ti != staged_ti) {
        stage_points(xs, a.x + (long long)blockIdx.z * a.strideX, r0, a.n, m, tid, pscale);
        staged_ti = ti;
    }
    stage_points(ys, a.y + (long long)blockIdx.z * a.strideX, c0, a.mcols, m, tid, pscale);
    __syncthreads();

    double h2[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) h2[i][j] = 0.0;
    for (int j = 0; j < m.d; ++j) {
        const double2 xa = *reinterpret_cast<const double2*>(&xs[j * CTP + ty * 4]);
        const double2 xb = *reinterpret_cast<const double2*>(&xs[j * CTP + ty * 4 + 2]);
        const double2 ya = *reinterpret_cast<const double2*>(&ys[j * CTP + 2 * tx]);
        const double2 yb = *reinterpret_cast<const double2*>(&ys[j * CTP + 32 + 2 * tx]);
        const double xr[4] = {xa.x, xa.y, xb.x, xb.y};
        const double yc[4] = {ya.x, ya.y, yb.x, yb.y};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const double df = xr[i] - yc[k];
                h2[i][k] = fma(df, df, h2[i][k]);
            }
    }
    double v[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const double t = sqrt_seeded(h2[i][k]);  // c h (covariance) or h (distance)
            v[i][k] = a.dist_only ? t : mr.cov_t(t);
        }
    const bool sym = a.mode != CM_RECT;
    const double diag_add = m.diag_add;
    const bool diag_tile = sym && ti == tj;
    // interior tiles (all but the edge and diagonal ones): no bounds or triangle tests, 16-byte stores
    const bool interior = !diag_tile && a.vec_ok && r0 + CT <= a.n && c0 + CT <= a.mcols;
    if (interior) {
        double* kp = Kb + (long long)(r0 + ty * 4) * a.ldk + c0 + 2 * tx;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            *reinterpret_cast<double2*>(kp) = make_double2(v[i][0], v[i][1]);
            *reinterpret_cast<double2*>(kp + 32) = make_double2(v[i][2], v[i][3]);
            kp += a.ldk;
        }
    } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int row = r0 + ty * 4 + i;
            if (row >= a.n) continue;
#pragma unroll
            for (int b = 0; b < 2; ++b) {
                const int col = c0 + 32 * b + 2 * tx;
                double v0 = v[i][2 * b], v1 = v[i][2 * b + 1];
                if (diag_tile && !a.dist_only) {
                    if (col == row) v0 += diag_add;
                    if (col + 1 == row) v1 += diag_add;
                }
                double* kp = Kb + (long long)row * a.ldk + col;
                const bool lower_only = a.mode == CM_SYM_LOWER && diag_tile;
                const bool w0 = col < a.mcols && (!lower_only || col <= row);
                const bool w1 = col + 1 < a.mcols && (!lower_only || col + 1 <= row);
                if (w0 && w1 && a.vec_ok) {
                    *reinterpret_cast<double2*>(kp) = make_double2(v0, v1);
                } else {
                    if (w0) kp[0] = v0;
                    if (w1) kp[1] = v1;
                }
            }
        }
    }
    if (a.mode == CM_SYM_FULL && !diag_tile) {
        // mirrored tile: K[col][row], each thread owns 4 consecutive rows -> 32 contiguous bytes per col
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int col = c0 + 32 * (k >> 1) + 2 * tx + (k & 1);
            if (col >= a.mcols) continue;
            const int row = r0 + ty * 4;
            double* kp = Kb + (long long)col * a.ldk + row;
            if (row + 3 < a.n && a.vec_ok) {
                *reinterpret_cast<double2*>(kp) = make_double2(v[0][k], v[1][k]);
                *reinterpret_cast<double2*>(kp + 2) = make_double2(v[2][k], v[3][k]);
            } else {
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    if (row + i < a.n) kp[i] = v[i][k];
            }
        }
    }
    // next tile of the list
    if (a.mode == CM_RECT) {
        if (++tj == a.tiles_n) { tj = 0; ++ti; }
    } else if (++tj > ti) {
        tj = 0;
        ++ti;
    }
    }  // tile loop
}

// mdev != nullptr: batched over `batch` parameter sets living in device memory (spec then only supplies
// p and d for the host-side checks); K advances by strideK per batch entry, x / y are shared.
int launch_matern_cov(const gpmp_cov_spec* spec, const MaternDev* mdev, int batch, long long strideK,
                      const double* x, int n, const double* y, int mcols, double* K, long long ldk, int mode,
                      int dist_only, cudaStream_t stream, long long strideX) {
    if (n <= 0 || mcols <= 0 || batch <= 0) return GPMP_OK;
    const bool same = (y == nullptr || y == x) ;
    if (mode != CM_RECT && !same) return GPMP_ERR_ARG;
    CovArgs a;
    int rc = make_matern_dev(spec, &a.m, same && !dist_only);
    if (rc) return rc;
    a.mdev = mdev;
    a.strideK = strideK;
    a.strideX = strideX;
    a.x = x;
    a.y = same ? x : y;
    a.K = K;
    a.ldk = ldk;
    a.n = n;
    a.mcols = same ? n : mcols;
    a.mode = mode;
    a.dist_only = dist_only;
    a.vec_ok = ((ldk & 1) == 0) && ((reinterpret_cast<uintptr_t>(K) & 15) == 0);
    const int tm = ceil_div(n, CT), tn = ceil_div(a.mcols, CT);
    a.tiles_n = tn;
    long long ntiles = mode == CM_RECT ? (long long)tm * tn : (long long)tm * (tm + 1) / 2;
    double bytes = mode == CM_SYM_LOWER ? 8.0 * n * (n + 1.0) / 2.0 : 8.0 * (double)n * a.mcols;
    bytes += 8.0 * (double)(n + (same ? 0 : a.mcols)) * spec->d;
    LaunchScope scope(KC_MATERN, bytes * batch, stream);
    a.ntiles = ntiles;
    dim3 grid((unsigned)ceil_div_ll(ntiles, COV_TILES_PER_CTA), 1, (unsigned)batch);
    const size_t cov_smem = (size_t)2 * spec->d * CTP * sizeof(double);  // x and y tiles: <= 33 KB at d = 32
    switch (dist_only ? 0 : spec->p) {
        case 0: matern_cov_kernel<0><<<grid, COV_THREADS, cov_smem, stream>>>(a); break;
        case 1: matern_cov_kernel<1><<<grid, COV_THREADS, cov_smem, stream>>>(a); break;
        case 2: matern_cov_kernel<2><<<grid, COV_THREADS, cov_smem, stream>>>(a); break;
        case 3: matern_cov_kernel<3><<<grid, COV_THREADS, cov_smem, stream>>>(a); break;
        case 4: matern_cov_kernel<4><<<grid, COV_THREADS, cov_smem, stream>>>(a); break;
        default: matern_cov_kernel<-1><<<grid, COV_THREADS, cov_smem, stream>>>(a); break;
    }
    GPMP_CHECK_LAUNCH();
    return GPMP_OK;
}

// ---------------------------------------------------------------------------------------------
// elementwise helpers (pairwise covariance, elementwise distance, maternp_kernel on an array)
// ---------------------------------------------------------------------------------------------
struct PairArgs {
    MaternDev m;
    const double* x; const double* y; double* out; int n; int dist_only;
};
__global__ void pairwise_kernel(const PairArgs a) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n) return;
    if (a.y == nullptr) {
        a.out[i] = a.dist_only ? 0.0 : a.m.sigma2;
        return;
    }
    double h2 = 0.0;
    for (int j = 0; j < a.m.d; ++j) {
        // reference: sqrt(sum((invrho * (x - y))**2)) (torch_backend.py:827-828)
        const double df = a.m.invrho[j] * (a.x[(long long)i * a.m.d + j] - a.y[(long long)i * a.m.d + j]);
        h2 = fma(df, df, h2);
    }
    const double h = sqrt(h2);
    a.out[i] = a.dist_only ? h : a.m.sigma2 * matern_k(a.m, h);
}

struct KernArgs {
    MaternDev m;
    const double* h; double* k; double* dk; long long count;
};
__global__ void maternp_elementwise_kernel(const KernArgs a) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long stride = (long long)gridDim.x * blockDim.x;
    for (; i < a.count; i += stride) {
        const double h = a.h[i];
        a.k[i] = matern_k(a.m, h);
        if (a.dk) {
            const double w = matern_dk_over_h(a.m, h);
            a.dk[i] = a.m.p == 0 ? w : w * h;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// K4: gradient contraction.  out[0] = sum G_ik Kc_ik, out[1] = tr(G) (same set), out[2+j] =
// sum G_ik dKc_ik/dloginvrho_j, with Kc = sigma2 k(D) regenerated per tile.  In the symmetric mode
// G = Kinv - U^T U is assembled on the fly from the lower triangle of Kinv and the r rows of U, and
// only lower tiles are visited (off-diagonal entries weighted twice).
// ---------------------------------------------------------------------------------------------
constexpr int CONTRACT_MAXR = GPMP_MAX_Q + 1;
constexpr int CONTRACT_TILES_PER_CTA = 8;

struct ContractArgs {
    MaternDev m;
    const double* x; const double* y;
    const double* G; long long ldg;
    const double* Ut; long long ldu; int r;  // optional low-rank correction rows (r x n)
    int n, mcols, sym, same_set, tiles_n;
    int dist_only;    // vjp of the scaled distance itself: weight G_ik / h (0 at h == 0)
    int tile_off;     // sym mode: index of the first lower tile visited (row-range restricted contraction)
    long long ntiles; // tiles visited per matrix
    double* partial;  // [nblocks][2 + d]
    // batched form (blockIdx.z = entry): per-entry kernel parameters and element strides
    const MaternDev* mdev; long long strideG, strideU, strideX;
};

// Shared memory (dynamic): x tile [d][CT], y tile [d][CT], U rows at the tile's rows [r][CT] and at its columns
// [r][CT] -- every entry of the tile reads its 2 r correction values from there, not from global memory.
template <int P>
__global__ void __launch_bounds__(COV_THREADS, 2) contract_kernel(const __grid_constant__ ContractArgs a) {
    extern __shared__ __align__(16) double csm[];
    __shared__ double wacc[COV_THREADS / 32][GPMP_MAX_DIM + 2];
    __shared__ MaternDev msh;
    __shared__ double etab[EXP_TAB];
    stage_exp_tab(etab);
    if (a.mdev) stage_matern(&msh, a.mdev + blockIdx.z, a.m);
    const MaternDev& m = a.mdev ? msh : a.m;
    MaternRegs<P> mr;
    mr.init(&m, etab);
    const int d = m.d, nr = a.Ut ? a.r : 0;
    double* xs = csm;
    double* ys = csm + (size_t)d * CTP;
    double (*ur)[CT] = reinterpret_cast<double (*)[CT]>(csm + (size_t)2 * d * CTP);
    double (*uc)[CT] = reinterpret_cast<double (*)[CT]>(csm + (size_t)2 * d * CTP + (size_t)nr * CT);
    const long long zb = blockIdx.z;
    const double* __restrict__ Gz = a.G + zb * a.strideG;
    const double* __restrict__ Uz = a.Ut ? a.Ut + zb * a.strideU : nullptr;
    // A CTA walks CONTRACT_TILES_PER_CTA consecutive tiles of the tile list and keeps its running sums to itself
    // (sK, sTr in registers, the d per-dimension sums per thread in shared memory): one block reduction and one
    // partial record per run of tiles instead of per tile.
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4, lane = tid & 31, warp = tid >> 5;
    double* dacc = csm + (size_t)2 * d * CTP + (size_t)2 * nr * CT;  // [d][COV_THREADS]
    for (int j = 0; j < d; ++j) dacc[j * COV_THREADS + tid] = 0.0;
    const long long t_first = (long long)a.tile_off + (long long)blockIdx.x * CONTRACT_TILES_PER_CTA;
    const long long t_end = (long long)a.tile_off + a.ntiles;
    int ti, tj;
    if (a.sym) tri_decode((int)t_first, ti, tj);
    else { ti = (int)(t_first / a.tiles_n); tj = (int)(t_first - (long long)ti * a.tiles_n); }
    // points pre-scaled by c / rho_j (plain 1 / rho_j for the distance vjp): the accumulated distance is t = c h,
    // and the squared differences of the second pass carry c^2, which is folded into the weights
    const double pscale = a.dist_only ? 1.0 : m.c;
    const double inv_c = 1.0 / pscale, inv_c2 = inv_c * inv_c;
    double sK = 0.0, sTr = 0.0;
    int staged_ti = -1;
    for (int it = 0; it < CONTRACT_TILES_PER_CTA && t_first + it < t_end; ++it) {
    const int r0 = ti * CT, c0 = tj * CT;
    if (it > 0) __syncthreads();  // every warp is done with the previous tiles
    if (ti != staged_ti) {
        stage_points(xs, a.x + zb * a.strideX, r0, a.n, m, tid, pscale);
        for (int e = tid; e < nr * CT; e += COV_THREADS) {
            const int q = e / CT, i = e - q * CT;
            ur[q][i] = r0 + i < a.n ? Uz[(long long)q * a.ldu + r0 + i] : 0.0;
        }
        staged_ti = ti;
    }
    stage_points(ys, a.y + zb * a.strideX, c0, a.mcols, m, tid, pscale);
    for (int e = tid; e < nr * CT; e += COV_THREADS) {
        const int q = e / CT, i = e - q * CT;
        uc[q][i] = c0 + i < a.mcols ? Uz[(long long)q * a.ldu + c0 + i] : 0.0;
    }
    __syncthreads();

    double h2[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int k = 0; k < 4; ++k) h2[i][k] = 0.0;
    for (int j = 0; j < d; ++j) {
        const double2 xa = *reinterpret_cast<const double2*>(&xs[j * CTP + ty * 4]);
        const double2 xb = *reinterpret_cast<const double2*>(&xs[j * CTP + ty * 4 + 2]);
        const double2 ya = *reinterpret_cast<const double2*>(&ys[j * CTP + 2 * tx]);
        const double2 yb = *reinterpret_cast<const double2*>(&ys[j * CTP + 32 + 2 * tx]);
        const double xr[4] = {xa.x, xa.y, xb.x, xb.y};
        const double yc[4] = {ya.x, ya.y, yb.x, yb.y};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const double df = xr[i] - yc[k];
                h2[i][k] = fma(df, df, h2[i][k]);
            }
    }
    // G tile entries of this thread: rows r0 + 4 ty + i, columns c0 + {2 tx, 2 tx + 1, 32 + 2 tx, 33 + 2 tx}
    const bool diag_tile = a.sym && ti == tj;
    // interior: no bounds, no triangle mask, no trace entries (a same-set tile on the diagonal has them even in the
    // rectangular form), 16-byte loads
    const bool interior = !(ti == tj && (a.sym || a.same_set)) && r0 + CT <= a.n && c0 + CT <= a.mcols &&
                          (a.ldg & 1) == 0 && (reinterpret_cast<uintptr_t>(Gz) & 15) == 0;
    double gv[4][4];
    if (interior) {
        const double* gp = Gz + (long long)(r0 + ty * 4) * a.ldg + c0 + 2 * tx;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const double2 g0 = *reinterpret_cast<const double2*>(gp);
            const double2 g1 = *reinterpret_cast<const double2*>(gp + 32);
            gv[i][0] = g0.x; gv[i][1] = g0.y; gv[i][2] = g1.x; gv[i][3] = g1.y;
            gp += a.ldg;
        }
    } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int row = r0 + ty * 4 + i;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int col = c0 + 32 * (k >> 1) + 2 * tx + (k & 1);
                const bool live = row < a.n && col < a.mcols && !(diag_tile && col > row);
                gv[i][k] = live ? Gz[(long long)row * a.ldg + col] : 0.0;
            }
        }
    }
    // low-rank correction G - U^T U from the staged rows (dead entries keep 0 through the weight below)
    for (int q = 0; q < nr; ++q) {
        const double2 ua = *reinterpret_cast<const double2*>(&ur[q][ty * 4]);
        const double2 ub = *reinterpret_cast<const double2*>(&ur[q][ty * 4 + 2]);
        const double2 va = *reinterpret_cast<const double2*>(&uc[q][2 * tx]);
        const double2 vb = *reinterpret_cast<const double2*>(&uc[q][32 + 2 * tx]);
        const double uu[4] = {ua.x, ua.y, ub.x, ub.y};
        const double vv[4] = {va.x, va.y, vb.x, vb.y};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int k = 0; k < 4; ++k) gv[i][k] = fma(-uu[i], vv[k], gv[i][k]);
    }
    // weights w = G_ik * sigma2 * k'(h)/h / c^2 (x2 for strictly-lower entries in sym mode)
    double w[4][4];
    const double off_mult = a.sym ? 2.0 : 1.0;  // off-diagonal tiles of the symmetric form count twice
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int row = r0 + ty * 4 + i;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int col = c0 + 32 * (k >> 1) + 2 * tx + (k & 1);
            double mult = off_mult, g = gv[i][k];
            if (!interior) {
                const bool live = row < a.n && col < a.mcols && !(diag_tile && col > row);
                if (a.sym && col >= row) mult = 1.0;
                if (!live) { g = 0.0; mult = 0.0; }
                if (a.same_set && live && row == col) sTr += g;
            }
            const double t = sqrt_seeded(h2[i][k]);
            double dkh, kc;
            if (a.dist_only) {
                dkh = t > 0.0 ? 1.0 / t : 0.0;  // reference: custom_sqrt has zero gradient at 0
                kc = 0.0;
            } else {
                mr.cov_and_dk_t(t, kc, dkh);  // both already carry sigma2
                const bool p0 = P >= 0 ? (P == 0) : (m.p == 0);
                if (p0) dkh = t > 0.0 ? dkh / (t * inv_c) : 0.0;  // reference: masked sqrt gradient at 0
            }
            const double mg = mult * g;
            sK = fma(mg, kc, sK);
            w[i][k] = mg * (dkh * inv_c2);
        }
    }
    // per-dimension sums of this tile, added to the thread's running sums
    for (int j = 0; j < d; ++j) {
        const double2 xa = *reinterpret_cast<const double2*>(&xs[j * CTP + ty * 4]);
        const double2 xb = *reinterpret_cast<const double2*>(&xs[j * CTP + ty * 4 + 2]);
        const double2 ya = *reinterpret_cast<const double2*>(&ys[j * CTP + 2 * tx]);
        const double2 yb = *reinterpret_cast<const double2*>(&ys[j * CTP + 32 + 2 * tx]);
        const double xr[4] = {xa.x, xa.y, xb.x, xb.y};
        const double yc[4] = {ya.x, ya.y, yb.x, yb.y};
        double s0 = 0.0, s1 = 0.0;
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int k = 0; k < 4; k += 2) {
                const double d0 = xr[i] - yc[k], d1 = xr[i] - yc[k + 1];
                s0 = fma(w[i][k], d0 * d0, s0);
                s1 = fma(w[i][k + 1], d1 * d1, s1);
            }
        dacc[j * COV_THREADS + tid] += s0 + s1;
    }
    // next tile of the list
    if (a.sym) {
        if (++tj > ti) { tj = 0; ++ti; }
    } else if (++tj == a.tiles_n) {
        tj = 0;
        ++ti;
    }
    }  // tile loop
    // block reduction of the 2 + d running sums (fixed order: deterministic)
    {
        double t0 = warp_sum(sK), t1 = warp_sum(sTr);
        if (lane == 0) { wacc[warp][0] = t0; wacc[warp][1] = t1; }
    }
    for (int j = 0; j < d; ++j) {
        const double sj = warp_sum(dacc[j * COV_THREADS + tid]);
        if (lane == 0) wacc[warp][2 + j] = sj;
    }
    __syncthreads();
    if (tid < 2 + d) {
        double t = 0.0;
#pragma unroll
        for (int wv = 0; wv < COV_THREADS / 32; ++wv) t += wacc[wv][tid];
        a.partial[((long long)blockIdx.z * gridDim.x + blockIdx.x) * (2 + d) + tid] = t;
    }
}

// deterministic final reduction of the per-tile partials and assembly of d value / d covparam
struct ContractFinalArgs {
    const double* partial; long long nblocks; int d, noise;
    double diag_add, half;  // half = 0.5 for the likelihood (0.5 tr(M dK)), 1.0 for a plain vjp
    double* grad;           // [1 + noise + d]
    const MaternDev* mdev;  // batched form: one block per entry, diag_add from the entry's parameters
};
__global__ void contract_final_kernel(const ContractFinalArgs a) {
    __shared__ double red[40];
    const int nv = 2 + a.d;
    __shared__ double tot[GPMP_MAX_DIM + 2];
    const double* __restrict__ partial = a.partial + (long long)blockIdx.x * a.nblocks * nv;
    double* __restrict__ grad = a.grad + (long long)blockIdx.x * (1 + a.noise + a.d);
    const double diag_add = a.mdev ? a.mdev[blockIdx.x].diag_add : a.diag_add;
    for (int v = 0; v < nv; ++v) {
        double s = 0.0;
        for (long long b = threadIdx.x; b < a.nblocks; b += blockDim.x) s += partial[b * nv + v];
        s = block_sum(s, red);
        if (threadIdx.x == 0) tot[v] = s;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        if (a.noise) {
            grad[0] = a.half * tot[0];
            grad[1] = a.half * diag_add * tot[1];
        } else {
            grad[0] = a.half * (tot[0] + diag_add * tot[1]);
        }
        for (int j = 0; j < a.d; ++j) grad[1 + a.noise + j] = a.half * tot[2 + j];
    }
}

size_t contract_workspace_bytes(int n, int mcols, int d) {
    long long tm = ceil_div(n, CT), tn = ceil_div(mcols, CT);
    return (size_t)(tm * tn) * (2 + d) * sizeof(double);
}

int launch_contract(const gpmp_cov_spec* spec, const double* x, int n, const double* y, int mcols, const double* G,
                    long long ldg, const double* Ut, long long ldu, int r, int sym, int dist_only, double half,
                    double* grad, void* partial, size_t partial_bytes, cudaStream_t stream, int tile_row0,
                    int tile_row1, const ContractBatch* cb) {
    const bool same = (y == nullptr || y == x);
    const int batch = cb ? cb->batch : 1;
    if (batch <= 0) return GPMP_OK;
    if (sym && !same) return GPMP_ERR_ARG;
    if (r > CONTRACT_MAXR) return GPMP_ERR_ARG;
    ContractArgs a;
    int rc = make_matern_dev(spec, &a.m, same);
    if (rc) return rc;
    a.x = x; a.y = same ? x : y;
    a.G = G; a.ldg = ldg; a.Ut = Ut; a.ldu = ldu; a.r = r;
    a.n = n; a.mcols = same ? n : mcols; a.sym = sym; a.same_set = (same && !dist_only) ? 1 : 0;
    a.dist_only = dist_only;
    const int tm = ceil_div(n, CT), tn = ceil_div(a.mcols, CT);
    a.tiles_n = tn;
    long long nblocks = sym ? (long long)tm * (tm + 1) / 2 : (long long)tm * tn;
    a.tile_off = 0;
    if (sym && (tile_row0 > 0 || tile_row1 >= 0)) {
        // lower tiles of the tile rows [tile_row0, tile_row1) only (row-partitioned contraction)
        if (tile_row1 < 0 || tile_row1 > tm) tile_row1 = tm;
        if (tile_row0 >= tile_row1) tile_row1 = tile_row0;
        a.tile_off = (int)((long long)tile_row0 * (tile_row0 + 1) / 2);
        nblocks = (long long)tile_row1 * (tile_row1 + 1) / 2 - a.tile_off;
    }
    a.ntiles = nblocks;
    nblocks = ceil_div_ll(nblocks, CONTRACT_TILES_PER_CTA);  // one CTA (and one partial record) per run of tiles
    if ((size_t)nblocks * batch * (2 + spec->d) * sizeof(double) > partial_bytes) return GPMP_ERR_WORKSPACE;
    a.partial = static_cast<double*>(partial);
    a.mdev = cb ? cb->mdev : nullptr;
    a.strideG = cb ? cb->strideG : 0; a.strideU = cb ? cb->strideU : 0; a.strideX = cb ? cb->strideX : 0;
    const dim3 cgrid((unsigned)nblocks, 1, (unsigned)batch);
    {
        double bytes = sym ? 8.0 * n * (n + 1.0) / 2.0 : 8.0 * (double)n * a.mcols;
        LaunchScope scope(KC_CONTRACT, bytes, stream);
        // x / y tiles, the U rows of the tile and the per-thread running sums of the d length-scale derivatives
        // (d x 256 doubles): 14 KB at d = 4, 28 KB at d = 8, 130 KB at the limits d = 32, r = 32 -- launches
        // above 48 KB opt in to the larger carve-out
        const size_t smem = ((size_t)2 * spec->d * CTP + (size_t)2 * (Ut ? r : 0) * CT + (size_t)spec->d * COV_THREADS) * sizeof(double);
        auto launch = [&](auto kern) -> int {
            if (smem > 48 * 1024 &&
                cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
                return GPMP_ERR_CUDA;
            kern<<<cgrid, COV_THREADS, smem, stream>>>(a);
            return GPMP_OK;
        };
        if (nblocks > 0) {
            int lrc;
            switch (dist_only ? 0 : spec->p) {
                case 0: lrc = launch(contract_kernel<0>); break;
                case 1: lrc = launch(contract_kernel<1>); break;
                case 2: lrc = launch(contract_kernel<2>); break;
                case 3: lrc = launch(contract_kernel<3>); break;
                case 4: lrc = launch(contract_kernel<4>); break;
                default: lrc = launch(contract_kernel<-1>); break;
            }
            if (lrc) return lrc;
        }
        GPMP_CHECK_LAUNCH();
    }
    ContractFinalArgs f;
    f.partial = a.partial; f.nblocks = nblocks; f.d = spec->d; f.noise = spec->noise;
    f.diag_add = a.m.diag_add; f.half = half; f.grad = grad; f.mdev = a.mdev;
    {
        LaunchScope scope(KC_SMALL, 0.0, stream);
        contract_final_kernel<<<batch, 256, 0, stream>>>(f);
        GPMP_CHECK_LAUNCH();
    }
    return GPMP_OK;
}

int launch_pairwise(const gpmp_cov_spec* spec, const double* x, const double* y, int n, double* out, int dist_only,
                    cudaStream_t stream) {
    if (n <= 0) return GPMP_OK;
    PairArgs a;
    int rc = make_matern_dev(spec, &a.m, false);
    if (rc) return rc;
    a.x = x; a.y = (y == x) ? nullptr : y; a.out = out; a.n = n; a.dist_only = dist_only;
    LaunchScope scope(KC_SMALL, 0.0, stream);
    pairwise_kernel<<<ceil_div(n, 256), 256, 0, stream>>>(a);
    GPMP_CHECK_LAUNCH();
    return GPMP_OK;
}

int launch_maternp_elementwise(int p, const double* h, double* k, double* dk, long long count, cudaStream_t stream) {
    if (count <= 0) return GPMP_OK;
    gpmp_cov_spec s;
    s.p = p; s.d = 1; s.noise = 0; s.reserved = 0; s.log_sigma2 = 0.0; s.log_tau2 = 0.0;
    for (int j = 0; j < GPMP_MAX_DIM; ++j) s.loginvrho[j] = 0.0;
    KernArgs a;
    int rc = make_matern_dev(&s, &a.m, false);
    if (rc) return rc;
    a.h = h; a.k = k; a.dk = dk; a.count = count;
    long long blocks = ceil_div_ll(count, 256);
    if (blocks > 148 * 16) blocks = 148 * 16;
    LaunchScope scope(KC_SMALL, 0.0, stream);
    maternp_elementwise_kernel<<<(unsigned)blocks, 256, 0, stream>>>(a);
    GPMP_CHECK_LAUNCH();
    return GPMP_OK;
}

}  // namespace gpmp
