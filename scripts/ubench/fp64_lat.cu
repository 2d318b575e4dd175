// Development microbenchmark: latency / throughput of the FP64 instructions the tile kernels depend on.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_lat fp64_lat.cu
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ double frsqrt(double x) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double h = 0.5 * x;
    y = y * fma(-h * y, y, 1.5);
    y = y * fma(-h * y, y, 1.5);
    return y;
}
template <int ILP>
__global__ void k_dfma(double* out, long long* clk, int iters, double a, double b) {
    double x[ILP];
    for (int i = 0; i < ILP; ++i) x[i] = threadIdx.x + i;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int i = 0; i < ILP; ++i) x[i] = fma(x[i], a, b);
    long long t1 = clock64();
    double s = 0;
    for (int i = 0; i < ILP; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) clk[0] = t1 - t0;
}
template <int ILP>
__global__ void k_dmma(double* out, long long* clk, int iters, double a, double b) {
    double c0[ILP], c1[ILP];
    for (int i = 0; i < ILP; ++i) c0[i] = c1[i] = threadIdx.x + i;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int i = 0; i < ILP; ++i) dmma(c0[i], c1[i], a, b);
    long long t1 = clock64();
    double s = 0;
    for (int i = 0; i < ILP; ++i) s += c0[i] + c1[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) clk[0] = t1 - t0;
}
__global__ void k_rsqrt(double* out, long long* clk, int iters, double a) {
    double x = 2.0 + threadIdx.x;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) x = frsqrt(x) + a;
    long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) clk[0] = t1 - t0;
}
__global__ void k_sqrtdiv(double* out, long long* clk, int iters, double a) {
    double x = 2.0 + threadIdx.x;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) x = 1.0 / sqrt(x) + a;
    long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) clk[0] = t1 - t0;
}
__global__ void k_sync(double* out, long long* clk, int iters) {
    __shared__ double s[1024];
    double x = threadIdx.x;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        s[threadIdx.x] = x;
        __syncthreads();
        x += s[(threadIdx.x + 33) % blockDim.x];
        __syncthreads();
    }
    long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) clk[0] = t1 - t0;
}
__global__ void k_shfl(double* out, long long* clk, int iters) {
    double x = threadIdx.x;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) x = fma(__shfl_sync(0xffffffffu, x, (it & 31)), 1.0000001, x);
    long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) clk[0] = t1 - t0;
}
int main() {
    double* out; long long* clk; long long h;
    cudaMalloc(&out, 1 << 22); cudaMalloc(&clk, 64);
    const int it = 4096;
#define RUN(name, call, per) call; cudaDeviceSynchronize(); call; cudaDeviceSynchronize(); \
    cudaMemcpy(&h, clk, 8, cudaMemcpyDeviceToHost); printf("%-44s %8.2f clk per %s\n", name, (double)h / it, per);
    RUN("DFMA dependent chain (1 warp)", (k_dfma<1><<<1, 32>>>(out, clk, it, 1.0000001, 1e-9)), "fma");
    RUN("DFMA 8 chains (1 warp)", (k_dfma<8><<<1, 32>>>(out, clk, it, 1.0000001, 1e-9)), "8 fma");
    RUN("DFMA 8 chains (16 warps)", (k_dfma<8><<<1, 512>>>(out, clk, it, 1.0000001, 1e-9)), "8 fma/warp");
    RUN("DFMA 8 chains (32 warps)", (k_dfma<8><<<1, 1024>>>(out, clk, it, 1.0000001, 1e-9)), "8 fma/warp");
    RUN("DMMA dependent chain (1 warp)", (k_dmma<1><<<1, 32>>>(out, clk, it, 1.0000001, 1e-9)), "mma");
    RUN("DMMA 4 chains (1 warp)", (k_dmma<4><<<1, 32>>>(out, clk, it, 1.0000001, 1e-9)), "4 mma");
    RUN("DMMA 8 chains (1 warp)", (k_dmma<8><<<1, 32>>>(out, clk, it, 1.0000001, 1e-9)), "8 mma");
    RUN("DMMA 8 chains (4 warps)", (k_dmma<8><<<1, 128>>>(out, clk, it, 1.0000001, 1e-9)), "8 mma/warp");
    RUN("DMMA 8 chains (16 warps)", (k_dmma<8><<<1, 512>>>(out, clk, it, 1.0000001, 1e-9)), "8 mma/warp");
    RUN("fast rsqrt chain", (k_rsqrt<<<1, 32>>>(out, clk, it, 1.5)), "rsqrt+add");
    RUN("1/sqrt chain", (k_sqrtdiv<<<1, 32>>>(out, clk, it, 1.5)), "div+sqrt+add");
    RUN("smem write + 2 bar.sync + read (512 thr)", (k_sync<<<1, 512>>>(out, clk, it)), "round");
    RUN("smem write + 2 bar.sync + read (128 thr)", (k_sync<<<1, 128>>>(out, clk, it)), "round");
    RUN("shfl + dfma chain", (k_shfl<<<1, 32>>>(out, clk, it)), "step");
    return 0;
}
