// Emulates the consumer loop of the state sweep: per batch 26 DMMA in 2 chains, F dependent DFMAs (+SH shuffles), 26 DMMA in 13 chains of 2.
#include <cstdio>
#include <cuda_runtime.h>
#define DMMA(c0, c1, a, b) asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b))
template <int F, int SH, int LDS>
__global__ void __launch_bounds__(512, 1) k(double *out, long long *cyc, int iters, double a, double b) {
    __shared__ double sm[8 * 104];
    for (int i = threadIdx.x; i < 8 * 104; i += blockDim.x) sm[i] = 1e-3 * i;
    __syncthreads();
    double x[26];
#pragma unroll
    for (int i = 0; i < 26; ++i) x[i] = threadIdx.x * 1e-3 + i;
    const int lane = threadIdx.x & 31, c = lane & 3, n = lane >> 2;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        double ga0 = 0, ga1 = 0, gb0 = 0, gb1 = 0;
#pragma unroll
        for (int t = 0; t < 13; ++t) {
            double v0 = b, v1 = b;
            if (LDS) { const double2 v = *reinterpret_cast<const double2 *>(sm + n * 104 + 8 * t + 2 * c); v0 = v.x; v1 = v.y; }
            DMMA(ga0, ga1, x[2 * t], v0);
            DMMA(gb0, gb1, x[2 * t + 1], v1);
        }
        double e = ga0 + gb0, f = ga1 + gb1;
#pragma unroll
        for (int s = 0; s < SH; ++s) { e += __shfl_xor_sync(0xffffffffu, e, 1 + (s & 1)); }
#pragma unroll
        for (int i = 0; i < F; ++i) e = fma(e, 0.999, f);
        const double ea0 = e * 1e-9, ea1 = f * 1e-9;
#pragma unroll
        for (int t = 0; t < 13; ++t) {
            double y0 = a, y1 = a;
            if (LDS) { y0 = sm[c * 104 + 8 * t + n]; y1 = sm[(4 + c) * 104 + 8 * t + n]; }
            DMMA(x[2 * t], x[2 * t + 1], ea0, y0);
            DMMA(x[2 * t], x[2 * t + 1], ea1, y1);
        }
    }
    long long t1 = clock64();
    double r = 0;
#pragma unroll
    for (int i = 0; i < 26; ++i) r += x[i];
    if (r == 123.456) out[0] = r;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
template <int F, int SH, int LDS>
void run(int warps) {
    double *out; long long *cyc, h;
    cudaMalloc(&out, 8); cudaMalloc(&cyc, 8);
    const int iters = 2000;
    k<F, SH, LDS><<<148, warps * 32>>>(out, cyc, iters, 1e-3, 1e-3);
    k<F, SH, LDS><<<148, warps * 32>>>(out, cyc, iters, 1e-3, 1e-3);
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    const double per_batch = (double)h / iters;               // per warp
    const double util = 52.0 * 16.0 * (warps / 4.0) / per_batch;
    printf("warps %2d F %2d SH %d LDS %d : %.0f clk per warp-batch, DMMA pipe utilisation %.1f%%\n", warps, F, SH, LDS, per_batch, 100 * util);
    cudaFree(out); cudaFree(cyc);
}
int main() {
    run<0, 0, 0>(4); run<0, 0, 0>(12); run<30, 6, 0>(4); run<30, 6, 0>(8); run<30, 6, 0>(12); run<30, 6, 1>(12); run<30, 6, 1>(16);
    run<60, 6, 1>(12); run<15, 6, 1>(12); run<30, 0, 1>(12); run<0, 0, 1>(12); run<60, 6, 1>(16);
    return 0;
}
