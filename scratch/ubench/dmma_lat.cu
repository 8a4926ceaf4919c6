// DMMA.8x8x4 latency / throughput microbenchmark: W warps per SM, C independent chains per warp, optional LDS traffic
#include <cstdio>
#include <cuda_runtime.h>
template <int C, int LDS>
__global__ void k(double *out, long long *cyc, int iters, double a, double b) {
    __shared__ double sm[4096];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = 1e-3 * i;
    __syncthreads();
    double c[C][2];
#pragma unroll
    for (int i = 0; i < C; ++i) { c[i][0] = threadIdx.x + i; c[i][1] = 0.5 * i; }
    const double *p = sm + (threadIdx.x & 31) * 2;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < C; ++i) {
            double bb = b;
            if (LDS) { bb = *(volatile double *)(p + ((it * C + i) & 31) * 64); }
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(bb));
        }
    }
    long long t1 = clock64();
    double r = 0;
#pragma unroll
    for (int i = 0; i < C; ++i) r += c[i][0] + c[i][1];
    if (r == 123.456) out[0] = r;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
template <int C, int LDS>
void run(int warps, const char *name) {
    double *out; long long *cyc, h;
    cudaMalloc(&out, 8); cudaMalloc(&cyc, 8);
    const int iters = 4096;
    k<C, LDS><<<148, warps * 32>>>(out, cyc, iters, 1e-3, 1e-3);
    k<C, LDS><<<148, warps * 32>>>(out, cyc, iters, 1e-3, 1e-3);
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    double per = (double)h / (iters * C);
    printf("%-8s warps/SM %2d chains %d lds %d : %.2f clk per DMMA per warp, %.2f clk per DMMA per SM (=> %.1f TFLOP/s at 1.9 GHz)\n", name,
           warps, C, LDS, per, per / warps, 512.0 * 148 * 1.9e9 / (per / warps) / 1e12);
    cudaFree(out); cudaFree(cyc);
}
int main() {
    run<1, 0>(1, "lat");  run<2, 0>(1, "lat"); run<4, 0>(1, "lat"); run<8, 0>(1, "lat");
    run<1, 0>(4, "w4");  run<2, 0>(4, "w4"); run<4, 0>(4, "w4"); run<8, 0>(4, "w4");
    run<2, 0>(12, "w12"); run<4, 0>(12, "w12"); run<8, 0>(12, "w12");
    run<2, 0>(16, "w16"); run<4, 0>(16, "w16");
    run<2, 1>(12, "w12+lds"); run<4, 1>(12, "w12+lds"); run<8, 1>(12, "w12+lds"); run<4, 1>(16, "w16+lds"); run<8, 1>(32, "w32+lds"); run<8, 0>(32, "w32");
    return 0;
}
