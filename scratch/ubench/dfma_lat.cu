// Scalar FP64 (DFMA) latency and throughput on one SM: W warps per SM, C independent dependent-chains per thread.
#include <cstdio>
#include <cuda_runtime.h>
template <int C>
__global__ void k(double *out, long long *cyc, int iters, double a, double b) {
    double c[C];
#pragma unroll
    for (int i = 0; i < C; ++i) c[i] = threadIdx.x * 1e-3 + i;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int i = 0; i < C; ++i) c[i] = fma(c[i], a, b);
    }
    long long t1 = clock64();
    double r = 0;
#pragma unroll
    for (int i = 0; i < C; ++i) r += c[i];
    if (r == 123.456) out[0] = r;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
template <int C>
void run(int warps) {
    double *out; long long *cyc, h;
    cudaMalloc(&out, 8); cudaMalloc(&cyc, 8);
    const int iters = 2048;
    k<C><<<148, warps * 32>>>(out, cyc, iters, 0.999999, 1e-9);
    k<C><<<148, warps * 32>>>(out, cyc, iters, 0.999999, 1e-9);
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    const double per = (double)h / (iters * 8.0);            // clocks per round of C DFMAs per thread
    printf("warps/SM %2d chains %d : %.1f clk per dependent step, %.2f clk per DFMA warp-instruction per SM sub-partition\n",
           warps, C, per, per / C / (warps / 4.0 < 1 ? 1 : warps / 4.0));
    cudaFree(out); cudaFree(cyc);
}
int main() {
    run<1>(1); run<2>(1); run<4>(1); run<8>(1);
    run<1>(4); run<2>(4); run<4>(4); run<8>(4);
    run<1>(16); run<2>(16); run<4>(16); run<8>(16);
    run<4>(32); run<8>(64);
    printf("status %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
