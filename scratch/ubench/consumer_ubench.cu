// Consumer loop of state_sweep_2p.cu in isolation (static shared-memory stage, no barriers, no producers): which part
// of the per-batch work keeps the FP64 tensor pipe from its 16-clock issue interval per SM sub-partition?
//   REC    0: e = omega * g (no serial chain)   1: the full 8-step recurrence (30 scalar FP64 + 1 DMMA + 2 SHFL.64)
//          2: the recurrence with its scalar arithmetic in FP32 (upper bound if it did not use the FP64 pipe)
//   GATHER 0: no all-gather shuffles            1: the 6 SHFL.64 of the kernel
//   LDS    0: B operands from registers         1: from shared memory, as in the kernel
//   TILES  row tiles of 8 rows per warp (1 = the kernel; 2 = two independent tiles interleaved in one warp)
#include <cstdio>
#include <cuda_runtime.h>
#define NT3 13
#define YST 104
__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmma_v(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
template <int REC, int GATHER, int LDS, int TILES, int NTHR = 512, int CH = 2>
__global__ void __launch_bounds__(TILES == 2 ? 256 : NTHR, 1) k(double *out, long long *cyc, int iters, double areg, double breg) {
    __shared__ __align__(16) double sy[8 * YST + 16];
    __shared__ __align__(16) double som[48 * 8];
    __shared__ double Gb[64];
    __shared__ __align__(16) double xch[16 * 32 * 2];
    for (int i = threadIdx.x; i < 8 * YST + 16; i += blockDim.x) sy[i] = 1e-3 * ((i * 7) % 13) - 5e-3;
    for (int i = threadIdx.x; i < 48 * 8; i += blockDim.x) som[i] = 1e-4 * (1 + i % 5);
    for (int i = threadIdx.x; i < 64; i += blockDim.x) Gb[i] = 1e-2 * (1 + i % 3);
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, c = lane & 3, n = lane >> 2;
    double x[TILES][2 * NT3];
#pragma unroll
    for (int tl = 0; tl < TILES; ++tl)
#pragma unroll
        for (int i = 0; i < 2 * NT3; ++i) x[tl][i] = threadIdx.x * 1e-3 + i + tl;
    const int gslot = (warp * 8 + n) % 40;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        double ea0[TILES], ea1[TILES];
#pragma unroll
        for (int tl = 0; tl < TILES; ++tl) {
            double om[8];
            const double2 *po = reinterpret_cast<const double2 *>(som + gslot * 8);
#pragma unroll
            for (int i = 0; i < 4; ++i) { const double2 v = po[i]; om[2 * i] = v.x; om[2 * i + 1] = v.y; }
            double ga0 = 0.0, ga1 = 0.0, gb0 = 0.0, gb1 = 0.0;
            const double *yrow = sy + n * YST + (n >> 1) * 4 + 2 * c;
#pragma unroll
            for (int t = 0; t < NT3; ++t) {
                double v0 = breg, v1 = breg;
                if (LDS) { const double2 v = *reinterpret_cast<const double2 *>(yrow + 8 * t); v0 = v.x; v1 = v.y; }
                dmma(ga0, ga1, x[tl][2 * t], v0);
                if (CH == 2) dmma(gb0, gb1, x[tl][2 * t + 1], v1); else dmma(ga0, ga1, x[tl][2 * t + 1], v1);
            }
            if (CH == 2) { ga0 += gb0; ga1 += gb1; }
            double gq[8];
            if (GATHER) {
                const double o0 = __shfl_xor_sync(0xffffffffu, ga0, 1), o1 = __shfl_xor_sync(0xffffffffu, ga1, 1);
                double q0, q1, q2, q3;
                if (c & 1) { q0 = o0; q1 = o1; q2 = ga0; q3 = ga1; } else { q0 = ga0; q1 = ga1; q2 = o0; q3 = o1; }
                const double r0 = __shfl_xor_sync(0xffffffffu, q0, 2), r1 = __shfl_xor_sync(0xffffffffu, q1, 2);
                const double r2 = __shfl_xor_sync(0xffffffffu, q2, 2), r3 = __shfl_xor_sync(0xffffffffu, q3, 2);
                if (c & 2) { gq[0] = r0; gq[1] = r1; gq[2] = r2; gq[3] = r3; gq[4] = q0; gq[5] = q1; gq[6] = q2; gq[7] = q3; }
                else { gq[0] = q0; gq[1] = q1; gq[2] = q2; gq[3] = q3; gq[4] = r0; gq[5] = r1; gq[6] = r2; gq[7] = r3; }
            } else {
#pragma unroll
                for (int i = 0; i < 8; ++i) gq[i] = (i & 1) ? ga1 : ga0;
            }
            double e[8];
            if (REC == 1) {
                e[0] = om[0] * gq[0];
                e[1] = om[1] * (gq[1] - Gb[8] * e[0]);
                e[2] = om[2] * (gq[2] - Gb[16] * e[0] - Gb[17] * e[1]);
                e[3] = om[3] * (gq[3] - Gb[24] * e[0] - Gb[25] * e[1] - Gb[26] * e[2]);
                const double ea = (c == 0) ? e[0] : (c == 1) ? e[1] : (c == 2) ? e[2] : e[3];
                double k0 = 0.0, k1 = 0.0;
                dmma(k0, k1, ea, Gb[(4 + (n & 3)) * 8 + c]);
                const double o0 = __shfl_xor_sync(0xffffffffu, k0, 1), o1 = __shfl_xor_sync(0xffffffffu, k1, 1);
                if (c & 1) { gq[4] -= o0; gq[5] -= o1; gq[6] -= k0; gq[7] -= k1; }
                else { gq[4] -= k0; gq[5] -= k1; gq[6] -= o0; gq[7] -= o1; }
                e[4] = om[4] * gq[4];
                e[5] = om[5] * (gq[5] - Gb[44] * e[4]);
                e[6] = om[6] * (gq[6] - Gb[52] * e[4] - Gb[53] * e[5]);
                e[7] = om[7] * (gq[7] - Gb[60] * e[4] - Gb[61] * e[5] - Gb[62] * e[6]);
            } else if (REC == 4) {
                // all-gather through shared memory: every lane stores its two dots, the row's lanes read four back
                double2 *xg = reinterpret_cast<double2 *>(xch) + (warp * 32 + (lane & ~3));
                xg[c] = make_double2(ga0, ga1);
                __syncwarp();
                const double2 t0 = xg[0], t1 = xg[1];
                gq[0] = t0.x; gq[1] = t0.y; gq[2] = t1.x; gq[3] = t1.y;
                e[0] = om[0] * gq[0];
                e[1] = om[1] * (gq[1] - Gb[8] * e[0]);
                e[2] = om[2] * (gq[2] - Gb[16] * e[0] - Gb[17] * e[1]);
                e[3] = om[3] * (gq[3] - Gb[24] * e[0] - Gb[25] * e[1] - Gb[26] * e[2]);
                const double ea = -((c == 0) ? e[0] : (c == 1) ? e[1] : (c == 2) ? e[2] : e[3]);
                double k0 = ga0, k1 = ga1;
                dmma(k0, k1, ea, (n >= 4) ? Gb[n * 8 + c] : 0.0);
                __syncwarp();
                xg[c] = make_double2(k0, k1);
                __syncwarp();
                const double2 t2 = xg[2], t3 = xg[3];
                gq[4] = t2.x; gq[5] = t2.y; gq[6] = t3.x; gq[7] = t3.y;
                e[4] = om[4] * gq[4];
                e[5] = om[5] * (gq[5] - Gb[44] * e[4]);
                e[6] = om[6] * (gq[6] - Gb[52] * e[4] - Gb[53] * e[5]);
                e[7] = om[7] * (gq[7] - Gb[60] * e[4] - Gb[61] * e[5] - Gb[62] * e[6]);
            } else if (REC == 3) {
                e[0] = om[0] * gq[0];
                e[1] = om[1] * (gq[1] - Gb[8] * e[0]);
                e[2] = om[2] * (gq[2] - Gb[16] * e[0] - Gb[17] * e[1]);
                e[3] = om[3] * (gq[3] - Gb[24] * e[0] - Gb[25] * e[1] - Gb[26] * e[2]);
                const double ea = -((c == 0) ? e[0] : (c == 1) ? e[1] : (c == 2) ? e[2] : e[3]);
                double k0 = ga0, k1 = ga1;              // lanes c >= 2 hold g[4..7] in fragment layout
                dmma(k0, k1, ea, (n >= 4) ? Gb[n * 8 + c] : 0.0);
                const double a0 = __shfl_xor_sync(0xffffffffu, k0, 2), a1 = __shfl_xor_sync(0xffffffffu, k1, 2);
                const double m0 = (c & 2) ? k0 : a0, m1 = (c & 2) ? k1 : a1;     // c even side: obs (4,5) ; odd side: (6,7)
                const double b0 = __shfl_xor_sync(0xffffffffu, m0, 1), b1 = __shfl_xor_sync(0xffffffffu, m1, 1);
                if (c & 1) { gq[4] = b0; gq[5] = b1; gq[6] = m0; gq[7] = m1; } else { gq[4] = m0; gq[5] = m1; gq[6] = b0; gq[7] = b1; }
                e[4] = om[4] * gq[4];
                e[5] = om[5] * (gq[5] - Gb[44] * e[4]);
                e[6] = om[6] * (gq[6] - Gb[52] * e[4] - Gb[53] * e[5]);
                e[7] = om[7] * (gq[7] - Gb[60] * e[4] - Gb[61] * e[5] - Gb[62] * e[6]);
            } else if (REC == 2) {
                float ef[8], gf[8], of[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) { gf[i] = (float)gq[i]; of[i] = (float)om[i]; }
                const float G8 = (float)Gb[8], G16 = (float)Gb[16], G17 = (float)Gb[17], G24 = (float)Gb[24], G25 = (float)Gb[25], G26 = (float)Gb[26];
                ef[0] = of[0] * gf[0];
                ef[1] = of[1] * (gf[1] - G8 * ef[0]);
                ef[2] = of[2] * (gf[2] - G16 * ef[0] - G17 * ef[1]);
                ef[3] = of[3] * (gf[3] - G24 * ef[0] - G25 * ef[1] - G26 * ef[2]);
                ef[4] = of[4] * (gf[4] - G8 * ef[3]);
                ef[5] = of[5] * (gf[5] - G8 * ef[4]);
                ef[6] = of[6] * (gf[6] - G16 * ef[4] - G17 * ef[5]);
                ef[7] = of[7] * (gf[7] - G24 * ef[4] - G25 * ef[5] - G26 * ef[6]);
#pragma unroll
                for (int i = 0; i < 8; ++i) e[i] = (double)ef[i];
            } else {
#pragma unroll
                for (int i = 0; i < 8; ++i) e[i] = om[i] * gq[i];
            }
            ea0[tl] = -((c == 0) ? e[0] : (c == 1) ? e[1] : (c == 2) ? e[2] : e[3]);
            ea1[tl] = -((c == 0) ? e[4] : (c == 1) ? e[5] : (c == 2) ? e[6] : e[7]);
        }
#pragma unroll
        for (int tl = 0; tl < TILES; ++tl) {
            const double *y0p = sy + c * YST + (c >> 1) * 4 + n;
            const double *y1p = sy + (4 + c) * YST + ((4 + c) >> 1) * 4 + n;
#pragma unroll
            for (int t = 0; t < NT3; ++t) dmma_v(x[tl][2 * t], x[tl][2 * t + 1], ea0[tl], LDS ? y0p[8 * t] : areg);
#pragma unroll
            for (int t = 0; t < NT3; ++t) dmma_v(x[tl][2 * t], x[tl][2 * t + 1], ea1[tl], LDS ? y1p[8 * t] : areg);
        }
    }
    long long t1 = clock64();
    double r = 0;
#pragma unroll
    for (int tl = 0; tl < TILES; ++tl)
#pragma unroll
        for (int i = 0; i < 2 * NT3; ++i) r += x[tl][i];
    if (r == 123.456) out[0] = r;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
template <int REC, int GATHER, int LDS, int TILES, int NTHR = 512, int CH = 2>
void run(int warps) {
    double *out; long long *cyc, h;
    cudaMalloc(&out, 8); cudaMalloc(&cyc, 8);
    const int iters = 3000;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<REC, GATHER, LDS, TILES, NTHR, CH><<<148, warps * 32>>>(out, cyc, iters, 1e-3, 1e-3);
    cudaEventRecord(e0);
    k<REC, GATHER, LDS, TILES, NTHR, CH><<<148, warps * 32>>>(out, cyc, iters, 1e-3, 1e-3);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    const double per_batch = (double)h / iters;               // clocks per loop iteration of one warp
    const double dm = (52.0 + ((REC == 1 || REC == 3 || REC == 4) ? 1 : 0)) * TILES;   // DMMAs per iteration per warp
    const double util = dm * 16.0 * (warps / 4.0) / per_batch;
    const double tf = dm * 512.0 * warps * 148.0 * iters / (ms * 1e-3) / 1e12;
    printf("warps %2d tiles %d REC %d GATHER %d LDS %d NTHR %d CH %d : %6.0f clk per warp-iteration, DMMA pipe %5.1f%% of 1/16clk/SMSP, %.1f TFLOP/s (rows/SM %3d -> %5.1f clk per row-batch)\n",
           warps, TILES, REC, GATHER, LDS, NTHR, CH, per_batch, 100 * util, tf, warps * 8 * TILES, per_batch / (warps * 8.0 * TILES));
    cudaFree(out); cudaFree(cyc);
}
int main() {
    cudaError_t e = cudaGetLastError();
    for (int w : {12, 16}) run<3, 1, 1, 1, 512, 1>(w);
    for (int w : {12, 16}) run<4, 0, 1, 1, 512, 1>(w);       // all-gather through shared memory
    e = cudaDeviceSynchronize();
    printf("status %s\n", cudaGetErrorString(e));
    return 0;
}
