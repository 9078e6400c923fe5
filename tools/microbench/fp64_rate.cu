// Micro-benchmark: sustained FP64 issue rate of one SM as a function of the
// instruction (DFMA / DADD / DMUL), the operand pattern and the resident warps.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_rate fp64_rate.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void k(double *out, double a, double b, int iters) {
    double acc[8], w[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        acc[r] = threadIdx.x * 1e-9 + r;
        w[r] = a + r * 1e-3;
    }
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                if (MODE == 0) acc[r] = fma(b, w[(r + u) % 8], acc[r]);        // FIR pattern
                if (MODE == 1) acc[r] = acc[r] + w[(r + u) % 8];               // DADD
                if (MODE == 2) acc[r] = acc[r] * w[(r + u) % 8];               // DMUL
                if (MODE == 3) acc[r] = fma(acc[r], b, w[(r + u) % 8]);        // FMA, other operand order
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int r = 0; r < 8; ++r) s += acc[r];
    if (s == 12345.678) out[0] = s;
}

__constant__ double ctaps[2048];

// FIR-like mix: 8 DFMA per window load (LDS.64, conflict-free stride 9) and,
// for TAPS, one broadcast LDS.128 per 16 DFMA.
template <int TAPS, bool PINGPONG>
__global__ void kfir(double *out, int iters) {
    __shared__ __align__(16) double xs[32 * 9 + 9 * 64 + 64];
    __shared__ __align__(16) double gs[512];
    for (int i = threadIdx.x; i < 32 * 9 + 9 * 64 + 64; i += blockDim.x) xs[i] = 1.0 + i * 1e-9;
    for (int i = threadIdx.x; i < 512; i += blockDim.x) gs[i] = 1e-3 + i * 1e-9;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    double acc[8], w[8], wb[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) acc[r] = 0.0;
    for (int i = 0; i < iters; ++i) {
        const double *px = xs + lane * 9;
        const double2 *gp = reinterpret_cast<const double2 *>(gs);
#pragma unroll
        for (int r = 0; r < 8; ++r) w[r] = px[r];
        for (int b = 0; b < 64; ++b) {        // 64 blocks of 8 taps
            if (PINGPONG) {
#pragma unroll
                for (int r = 0; r < 8; ++r) wb[r] = px[9 + r];
            }
#pragma unroll
            for (int u = 0; u < 8; u += 2) {
                double2 g2 = make_double2(1e-3, 2e-3);
                if (TAPS == 1) g2 = gp[u / 2];
                if (TAPS == 2) g2 = make_double2(gs[(b & 63) * 8 + u], gs[(b & 63) * 8 + u + 1]);
                if (TAPS == 3) {
                    const int wq = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0) & 15;
                    g2 = make_double2(ctaps[wq * 80 + (b % 10) * 8 + u], ctaps[wq * 80 + (b % 10) * 8 + u + 1]);
                }
                if (PINGPONG) {
#pragma unroll
                    for (int r = 0; r < 8; ++r) acc[r] = fma(g2.x, r + u < 8 ? w[(r + u) % 8] : wb[(r + u) % 8], acc[r]);
#pragma unroll
                    for (int r = 0; r < 8; ++r) acc[r] = fma(g2.y, r + u + 1 < 8 ? w[(r + u + 1) % 8] : wb[(r + u + 1) % 8], acc[r]);
                } else {
#pragma unroll
                    for (int r = 0; r < 8; ++r) acc[r] = fma(g2.x, w[(r + u) % 8], acc[r]);
                    w[u] = px[u + 9];
#pragma unroll
                    for (int r = 0; r < 8; ++r) acc[r] = fma(g2.y, w[(r + u + 1) % 8], acc[r]);
                    w[u + 1] = px[u + 1 + 9];
                }
            }
            if (PINGPONG) {
#pragma unroll
                for (int r = 0; r < 8; ++r) w[r] = wb[r];
            }
            gp += 4;
            px += 9;
        }
    }
    double s = 0;
#pragma unroll
    for (int r = 0; r < 8; ++r) s += acc[r];
    if (s == 12345.678) out[0] = s;
}

template <int TAPS, bool PP>
void runfir(const char *name, int threads, int nsm) {
    double *out;
    cudaMalloc(&out, 8);
    const int iters = 60;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    kfir<TAPS, PP><<<nsm, threads>>>(out, 2);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    kfir<TAPS, PP><<<nsm, threads>>>(out, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    int clk;
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const double inst = (double)iters * 64 * 64 * threads;
    printf("%-28s threads/SM %4d: %.3f ms  %.1f DFMA lane-ops/clk/SM\n", name, threads, ms,
           inst / (ms * 1e-3 * clk * 1e3));
    cudaFree(out);
}

template <int MODE>
void run(const char *name, int threads, int nsm) {
    double *out;
    cudaMalloc(&out, 8);
    const int iters = 4000;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    k<MODE><<<nsm, threads>>>(out, 1.0000001, 0.9999999, 10);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<MODE><<<nsm, threads>>>(out, 1.0000001, 0.9999999, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    int clk;
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const double inst = (double)iters * 64 * threads;          // thread-instructions per SM
    const double per_clk = inst / (ms * 1e-3 * clk * 1e3);
    printf("%-28s threads/SM %4d: %.3f ms  %.1f lane-ops/clk/SM (at %d MHz nominal)\n", name, threads,
           ms, per_clk, clk / 1000);
    cudaFree(out);
}

int main() {
    int nsm;
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
    for (int t : {128, 256, 512, 1024}) {
        run<0>("DFMA acc=fma(b,w,acc)", t, nsm);
        run<3>("DFMA acc=fma(acc,b,w)", t, nsm);
        run<1>("DADD", t, nsm);
        run<2>("DMUL", t, nsm);
    }
    for (int t : {256, 512}) {
        runfir<0, false>("FIR window LDS.64", t, nsm);
        runfir<1, false>("FIR window + tap LDS.128", t, nsm);
        runfir<2, false>("FIR window + tap LDS.64 x2", t, nsm);
        runfir<3, false>("FIR window + tap const", t, nsm);
    }
    return 0;
}
