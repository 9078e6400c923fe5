// Micro-benchmark: packed FP32 (FFMA2 / FADD2, sm_100a) against scalar FFMA / FADD:
// does a packed instruction deliver two results per issue slot at the scalar rate?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp32x2_rate fp32x2_rate.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void k(float *out, float a, float b, int iters) {
    float2 acc[8], w[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        acc[r] = make_float2(threadIdx.x * 1e-9f + r, r * 0.5f);
        w[r] = make_float2(a + r * 1e-3f, b + r * 1e-3f);
    }
    const float2 bb = make_float2(b, b);
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                const float2 ww = w[(r + u) % 8];
                if (MODE == 0) {                       // scalar FFMA x2
                    acc[r].x = fmaf(b, ww.x, acc[r].x);
                    acc[r].y = fmaf(b, ww.y, acc[r].y);
                }
                if (MODE == 1) acc[r] = __ffma2_rn(bb, ww, acc[r]);     // FFMA2
                if (MODE == 2) {                       // scalar FADD x2
                    acc[r].x += ww.x;
                    acc[r].y += ww.y;
                }
                if (MODE == 3) acc[r] = __fadd2_rn(acc[r], ww);        // FADD2
            }
        }
    }
    float s = 0;
#pragma unroll
    for (int r = 0; r < 8; ++r) s += acc[r].x + acc[r].y;
    if (s == 12345.678f) out[0] = s;
}

template <int MODE>
void run(const char *name, int threads) {
    float *out;
    cudaMalloc(&out, 4);
    const int iters = 4000, blocks = 148;
    k<MODE><<<blocks, threads>>>(out, 1.0f, 1e-7f, 10);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<MODE><<<blocks, threads>>>(out, 1.0f, 1e-7f, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    int clk_khz;
    cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    const double results = 2.0 * 64.0 * iters * threads;        // per SM (one CTA per SM)
    const double clocks = ms * 1e-3 * clk_khz * 1e3;
    printf("%-12s threads/SM %4d  %7.3f ms  %6.1f FP32 results/clk/SM (nominal clock %d MHz)\n",
           name, threads, ms, results / clocks, clk_khz / 1000);
    cudaFree(out);
}

int main() {
    for (int threads : {256, 512, 1024}) {
        run<0>("FFMA x2", threads);
        run<1>("FFMA2", threads);
        run<2>("FADD x2", threads);
        run<3>("FADD2", threads);
    }
    return 0;
}
