// Micro-benchmark: sustained rate of the FP64 tensor-core MMA (mma.sync ... f64)
// on one SM, operands from registers and from shared memory, against the DFMA
// stream of fp64_rate.cu.  Question it answers: can a Toeplitz-GEMM form of the
// decimating FIR (north_star: "tensor cores only if shown to win") deliver more
// FMA lanes per clock than the 43-50 of 64 a three-operand DFMA stream sustains?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_rate dmma_rate.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma884(double (&c)[2], double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c[0]), "+d"(c[1])
                 : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmma1684(double (&c)[4], const double (&a)[2], double b) {
    asm volatile(
        "mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
        : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
        : "d"(a[0]), "d"(a[1]), "d"(b));
}
__device__ __forceinline__ void dmma1688(double (&c)[4], const double (&a)[4], const double (&b)[2]) {
    asm volatile(
        "mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, "
        "{%0,%1,%2,%3};"
        : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
        : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}
__device__ __forceinline__ void dmma16816(double (&c)[4], const double (&a)[8], const double (&b)[4]) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, "
        "{%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
        : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
        : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
          "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}

// MODE 0: m8n8k4, 8 independent accumulator tiles, operands in registers
// MODE 1: m16n8k4   MODE 2: m16n8k8   MODE 3: m16n8k16   (4 independent tiles)
template <int MODE>
__global__ void kreg(double *out, int iters) {
    const double a0 = 1.0 + threadIdx.x * 1e-9, b0 = 1.0 - threadIdx.x * 1e-9;
    double s = 0;
    if (MODE == 0) {
        double c[8][2] = {};
        for (int i = 0; i < iters; ++i) {
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int t = 0; t < 8; ++t) dmma884(c[t], a0 + u, b0 + t);
        }
#pragma unroll
        for (int t = 0; t < 8; ++t) s += c[t][0] + c[t][1];
    } else {
        double c[4][4] = {};
        double a[8], b[4];
#pragma unroll
        for (int j = 0; j < 8; ++j) a[j] = a0 + j;
#pragma unroll
        for (int j = 0; j < 4; ++j) b[j] = b0 + j;
        for (int i = 0; i < iters; ++i) {
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    if (MODE == 1) {
                        double aa[2] = {a[u], a[t]};
                        dmma1684(c[t], aa, b[u]);
                    }
                    if (MODE == 2) {
                        double aa[4] = {a[u], a[t], a[4 + u], a[4 + t]};
                        double bb[2] = {b[u], b[t]};
                        dmma1688(c[t], aa, bb);
                    }
                    if (MODE == 3) dmma16816(c[t], a, b);
                }
        }
#pragma unroll
        for (int t = 0; t < 4; ++t) s += c[t][0] + c[t][1] + c[t][2] + c[t][3];
    }
    if (s == 12345.678) out[0] = s;
}

// Toeplitz-FIR operand pattern with m16n8k8 (A = data windows, row-major 16 x 8,
// from shared memory with a row stride of LDROW doubles; B = tap matrix 8 x 8
// fragment from shared memory, shared by NT row tiles):
//   per k-step: 2 LDS.64 (B) + NT * 4 LDS.64 (A) for NT MMAs of 1024 FMA.
template <int NT, int LDROW>
__global__ void ktoep(double *out, int iters, int ksteps) {
    extern __shared__ __align__(16) double sm[];
    double *xs = sm;                    // data: (16 * NT) rows * LDROW + ksteps * 8
    double *ts = sm + 16 * NT * LDROW + ksteps * 8 + 64;   // taps: ksteps * 64
    const int total = 16 * NT * LDROW + ksteps * 8 + 64 + ksteps * 64;
    for (int i = threadIdx.x; i < total; i += blockDim.x) sm[i] = 1.0 + i * 1e-9;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int g = lane >> 2, q = lane & 3;
    double c[NT][4] = {};
    for (int i = 0; i < iters; ++i) {
        for (int k = 0; k < ksteps; ++k) {
            double b[2];
            b[0] = ts[k * 64 + q * 8 + g];
            b[1] = ts[k * 64 + (q + 4) * 8 + g];
#pragma unroll
            for (int t = 0; t < NT; ++t) {
                const double *px = xs + (t * 16 + g) * LDROW + k * 8 + q;
                double a[4];
                a[0] = px[0];
                a[1] = px[8 * LDROW];
                a[2] = px[4];
                a[3] = px[8 * LDROW + 4];
                dmma1688(c[t], a, b);
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int t = 0; t < NT; ++t) s += c[t][0] + c[t][1] + c[t][2] + c[t][3];
    if (s == 12345.678) out[0] = s;
}

// same with m8n8k4: per k-step 1 LDS.64 (B) + NT LDS.64 (A) for NT MMAs of 256 FMA
template <int NT, int LDROW>
__global__ void ktoep4(double *out, int iters, int ksteps) {
    extern __shared__ __align__(16) double sm[];
    double *xs = sm;
    double *ts = sm + 8 * NT * LDROW + ksteps * 4 + 64;
    const int total = 8 * NT * LDROW + ksteps * 4 + 64 + ksteps * 32;
    for (int i = threadIdx.x; i < total; i += blockDim.x) sm[i] = 1.0 + i * 1e-9;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int g = lane >> 2, q = lane & 3;
    double c[NT][2] = {};
    for (int i = 0; i < iters; ++i) {
#pragma unroll 4
        for (int k = 0; k < ksteps; ++k) {
            const double b = ts[k * 32 + q * 8 + g];
#pragma unroll
            for (int t = 0; t < NT; ++t) {
                const double a = xs[(t * 8 + g) * LDROW + k * 4 + q];
                dmma884(c[t], a, b);
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int t = 0; t < NT; ++t) s += c[t][0] + c[t][1];
    if (s == 12345.678) out[0] = s;
}

static double clk_khz() {
    int clk;
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    return (double)clk;
}

template <typename F>
static float timed(F launch) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    launch(2);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    launch(0);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) printf("  CUDA error: %s\n", cudaGetErrorString(err));
    return ms;
}

template <int MODE>
static void runreg(const char *name, int threads, int nsm) {
    double *out;
    cudaMalloc(&out, 8);
    const int iters = 2000;
    const float ms = timed([&](int it) { kreg<MODE><<<nsm, threads>>>(out, it ? it : iters); });
    const double fma_per_mma = MODE == 0 ? 256 : MODE == 1 ? 512 : MODE == 2 ? 1024 : 2048;
    const double mmas = (double)iters * (MODE == 0 ? 32 : 16) * (threads / 32);
    printf("%-34s threads/SM %4d: %.3f ms  %.1f FMA lanes/clk/SM\n", name, threads, ms,
           mmas * fma_per_mma / (ms * 1e-3 * clk_khz() * 1e3));
    cudaFree(out);
}

template <int NT, int LDROW>
static void runtoep(const char *name, int threads, int nsm) {
    double *out;
    cudaMalloc(&out, 8);
    const int iters = 200, ksteps = 32;
    const size_t smem = (size_t)(16 * NT * LDROW + ksteps * 8 + 64 + ksteps * 64) * 8;
    cudaFuncSetAttribute(ktoep<NT, LDROW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const float ms = timed(
        [&](int it) { ktoep<NT, LDROW><<<nsm, threads, smem>>>(out, it ? it : iters, ksteps); });
    const double mmas = (double)iters * ksteps * NT * (threads / 32);
    printf("%-34s threads/SM %4d: %.3f ms  %.1f FMA lanes/clk/SM (smem %zu B)\n", name, threads, ms,
           mmas * 1024 / (ms * 1e-3 * clk_khz() * 1e3), smem);
    cudaFree(out);
}

template <int NT, int LDROW>
static void runtoep4(const char *name, int threads, int nsm) {
    double *out;
    cudaMalloc(&out, 8);
    const int iters = 200, ksteps = 64;
    const size_t smem = (size_t)(8 * NT * LDROW + ksteps * 4 + 64 + ksteps * 32) * 8;
    cudaFuncSetAttribute(ktoep4<NT, LDROW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const float ms = timed(
        [&](int it) { ktoep4<NT, LDROW><<<nsm, threads, smem>>>(out, it ? it : iters, ksteps); });
    const double mmas = (double)iters * ksteps * NT * (threads / 32);
    printf("%-34s threads/SM %4d: %.3f ms  %.1f FMA lanes/clk/SM (smem %zu B)\n", name, threads, ms,
           mmas * 256 / (ms * 1e-3 * clk_khz() * 1e3), smem);
    cudaFree(out);
}

int main() {
    int nsm;
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
    printf("SMs %d, nominal clock %.0f MHz\n", nsm, clk_khz() / 1e3);
    for (int t : {128, 256, 512, 1024}) {
        runreg<0>("DMMA m8n8k4 reg operands", t, nsm);
        runreg<1>("DMMA m16n8k4 reg operands", t, nsm);
        runreg<2>("DMMA m16n8k8 reg operands", t, nsm);
        runreg<3>("DMMA m16n8k16 reg operands", t, nsm);
    }
    for (int t : {256, 512}) {
        runtoep<1, 25>("Toeplitz m16n8k8 NT=1 ld=25", t, nsm);
        runtoep<2, 25>("Toeplitz m16n8k8 NT=2 ld=25", t, nsm);
        runtoep<4, 25>("Toeplitz m16n8k8 NT=4 ld=25", t, nsm);
        runtoep<4, 26>("Toeplitz m16n8k8 NT=4 ld=26", t, nsm);
        runtoep<4, 28>("Toeplitz m16n8k8 NT=4 ld=28", t, nsm);
        runtoep4<4, 25>("Toeplitz m8n8k4 NT=4 ld=25", t, nsm);
        runtoep4<8, 25>("Toeplitz m8n8k4 NT=8 ld=25", t, nsm);
        runtoep4<8, 28>("Toeplitz m8n8k4 NT=8 ld=28", t, nsm);
    }
    return 0;
}
