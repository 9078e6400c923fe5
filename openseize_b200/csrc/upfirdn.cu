// Polyphase resampling.  Replaces scipy.signal.resample_poly -> upfirdn as
// called per halo'd chunk by nm.polyphase_resample (reference
// core/numerical.py:610,631).  With scipy's bookkeeping folded (h' = h*up,
// half = (K-1)/2, n_pre_pad / n_pre_remove cancel to a centred filter):
//
//     y[j] = sum_k h'[j*down + half - k*up] * x[k]         (global indices)
//
// upfirdn_dec_kernel   up == 1 (decimation, the hot case).  The K-tap sum is
//     split by input phase p = tap index mod down: each phase is a stride-1
//     FIR over the decimated sequence X_p[m] = x[m*down + p], so a thread that
//     owns R consecutive outputs slides a register window along X_p (one LDS
//     per R DFMA).  The input tile is staged phase-major in shared memory
//     with a skew that makes the stride-R window reads conflict free; the 8
//     warps of a CTA split the (phase, tap) work of the same 32*R outputs and
//     reduce through shared memory.
// upfirdn_general_kernel   any up/down; one thread per output sample.
#include <vector>

#include "common.cuh"

namespace osz {

constexpr int UFD_NT = 256;
constexpr int UFD_NW = UFD_NT / 32;

template <int R>
__device__ __forceinline__ int ufd_phys(int m) {
    return m + m / R;
}

// R outputs per lane; tile = 32*R outputs per CTA.
template <int R>
__global__ void __launch_bounds__(UFD_NT, 2)
upfirdn_dec_kernel(const double *__restrict__ x, int64_t ldx, int64_t x_first, int64_t x_len,
                   int64_t out_first, int64_t n_out, int K, int M, int Q /* taps per phase */,
                   int ldm /* smem row length, odd */, int half,
                   const double *__restrict__ gphase /* [M][Q] reversed taps by phase */,
                   double *__restrict__ y, int64_t ldy) {
    constexpr int TO = 32 * R;
    extern __shared__ __align__(16) double smem[];
    double *xs = smem;                       // [M][ldm]
    double *gs = smem + (size_t)M * ldm;     // [M][Q]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t row = blockIdx.y;
    const int64_t o0 = (int64_t)blockIdx.x * TO;              // tile-relative to out_first
    // global index of the first input sample of this tile
    const int64_t in0 = (out_first + o0) * M + half - (K - 1);
    const double *xr = x + row * ldx;

    // stage [TO + Q] decimated samples of every phase, phase-major, skewed;
    // eight independent loads in flight per thread.  Tiles that lie wholly
    // inside the supplied window (all but the recording's edges) skip the
    // per-element bounds arithmetic.
    const int n_in = (TO + Q) * M;
    {
        constexpr int U = 8;
        const int dp = UFD_NT % M, dm = UFD_NT / M;
        int p = tid % M, m = tid / M;
        const int64_t rel0 = in0 - x_first;
        const bool interior = rel0 >= 0 && rel0 + n_in <= x_len;
        const double *src = xr + rel0 + tid;
        for (int e0 = tid; e0 < n_in; e0 += U * UFD_NT, src += U * UFD_NT) {
            double val[U];
            if (interior && e0 + (U - 1) * UFD_NT < n_in) {
#pragma unroll
                for (int u = 0; u < U; ++u) val[u] = ld_stream(src + u * UFD_NT);
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    xs[p * ldm + ufd_phys<R>(m)] = val[u];
                    p += dp;
                    m += dm;
                    if (p >= M) {
                        p -= M;
                        m += 1;
                    }
                }
            } else {
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int e = e0 + u * UFD_NT;
                    const int64_t g = rel0 + e;
                    val[u] = (e < n_in && g >= 0 && g < x_len) ? ld_stream(xr + g) : 0.0;
                }
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    if (e0 + u * UFD_NT < n_in) xs[p * ldm + ufd_phys<R>(m)] = val[u];
                    p += dp;
                    m += dm;
                    if (p >= M) {
                        p -= M;
                        m += 1;
                    }
                }
            }
        }
    }
    for (int i = tid; i < M * Q; i += UFD_NT) gs[i] = gphase[i];
    __syncthreads();

    // this warp's share of the M*Q (phase, tap) slots
    const int slots = M * Q;
    const int s_lo = (int)(((int64_t)slots * warp) / UFD_NW);
    const int s_hi = (int)(((int64_t)slots * (warp + 1)) / UFD_NW);

    double acc[R];
#pragma unroll
    for (int r = 0; r < R; ++r) acc[r] = 0.0;

    // Shared-memory position of window element i of lane l is
    //   l*(R+1) + T(i),  T(i) = i + i/R   (ufd_phys of l*R + i, R a power of two).
    // Along a run of taps the index advances by one per step, so with
    // rem = (first index) % R fixed for the run, T(j+d) - T(j) = d + [rem + d >= R]
    // (+1 more for d >= R): every load is `one of two base pointers + immediate`.
    int s = s_lo;
    while (s < s_hi) {
        const int p = s / Q, q0 = s - p * Q;
        int q1 = Q;
        if (p * Q + q1 > s_hi) q1 = s_hi - p * Q;
        const double *gp = gs + p * Q;
        const unsigned uq = (unsigned)q0;
        const unsigned rem = uq % R;
        const double *pa = xs + p * ldm + lane * (R + 1) + (uq + uq / R);   // T(q0), crossing not passed
        const double *pb = pa + 1;                                          // crossing passed
        bool c[R];
#pragma unroll
        for (int u = 0; u < R; ++u) c[u] = rem + u >= R;
        double w[R];
#pragma unroll
        for (int r = 0; r < R; ++r) w[r] = (c[r] ? pb : pa)[r];
        int q = q0;
        for (; q + R <= q1; q += R) {
#pragma unroll
            for (int u = 0; u < R; ++u) {
                const double g = gp[q + u];
#pragma unroll
                for (int r = 0; r < R; ++r) acc[r] = fma(g, w[(r + u) % R], acc[r]);
                w[u] = (c[u] ? pb : pa)[u + R + 1];
            }
            pa += R + 1;
            pb += R + 1;
        }
        // remainder (< R taps): same rotation, guarded
#pragma unroll
        for (int u = 0; u < R; ++u) {
            if (q + u < q1) {
                const double g = gp[q + u];
#pragma unroll
                for (int r = 0; r < R; ++r) acc[r] = fma(g, w[(r + u) % R], acc[r]);
                w[u] = (c[u] ? pb : pa)[u + R + 1];
            }
        }
        s = p * Q + q1;
    }
    __syncthreads();   // tile no longer needed: reuse it for the cross-warp sum
    double *red = smem;   // [NW][32 lanes][R + 1]: the +1 keeps the stride-R stores conflict free
    constexpr int LDR = 32 * (R + 1);
#pragma unroll
    for (int r = 0; r < R; ++r) red[warp * LDR + lane * (R + 1) + r] = acc[r];
    __syncthreads();
    for (int o = tid; o < TO; o += UFD_NT) {
        const int idx = (o / R) * (R + 1) + (o % R);
        double sum = 0.0;
#pragma unroll
        for (int w8 = 0; w8 < UFD_NW; ++w8) sum += red[w8 * LDR + idx];
        if (o0 + o < n_out) st_stream(y + row * ldy + o0 + o, sum);
    }
}

__global__ void upfirdn_general_kernel(const double *__restrict__ x, int64_t ldx, int64_t x_first,
                                       int64_t x_len, int64_t out_first, int64_t n_out, int K,
                                       int L, int M, int half,
                                       const double *__restrict__ h /* scaled by up */,
                                       double *__restrict__ y, int64_t ldy) {
    const int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= n_out) return;
    const int64_t row = blockIdx.y;
    const int64_t t = (out_first + o) * M + half;
    const int64_t k0 = t / L;                 // newest contributing input sample
    const int phi = (int)(t - k0 * L);
    const double *xr = x + row * ldx;
    double acc = 0.0;
    // scipy's upfirdn accumulates oldest input first: keep the same order
    const int qmax = (K - 1 - phi) / L;
    for (int q = qmax; q >= 0; --q) {
        const int64_t g = k0 - q - x_first;
        if (g >= 0 && g < x_len) acc = fma(ldg(h + phi + q * L), ldg(xr + g), acc);
    }
    y[row * ldy + o] = acc;
}

}  // namespace osz

using namespace osz;

struct osz_upfirdn_plan {
    int K = 0, up = 1, down = 1, half = 0;
    int Q = 0;                     // taps per phase (decimator)
    int R = 0;                     // outputs per lane (0: general kernel)
    int ldm = 0;
    size_t smem = 0;
    double *d_h = nullptr;         // h * up
    double *d_gphase = nullptr;    // [down][Q]
};

template <int R>
static int launch_dec(const osz_upfirdn_plan *p, const double *x, int64_t ldx, int64_t rows,
                      int64_t x_first, int64_t x_len, int64_t out_first, int64_t n_out, double *y,
                      int64_t ldy, cudaStream_t st) {
    OSZ_CUDA(cudaFuncSetAttribute(upfirdn_dec_kernel<R>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)p->smem));
    dim3 grid((unsigned)((n_out + 32 * R - 1) / (32 * R)), (unsigned)rows);
    upfirdn_dec_kernel<R><<<grid, UFD_NT, p->smem, st>>>(x, ldx, x_first, x_len, out_first, n_out,
                                                         p->K, p->down, p->Q, p->ldm, p->half,
                                                         p->d_gphase, y, ldy);
    OSZ_LAUNCHED("upfirdn_dec_kernel");
    return OSZ_OK;
}

static size_t dec_smem(int R, int M, int Q, int *ldm_out) {
    const int TO = 32 * R;
    int ldm = (TO + Q + R) + (TO + Q + R) / R + 1;
    if ((ldm & 1) == 0) ++ldm;
    *ldm_out = ldm;
    size_t tile = ((size_t)M * ldm + (size_t)M * Q) * 8;
    size_t red = (size_t)UFD_NW * 32 * (R + 1) * 8;
    return tile > red ? tile : red;
}

extern "C" {

int osz_upfirdn_plan_create(osz_upfirdn_plan **out, const double *h, int K, int up, int down) {
    if (!out || !h || K < 1 || up < 1 || down < 1)
        return fail(OSZ_ERR_ARG, "osz_upfirdn_plan_create: bad arguments");
    osz_upfirdn_plan *p = new osz_upfirdn_plan();
    p->K = K;
    p->up = up;
    p->down = down;
    p->half = (K - 1) / 2;
    std::vector<double> hs(K);
    for (int i = 0; i < K; ++i) hs[i] = h[i] * (double)up;      // scipy: h = h * up
    bool ok = cudaMalloc(&p->d_h, (size_t)K * 8) == cudaSuccess &&
              cudaMemcpy(p->d_h, hs.data(), (size_t)K * 8, cudaMemcpyHostToDevice) == cudaSuccess;
    if (ok && up == 1 && down >= 2) {
        const int M = down;
        p->Q = (K + M - 1) / M;
        // largest R in {16, 8, 4} whose tile fits ~100 KB (two CTAs per SM)
        for (int R : {16, 8, 4}) {
            int ldm = 0;
            size_t s = dec_smem(R, M, p->Q, &ldm);
            if (s <= 100 * 1024) {
                p->R = R;
                p->ldm = ldm;
                p->smem = s;
                break;
            }
        }
        if (p->R) {
            // gphase[p][q] = g[q*M + p], g[j] = h'[K-1-j]
            std::vector<double> gp((size_t)M * p->Q, 0.0);
            for (int j = 0; j < K; ++j) gp[(size_t)(j % M) * p->Q + j / M] = hs[K - 1 - j];
            ok = cudaMalloc(&p->d_gphase, gp.size() * 8) == cudaSuccess &&
                 cudaMemcpy(p->d_gphase, gp.data(), gp.size() * 8, cudaMemcpyHostToDevice) ==
                     cudaSuccess;
        }
    }
    if (!ok) {
        osz_upfirdn_plan_destroy(p);
        return fail(OSZ_ERR_CUDA, "osz_upfirdn_plan_create: device upload failed");
    }
    *out = p;
    return OSZ_OK;
}

int osz_upfirdn_plan_destroy(osz_upfirdn_plan *p) {
    if (!p) return OSZ_OK;
    cudaFree(p->d_h);
    cudaFree(p->d_gphase);
    delete p;
    return OSZ_OK;
}

int osz_upfirdn_exec_f64(const osz_upfirdn_plan *p, const double *x, int64_t ldx, int64_t rows,
                         int64_t x_first, int64_t x_len, int64_t out_first, int64_t n_out,
                         double *y, int64_t ldy, void *stream) {
    if (!p || !x || !y) return fail(OSZ_ERR_ARG, "osz_upfirdn_exec_f64: null argument");
    if (rows <= 0 || n_out <= 0) return OSZ_OK;
    if (rows > 65535) return fail(OSZ_ERR_UNSUPPORTED, "upfirdn: more than 65535 rows per call");
    cudaStream_t st = as_stream(stream);
    switch (p->R) {
        case 16: return launch_dec<16>(p, x, ldx, rows, x_first, x_len, out_first, n_out, y, ldy, st);
        case 8: return launch_dec<8>(p, x, ldx, rows, x_first, x_len, out_first, n_out, y, ldy, st);
        case 4: return launch_dec<4>(p, x, ldx, rows, x_first, x_len, out_first, n_out, y, ldy, st);
        default: break;
    }
    dim3 grid((unsigned)((n_out + 255) / 256), (unsigned)rows);
    upfirdn_general_kernel<<<grid, 256, 0, st>>>(x, ldx, x_first, x_len, out_first, n_out, p->K,
                                                 p->up, p->down, p->half, p->d_h, y, ldy);
    OSZ_LAUNCHED("upfirdn_general_kernel");
    return OSZ_OK;
}

}  // extern "C"
