// Generic-nfft windowed DFT (any nfft >= 2): the path for segment lengths the
// shared-memory power-of-two kernels do not cover -- e.g. the reference's
// default psd resolution of 0.5 Hz gives nfft = 2 fs = 10000 at 5 kHz
// (spectra/estimators.py:144).  Same contract as spectra.cu.
//
// Two segments of a row are packed as real / imaginary part of one complex
// sequence, transformed by a batched mixed-radix Stockham FFT that runs one
// kernel per pass through global memory (radices 16/8/4/2 and 3/5/7/11/13),
// then untangled.  Lengths with a prime factor above 13 go through Bluestein's
// chirp-z identity on a power-of-two transform.  Throughput is a fraction of
// the shared-memory path (every pass round-trips HBM); correctness and
// coverage are what this path is for.
#include <math.h>

#include <vector>

#include "common.cuh"
#include "fft_core.cuh"

struct osz_spec_plan;
int osz_spec_plan_nfft(const osz_spec_plan *p);
int osz_spec_plan_stride(const osz_spec_plan *p);
int osz_spec_plan_detrend(const osz_spec_plan *p);
double osz_spec_plan_norm(const osz_spec_plan *p);
const double *osz_spec_plan_window(const osz_spec_plan *p);

namespace osz {

enum { GEN_ACCUM = 0, GEN_PGRAM = 1, GEN_STFT = 2 };

// ---- one Stockham pass: radix R, sub-transform length Ns, batch of n-point rows
template <int R>
__global__ void gen_pass_kernel(const double2 *__restrict__ in, double2 *__restrict__ out, int n,
                                int Ns, const double2 *__restrict__ tw, long long total) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int per = n / R;
    const long long b = idx / per;
    const int j = (int)(idx - b * per);
    const int k = j % Ns;
    const double2 *src = in + b * n + j;
    double2 v[R];
#pragma unroll
    for (int r = 0; r < R; ++r) v[r] = src[(long long)r * per];
    const long long tstep = (long long)k * (n / (Ns * R));   // r * tstep < n
#pragma unroll
    for (int r = 1; r < R; ++r) v[r] = cmul(v[r], ldg(tw + r * tstep));
    if (R == 2 || R == 4 || R == 8 || R == 16) {
        bfly<(R == 2 || R == 4 || R == 8 || R == 16) ? R : 2>(v);
    } else {
        // small odd radix: direct DFT with the R-th roots of unity
        double2 o[R];
        const int root = n / R;
#pragma unroll
        for (int q = 0; q < R; ++q) {
            double2 acc = v[0];
#pragma unroll
            for (int r = 1; r < R; ++r) acc = cadd(acc, cmul(v[r], ldg(tw + ((q * r) % R) * root)));
            o[q] = acc;
        }
#pragma unroll
        for (int q = 0; q < R; ++q) v[q] = o[q];
    }
    double2 *dst = out + b * n + (long long)(j - k) * R + k;
#pragma unroll
    for (int r = 0; r < R; ++r) dst[(long long)r * Ns] = v[r];
}

// ---- load: detrend + window two segments into one complex row of length ldz,
//      optionally multiplied by conj(chirp) (Bluestein), zero padded to ldz
__global__ void gen_load_kernel(const double *__restrict__ x, int64_t ldx, int64_t rows,
                                int64_t nseg, int64_t seg0, int nfft, int64_t stride, int detrend,
                                const double *__restrict__ win, const double2 *__restrict__ chirp,
                                double2 *__restrict__ z, int ldz) {
    __shared__ double red[4][32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
    const int64_t pair = blockIdx.x, row = blockIdx.y;
    const int64_t sa = seg0 + 2 * pair;
    const bool has_b = sa + 1 < nseg;
    const double *xa = x + row * ldx + sa * stride;
    const double *xb = xa + stride;
    double s[4] = {0.0, 0.0, 0.0, 0.0};
    const double tbar = 0.5 * (nfft - 1);
    if (detrend != OSZ_DETREND_NONE) {
        for (int i = tid; i < nfft; i += blockDim.x) {
            const double a = xa[i], b = has_b ? xb[i] : 0.0, tc = (double)i - tbar;
            s[0] += a;
            s[1] += b;
            s[2] = fma(tc, a, s[2]);
            s[3] = fma(tc, b, s[3]);
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) s[q] += __shfl_xor_sync(0xffffffffu, s[q], o);
            if (lane == 0) red[q][warp] = s[q];
        }
        __syncthreads();
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            double t = 0.0;
            for (int w = 0; w < nw; ++w) t += red[q][w];
            s[q] = t;
        }
    }
    const double ma = s[0] / nfft, mb = s[1] / nfft;
    double ka = 0.0, kb = 0.0;
    if (detrend == OSZ_DETREND_LINEAR) {
        const double stt = (double)nfft * ((double)nfft * nfft - 1.0) / 12.0;
        ka = s[2] / stt;
        kb = s[3] / stt;
    }
    double2 *zr = z + (pair * rows + row) * (int64_t)ldz;
    for (int i = tid; i < ldz; i += blockDim.x) {
        double2 val = make_double2(0.0, 0.0);
        if (i < nfft) {
            const double tc = (double)i - tbar, w = win[i];
            double a = xa[i], b = has_b ? xb[i] : 0.0;
            if (detrend != OSZ_DETREND_NONE) {
                a -= fma(ka, tc, ma);
                b -= fma(kb, tc, mb);
            }
            val = make_double2(a * w, b * w);
            if (chirp) {
                const double2 c = chirp[i];
                val = cmul(val, make_double2(c.x, -c.y));
            }
        }
        zr[i] = val;
    }
}

// Bluestein middle step: P = A * B, conjugated so that the next forward
// transform yields conj(IFFT)
__global__ void gen_mul_conj_kernel(double2 *__restrict__ a, const double2 *__restrict__ B, int m,
                                    long long total) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const double2 p = cmul(a[idx], B[idx % m]);
    a[idx] = make_double2(p.x, -p.y);
}

// ---- finish: untangle the two segments and emit.  Z[k] is zr[k] for the plain
//      path; for Bluestein Z[k] = conj(chirp[k]) * conj(zr[k]) / M.
__global__ void gen_finish_kernel(const double2 *__restrict__ z, int ldz, int64_t rows,
                                  int64_t nseg, int64_t seg0, int64_t npairs, int nfft,
                                  const double2 *__restrict__ chirp, double inv_m, double norm,
                                  int mode, double *__restrict__ out, int64_t ldp) {
    const int nf = nfft / 2 + 1;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = (long long)npairs * rows * nf;
    if (idx >= total) return;
    const int k = (int)(idx % nf);
    const int64_t row = (idx / nf) % rows;
    const int64_t pair = idx / ((long long)nf * rows);
    const double2 *zr = z + (pair * rows + row) * (int64_t)ldz;
    const int kn = k == 0 ? 0 : nfft - k;
    double2 zk = zr[k], zn = zr[kn];
    if (chirp) {
        const double2 ck = chirp[k], cn = chirp[kn];
        zk = cmul(make_double2(ck.x, -ck.y), make_double2(zk.x * inv_m, -zk.y * inv_m));
        zn = cmul(make_double2(cn.x, -cn.y), make_double2(zn.x * inv_m, -zn.y * inv_m));
    }
    const double2 xa = make_double2(0.5 * (zk.x + zn.x), 0.5 * (zk.y - zn.y));
    const double2 xb = make_double2(0.5 * (zk.y + zn.y), -0.5 * (zk.x - zn.x));
    const int64_t sa = seg0 + 2 * pair;
    const bool has_b = sa + 1 < nseg;
    const bool edge = k == 0 || (nfft % 2 == 0 && k == nfft / 2);
    const double f = edge ? norm : 2.0 * norm;
    if (mode == GEN_STFT) {
        const double amp = sqrt(norm);
        double2 *o = reinterpret_cast<double2 *>(out);
        o[(sa * rows + row) * nf + k] = make_double2(xa.x * amp, xa.y * amp);
        if (has_b) o[((sa + 1) * rows + row) * nf + k] = make_double2(xb.x * amp, xb.y * amp);
    } else if (mode == GEN_PGRAM) {
        out[(sa * rows + row) * nf + k] = f * (xa.x * xa.x + xa.y * xa.y);
        if (has_b) out[((sa + 1) * rows + row) * nf + k] = f * (xb.x * xb.x + xb.y * xb.y);
    } else {
        double p = xa.x * xa.x + xa.y * xa.y;
        if (has_b) p += xb.x * xb.x + xb.y * xb.y;
        atomicAdd(out + row * ldp + k, f * p);
    }
}

struct GenFft {
    int n = 0;
    std::vector<int> radices;
    double2 *d_tw = nullptr;   // exp(-2 pi i m / n), m < n
};

static bool factor(int n, std::vector<int> &radices) {
    const int cand[] = {16, 8, 4, 2, 3, 5, 7, 11, 13};
    for (int c : cand)
        while (n % c == 0) {
            radices.push_back(c);
            n /= c;
        }
    return n == 1;
}

static int make_fft(GenFft &f, int n) {
    f.n = n;
    f.radices.clear();
    if (!factor(n, f.radices)) return fail(OSZ_ERR_UNSUPPORTED, "generic fft: unfactorable length");
    std::vector<double> tw(2 * (size_t)n);
    const long double two_pi = 6.283185307179586476925286766559005768L;
    for (int m = 0; m < n; ++m) {
        tw[2 * m] = (double)cosl(two_pi * m / n);
        tw[2 * m + 1] = (double)(-sinl(two_pi * m / n));
    }
    if (cudaMalloc(&f.d_tw, tw.size() * 8) != cudaSuccess ||
        cudaMemcpy(f.d_tw, tw.data(), tw.size() * 8, cudaMemcpyHostToDevice) != cudaSuccess)
        return fail(OSZ_ERR_CUDA, "generic fft: twiddle upload failed");
    return OSZ_OK;
}

template <int R>
static int launch_pass(const double2 *in, double2 *out, int n, int Ns, const double2 *tw,
                       long long batch, cudaStream_t st) {
    const long long total = batch * (n / R);
    gen_pass_kernel<R><<<(unsigned)((total + 127) / 128), 128, 0, st>>>(in, out, n, Ns, tw, total);
    OSZ_LAUNCHED("gen_pass_kernel");
    return OSZ_OK;
}

// Forward transform of `batch` rows of length f.n; result pointer in *res.
static int run_fft(const GenFft &f, double2 *a, double2 *b, long long batch, cudaStream_t st,
                   double2 **res) {
    int Ns = 1;
    double2 *in = a, *out = b;
    for (int R : f.radices) {
        int rc;
        switch (R) {
            case 16: rc = launch_pass<16>(in, out, f.n, Ns, f.d_tw, batch, st); break;
            case 8: rc = launch_pass<8>(in, out, f.n, Ns, f.d_tw, batch, st); break;
            case 4: rc = launch_pass<4>(in, out, f.n, Ns, f.d_tw, batch, st); break;
            case 2: rc = launch_pass<2>(in, out, f.n, Ns, f.d_tw, batch, st); break;
            case 3: rc = launch_pass<3>(in, out, f.n, Ns, f.d_tw, batch, st); break;
            case 5: rc = launch_pass<5>(in, out, f.n, Ns, f.d_tw, batch, st); break;
            case 7: rc = launch_pass<7>(in, out, f.n, Ns, f.d_tw, batch, st); break;
            case 11: rc = launch_pass<11>(in, out, f.n, Ns, f.d_tw, batch, st); break;
            default: rc = launch_pass<13>(in, out, f.n, Ns, f.d_tw, batch, st); break;
        }
        if (rc != OSZ_OK) return rc;
        Ns *= R;
        double2 *t = in;
        in = out;
        out = t;
    }
    *res = in;
    return OSZ_OK;
}

struct GenState {
    int nfft = 0;
    bool bluestein = false;
    int m = 0;                    // transform length (nfft, or the Bluestein power of two)
    GenFft fft;
    double2 *d_chirp = nullptr;   // exp(+i pi j^2 / nfft), j < nfft
    double2 *d_B = nullptr;       // FFT_m of the wrapped chirp
    double2 *buf[2] = {nullptr, nullptr};
    size_t cap = 0;               // complex elements per buffer
};

}  // namespace osz

using namespace osz;

void osz_generic_destroy(void *state) {
    GenState *g = static_cast<GenState *>(state);
    if (!g) return;
    cudaFree(g->fft.d_tw);
    cudaFree(g->d_chirp);
    cudaFree(g->d_B);
    cudaFree(g->buf[0]);
    cudaFree(g->buf[1]);
    delete g;
}

int osz_generic_create(void **state, int nfft) {
    GenState *g = new GenState();
    g->nfft = nfft;
    std::vector<int> probe;
    g->bluestein = !factor(nfft, probe);
    g->m = nfft;
    if (g->bluestein) {
        g->m = 1;
        while (g->m < 2 * nfft - 1) g->m <<= 1;
    }
    int rc = make_fft(g->fft, g->m);
    if (rc != OSZ_OK) {
        osz_generic_destroy(g);
        return rc;
    }
    if (g->bluestein) {
        // chirp c[j] = exp(+i pi j^2 / n) with j^2 reduced mod 2n exactly
        const long double pi = 3.141592653589793238462643383279502884L;
        std::vector<double> c(2 * (size_t)nfft), wrapped(2 * (size_t)g->m, 0.0);
        for (int j = 0; j < nfft; ++j) {
            const long long q = ((long long)j * j) % (2LL * nfft);
            const long double a = pi * (long double)q / (long double)nfft;
            c[2 * j] = (double)cosl(a);
            c[2 * j + 1] = (double)sinl(a);
            wrapped[2 * j] = c[2 * j];
            wrapped[2 * j + 1] = c[2 * j + 1];
            if (j > 0) {
                wrapped[2 * (size_t)(g->m - j)] = c[2 * j];
                wrapped[2 * (size_t)(g->m - j) + 1] = c[2 * j + 1];
            }
        }
        double2 *tmp = nullptr, *res = nullptr;
        bool ok = cudaMalloc(&g->d_chirp, c.size() * 8) == cudaSuccess &&
                  cudaMemcpy(g->d_chirp, c.data(), c.size() * 8, cudaMemcpyHostToDevice) ==
                      cudaSuccess &&
                  cudaMalloc(&g->d_B, wrapped.size() * 8) == cudaSuccess &&
                  cudaMalloc(&tmp, wrapped.size() * 8) == cudaSuccess &&
                  cudaMemcpy(tmp, wrapped.data(), wrapped.size() * 8, cudaMemcpyHostToDevice) ==
                      cudaSuccess;
        if (ok) {
            double2 *scratch = g->d_B;
            rc = run_fft(g->fft, tmp, scratch, 1, 0, &res);
            ok = rc == OSZ_OK && cudaDeviceSynchronize() == cudaSuccess;
            if (ok && res != g->d_B)
                ok = cudaMemcpy(g->d_B, res, wrapped.size() * 8, cudaMemcpyDeviceToDevice) ==
                     cudaSuccess;
        }
        cudaFree(tmp);
        if (!ok) {
            osz_generic_destroy(g);
            return fail(OSZ_ERR_CUDA, "generic spectra: Bluestein setup failed");
        }
    }
    *state = g;
    return OSZ_OK;
}

int osz_generic_exec(void *state, const osz_spec_plan *p, int mode, const double *x, int64_t ldx,
                     int64_t rows, int64_t nseg, double *out, int64_t ldp, cudaStream_t st) {
    GenState *g = static_cast<GenState *>(state);
    if (!g) return fail(OSZ_ERR_ARG, "generic spectra: no state");
    const int nfft = g->nfft, m = g->m;
    const int64_t stride = osz_spec_plan_stride(p);
    const int64_t npairs_all = (nseg + 1) / 2;
    // workspace: two ping-pong buffers, <= 256 MiB each, sized for whole pairs
    const size_t per_pair = (size_t)rows * m;
    size_t want = per_pair * (size_t)npairs_all;
    const size_t cap_max = (size_t)(256u << 20) / 16;
    if (want > cap_max) want = (cap_max / per_pair) * per_pair;
    if (want < per_pair) want = per_pair;
    if (g->cap < want) {
        cudaFree(g->buf[0]);
        cudaFree(g->buf[1]);
        g->buf[0] = g->buf[1] = nullptr;
        g->cap = 0;
        if (cudaMalloc(&g->buf[0], want * 16) != cudaSuccess ||
            cudaMalloc(&g->buf[1], want * 16) != cudaSuccess)
            return fail(OSZ_ERR_ALLOC, "generic spectra: workspace allocation failed");
        g->cap = want;
    }
    const int64_t pairs_per_batch = (int64_t)(g->cap / per_pair);
    for (int64_t p0 = 0; p0 < npairs_all; p0 += pairs_per_batch) {
        const int64_t np = npairs_all - p0 < pairs_per_batch ? npairs_all - p0 : pairs_per_batch;
        if (np > 2147483647LL || rows > 65535)
            return fail(OSZ_ERR_UNSUPPORTED, "generic spectra: grid too large");
        dim3 grid((unsigned)np, (unsigned)rows);
        gen_load_kernel<<<grid, 256, 0, st>>>(x, ldx, rows, nseg, 2 * p0, nfft, stride,
                                              osz_spec_plan_detrend(p), osz_spec_plan_window(p),
                                              g->bluestein ? g->d_chirp : nullptr, g->buf[0], m);
        OSZ_LAUNCHED("gen_load_kernel");
        const long long batch = (long long)np * rows;
        double2 *res = nullptr;
        int rc = run_fft(g->fft, g->buf[0], g->buf[1], batch, st, &res);
        if (rc != OSZ_OK) return rc;
        if (g->bluestein) {
            const long long total = batch * m;
            gen_mul_conj_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(res, g->d_B, m,
                                                                                 total);
            OSZ_LAUNCHED("gen_mul_conj_kernel");
            double2 *other = res == g->buf[0] ? g->buf[1] : g->buf[0];
            rc = run_fft(g->fft, res, other, batch, st, &res);
            if (rc != OSZ_OK) return rc;
        }
        const long long total = (long long)np * rows * (nfft / 2 + 1);
        gen_finish_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(
            res, m, rows, nseg, 2 * p0, np, nfft, g->bluestein ? g->d_chirp : nullptr,
            1.0 / (double)m, osz_spec_plan_norm(p), mode, out, ldp);
        OSZ_LAUNCHED("gen_finish_kernel");
    }
    return OSZ_OK;
}
