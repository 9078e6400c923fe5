// Shared-memory mixed-radix windowed DFT for even segment lengths whose half
// factors into 2, 3 and 5 -- the lengths the reference's DEFAULT resolution
// produces (psd(..., resolution=0.5) gives nfft = 2 fs: 1000, 2000, 2400,
// 10000, ..., spectra/estimators.py:144).  Same contract as spectra.cu; the
// global-memory path of spectra_generic.cu remains for everything else (odd
// lengths, other prime factors, segments that do not fit in shared memory).
//
// One real segment of nfft samples is transformed as a complex sequence of
// n2 = nfft/2 points z[j] = x[2j] + i x[2j+1] -- which is exactly how the
// samples lie in memory -- by a Stockham FFT that ping-pongs between two
// shared-memory buffers (radices 16/8/4/2/5/3, one pass per radix), then
// untangled:  X[k] = (Z[k] + conj Z[n2-k])/2 - i/2 W_nfft^k (Z[k] - conj Z[n2-k]).
// Each pass reads its base twiddle exp(-2 pi i k / (Ns R)) from a per-pass table
// in shared memory indexed by k (consecutive lanes, consecutive entries: no bank
// conflicts -- a single table indexed by k * n2/(Ns R) collided 4-8 ways) and
// forms the higher powers by repeated multiplication; the untangling twiddles
// come from a two-level table, exp(-2 pi i k / nfft) = A[k >> 6] * B[k & 63].  A CTA walks a run of consecutive segments of one
// row; Welch accumulates the one-sided periodograms in shared memory and adds
// them to psd_sum once at the end.
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

enum { MIX_ACCUM = 0, MIX_PGRAM = 1, MIX_STFT = 2 };
// threads per CTA: 256, or 512 for segments so long that only one CTA fits an SM
constexpr int MIX_MAXPASS = 16;

struct MixParams {
    int n, n2, npass, tw_len, ua_len;
    int radix[MIX_MAXPASS];
    int tw_off[MIX_MAXPASS];     // pass p: base twiddles exp(-2 pi i k / (Ns R)), k < Ns, at tw_off[p]
};

// exp(-2 pi i m / period) from the two-level table [A: ceil(period/64)][B: 64]
__device__ __forceinline__ double2 mix_tw(const double2 *A, const double2 *B, int m) {
    return cmul(A[m >> 6], B[m & 63]);
}

template <int R>
__device__ __forceinline__ void mix_bfly(double2 *v) {
    if constexpr (R == 2 || R == 4 || R == 8 || R == 16) {
        bfly<R>(v);
    } else if constexpr (R == 3) {
        const double c = 0.86602540378443864676;   // sin(2 pi / 3)
        const double2 t = cadd(v[1], v[2]);
        const double2 m = make_double2(fma(-0.5, t.x, v[0].x), fma(-0.5, t.y, v[0].y));
        const double2 d = csub(v[1], v[2]);
        const double2 s = make_double2(c * d.y, -c * d.x);        // -i c d
        v[0] = cadd(v[0], t);
        v[1] = cadd(m, s);
        v[2] = csub(m, s);
    } else {
        static_assert(R == 5, "radix");
        const double c1 = 0.30901699437494742410, c2 = -0.80901699437494742410;   // cos 2pi/5, 4pi/5
        const double s1 = 0.95105651629515357212, s2 = 0.58778525229247312917;    // sin 2pi/5, 4pi/5
        const double2 a1 = cadd(v[1], v[4]), a2 = cadd(v[2], v[3]);
        const double2 b1 = csub(v[1], v[4]), b2 = csub(v[2], v[3]);
        const double2 p1 = make_double2(fma(c1, a1.x, fma(c2, a2.x, v[0].x)),
                                        fma(c1, a1.y, fma(c2, a2.y, v[0].y)));
        const double2 p2 = make_double2(fma(c2, a1.x, fma(c1, a2.x, v[0].x)),
                                        fma(c2, a1.y, fma(c1, a2.y, v[0].y)));
        const double2 q1 = make_double2(fma(s1, b1.x, s2 * b2.x), fma(s1, b1.y, s2 * b2.y));
        const double2 q2 = make_double2(fma(s2, b1.x, -s1 * b2.x), fma(s2, b1.y, -s1 * b2.y));
        v[0] = cadd(v[0], cadd(a1, a2));
        v[1] = make_double2(p1.x + q1.y, p1.y - q1.x);            // p1 - i q1
        v[4] = make_double2(p1.x - q1.y, p1.y + q1.x);
        v[2] = make_double2(p2.x + q2.y, p2.y - q2.x);
        v[3] = make_double2(p2.x - q2.y, p2.y + q2.x);
    }
}

// One Stockham pass over the CTA's n2 points: radix R, sub-transform length Ns.
template <int R, int MIX_NT>
__device__ __forceinline__ void mix_pass(const double2 *__restrict__ src, double2 *__restrict__ dst,
                                         int n2, int Ns, const double2 *TW, int tid) {
    const int per = n2 / R;
    for (int j = tid; j < per; j += MIX_NT) {
        const int k = j % Ns;
        double2 v[R];
#pragma unroll
        for (int r = 0; r < R; ++r) v[r] = src[j + r * per];
        if (Ns > 1) {
            const double2 w1 = TW[k];
            double2 w = w1;
#pragma unroll
            for (int r = 1; r < R; ++r) {
                v[r] = cmul(v[r], w);
                if (r + 1 < R) w = cmul(w, w1);
            }
        }
        mix_bfly<R>(v);
        double2 *d = dst + (j - k) * R + k;
#pragma unroll
        for (int r = 0; r < R; ++r) d[r * Ns] = v[r];
    }
}

template <int MODE, int MIX_NT>
__global__ void __launch_bounds__(MIX_NT, MIX_NT == 256 ? 3 : 1)
spec_mixed_kernel(const __grid_constant__ MixParams prm, const double *__restrict__ x, int64_t ldx,
                  int64_t rows, int64_t nseg, int64_t stride, int detrend,
                  const double *__restrict__ win, const double2 *__restrict__ tables, double norm,
                  double *__restrict__ out, int64_t ldp, int64_t segs_per_cta) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int n = prm.n, n2 = prm.n2;
    double2 *bufA = reinterpret_cast<double2 *>(smem_raw);
    double2 *bufB = bufA + n2;
    double2 *TW = bufB + n2;                 // per-pass base twiddles
    double2 *UA = TW + prm.tw_len;           // exp(-2 pi i 64 h / n)
    double2 *UB = UA + prm.ua_len;           // exp(-2 pi i l / n)
    double *accs = reinterpret_cast<double *>(UB + 64);   // n2 + 1 (Welch only)
    __shared__ double red[4 * 32];

    const int tid = threadIdx.x;
    const int64_t row = blockIdx.y;
    const int64_t s0 = (int64_t)blockIdx.x * segs_per_cta;
    int64_t s1 = s0 + segs_per_cta;
    if (s1 > nseg) s1 = nseg;
    if (s0 >= s1) return;
    const int ntab = prm.tw_len + prm.ua_len + 64;
    for (int i = tid; i < ntab; i += MIX_NT) TW[i] = ldg(tables + i);
    if (MODE == MIX_ACCUM)
        for (int i = tid; i <= n2; i += MIX_NT) accs[i] = 0.0;
    const double *xr = x + row * ldx;
    const double amp = sqrt(norm);
    const int64_t nf = n2 + 1;

    for (int64_t s = s0; s < s1; ++s) {
        const double *xs = xr + s * stride;
        double *raw = reinterpret_cast<double *>(bufA);      // z[j] = (x[2j], x[2j+1]) in place
        __syncthreads();                                      // previous segment fully consumed
        for (int i = tid; i < n; i += MIX_NT) raw[i] = ldg(xs + i);
        __syncthreads();
        double mean = 0.0, slope = 0.0;
        const double tbar = 0.5 * (double)(n - 1);
        if (detrend != OSZ_DETREND_NONE) {
            double sums[2] = {0.0, 0.0};
            for (int i = tid; i < n; i += MIX_NT) {
                const double v = raw[i];
                sums[0] += v;
                sums[1] = fma((double)i - tbar, v, sums[1]);
            }
            // block reduction (256 threads): warp shuffles, then 8 partials
#pragma unroll
            for (int q = 0; q < 2; ++q) {
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) sums[q] += __shfl_xor_sync(0xffffffffu, sums[q], o);
            }
            if ((tid & 31) == 0) {
                red[tid >> 5] = sums[0];
                red[32 + (tid >> 5)] = sums[1];
            }
            __syncthreads();
            double t0 = 0.0, t1 = 0.0;
#pragma unroll
            for (int w = 0; w < MIX_NT / 32; ++w) {
                t0 += red[w];
                t1 += red[32 + w];
            }
            mean = t0 / (double)n;
            if (detrend == OSZ_DETREND_LINEAR) {
                const double dn = (double)n;
                slope = t1 / (dn * (dn * dn - 1.0) / 12.0);
            }
        }
        for (int i = tid; i < n; i += MIX_NT) {
            double v = raw[i];
            if (detrend != OSZ_DETREND_NONE) v -= fma(slope, (double)i - tbar, mean);
            raw[i] = v * ldg(win + i);
        }
        __syncthreads();
        // ---- passes
        const double2 *src = bufA;
        double2 *dst = bufB;
        int Ns = 1;
        for (int pi = 0; pi < prm.npass; ++pi) {
            const int R = prm.radix[pi];
            switch (R) {
                case 16: mix_pass<16, MIX_NT>(src, dst, n2, Ns, TW + prm.tw_off[pi], tid); break;
                case 8: mix_pass<8, MIX_NT>(src, dst, n2, Ns, TW + prm.tw_off[pi], tid); break;
                case 4: mix_pass<4, MIX_NT>(src, dst, n2, Ns, TW + prm.tw_off[pi], tid); break;
                case 2: mix_pass<2, MIX_NT>(src, dst, n2, Ns, TW + prm.tw_off[pi], tid); break;
                case 5: mix_pass<5, MIX_NT>(src, dst, n2, Ns, TW + prm.tw_off[pi], tid); break;
                default: mix_pass<3, MIX_NT>(src, dst, n2, Ns, TW + prm.tw_off[pi], tid); break;
            }
            Ns *= R;
            __syncthreads();
            const double2 *t = src;
            src = dst;
            dst = const_cast<double2 *>(t);
        }
        // ---- untangle the half-size transform into the nfft/2 + 1 bins and emit
        for (int k = tid; k <= n2; k += MIX_NT) {
            const double2 zk = src[k == n2 ? 0 : k];
            const double2 zc = src[k == 0 ? 0 : n2 - k];
            const double2 e = make_double2(0.5 * (zk.x + zc.x), 0.5 * (zk.y - zc.y));   // even part
            const double2 d = make_double2(0.5 * (zk.x - zc.x), 0.5 * (zk.y + zc.y));   // (Zk - conj Zc)/2
            const double2 w = mix_tw(UA, UB, k);
            const double2 wd = cmul(w, d);
            const double2 X = make_double2(e.x + wd.y, e.y - wd.x);                     // e - i w d
            if (MODE == MIX_STFT) {
                double2 *o = reinterpret_cast<double2 *>(out);
                o[(s * rows + row) * nf + k] = make_double2(X.x * amp, X.y * amp);
            } else {
                const double f = (k == 0 || k == n2) ? norm : 2.0 * norm;
                const double pw = f * fma(X.x, X.x, X.y * X.y);
                if (MODE == MIX_PGRAM)
                    out[(s * rows + row) * nf + k] = pw;
                else
                    accs[k] += pw;
            }
        }
    }
    if (MODE == MIX_ACCUM) {
        __syncthreads();
        double *o = out + row * ldp;
        for (int i = tid; i <= n2; i += MIX_NT) atomicAdd(o + i, accs[i]);
    }
}

struct MixState {
    MixParams prm;
    double2 *d_tables = nullptr;
    size_t smem_seg = 0, smem_acc = 0;     // dynamic shared memory without / with accumulators
};

}  // namespace osz

using namespace osz;

void osz_mixed_destroy(void *state) {
    MixState *m = static_cast<MixState *>(state);
    if (!m) return;
    cudaFree(m->d_tables);
    delete m;
}

// Returns OSZ_OK with *state == nullptr when nfft is not eligible.
int osz_mixed_create(void **state, int nfft) {
    *state = nullptr;
    if (nfft < 16 || (nfft & 1)) return OSZ_OK;
    int n2 = nfft / 2, rest = n2;
    MixState *m = new MixState();
    m->prm.n = nfft;
    m->prm.n2 = n2;
    m->prm.npass = 0;
    // odd radices first: while the sub-transform length Ns is still short a pass scatters
    // with stride R (in 16-byte elements), which is bank-conflict free only for odd R; by the
    // time the power-of-two radices run, Ns >= 32 and consecutive lanes write consecutive slots
    const int cand[] = {5, 3, 16, 8, 4, 2};
    for (int c : cand)
        while (rest % c == 0 && m->prm.npass < MIX_MAXPASS) {
            m->prm.radix[m->prm.npass++] = c;
            rest /= c;
        }
    m->prm.ua_len = (n2 + 1 + 63) / 64;
    {
        int Ns = 1, off = 0;
        for (int pi = 0; pi < m->prm.npass; ++pi) {
            m->prm.tw_off[pi] = off;
            off += Ns;
            Ns *= m->prm.radix[pi];
        }
        m->prm.tw_len = off;
    }
    const size_t ntab = (size_t)m->prm.tw_len + m->prm.ua_len + 64;
    m->smem_seg = ((size_t)2 * n2 + ntab) * 16;
    m->smem_acc = m->smem_seg + ((size_t)n2 + 2) * 8;
    if (rest != 1 || m->smem_seg > 220 * 1024) {
        delete m;
        return OSZ_OK;
    }
    std::vector<double> t(2 * ntab);
    const long double two_pi = 6.283185307179586476925286766559005768L;
    size_t at = 0;
    auto emit = [&](int count, long double step, long double period) {
        for (int i = 0; i < count; ++i) {
            const long double a = -two_pi * step * i / period;
            t[at++] = (double)cosl(a);
            t[at++] = (double)sinl(a);
        }
    };
    {
        int Ns = 1;
        for (int pi = 0; pi < m->prm.npass; ++pi) {
            emit(Ns, 1.0L, (long double)Ns * m->prm.radix[pi]);
            Ns *= m->prm.radix[pi];
        }
    }
    emit(m->prm.ua_len, 64.0L, (long double)nfft);
    emit(64, 1.0L, (long double)nfft);
    if (cudaMalloc(&m->d_tables, t.size() * 8) != cudaSuccess ||
        cudaMemcpy(m->d_tables, t.data(), t.size() * 8, cudaMemcpyHostToDevice) != cudaSuccess) {
        osz_mixed_destroy(m);
        return fail(OSZ_ERR_CUDA, "mixed-radix spectra: table upload failed");
    }
    *state = m;
    return OSZ_OK;
}

// Whether this state can run `mode` (the Welch accumulators need more shared memory).
int osz_mixed_can(void *state, int mode) {
    MixState *m = static_cast<MixState *>(state);
    if (!m) return 0;
    return (mode == MIX_ACCUM ? m->smem_acc : m->smem_seg) <= 220 * 1024;
}

template <int MODE, int NT>
static int launch_mixed(const MixState *m, const osz_spec_plan *p, const double *x, int64_t ldx,
                        int64_t rows, int64_t nseg, double *out, int64_t ldp, cudaStream_t st) {
    const size_t smem = MODE == MIX_ACCUM ? m->smem_acc : m->smem_seg;
    OSZ_CUDA(cudaFuncSetAttribute(spec_mixed_kernel<MODE, NT>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    // CTAs resident per SM by shared memory; ~4 waves of work, at least 4 segments per
    // CTA for Welch so the accumulator flush amortises
    int64_t per_sm = (int64_t)(220 * 1024 / smem);
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 8) per_sm = 8;
    const int64_t target = (int64_t)sm_count() * per_sm * 4;
    int64_t per_row = (target + rows - 1) / rows;
    if (per_row < 1) per_row = 1;
    int64_t spc = (nseg + per_row - 1) / per_row;
    if (MODE == MIX_ACCUM && spc < 4) spc = 4;
    if (spc < 1) spc = 1;
    const int64_t gx = (nseg + spc - 1) / spc;
    if (rows > 65535) return fail(OSZ_ERR_UNSUPPORTED, "spectra: more than 65535 rows per call");
    dim3 grid((unsigned)gx, (unsigned)rows);
    spec_mixed_kernel<MODE, NT><<<grid, NT, smem, st>>>(
        m->prm, x, ldx, rows, nseg, osz_spec_plan_stride(p), osz_spec_plan_detrend(p),
        osz_spec_plan_window(p), m->d_tables, osz_spec_plan_norm(p), out, ldp, spc);
    OSZ_LAUNCHED("spec_mixed_kernel");
    return OSZ_OK;
}

template <int MODE>
static int launch_mixed_nt(const MixState *m, const osz_spec_plan *p, const double *x, int64_t ldx,
                           int64_t rows, int64_t nseg, double *out, int64_t ldp, cudaStream_t st) {
    const size_t smem = MODE == MIX_ACCUM ? m->smem_acc : m->smem_seg;
    if (smem > 110 * 1024) return launch_mixed<MODE, 512>(m, p, x, ldx, rows, nseg, out, ldp, st);
    return launch_mixed<MODE, 256>(m, p, x, ldx, rows, nseg, out, ldp, st);
}

int osz_mixed_exec(void *state, const osz_spec_plan *p, int mode, const double *x, int64_t ldx,
                   int64_t rows, int64_t nseg, double *out, int64_t ldp, cudaStream_t st) {
    const MixState *m = static_cast<const MixState *>(state);
    if (mode == MIX_ACCUM) return launch_mixed_nt<MIX_ACCUM>(m, p, x, ldx, rows, nseg, out, ldp, st);
    if (mode == MIX_PGRAM) return launch_mixed_nt<MIX_PGRAM>(m, p, x, ldx, rows, nseg, out, ldp, st);
    return launch_mixed_nt<MIX_STFT>(m, p, x, ldx, rows, nseg, out, ldp, st);
}
