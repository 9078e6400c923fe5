// Cascaded biquads (DF2T) as a time-parallel scan.  Replaces the per-chunk
// scipy.signal.sosfilt calls of nm.sosfilt / nm.sosfiltfilt (reference
// core/numerical.py:334,399,402,410) and scipy.signal.lfilter for second
// order (b, a) (:445,508,511,519).
//
// One section in state-space form (derived from the DF2T recurrence
//   y = b0 x + z0 ; z0' = b1 x - a1 y + z1 ; z1' = b2 x - a2 y ):
//   s' = A s + B x,  y = b0 x + s[0],  A = [[-a1, 1], [-a2, 0]].
//
// One CTA owns one row and walks it in blocks of 8192 samples; the 256
// threads each own 32 consecutive samples in registers.  Per section:
//   1. every thread runs the recurrence over its 32 samples from a ZERO state
//      (the thread that holds the first real sample starts from the carried
//      state instead) -> local outputs + local final state f;
//   2. the true state at every thread boundary is the scan of
//      e_p = M e_{p-1} + f_p, M = A^32: Kogge-Stone over the warp with the
//      precomputed powers M^(2^k), then a serial combine over the 8 warps;
//   3. every thread adds the zero-input response of its entering state to its
//      32 outputs: y_i += g0[i]*S0 + g1[i]*S1, (g0, g1)[i] = row 0 of A^i.
// The corrected outputs are the next section's inputs.  The carried state
// after the last sample is written back, exactly as the reference carries `z`
// between chunks.  All per-filter tables ride in the kernel parameter block
// (constant bank), so concurrent plans never share mutable device state.
#include <math.h>
#include <stdlib.h>

#include <map>
#include <mutex>
#include <type_traits>
#include <vector>

#include "common.cuh"
#include "sos_core.cuh"
#include "sos_tile.cuh"

namespace osz {

// T = samples per thread: 32 for long cascades (less scan work per sample), 16
// for one or two sections, where the kernel is bound by memory latency and the
// smaller footprint (32 registers of samples, 35 KB of shared memory) lets four
// CTAs share an SM instead of two.
// PF: the loads of block blk + 1 are issued into registers before block blk is scanned, so
// a CTA always has loads in flight.  T = 32 with PF takes 64 more registers (one CTA per
// SM): the build for launches with no more CTAs than SMs (few rows), where a lone CTA per
// SM otherwise leaves HBM idle through its scan phase.
// TIO: the samples' type in memory (double, or float for the float32 I/O mode: the
// recurrence, the scan and the carried state stay float64 -- SURVEY.md 8d).
template <bool WRITE, int T, bool PF = false, typename TIO = double>
__global__ void __launch_bounds__(SOS_NT, (T == 32 ? (PF ? 1 : 2) : (PF ? 2 : 4)))
sos_scan_kernel(const __grid_constant__ SosParams prm, const TIO *__restrict__ x, int64_t ldx,
                int64_t n_total, int reverse, const double *__restrict__ state_in,
                double *__restrict__ state, TIO *__restrict__ y, int64_t ldy,
                const double *__restrict__ lanepow /* [sec][32][4]: A^(T*(lane+1)) */,
                int64_t span_len, int64_t settle,
                const double *__restrict__ span_in /* exact split, pass 2: entering states */,
                double *__restrict__ span_out /* exact split, pass 1: final states */) {
    constexpr int BLK = SOS_NT * T;        // samples per CTA iteration
    constexpr int LD = T + 1;              // padded shared-memory row
    constexpr int LOGT = T == 32 ? 5 : 4;
    static_assert(T == 32 || T == 16, "T");
    extern __shared__ __align__(16) double buf[];   // SOS_NT * LD
    __shared__ double wtot[2][SOS_NT / 32][2];   // double-buffered by section parity
    __shared__ double carry[SOS_MAXSEC][2];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t row = blockIdx.x;
    const int nsec = prm.nsec;
    double *st = state + row * nsec * 2;

    // Time split: span k of a row covers logical samples [a, b).  Spans after
    // the first warm the filter up from a zero state over the `settle` samples
    // before `a` (a stable cascade has forgotten its starting state by then, to
    // ~1e-24 relative) and store nothing for them.  Span 0 starts from the
    // carried state; the last span writes the state that is carried on.
    // Exact split (span_in / span_out): no warm-up; pass 1 runs every span from
    // rest (span 0 from the carried state) and publishes its final state, the
    // host-launched combine kernel turns those into entering states, pass 2
    // starts every span from its true entering state.
    const bool exact = span_in != nullptr || span_out != nullptr;
    const int64_t span = blockIdx.y;
    const int64_t a = span * span_len;
    int64_t b = a + span_len;
    if (b > n_total) b = n_total;
    const int64_t w0 = (span == 0 || exact) ? a : a - settle;   // first logical sample processed
    const int64_t n = b - w0;                           // samples this CTA runs through
    const int64_t keep = a - w0;                        // local index of the first stored sample
    // logical sample s (local) <-> global index: forward w0 + s, reverse n_total-1-(w0+s)
    const TIO *xr = x + row * ldx + (reverse ? n_total - 1 - w0 : w0);
    TIO *yr = WRITE ? y + row * ldy + (reverse ? n_total - 1 - w0 : w0) : nullptr;

    if (tid < nsec * 2) {
        double c0 = 0.0;
        if (span == 0)
            c0 = state_in[row * nsec * 2 + tid];
        else if (span_in)
            c0 = span_in[(row * gridDim.y + span) * nsec * 2 + tid];
        carry[tid >> 1][tid & 1] = c0;
    }

    const int64_t nblk = (n + BLK - 1) / BLK;
    const int64_t first_len = n - (nblk - 1) * BLK;

    double nxt[PF ? T : 1];
    auto prefetch = [&](int64_t blk) {          // a full block (blk >= 1)
        const int64_t p0 = first_len + (blk - 1) * BLK;
        const TIO *src = reverse ? xr - p0 - tid : xr + p0 + tid;
#pragma unroll
        for (int it = 0; it < (PF ? T : 1); ++it)
            nxt[it] = ld_stream(reverse ? src - it * SOS_NT : src + it * SOS_NT);
    };
    for (int64_t blk = 0; blk < nblk; ++blk) {
        // The first block is the short one and sits at the END of the BLK
        // slots, behind `off` virtual zero samples that keep a zero state.
        const int off = blk == 0 ? (int)(BLK - first_len) : 0;
        const int64_t pos0 = blk == 0 ? 0 : first_len + (blk - 1) * BLK;
        __syncthreads();   // carry[] visible / buf free
        if (blk != 0 && PF) {
#pragma unroll
            for (int it = 0; it < (PF ? T : 1); ++it) {
                const int e = tid + it * SOS_NT;
                buf[(e >> LOGT) * LD + (e & (T - 1))] = nxt[it];
            }
        } else if (blk != 0) {
            // full block: T independent coalesced loads in flight per thread
            const TIO *src = reverse ? xr - pos0 - tid : xr + pos0 + tid;
            double tmp[T];
#pragma unroll
            for (int it = 0; it < T; ++it)
                tmp[it] = ld_stream(reverse ? src - it * SOS_NT : src + it * SOS_NT);
#pragma unroll
            for (int it = 0; it < T; ++it) {
                const int e = tid + it * SOS_NT;
                buf[(e >> LOGT) * LD + (e & (T - 1))] = tmp[it];
            }
        } else {
#pragma unroll 8
            for (int e = tid; e < BLK; e += SOS_NT) {
                double val = 0.0;
                if (e >= off) {
                    const int64_t s = pos0 + (e - off);
                    val = (double)ld_stream(reverse ? xr - s : xr + s);
                }
                buf[(e >> LOGT) * LD + (e & (T - 1))] = val;
            }
        }
        if (PF && blk + 1 < nblk) prefetch(blk + 1);
        __syncthreads();
        double v[T];
#pragma unroll
        for (int i = 0; i < T; ++i) v[i] = buf[tid * LD + i];

        sos_scan_block<T, 0>(prm, v, blk != 0, off, carry, wtot, lanepow, tid, lane, warp);
        if (WRITE && pos0 + (BLK - off) > keep) {       // block holds samples to store
#pragma unroll
            for (int i = 0; i < T; ++i) buf[tid * LD + i] = v[i];
            __syncthreads();
            if (blk != 0 && pos0 >= keep) {
                TIO *dst = reverse ? yr - pos0 - tid : yr + pos0 + tid;
#pragma unroll
                for (int it = 0; it < T; ++it) {
                    const int e = tid + it * SOS_NT;
                    st_stream(reverse ? dst - it * SOS_NT : dst + it * SOS_NT,
                              (TIO)buf[(e >> LOGT) * LD + (e & (T - 1))]);
                }
            } else {
#pragma unroll 4
                for (int e = tid; e < BLK; e += SOS_NT) {
                    const int64_t s = pos0 + (e - off);
                    if (e >= off && s >= keep)
                        st_stream(reverse ? yr - s : yr + s,
                                  (TIO)buf[(e >> LOGT) * LD + (e & (T - 1))]);
                }
            }
        }
    }
    __syncthreads();
    if (tid < nsec * 2) {
        if (span_out)
            span_out[(row * gridDim.y + span) * nsec * 2 + tid] = carry[tid >> 1][tid & 1];
        else if (span == gridDim.y - 1)
            st[tid] = carry[tid >> 1][tid & 1];
    }
}

// Exact split, between the passes: entering state of span k+1 = Phi e_k + f_k,
// Phi = (zero-input transition of the whole cascade)^(span length); span 0 ran
// from the true carried state, so f_0 is already the state entering span 1.
__global__ void sos_combine_kernel(const double *__restrict__ phi /* ns2 x ns2 */,
                                   const double *__restrict__ f, double *__restrict__ e,
                                   int64_t rows, int nspan, int ns2) {
    const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= rows) return;
    double cur[2 * SOS_MAXSEC], nxt[2 * SOS_MAXSEC];
    const double *fr = f + row * nspan * ns2;
    double *er = e + row * nspan * ns2;
    for (int i = 0; i < ns2; ++i) cur[i] = fr[i];
    for (int k = 1; k < nspan; ++k) {
        for (int i = 0; i < ns2; ++i) er[k * ns2 + i] = cur[i];
        if (k + 1 == nspan) break;
        for (int i = 0; i < ns2; ++i) {
            double acc = fr[k * ns2 + i];
            for (int j = 0; j < ns2; ++j) acc = fma(phi[i * ns2 + j], cur[j], acc);
            nxt[i] = acc;
        }
        for (int i = 0; i < ns2; ++i) cur[i] = nxt[i];
    }
}

// Entering state of every time span but the first, WITHOUT running the recurrence: the
// state a cascade holds after a run of samples is a linear functional of them,
//   s = sum_d W[d] * x[last - d],   W[d] = T^d b   (T: one-step zero-input transition,
//   b: the state one unit sample leaves behind),
// and W decays like the filter's impulse response, so the `settle` samples before a
// span fix its entering state to ~1e-18 -- the same truncation as re-filtering them
// from rest (the warm-up split), but 2 nsec FMAs per sample, fully parallel, instead of
// the scan.  One CTA per (span >= 1, row); W: [settle][ns2] doubles.
template <int NS2>
__global__ void __launch_bounds__(256)
sos_entering_kernel(const double *__restrict__ W, int64_t settle, const double *__restrict__ x,
                    int64_t ldx, int64_t n_total, int reverse, int64_t span_len,
                    double *__restrict__ span_e /* [rows][nspan][NS2] */, int nspan, int span_off) {
    const int64_t span = blockIdx.x + 1, row = blockIdx.y;
    const int64_t a = span * span_len;               // first logical sample of the span
    const double *xr = x + row * ldx;
    double acc[NS2];
#pragma unroll
    for (int c = 0; c < NS2; ++c) acc[c] = 0.0;
    const int64_t m = settle < a ? settle : a;
    // sample d back from the span is logical index a - 1 - d: global a - 1 - d forward,
    // n_total - a + d reversed
    const double *x0 = xr + (reverse ? n_total - a : a - 1);
    const int64_t dir = reverse ? 1 : -1;
    constexpr int U = 8;                             // independent loads in flight per thread
    int64_t d = threadIdx.x;
    for (; d + (U - 1) * (int64_t)blockDim.x < m; d += U * (int64_t)blockDim.x) {
        double v[U], w[U][NS2];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t du = d + u * (int64_t)blockDim.x;
            v[u] = ld_stream(x0 + dir * du);
#pragma unroll
            for (int c = 0; c < NS2; ++c) w[u][c] = ldg(W + du * NS2 + c);
        }
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
            for (int c = 0; c < NS2; ++c) acc[c] = fma(w[u][c], v[u], acc[c]);
    }
    for (; d < m; d += blockDim.x) {
        const double v = ld_stream(x0 + dir * d);
#pragma unroll
        for (int c = 0; c < NS2; ++c) acc[c] = fma(ldg(W + d * NS2 + c), v, acc[c]);
    }
    __shared__ double part[8][NS2];
#pragma unroll
    for (int c = 0; c < NS2; ++c) {
        double v = acc[c];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5][c] = v;
    }
    __syncthreads();
    if (threadIdx.x < NS2) {
        double v = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) v += part[w][threadIdx.x];
        span_e[(row * nspan + span - span_off) * NS2 + threadIdx.x] = v;
    }
}

__global__ void sos_copy_state_kernel(const double *__restrict__ src, double *__restrict__ dst,
                                      int64_t count) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) dst[i] = src[i];
}

template <typename TIO>
__global__ void sos_state_from_sample_kernel(SosZi zi, int nsec, const TIO *__restrict__ x,
                                             int64_t ldx, int64_t rows, int64_t sample,
                                             double *__restrict__ state) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * nsec * 2) return;
    const int64_t row = i / (nsec * 2);
    const int sj = (int)(i % (nsec * 2));
    state[i] = zi.zi[sj >> 1][sj & 1] * (double)x[row * ldx + sample];
}

}  // namespace osz

using namespace osz;

struct osz_sos_plan {
    SosParams prm;
    int T = 32;                     // samples per thread of the kernel this plan uses
    bool prefetch = false;          // T = 16: register prefetch of the next block
    int64_t settle = -1;            // samples after which the start state is forgotten (-1: never)
    double *d_lanepow = nullptr;    // [sec][32][4]: A^(T*(lane+1))
    double *T16_lanepow = nullptr;  // the same for 16 samples per thread (sosdec.cu)
    // Plans are cached process-wide by their coefficients, so two producers with the
    // same filter may run one plan on two streams at once: a launch owns no mutable
    // plan state.  The per-call scratch (state copy of a time-split launch, per-span
    // states of the exact split) comes from the stream-ordered allocator
    // (cudaMallocAsync / cudaFreeAsync on the launching stream: no device
    // synchronisation, no sharing between streams); Phi = T^span_len is immutable once
    // built and kept per span length under a mutex.
    mutable std::mutex mu;
    mutable std::map<int64_t, double *> phi;   // span length -> device (2 nsec)^2 matrix
    std::vector<long double> Tmat;  // (2 nsec)^2 one-step zero-input transition, row major
    double *d_weights = nullptr;    // [settle][2 nsec]: W[d] = T^d b (sos_entering_kernel)
    SosTileTab *d_tiletab = nullptr;   // one section: tables of the tiled look-back scan (sos_tile.cuh)
};

namespace {
struct M2 {
    long double a, b, c, d;
};
M2 mul(const M2 &x, const M2 &y) {
    return {x.a * y.a + x.b * y.c, x.a * y.b + x.b * y.d, x.c * y.a + x.d * y.c,
            x.c * y.b + x.d * y.d};
}
}  // namespace

// Samples after which the cascade has forgotten its start state: the smallest n
// with ||T^n||_inf < 1e-18 (T = one-step zero-input transition of the whole
// cascade), found on the ladder T^(2^k) in long double and refined bit by bit.
// Exact in the pole multiplicities -- transients of m coinciding poles decay like
// n^(m-1) r^n, which a bound from the largest pole radius alone misses.  -1 when
// the norm never gets there (a pole on or outside the unit circle).
static int64_t settle_samples(const std::vector<long double> &T, int n) {
    const long double tol = 1e-18L;
    typedef std::vector<long double> Mat;
    auto mul = [n](const Mat &A, const Mat &B) {
        Mat C((size_t)n * n, 0.0L);
        for (int i = 0; i < n; ++i)
            for (int k = 0; k < n; ++k) {
                const long double a = A[(size_t)i * n + k];
                if (a == 0.0L) continue;
                for (int j = 0; j < n; ++j) C[(size_t)i * n + j] += a * B[(size_t)k * n + j];
            }
        return C;
    };
    auto norm = [n](const Mat &A) {
        long double best = 0.0L;
        for (int i = 0; i < n; ++i) {
            long double row = 0.0L;
            for (int j = 0; j < n; ++j) row += fabsl(A[(size_t)i * n + j]);
            if (!(row <= best)) best = row;      // also catches NaN
        }
        return best;
    };
    std::vector<Mat> ladder(1, T);               // ladder[k] = T^(2^k)
    while (true) {
        const long double v = norm(ladder.back());
        if (!(v == v) || v > 1e300L) return -1;
        if (v < tol) break;
        if (ladder.size() > 40) return -1;
        ladder.push_back(mul(ladder.back(), ladder.back()));
    }
    const int k = (int)ladder.size() - 1;
    if (k == 0) return 2 * n;
    Mat acc = ladder[k - 1];
    int64_t steps = (int64_t)1 << (k - 1);
    for (int j = k - 2; j >= 0; --j) {
        Mat cand = mul(acc, ladder[j]);
        if (norm(cand) >= tol) {
            acc.swap(cand);
            steps += (int64_t)1 << j;
        }
    }
    return steps + steps / 16 + 64;
}

extern "C" {

int osz_sos_plan_create(osz_sos_plan **out, const double *sos, int nsec) {
    if (!out || !sos || nsec < 1) return fail(OSZ_ERR_ARG, "osz_sos_plan_create: bad arguments");
    if (nsec > SOS_MAXSEC)
        return fail(OSZ_ERR_UNSUPPORTED, "osz_sos_plan_create: more than 16 sections per plan "
                                         "(split the cascade into several plans)");
    osz_sos_plan *p = new osz_sos_plan();
    p->prm.nsec = nsec;
    p->prm.pad_ = 0;
    {
        // One CTA per row: with rows CTAs in the grid the smaller footprint of
        // T = 16 buys no extra residency, and it measured slower (notch, 256 rows:
        // 1.02 ms vs 0.92 ms; 8 sections: 3.67 vs 2.74 ms).  T = 32 is the default;
        // OSZ_SOS_T=16 selects the other build (tests exercise both).
        const char *e = getenv("OSZ_SOS_T");
        p->T = e ? atoi(e) : 32;
        if (p->T != 16 && p->T != 32) p->T = 32;
        const char *pf = getenv("OSZ_SOS_PF");
        p->prefetch = pf ? atoi(pf) != 0 : false;
    }
    std::vector<double> lanepow((size_t)nsec * 32 * 4), lanepow16((size_t)nsec * 32 * 4);
    double rmax = 0.0;              // largest pole radius of the cascade
    for (int s = 0; s < nsec; ++s) {
        {
            const double a0 = sos[6 * s + 3], a1 = sos[6 * s + 4] / a0, a2 = sos[6 * s + 5] / a0;
            const double disc = a1 * a1 - 4.0 * a2;
            double r;
            if (disc < 0.0) {
                r = sqrt(a2);
            } else {
                const double sq = sqrt(disc);
                r = fmax(fabs((-a1 + sq) * 0.5), fabs((-a1 - sq) * 0.5));
            }
            if (r > rmax) rmax = r;
        }
        const double *r = sos + 6 * s;
        const long double a0 = r[3];
        if (a0 == 0.0L) {
            delete p;
            return fail(OSZ_ERR_ARG, "osz_sos_plan_create: a0 == 0");
        }
        SosSec &c = p->prm.sec[s];
        c.b0 = (double)(r[0] / a0);
        c.b1 = (double)(r[1] / a0);
        c.b2 = (double)(r[2] / a0);
        c.a1 = (double)(r[4] / a0);
        c.a2 = (double)(r[5] / a0);
        const M2 A = {-(long double)c.a1, 1.0L, -(long double)c.a2, 0.0L};
        M2 pw = {1.0L, 0.0L, 0.0L, 1.0L};   // A^i
        for (int i = 0; i < SOS_T; ++i) {
            if (i == 16) {
                c.A16[0] = (double)pw.a;
                c.A16[1] = (double)pw.b;
                c.A16[2] = (double)pw.c;
                c.A16[3] = (double)pw.d;
            }
            if (i == 8) {
                c.A8[0] = (double)pw.a;
                c.A8[1] = (double)pw.b;
                c.A8[2] = (double)pw.c;
                c.A8[3] = (double)pw.d;
            }
            c.g0[i] = (double)pw.a;
            c.g1[i] = (double)pw.b;
            pw = mul(A, pw);
        }
        M2 M = pw;                           // A^32
        M2 q = M;
        for (int k = 0; k < 5; ++k) {        // M^(2^k)
            c.P[k][0] = (double)q.a;
            c.P[k][1] = (double)q.b;
            c.P[k][2] = (double)q.c;
            c.P[k][3] = (double)q.d;
            q = mul(q, q);
        }
        c.Q[0] = (double)q.a;                // M^32
        c.Q[1] = (double)q.b;
        c.Q[2] = (double)q.c;
        c.Q[3] = (double)q.d;
        M2 Mt = M;                           // thread transition A^T
        if (p->T == 16) Mt = {c.A16[0], c.A16[1], c.A16[2], c.A16[3]};
        M2 lp = Mt;                          // Mt^(lane+1)
        for (int l = 0; l < 32; ++l) {
            double *d = &lanepow[((size_t)s * 32 + l) * 4];
            d[0] = (double)lp.a;
            d[1] = (double)lp.b;
            d[2] = (double)lp.c;
            d[3] = (double)lp.d;
            lp = mul(Mt, lp);
        }
        const M2 M16 = {c.A16[0], c.A16[1], c.A16[2], c.A16[3]};
        lp = M16;
        for (int l = 0; l < 32; ++l) {
            double *d = &lanepow16[((size_t)s * 32 + l) * 4];
            d[0] = (double)lp.a;
            d[1] = (double)lp.b;
            d[2] = (double)lp.c;
            d[3] = (double)lp.d;
            lp = mul(M16, lp);
        }
    }
    for (int s = nsec; s < SOS_MAXSEC; ++s) p->prm.sec[s] = SosSec{};
    {
        // one-step zero-input transition of the whole cascade (the next section's
        // input is this section's output): column j = step(unit state j, x = 0)
        const int ns2 = 2 * nsec;
        p->Tmat.assign((size_t)ns2 * ns2, 0.0L);
        for (int j = 0; j < ns2; ++j) {
            long double xin = 0.0L;
            for (int s2 = 0; s2 < nsec; ++s2) {
                const SosSec &c = p->prm.sec[s2];
                const long double z0 = (j == 2 * s2) ? 1.0L : 0.0L;
                const long double z1 = (j == 2 * s2 + 1) ? 1.0L : 0.0L;
                const long double yv = (long double)c.b0 * xin + z0;
                p->Tmat[(size_t)(2 * s2) * ns2 + j] =
                    (long double)c.b1 * xin - (long double)c.a1 * yv + z1;
                p->Tmat[(size_t)(2 * s2 + 1) * ns2 + j] =
                    (long double)c.b2 * xin - (long double)c.a2 * yv;
                xin = yv;
            }
        }
    }
    (void)rmax;
    p->settle = settle_samples(p->Tmat, 2 * nsec);
    std::vector<double> weights;
    if (nsec <= 2 && p->settle > 0 && p->settle <= (1 << 20)) {
        // b: the state one unit input sample leaves behind from rest; W[d] = T^d b
        const int ns2 = 2 * nsec;
        std::vector<long double> w(ns2), nxt(ns2);
        long double xin = 1.0L;
        for (int s2 = 0; s2 < nsec; ++s2) {
            const SosSec &c = p->prm.sec[s2];
            const long double yv = (long double)c.b0 * xin;
            w[2 * s2] = (long double)c.b1 * xin - (long double)c.a1 * yv;
            w[2 * s2 + 1] = (long double)c.b2 * xin - (long double)c.a2 * yv;
            xin = yv;
        }
        weights.resize((size_t)p->settle * ns2);
        for (int64_t d = 0; d < p->settle; ++d) {
            for (int i = 0; i < ns2; ++i) weights[(size_t)d * ns2 + i] = (double)w[i];
            for (int i = 0; i < ns2; ++i) {
                long double acc = 0.0L;
                for (int j = 0; j < ns2; ++j) acc += p->Tmat[(size_t)i * ns2 + j] * w[j];
                nxt[i] = acc;
            }
            w.swap(nxt);
        }
    }
    if (nsec == 1) {
        // tiled look-back scan: A^(16 p) per thread and Phi^j, Phi = A^TILE, per look-back lane
        std::vector<SosTileTab> tab(1);
        const SosSec &c = p->prm.sec[0];
        const M2 A = {-(long double)c.a1, 1.0L, -(long double)c.a2, 0.0L};
        M2 A16 = {1.0L, 0.0L, 0.0L, 1.0L};
        for (int i = 0; i < TILE_T; ++i) A16 = mul(A, A16);
        M2 pw = {1.0L, 0.0L, 0.0L, 1.0L};
        for (int t = 0; t < SOS_NT; ++t) {
            tab[0].thr[t][0] = (double)pw.a;
            tab[0].thr[t][1] = (double)pw.b;
            tab[0].thr[t][2] = (double)pw.c;
            tab[0].thr[t][3] = (double)pw.d;
            pw = mul(A16, pw);
        }
        const M2 Phi = pw;                  // A^(16 * 256)
        pw = {1.0L, 0.0L, 0.0L, 1.0L};
        for (int j = 0; j <= 32; ++j) {
            tab[0].phi[j][0] = (double)pw.a;
            tab[0].phi[j][1] = (double)pw.b;
            tab[0].phi[j][2] = (double)pw.c;
            tab[0].phi[j][3] = (double)pw.d;
            pw = mul(Phi, pw);
        }
        if (cudaMalloc(&p->d_tiletab, sizeof(SosTileTab)) != cudaSuccess ||
            cudaMemcpy(p->d_tiletab, tab.data(), sizeof(SosTileTab), cudaMemcpyHostToDevice) !=
                cudaSuccess) {
            osz_sos_plan_destroy(p);
            return fail(OSZ_ERR_CUDA, "osz_sos_plan_create: device upload failed");
        }
    }
    if (cudaMalloc(&p->d_lanepow, lanepow.size() * 8) != cudaSuccess ||
        cudaMemcpy(p->d_lanepow, lanepow.data(), lanepow.size() * 8, cudaMemcpyHostToDevice) !=
            cudaSuccess ||
        cudaMalloc(&p->T16_lanepow, lanepow16.size() * 8) != cudaSuccess ||
        cudaMemcpy(p->T16_lanepow, lanepow16.data(), lanepow16.size() * 8,
                   cudaMemcpyHostToDevice) != cudaSuccess ||
        (!weights.empty() &&
         (cudaMalloc(&p->d_weights, weights.size() * 8) != cudaSuccess ||
          cudaMemcpy(p->d_weights, weights.data(), weights.size() * 8, cudaMemcpyHostToDevice) !=
              cudaSuccess))) {
        osz_sos_plan_destroy(p);
        return fail(OSZ_ERR_CUDA, "osz_sos_plan_create: device upload failed");
    }
    *out = p;
    return OSZ_OK;
}

int osz_sos_plan_destroy(osz_sos_plan *p) {
    if (!p) return OSZ_OK;
    cudaFree(p->d_lanepow);
    cudaFree(p->T16_lanepow);
    cudaFree(p->d_weights);
    cudaFree(p->d_tiletab);
    for (auto &kv : p->phi) cudaFree(kv.second);
    delete p;
    return OSZ_OK;
}

// Phi = T^span_len (long double repeated squaring on the host), uploaded once per
// span length; immutable afterwards, so concurrent launches may share it.
static int sos_phi(const osz_sos_plan *p, int64_t span_len, const double **out) {
    std::lock_guard<std::mutex> lk(p->mu);
    auto it = p->phi.find(span_len);
    if (it != p->phi.end()) {
        *out = it->second;
        return OSZ_OK;
    }
    const int ns2 = 2 * p->prm.nsec;
    const size_t nn = (size_t)ns2 * ns2;
    std::vector<long double> acc(nn, 0.0L), base(p->Tmat), tmp(nn);
    for (int i = 0; i < ns2; ++i) acc[(size_t)i * ns2 + i] = 1.0L;
    auto matmul = [&](const std::vector<long double> &A, const std::vector<long double> &B,
                      std::vector<long double> &C) {
        for (int i = 0; i < ns2; ++i)
            for (int j = 0; j < ns2; ++j) {
                long double sum = 0.0L;
                for (int k = 0; k < ns2; ++k) sum += A[(size_t)i * ns2 + k] * B[(size_t)k * ns2 + j];
                C[(size_t)i * ns2 + j] = sum;
            }
    };
    for (int64_t e = span_len; e; e >>= 1) {
        if (e & 1) {
            matmul(base, acc, tmp);
            acc.swap(tmp);
        }
        matmul(base, base, tmp);
        base.swap(tmp);
    }
    std::vector<double> phi(nn);
    for (size_t i = 0; i < nn; ++i) phi[i] = (double)acc[i];
    double *d = nullptr;
    OSZ_CUDA(cudaMalloc(&d, nn * 8));
    // pageable source, default stream: complete on return (first use of a span length only)
    cudaError_t err = cudaMemcpy(d, phi.data(), nn * 8, cudaMemcpyHostToDevice);
    if (err != cudaSuccess) {
        cudaFree(d);
        return fail(OSZ_ERR_CUDA, std::string("osz_sos_exec_f64: Phi upload: ") +
                                      cudaGetErrorString(err));
    }
    if (p->phi.size() >= 1024) {    // a handful per workload; a pathological stream of
        cudaDeviceSynchronize();    // distinct lengths drains the device before it evicts
        for (auto &kv : p->phi) cudaFree(kv.second);
        p->phi.clear();
    }
    p->phi[span_len] = d;
    *out = d;
    return OSZ_OK;
}

}  // extern "C"

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link against libcuda)
typedef CUresult (*osz_tmap_encode_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *,
                                       const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                       const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                       CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static osz_tmap_encode_fn tmap_encoder() {
    static osz_tmap_encode_fn fn = [] {
        void *ptr = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) !=
                cudaSuccess || q != cudaDriverEntryPointSuccess)
            ptr = nullptr;
        return reinterpret_cast<osz_tmap_encode_fn>(ptr);
    }();
    return fn;
}

// Rows of samples as [rows][lines][16] with 256-line boxes; 128-byte lines and swizzle for
// float64, 64-byte lines and swizzle for float32 (sos_tile_tma_kernel).  `base` must be
// 16-byte aligned and the row pitch a multiple of 16 bytes.
template <typename TIO>
static bool tile_tensor_map(CUtensorMap *map, const TIO *base, int64_t ld, int64_t rows,
                            int64_t lines) {
    osz_tmap_encode_fn enc = tmap_encoder();
    if (!enc) return false;
    const bool f64 = sizeof(TIO) == 8;
    const cuuint64_t dims[3] = {(cuuint64_t)TILE_T, (cuuint64_t)lines, (cuuint64_t)rows};
    const cuuint64_t strides[2] = {(cuuint64_t)TILE_T * sizeof(TIO), (cuuint64_t)ld * sizeof(TIO)};
    const cuuint32_t box[3] = {(cuuint32_t)TILE_T, (cuuint32_t)SOS_NT, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    return enc(map, f64 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3,
               const_cast<TIO *>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               f64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
               CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <typename TIO>
static bool tile_tma_ok(const TIO *x, int64_t ldx, const TIO *y, int64_t ldy, int64_t n,
                        int64_t ntile, int reverse) {
    if (ntile < 2) return false;
    const int64_t first_len = n - (ntile - 1) * TILE;
    const int64_t shift = reverse ? 0 : first_len;
    const int64_t per16 = 16 / (int64_t)sizeof(TIO);          // elements per 16 bytes
    auto ok = [&](const TIO *p, int64_t ld) {
        return (reinterpret_cast<uintptr_t>(p + shift) & 15) == 0 && ld % per16 == 0 && ld >= n;
    };
    return ok(x, ldx) && (!y || ok(y, ldy));
}

// Launch of a tile kernel: an ordinary launch with ticket dealing, a COOPERATIVE launch with
// static dealing (the CTAs wait on each other's tiles, so a partly resident grid could wait
// for CTAs that another kernel keeps off the SMs; the cooperative launch starts only when
// the whole grid fits).
template <typename... KArgs, typename... Args>
static cudaError_t tile_launch(int dynamic, void (*kernel)(KArgs...), unsigned ctas, int smem,
                               cudaStream_t st, Args... args) {
    // (`dynamic` is the kernel's LAST parameter and is appended here)
    if (!dynamic) {
        int stat = 0;
        void *ptrs[] = {(void *)&args..., (void *)&stat};
        const cudaError_t err = cudaLaunchCooperativeKernel((const void *)kernel, dim3(ctas),
                                                            dim3(SOS_NT), ptrs, (size_t)smem, st);
        if (err == cudaSuccess) return err;
        // the context cannot hold the whole grid at once (MPS share, green context, ...):
        // ticket dealing does not need it to
        (void)cudaGetLastError();
    }
    kernel<<<ctas, SOS_NT, smem, st>>>(args..., 1);
    return cudaGetLastError();
}

template <typename TIO>
static int sos_exec_t(const osz_sos_plan *p, const TIO *x, int64_t ldx, int64_t rows, int64_t n,
                      int reverse, double *state, TIO *y, int64_t ldy, void *stream,
                      const double *zi_host = nullptr /* start from zi * first sample */) {
    if (!p || !x || !state) return fail(OSZ_ERR_ARG, "osz_sos_exec: null argument");
    if (rows <= 0 || n <= 0) return OSZ_OK;
    cudaStream_t st = as_stream(stream);
    const int smem = SOS_NT * (p->T + 1) * 8;
    const int64_t BLK = (int64_t)SOS_NT * p->T;
    const int ns2 = 2 * p->prm.nsec;
    // One section: tiles of 4096 samples with a decoupled look-back (sos_tile.cuh): every
    // sample read once, any number of rows fills the GPU.  OSZ_SOS_TILE=0 turns it off.
    // (read per call: the tests switch it within one process)
    const char *tile_env = getenv("OSZ_SOS_TILE");
    const int tile_mode = tile_env ? atoi(tile_env) : -1;
    if (p->d_tiletab && tile_mode != 0) {
        const int64_t ntile = (n + TILE - 1) / TILE;
        // tickets where a row has few tiles in flight (predecessors mostly finished rounds
        // ago: 256 rows 0.676 against 0.706 ms), static rounds below (32 rows 0.122 against
        // 0.138 ms); OSZ_SOS_TILE_DEAL=0/1 forces static / tickets
        const char *deal_env = getenv("OSZ_SOS_TILE_DEAL");
        // (a state-only pass -- the look-ahead, one round of tiles -- always takes tickets: an
        //  ordinary launch can start in the tail of the kernel before it, a cooperative one not)
        const int dynamic = deal_env ? (atoi(deal_env) != 0) : (rows > (int64_t)sm_count() || !y);
        SosParams1 prm1;
        prm1.nsec = 1;
        prm1.pad_ = 0;
        prm1.sec[0] = p->prm.sec[0];
        if (rows * ntile < ((int64_t)1 << 31) && rows < ((int64_t)1 << 24)) {
            const int64_t total = rows * ntile;
            const char *tma_env = getenv("OSZ_SOS_TILE_TMA");
            if ((!tma_env || atoi(tma_env) != 0) && tile_tma_ok<TIO>(x, ldx, y, ldy, n, ntile, reverse)) {
                // float64, 16-byte aligned tiles: boxes moved by the TMA
                const int64_t first_len = n - (ntile - 1) * TILE;
                const int64_t shift = reverse ? 0 : first_len;
                const int64_t lines = (ntile - 1) * SOS_NT;
                alignas(64) CUtensorMap mx, my;
                const TIO *xd = x;
                TIO *yd = y;
                bool ok = tile_tensor_map<TIO>(&mx, xd + shift, ldx, rows, lines);
                if (ok && y) ok = tile_tensor_map<TIO>(&my, yd + shift, ldy, rows, lines);
                if (ok) {
                    if (!y) my = mx;
                    char *scr = nullptr;
                    const size_t bytes = 16 + (size_t)total * 32;
                    OSZ_CUDA(scratch_alloc((void **)&scr, bytes, st));
                    OSZ_CUDA(cudaMemsetAsync(scr, 0xFF, bytes, st));   // nothing published, rank -1
                    unsigned *ticket = reinterpret_cast<unsigned *>(scr);
                    double2 *agg = reinterpret_cast<double2 *>(scr + 16);
                    double2 *incl = agg + total;
                    const int tsmem = 2 * TILE * (int)sizeof(TIO) + 1024;
                    int64_t ctas = total;
                    cudaError_t lerr = cudaSuccess;
#define OSZ_TMA_LAUNCH(W)                                                                        \
    do {                                                                                         \
        static const int per_sm = [] {   /* the CTAs wait on each other: co-resident grid */      \
            int v = 0;                                                                           \
            if (cudaFuncSetAttribute(sos_tile_tma_kernel<W, TIO>,                                \
                                     cudaFuncAttributeMaxDynamicSharedMemorySize,                \
                                     2 * TILE * (int)sizeof(TIO) + 1024) != cudaSuccess ||       \
                cudaOccupancyMaxActiveBlocksPerMultiprocessor(                                   \
                    &v, sos_tile_tma_kernel<W, TIO>, SOS_NT,                                     \
                    2 * TILE * (int)sizeof(TIO) + 1024) != cudaSuccess)                          \
                v = 0;                                                                           \
            return v;                                                                            \
        }();                                                                                     \
        if (per_sm < 1) lerr = cudaErrorLaunchOutOfResources;                                    \
        /* (the attribute is per device: a process that drives several sets it on each) */       \
        if (lerr == cudaSuccess)                                                                 \
            lerr = cudaFuncSetAttribute(sos_tile_tma_kernel<W, TIO>,                             \
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, tsmem);     \
        if (lerr == cudaSuccess) {                                                               \
            if (ctas > (int64_t)per_sm * sm_count()) ctas = (int64_t)per_sm * sm_count();        \
            lerr = tile_launch(dynamic, sos_tile_tma_kernel<W, TIO>, (unsigned)ctas, tsmem, st,  \
                               prm1,                                                             \
                               mx, my, (const SosTileTab *)p->d_tiletab, xd, ldx, (int)rows, n,  \
                               reverse, (const double *)state, state, yd, ldy,                   \
                               (const double *)p->T16_lanepow, ticket, agg, incl, (int)ntile,    \
                               (int)(zi_host != nullptr), zi_host ? zi_host[0] : 0.0,            \
                               zi_host ? zi_host[1] : 0.0);                                      \
        }                                                                                        \
    } while (0)
                    if (y) OSZ_TMA_LAUNCH(true);
                    else OSZ_TMA_LAUNCH(false);
#undef OSZ_TMA_LAUNCH
                    cudaFreeAsync(scr, st);
                    if (lerr != cudaSuccess)
                        return fail(OSZ_ERR_CUDA, std::string("sos_tile_tma_kernel launch: ") +
                                                      cudaGetErrorString(lerr));
                    g_launches.fetch_add(1, std::memory_order_relaxed);
                    return OSZ_OK;
                }
            }
        }
        // Unaligned rows / float32 samples: plain loads and stores through a transposing
        // buffer.  Measured on B200 (notch, 1e6 samples): ahead of one CTA per row below
        // ~48 rows (32 rows 0.176 against 0.218 ms), behind it above (256 rows 1.09 / 0.81).
        if (rows * ntile < ((int64_t)1 << 31) && rows < ((int64_t)1 << 24) &&
            (tile_mode > 0 || rows * 3 <= sm_count())) {
            const int64_t total = rows * ntile;
            const size_t flag_bytes = (((size_t)total + 4) * 4 + 15) & ~(size_t)15;
            char *scr = nullptr;
            OSZ_CUDA(scratch_alloc((void **)&scr, flag_bytes + (size_t)total * 32, st));
            OSZ_CUDA(cudaMemsetAsync(scr, 0, flag_bytes, st));
            unsigned *ticket = reinterpret_cast<unsigned *>(scr);
            unsigned *flag = ticket + 4;
            double2 *agg = reinterpret_cast<double2 *>(scr + flag_bytes);
            double2 *incl = agg + total;
            const int tsmem = SOS_NT * (TILE_T + 1) * (int)sizeof(TIO);
            int64_t ctas = total;
#define OSZ_TILE_LAUNCH(W, YY, LDY)                                                             \
    do {                                                                                        \
        static const int per_sm = [] {   /* the CTAs wait on each other: co-resident grid */     \
            int v = 0;                                                                          \
            if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&v, sos_tile_kernel<W, TIO>,      \
                                                              SOS_NT, tsmem) != cudaSuccess)    \
                v = 0;                                                                          \
            return v;                                                                           \
        }();                                                                                    \
        if (per_sm < 1) return fail(OSZ_ERR_CUDA, "sos_tile_kernel: no CTA fits an SM");        \
        if (ctas > (int64_t)per_sm * sm_count()) ctas = (int64_t)per_sm * sm_count();           \
        tile_launch(dynamic, sos_tile_kernel<W, TIO>, (unsigned)ctas, tsmem, st, prm1,          \
                    (const SosTileTab *)p->d_tiletab, x, ldx, (int)rows, n, reverse,            \
                    (const double *)state, state, YY, (int64_t)(LDY),                           \
                    (const double *)p->T16_lanepow, ticket, flag, agg, incl, (int)ntile,        \
                    (int)(zi_host != nullptr), zi_host ? zi_host[0] : 0.0,                      \
                    zi_host ? zi_host[1] : 0.0);                                                \
    } while (0)
            if (y) OSZ_TILE_LAUNCH(true, y, ldy);
            else OSZ_TILE_LAUNCH(false, (TIO *)nullptr, 0);
#undef OSZ_TILE_LAUNCH
            const cudaError_t lerr = cudaGetLastError();
            cudaFreeAsync(scr, st);
            if (lerr != cudaSuccess)
                return fail(OSZ_ERR_CUDA, std::string("sos_tile_kernel launch: ") +
                                              cudaGetErrorString(lerr));
            g_launches.fetch_add(1, std::memory_order_relaxed);
            return OSZ_OK;
        }
    }
    if (zi_host) {
        // generic kernels: the start state zi * (first sample processed) as its own launch
        SosZi z;
        for (int sct = 0; sct < p->prm.nsec; ++sct) {
            z.zi[sct][0] = zi_host[2 * sct];
            z.zi[sct][1] = zi_host[2 * sct + 1];
        }
        const int64_t count = rows * p->prm.nsec * 2;
        sos_state_from_sample_kernel<TIO><<<(unsigned)((count + 255) / 256), 256, 0, st>>>(
            z, p->prm.nsec, x, ldx, rows, reverse ? n - 1 : 0, state);
        OSZ_LAUNCHED("sos_state_from_sample_kernel");
    }
    // Spans per row.  Splitting pays only while one CTA per row leaves SMs idle
    // (rows <= SM count).  Two ways to cut a row:
    //   warm-up: every later span re-filters the `settle` samples before it from
    //            rest (cost n/k + settle per CTA); 64 rows x 8 sections 1.83 -> 1.32 ms
    //            with 2 spans, but useless on a full GPU (256 rows 2.76 -> 3.27 ms)
    //            and impossible when the filter decays slowly (settle ~ n);
    //   exact:   pass 1 reduces every span to its final state from rest, a combine
    //            kernel composes them with Phi = T^(span length), pass 2 filters
    //            every span from its true entering state (cost ~2.1 n/k).
    // OSZ_SOS_SPLIT / OSZ_SOS_EXACT force a span count of either kind.
    //   weights: (one or two sections) the entering state of every later span as a
    //            weighted sum of the `settle` samples before it (sos_entering_kernel:
    //            2 nsec FMAs per sample, no recurrence), then every span from its own
    //            entering state: cost n/k per CTA plus a small parallel pre-pass -- what
    //            lets 32 rows (one GPU's share of a 256-channel recording split over 8)
    //            fill the SMs twice -- measured slower than the warm-up split, see below.
    // OSZ_SOS_SPLIT / OSZ_SOS_EXACT / OSZ_SOS_WEIGHTS force a span count of a kind.
    int64_t nspan = 1;
    bool exact = false, weighted = false;
    static const int forced_weights = [] {
        // Measured on B200 (notch, 1e6 samples; profiles/r02_kernel_bench.md): 32 rows
        // 0.262 ms against 0.227 ms for the warm-up split, 64 rows 0.357 / 0.277, 128 rows
        // 0.547 / 0.470 -- the pre-pass reads 24 bytes per sample (sample + two weights)
        // and that costs more than re-scanning `settle` samples does.  Opt-in.
        const char *e = getenv("OSZ_SOS_WEIGHTS");
        return e ? atoi(e) : 0;             // -1: automatic, 0: off, k: k spans
    }();
    if (std::is_same<TIO, double>::value && y && p->d_weights && forced_weights != 0 &&
        n >= 2 * p->settle) {
        const int64_t min_span = p->settle > 8 * BLK ? p->settle : 8 * BLK;
        int64_t k = forced_weights > 0 ? forced_weights : (2 * (int64_t)sm_count()) / rows;
        if (k > n / min_span) k = n / min_span;
        if (k > 64) k = 64;
        if (k >= 2) {
            nspan = k;
            weighted = true;
        }
    }
    if (y && !weighted) {           // a state-only pass wants the last span only
        static const int forced = [] {
            const char *e = getenv("OSZ_SOS_SPLIT");
            return e ? atoi(e) : 0;
        }();
        static const int forced_exact = [] {
            const char *e = getenv("OSZ_SOS_EXACT");
            return e ? atoi(e) : 0;
        }();
        // Predicted time of a launch, in units of "samples one CTA filters alone":
        // per-CTA work times the slow-down once CTAs outnumber the SMs (a second
        // co-resident CTA adds ~35 %; fitted on B200 to 4...128 rows x {1, 8} sections,
        // profiles/r01_kernel_bench.md).
        const double sms = (double)sm_count();
        auto slowdown = [&](double ctas) {
            if (ctas <= sms) return 1.0;
            const double resident = ctas < 2 * sms ? ctas : 2 * sms;   // two CTAs per SM fit
            const double waves = ceil(ctas / (2 * sms));
            return waves * resident / (sms + 0.35 * (resident - sms));
        };
        // fixed cost of the exact split (two extra launches, the combine loop), in samples
        const double ms_per_msample = 0.36 + 0.172 * p->prm.nsec;
        int64_t kmax_warm = p->settle > 0 ? n / (2 * p->settle) : 0;
        int64_t kmax_exact = n / (8 * BLK);          // spans of at least eight blocks
        if (forced > 0) {
            kmax_exact = 0;
            kmax_warm = kmax_warm < forced ? kmax_warm : forced;
        } else if (forced_exact > 0) {
            kmax_warm = 0;
            kmax_exact = kmax_exact < forced_exact ? kmax_exact : forced_exact;
        }
        if (kmax_warm > 64) kmax_warm = 64;
        if (kmax_exact > 64) kmax_exact = 64;
        if (forced > 0) {
            if (kmax_warm >= 2) nspan = kmax_warm;
        } else if (forced_exact > 0) {
            if (kmax_exact >= 2) {
                nspan = kmax_exact;
                exact = true;
            }
        } else {
            double best = (double)n * slowdown((double)rows);
            const int64_t kmax = kmax_warm > kmax_exact ? kmax_warm : kmax_exact;
            for (int64_t k = 2; k <= kmax; ++k) {
                const double slow = slowdown((double)rows * k);
                if (k <= kmax_warm) {
                    const double t = ((double)n / k + (double)p->settle) * slow;
                    if (t < best) {
                        best = t;
                        nspan = k;
                        exact = false;
                    }
                }
                if (k <= kmax_exact) {
                    const double fixed = (0.06 + 4e-6 * k * ns2 * ns2) / ms_per_msample * 1e6;
                    const double t = 2.1 * (double)n / k * slow + fixed;
                    if (t < best) {
                        best = t;
                        nspan = k;
                        exact = true;
                    }
                }
            }
        }
    }
    const int64_t span_len = (n + nspan - 1) / nspan;
    const double *state_in = state;
    // per-call scratch from the stream-ordered allocator: [state copy | span finals | span entering]
    double *scratch = nullptr;
    const int64_t n_copy = nspan > 1 ? rows * ns2 : 0;
    const int64_t n_span = exact || weighted ? rows * nspan * ns2 : 0;
    if (n_copy + 2 * n_span > 0)
        OSZ_CUDA(scratch_alloc((void **)&scratch, (size_t)(n_copy + 2 * n_span) * 8, st));
    struct Release {                 // freed in stream order after the kernels below
        double *ptr;
        cudaStream_t st;
        ~Release() {
            if (ptr) cudaFreeAsync(ptr, st);
        }
    } release{scratch, st};
    if (nspan > 1) {
        // the last span writes the carried state while span 0 may still read it
        sos_copy_state_kernel<<<(unsigned)((n_copy + 255) / 256), 256, 0, st>>>(state, scratch,
                                                                               n_copy);
        OSZ_LAUNCHED("sos_copy_state_kernel");
        state_in = scratch;
    }
    double *span_f = nullptr, *span_e = nullptr;
    const double *d_phi = nullptr;
    if (exact) {
        span_f = scratch + n_copy;
        span_e = span_f + n_span;
        const int rc = sos_phi(p, span_len, &d_phi);
        if (rc != OSZ_OK) return rc;
    }
    if (weighted) {
        span_e = scratch + n_copy;
        const dim3 g2((unsigned)(nspan - 1), (unsigned)rows);
        const double *xd = reinterpret_cast<const double *>(x);      // (TIO is double here)
        if (ns2 == 2)
            sos_entering_kernel<2><<<g2, 256, 0, st>>>(p->d_weights, p->settle, xd, ldx, n, reverse,
                                                       span_len, span_e, (int)nspan, 0);
        else
            sos_entering_kernel<4><<<g2, 256, 0, st>>>(p->d_weights, p->settle, xd, ldx, n, reverse,
                                                       span_len, span_e, (int)nspan, 0);
        OSZ_LAUNCHED("sos_entering_kernel");
    }
    const dim3 grid((unsigned)rows, (unsigned)nspan);
    static const int lone_ok = [] {
        // measured on B200 (notch, 32 rows x 4 spans = 128 CTAs): 0.242 ms with the
        // prefetching build against 0.224 ms without -- the 255-register build loses more
        // to its lower issue rate than the earlier loads win; opt-in
        const char *e = getenv("OSZ_SOS_LONE");
        return e ? atoi(e) : 0;
    }();
    const bool lone = lone_ok && rows * nspan <= (int64_t)sm_count();
#define OSZ_SOS_LAUNCH(W, TT, YY, SIN, SOUT)                                                    \
    do {                                                                                        \
        if (TT == 32 && lone) {                                                                 \
            OSZ_CUDA(cudaFuncSetAttribute(sos_scan_kernel<W, 32, true, TIO>,                         \
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, smem));  \
            sos_scan_kernel<W, 32, true, TIO><<<grid, SOS_NT, smem, st>>>(                           \
                p->prm, x, ldx, n, reverse, state_in, state, (TIO *)(YY), ldy, p->d_lanepow, span_len,   \
                p->settle, SIN, SOUT);                                                          \
        } else if (TT == 16 && p->prefetch) {                                                          \
            OSZ_CUDA(cudaFuncSetAttribute(sos_scan_kernel<W, 16, true, TIO>,                         \
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, smem));  \
            sos_scan_kernel<W, 16, true, TIO><<<grid, SOS_NT, smem, st>>>(                           \
                p->prm, x, ldx, n, reverse, state_in, state, (TIO *)(YY), ldy, p->d_lanepow, span_len,   \
                p->settle, SIN, SOUT);                                                          \
        } else {                                                                                \
            OSZ_CUDA(cudaFuncSetAttribute(sos_scan_kernel<W, TT, false, TIO>,                               \
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, smem));  \
            sos_scan_kernel<W, TT, false, TIO><<<grid, SOS_NT, smem, st>>>(                                 \
                p->prm, x, ldx, n, reverse, state_in, state, (TIO *)(YY), ldy, p->d_lanepow, span_len,   \
                p->settle, SIN, SOUT);                                                          \
        }                                                                                       \
        OSZ_LAUNCHED("sos_scan_kernel");                                                        \
    } while (0)
    if (exact) {
        if (p->T == 16) OSZ_SOS_LAUNCH(false, 16, nullptr, nullptr, span_f);
        else OSZ_SOS_LAUNCH(false, 32, nullptr, nullptr, span_f);
        sos_combine_kernel<<<(unsigned)((rows + 63) / 64), 64, 0, st>>>(d_phi, span_f, span_e,
                                                                       rows, (int)nspan, ns2);
        OSZ_LAUNCHED("sos_combine_kernel");
        if (p->T == 16) OSZ_SOS_LAUNCH(true, 16, y, span_e, nullptr);
        else OSZ_SOS_LAUNCH(true, 32, y, span_e, nullptr);
    } else if (weighted) {
        if (p->T == 16) OSZ_SOS_LAUNCH(true, 16, y, span_e, nullptr);
        else OSZ_SOS_LAUNCH(true, 32, y, span_e, nullptr);
    } else if (y) {
        if (p->T == 16) OSZ_SOS_LAUNCH(true, 16, y, nullptr, nullptr);
        else OSZ_SOS_LAUNCH(true, 32, y, nullptr, nullptr);
    } else {
        if (p->T == 16) OSZ_SOS_LAUNCH(false, 16, nullptr, nullptr, nullptr);
        else OSZ_SOS_LAUNCH(false, 32, nullptr, nullptr, nullptr);
    }
#undef OSZ_SOS_LAUNCH
    return OSZ_OK;
}

extern "C" {

int osz_sos_exec_f64(const osz_sos_plan *p, const double *x, int64_t ldx, int64_t rows, int64_t n,
                     int reverse, double *state, double *y, int64_t ldy, void *stream) {
    return sos_exec_t<double>(p, x, ldx, rows, n, reverse, state, y, ldy, stream);
}
// Look-ahead pass of the forward-backward filters (numerical.py:397-399, :508-509) in one
// call: the state left by filtering x from zi * (its first sample processed).
int osz_sos_lookahead_f64(const osz_sos_plan *p, const double *zi_host, const double *x,
                          int64_t ldx, int64_t rows, int64_t n, int reverse, double *state,
                          void *stream) {
    if (!zi_host) return fail(OSZ_ERR_ARG, "osz_sos_lookahead_f64: null zi");
    return sos_exec_t<double>(p, x, ldx, rows, n, reverse, state, nullptr, 0, stream, zi_host);
}
int osz_sos_lookahead_f32(const osz_sos_plan *p, const double *zi_host, const float *x,
                          int64_t ldx, int64_t rows, int64_t n, int reverse, double *state,
                          void *stream) {
    if (!zi_host) return fail(OSZ_ERR_ARG, "osz_sos_lookahead_f32: null zi");
    return sos_exec_t<float>(p, x, ldx, rows, n, reverse, state, nullptr, 0, stream, zi_host);
}
// float32 I/O: float samples in and out; recurrence, scan and carried state in float64
int osz_sos_exec_f32(const osz_sos_plan *p, const float *x, int64_t ldx, int64_t rows, int64_t n,
                     int reverse, double *state, float *y, int64_t ldy, void *stream) {
    return sos_exec_t<float>(p, x, ldx, rows, n, reverse, state, y, ldy, stream);
}

int osz_sos_tail_state_f64(const osz_sos_plan *p, const double *x, int64_t ldx, int64_t rows,
                           int64_t n, int reverse, double *state, void *stream) {
    if (!p || !x || !state) return fail(OSZ_ERR_ARG, "osz_sos_tail_state_f64: null argument");
    if (rows <= 0) return OSZ_OK;
    if (!p->d_weights || p->settle <= 0 || n < p->settle)
        return fail(OSZ_ERR_UNSUPPORTED, "osz_sos_tail_state_f64: needs a cascade of one or two "
                                         "sections and at least `settle` samples");
    const dim3 grid(1, (unsigned)rows);
    cudaStream_t st = as_stream(stream);
    if (p->prm.nsec == 1)
        sos_entering_kernel<2><<<grid, 256, 0, st>>>(p->d_weights, p->settle, x, ldx, n, reverse, n,
                                                     state, 1, 1);
    else
        sos_entering_kernel<4><<<grid, 256, 0, st>>>(p->d_weights, p->settle, x, ldx, n, reverse, n,
                                                     state, 1, 1);
    OSZ_LAUNCHED("sos_entering_kernel");
    return OSZ_OK;
}

int64_t osz_sos_plan_settle(const osz_sos_plan *p) { return p ? p->settle : -1; }
int osz_sos_plan_has_weights(const osz_sos_plan *p) { return p && p->d_weights ? 1 : 0; }

// internal (sosdec.cu): the plan's kernel parameter block
int osz_sos_plan_params(const osz_sos_plan *p, SosParams *prm, const double **lanepow,
                        int64_t *settle) {
    if (!p) return fail(OSZ_ERR_ARG, "osz_sos_plan_params: null plan");
    if (p->T16_lanepow == nullptr)
        return fail(OSZ_ERR_UNSUPPORTED, "osz_sos_plan_params: no T = 16 tables");
    *prm = p->prm;
    *lanepow = p->T16_lanepow;
    *settle = p->settle;
    return OSZ_OK;
}

}  // extern "C"

template <typename TIO>
static int sos_state_from_sample_t(const osz_sos_plan *p, const double *zi, const TIO *x,
                                   int64_t ldx, int64_t rows, int64_t sample, double *state,
                                   void *stream) {
    if (!p || !zi || !x || !state)
        return fail(OSZ_ERR_ARG, "osz_sos_state_from_sample: null argument");
    if (rows <= 0) return OSZ_OK;
    SosZi z{};
    for (int s = 0; s < p->prm.nsec; ++s) {
        z.zi[s][0] = zi[2 * s];
        z.zi[s][1] = zi[2 * s + 1];
    }
    const int64_t total = rows * p->prm.nsec * 2;
    sos_state_from_sample_kernel<TIO><<<(unsigned)((total + 255) / 256), 256, 0,
                                        as_stream(stream)>>>(z, p->prm.nsec, x, ldx, rows, sample,
                                                             state);
    OSZ_LAUNCHED("sos_state_from_sample_kernel");
    return OSZ_OK;
}

extern "C" {

int osz_sos_state_from_sample_f64(const osz_sos_plan *p, const double *zi, const double *x,
                                  int64_t ldx, int64_t rows, int64_t sample, double *state,
                                  void *stream) {
    return sos_state_from_sample_t<double>(p, zi, x, ldx, rows, sample, state, stream);
}
int osz_sos_state_from_sample_f32(const osz_sos_plan *p, const double *zi, const float *x,
                                  int64_t ldx, int64_t rows, int64_t sample, double *state,
                                  void *stream) {
    return sos_state_from_sample_t<float>(p, zi, x, ldx, rows, sample, state, stream);
}

}  // extern "C"
