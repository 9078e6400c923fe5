// Windowed DFT of sliding segments: Welch accumulate, per-segment periodogram
// and STFT.  Replaces, per segment, detrend -> window -> rfft -> scale
// (reference core/numerical.py:691-716) and |X|^2 with one-sided doubling
// (:781-794), over the windows of _spectra_estimatives (:817-849).
//
// Power-of-two path (256 <= nfft <= 8192): TWO consecutive segments of a row
// are transformed as the real and imaginary part of one nfft-point complex FFT
// held in registers + shared memory (fft_core.cuh).  For Welch the two
// segments' periodograms are summed anyway, and
//     |X_a[k]|^2 + |X_b[k]|^2 = (|Z[k]|^2 + |Z[N-k]|^2) / 2 ,
// so the kernel only accumulates |Z|^2 per bin in registers over all of a
// row's segment pairs and folds bins k and N-k when it adds the result into
// psd_sum -- no untangling pass, no per-segment memory traffic.
// The STFT / periodogram kernels untangle X_a, X_b through shared memory.
//
// Other nfft (not a power of two, or outside 256..8192) use the generic path
// in spectra_generic.cu.
#include <stdlib.h>

#include <type_traits>
#include <vector>

#include "common.cuh"
#include "fft_core.cuh"

namespace osz {

enum { SPEC_ACCUM = 0, SPEC_PGRAM = 1, SPEC_STFT = 2 };

// Sum of up to four doubles over the CTA (NT threads).  `red` is 4*32 doubles.
template <int NV, int NT>
__device__ __forceinline__ void block_sum(double (&val)[NV], double *red, int tid) {
    constexpr int LANES = NT < 32 ? NT : 32;            // N = 256 runs 16 threads
    constexpr unsigned MASK = NT < 32 ? ((1u << NT) - 1u) : 0xffffffffu;
    constexpr int NW = NT < 32 ? 1 : NT / 32;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
#pragma unroll
        for (int o = LANES / 2; o > 0; o >>= 1) val[i] += __shfl_xor_sync(MASK, val[i], o);
    }
    if (NW == 1) return;                                 // every lane holds the sum
    if ((tid & 31) == 0) {
#pragma unroll
        for (int i = 0; i < NV; ++i) red[i * 32 + (tid >> 5)] = val[i];
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < NW; ++w) s += red[i * 32 + w];   // same order in every thread
        val[i] = s;
    }
    __syncthreads();
}

// Load segments a (-> .x) and b (-> .y), detrend and window them.
template <int LOG2N, int DETREND>
__device__ __forceinline__ void load_pair(double2 (&v)[16], const double *__restrict__ xa,
                                          bool has_b, int64_t stride,
                                          const double *__restrict__ win, double *red, int tid) {
    using C = FftCfg<LOG2N>;
    constexpr int N = C::N, NT = C::NT;
#pragma unroll
    for (int r = 0; r < 16; ++r) {
        const int i = tid + r * NT;
        v[r].x = ldg(xa + i);
        v[r].y = has_b ? ldg(xa + stride + i) : 0.0;
    }
    if (DETREND == OSZ_DETREND_CONSTANT) {
        double s[2] = {0.0, 0.0};
#pragma unroll
        for (int r = 0; r < 16; ++r) {
            s[0] += v[r].x;
            s[1] += v[r].y;
        }
        block_sum<2, NT>(s, red, tid);
        const double ma = s[0] / N, mb = s[1] / N;
#pragma unroll
        for (int r = 0; r < 16; ++r) {
            const double w = ldg(win + tid + r * NT);
            v[r].x = (v[r].x - ma) * w;
            v[r].y = (v[r].y - mb) * w;
        }
    } else if (DETREND == OSZ_DETREND_LINEAR) {
        // least-squares line over t = 0..N-1 (scipy.signal.detrend type='linear')
        double s[4] = {0.0, 0.0, 0.0, 0.0};
        const double tbar = 0.5 * (N - 1);
#pragma unroll
        for (int r = 0; r < 16; ++r) {
            const double tc = (double)(tid + r * NT) - tbar;
            s[0] += v[r].x;
            s[1] += v[r].y;
            s[2] = fma(tc, v[r].x, s[2]);
            s[3] = fma(tc, v[r].y, s[3]);
        }
        block_sum<4, NT>(s, red, tid);
        const double stt = (double)N * ((double)N * N - 1.0) / 12.0;   // sum (t - tbar)^2
        const double ma = s[0] / N, mb = s[1] / N;
        const double ka = s[2] / stt, kb = s[3] / stt;
#pragma unroll
        for (int r = 0; r < 16; ++r) {
            const double tc = (double)(tid + r * NT) - tbar;
            const double w = ldg(win + tid + r * NT);
            v[r].x = (v[r].x - fma(ka, tc, ma)) * w;
            v[r].y = (v[r].y - fma(kb, tc, mb)) * w;
        }
    } else {
#pragma unroll
        for (int r = 0; r < 16; ++r) {
            const double w = ldg(win + tid + r * NT);
            v[r].x *= w;
            v[r].y *= w;
        }
    }
}

// Welch: one CTA walks `pairs_per_cta` consecutive segment pairs of one row.
template <int LOG2N, int DETREND>
__global__ void __launch_bounds__(FftCfg<LOG2N>::NT, (LOG2N <= 12 ? 2 : 1))
welch_accum_kernel(const double *__restrict__ x, int64_t ldx, int64_t nseg, int64_t stride,
                   const double *__restrict__ win, const double2 *__restrict__ tw, double norm,
                   double *__restrict__ psd_sum, int64_t ldp, int64_t pairs_per_cta) {
    using C = FftCfg<LOG2N>;
    constexpr int N = C::N, NT = C::NT;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double2 *sm = reinterpret_cast<double2 *>(smem_raw);
    // per-thread |Z|^2 accumulators live in shared memory ([r][tid], conflict
    // free): in registers they pushed the FFT over the 128-register budget of
    // two CTAs per SM and spilled (profiles/r01_ncu_summary.md)
    double *accs = reinterpret_cast<double *>(smem_raw + C::SMEM_BYTES);
    __shared__ double red[4 * 32];

    const int tid = threadIdx.x;
    const int64_t row = blockIdx.y;
    const int64_t npairs = (nseg + 1) / 2;
    const int64_t p0 = (int64_t)blockIdx.x * pairs_per_cta;
    int64_t p1 = p0 + pairs_per_cta;
    if (p1 > npairs) p1 = npairs;
    const double *xr = x + row * ldx;
    const FftTw ftw = fft_load_tw<LOG2N>(tw, tid);

#pragma unroll
    for (int r = 0; r < 16; ++r) accs[r * NT + tid] = 0.0;

    for (int64_t p = p0; p < p1; ++p) {
        double2 v[16];
        if (p + 1 < p1) {
            // pull the next pair's samples towards L2 while this pair is transformed
            // (the loads below otherwise wait on DRAM: long_scoreboard was the top stall)
            const double *nx = xr + 2 * (p + 1) * stride;
            const int64_t span_elems = stride + N;              // samples the next pair reads
            for (int64_t e = (int64_t)tid * 16; e < span_elems; e += (int64_t)NT * 16)
                asm volatile("prefetch.global.L2 [%0];" ::"l"(nx + e));
        }
        load_pair<LOG2N, DETREND>(v, xr + 2 * p * stride, 2 * p + 1 < nseg, stride, win, red, tid);
        fft_r2r<LOG2N>(v, sm, ftw, tid);
#pragma unroll
        for (int r = 0; r < 16; ++r)
            accs[r * NT + tid] = fma(v[r].x, v[r].x, fma(v[r].y, v[r].y, accs[r * NT + tid]));
    }
    if (p0 >= p1) return;
    // fold k and N-k:  psd[k] += norm * (A[k] + A[N-k]) for 0<k<N/2 (this is the
    // one-sided doubling), psd[0] += norm*A[0], psd[N/2] += norm*A[N/2].
    double *out = psd_sum + row * ldp;
#pragma unroll
    for (int r = 0; r < 16; ++r) {
        const int idx = tid + r * NT;
        const int bin = idx <= N / 2 ? idx : N - idx;
        atomicAdd(out + bin, accs[r * NT + tid] * norm);
    }
}

// Ping-pong Welch (256 <= nfft <= 4096): one persistent CTA per SM, two thread
// groups, each walking its own contiguous run of segment pairs (row-major over
// (row, pair)); the groups alternate on the FP64 pipe (fft_core.cuh
// SyncPingPong), so one group's butterflies overlap the other's shared-memory
// exchanges.  A pair's samples (stride + nfft contiguous doubles) arrive by one
// TMA bulk copy into the group's exchange buffer, issued as soon as the previous
// pair's last pass has read its inputs.  |Z|^2 accumulates in shared memory and
// is folded into psd_sum whenever the run crosses into another row.
template <int LOG2N>
struct WelchPP {
    using C = FftCfg<LOG2N>;
    static constexpr int OFF_BAR = C::TW_TOTAL * 16;
    static constexpr int OFF_RED = OFF_BAR + 16;                       // [group][parity][2][8 warps]
    static constexpr int OFF_G = (OFF_RED + 2 * 2 * 2 * 8 * 8 + 127) & ~127;
    static constexpr int GROUP_BYTES = C::SMEM_BYTES + C::N * 8;      // exchange + accumulators
    static constexpr int SMEM = OFF_G + 2 * GROUP_BYTES;
    static constexpr int CTAS_PER_SM = 4096 / C::N;                   // 512 threads per SM
};

template <int LOG2N, int DETREND, bool TOKEN = true>
__global__ void __launch_bounds__(2 * FftCfg<LOG2N>::NT, WelchPP<LOG2N>::CTAS_PER_SM)
welch_pp_kernel(const double *__restrict__ x, int64_t ldx, int64_t nseg, int64_t stride,
                const double *__restrict__ win, const double2 *__restrict__ tw, double norm,
                double *__restrict__ psd_sum, int64_t ldp, int64_t npairs, int64_t nwork,
                int64_t per_group, int lag, int zero) {
    using C = FftCfg<LOG2N>;
    using L = WelchPP<LOG2N>;
    using Sync = typename std::conditional<TOKEN, SyncPingPong<LOG2N>, SyncGroups<LOG2N>>::type;
    constexpr int N = C::N, NT = C::NT, NW = NT / 32 > 0 ? NT / 32 : 1;
    static_assert(NT >= 32, "ping-pong Welch needs whole warps per group");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int g = threadIdx.x / NT;
    const int tid = threadIdx.x - g * NT;
    double2 *tw_sm = reinterpret_cast<double2 *>(smem_raw);
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem_raw + L::OFF_BAR) + g;
    double *red = reinterpret_cast<double *>(smem_raw + L::OFF_RED) + g * 32;
    double2 *sm = reinterpret_cast<double2 *>(smem_raw + L::OFF_G + g * L::GROUP_BYTES);
    double *sx = reinterpret_cast<double *>(sm);
    double *accs = reinterpret_cast<double *>(smem_raw + L::OFF_G + g * L::GROUP_BYTES +
                                              C::SMEM_BYTES);

    for (int i = threadIdx.x; i < C::TW_TOTAL; i += 2 * NT) tw_sm[i] = ldg(tw + i);
    if (tid == 0) {
        mbar_init(bar, 1);
        fence_barrier_init();
    }
#pragma unroll
    for (int r = 0; r < 16; ++r) accs[r * NT + tid] = 0.0;
    __syncthreads();
    const Sync sync{g, tid, zero, tw_sm};

    // this group's run of work items [w0, w1); item w = row * npairs + pair
    const int64_t w0 = ((int64_t)blockIdx.x * 2 + g) * per_group;
    int64_t w1 = w0 + per_group;
    if (w1 > nwork) w1 = nwork;
    int64_t row = w0 < nwork ? w0 / npairs : 0;
    int64_t pair = w0 < nwork ? w0 - row * npairs : 0;

    auto issue = [&](int64_t w_, int64_t row_, int64_t pair_) {
        if (w_ >= w1) return;
        const int64_t need = 2 * pair_ + 1 < nseg ? stride + N : N;
        tma_fetch_span(sx, x + row_ * ldx + 2 * pair_ * stride, need, bar);
    };
    auto flush = [&](int64_t row_) {
        // fold k and N-k:  psd[k] += norm * (A[k] + A[N-k]) for 0 < k < N/2 (the
        // one-sided doubling), psd[0] += norm * A[0], psd[N/2] += norm * A[N/2]
        double *out = psd_sum + row_ * ldp;
#pragma unroll
        for (int r = 0; r < 16; ++r) {
            const int idx = tid + r * NT;
            const int bin = idx <= N / 2 ? idx : N - idx;
            atomicAdd(out + bin, accs[r * NT + tid] * norm);
            accs[r * NT + tid] = 0.0;
        }
    };
    if (tid == 0) issue(w0, row, pair);
    sync.prime();
    if (g == 1)
        for (int i = 0; i < lag; ++i) sync.idle_turn();

    for (int64_t it = 0; it < per_group; ++it) {
        const int64_t w = w0 + it;
        const bool live = w < w1;
        int64_t rown = row, pairn = pair + 1;
        if (pairn >= npairs) {
            pairn = 0;
            ++rown;
        }
        double2 v[16];
        double wv[16];
#pragma unroll
        for (int r = 0; r < 16; ++r) wv[r] = ldg(win + tid + r * NT);
        if (live) {
            const double *xa = sx + span_mis(x + row * ldx + 2 * pair * stride) + tid;
            const double *xb = xa + stride;
            while (!mbar_try_wait(bar, (uint32_t)(it & 1))) {
            }
            if (2 * pair + 1 < nseg) {
                if (stride == N / 2) {
                    // 50 % overlap: the second segment's first half is the first one's second
#pragma unroll
                    for (int r = 0; r < 16; ++r) v[r].x = xa[r * NT];
#pragma unroll
                    for (int r = 0; r < 8; ++r) v[r].y = v[r + 8].x;
#pragma unroll
                    for (int r = 8; r < 16; ++r) v[r].y = xa[(r + 8) * NT];
                } else {
#pragma unroll
                    for (int r = 0; r < 16; ++r) v[r] = make_double2(xa[r * NT], xb[r * NT]);
                }
            } else {
#pragma unroll
                for (int r = 0; r < 16; ++r) v[r] = make_double2(xa[r * NT], 0.0);
            }
        } else {
#pragma unroll
            for (int r = 0; r < 16; ++r) v[r] = make_double2(0.0, 0.0);
        }
        if (DETREND != OSZ_DETREND_NONE) {
            // segment sums: per thread, per warp (shuffles), per group (shared memory)
            constexpr int NV = DETREND == OSZ_DETREND_LINEAR ? 4 : 2;
            double s[4] = {0.0, 0.0, 0.0, 0.0};
            const double tbar = 0.5 * (N - 1);
#pragma unroll
            for (int r = 0; r < 16; ++r) {
                s[0] += v[r].x;
                s[1] += v[r].y;
                if (DETREND == OSZ_DETREND_LINEAR) {
                    const double tc = (double)(tid + r * NT) - tbar;
                    s[2] = fma(tc, v[r].x, s[2]);
                    s[3] = fma(tc, v[r].y, s[3]);
                }
            }
#pragma unroll
            for (int i = 0; i < NV; ++i) {
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) s[i] += __shfl_xor_sync(0xffffffffu, s[i], o);
            }
            double *rd = red;                       // (single buffer: see the barrier below)
            if ((tid & 31) == 0) {
#pragma unroll
                for (int i = 0; i < NV; ++i) rd[i * 8 + (tid >> 5)] = s[i];
            }
            sync.group();
#pragma unroll
            for (int i = 0; i < NV; ++i) {
                double t = 0.0;
#pragma unroll
                for (int q = 0; q < NW; ++q) t += rd[i * 8 + q];   // same order in every thread
                s[i] = t;
            }
            // (the next write to `red` comes after this item's exchange barriers)
            const int tk = sync.take();
            const double ma = Sync::tie(s[0] / N, tk), mb = Sync::tie(s[1] / N, tk);
            if (DETREND == OSZ_DETREND_CONSTANT) {
#pragma unroll
                for (int r = 0; r < 16; ++r) {
                    v[r].x = (v[r].x - ma) * wv[r];
                    v[r].y = (v[r].y - mb) * wv[r];
                }
            } else {
                // least-squares line over t = 0..N-1 (scipy.signal.detrend type='linear')
                const double stt = (double)N * ((double)N * N - 1.0) / 12.0;   // sum (t - tbar)^2
                const double ka = s[2] / stt, kb = s[3] / stt;
#pragma unroll
                for (int r = 0; r < 16; ++r) {
                    const double tc = (double)(tid + r * NT) - tbar;
                    v[r].x = (v[r].x - fma(ka, tc, ma)) * wv[r];
                    v[r].y = (v[r].y - fma(kb, tc, mb)) * wv[r];
                }
            }
        } else {
            const int tk = sync.take();
#pragma unroll
            for (int r = 0; r < 16; ++r) {
                const double w_ = Sync::tie(wv[r], tk);
                v[r].x *= w_;
                v[r].y *= w_;
            }
        }
        bfly<16>(v);
        sync.release(v);
        fft_r2r_tail<LOG2N, Sync, true>(v, sm, tid, sync, [&]() {
            sync.group();                 // every thread of the group has read its inputs
            if (tid == 0) {
                fence_proxy_async();
                issue(w + 1, rown, pairn);
            }
        });
#pragma unroll
        for (int r = 0; r < 16; ++r)
            accs[r * NT + tid] = fma(v[r].x, v[r].x, fma(v[r].y, v[r].y, accs[r * NT + tid]));
        if (live && (rown != row || w + 1 >= w1)) flush(row);
        row = rown;
        pair = pairn;
    }
    if (g == 0)
        for (int i = 0; i < lag; ++i) sync.idle_turn();
    sync.drain();
}

// float32-compute variant of the ping-pong Welch kernel (opt-in, the plan's
// compute mode): float64 samples in, float64 sums out, window product and
// transform in float32 (namespace oszf of fft_core.cuh).  The float64 kernel is
// bound by the FP64 pipe at ~25 % of the 8-bytes-per-sample HBM roofline; the
// FP32 pipe issues at twice that rate and the exchange buffers halve.
//   * Samples are centred in float64 BEFORE they are narrowed: d = x - c with c
//     the first sample of the pair's span, so a DC offset far above the signal's
//     fluctuation costs no float32 precision.  The detrend sums run in float64
//     over the d (exact input to the float32 stage), the fitted mean / line is
//     then removed in float32.
//   * |Z|^2 is formed in float32 and accumulated in float64 REGISTERS (the
//     float32 transform leaves room for them), folded into psd_sum whenever the
//     run crosses into another row.
//   * The span has its own staging buffer, so the next pair's TMA copy is issued
//     as soon as this pair's samples are in registers.
template <int LOG2N>
struct WelchPPC32 {
    using C = oszf::FftCfg<LOG2N>;
    static constexpr int OFF_BAR = C::TW_TOTAL * 8;
    static constexpr int OFF_RED = OFF_BAR + 16;                       // [group][4][8 warps]
    static constexpr int OFF_WIN = (OFF_RED + 2 * 4 * 8 * 8 + 127) & ~127;
    static constexpr int OFF_G = (OFF_WIN + C::N * 4 + 127) & ~127;
    static constexpr int STAGE_BYTES = ((2 * C::N + 2) * 8 + 127) & ~127;   // stride + N (+ misalignment)
    static constexpr int GROUP_BYTES = C::SMEM_BYTES + STAGE_BYTES;
    static constexpr int SMEM = OFF_G + 2 * GROUP_BYTES;
    // 512 threads per SM where shared memory allows (it does down to nfft 1024)
    static constexpr int FIT = (228 * 1024) / (SMEM + 1024);
    static constexpr int CTAS_PER_SM = 4096 / C::N < FIT ? 4096 / C::N : FIT;
};

// TIO: the samples' type in memory -- double, or float for the float32 I/O mode (4 bytes
// per sample of HBM traffic instead of 8; sums and the centring stay float64).
template <int LOG2N, int DETREND, bool TOKEN, typename TIO>
__global__ void __launch_bounds__(2 * oszf::FftCfg<LOG2N>::NT, WelchPPC32<LOG2N>::CTAS_PER_SM)
welch_pp_c32_kernel(const TIO *__restrict__ x, int64_t ldx, int64_t nseg, int64_t stride,
                    const float *__restrict__ win, const float2 *__restrict__ tw, double norm,
                    double *__restrict__ psd_sum, int64_t ldp, int64_t npairs, int64_t nwork,
                    int64_t per_group, int lag, int zero) {
    using C = oszf::FftCfg<LOG2N>;
    using L = WelchPPC32<LOG2N>;
    using Sync = typename std::conditional<TOKEN, oszf::SyncPingPong<LOG2N>,
                                           oszf::SyncGroups<LOG2N>>::type;
    constexpr int N = C::N, NT = C::NT, NW = NT / 32;
    static_assert(NT >= 32, "ping-pong Welch needs whole warps per group");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int g = threadIdx.x / NT;
    const int tid = threadIdx.x - g * NT;
    float2 *tw_sm = reinterpret_cast<float2 *>(smem_raw);
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem_raw + L::OFF_BAR) + g;
    double *red = reinterpret_cast<double *>(smem_raw + L::OFF_RED) + g * 32;
    float *win_sm = reinterpret_cast<float *>(smem_raw + L::OFF_WIN);
    float2 *sm = reinterpret_cast<float2 *>(smem_raw + L::OFF_G + g * L::GROUP_BYTES);
    TIO *sx = reinterpret_cast<TIO *>(smem_raw + L::OFF_G + g * L::GROUP_BYTES + C::SMEM_BYTES);

    for (int i = threadIdx.x; i < C::TW_TOTAL; i += 2 * NT) tw_sm[i] = ldg(tw + i);
    for (int i = threadIdx.x; i < N; i += 2 * NT) win_sm[i] = ldg(win + i);
    if (tid == 0) {
        mbar_init(bar, 1);
        fence_barrier_init();
    }
    __syncthreads();
    const Sync sync{g, tid, zero, tw_sm};
    double acc[16];
#pragma unroll
    for (int r = 0; r < 16; ++r) acc[r] = 0.0;

    // this group's run of work items [w0, w1); item w = row * npairs + pair
    const int64_t w0 = ((int64_t)blockIdx.x * 2 + g) * per_group;
    int64_t w1 = w0 + per_group;
    if (w1 > nwork) w1 = nwork;
    int64_t row = w0 < nwork ? w0 / npairs : 0;
    int64_t pair = w0 < nwork ? w0 - row * npairs : 0;

    auto issue = [&](int64_t w_, int64_t row_, int64_t pair_) {
        if (w_ >= w1) return;
        const int64_t need = 2 * pair_ + 1 < nseg ? stride + N : N;
        tma_fetch_span(sx, x + row_ * ldx + 2 * pair_ * stride, need, bar);
    };
    auto flush = [&](int64_t row_) {
        // fold k and N-k:  psd[k] += norm * (A[k] + A[N-k]) for 0 < k < N/2 (the
        // one-sided doubling), psd[0] += norm * A[0], psd[N/2] += norm * A[N/2]
        double *out = psd_sum + row_ * ldp;
#pragma unroll
        for (int r = 0; r < 16; ++r) {
            const int idx = tid + r * NT;
            const int bin = idx <= N / 2 ? idx : N - idx;
            atomicAdd(out + bin, acc[r] * norm);
            acc[r] = 0.0;
        }
    };
    if (tid == 0) issue(w0, row, pair);
    sync.prime();
    if (g == 1)
        for (int i = 0; i < lag; ++i) sync.idle_turn();

    for (int64_t it = 0; it < per_group; ++it) {
        const int64_t w = w0 + it;
        const bool live = w < w1;
        int64_t rown = row, pairn = pair + 1;
        if (pairn >= npairs) {
            pairn = 0;
            ++rown;
        }
        float2 v[16];
        double s[4] = {0.0, 0.0, 0.0, 0.0};
        constexpr double tbar = 0.5 * (N - 1);
        if (live) {
            const TIO *x0 = sx + span_mis(x + row * ldx + 2 * pair * stride);
            const TIO *xa = x0 + tid;
            const TIO *xb = xa + stride;
            while (!mbar_try_wait(bar, (uint32_t)(it & 1))) {
            }
            const double c = DETREND != OSZ_DETREND_NONE ? (double)x0[0] : 0.0;
            const bool has_b = 2 * pair + 1 < nseg;
            if (has_b && stride == N / 2) {
                // 50 % overlap: the second segment's first half is the first one's second
                double d[24];
#pragma unroll
                for (int r = 0; r < 24; ++r) d[r] = (double)xa[r * NT] - c;
#pragma unroll
                for (int r = 0; r < 16; ++r) v[r] = make_float2((float)d[r], (float)d[r + 8]);
                if (DETREND != OSZ_DETREND_NONE) {
#pragma unroll
                    for (int r = 0; r < 16; ++r) {
                        s[0] += d[r];
                        s[1] += d[r + 8];
                        if (DETREND == OSZ_DETREND_LINEAR) {
                            const double tc = (double)(tid + r * NT) - tbar;
                            s[2] = fma(tc, d[r], s[2]);
                            s[3] = fma(tc, d[r + 8], s[3]);
                        }
                    }
                }
            } else {
#pragma unroll
                for (int r = 0; r < 16; ++r) {
                    const double da = (double)xa[r * NT] - c;
                    const double db = has_b ? (double)xb[r * NT] - c : 0.0;
                    v[r] = make_float2((float)da, (float)db);
                    if (DETREND != OSZ_DETREND_NONE) {
                        s[0] += da;
                        s[1] += db;
                        if (DETREND == OSZ_DETREND_LINEAR) {
                            const double tc = (double)(tid + r * NT) - tbar;
                            s[2] = fma(tc, da, s[2]);
                            s[3] = fma(tc, db, s[3]);
                        }
                    }
                }
            }
        } else {
#pragma unroll
            for (int r = 0; r < 16; ++r) v[r] = make_float2(0.0f, 0.0f);
        }
        if (DETREND != OSZ_DETREND_NONE) {
            // segment sums: per warp (shuffles), per group (shared memory)
            constexpr int NV = DETREND == OSZ_DETREND_LINEAR ? 4 : 2;
#pragma unroll
            for (int i = 0; i < NV; ++i) {
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) s[i] += __shfl_xor_sync(0xffffffffu, s[i], o);
            }
            if ((tid & 31) == 0) {
#pragma unroll
                for (int i = 0; i < NV; ++i) red[i * 8 + (tid >> 5)] = s[i];
            }
        }
        // every thread of the group holds its samples: the staging buffer is free
        // for the next pair (and `red` is complete)
        sync.group();
        if (tid == 0) {
            fence_proxy_async();
            issue(w + 1, rown, pairn);
        }
        if (DETREND != OSZ_DETREND_NONE) {
            constexpr int NV = DETREND == OSZ_DETREND_LINEAR ? 4 : 2;
#pragma unroll
            for (int i = 0; i < NV; ++i) {
                double t = 0.0;
#pragma unroll
                for (int q = 0; q < NW; ++q) t += red[i * 8 + q];   // same order in every thread
                s[i] = t;
            }
            // (the next write to `red` comes after this item's exchange barriers)
            const float ma0 = (float)(s[0] / N), mb0 = (float)(s[1] / N);
            const int tk = sync.take();
            const float ma = Sync::tie(ma0, tk), mb = Sync::tie(mb0, tk);
            if (DETREND == OSZ_DETREND_CONSTANT) {
#pragma unroll
                for (int r = 0; r < 16; ++r) {
                    const float wv = win_sm[tid + r * NT];
                    v[r].x = (v[r].x - ma) * wv;
                    v[r].y = (v[r].y - mb) * wv;
                }
            } else {
                // least-squares line over t = 0..N-1 (scipy.signal.detrend type='linear')
                constexpr double stt = (double)N * ((double)N * N - 1.0) / 12.0;   // sum (t - tbar)^2
                const float ka = (float)(s[2] / stt), kb = (float)(s[3] / stt);
#pragma unroll
                for (int r = 0; r < 16; ++r) {
                    const float tc = (float)(tid + r * NT) - (float)tbar;
                    const float wv = win_sm[tid + r * NT];
                    v[r].x = (v[r].x - fmaf(ka, tc, ma)) * wv;
                    v[r].y = (v[r].y - fmaf(kb, tc, mb)) * wv;
                }
            }
        } else {
            const int tk = sync.take();
#pragma unroll
            for (int r = 0; r < 16; ++r) {
                const float w_ = Sync::tie(win_sm[tid + r * NT], tk);
                v[r].x *= w_;
                v[r].y *= w_;
            }
        }
        oszf::bfly<16>(v);
        sync.release(v);
        oszf::fft_r2r_tail<LOG2N, Sync, true>(v, sm, tid, sync);
#pragma unroll
        for (int r = 0; r < 16; ++r) acc[r] += (double)fmaf(v[r].x, v[r].x, v[r].y * v[r].y);
        if (live && (rown != row || w + 1 >= w1)) flush(row);
        row = rown;
        pair = pairn;
    }
    if (g == 0)
        for (int i = 0; i < lag; ++i) sync.idle_turn();
    sync.drain();
}

// Per-segment outputs: one CTA per (segment pair, row).
template <int LOG2N, int DETREND, int MODE>
__global__ void __launch_bounds__(FftCfg<LOG2N>::NT, (LOG2N <= 12 ? 2 : 1))
spec_segments_kernel(const double *__restrict__ x, int64_t ldx, int64_t rows, int64_t nseg,
                     int64_t stride, const double *__restrict__ win,
                     const double2 *__restrict__ tw, double norm, double *__restrict__ out) {
    using C = FftCfg<LOG2N>;
    constexpr int N = C::N, NT = C::NT, NF = N / 2 + 1;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double2 *sm = reinterpret_cast<double2 *>(smem_raw);
    __shared__ double red[4 * 32];

    const int tid = threadIdx.x;
    const int64_t row = blockIdx.y;
    const int64_t sa = (int64_t)blockIdx.x * 2;
    const bool has_b = sa + 1 < nseg;
    const FftTw ftw = fft_load_tw<LOG2N>(tw, tid);

    double2 v[16];
    load_pair<LOG2N, DETREND>(v, x + row * ldx + sa * stride, has_b, stride, win, red, tid);
    fft_r2r<LOG2N>(v, sm, ftw, tid);
    __syncthreads();
#pragma unroll
    for (int r = 0; r < 16; ++r) sm[fft_phys(tid + r * NT)] = v[r];
    __syncthreads();

    const double amp = sqrt(norm);
    // bins k = tid + m*NT, m = 0..7 (k < N/2), plus k = N/2 on thread 0
#pragma unroll
    for (int m = 0; m <= 8; ++m) {
        if (m == 8 && tid != 0) break;
        const int k = tid + m * NT;
        const double2 zk = v[m];
        const double2 zn = sm[fft_phys((N - k) & (N - 1))];
        // X_a = (Z[k] + conj Z[N-k]) / 2 ; X_b = (Z[k] - conj Z[N-k]) / (2i)
        const double2 xa = make_double2(0.5 * (zk.x + zn.x), 0.5 * (zk.y - zn.y));
        const double2 xb = make_double2(0.5 * (zk.y + zn.y), -0.5 * (zk.x - zn.x));
        if (MODE == SPEC_STFT) {
            double2 *o = reinterpret_cast<double2 *>(out);
            o[(sa * rows + row) * NF + k] = make_double2(xa.x * amp, xa.y * amp);
            if (has_b) o[((sa + 1) * rows + row) * NF + k] = make_double2(xb.x * amp, xb.y * amp);
        } else {
            const double f = (k == 0 || k == N / 2) ? norm : 2.0 * norm;
            out[(sa * rows + row) * NF + k] = f * (xa.x * xa.x + xa.y * xa.y);
            if (has_b) out[((sa + 1) * rows + row) * NF + k] = f * (xb.x * xb.x + xb.y * xb.y);
        }
    }
}

// float32-compute variant of spec_segments_kernel (opt-in, the plan's compute
// mode; nfft 512 .. 4096): float64 samples in, float64 / complex128 out, the
// window product, the transform and the untangling in float32.  As in
// welch_pp_c32_kernel the samples are centred in float64 on the pair's first
// sample before they are narrowed and the detrend sums run in float64.
template <int LOG2N, int DETREND, int MODE>
__global__ void __launch_bounds__(oszf::FftCfg<LOG2N>::NT, 3)
spec_segments_c32_kernel(const double *__restrict__ x, int64_t ldx, int64_t rows, int64_t nseg,
                         int64_t stride, const float *__restrict__ win,
                         const float2 *__restrict__ tw, double norm, double *__restrict__ out) {
    using C = oszf::FftCfg<LOG2N>;
    constexpr int N = C::N, NT = C::NT, NF = N / 2 + 1;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float2 *sm = reinterpret_cast<float2 *>(smem_raw);
    __shared__ double red[4 * 32];

    const int tid = threadIdx.x;
    const int64_t row = blockIdx.y;
    const int64_t sa = (int64_t)blockIdx.x * 2;
    const bool has_b = sa + 1 < nseg;
    const oszf::FftTw ftw = oszf::fft_load_tw<LOG2N>(tw, tid);
    const double *xa = x + row * ldx + sa * stride;

    float2 v[16];
    double s[4] = {0.0, 0.0, 0.0, 0.0};
    constexpr double tbar = 0.5 * (N - 1);
    const double c = DETREND != OSZ_DETREND_NONE ? ldg(xa) : 0.0;
#pragma unroll
    for (int r = 0; r < 16; ++r) {
        const int i = tid + r * NT;
        const double da = ldg(xa + i) - c;
        const double db = has_b ? ldg(xa + stride + i) - c : 0.0;
        v[r] = make_float2((float)da, (float)db);
        if (DETREND != OSZ_DETREND_NONE) {
            s[0] += da;
            s[1] += db;
            if (DETREND == OSZ_DETREND_LINEAR) {
                const double tc = (double)i - tbar;
                s[2] = fma(tc, da, s[2]);
                s[3] = fma(tc, db, s[3]);
            }
        }
    }
    if (DETREND == OSZ_DETREND_NONE) {
#pragma unroll
        for (int r = 0; r < 16; ++r) {
            const float w = ldg(win + tid + r * NT);
            v[r].x *= w;
            v[r].y *= w;
        }
    } else {
        block_sum<4, NT>(s, red, tid);
        const float ma = (float)(s[0] / N), mb = (float)(s[1] / N);
        float ka = 0.0f, kb = 0.0f;
        if (DETREND == OSZ_DETREND_LINEAR) {
            constexpr double stt = (double)N * ((double)N * N - 1.0) / 12.0;   // sum (t - tbar)^2
            ka = (float)(s[2] / stt);
            kb = (float)(s[3] / stt);
        }
#pragma unroll
        for (int r = 0; r < 16; ++r) {
            const float tc = (float)(tid + r * NT) - (float)tbar;
            const float w = ldg(win + tid + r * NT);
            v[r].x = (v[r].x - fmaf(ka, tc, ma)) * w;
            v[r].y = (v[r].y - fmaf(kb, tc, mb)) * w;
        }
    }
    oszf::fft_r2r<LOG2N>(v, sm, ftw, tid);
    __syncthreads();
#pragma unroll
    for (int r = 0; r < 16; ++r) sm[oszf::fft_phys(tid + r * NT)] = v[r];
    __syncthreads();

    const float amp = (float)sqrt(norm);
    // bins k = tid + m*NT, m = 0..7 (k < N/2), plus k = N/2 on thread 0
#pragma unroll
    for (int m = 0; m <= 8; ++m) {
        if (m == 8 && tid != 0) break;
        const int k = tid + m * NT;
        const float2 zk = v[m];
        const float2 zn = sm[oszf::fft_phys((N - k) & (N - 1))];
        // X_a = (Z[k] + conj Z[N-k]) / 2 ; X_b = (Z[k] - conj Z[N-k]) / (2i)
        const float2 fa = make_float2(0.5f * (zk.x + zn.x), 0.5f * (zk.y - zn.y));
        const float2 fb = make_float2(0.5f * (zk.y + zn.y), -0.5f * (zk.x - zn.x));
        if (MODE == SPEC_STFT) {
            double2 *o = reinterpret_cast<double2 *>(out);
            o[(sa * rows + row) * NF + k] = make_double2((double)(fa.x * amp), (double)(fa.y * amp));
            if (has_b)
                o[((sa + 1) * rows + row) * NF + k] =
                    make_double2((double)(fb.x * amp), (double)(fb.y * amp));
        } else {
            const double f = (k == 0 || k == N / 2) ? norm : 2.0 * norm;
            out[(sa * rows + row) * NF + k] = f * (double)(fa.x * fa.x + fa.y * fa.y);
            if (has_b)
                out[((sa + 1) * rows + row) * NF + k] = f * (double)(fb.x * fb.x + fb.y * fb.y);
        }
    }
}

// One zero-padded segment per row (periodogram / modified_dft with nfft > n,
// reference numerical.py:688-699): out[r][i] = (x[r][i] - trend_r(i)) * win[i]
// for i < n, 0 for n <= i < nfft.  One CTA per row; the transform itself then
// runs with a unit window and no detrending.
__global__ void __launch_bounds__(256)
spec_prepare_kernel(const double *__restrict__ x, int64_t ldx, int64_t n, int64_t nfft,
                    const double *__restrict__ win, int detrend, double *__restrict__ out,
                    int64_t ldo) {
    __shared__ double red[4 * 32];
    const int tid = threadIdx.x;
    const double *xr = x + (int64_t)blockIdx.x * ldx;
    double *orow = out + (int64_t)blockIdx.x * ldo;
    double s[4] = {0.0, 0.0, 0.0, 0.0};
    const double tbar = 0.5 * (double)(n - 1);
    if (detrend != OSZ_DETREND_NONE) {
        for (int64_t i = tid; i < n; i += 256) {
            const double v = xr[i];
            s[0] += v;
            s[2] = fma((double)i - tbar, v, s[2]);
        }
        block_sum<4, 256>(s, red, tid);
    }
    const double mean = s[0] / (double)n;
    double slope = 0.0;
    if (detrend == OSZ_DETREND_LINEAR && n > 1) {
        const double dn = (double)n;
        slope = s[2] / (dn * (dn * dn - 1.0) / 12.0);
    }
    for (int64_t i = tid; i < nfft; i += 256) {
        double v = 0.0;
        if (i < n) {
            v = xr[i];
            if (detrend != OSZ_DETREND_NONE) v -= fma(slope, (double)i - tbar, mean);
            v *= win[i];
        }
        orow[i] = v;
    }
}

}  // namespace osz

using namespace osz;

struct osz_spec_plan {
    int nfft = 0, stride = 0, detrend = 0, path = 0, log2n = 0;
    int compute = OSZ_COMPUTE_F64;
    double norm = 0.0;
    double *d_win = nullptr;
    double2 *d_tw = nullptr;
    float *d_winf = nullptr;    // float32 compute (osz_spec_plan_set_compute)
    float2 *d_twf = nullptr;
    void *generic = nullptr;    // spectra_generic.cu state
    void *mixed = nullptr;      // spectra_mixed.cu state (shared-memory mixed radix), may be null
};

// generic path (spectra_generic.cu)
int osz_generic_create(void **state, int nfft);
void osz_generic_destroy(void *state);
int osz_generic_exec(void *state, const osz_spec_plan *p, int mode, const double *x, int64_t ldx,
                     int64_t rows, int64_t nseg, double *out, int64_t ldp, cudaStream_t st);
// shared-memory mixed-radix path (spectra_mixed.cu)
int osz_mixed_create(void **state, int nfft);
void osz_mixed_destroy(void *state);
int osz_mixed_can(void *state, int mode);
int osz_mixed_exec(void *state, const osz_spec_plan *p, int mode, const double *x, int64_t ldx,
                   int64_t rows, int64_t nseg, double *out, int64_t ldp, cudaStream_t st);
// accessors used by the generic path
int osz_spec_plan_nfft(const osz_spec_plan *p) { return p->nfft; }
int osz_spec_plan_stride(const osz_spec_plan *p) { return p->stride; }
int osz_spec_plan_detrend(const osz_spec_plan *p) { return p->detrend; }
double osz_spec_plan_norm(const osz_spec_plan *p) { return p->norm; }
const double *osz_spec_plan_window(const osz_spec_plan *p) { return p->d_win; }

template <int LOG2N, int DETREND>
static int launch_welch(const osz_spec_plan *p, const double *x, int64_t ldx, int64_t rows,
                        int64_t nseg, double *psd, int64_t ldp, cudaStream_t st) {
    using C = FftCfg<LOG2N>;
    constexpr int SMEM = C::SMEM_BYTES + C::N * 8;   // FFT exchange + accumulators
    OSZ_CUDA(cudaFuncSetAttribute(welch_accum_kernel<LOG2N, DETREND>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    const int64_t npairs = (nseg + 1) / 2;
    // enough CTAs for ~4 waves of 2 CTAs/SM, but at least 4 pairs per CTA so the
    // register accumulators amortise the atomics
    const int64_t target = (int64_t)sm_count() * 8;
    int64_t per_row = (target + rows - 1) / rows;
    if (per_row < 1) per_row = 1;
    int64_t ppc = (npairs + per_row - 1) / per_row;
    if (ppc < 4) ppc = 4;
    const int64_t gx = (npairs + ppc - 1) / ppc;
    dim3 grid((unsigned)gx, (unsigned)rows);
    welch_accum_kernel<LOG2N, DETREND><<<grid, C::NT, SMEM, st>>>(
        x, ldx, nseg, p->stride, p->d_win, p->d_tw, p->norm, psd, ldp, ppc);
    OSZ_LAUNCHED("welch_accum_kernel");
    return OSZ_OK;
}

template <int LOG2N, int DETREND>
static int launch_welch_pp(const osz_spec_plan *p, const double *x, int64_t ldx, int64_t rows,
                           int64_t nseg, double *psd, int64_t ldp, cudaStream_t st) {
    using C = FftCfg<LOG2N>;
    using L = WelchPP<LOG2N>;
    static const int token = [] {
        // measured (256 x 1e6): nfft 4096 206 G samples/s with the token, 200 without;
        // nfft 1024 (four CTAs per SM) 209 with, 227 without
        const char *e = getenv("OSZ_WELCH64_TOKEN");
        return e ? atoi(e) : (LOG2N >= 11 ? 1 : 0);
    }();
    auto kern = token ? welch_pp_kernel<LOG2N, DETREND, true> : welch_pp_kernel<LOG2N, DETREND, false>;
    OSZ_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::SMEM));
    const int64_t npairs = (nseg + 1) / 2;
    const int64_t nwork = npairs * rows;
    int64_t grid = (nwork + 1) / 2;
    if (grid > (int64_t)sm_count() * L::CTAS_PER_SM) grid = (int64_t)sm_count() * L::CTAS_PER_SM;
    const int64_t per_group = (nwork + 2 * grid - 1) / (2 * grid);
    static const int lag = [] {
        const char *e = getenv("OSZ_WELCH_LAG");
        return e ? atoi(e) : 0;
    }();
    kern<<<(unsigned)grid, 2 * C::NT, L::SMEM, st>>>(
        x, ldx, nseg, p->stride, p->d_win, p->d_tw, p->norm, psd, ldp, npairs, nwork, per_group,
        lag, 0);
    OSZ_LAUNCHED("welch_pp_kernel");
    return OSZ_OK;
}

template <int LOG2N, int DETREND, typename TIO>
static int launch_welch_pp_c32(const osz_spec_plan *p, const TIO *x, int64_t ldx, int64_t rows,
                               int64_t nseg, double *psd, int64_t ldp, cudaStream_t st) {
    using C = oszf::FftCfg<LOG2N>;
    using L = WelchPPC32<LOG2N>;
    static_assert(L::SMEM * L::CTAS_PER_SM + 1024 * L::CTAS_PER_SM <= 228 * 1024,
                  "float32-compute Welch: shared memory");
    static const int token = [] {
        const char *e = getenv("OSZ_WELCH32_TOKEN");
        return e ? atoi(e) : 0;
    }();
    auto kern = token ? welch_pp_c32_kernel<LOG2N, DETREND, true, TIO>
                      : welch_pp_c32_kernel<LOG2N, DETREND, false, TIO>;
    OSZ_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::SMEM));
    const int64_t npairs = (nseg + 1) / 2;
    const int64_t nwork = npairs * rows;
    int64_t grid = (nwork + 1) / 2;
    if (grid > (int64_t)sm_count() * L::CTAS_PER_SM) grid = (int64_t)sm_count() * L::CTAS_PER_SM;
    const int64_t per_group = (nwork + 2 * grid - 1) / (2 * grid);
    static const int lag = [] {
        const char *e = getenv("OSZ_WELCH_LAG");
        return e ? atoi(e) : 0;
    }();
    kern<<<(unsigned)grid, 2 * C::NT, L::SMEM, st>>>(
        x, ldx, nseg, p->stride, p->d_winf, p->d_twf, p->norm, psd, ldp, npairs, nwork, per_group,
        lag, 0);
    OSZ_LAUNCHED("welch_pp_c32_kernel");
    return OSZ_OK;
}

template <int LOG2N, int DETREND, int MODE>
static int launch_segments(const osz_spec_plan *p, const double *x, int64_t ldx, int64_t rows,
                           int64_t nseg, double *out, cudaStream_t st) {
    using C = FftCfg<LOG2N>;
    OSZ_CUDA(cudaFuncSetAttribute(spec_segments_kernel<LOG2N, DETREND, MODE>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
    dim3 grid((unsigned)((nseg + 1) / 2), (unsigned)rows);
    spec_segments_kernel<LOG2N, DETREND, MODE><<<grid, C::NT, C::SMEM_BYTES, st>>>(
        x, ldx, rows, nseg, p->stride, p->d_win, p->d_tw, p->norm, out);
    OSZ_LAUNCHED("spec_segments_kernel");
    return OSZ_OK;
}

template <int LOG2N, int DETREND, int MODE>
static int launch_segments_c32(const osz_spec_plan *p, const double *x, int64_t ldx, int64_t rows,
                               int64_t nseg, double *out, cudaStream_t st) {
    using C = oszf::FftCfg<LOG2N>;
    OSZ_CUDA(cudaFuncSetAttribute(spec_segments_c32_kernel<LOG2N, DETREND, MODE>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
    dim3 grid((unsigned)((nseg + 1) / 2), (unsigned)rows);
    spec_segments_c32_kernel<LOG2N, DETREND, MODE><<<grid, C::NT, C::SMEM_BYTES, st>>>(
        x, ldx, rows, nseg, p->stride, p->d_winf, p->d_twf, p->norm, out);
    OSZ_LAUNCHED("spec_segments_c32_kernel");
    return OSZ_OK;
}

template <int LOG2N, int DETREND>
static int dispatch_mode(const osz_spec_plan *p, int mode, const double *x, int64_t ldx,
                         int64_t rows, int64_t nseg, double *out, int64_t ldp, cudaStream_t st) {
    if (mode == SPEC_ACCUM) {
        static const int pp = [] {
            const char *e = getenv("OSZ_WELCH_PP");
            return e ? atoi(e) : 1;
        }();
        if constexpr (LOG2N >= 9 && LOG2N <= 12) {
            if (p->compute == OSZ_COMPUTE_F32 && p->d_winf && p->d_twf)
                return launch_welch_pp_c32<LOG2N, DETREND, double>(p, x, ldx, rows, nseg, out, ldp,
                                                                   st);
            if (pp) return launch_welch_pp<LOG2N, DETREND>(p, x, ldx, rows, nseg, out, ldp, st);
        }
        return launch_welch<LOG2N, DETREND>(p, x, ldx, rows, nseg, out, ldp, st);
    }
    if constexpr (LOG2N >= 9 && LOG2N <= 12) {
        if (p->compute == OSZ_COMPUTE_F32 && p->d_winf && p->d_twf) {
            if (mode == SPEC_PGRAM)
                return launch_segments_c32<LOG2N, DETREND, SPEC_PGRAM>(p, x, ldx, rows, nseg, out, st);
            return launch_segments_c32<LOG2N, DETREND, SPEC_STFT>(p, x, ldx, rows, nseg, out, st);
        }
    }
    if (mode == SPEC_PGRAM)
        return launch_segments<LOG2N, DETREND, SPEC_PGRAM>(p, x, ldx, rows, nseg, out, st);
    return launch_segments<LOG2N, DETREND, SPEC_STFT>(p, x, ldx, rows, nseg, out, st);
}

template <int LOG2N>
static int dispatch_detrend(const osz_spec_plan *p, int mode, const double *x, int64_t ldx,
                            int64_t rows, int64_t nseg, double *out, int64_t ldp,
                            cudaStream_t st) {
    switch (p->detrend) {
        case OSZ_DETREND_NONE:
            return dispatch_mode<LOG2N, OSZ_DETREND_NONE>(p, mode, x, ldx, rows, nseg, out, ldp, st);
        case OSZ_DETREND_CONSTANT:
            return dispatch_mode<LOG2N, OSZ_DETREND_CONSTANT>(p, mode, x, ldx, rows, nseg, out, ldp,
                                                              st);
        default:
            return dispatch_mode<LOG2N, OSZ_DETREND_LINEAR>(p, mode, x, ldx, rows, nseg, out, ldp,
                                                            st);
    }
}

static int spec_exec(const osz_spec_plan *p, int mode, const double *x, int64_t ldx, int64_t rows,
                     int64_t nseg, double *out, int64_t ldp, void *stream) {
    if (!p || !x || !out) return fail(OSZ_ERR_ARG, "spectra exec: null argument");
    if (rows <= 0 || nseg <= 0) return OSZ_OK;
    if (rows > 65535) return fail(OSZ_ERR_UNSUPPORTED, "spectra: more than 65535 rows per call");
    cudaStream_t st = as_stream(stream);
    if (p->path == 2) {
        static const int use_mixed = [] {
            const char *e = getenv("OSZ_SPEC_MIXED");
            return e ? atoi(e) : 1;
        }();
        if (use_mixed && p->mixed && osz_mixed_can(p->mixed, mode))
            return osz_mixed_exec(p->mixed, p, mode, x, ldx, rows, nseg, out, ldp, st);
        return osz_generic_exec(p->generic, p, mode, x, ldx, rows, nseg, out, ldp, st);
    }
    switch (p->log2n) {
        case 8: return dispatch_detrend<8>(p, mode, x, ldx, rows, nseg, out, ldp, st);
        case 9: return dispatch_detrend<9>(p, mode, x, ldx, rows, nseg, out, ldp, st);
        case 10: return dispatch_detrend<10>(p, mode, x, ldx, rows, nseg, out, ldp, st);
        case 11: return dispatch_detrend<11>(p, mode, x, ldx, rows, nseg, out, ldp, st);
        case 12: return dispatch_detrend<12>(p, mode, x, ldx, rows, nseg, out, ldp, st);
        case 13: return dispatch_detrend<13>(p, mode, x, ldx, rows, nseg, out, ldp, st);
    }
    return fail(OSZ_ERR_UNSUPPORTED, "spectra: unsupported nfft");
}

// float32 I/O: float samples in, float64 sums out; plans whose Welch accumulation runs in
// float32 arithmetic (osz_spec_plan_set_compute, power-of-two nfft 512 ... 4096).
template <int LOG2N>
static int welch_f32_detrend(const osz_spec_plan *p, const float *x, int64_t ldx, int64_t rows,
                             int64_t nseg, double *psd, int64_t ldp, cudaStream_t st) {
    switch (p->detrend) {
        case OSZ_DETREND_NONE:
            return launch_welch_pp_c32<LOG2N, OSZ_DETREND_NONE, float>(p, x, ldx, rows, nseg, psd,
                                                                       ldp, st);
        case OSZ_DETREND_CONSTANT:
            return launch_welch_pp_c32<LOG2N, OSZ_DETREND_CONSTANT, float>(p, x, ldx, rows, nseg,
                                                                           psd, ldp, st);
        default:
            return launch_welch_pp_c32<LOG2N, OSZ_DETREND_LINEAR, float>(p, x, ldx, rows, nseg, psd,
                                                                         ldp, st);
    }
}

extern "C" {

int osz_spec_plan_create(osz_spec_plan **out, int nfft, int stride, const double *window,
                         int detrend, double norm) {
    if (!out || !window || nfft < 2 || stride < 1 || stride > nfft)
        return fail(OSZ_ERR_ARG, "osz_spec_plan_create: bad arguments");
    if (detrend < OSZ_DETREND_NONE || detrend > OSZ_DETREND_LINEAR)
        return fail(OSZ_ERR_ARG, "osz_spec_plan_create: unknown detrend");
    osz_spec_plan *p = new osz_spec_plan();
    p->nfft = nfft;
    p->stride = stride;
    p->detrend = detrend;
    p->norm = norm;
    int log2n = 0;
    while ((1 << log2n) < nfft) ++log2n;
    const bool pow2 = (1 << log2n) == nfft && log2n >= 8 && log2n <= 13;
    bool ok = cudaMalloc(&p->d_win, (size_t)nfft * 8) == cudaSuccess &&
              cudaMemcpy(p->d_win, window, (size_t)nfft * 8, cudaMemcpyHostToDevice) == cudaSuccess;
    if (ok && pow2) {
        p->path = 1;
        p->log2n = log2n;
        std::vector<double> tw = make_fft_twiddles(log2n);
        ok = cudaMalloc(&p->d_tw, tw.size() * 8) == cudaSuccess &&
             cudaMemcpy(p->d_tw, tw.data(), tw.size() * 8, cudaMemcpyHostToDevice) == cudaSuccess;
    } else if (ok) {
        p->path = 2;
        int rc = osz_generic_create(&p->generic, nfft);
        if (rc == OSZ_OK) rc = osz_mixed_create(&p->mixed, nfft);
        if (rc != OSZ_OK) {
            osz_spec_plan_destroy(p);
            return rc;
        }
    }
    if (!ok) {
        osz_spec_plan_destroy(p);
        return fail(OSZ_ERR_CUDA, "osz_spec_plan_create: device upload failed");
    }
    *out = p;
    return OSZ_OK;
}

int osz_spec_plan_destroy(osz_spec_plan *p) {
    if (!p) return OSZ_OK;
    cudaFree(p->d_win);
    cudaFree(p->d_tw);
    cudaFree(p->d_winf);
    cudaFree(p->d_twf);
    if (p->generic) osz_generic_destroy(p->generic);
    if (p->mixed) osz_mixed_destroy(p->mixed);
    delete p;
    return OSZ_OK;
}

int osz_spec_plan_path(const osz_spec_plan *p) { return p ? p->path : 0; }

int osz_spec_plan_set_compute(osz_spec_plan *p, int compute, const double *window) {
    if (!p || (compute != OSZ_COMPUTE_F64 && compute != OSZ_COMPUTE_F32))
        return fail(OSZ_ERR_ARG, "osz_spec_plan_set_compute: bad arguments");
    if (compute == OSZ_COMPUTE_F64) {
        p->compute = OSZ_COMPUTE_F64;
        return OSZ_OK;
    }
    // float32 arithmetic exists for the Welch accumulation at nfft = 512 .. 4096;
    // every other plan keeps computing in float64
    if (p->path != 1 || p->log2n < 9 || p->log2n > 12) return OSZ_OK;
    if (!p->d_winf) {
        if (!window) return fail(OSZ_ERR_ARG, "osz_spec_plan_set_compute: window needed");
        std::vector<float> wf((size_t)p->nfft);
        for (int i = 0; i < p->nfft; ++i) wf[(size_t)i] = (float)window[i];
        std::vector<float> twf = oszf::make_fft_twiddles(p->log2n);
        const bool ok =
            cudaMalloc(&p->d_winf, wf.size() * 4) == cudaSuccess &&
            cudaMemcpy(p->d_winf, wf.data(), wf.size() * 4, cudaMemcpyHostToDevice) == cudaSuccess &&
            cudaMalloc(&p->d_twf, twf.size() * 4) == cudaSuccess &&
            cudaMemcpy(p->d_twf, twf.data(), twf.size() * 4, cudaMemcpyHostToDevice) == cudaSuccess;
        if (!ok) return fail(OSZ_ERR_CUDA, "osz_spec_plan_set_compute: device upload failed");
    }
    p->compute = OSZ_COMPUTE_F32;
    return OSZ_OK;
}
int osz_spec_plan_compute(const osz_spec_plan *p) { return p ? p->compute : 0; }

int osz_welch_accum_f64(const osz_spec_plan *p, const double *x, int64_t ldx, int64_t rows,
                        int64_t nseg, double *psd_sum, int64_t ldp, void *stream) {
    return spec_exec(p, SPEC_ACCUM, x, ldx, rows, nseg, psd_sum, ldp, stream);
}
int osz_welch_accum_f32(const osz_spec_plan *p, const float *x, int64_t ldx, int64_t rows,
                        int64_t nseg, double *psd_sum, int64_t ldp, void *stream) {
    if (!p || !x || !psd_sum) return fail(OSZ_ERR_ARG, "osz_welch_accum_f32: null argument");
    if (rows <= 0 || nseg <= 0) return OSZ_OK;
    if (p->path != 1 || p->compute != OSZ_COMPUTE_F32 || !p->d_winf || !p->d_twf)
        return fail(OSZ_ERR_UNSUPPORTED, "osz_welch_accum_f32: needs a plan set to float32 "
                                         "arithmetic (power-of-two nfft 512 ... 4096)");
    cudaStream_t st = as_stream(stream);
    switch (p->log2n) {
        case 9: return welch_f32_detrend<9>(p, x, ldx, rows, nseg, psd_sum, ldp, st);
        case 10: return welch_f32_detrend<10>(p, x, ldx, rows, nseg, psd_sum, ldp, st);
        case 11: return welch_f32_detrend<11>(p, x, ldx, rows, nseg, psd_sum, ldp, st);
        case 12: return welch_f32_detrend<12>(p, x, ldx, rows, nseg, psd_sum, ldp, st);
    }
    return fail(OSZ_ERR_UNSUPPORTED, "osz_welch_accum_f32: unsupported nfft");
}
int osz_periodogram_f64(const osz_spec_plan *p, const double *x, int64_t ldx, int64_t rows,
                        int64_t nseg, double *out, void *stream) {
    return spec_exec(p, SPEC_PGRAM, x, ldx, rows, nseg, out, 0, stream);
}
int osz_spec_prepare_f64(const double *x, int64_t ldx, int64_t rows, int64_t n, int64_t nfft,
                         const double *window_dev, int detrend, double *out, int64_t ldo,
                         void *stream) {
    if (!x || !window_dev || !out || n < 1 || nfft < n)
        return fail(OSZ_ERR_ARG, "osz_spec_prepare_f64: bad arguments");
    if (detrend < OSZ_DETREND_NONE || detrend > OSZ_DETREND_LINEAR)
        return fail(OSZ_ERR_ARG, "osz_spec_prepare_f64: unknown detrend");
    if (rows <= 0) return OSZ_OK;
    spec_prepare_kernel<<<(unsigned)rows, 256, 0, as_stream(stream)>>>(x, ldx, n, nfft, window_dev,
                                                                      detrend, out, ldo);
    OSZ_LAUNCHED("spec_prepare_kernel");
    return OSZ_OK;
}
int osz_stft_f64(const osz_spec_plan *p, const double *x, int64_t ldx, int64_t rows, int64_t nseg,
                 double *out, void *stream) {
    return spec_exec(p, SPEC_STFT, x, ldx, rows, nseg, out, 0, stream);
}

}  // extern "C"
