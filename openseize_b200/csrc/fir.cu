// FIR filtering of halo'd chunks: the per-chunk arithmetic of nm.oaconvolve
// (reference core/numerical.py:229-269) as two sm_100a kernels.
//
//   fir_direct_kernel   short filters: chunk-plus-halo tile staged into shared
//                       memory by a 1-D TMA bulk copy, taps in shared memory,
//                       15 outputs per thread on a register sliding window
//                       (one LDS per 15 DFMA).
//   fir_fft_kernel      long filters: overlap-save blocks of N = 4096 / 8192.
//                       Two consecutive blocks of a row ride as the real and
//                       imaginary part of ONE complex transform (real taps =>
//                       the two convolutions never mix), forward FFT, multiply
//                       by H, inverse FFT, all register-to-register through
//                       the shared-memory FFT of fft_core.cuh.  Blocks are
//                       independent CTAs -- the scatter-free dual of the
//                       reference's overlap-add (same sums, no carried tail).
//
// Both compute  y[r][i] = sum_k taps[k] * x[r][i + K-1-k]  (header contract).
#include <stdlib.h>

#include <type_traits>
#include <vector>

#include "common.cuh"
#include "fft_core.cuh"

namespace osz {

// ------------------------------------------------------------------ direct
constexpr int FIR_NT = 128;   // threads per CTA
constexpr int FIR_R = 15;     // outputs per thread (odd: conflict-free LDS.64 at stride R)
constexpr int FIR_TILE = FIR_NT * FIR_R;
constexpr int FIR_DIRECT_MAX_TAPS = 1024;
constexpr int FIR_FFT_MAX_TAPS = 2049;     // one N = 8192 block: (N-K+1)/N >= 3/4

__global__ void __launch_bounds__(FIR_NT, 4)
fir_direct_kernel(const double *__restrict__ x, int64_t ldx, int64_t n_out, int ntaps, int kpad,
                  const double *__restrict__ taps_rev /* kpad, zero padded */, double *__restrict__ y,
                  int64_t ldy) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem_raw);
    double *gs = reinterpret_cast<double *>(smem_raw + 16);          // kpad taps
    const int xs_off = (16 + kpad * 8 + 127) & ~127;
    double *xs = reinterpret_cast<double *>(smem_raw + xs_off);      // tile + halo
    const int xs_len = FIR_TILE + kpad + FIR_R + 2;

    const int tid = threadIdx.x;
    const int64_t row = blockIdx.y;
    const int64_t tile0 = (int64_t)blockIdx.x * FIR_TILE;
    const double *xr = x + row * ldx + tile0;
    const int64_t span_left = n_out + ntaps - 1 - tile0;             // readable from xr
    const int mis = (int)((reinterpret_cast<uintptr_t>(xr) >> 3) & 1);
    int64_t need = FIR_TILE + ntaps - 1;
    if (need > span_left) need = span_left;
    const int cnt = (int)((need + mis + 1) & ~(int64_t)1);           // even element count

    if (tid == 0) {
        mbar_init(bar, 1);
        fence_barrier_init();
    }
    __syncthreads();
    if (tid == 0) {
        mbar_expect_tx(bar, (uint32_t)cnt * 8u);
        tma_load_1d(xs, xr - mis, (uint32_t)cnt * 8u, bar);
    }
    // meanwhile: taps, and zeros behind the TMA region (padded taps multiply them)
    for (int i = tid; i < kpad; i += FIR_NT) gs[i] = taps_rev[i];
    for (int i = cnt + tid; i < xs_len; i += FIR_NT) xs[i] = 0.0;
    __syncthreads();
    mbar_wait(bar, 0);

    const double *xt = xs + tid * FIR_R + mis;
    double acc[FIR_R], w[FIR_R];
#pragma unroll
    for (int r = 0; r < FIR_R; ++r) {
        acc[r] = 0.0;
        w[r] = xt[r];
    }
    for (int j0 = 0; j0 < kpad; j0 += FIR_R) {
#pragma unroll
        for (int u = 0; u < FIR_R; ++u) {
            const double g = gs[j0 + u];
#pragma unroll
            for (int r = 0; r < FIR_R; ++r) acc[r] = fma(g, w[(r + u) % FIR_R], acc[r]);
            w[u] = xt[j0 + u + FIR_R];
        }
    }
    double *yr = y + row * ldy + tile0 + tid * FIR_R;
    const int64_t left = n_out - (tile0 + tid * FIR_R);
#pragma unroll
    for (int r = 0; r < FIR_R; ++r)
        if (r < left) st_stream(yr + r, acc[r]);
}

// --------------------------------------------------------------------- FFT
// MINB = CTAs per SM the register budget is sized for (2: 128 regs, 3: 80 regs).
template <int LOG2N, int MINB>
__global__ void __launch_bounds__(FftCfg<LOG2N>::NT, MINB)
fir_fft_kernel(const double *__restrict__ x, int64_t ldx, int64_t n_out, int ntaps,
               const double2 *__restrict__ H /* N, scaled 1/N */, const double2 *__restrict__ tw,
               double *__restrict__ y, int64_t ldy, int accumulate) {
    using C = FftCfg<LOG2N>;
    constexpr int N = C::N, NT = C::NT;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double2 *sm = reinterpret_cast<double2 *>(smem_raw);

    const int tid = threadIdx.x;
    const int64_t row = blockIdx.y;
    const int64_t step = N - ntaps + 1;
    const int64_t span = n_out + ntaps - 1;
    const int64_t base_a = (int64_t)blockIdx.x * 2 * step;     // block b starts base_a + step
    const double *pa = x + row * ldx + base_a + tid;
    const double *pb = pa + step;
    const int64_t la = span - base_a, lb = la - step;
    const int lim_a = (int)(la > N ? N : la), lim_b = (int)(lb > N ? N : (lb < 0 ? 0 : lb));
    const FftTw ftw = fft_load_tw<LOG2N>(tw, tid);

    double2 v[16];
#pragma unroll
    for (int r = 0; r < 16; ++r) {
        const int i = tid + r * NT;
        v[r].x = i < lim_a ? ld_stream(pa + r * NT) : 0.0;
        v[r].y = i < lim_b ? ld_stream(pb + r * NT) : 0.0;
    }
    fft_r2r<LOG2N>(v, sm, ftw, tid);
    const double2 *hp = H + tid;
#pragma unroll
    for (int r = 0; r < 16; ++r) {
        const double2 p = cmul(v[r], ldg(hp + r * NT));
        v[r] = make_double2(p.y, p.x);  // swap: inverse transform via the forward kernel
    }
    // The inverse transform re-reads its base twiddles with volatile loads:
    // otherwise ptxas keeps the 30 twiddle powers of the forward transform
    // alive for reuse and spills ~300 B per thread (profiles/r01_ncu_summary.md).
    FftTw ftw2;
    {
        const volatile double2 *vt = tw;
        ftw2.m1.x = vt[C::OFF_M1 + (tid & 15)].x;
        ftw2.m1.y = vt[C::OFF_M1 + (tid & 15)].y;
        ftw2.m2.x = vt[C::OFF_M2 + (tid & (16 * C::R1 - 1))].x;
        ftw2.m2.y = vt[C::OFF_M2 + (tid & (16 * C::R1 - 1))].y;
        ftw2.last.x = vt[C::OFF_L + tid].x;
        ftw2.last.y = vt[C::OFF_L + tid].y;
    }
    fft_r2r<LOG2N>(v, sm, ftw2, tid);
    const int k1 = ntaps - 1;
    double *ya = y + row * ldy + base_a + tid - k1;
    double *yb = ya + step;
    const int64_t oa = n_out - base_a + k1, ob = oa - step;
    const int out_a = (int)(oa > N ? N : oa), out_b = (int)(ob > N ? N : (ob < 0 ? 0 : ob));
#pragma unroll
    for (int r = 0; r < 16; ++r) {
        const int i = tid + r * NT;
        if (i >= k1) {
            // after the swap back: real part = v.y, imaginary part = v.x
            if (accumulate) {   // a later partition of a filter longer than one block allows
                if (i < out_a) ya[r * NT] += v[r].y;
                if (i < out_b) yb[r * NT] += v[r].x;
            } else {
                if (i < out_a) st_stream(ya + r * NT, v[r].y);
                if (i < out_b) st_stream(yb + r * NT, v[r].x);
            }
        }
    }
}

// Ping-pong variant (N = 4096): one persistent CTA per SM, two thread groups,
// each transforming its own pair of blocks; the groups alternate on the FP64
// pipe (fft_core.cuh SyncPingPong) so one group's butterflies overlap the other
// group's shared-memory exchanges and global traffic.  Work item w =
// row * npairs + pair; group g of CTA c takes w = 2c + g, then strides by
// 2 * gridDim.x.  Every group runs the same number of iterations (the token
// protocol needs matching acquire/release counts); items past the end compute
// on zeros and store nothing.
//
// The item's input span (both blocks: step + N contiguous samples) is fetched
// by ONE 1-D TMA bulk copy into the group's own exchange buffer, issued as
// soon as the previous item's last pass has read its inputs, so the DRAM
// latency hides behind that item's last butterflies and its stores.  H is kept
// as a half table in shared memory (real taps: H[N-k] = conj H[k]).
template <int LOG2N>
struct FirPP {
    using C = FftCfg<LOG2N>;
    static constexpr int NH = C::N / 2 + 1;
    static constexpr int OFF_H = C::TW_TOTAL * 16;
    static constexpr int OFF_BAR = OFF_H + NH * 16;
    static constexpr int OFF_X = (OFF_BAR + 16 + 127) & ~127;
    static constexpr int SMEM = OFF_X + 2 * C::SMEM_BYTES;
};

template <int LOG2N, bool ACC, int POLICY, bool TOKEN = true>
__global__ void __launch_bounds__(2 * FftCfg<LOG2N>::NT, 1)
fir_fft_pp_kernel(const double *__restrict__ x, int64_t ldx, int64_t n_out, int ntaps,
                  const double2 *__restrict__ H, const double2 *__restrict__ tw,
                  double *__restrict__ y, int64_t ldy, int64_t npairs, int64_t nwork, int iters,
                  int lag, int zero) {
    using C = FftCfg<LOG2N>;
    using L = FirPP<LOG2N>;
    using Sync = typename std::conditional<TOKEN, SyncPingPong<LOG2N>, SyncGroups<LOG2N>>::type;
    constexpr int N = C::N, NT = C::NT;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int g = threadIdx.x / NT;
    const int tid = threadIdx.x - g * NT;
    double2 *tw_sm = reinterpret_cast<double2 *>(smem_raw);
    double2 *h_sm = reinterpret_cast<double2 *>(smem_raw + L::OFF_H);
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem_raw + L::OFF_BAR) + g;
    double2 *sm = reinterpret_cast<double2 *>(smem_raw + L::OFF_X) + g * C::SMEM_ELEMS;
    double *sx = reinterpret_cast<double *>(sm);      // the exchange buffer as the item's input span

    for (int i = threadIdx.x; i < C::TW_TOTAL; i += 2 * NT) tw_sm[i] = ldg(tw + i);
    for (int i = threadIdx.x; i < L::NH; i += 2 * NT) h_sm[i] = ldg(H + i);
    if (tid == 0) {
        mbar_init(bar, 1);
        fence_barrier_init();
    }
    __syncthreads();
    const Sync sync{g, tid, zero, tw_sm};

    const int64_t step = N - ntaps + 1;
    const int64_t span = n_out + ntaps - 1;
    const int k1 = ntaps - 1;
    const int64_t dw = 2 * (int64_t)gridDim.x;
    const int64_t drow = dw / npairs, dpair = dw - drow * npairs;
    int64_t w = 2 * (int64_t)blockIdx.x + g;
    int64_t row = w / npairs, pair = w - row * npairs;

    // thread 0 of the group: fetch the span of item (w_, row_, pair_) into sx
    auto issue = [&](int64_t w_, int64_t row_, int64_t pair_) {
        if (w_ >= nwork) return;
        const double *src = x + row_ * ldx + pair_ * 2 * step;
        const int64_t left = span - pair_ * 2 * step;           // samples readable from src
        tma_fetch_span(sx, src, left < step + N ? left : step + N, bar);
    };
    if (tid == 0) issue(w, row, pair);
    sync.prime();
    // Group 1 runs two turns behind group 0, so that a group's longest stretch
    // without FP64 work (stores of one item + input of the next) coincides with
    // the other group's longest FP64 phase (last pass + H + first inverse pass).
    if (g == 1)
        for (int i = 0; i < lag; ++i) sync.idle_turn();

    for (int it = 0; it < iters; ++it) {
        const bool live = w < nwork;
        const int64_t base_a = pair * 2 * step;
        const int64_t left = span - base_a;
        int64_t wn = w + dw, rown = row + drow, pairn = pair + dpair;
        if (pairn >= npairs) {
            pairn -= npairs;
            ++rown;
        }

        double2 v[16];
        if (live) {
            const double *xa = sx + span_mis(x + row * ldx + base_a) + tid, *xb = xa + step;
            while (!mbar_try_wait(bar, it & 1)) {
            }
            if (left >= step + N) {
#pragma unroll
                for (int r = 0; r < 16; ++r) v[r] = make_double2(xa[r * NT], xb[r * NT]);
            } else {
                const int lim_a = (int)(left > N ? N : left);
                const int lim_b = (int)(left - step > N ? N : (left < step ? 0 : left - step));
#pragma unroll
                for (int r = 0; r < 16; ++r) {
                    const int i = tid + r * NT;
                    v[r].x = i < lim_a ? xa[r * NT] : 0.0;
                    v[r].y = i < lim_b ? xb[r * NT] : 0.0;
                }
            }
        } else {
#pragma unroll
            for (int r = 0; r < 16; ++r) v[r] = make_double2(0.0, 0.0);
        }
        if (POLICY & 1) sync.acquire_first(v);
        bfly<16>(v);
        if (POLICY & 1) sync.release(v);
        fft_r2r_tail<LOG2N, Sync, (POLICY & 2) == 0>(v, sm, tid, sync);
        // still holding the token: multiply by H and run the inverse transform's
        // first (twiddle-free) pass
        {
            const double2 *hlo = h_sm + tid, *hhi = h_sm + 8 * NT - tid;
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                const double2 p = cmul(v[r], hlo[r * NT]);
                v[r] = make_double2(p.y, p.x);  // swap: inverse transform via the forward kernel
            }
#pragma unroll
            for (int r = 8; r < 16; ++r) {
                const double2 h = hhi[(8 - r) * NT];    // H[k] = conj H[N - k], N - k = (16-r) NT - tid
                const double2 a = v[r];
                v[r] = make_double2(fma(a.y, h.x, -a.x * h.y), fma(a.x, h.x, a.y * h.y));
            }
        }
        bfly<16>(v);
        if (POLICY & 2) sync.release(v);
        fft_r2r_tail<LOG2N, Sync, true>(v, sm, tid, sync, [&]() {
            sync.group();                 // every thread of the group has read its inputs
            if (tid == 0) {
                fence_proxy_async();
                issue(wn, rown, pairn);
            }
        });

        if (live) {
            double *ya = y + row * ldy + base_a + tid - k1;
            double *yb = ya + step;
            const int64_t oa = n_out - base_a + k1, ob = oa - step;
            // after the swap back: real part = v.y, imaginary part = v.x
            if (ob >= N) {
#pragma unroll
                for (int r = 0; r < 16; ++r) {
                    if (tid + r * NT >= k1) {
                        if (ACC) {   // a later partition of a filter longer than one block
                            ya[r * NT] += v[r].y;
                            yb[r * NT] += v[r].x;
                        } else {
                            st_stream(ya + r * NT, v[r].y);
                            st_stream(yb + r * NT, v[r].x);
                        }
                    }
                }
            } else {
                const int out_a = (int)(oa > N ? N : oa), out_b = (int)(ob < 0 ? 0 : ob);
#pragma unroll
                for (int r = 0; r < 16; ++r) {
                    const int i = tid + r * NT;
                    if (i >= k1) {
                        if (ACC) {
                            if (i < out_a) ya[r * NT] += v[r].y;
                            if (i < out_b) yb[r * NT] += v[r].x;
                        } else {
                            if (i < out_a) st_stream(ya + r * NT, v[r].y);
                            if (i < out_b) st_stream(yb + r * NT, v[r].x);
                        }
                    }
                }
            }
        }
        w = wn;
        row = rown;
        pair = pairn;
    }
    if (g == 0)
        for (int i = 0; i < lag; ++i) sync.idle_turn();
    sync.drain();
}

// float32-compute variant of the ping-pong kernel (opt-in, OSZ_FIR_FFT_F32):
// float64 samples in and out, both transforms and the H product in float32
// (namespace oszf of fft_core.cuh).  The FP32 pipe issues at twice the FP64 rate
// and the exchange buffers halve, so the kernel becomes HBM bound at the
// float64 roofline's 16 bytes per sample; the result differs from the float64
// path by ~1e-6 of the output peak (north_star's float32 tolerance is 1e-5).
// The input span has its own staging buffer here (the float2 exchange buffer is
// too small for it), so the next item's TMA copy is issued as soon as this
// item's samples are in registers and has the whole item to land.
template <int LOG2N>
struct FirPPC32 {
    using C = oszf::FftCfg<LOG2N>;
    static constexpr int NH = C::N / 2 + 1;
    static constexpr int OFF_H = C::TW_TOTAL * 8;
    static constexpr int OFF_BAR = (OFF_H + NH * 8 + 15) & ~15;
    static constexpr int OFF_X = (OFF_BAR + 16 + 127) & ~127;
    static constexpr int STAGE_BYTES = 2 * C::N * 8;                  // step + N doubles at most
    static constexpr int GROUP_BYTES = C::SMEM_BYTES + STAGE_BYTES;
    static constexpr int SMEM = OFF_X + 2 * GROUP_BYTES;
};

// TIO: the samples' type in memory -- double (float64 samples in and out, the default
// of the float32 ARITHMETIC mode) or float (the float32 I/O mode: 8 bytes per sample of
// HBM traffic instead of 16).
template <int LOG2N, bool ACC, bool TOKEN, typename TIO>
__global__ void __launch_bounds__(2 * oszf::FftCfg<LOG2N>::NT, 1)
fir_fft_pp_c32_kernel(const TIO *__restrict__ x, int64_t ldx, int64_t n_out, int ntaps,
                      const float2 *__restrict__ H, const float2 *__restrict__ tw,
                      TIO *__restrict__ y, int64_t ldy, int64_t npairs, int64_t nwork,
                      int iters, int lag, int zero) {
    using C = oszf::FftCfg<LOG2N>;
    using L = FirPPC32<LOG2N>;
    using Sync = typename std::conditional<TOKEN, oszf::SyncPingPong<LOG2N>,
                                           oszf::SyncGroups<LOG2N>>::type;
    constexpr int N = C::N, NT = C::NT;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int g = threadIdx.x / NT;
    const int tid = threadIdx.x - g * NT;
    float2 *tw_sm = reinterpret_cast<float2 *>(smem_raw);
    float2 *h_sm = reinterpret_cast<float2 *>(smem_raw + L::OFF_H);
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem_raw + L::OFF_BAR) + g;
    float2 *sm = reinterpret_cast<float2 *>(smem_raw + L::OFF_X + g * L::GROUP_BYTES);
    TIO *sx = reinterpret_cast<TIO *>(smem_raw + L::OFF_X + g * L::GROUP_BYTES + C::SMEM_BYTES);

    for (int i = threadIdx.x; i < C::TW_TOTAL; i += 2 * NT) tw_sm[i] = ldg(tw + i);
    for (int i = threadIdx.x; i < L::NH; i += 2 * NT) h_sm[i] = ldg(H + i);
    if (tid == 0) {
        mbar_init(bar, 1);
        fence_barrier_init();
    }
    __syncthreads();
    const Sync sync{g, tid, zero, tw_sm};

    const int64_t step = N - ntaps + 1;
    const int64_t span = n_out + ntaps - 1;
    const int k1 = ntaps - 1;
    const int64_t dw = 2 * (int64_t)gridDim.x;
    const int64_t drow = dw / npairs, dpair = dw - drow * npairs;
    int64_t w = 2 * (int64_t)blockIdx.x + g;
    int64_t row = w / npairs, pair = w - row * npairs;

    auto issue = [&](int64_t w_, int64_t row_, int64_t pair_) {
        if (w_ >= nwork) return;
        const TIO *src = x + row_ * ldx + pair_ * 2 * step;
        const int64_t left = span - pair_ * 2 * step;
        tma_fetch_span(sx, src, left < step + N ? left : step + N, bar);
    };
    if (tid == 0) issue(w, row, pair);
    sync.prime();
    if (g == 1)
        for (int i = 0; i < lag; ++i) sync.idle_turn();

    for (int it = 0; it < iters; ++it) {
        const bool live = w < nwork;
        const int64_t base_a = pair * 2 * step;
        const int64_t left = span - base_a;
        int64_t wn = w + dw, rown = row + drow, pairn = pair + dpair;
        if (pairn >= npairs) {
            pairn -= npairs;
            ++rown;
        }
        float2 v[16];
        if (live) {
            const TIO *xa = sx + span_mis(x + row * ldx + base_a) + tid, *xb = xa + step;
            while (!mbar_try_wait(bar, it & 1)) {
            }
            if (left >= step + N) {
#pragma unroll
                for (int r = 0; r < 16; ++r)
                    v[r] = make_float2((float)xa[r * NT], (float)xb[r * NT]);
            } else {
                const int lim_a = (int)(left > N ? N : left);
                const int lim_b = (int)(left - step > N ? N : (left < step ? 0 : left - step));
#pragma unroll
                for (int r = 0; r < 16; ++r) {
                    const int i = tid + r * NT;
                    v[r].x = i < lim_a ? (float)xa[r * NT] : 0.0f;
                    v[r].y = i < lim_b ? (float)xb[r * NT] : 0.0f;
                }
            }
        } else {
#pragma unroll
            for (int r = 0; r < 16; ++r) v[r] = make_float2(0.0f, 0.0f);
        }
        // the staging buffer is free once every thread of the group holds its samples
        sync.group();
        if (tid == 0) {
            fence_proxy_async();
            issue(wn, rown, pairn);
        }
        sync.acquire_first(v);
        oszf::bfly<16>(v);
        sync.release(v);
        oszf::fft_r2r_tail<LOG2N, Sync, false>(v, sm, tid, sync);
        {
            const float2 *hlo = h_sm + tid, *hhi = h_sm + 8 * NT - tid;
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                const float2 p = oszf::cmul(v[r], hlo[r * NT]);
                v[r] = make_float2(p.y, p.x);
            }
#pragma unroll
            for (int r = 8; r < 16; ++r) {
                const float2 h = hhi[(8 - r) * NT];
                const float2 a = v[r];
                v[r] = make_float2(fmaf(a.y, h.x, -a.x * h.y), fmaf(a.x, h.x, a.y * h.y));
            }
        }
        oszf::bfly<16>(v);
        sync.release(v);
        oszf::fft_r2r_tail<LOG2N, Sync, true>(v, sm, tid, sync);

        if (live) {
            TIO *ya = y + row * ldy + base_a + tid - k1;
            TIO *yb = ya + step;
            const int64_t oa = n_out - base_a + k1, ob = oa - step;
            const int out_a = (int)(oa > N ? N : oa), out_b = (int)(ob < 0 ? 0 : (ob > N ? N : ob));
#pragma unroll
            for (int r = 0; r < 16; ++r) {
                const int i = tid + r * NT;
                if (i >= k1) {
                    if (ACC) {
                        if (i < out_a) ya[r * NT] += (TIO)v[r].y;
                        if (i < out_b) yb[r * NT] += (TIO)v[r].x;
                    } else {
                        if (i < out_a) st_stream(ya + r * NT, (TIO)v[r].y);
                        if (i < out_b) st_stream(yb + r * NT, (TIO)v[r].x);
                    }
                }
            }
        }
        w = wn;
        row = rown;
        pair = pairn;
    }
    if (g == 0)
        for (int i = 0; i < lag; ++i) sync.idle_turn();
    sync.drain();
}

}  // namespace osz

using namespace osz;

struct osz_fir_plan {
    int ntaps = 0;
    int algo = OSZ_FIR_DIRECT;
    int kpad = 0;
    int log2n = 0;
    // filters longer than FIR_FFT_MAX_TAPS: partitions of the taps, each a plan of
    // its own; exec adds their valid convolutions (y = sum_p h_p * x shifted)
    std::vector<osz_fir_plan *> parts;
    std::vector<int> part_end;      // exclusive end tap index of each partition
    double *d_taps_rev = nullptr;   // direct
    double2 *d_H = nullptr;         // fft
    double2 *d_tw = nullptr;
    float2 *d_Hf = nullptr;         // float32-compute fft (OSZ_FIR_FFT_F32)
    float2 *d_twf = nullptr;
};

template <int LOG2N, int MINB>
static int launch_fir_fft(const osz_fir_plan *p, const double *x, int64_t ldx, int64_t rows,
                          int64_t n_out, double *y, int64_t ldy, cudaStream_t st, int accumulate) {
    using C = FftCfg<LOG2N>;
    OSZ_CUDA(cudaFuncSetAttribute(fir_fft_kernel<LOG2N, MINB>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
    const int64_t step = C::N - p->ntaps + 1;
    const int64_t nblocks = (n_out + step - 1) / step;
    const int64_t npairs = (nblocks + 1) / 2;
    if (rows > 65535) return fail(OSZ_ERR_UNSUPPORTED, "fir: more than 65535 rows per call");
    dim3 grid((unsigned)npairs, (unsigned)rows);
    fir_fft_kernel<LOG2N, MINB><<<grid, C::NT, C::SMEM_BYTES, st>>>(
        x, ldx, n_out, p->ntaps, p->d_H, p->d_tw, y, ldy, accumulate);
    OSZ_LAUNCHED("fir_fft_kernel");
    return OSZ_OK;
}

template <int LOG2N, bool ACC, int POLICY>
static int launch_fir_fft_pp(const osz_fir_plan *p, const double *x, int64_t ldx, int64_t rows,
                             int64_t n_out, double *y, int64_t ldy, cudaStream_t st) {
    using C = FftCfg<LOG2N>;
    constexpr int SMEM = FirPP<LOG2N>::SMEM;
    // (the token-free build measured slower in float64: 223 vs 257 G samples/s at 113 taps,
    // 201 vs 222 at 671 -- profiles/r01_ncu_summary.md D -- and is not instantiated)
    auto kern = fir_fft_pp_kernel<LOG2N, ACC, POLICY, true>;
    OSZ_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    const int64_t step = C::N - p->ntaps + 1;
    const int64_t nblocks = (n_out + step - 1) / step;
    const int64_t npairs = (nblocks + 1) / 2;
    const int64_t nwork = npairs * rows;
    int64_t grid = (nwork + 1) / 2;
    if (grid > sm_count()) grid = sm_count();
    const int iters = (int)((nwork + 2 * grid - 1) / (2 * grid));
    static const int lag = [] {
        const char *e = getenv("OSZ_PP_LAG");
        return e ? atoi(e) : 2;
    }();
    kern<<<(unsigned)grid, 2 * C::NT, SMEM, st>>>(
        x, ldx, n_out, p->ntaps, p->d_H, p->d_tw, y, ldy, npairs, nwork, iters, lag, 0);
    OSZ_LAUNCHED("fir_fft_pp_kernel");
    return OSZ_OK;
}

template <int LOG2N, bool ACC, typename TIO>
static int launch_fir_fft_pp_c32(const osz_fir_plan *p, const TIO *x, int64_t ldx, int64_t rows,
                                 int64_t n_out, TIO *y, int64_t ldy, cudaStream_t st) {
    using C = oszf::FftCfg<LOG2N>;
    constexpr int SMEM = FirPPC32<LOG2N>::SMEM;
    static_assert(SMEM <= 227 * 1024, "float32-compute FIR: shared memory");
    static const int token = [] {
        const char *e = getenv("OSZ_FIR32_TOKEN");
        return e ? atoi(e) : 0;
    }();
    auto kern = token ? fir_fft_pp_c32_kernel<LOG2N, ACC, true, TIO>
                      : fir_fft_pp_c32_kernel<LOG2N, ACC, false, TIO>;
    OSZ_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    const int64_t step = C::N - p->ntaps + 1;
    const int64_t nblocks = (n_out + step - 1) / step;
    const int64_t npairs = (nblocks + 1) / 2;
    const int64_t nwork = npairs * rows;
    int64_t grid = (nwork + 1) / 2;
    if (grid > sm_count()) grid = sm_count();
    const int iters = (int)((nwork + 2 * grid - 1) / (2 * grid));
    kern<<<(unsigned)grid, 2 * C::NT, SMEM, st>>>(
        x, ldx, n_out, p->ntaps, p->d_Hf, p->d_twf, y, ldy, npairs, nwork, iters, 2, 0);
    OSZ_LAUNCHED("fir_fft_pp_c32_kernel");
    return OSZ_OK;
}

template <int LOG2N, bool ACC>
static int launch_fir_fft_pp_policy(const osz_fir_plan *p, const double *x, int64_t ldx,
                                    int64_t rows, int64_t n_out, double *y, int64_t ldy,
                                    cudaStream_t st) {
    static const int policy = [] {
        const char *e = getenv("OSZ_PP_POLICY");
        return e ? atoi(e) : 3;
    }();
    switch (policy) {
        case 0: return launch_fir_fft_pp<LOG2N, ACC, 0>(p, x, ldx, rows, n_out, y, ldy, st);
        case 1: return launch_fir_fft_pp<LOG2N, ACC, 1>(p, x, ldx, rows, n_out, y, ldy, st);
        case 2: return launch_fir_fft_pp<LOG2N, ACC, 2>(p, x, ldx, rows, n_out, y, ldy, st);
        default: return launch_fir_fft_pp<LOG2N, ACC, 3>(p, x, ldx, rows, n_out, y, ldy, st);
    }
}

extern "C" {

int osz_fir_plan_create(osz_fir_plan **out, const double *taps, int ntaps, int algo) {
    if (!out || !taps || ntaps < 1) return fail(OSZ_ERR_ARG, "osz_fir_plan_create: bad arguments");
    bool c32 = false;
    if (algo == OSZ_FIR_FFT_F32) {
        // float32 compute exists for one N = 4096 block; longer filters stay in float64
        c32 = ntaps <= 1025;
        algo = OSZ_FIR_FFT;
    }
    if (algo == OSZ_FIR_AUTO) algo = ntaps <= 24 ? OSZ_FIR_DIRECT : OSZ_FIR_FFT;
    if (algo == OSZ_FIR_DIRECT && ntaps > FIR_DIRECT_MAX_TAPS) algo = OSZ_FIR_FFT;
    osz_fir_plan *p = new osz_fir_plan();
    p->ntaps = ntaps;
    p->algo = algo;
    if (algo == OSZ_FIR_DIRECT) {
        p->kpad = ((ntaps + FIR_R - 1) / FIR_R) * FIR_R;
        std::vector<double> rev(p->kpad, 0.0);
        for (int j = 0; j < ntaps; ++j) rev[j] = taps[ntaps - 1 - j];
        if (cudaMalloc(&p->d_taps_rev, rev.size() * 8) != cudaSuccess ||
            cudaMemcpy(p->d_taps_rev, rev.data(), rev.size() * 8, cudaMemcpyHostToDevice) !=
                cudaSuccess) {
            osz_fir_plan_destroy(p);
            return fail(OSZ_ERR_CUDA, "osz_fir_plan_create: device upload failed");
        }
    } else if (algo == OSZ_FIR_FFT) {
        // block size: keep the overlap-save efficiency (N-K+1)/N at or above 3/4
        if (ntaps <= 1025)
            p->log2n = 12;
        else if (ntaps <= 2049)
            p->log2n = 13;
        else {
            // partition the taps; every part is an ordinary N = 8192 plan
            for (int t0 = 0; t0 < ntaps; t0 += FIR_FFT_MAX_TAPS) {
                const int t1 = t0 + FIR_FFT_MAX_TAPS < ntaps ? t0 + FIR_FFT_MAX_TAPS : ntaps;
                osz_fir_plan *part = nullptr;
                int rc = osz_fir_plan_create(&part, taps + t0, t1 - t0, OSZ_FIR_FFT);
                if (rc != OSZ_OK) {
                    osz_fir_plan_destroy(p);
                    return rc;
                }
                p->parts.push_back(part);
                p->part_end.push_back(t1);
            }
            *out = p;
            return OSZ_OK;
        }
        const int N = 1 << p->log2n;
        // H[k] = (1/N) sum_j taps[j] exp(-2 pi i j k / N), long double accumulation
        std::vector<long double> cs(N), sn(N);
        const long double two_pi = 6.283185307179586476925286766559005768L;
        for (int m = 0; m < N; ++m) {
            cs[m] = cosl(two_pi * m / N);
            sn[m] = sinl(two_pi * m / N);
        }
        std::vector<double> H(2 * (size_t)N);
        for (int k = 0; k < N; ++k) {
            long double re = 0, im = 0;
            for (int j = 0; j < ntaps; ++j) {
                const int m = (int)(((long long)j * k) & (N - 1));
                re += taps[j] * cs[m];
                im -= taps[j] * sn[m];
            }
            H[2 * k] = (double)(re / N);
            H[2 * k + 1] = (double)(im / N);
        }
        if (c32 && p->log2n == 12) {
            std::vector<float> Hf(H.size()), twf = oszf::make_fft_twiddles(12);
            for (size_t i = 0; i < H.size(); ++i) Hf[i] = (float)H[i];
            if (cudaMalloc(&p->d_Hf, Hf.size() * 4) != cudaSuccess ||
                cudaMemcpy(p->d_Hf, Hf.data(), Hf.size() * 4, cudaMemcpyHostToDevice) !=
                    cudaSuccess ||
                cudaMalloc(&p->d_twf, twf.size() * 4) != cudaSuccess ||
                cudaMemcpy(p->d_twf, twf.data(), twf.size() * 4, cudaMemcpyHostToDevice) !=
                    cudaSuccess) {
                osz_fir_plan_destroy(p);
                return fail(OSZ_ERR_CUDA, "osz_fir_plan_create: device upload failed");
            }
            p->algo = OSZ_FIR_FFT_F32;
        }
        std::vector<double> tw = make_fft_twiddles(p->log2n);
        if (cudaMalloc(&p->d_H, H.size() * 8) != cudaSuccess ||
            cudaMemcpy(p->d_H, H.data(), H.size() * 8, cudaMemcpyHostToDevice) != cudaSuccess ||
            cudaMalloc(&p->d_tw, tw.size() * 8) != cudaSuccess ||
            cudaMemcpy(p->d_tw, tw.data(), tw.size() * 8, cudaMemcpyHostToDevice) != cudaSuccess) {
            osz_fir_plan_destroy(p);
            return fail(OSZ_ERR_CUDA, "osz_fir_plan_create: device upload failed");
        }
    } else {
        delete p;
        return fail(OSZ_ERR_ARG, "osz_fir_plan_create: unknown algorithm");
    }
    *out = p;
    return OSZ_OK;
}

int osz_fir_plan_destroy(osz_fir_plan *p) {
    if (!p) return OSZ_OK;
    for (osz_fir_plan *part : p->parts) osz_fir_plan_destroy(part);
    cudaFree(p->d_taps_rev);
    cudaFree(p->d_H);
    cudaFree(p->d_tw);
    cudaFree(p->d_Hf);
    cudaFree(p->d_twf);
    delete p;
    return OSZ_OK;
}

int osz_fir_plan_algo(const osz_fir_plan *p) { return p ? p->algo : 0; }

static int fir_exec(const osz_fir_plan *p, const double *x, int64_t ldx, int64_t rows,
                    int64_t n_out, double *y, int64_t ldy, cudaStream_t st, int accumulate);

int osz_fir_exec_f64(const osz_fir_plan *p, const double *x, int64_t ldx, int64_t rows,
                     int64_t n_out, double *y, int64_t ldy, void *stream) {
    if (!p || !x || !y) return fail(OSZ_ERR_ARG, "osz_fir_exec_f64: null argument");
    if (rows <= 0 || n_out <= 0) return OSZ_OK;
    cudaStream_t st = as_stream(stream);
    if (!p->parts.empty()) {
        // taps [t0, t1) of the filter see the span shifted by ntaps - t1
        for (size_t i = 0; i < p->parts.size(); ++i) {
            int rc = fir_exec(p->parts[i], x + (p->ntaps - p->part_end[i]), ldx, rows, n_out, y,
                              ldy, st, i > 0);
            if (rc != OSZ_OK) return rc;
        }
        return OSZ_OK;
    }
    return fir_exec(p, x, ldx, rows, n_out, y, ldy, st, 0);
}

// float32 I/O: float samples in and out, the float32-arithmetic overlap-save kernel
// (plans created with OSZ_FIR_FFT_F32, 25 ... 1025 taps).
int osz_fir_exec_f32(const osz_fir_plan *p, const float *x, int64_t ldx, int64_t rows,
                     int64_t n_out, float *y, int64_t ldy, void *stream) {
    if (!p || !x || !y) return fail(OSZ_ERR_ARG, "osz_fir_exec_f32: null argument");
    if (rows <= 0 || n_out <= 0) return OSZ_OK;
    if (p->algo != OSZ_FIR_FFT_F32 || p->log2n != 12 || !p->parts.empty())
        return fail(OSZ_ERR_UNSUPPORTED, "osz_fir_exec_f32: needs a float32-arithmetic FFT plan "
                                         "(OSZ_FIR_FFT_F32, at most 1025 taps)");
    return launch_fir_fft_pp_c32<12, false, float>(p, x, ldx, rows, n_out, y, ldy,
                                                   as_stream(stream));
}

}  // extern "C"

static int fir_exec(const osz_fir_plan *p, const double *x, int64_t ldx, int64_t rows,
                    int64_t n_out, double *y, int64_t ldy, cudaStream_t st, int accumulate) {
    if (p->algo == OSZ_FIR_DIRECT) {
        const int xs_off = (16 + p->kpad * 8 + 127) & ~127;
        const int smem = xs_off + (FIR_TILE + p->kpad + FIR_R + 2) * 8;
        OSZ_CUDA(cudaFuncSetAttribute(fir_direct_kernel,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
        if (rows > 65535) return fail(OSZ_ERR_UNSUPPORTED, "fir: more than 65535 rows per call");
        dim3 grid((unsigned)((n_out + FIR_TILE - 1) / FIR_TILE), (unsigned)rows);
        fir_direct_kernel<<<grid, FIR_NT, smem, st>>>(x, ldx, n_out, p->ntaps, p->kpad,
                                                     p->d_taps_rev, y, ldy);
        OSZ_LAUNCHED("fir_direct_kernel");
        return OSZ_OK;
    }
    if (p->algo == OSZ_FIR_FFT_F32 && p->log2n == 12)
        return accumulate
                   ? launch_fir_fft_pp_c32<12, true, double>(p, x, ldx, rows, n_out, y, ldy, st)
                   : launch_fir_fft_pp_c32<12, false, double>(p, x, ldx, rows, n_out, y, ldy, st);
    if (p->log2n == 12) {
        // Two CTAs per SM at 128 registers beat three at 80 (104 B of spills):
        // 217 vs 184 G samples/s at 113 taps (profiles/r01_kernel_bench.md).
        // OSZ_FIR_MINB=3 selects the other build for re-measurement.
        static const int minb = [] {
            const char *e = getenv("OSZ_FIR_MINB");
            return e ? atoi(e) : 2;
        }();
        static const int pp = [] {
            const char *e = getenv("OSZ_FIR_PP");
            return e ? atoi(e) : 1;
        }();
        if (pp)
            return accumulate
                       ? launch_fir_fft_pp_policy<12, true>(p, x, ldx, rows, n_out, y, ldy, st)
                       : launch_fir_fft_pp_policy<12, false>(p, x, ldx, rows, n_out, y, ldy, st);
        if (minb == 3) return launch_fir_fft<12, 3>(p, x, ldx, rows, n_out, y, ldy, st, accumulate);
        return launch_fir_fft<12, 2>(p, x, ldx, rows, n_out, y, ldy, st, accumulate);
    }
    return launch_fir_fft<13, 1>(p, x, ldx, rows, n_out, y, ldy, st, accumulate);
}
