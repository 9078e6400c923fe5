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
#include <stdlib.h>

#include <vector>

#include "common.cuh"
#include "sos_core.cuh"
#include "ufd_mma.cuh"

namespace osz {

constexpr int UFD_NT = 256;
constexpr int UFD_NW = UFD_NT / 32;

template <int R>
__device__ __forceinline__ int ufd_phys(int m) {
    return m + m / R;
}

// R outputs per lane; tile = 32*R outputs per CTA.  T = double, or float for the
// opt-in float32 arithmetic (float64 samples in and out: they are narrowed when the
// tile is staged, taps, windows and partial sums are float32, the result is widened
// when it is stored; half the shared memory, twice the FMA rate).
// TIO: the samples' type in memory -- double, or float for the float32 I/O mode.
template <int R, typename T, typename TIO = double>
__global__ void __launch_bounds__(UFD_NT, 2)
upfirdn_dec_kernel(const TIO *__restrict__ x, int64_t ldx, int64_t x_first, int64_t x_len,
                   int64_t out_first, int64_t n_out, int K, int M, int Q /* taps per phase */,
                   int ldm /* smem row length, odd */, int half,
                   const T *__restrict__ gphase /* [M][Q] reversed taps by phase */,
                   TIO *__restrict__ y, int64_t ldy) {
    constexpr int TO = 32 * R;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *smem = reinterpret_cast<T *>(smem_raw);
    T *xs = smem;                            // [M][ldm]
    T *gs = smem + (size_t)M * ldm;          // [M][Q]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t row = blockIdx.y;
    const int64_t o0 = (int64_t)blockIdx.x * TO;              // tile-relative to out_first
    // global index of the first input sample of this tile
    const int64_t in0 = (out_first + o0) * M + half - (K - 1);
    const TIO *xr = x + row * ldx;

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
        const TIO *src = xr + rel0 + tid;
        for (int e0 = tid; e0 < n_in; e0 += U * UFD_NT, src += U * UFD_NT) {
            TIO val[U];
            if (interior && e0 + (U - 1) * UFD_NT < n_in) {
#pragma unroll
                for (int u = 0; u < U; ++u) val[u] = ld_stream(src + u * UFD_NT);
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    xs[p * ldm + ufd_phys<R>(m)] = (T)val[u];
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
                    val[u] = (e < n_in && g >= 0 && g < x_len) ? ld_stream(xr + g) : (TIO)0;
                }
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    if (e0 + u * UFD_NT < n_in) xs[p * ldm + ufd_phys<R>(m)] = (T)val[u];
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

    T acc[R];
#pragma unroll
    for (int r = 0; r < R; ++r) acc[r] = (T)0;

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
        const T *gp = gs + p * Q;
        const unsigned uq = (unsigned)q0;
        const unsigned rem = uq % R;
        const T *pa = xs + p * ldm + lane * (R + 1) + (uq + uq / R);   // T(q0), crossing not passed
        const T *pb = pa + 1;                                          // crossing passed
        bool c[R];
#pragma unroll
        for (int u = 0; u < R; ++u) c[u] = rem + u >= R;
        T w[R];
#pragma unroll
        for (int r = 0; r < R; ++r) w[r] = (c[r] ? pb : pa)[r];
        int q = q0;
        for (; q + R <= q1; q += R) {
#pragma unroll
            for (int u = 0; u < R; ++u) {
                const T g = gp[q + u];
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
                const T g = gp[q + u];
#pragma unroll
                for (int r = 0; r < R; ++r) acc[r] = fma(g, w[(r + u) % R], acc[r]);
                w[u] = (c[u] ? pb : pa)[u + R + 1];
            }
        }
        s = p * Q + q1;
    }
    __syncthreads();   // tile no longer needed: reuse it for the cross-warp sum
    T *red = smem;   // [NW][32 lanes][R + 1]: the +1 keeps the stride-R stores conflict free
    constexpr int LDR = 32 * (R + 1);
#pragma unroll
    for (int r = 0; r < R; ++r) red[warp * LDR + lane * (R + 1) + r] = acc[r];
    __syncthreads();
    for (int o = tid; o < TO; o += UFD_NT) {
        const int idx = (o / R) * (R + 1) + (o % R);
        T sum = (T)0;
#pragma unroll
        for (int w8 = 0; w8 < UFD_NW; ++w8) sum += red[w8 * LDR + idx];
        if (o0 + o < n_out) st_stream(y + row * ldy + o0 + o, (TIO)sum);
    }
}

// ---- double-buffered decimator -------------------------------------------------
// Same arithmetic as upfirdn_dec_kernel, reorganised so the FP64 pipe never
// waits for staging (the first kernel alternated a memory-bound staging phase
// and an FP64-bound FIR phase per tile, and two co-resident CTAs drift into
// lock step: FP64 pipe 29-45 %, profiles/r01_ncu_summary.md):
//   * one persistent 512-thread CTA per SM walks tiles of 32*R outputs;
//   * the next tile is scattered phase-major into the second shared-memory
//     buffer by 8-byte cp.async (LDGSTS, zero-fill outside the supplied window)
//     while the 16 warps run the FIR of the current tile; every thread's
//     scatter offsets are the same for every tile and live in registers;
//   * the (phase, tap) work is split between warps in blocks of R taps, so a
//     run always starts on a window boundary (no per-load pointer select) and
//     taps are read two at a time (LDS.128, broadcast).
// Instruction mix of the inner loop: 64 DFMA per 4 + 8 loads.
constexpr int UFD2_NT = 512;
constexpr int UFD2_NW = UFD2_NT / 32;
constexpr int UFD2_MAXE = 26;            // staged elements per thread and tile (register table)
// Taps of the double-buffered kernel live in constant memory, one slot per plan
// (first fit, released when the plan is destroyed; plans that find no room use
// the first kernel).  [M][QB*R] reversed taps by phase, zero padded.
constexpr int UFD2_CONST_DOUBLES = 7168;   // 56 KB of the 64 KB constant bank
__constant__ double c_ufd_taps[UFD2_CONST_DOUBLES];

template <int R>
__global__ void __launch_bounds__(UFD2_NT, 1)
upfirdn_dec2_kernel(const double *__restrict__ x, int64_t ldx, int64_t x_first, int64_t x_len,
                    int64_t out_first, int64_t n_out, int K, int M, int QB /* R-tap blocks per phase */,
                    int ldm /* doubles per phase row, odd */, int half,
                    const double *__restrict__ gphase /* [M][QB*R] reversed taps, zero padded */,
                    double *__restrict__ y, int64_t ldy, int tiles_per_row, int64_t ntiles,
                    int tap_slot /* offset into c_ufd_taps */) {
    constexpr int TO = 32 * R;
    constexpr int LDR = 32 * (R + 1);
    extern __shared__ __align__(16) double smem[];
    const int Qp = QB * R;
    double *red = smem;                                       // [NW][32 lanes][R + 1]
    double *bufs = red + UFD2_NW * LDR;                       // 2 x [M][ldm]
    const int tile_elems = M * ldm;

    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);   // warp-uniform for the compiler
    const int n_in = (TO + Qp) * M;

    // byte offset (within a buffer) of staged element tid + k * NT: phase e % M,
    // decimated index e / M
    uint32_t off[UFD2_MAXE];
#pragma unroll
    for (int k = 0; k < UFD2_MAXE; ++k) {
        const int e = tid + k * UFD2_NT;
        const int pe = e % M, me = e / M;
        off[k] = (uint32_t)(pe * ldm + ufd_phys<R>(me)) * 8u;
    }
    const uint32_t sbase = smem_u32(bufs);

    auto stage = [&](int64_t t, int which) {
        if (t >= ntiles) return;
        const int64_t row = t / tiles_per_row;
        const int64_t o0 = (t - row * tiles_per_row) * TO;
        const int64_t rel0 = (out_first + o0) * M + half - (K - 1) - x_first;
        const double *xr = x + row * ldx;
        const uint32_t dst = sbase + (uint32_t)which * (uint32_t)tile_elems * 8u;
        if (rel0 >= 0 && rel0 + n_in <= x_len) {
            const double *src = xr + rel0 + tid;
#pragma unroll
            for (int k = 0; k < UFD2_MAXE; ++k)
                if (tid + k * UFD2_NT < n_in) cp_async8_zfill(dst + off[k], src + k * UFD2_NT, true);
        } else {
#pragma unroll
            for (int k = 0; k < UFD2_MAXE; ++k) {
                const int e = tid + k * UFD2_NT;
                const int64_t gi = rel0 + e;
                const bool ok = gi >= 0 && gi < x_len;
                if (e < n_in) cp_async8_zfill(dst + off[k], ok ? xr + gi : xr, ok);
            }
        }
    };

    // Taps come from constant memory through the
    // UNIFORM datapath: a DFMA whose multiplier is a uniform register reads two
    // vector operands instead of three, and three-vector-operand DFMA streams
    // sustain only ~46 of 64 lanes/clk/SM on B200 (tools/microbench/fp64_rate.cu).
    // every warp runs the same number of (phase, R-tap block) units, so the
    // control flow below is warp- AND block-uniform
    const int units = M * QB;
    const int upw = (units + UFD2_NW - 1) / UFD2_NW;
    const int ubase = __shfl_sync(0xffffffffu, warp * upw, 0);
    const int p_first = ubase / QB, b_first = ubase - p_first * QB;

    stage(blockIdx.x, 0);
    cp_async_commit();
    int parity = 0;
    for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x, parity ^= 1) {
        const double *xs = bufs + (size_t)parity * tile_elems;
        stage(t + gridDim.x, parity ^ 1);
        cp_async_commit();
        cp_async_wait<1>();
        __syncthreads();

        double acc[R];
#pragma unroll
        for (int r = 0; r < R; ++r) acc[r] = 0.0;
        // Two register windows alternate: while the taps of one block run against
        // (wa | wb), the block after it is already loading into the other set.
        // Window element d of block b, lane l, phase p: xs[p*ldm + (l + b)*(R+1) + d].
        {
            double wa[R], wb[R];
            int pcur = p_first, bcur = b_first;
            const int gidx = tap_slot + ubase * R;        // warp-uniform tap index
            for (int i = 0; i < upw; i += 2) {
                if (ubase + i >= units) break;
                {
                    const double *px = xs + pcur * ldm + (lane + bcur) * (R + 1);
                    if (i == 0 || bcur == 0) {
#pragma unroll
                        for (int r = 0; r < R; ++r) wa[r] = px[r];
                    }
#pragma unroll
                    for (int r = 0; r < R; ++r) wb[r] = px[R + 1 + r];
#pragma unroll
                    for (int u = 0; u < R; ++u) {
                        const double g = c_ufd_taps[gidx + i * R + u];
#pragma unroll
                        for (int r = 0; r < R; ++r)
                            acc[r] = fma(g, r + u < R ? wa[(r + u) % R] : wb[(r + u) % R], acc[r]);
                    }
                    if (++bcur == QB) {
                        bcur = 0;
                        ++pcur;
                    }
                }
                if (i + 1 >= upw || ubase + i + 1 >= units) break;
                {
                    const double *px = xs + pcur * ldm + (lane + bcur) * (R + 1);
                    if (bcur == 0) {
#pragma unroll
                        for (int r = 0; r < R; ++r) wb[r] = px[r];
                    }
#pragma unroll
                    for (int r = 0; r < R; ++r) wa[r] = px[R + 1 + r];
#pragma unroll
                    for (int u = 0; u < R; ++u) {
                        const double g = c_ufd_taps[gidx + (i + 1) * R + u];
#pragma unroll
                        for (int r = 0; r < R; ++r)
                            acc[r] = fma(g, r + u < R ? wb[(r + u) % R] : wa[(r + u) % R], acc[r]);
                    }
                    if (++bcur == QB) {
                        bcur = 0;
                        ++pcur;
                    }
                }
            }
        }
#pragma unroll
        for (int r = 0; r < R; ++r) red[warp * LDR + lane * (R + 1) + r] = acc[r];
        __syncthreads();
        const int64_t row = t / tiles_per_row;
        const int64_t o0 = (t - row * tiles_per_row) * TO;
        for (int o = tid; o < TO; o += UFD2_NT) {
            const int idx = (o / R) * (R + 1) + (o % R);
            double sum = 0.0;
#pragma unroll
            for (int w8 = 0; w8 < UFD2_NW; ++w8) sum += red[w8 * LDR + idx];
            if (o0 + o < n_out) st_stream(y + row * ldy + o0 + o, sum);
        }
        // (the next iteration's barrier orders these reads before red is rewritten;
        //  this tile's buffer is refilled only after that barrier as well)
    }
    cp_async_wait<0>();
}

// ---- decimator on the FP64 tensor cores (DMMA) -----------------------------------
// north_star allows tensor cores for an FIR "only if a Toeplitz-GEMM formulation is
// shown to win".  Measured on B200 (tools/microbench/dmma_rate.cu, profiles/
// r02_microbench.md): mma.sync.m8n8k4.f64 sustains 63 of 64 FMA lanes/clk/SM with
// both operands streamed from shared memory, a DFMA FIR loop 43-50 (its third
// register operand and one LDS per 8 DFMA are what cap it).  So the decimating
// filter is evaluated as a banded Toeplitz product:
//
//   C[i][t] = out[J0 + i*S + 8*tau + t]              (8 x 8 tile, i = segment of S outputs)
//           = sum_w  x[base + (i*S + 8*tau)*M + w] * g[w - t*M],      0 <= w < K + 7*M
//
// with the sum over w taken phase by phase, four taps of one phase per MMA:
//   w = p + (4*s + q)*M     ->   A[i][q] = Xp[i*S + 8*tau + 4*s + q],  Xp[n] = x[base + p + n*M]
//                                B[q][t] = gp[p][4*s + q - t]          (zero outside the taps)
// A depends on (tau, s) only through 2*tau + s, so a warp that owns tiles tau and
// tau+1 loads ONE new A fragment per k-step and reuses it two steps later; B is one
// LDS.64 per k-step shared by both tiles: 2 LDS per 2 DMMA (512 FMA).
// The 8 rows of A are 8 segments of ONE channel's time axis: the input tile stays
// time-contiguous in shared memory (no phase scatter), segment i at pitch
// P = S*M + pad doubles, pad chosen so that the 16 lanes of a half warp
// (g*P + q*M) hit 16 different banks.  Useful fraction of the MMA work:
// K / sum_p 4*ceil((Q_p + 7) / 4)  (86 % at 1231 taps, M = 25; 70 % at 561 taps).
template <int WT, int KS, int RB>
__global__ void __launch_bounds__(WT *KS * 32 + 32, 1)
upfirdn_mma_kernel(const __grid_constant__ UfdMmaGeom gm, const double *__restrict__ x, int64_t ldx,
                   int64_t x_first, int64_t x_len, int64_t out_first, int64_t n_out,
                   const double *__restrict__ gpad /* [M][ldq]: 7 zeros, taps of the phase, zeros */,
                   double *__restrict__ y, int64_t ldy, int tiles_per_row, int64_t ntiles) {
    // One persistent CTA per SM walks (row, tile) items with two shared-memory tile
    // buffers.  WT * KS MMA warps (each: four 8 x 8 output tiles of its segment quarter,
    // 1 / KS of the k-steps) never meet at a CTA-wide barrier: they hand their partial
    // sums to ONE epilogue warp through named barriers and go on to the next tile, whose
    // data is already in the other buffer.  The epilogue warp sums the KS partials,
    // stores the tile's 8 S outputs and refills the buffer just released two tiles
    // ahead: interior tiles by 1-D TMA bulk copies (one per segment, completing on an
    // mbarrier), tiles that touch the edge of the supplied window by 8-byte cp.async
    // with zero fill.
    constexpr int NMMA = WT * KS * 32;                       // MMA threads
    constexpr int NT = NMMA + 32;
    extern __shared__ __align__(16) double smem_mma[];
    __shared__ __align__(8) uint64_t bars[2];
    const int M = gm.M, SM = gm.SM, P = gm.P, S = gm.S;
    const int nseg = (gm.total_len + SM - 1) / SM;
    const int tile_elems = nseg * P + 2;                     // + 2: room for the alignment shift
    double *bufs = smem_mma;                                 // 2 x tile_elems
    double *gs = bufs + 2 * (size_t)tile_elems;              // M * ldq
    double *red = gs + (size_t)M * gm.ldq;                   // RB x KS x 8 S partial sums

    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const bool epilogue = warp == WT * KS;
    const uint32_t sbase = smem_u32(bufs);
    const bool tma_pitch = (P & 1) == 0 && (SM & 1) == 0;

    // where tile t's window starts, relative to the supplied window of its row
    auto tile_rel0 = [&](int64_t t, int64_t &row) {
        row = t / tiles_per_row;
        const int64_t o0 = (t - row * tiles_per_row) * (8 * S);
        return (out_first + o0) * M + gm.half - (gm.K - 1) - x_first;
    };
    // bulk-copy eligible: strictly inside the window (the copies run up to one element
    // past either end of a segment's span for 16-byte alignment)
    auto tile_tma = [&](int64_t rel0) {
        return tma_pitch && rel0 >= 1 && rel0 + gm.total_len + 2 <= x_len;
    };
    // epilogue warp: fill buffer `which` with tile t; completes one phase of bars[which]
    auto stage = [&](int64_t t, int which) {
        if (t >= ntiles) return;
        int64_t row;
        const int64_t rel0 = tile_rel0(t, row);
        const double *xr = x + row * ldx;
        if (tile_tma(rel0)) {
            if (lane == 0) {
                const double *src0 = xr + rel0;
                const int mis = span_mis(src0);
                uint32_t bytes = 0;
                for (int seg = 0; seg < nseg; ++seg) {
                    const int len = min(SM, gm.total_len - seg * SM);
                    bytes += (uint32_t)((len + mis + 1) & ~1) * 8u;
                }
                mbar_expect_tx(&bars[which], bytes);
                double *dst0 = bufs + (size_t)which * tile_elems;
                for (int seg = 0; seg < nseg; ++seg) {
                    const int len = min(SM, gm.total_len - seg * SM);
                    tma_load_1d(dst0 + (size_t)seg * P, src0 - mis + (size_t)seg * SM,
                                (uint32_t)((len + mis + 1) & ~1) * 8u, &bars[which]);
                }
            }
            return;
        }
        const uint32_t dst0 = sbase + (uint32_t)which * (uint32_t)tile_elems * 8u;
        for (int seg = 0; seg < nseg; ++seg) {
            const int len = min(SM, gm.total_len - seg * SM);
            const int64_t g0 = rel0 + (int64_t)seg * SM;
            const uint32_t dst = dst0 + (uint32_t)(seg * P) * 8u;
            for (int e = lane; e < len; e += 32) {
                const int64_t gi = g0 + e;
                const bool ok = gi >= 0 && gi < x_len;
                cp_async8_zfill(dst + e * 8u, ok ? xr + gi : xr, ok);
            }
        }
        cp_async_commit();
        cp_async_wait<0>();
        __syncwarp();
        if (lane == 0) {
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bars[which]))
                         : "memory");
        }
    };

    if (tid == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        fence_barrier_init();
    }
    for (int i = tid; i < M * gm.ldq; i += NT) gs[i] = gpad[i];
    __syncthreads();

    // named barriers: 1 + p  "partials of a tile of parity p are in red[p]" (MMA warps
    // arrive, the epilogue warp waits); 3 + p  "red[p] has been summed" (the epilogue
    // warp arrives, MMA warps wait before they write red[p] again, two tiles later)
    if (epilogue) {
        stage(blockIdx.x, 0);
        stage((int64_t)blockIdx.x + gridDim.x, 1);
        asm volatile("bar.arrive 3, %0;" ::"n"(NT) : "memory");      // red starts free
        if (RB == 2) asm volatile("bar.arrive 4, %0;" ::"n"(NT) : "memory");
        int parity = 0;
        for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x, parity ^= 1) {
            if (parity == 0) named_bar_sync_id(1, NT); else named_bar_sync_id(2, NT);
            // every MMA warp is done with buffer `parity`: sum, store, refill
            int64_t row;
            tile_rel0(t, row);
            const int64_t o0 = (t - row * tiles_per_row) * (8 * S);
            const double *r0 = red + (size_t)(RB == 2 ? parity : 0) * KS * 8 * S;
            for (int o = lane; o < 8 * S; o += 32) {
                double sum = r0[o];
#pragma unroll
                for (int k = 1; k < KS; ++k) sum += r0[(size_t)k * 8 * S + o];
                if (o0 + o < n_out) st_stream(y + row * ldy + o0 + o, sum);
            }
            if (RB == 1 || parity == 0) asm volatile("bar.arrive 3, %0;" ::"n"(NT) : "memory");
            else asm volatile("bar.arrive 4, %0;" ::"n"(NT) : "memory");
            stage(t + 2 * (int64_t)gridDim.x, parity);
        }
        return;
    }

    const int wt = warp % WT, wk = warp / WT;
    const int g = lane >> 2, q = lane & 3;
    const int pad = P - SM;
    const int logS = gm.logS;
    const int k_lo = (int)(((long)gm.ktotal * wk) / KS), k_hi = (int)(((long)gm.ktotal * (wk + 1)) / KS);
    const int oa = g * S + 32 * wt + 2 * q;          // this thread's outputs: oa + 8 j + {0, 1}
    int parity = 0;
    uint32_t phases = 0u;                            // mbarrier phase bit of either buffer
    for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x, parity ^= 1) {
        int64_t row;
        const int64_t rel0 = tile_rel0(t, row);
        mbar_wait(&bars[parity], (phases >> parity) & 1u);
        phases ^= 1u << parity;
        const int mis = tile_tma(rel0) ? span_mis(x + row * ldx + rel0) : 0;
        // ---- banded Toeplitz product on the tensor cores.  Sample n of phase p in
        //      segment g's window sits at offset p + n*M, which lies n / S segments on
        //      (p < M): one pad per segment crossed
        const double *xrow = bufs + (size_t)parity * tile_elems + (size_t)g * P + mis;
        auto fetch = [&](int p, int n) { return xrow[p + n * M + (n >> logS) * pad]; };
        double c[8];
        ufd_mma_ksteps(gm, gs, k_lo, k_hi, wt, g, q, fetch, c);
        // ---- partial sums to red[parity] once the epilogue warp has summed its previous
        //      contents, then on to the next tile
        if (RB == 1 || parity == 0) named_bar_sync_id(3, NT); else named_bar_sync_id(4, NT);
        double *rd = red + ((size_t)(RB == 2 ? parity : 0) * KS + wk) * 8 * S;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            rd[oa + 8 * j] = c[2 * j];
            rd[oa + 8 * j + 1] = c[2 * j + 1];
        }
        __threadfence_block();
        if (parity == 0) asm volatile("bar.arrive 1, %0;" ::"n"(NT) : "memory");
        else asm volatile("bar.arrive 2, %0;" ::"n"(NT) : "memory");
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
    float *d_gphasef = nullptr;    // the same taps in float32 (osz_upfirdn_plan_set_compute)
    size_t smemf = 0;              // ... with its own tile geometry: half the bytes per
    int Rf = 0, ldmf = 0;          //     sample lets a larger R fit
    int compute = 0;               // OSZ_COMPUTE_F64 / OSZ_COMPUTE_F32
    // double-buffered kernel
    bool dec2 = false;
    int ldm2 = 0, QB = 0;
    size_t smem2 = 0;
    double *d_gphase2 = nullptr;   // [down][QB*8], zero padded
    int tap_slot = -1, tap_len = 0; // the same taps in constant memory (c_ufd_taps)
    // tensor-core (DMMA) decimator
    bool mma = false;
    int kernel = OSZ_UFD_AUTO;     // osz_upfirdn_plan_set_kernel
    UfdMmaGeom mg{};
    int mma_wt = 0, mma_ks = 2, mma_rb = 1;   // tiles per warp group, k-splits, partial-sum buffers
    double mma_useful = 0.0;       // K / (4 * k-steps): useful share of the MMA work
    size_t smem_mma = 0;
    double *d_gpad = nullptr;      // [down][ldq]: window-order taps g[j] = h'[K-1-j]
    double *d_gpad_rev = nullptr;  // the same table for a reversed time axis (taps h')
    double *d_g = nullptr;         // g[j], K doubles
};

template <int WT, int KS, int RB>
static int launch_mma(const osz_upfirdn_plan *p, const double *x, int64_t ldx, int64_t rows,
                      int64_t x_first, int64_t x_len, int64_t out_first, int64_t n_out, double *y,
                      int64_t ldy, cudaStream_t st) {
    OSZ_CUDA(cudaFuncSetAttribute(upfirdn_mma_kernel<WT, KS, RB>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p->smem_mma));
    const int64_t per_tile = 8 * (int64_t)p->mg.S;
    const int64_t tiles_per_row = (n_out + per_tile - 1) / per_tile;
    const int64_t ntiles = tiles_per_row * rows;
    const int64_t grid = ntiles < sm_count() ? ntiles : sm_count();
    upfirdn_mma_kernel<WT, KS, RB><<<(unsigned)grid, WT * KS * 32 + 32, p->smem_mma, st>>>(
        p->mg, x, ldx, x_first, x_len, out_first, n_out, p->d_gpad, y, ldy, (int)tiles_per_row,
        ntiles);
    OSZ_LAUNCHED("upfirdn_mma_kernel");
    return OSZ_OK;
}

// Geometry of the tensor-core decimator for K taps, decimation M: the largest
// segment length S in {64, 48, 32, 16} whose two tile buffers (8 segments + reach
// each), padded tap table and k-split scratch fit one CTA per SM.
static bool mma_geometry(int K, int M, int ksplit, int red_bufs, UfdMmaGeom *gm, int *wt_out,
                         size_t *smem_out) {
    int smax = 0;
    long total_steps = 0;
    const std::vector<int> ks = ufd_ksteps(K, M, &smax, &total_steps);
    for (int S : {32}) {
        const int WT = S / 32;
        const int SM = S * M;
        // in-segment sample offsets run up to r_max (newest fragment of the last batch)
        const int nmax = 32 * (WT - 1) + 4 * (smax + 6) + 3;
        const int row_len = (M - 1) + nmax * M + 1;
        if (row_len > 4 * SM) continue;                   // at most three pad crossings
        const int total_len = 7 * SM + row_len;
        const int P = SM + ufd_best_pad(SM, M);
        const int nseg = (total_len + SM - 1) / SM;
        const int ldq = 7 + 4 * (smax + 1) + 4;
        const size_t smem =
            (2 * ((size_t)nseg * P + 2) + (size_t)M * ldq + (size_t)red_bufs * ksplit * 8 * S) * 8;
        if (smem > 225 * 1024) continue;
        gm->K = K;
        gm->M = M;
        gm->half = (K - 1) / 2;
        gm->S = S;
        gm->logS = S == 64 ? 6 : 5;
        gm->SM = SM;
        gm->P = P;
        gm->total_len = total_len;
        gm->ldq = ldq;
        gm->p_rem = (K - 1) % M;
        gm->ks_hi = ks[0];
        gm->ks_lo = ks[M - 1];
        gm->ktotal = (int)total_steps;
        *wt_out = WT;
        *smem_out = smem;
        return true;
    }
    return false;
}

template <int R>
static int launch_dec(const osz_upfirdn_plan *p, const double *x, int64_t ldx, int64_t rows,
                      int64_t x_first, int64_t x_len, int64_t out_first, int64_t n_out, double *y,
                      int64_t ldy, cudaStream_t st) {
    dim3 grid((unsigned)((n_out + 32 * R - 1) / (32 * R)), (unsigned)rows);
    OSZ_CUDA(cudaFuncSetAttribute(upfirdn_dec_kernel<R, double>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p->smem));
    upfirdn_dec_kernel<R, double><<<grid, UFD_NT, p->smem, st>>>(
        x, ldx, x_first, x_len, out_first, n_out, p->K, p->down, p->Q, p->ldm, p->half,
        p->d_gphase, y, ldy);
    OSZ_LAUNCHED("upfirdn_dec_kernel");
    return OSZ_OK;
}

template <int R, typename TIO>
static int launch_dec_f32(const osz_upfirdn_plan *p, const TIO *x, int64_t ldx, int64_t rows,
                          int64_t x_first, int64_t x_len, int64_t out_first, int64_t n_out,
                          TIO *y, int64_t ldy, cudaStream_t st) {
    dim3 grid((unsigned)((n_out + 32 * R - 1) / (32 * R)), (unsigned)rows);
    OSZ_CUDA(cudaFuncSetAttribute(upfirdn_dec_kernel<R, float, TIO>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p->smemf));
    upfirdn_dec_kernel<R, float, TIO><<<grid, UFD_NT, p->smemf, st>>>(
        x, ldx, x_first, x_len, out_first, n_out, p->K, p->down, p->Q, p->ldmf, p->half,
        p->d_gphasef, y, ldy);
    OSZ_LAUNCHED("upfirdn_dec_kernel<float>");
    return OSZ_OK;
}

// first-fit allocator over c_ufd_taps (a handful of live plans at most)
#include <mutex>
static std::mutex g_slot_mu;
static std::vector<std::pair<int, int>> g_slots;   // (offset, length), sorted by offset
static int ufd_slot_alloc(int len) {
    std::lock_guard<std::mutex> lk(g_slot_mu);
    len = (len + 1) & ~1;
    int at = 0;
    size_t i = 0;
    for (; i < g_slots.size(); ++i) {
        if (g_slots[i].first - at >= len) break;
        at = g_slots[i].first + g_slots[i].second;
    }
    if (at + len > UFD2_CONST_DOUBLES) return -1;
    g_slots.insert(g_slots.begin() + i, std::make_pair(at, len));
    return at;
}
static void ufd_slot_free(int off, int) {
    std::lock_guard<std::mutex> lk(g_slot_mu);
    for (size_t i = 0; i < g_slots.size(); ++i)
        if (g_slots[i].first == off) {
            g_slots.erase(g_slots.begin() + i);
            return;
        }
}

template <int R>
static int launch_dec2(const osz_upfirdn_plan *p, const double *x, int64_t ldx, int64_t rows,
                       int64_t x_first, int64_t x_len, int64_t out_first, int64_t n_out, double *y,
                       int64_t ldy, cudaStream_t st) {
    OSZ_CUDA(cudaFuncSetAttribute(upfirdn_dec2_kernel<R>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p->smem2));
    const int64_t tiles_per_row = (n_out + 32 * R - 1) / (32 * R);
    const int64_t ntiles = tiles_per_row * rows;
    int64_t grid = ntiles < sm_count() ? ntiles : sm_count();
    upfirdn_dec2_kernel<R><<<(unsigned)grid, UFD2_NT, p->smem2, st>>>(
        x, ldx, x_first, x_len, out_first, n_out, p->K, p->down, p->QB, p->ldm2, p->half,
        p->d_gphase2, y, ldy, (int)tiles_per_row, ntiles, p->tap_slot);
    OSZ_LAUNCHED("upfirdn_dec2_kernel");
    return OSZ_OK;
}

static size_t dec_smem(int R, int M, int Q, int *ldm_out, size_t elem = 8) {
    const int TO = 32 * R;
    int ldm = (TO + Q + R) + (TO + Q + R) / R + 1;
    if ((ldm & 1) == 0) ++ldm;
    *ldm_out = ldm;
    size_t tile = ((size_t)M * ldm + (size_t)M * Q) * elem;
    size_t red = (size_t)UFD_NW * 32 * (R + 1) * elem;
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
            // double-buffered kernel: R = 8, taps padded to whole blocks of R per phase
            if (ok) {
                constexpr int R2 = 8, TO2 = 32 * R2;
                const int QB = (p->Q + R2 - 1) / R2, Qp = QB * R2;
                const int cols = TO2 + Qp + R2;                 // decimated samples per phase row
                int ldm2 = cols + cols / R2 + 1;
                if ((ldm2 & 1) == 0) ++ldm2;
                const size_t smem2 = (2 * (size_t)M * ldm2 + (size_t)UFD2_NW * 32 * (R2 + 1)) * 8;
                const int per_thread = ((TO2 + Qp) * M + UFD2_NT - 1) / UFD2_NT;
                const int slot = smem2 <= 225 * 1024 && per_thread <= UFD2_MAXE
                                     ? ufd_slot_alloc(M * Qp) : -1;
                if (slot >= 0) {
                    std::vector<double> gp2((size_t)M * Qp, 0.0);
                    for (int j = 0; j < K; ++j) gp2[(size_t)(j % M) * Qp + j / M] = hs[K - 1 - j];
                    ok = cudaMalloc(&p->d_gphase2, gp2.size() * 8) == cudaSuccess &&
                         cudaMemcpy(p->d_gphase2, gp2.data(), gp2.size() * 8,
                                    cudaMemcpyHostToDevice) == cudaSuccess;
                    p->tap_slot = slot;
                    p->tap_len = M * Qp;
                    ok = ok && cudaMemcpyToSymbol(c_ufd_taps, gp2.data(), gp2.size() * 8,
                                                  (size_t)slot * 8) == cudaSuccess;
                    p->dec2 = ok;
                    p->ldm2 = ldm2;
                    p->QB = QB;
                    p->smem2 = smem2;
                }
            }
        }
    }
    if (ok && up == 1 && down >= 2 && p->R) {
        static const int ksplit = [] {
            const char *e = getenv("OSZ_UFD_MMA_KS");
            const int v = e ? atoi(e) : 8;
            return v == 8 || v == 12 || v == 16 ? v : 8;
        }();
        p->mma_ks = ksplit;
        // one partial-sum buffer (default) keeps the CTA at 160 KB of shared memory, which
        // leaves room for a CTA of the biquad scan on the same SM (its forward pass runs
        // on another stream): OSZ_UFD_MMA_RED=2 double-buffers the partial sums
        static const int red_bufs = [] {
            const char *e = getenv("OSZ_UFD_MMA_RED");
            return e && atoi(e) == 2 ? 2 : 1;
        }();
        p->mma_rb = red_bufs;
        if (mma_geometry(K, down, p->mma_ks, p->mma_rb, &p->mg, &p->mma_wt, &p->smem_mma)) {
            // taps in the order the kernel walks its window: g[j] = h'[K-1-j]
            std::vector<double> g(K);
            for (int j = 0; j < K; ++j) g[j] = hs[K - 1 - j];
            const std::vector<double> gp = ufd_gpad(g.data(), K, down, p->mg.ldq);
            const std::vector<double> gpr = ufd_gpad(hs.data(), K, down, p->mg.ldq);
            ok = cudaMalloc(&p->d_gpad, gp.size() * 8) == cudaSuccess &&
                 cudaMemcpy(p->d_gpad, gp.data(), gp.size() * 8, cudaMemcpyHostToDevice) ==
                     cudaSuccess &&
                 cudaMalloc(&p->d_gpad_rev, gpr.size() * 8) == cudaSuccess &&
                 cudaMemcpy(p->d_gpad_rev, gpr.data(), gpr.size() * 8, cudaMemcpyHostToDevice) ==
                     cudaSuccess &&
                 cudaMalloc(&p->d_g, (size_t)K * 8) == cudaSuccess &&
                 cudaMemcpy(p->d_g, g.data(), (size_t)K * 8, cudaMemcpyHostToDevice) ==
                     cudaSuccess;
            p->mma = ok;
            p->mma_useful = (double)K / (4.0 * p->mg.ktotal);
        }
    }
    if (!ok) {
        osz_upfirdn_plan_destroy(p);
        return fail(OSZ_ERR_CUDA, "osz_upfirdn_plan_create: device upload failed");
    }
    *out = p;
    return OSZ_OK;
}

int osz_upfirdn_plan_set_compute(osz_upfirdn_plan *p, int compute) {
    if (!p || (compute != OSZ_COMPUTE_F64 && compute != OSZ_COMPUTE_F32))
        return fail(OSZ_ERR_ARG, "osz_upfirdn_plan_set_compute: bad arguments");
    if (compute == OSZ_COMPUTE_F64 || !p->R) {   // float32 exists for the decimating kernel
        p->compute = OSZ_COMPUTE_F64;
        return OSZ_OK;
    }
    if (!p->d_gphasef) {
        const size_t n = (size_t)p->down * p->Q;
        std::vector<double> gp(n);
        std::vector<float> gf(n);
        bool ok = cudaMemcpy(gp.data(), p->d_gphase, n * 8, cudaMemcpyDeviceToHost) == cudaSuccess;
        for (size_t i = 0; i < n; ++i) gf[i] = (float)gp[i];
        ok = ok && cudaMalloc(&p->d_gphasef, n * 4) == cudaSuccess &&
             cudaMemcpy(p->d_gphasef, gf.data(), n * 4, cudaMemcpyHostToDevice) == cudaSuccess;
        if (!ok) return fail(OSZ_ERR_CUDA, "osz_upfirdn_plan_set_compute: device upload failed");
        for (int R : {16, 8, 4}) {
            int ldm = 0;
            const size_t sm = dec_smem(R, p->down, p->Q, &ldm, 4);
            if (sm <= 100 * 1024) {
                p->Rf = R;
                p->ldmf = ldm;
                p->smemf = sm;
                break;
            }
        }
    }
    if (!p->Rf) return OSZ_OK;                   // (cannot happen when the float64 tile fits)
    p->compute = OSZ_COMPUTE_F32;
    return OSZ_OK;
}
int osz_upfirdn_plan_compute(const osz_upfirdn_plan *p) { return p ? p->compute : 0; }

// internal (sosdec.cu): the decimating filter's tap tables
int osz_upfirdn_plan_taps(const osz_upfirdn_plan *p, int *K, int *M, int *half,
                          const double **d_gpad_fwd, const double **d_gpad_rev,
                          const double **d_taps_g, int *ldq) {
    if (!p) return fail(OSZ_ERR_ARG, "osz_upfirdn_plan_taps: null plan");
    if (!p->mma) return fail(OSZ_ERR_UNSUPPORTED, "osz_upfirdn_plan_taps: not a decimating plan "
                                                  "with a tensor-core geometry");
    *K = p->K;
    *M = p->down;
    *half = p->half;
    *d_gpad_fwd = p->d_gpad;
    *d_gpad_rev = p->d_gpad_rev;
    *d_taps_g = p->d_g;
    *ldq = p->mg.ldq;
    return OSZ_OK;
}

static int ufd_use_mma_default() {
    static const int v = [] {
        const char *e = getenv("OSZ_UFD_MMA");
        return e ? atoi(e) : 1;
    }();
    return v;
}
int osz_upfirdn_plan_set_kernel(osz_upfirdn_plan *p, int kernel) {
    if (!p || kernel < OSZ_UFD_AUTO || kernel > OSZ_UFD_MMA)
        return fail(OSZ_ERR_ARG, "osz_upfirdn_plan_set_kernel: bad arguments");
    if (kernel == OSZ_UFD_MMA && !p->mma)
        return fail(OSZ_ERR_UNSUPPORTED, "osz_upfirdn_plan_set_kernel: no tensor-core geometry "
                                         "for this filter (up > 1, or the tile does not fit)");
    p->kernel = kernel;
    return OSZ_OK;
}
int osz_upfirdn_plan_kernel(const osz_upfirdn_plan *p) {
    if (!p) return 0;
    if (!p->R) return OSZ_UFD_GENERAL;
    // AUTO: the tensor-core kernel when enough of its MMA work is useful (long per-phase
    // filters: 8 outputs share a window, so a phase of Q taps costs Q + 7).  Measured on B200
    // (256 x 1e6): 1231 taps / M = 25 (86 % useful): 1.22 ms against 1.55 ms for the CUDA-core
    // kernel; 561 / 25 (70 %): 0.82 against 0.95; 449 / 20 (70 %, even M: the even segment
    // pitch the TMA staging needs leaves a 2-way bank conflict on the A fragments): 0.98
    // against 0.96 -- so 68 % is enough for odd M, even M wants 80 %.
    const double need = (p->down & 1) ? 0.68 : 0.8;
    const bool worth = p->mma && p->mma_useful >= need && (p->mg.P % 2 == 0);
    if (p->kernel == OSZ_UFD_MMA || (p->kernel == OSZ_UFD_AUTO && worth && ufd_use_mma_default()))
        return OSZ_UFD_MMA;
    return OSZ_UFD_POLYPHASE;
}

int osz_upfirdn_plan_destroy(osz_upfirdn_plan *p) {
    if (!p) return OSZ_OK;
    cudaFree(p->d_h);
    cudaFree(p->d_gphase);
    cudaFree(p->d_gphasef);
    cudaFree(p->d_gphase2);
    cudaFree(p->d_gpad);
    cudaFree(p->d_gpad_rev);
    cudaFree(p->d_g);
    if (p->tap_slot >= 0) ufd_slot_free(p->tap_slot, p->tap_len);
    delete p;
    return OSZ_OK;
}

// float32 I/O: float samples in and out, the float32-arithmetic decimating kernel (plans
// with up == 1 set to OSZ_COMPUTE_F32).
int osz_upfirdn_exec_f32(const osz_upfirdn_plan *p, const float *x, int64_t ldx, int64_t rows,
                         int64_t x_first, int64_t x_len, int64_t out_first, int64_t n_out,
                         float *y, int64_t ldy, void *stream) {
    if (!p || !x || !y) return fail(OSZ_ERR_ARG, "osz_upfirdn_exec_f32: null argument");
    if (rows <= 0 || n_out <= 0) return OSZ_OK;
    if (rows > 65535) return fail(OSZ_ERR_UNSUPPORTED, "upfirdn: more than 65535 rows per call");
    if (p->compute != OSZ_COMPUTE_F32 || !p->d_gphasef || !p->Rf)
        return fail(OSZ_ERR_UNSUPPORTED, "osz_upfirdn_exec_f32: needs a decimating plan set to "
                                         "float32 arithmetic");
    cudaStream_t st = as_stream(stream);
    switch (p->Rf) {
        case 16:
            return launch_dec_f32<16, float>(p, x, ldx, rows, x_first, x_len, out_first, n_out, y, ldy, st);
        case 8:
            return launch_dec_f32<8, float>(p, x, ldx, rows, x_first, x_len, out_first, n_out, y, ldy, st);
        default:
            return launch_dec_f32<4, float>(p, x, ldx, rows, x_first, x_len, out_first, n_out, y, ldy, st);
    }
}

int osz_upfirdn_exec_f64(const osz_upfirdn_plan *p, const double *x, int64_t ldx, int64_t rows,
                         int64_t x_first, int64_t x_len, int64_t out_first, int64_t n_out,
                         double *y, int64_t ldy, void *stream) {
    if (!p || !x || !y) return fail(OSZ_ERR_ARG, "osz_upfirdn_exec_f64: null argument");
    if (rows <= 0 || n_out <= 0) return OSZ_OK;
    if (rows > 65535) return fail(OSZ_ERR_UNSUPPORTED, "upfirdn: more than 65535 rows per call");
    cudaStream_t st = as_stream(stream);
    static const int use_dec2 = [] {
        // measured on B200 (256 x 1e6, profiles/r01_kernel_bench.md): 158.6 vs 164.9 G
        // samples/s for the fused 1231-tap filter, 255 vs 268 at 561 taps -- the
        // double-buffered kernel stays opt-in until its tap delivery is uniform
        const char *e = getenv("OSZ_UFD_DEC2");
        return e ? atoi(e) : 0;
    }();
    if (p->compute == OSZ_COMPUTE_F32 && p->d_gphasef) {
        switch (p->Rf) {
            case 16:
                return launch_dec_f32<16, double>(p, x, ldx, rows, x_first, x_len, out_first, n_out, y, ldy, st);
            case 8:
                return launch_dec_f32<8, double>(p, x, ldx, rows, x_first, x_len, out_first, n_out, y, ldy, st);
            default:
                return launch_dec_f32<4, double>(p, x, ldx, rows, x_first, x_len, out_first, n_out, y, ldy, st);
        }
    }
    if (osz_upfirdn_plan_kernel(p) == OSZ_UFD_MMA && !(p->dec2 && use_dec2)) {
#define OSZ_MMA_CASE(WT, KS)                                                                   \
    if (p->mma_wt == WT && p->mma_ks == KS) {                                                 \
        if (p->mma_rb == 1)                                                                   \
            return launch_mma<WT, KS, 1>(p, x, ldx, rows, x_first, x_len, out_first, n_out, y, ldy, st); \
        return launch_mma<WT, KS, 2>(p, x, ldx, rows, x_first, x_len, out_first, n_out, y, ldy, st);     \
    }
        OSZ_MMA_CASE(1, 8) OSZ_MMA_CASE(1, 12) OSZ_MMA_CASE(1, 16)
#undef OSZ_MMA_CASE
    }
    if (p->dec2 && use_dec2)
        return launch_dec2<8>(p, x, ldx, rows, x_first, x_len, out_first, n_out, y, ldy, st);
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
