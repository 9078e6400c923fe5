// Shared-memory radix-16 Stockham FFT, complex float64, N = 2^8 .. 2^13.
//
// One CTA of N/16 threads transforms one N-point complex sequence.  Each
// thread enters with 16 values in REGISTERS, v[r] = x[tid + r*N/16], and
// leaves with v[r] = X[tid + r*N/16] -- the same strided-natural layout, so a
// forward transform, a pointwise multiply and an inverse transform chain
// without touching memory in between, and global loads/stores on either side
// are coalesced.  Passes are [16, mid..., 16] with mid in {-, 2, 4, 8, 16,
// 16*2}; between passes data is exchanged through shared memory (AoS
// double2, one pad element per 16 so that both the stride-1 and the stride-16
// patterns are bank-conflict free).
//
// Twiddles: a thread's butterflies in a given pass always use the same base
// twiddle w = exp(-2 pi i k / (Ns R)) (k depends only on the thread index), so
// each thread loads its three base twiddles ONCE per kernel (FftTw) and forms
// w^2 .. w^15 with a depth-4 product tree in registers -- no table traffic in
// the transform loops (the first version read 30 twiddles per transform through
// L1 and was LSU bound: profiles/r01_ncu_summary.md).
#pragma once

#include <math.h>

#include <vector>

#include "common.cuh"

namespace osz {

__device__ __forceinline__ double2 cadd(double2 a, double2 b) {
    return make_double2(a.x + b.x, a.y + b.y);
}
__device__ __forceinline__ double2 csub(double2 a, double2 b) {
    return make_double2(a.x - b.x, a.y - b.y);
}
__device__ __forceinline__ double2 cmul(double2 a, double2 b) {
    return make_double2(fma(a.x, b.x, -a.y * b.y), fma(a.x, b.y, a.y * b.x));
}
// a * (-i)
__device__ __forceinline__ double2 mul_mi(double2 a) { return make_double2(a.y, -a.x); }

#define OSZ_SQRT1_2 0.70710678118654752440
#define OSZ_COS_PI_8 0.92387953251128675613
#define OSZ_SIN_PI_8 0.38268343236508977173

// a * W8^1 = a * (1 - i)/sqrt2
__device__ __forceinline__ double2 mul_w8_1(double2 a) {
    return make_double2((a.x + a.y) * OSZ_SQRT1_2, (a.y - a.x) * OSZ_SQRT1_2);
}
// a * W8^3 = a * (-1 - i)/sqrt2
__device__ __forceinline__ double2 mul_w8_3(double2 a) {
    return make_double2((a.y - a.x) * OSZ_SQRT1_2, -(a.x + a.y) * OSZ_SQRT1_2);
}

__device__ __forceinline__ void bfly2(double2 &a0, double2 &a1) {
    double2 t = a0;
    a0 = cadd(t, a1);
    a1 = csub(t, a1);
}

// forward 4-point DFT, natural order in and out
__device__ __forceinline__ void bfly4(double2 &a0, double2 &a1, double2 &a2, double2 &a3) {
    double2 t0 = cadd(a0, a2), t1 = csub(a0, a2);
    double2 t2 = cadd(a1, a3), t3 = mul_mi(csub(a1, a3));
    a0 = cadd(t0, t2);
    a1 = cadd(t1, t3);
    a2 = csub(t0, t2);
    a3 = csub(t1, t3);
}

// Forward R-point DFT of v[0..R), natural order in and out.
template <int R>
__device__ __forceinline__ void bfly(double2 *v);

template <>
__device__ __forceinline__ void bfly<2>(double2 *v) {
    bfly2(v[0], v[1]);
}
template <>
__device__ __forceinline__ void bfly<4>(double2 *v) {
    bfly4(v[0], v[1], v[2], v[3]);
}
template <>
__device__ __forceinline__ void bfly<8>(double2 *v) {
    // n = 4*n1 + n2, k = k1 + 2*k2
    bfly2(v[0], v[4]);
    bfly2(v[1], v[5]);
    bfly2(v[2], v[6]);
    bfly2(v[3], v[7]);
    v[5] = mul_w8_1(v[5]);
    v[6] = mul_mi(v[6]);
    v[7] = mul_w8_3(v[7]);
    bfly4(v[0], v[1], v[2], v[3]);  // k1 = 0 -> X[0], X[2], X[4], X[6]
    bfly4(v[4], v[5], v[6], v[7]);  // k1 = 1 -> X[1], X[3], X[5], X[7]
    double2 o1 = v[4], o2 = v[1], o3 = v[5], o4 = v[2], o5 = v[6], o6 = v[3];
    v[1] = o1;
    v[2] = o2;
    v[3] = o3;
    v[4] = o4;
    v[5] = o5;
    v[6] = o6;
}
template <>
__device__ __forceinline__ void bfly<16>(double2 *v) {
    // n = 4*n1 + n2, k = k1 + 4*k2.  Stage 1: DFT over n1 for each n2.
    bfly4(v[0], v[4], v[8], v[12]);
    bfly4(v[1], v[5], v[9], v[13]);
    bfly4(v[2], v[6], v[10], v[14]);
    bfly4(v[3], v[7], v[11], v[15]);
    // now v[n2 + 4*k1] = A[n2][k1]; twiddle by W16^(n2*k1)
    const double2 w1 = make_double2(OSZ_COS_PI_8, -OSZ_SIN_PI_8);
    const double2 w3 = make_double2(OSZ_SIN_PI_8, -OSZ_COS_PI_8);
    v[5] = cmul(v[5], w1);                                       // n2=1,k1=1 : W^1
    v[6] = mul_w8_1(v[6]);                                       // n2=2,k1=1 : W^2
    v[7] = cmul(v[7], w3);                                       // n2=3,k1=1 : W^3
    v[9] = mul_w8_1(v[9]);                                       // n2=1,k1=2 : W^2
    v[10] = mul_mi(v[10]);                                       // n2=2,k1=2 : W^4
    v[11] = mul_w8_3(v[11]);                                     // n2=3,k1=2 : W^6
    v[13] = cmul(v[13], w3);                                     // n2=1,k1=3 : W^3
    v[14] = mul_w8_3(v[14]);                                     // n2=2,k1=3 : W^6
    v[15] = cmul(v[15], make_double2(-OSZ_COS_PI_8, OSZ_SIN_PI_8));  // n2=3,k1=3 : W^9
    // Stage 2: DFT over n2 for each k1 -> X[k1 + 4*k2] lands in v[k2 + 4*k1]
    bfly4(v[0], v[1], v[2], v[3]);
    bfly4(v[4], v[5], v[6], v[7]);
    bfly4(v[8], v[9], v[10], v[11]);
    bfly4(v[12], v[13], v[14], v[15]);
    // 4x4 transpose of register names -> natural order
    double2 t;
#define OSZ_SWAP(a, b) \
    t = v[a];          \
    v[a] = v[b];       \
    v[b] = t;
    OSZ_SWAP(1, 4)
    OSZ_SWAP(2, 8)
    OSZ_SWAP(3, 12)
    OSZ_SWAP(6, 9)
    OSZ_SWAP(7, 13)
    OSZ_SWAP(11, 14)
#undef OSZ_SWAP
}

template <int LOG2N>
struct FftCfg {
    static_assert(LOG2N >= 8 && LOG2N <= 13, "shared-memory FFT supports N = 256 .. 8192");
    static constexpr int N = 1 << LOG2N;
    static constexpr int NT = N / 16;                 // threads per CTA
    static constexpr int MIDBITS = LOG2N - 8;
    static constexpr int R1 = MIDBITS >= 4 ? 16 : (1 << MIDBITS);   // first middle radix (1: none)
    static constexpr int R2 = MIDBITS > 4 ? (1 << (MIDBITS - 4)) : 1;  // second middle radix
    // base-twiddle table offsets (double2 elements): [mid1: 16][mid2: 16*R1][last: N/16]
    static constexpr int OFF_M1 = 0;
    static constexpr int OFF_M2 = 16;
    static constexpr int OFF_L = 16 + 16 * R1;
    static constexpr int TW_TOTAL = OFF_L + N / 16;
    static constexpr int SMEM_ELEMS = N + N / 16;     // padded double2 elements
    static constexpr int SMEM_BYTES = SMEM_ELEMS * 16;
};

__device__ __forceinline__ int fft_phys(int i) { return i + (i >> 4); }
// padded offset of a multiple of 16: a compile-time constant, so every shared
// memory access below is `one base register + immediate` (the first version
// recomputed fft_phys() per element and the 48 addresses it kept live across
// transforms spilled).
__host__ __device__ constexpr int fft_pad16(int x) { return x + x / 16; }

__device__ __forceinline__ double2 csqr(double2 a) {
    return make_double2(fma(a.x, a.x, -a.y * a.y), 2.0 * a.x * a.y);
}

// High powers of a base twiddle: they only depend on the table entry, so a
// ping-pong group computes them BEFORE it takes the FP64 token (this keeps the
// serial squaring chain off the token's critical path).
struct TwPre {
    double2 w4, w8, w12;
};
template <int R>
__device__ __forceinline__ TwPre twiddle_pre(double2 w1) {
    TwPre p;
    p.w4 = p.w8 = p.w12 = make_double2(1.0, 0.0);
    if constexpr (R >= 8) {
        p.w4 = csqr(csqr(w1));
        if constexpr (R >= 16) {
            p.w8 = csqr(p.w4);
            p.w12 = cmul(p.w8, p.w4);
        }
    }
    return p;
}

// v[r] *= w^r for r = 1 .. R-1, powers from a shallow product tree.
template <int R>
__device__ __forceinline__ void twiddle_pow(double2 *v, double2 w1, const TwPre &pre) {
    if constexpr (R >= 2) v[1] = cmul(v[1], w1);
    if constexpr (R >= 4) {
        const double2 w2 = csqr(w1), w3 = cmul(w2, w1);
        v[2] = cmul(v[2], w2);
        v[3] = cmul(v[3], w3);
        if constexpr (R >= 8) {
            const double2 w4 = pre.w4;
            v[4] = cmul(v[4], w4);
            v[5] = cmul(v[5], cmul(w4, w1));
            v[6] = cmul(v[6], cmul(w4, w2));
            v[7] = cmul(v[7], cmul(w4, w3));
            if constexpr (R >= 16) {
                const double2 w8 = pre.w8, w12 = pre.w12;
                v[8] = cmul(v[8], w8);
                v[9] = cmul(v[9], cmul(w8, w1));
                v[10] = cmul(v[10], cmul(w8, w2));
                v[11] = cmul(v[11], cmul(w8, w3));
                v[12] = cmul(v[12], w12);
                v[13] = cmul(v[13], cmul(w12, w1));
                v[14] = cmul(v[14], cmul(w12, w2));
                v[15] = cmul(v[15], cmul(w12, w3));
            }
        }
    }
}

// A thread's base twiddles for the middle passes and the last pass.
struct FftTw {
    double2 m1, m2, last;
};

template <int LOG2N>
__device__ __forceinline__ FftTw fft_load_tw(const double2 *__restrict__ tw, int tid) {
    using C = FftCfg<LOG2N>;
    FftTw t;
    t.m1 = ldg(tw + C::OFF_M1 + (tid & 15));
    t.m2 = ldg(tw + C::OFF_M2 + (tid & (16 * C::R1 - 1)));
    t.last = ldg(tw + C::OFF_L + tid);
    return t;
}

// ---- synchronisation policies ------------------------------------------------
// The transform alternates FP64 phases (twiddle + butterfly, registers only)
// with exchange phases (shared-memory stores/loads around a barrier).  Two
// CTAs that merely share an SM fall into lock step -- both in an FP64 phase,
// then both in an exchange phase, because the pipes are shared round-robin --
// and each pipe idles half of the time (profiles/r01_ncu_summary.md: FP64
// 50 % + LSU 55 %).  SyncPingPong runs TWO transforms in one CTA, one per
// thread group, and passes an FP64 token between the groups with named
// barriers, so that one group's twiddled butterflies overlap the other's
// exchange.  ptxas moves arithmetic freely across BAR instructions, so the
// token is tied to the data flow: the pass's base twiddle is read from shared
// memory AFTER the acquiring barrier (every product of the pass depends on it),
// and the releasing barrier's id is computed from the butterfly results.
struct SyncCta {
    FftTw t;
    __device__ __forceinline__ void group() const { __syncthreads(); }
    __device__ __forceinline__ double2 peek_m1() const { return t.m1; }
    __device__ __forceinline__ double2 peek_m2() const { return t.m2; }
    __device__ __forceinline__ double2 peek_last() const { return t.last; }
    __device__ __forceinline__ double2 acquire(double2 peeked) const { return peeked; }
    __device__ __forceinline__ void acquire_first(double2 (&)[16]) const {}
    __device__ __forceinline__ void release(double2 (&)[16]) const {}
};

// Two groups of N/16 threads; barrier ids: 1+g group-local, 3+g "group g may
// compute".  Group 1 calls prime() once before the main loop, group 0 drain()
// once after it, and both groups execute the same number of acquire/release
// pairs.  `tw_sm` is the CTA's shared-memory copy of the base-twiddle table,
// `zero` an opaque 0 (a kernel argument the compiler cannot fold).
template <int LOG2N>
struct SyncPingPong {
    using C = FftCfg<LOG2N>;
    static constexpr int GT = C::NT;
    int g, tid, zero;
    const double2 *tw_sm;
    __device__ __forceinline__ void group() const {
        asm volatile("bar.sync %0, %1;" ::"r"(1 + g), "n"(GT) : "memory");
    }
    // Wait for the token; returns an opaque 0 that exists only after the wait (a
    // clock read masked by `zero`).  XOR-ing it into an operand of the phase's
    // arithmetic is the data dependency that keeps ptxas from hoisting that
    // arithmetic above the barrier.  (A shared-memory read after the barrier
    // would do the same, but it queues behind the other group's exchange
    // traffic: 180 cycles per phase in the first version.)
    __device__ __forceinline__ int take() const {
        asm volatile("bar.sync %0, %1;" ::"r"(3 + g), "n"(2 * GT) : "memory");
        int c;
        asm volatile("mov.u32 %0, %%clock;" : "=r"(c)::"memory");
        return c & zero;
    }
    static __device__ __forceinline__ double tie(double a, int t) {
        return __hiloint2double(__double2hiint(a) ^ t, __double2loint(a));
    }
    __device__ __forceinline__ double2 taken(double2 w) const {
        const int t = take();
        return make_double2(tie(w.x, t), tie(w.y, t));
    }
    __device__ __forceinline__ double2 peek_m1() const { return tw_sm[C::OFF_M1 + (tid & 15)]; }
    __device__ __forceinline__ double2 peek_m2() const {
        return tw_sm[C::OFF_M2 + (tid & (16 * C::R1 - 1))];
    }
    __device__ __forceinline__ double2 peek_last() const { return tw_sm[C::OFF_L + tid]; }
    __device__ __forceinline__ double2 acquire(double2 peeked) const { return taken(peeked); }
    // Token for the twiddle-free first pass: tie the four values every
    // first-stage butterfly starts from.
    __device__ __forceinline__ void acquire_first(double2 (&v)[16]) const {
        const int t = take();
#pragma unroll
        for (int r = 0; r < 4; ++r) v[r] = make_double2(tie(v[r].x, t), tie(v[r].y, t));
    }
    // Pass the token on.  The barrier id is computed from the results of the
    // last butterfly stage (two outputs of each final 4-point butterfly), so the
    // arrive cannot be scheduled ahead of the arithmetic.
    __device__ __forceinline__ void release(double2 (&v)[16]) const {
        int t[4];
#pragma unroll
        for (int q = 0; q < 4; ++q)
            t[q] = __double2hiint(v[q].x) | __double2hiint(v[q].y) | __double2hiint(v[q + 4].x) |
                   __double2hiint(v[q + 4].y);
        const int u = ((t[0] | t[1]) | (t[2] | t[3])) & zero;
        asm volatile("bar.arrive %0, %1;" ::"r"(4 - g + u), "n"(2 * GT) : "memory");
    }
    // a turn without work: shifts this group's phase sequence against the other's
    __device__ __forceinline__ void idle_turn() const {
        asm volatile("bar.sync %0, %1;" ::"r"(3 + g), "n"(2 * GT) : "memory");
        asm volatile("bar.arrive %0, %1;" ::"r"(4 - g), "n"(2 * GT) : "memory");
    }
    __device__ __forceinline__ void prime() const {
        if (g == 1) asm volatile("bar.arrive 3, %0;" ::"n"(2 * GT) : "memory");
    }
    __device__ __forceinline__ void drain() const {
        if (g == 0) asm volatile("bar.sync 3, %0;" ::"n"(2 * GT) : "memory");
    }
};

// One middle pass (radix R, sub-transform length Ns) over the group's N points,
// in place in shared memory: every thread reads all its inputs, the group
// synchronises, then butterflies are written to their Stockham positions.
// All of a thread's butterflies share k = tid mod Ns, hence one base twiddle.
template <int N, int R, int NS, bool FIRST_MID, class Sync>
__device__ __forceinline__ void fft_mid_pass(double2 *sm, int tid, const Sync &sync) {
    constexpr int NT = N / 16;
    constexpr int PER = 16 / R;  // butterflies per thread
    static_assert(NT % NS == 0, "k must not depend on the butterfly index");
    static_assert(NT % 16 == 0 && (N / R) % 16 == 0 && NS % 16 == 0, "padding arithmetic");
    double2 v[16];
    const double2 *src = sm + fft_phys(tid);
#pragma unroll
    for (int q = 0; q < PER; ++q) {
#pragma unroll
        for (int r = 0; r < R; ++r) v[q * R + r] = src[fft_pad16(q * NT + r * (N / R))];
    }
    const double2 wp = FIRST_MID ? sync.peek_m1() : sync.peek_m2();
    const TwPre pre = twiddle_pre<R>(wp);
    sync.group();
    const double2 w1 = sync.acquire(wp);
    const int k = tid & (NS - 1);
    double2 *dst = sm + fft_phys((tid - k) * R + k);
#pragma unroll
    for (int q = 0; q < PER; ++q) {
        twiddle_pow<R>(v + q * R, w1, pre);
        bfly<R>(v + q * R);
    }
    sync.release(v);
#pragma unroll
    for (int q = 0; q < PER; ++q) {
#pragma unroll
        for (int r = 0; r < R; ++r) dst[fft_pad16(q * NT * R + r * NS)] = v[q * R + r];
    }
    sync.group();
}

// Forward FFT, registers to registers (see file header).  `sm` must hold
// FftCfg<LOG2N>::SMEM_ELEMS double2 (per group); `tid` is the index within the
// group.  Contains its own leading barrier, so it can be called back to back.
// The first pass (no twiddles) runs outside the FP64 token.  REL = false: the
// caller continues with FP64 work after the last pass and calls
// sync.release(v) itself.
struct FftNoHook {
    __device__ __forceinline__ void operator()() const {}
};

// `after_loads()` runs once the last pass has read its inputs: from then on the
// transform no longer touches `sm` (callers refill it for the next item there).
template <int LOG2N, class Sync, bool REL = true, class Hook = FftNoHook>
__device__ __forceinline__ void fft_r2r_tail(double2 (&v)[16], double2 *sm, int tid,
                                             const Sync &sync, const Hook &after_loads = Hook()) {
    using C = FftCfg<LOG2N>;
    constexpr int N = C::N, NT = C::NT;
    // (pass 1, a twiddle-free bfly<16>(v), has been done by the caller)
    sync.group();  // previous users of sm are done
    {
        double2 *dst = sm + 17 * tid;     // fft_phys(16 * tid + r) = 17 * tid + r
#pragma unroll
        for (int r = 0; r < 16; ++r) dst[r] = v[r];
    }
    sync.group();
    if constexpr (C::R1 > 1) fft_mid_pass<N, C::R1, 16, true>(sm, tid, sync);
    if constexpr (C::R2 > 1) fft_mid_pass<N, C::R2, 16 * C::R1, false>(sm, tid, sync);
    // last pass: radix 16, Ns = N/16, k = tid, output index tid + r*NT
    {
        const double2 *src = sm + fft_phys(tid);
#pragma unroll
        for (int r = 0; r < 16; ++r) v[r] = src[fft_pad16(r * NT)];
    }
    const double2 wp = sync.peek_last();
    const TwPre pre = twiddle_pre<16>(wp);
    after_loads();
    const double2 wl = sync.acquire(wp);
    twiddle_pow<16>(v, wl, pre);
    bfly<16>(v);
    if (REL) sync.release(v);
}

template <int LOG2N, class Sync, bool REL = true>
__device__ __forceinline__ void fft_r2r(double2 (&v)[16], double2 *sm, int tid, const Sync &sync) {
    bfly<16>(v);   // pass 1: radix 16, Ns = 1, no twiddles
    fft_r2r_tail<LOG2N, Sync, REL>(v, sm, tid, sync);
}

template <int LOG2N>
__device__ __forceinline__ void fft_r2r(double2 (&v)[16], double2 *sm, const FftTw &tw, int tid) {
    fft_r2r<LOG2N, SyncCta>(v, sm, tid, SyncCta{tw});
}

// Host: base-twiddle tables for FftCfg<log2n>, as (re, im) pairs.
inline std::vector<double> make_fft_twiddles(int log2n) {
    const int N = 1 << log2n;
    const int mid = log2n - 8;
    const int R1 = mid >= 4 ? 16 : (1 << mid);
    const int R2 = mid > 4 ? (1 << (mid - 4)) : 1;
    std::vector<double> t;
    auto emit = [&](int count, long double period) {
        for (int k = 0; k < count; ++k) {
            // exp(-2 pi i k / period); long double keeps the table correctly rounded
            long double a = -2.0L * 3.14159265358979323846264338327950288L * (long double)k / period;
            t.push_back((double)cosl(a));
            t.push_back((double)sinl(a));
        }
    };
    emit(16, 16.0L * R1);                 // mid pass 1: Ns = 16, radix R1
    emit(16 * R1, 16.0L * R1 * R2);       // mid pass 2: Ns = 16 R1, radix R2
    emit(N / 16, (long double)N);         // last pass: Ns = N/16, radix 16
    return t;
}

}  // namespace osz
