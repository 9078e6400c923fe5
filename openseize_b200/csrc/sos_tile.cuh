// Single-biquad scan over TILES of a row with a decoupled look-back, for launches that
// cannot fill the GPU with one CTA per row (a 32-channel block of a recording split over
// 8 GPUs; the short look-ahead pass of nm.sosfiltfilt / nm.filtfilt, reference
// core/numerical.py:399-403,508-512) and as a load-balanced alternative when they can.
//
// A row is cut into tiles of 4096 samples.  CTAs are persistent and draw tiles from a
// ticket counter in time-major order (tile t of every row before tile t + 1 of any), one
// draw ahead, so the tiles a tile depends on always belong to CTAs that are running.  A tile
//   1. loads its samples ONCE and scans them from rest (the same thread-level recurrence
//      and Kogge-Stone combine as sos_scan_block, 16 samples per thread);
//   2. publishes its aggregate (the state it would leave behind from rest);
//   3. looks back (warp 0, one predecessor per lane): the state entering the tile is
//         s_in = sum_j Phi^j agg(t-1-j)  +  Phi^d incl(t-1-d),   Phi = A^4096,
//      over the aggregates of the predecessors up to the nearest one whose INCLUSIVE state
//      (its true leaving state) is already published; then publishes its own inclusive
//      state Phi s_in + agg;
//   4. adds the zero-input response of each thread's true entering state
//      (rest-state + A^(16 p) s_in) to its 16 outputs and stores them.
// Every sample is read once and written once whatever the number of rows, with no warm-up
// re-filtering (the time split of sos_scan_kernel) and no second pass (its exact split).
// Tile 0 is the short one (right-aligned behind virtual zeros, as in sos_scan_kernel) and
// starts from the carried state, so its leaving state is inclusive at once.
// Two builds: sos_tile_tma_kernel (below; float64 samples on 16-byte aligned rows: tiles
// moved by the TMA, the next one in flight while this one is scanned) and sos_tile_kernel
// (any alignment, float32 samples: plain coalesced loads / stores through a transposing
// shared-memory buffer, flagged descriptors).
#pragma once

#include <cuda.h>

#include "sos_core.cuh"

namespace osz {

constexpr int TILE_T = 16;
constexpr int TILE = SOS_NT * TILE_T;       // 4096 samples

// Kernel parameter block of the one-section kernels: the head of SosParams (nsec, one
// section) -- 0.8 KB per launch instead of the 13 KB of the full 16-section block.
struct SosParams1 {
    int nsec;
    int pad_;
    SosSec sec[1];
};

struct SosTileTab {
    double thr[SOS_NT][4];      // A^(16 p): thread p's entering state per unit tile-entering state
    double phi[33][4];          // Phi^j, j = 0 .. 32
};

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned *p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_u32(unsigned *p, unsigned v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// flag values of a tile descriptor
constexpr unsigned TILE_NONE = 0, TILE_AGG = 1, TILE_INCL = 2;

template <bool WRITE, typename TIO>
__global__ void __launch_bounds__(SOS_NT, 4)
sos_tile_kernel(const __grid_constant__ SosParams1 prm1, const SosTileTab *__restrict__ tab,
                const TIO *__restrict__ x, int64_t ldx, int rows, int64_t n_total, int reverse,
                const double *state_in, double *state,   // may be the same array
                TIO *__restrict__ y, int64_t ldy,
                const double *__restrict__ lanepow /* [32][4]: A^(16 (lane + 1)) */,
                unsigned *__restrict__ ticket, unsigned *__restrict__ flag /* [rows][ntile] */,
                double2 *__restrict__ agg, double2 *__restrict__ incl, int ntile,
                int use_zi, double zi0, double zi1 /* start state = zi * first sample */,
                int dynamic /* tickets (any residency) / static round-robin (co-resident grid) */) {
    constexpr int T = TILE_T, LD = T + 1, LOGT = 4;
    extern __shared__ __align__(16) unsigned char tile_smem[];
    TIO *stage0 = reinterpret_cast<TIO *>(tile_smem);  // SOS_NT * LD elements of TIO
    __shared__ double wtot[2][SOS_NT / 32][2];
    __shared__ double carry[SOS_MAXSEC][2];
    __shared__ double s_in[2];
    __shared__ unsigned s_job[2];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // (sos_scan_block only touches sec[0 .. nsec-1] of the block it is handed)
    const SosParams &prm = reinterpret_cast<const SosParams &>(prm1);
    const SosSec &c = prm1.sec[0];
    const unsigned total = (unsigned)rows * (unsigned)ntile;
    const int64_t first_len = n_total - (int64_t)(ntile - 1) * TILE;

    double th[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) th[i] = ldg(&tab->thr[tid][i]);
    const double *lp = lanepow + lane * 4;
    const double lp0 = ldg(lp + 0), lp1 = ldg(lp + 1), lp2 = ldg(lp + 2), lp3 = ldg(lp + 3);

    // Jobs in time-major order, drawn from a ticket counter ONE job ahead (the draw for
    // the next job is issued when this one starts, so its latency is hidden).  Every tile a
    // job waits for has a smaller ticket, hence belongs to a CTA that is running: the oldest
    // unfinished job is always being worked on and waits for nothing unfinished, so the
    // scheme cannot deadlock whatever part of the grid is resident (another kernel may hold
    // SMs).  Drawing further ahead is what must not be done: a CTA then parks tiles behind
    // unrelated work and every tile depending on a parked one waits (measured: a constant
    // 0.43 ms for any row count <= 32 when drawing two ahead).
    // Static mode (dynamic == 0; the launch is COOPERATIVE, so the whole grid is resident):
    // the first draw is the CTA's rank and it walks rank, rank + G, ... -- all CTAs then
    // work on the same round of consecutive tiles at the same time, which is what keeps
    // the look-back waits short when a row has many tiles in flight (few rows).
    if (tid == 0) s_job[0] = atomicAdd(ticket, 1u);
    __syncthreads();
    unsigned job = s_job[0];
    const unsigned G = gridDim.x;
    int k = 0;
    while (job < total) {
        if (tid == 0 && dynamic) s_job[(k + 1) & 1] = atomicAdd(ticket, 1u);
        TIO *buf = stage0;
        const int t = (int)(job / (unsigned)rows);
        const int64_t row = (int64_t)(job - (unsigned)t * (unsigned)rows);
        const int64_t pos0 = t == 0 ? 0 : first_len + (int64_t)(t - 1) * TILE;
        const TIO *xr = x + row * ldx + (reverse ? n_total - 1 : 0);
        TIO *yr = WRITE ? y + row * ldy + (reverse ? n_total - 1 : 0) : nullptr;
        unsigned *fl = flag + row * ntile;
        double2 *ag = agg + row * ntile, *in = incl + row * ntile;
        double v[T];

        if (t == 0) {
            // ---- the short first tile: generic path, from the carried state
            const int off = (int)(TILE - first_len);
            if (tid < 2)
                carry[0][tid] = use_zi ? (tid ? zi1 : zi0) * (double)xr[0] : state_in[row * 2 + tid];
#pragma unroll 8
            for (int e = tid; e < TILE; e += SOS_NT) {
                TIO val = (TIO)0;
                if (e >= off) {
                    const int64_t s = e - off;
                    val = ld_stream(reverse ? xr - s : xr + s);
                }
                buf[(e >> LOGT) * LD + (e & (T - 1))] = val;
            }
            __syncthreads();
#pragma unroll
            for (int i = 0; i < T; ++i) v[i] = (double)buf[tid * LD + i];
            sos_scan_block<T, 0>(prm, v, false, off, carry, wtot, lanepow, tid, lane, warp);
            __syncthreads();                     // carry[] holds the leaving state
            if (tid == 0) {
                const double e0 = carry[0][0], e1 = carry[0][1];
                if (ntile == 1) {
                    state[row * 2 + 0] = e0;
                    state[row * 2 + 1] = e1;
                } else {
                    in[0] = make_double2(e0, e1);
                    __threadfence();
                    st_release_u32(fl, TILE_INCL);
                }
            }
            if (WRITE) {
#pragma unroll
                for (int i = 0; i < T; ++i) buf[tid * LD + i] = (TIO)v[i];
                __syncthreads();
#pragma unroll 4
                for (int e = tid; e < TILE; e += SOS_NT) {
                    const int64_t s = e - off;
                    if (e >= off)
                        st_stream(reverse ? yr - s : yr + s, buf[(e >> LOGT) * LD + (e & (T - 1))]);
                }
            }
        } else {
            // ---- a full tile
            {
                const TIO *src = reverse ? xr - pos0 - tid : xr + pos0 + tid;
                TIO tmp[T];
#pragma unroll
                for (int it = 0; it < T; ++it)
                    tmp[it] = ld_stream(reverse ? src - it * SOS_NT : src + it * SOS_NT);
#pragma unroll
                for (int it = 0; it < T; ++it) {
                    const int e = tid + it * SOS_NT;
                    buf[(e >> LOGT) * LD + (e & (T - 1))] = tmp[it];
                }
            }
            __syncthreads();
#pragma unroll
            for (int i = 0; i < T; ++i) v[i] = (double)buf[tid * LD + i];
            const double b0 = c.b0, b1 = c.b1, b2 = c.b2, na1 = -c.a1, na2 = -c.a2;
            // two independent 8-sample chains from rest
            double za0[2] = {0.0, 0.0}, za1[2] = {0.0, 0.0};
#pragma unroll
            for (int i = 0; i < 8; ++i) {
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const double xi = v[8 * j + i];
                    const double yi = fma(b0, xi, za0[j]);
                    za0[j] = fma(na1, yi, fma(b1, xi, za1[j]));
                    za1[j] = fma(na2, yi, b2 * xi);
                    v[8 * j + i] = yi;
                }
            }
            const double zs0 = za0[0], zs1 = za1[0];
            double f0 = fma(c.A8[0], zs0, c.A8[1] * zs1) + za0[1];
            double f1 = fma(c.A8[2], zs0, c.A8[3] * zs1) + za1[1];
            // warp-inclusive scan with M = A^16: M^(2^k) = {A16, P[0..3]}
#pragma unroll
            for (int kk = 0; kk < 5; ++kk) {
                const double *pm = kk == 0 ? c.A16 : c.P[kk - 1];
                const double g0 = __shfl_up_sync(0xffffffffu, f0, 1 << kk);
                const double g1 = __shfl_up_sync(0xffffffffu, f1, 1 << kk);
                if (lane >= (1 << kk)) {
                    f0 += fma(pm[0], g0, pm[1] * g1);
                    f1 += fma(pm[2], g0, pm[3] * g1);
                }
            }
            if (lane == 31) {
                wtot[0][warp][0] = f0;
                wtot[0][warp][1] = f1;
            }
            __syncthreads();
            const double *qm = c.P[4];           // transition over one warp: A^512
            const bool last = t == ntile - 1;
            if (warp == 0 && (WRITE || last)) {
                // aggregate of the tile = combine of the 8 warp totals
                double a0 = 0.0, a1 = 0.0;
#pragma unroll
                for (int u = 0; u < SOS_NT / 32; ++u) {
                    const double t0 = fma(qm[0], a0, qm[1] * a1) + wtot[0][u][0];
                    const double t1 = fma(qm[2], a0, qm[3] * a1) + wtot[0][u][1];
                    a0 = t0;
                    a1 = t1;
                }
                if (lane == 0 && !last) {
                    ag[t] = make_double2(a0, a1);
                    __threadfence();
                    st_release_u32(fl + t, TILE_AGG);
                }
                // ---- look-back: lane j inspects tile base - j
                double e0 = 0.0, e1 = 0.0;
                double w0 = 1.0, w1 = 0.0, w2 = 0.0, w3 = 1.0;      // Phi^(32 windows)
                int base = t - 1;
                while (true) {
                    const int idx = base - lane;
                    unsigned st = TILE_INCL;
                    unsigned incl_mask, first;
                    while (true) {
                        if (idx >= 0) st = ld_acquire_u32(fl + idx);
                        incl_mask = __ballot_sync(0xffffffffu, st == TILE_INCL);
                        const unsigned ready = __ballot_sync(0xffffffffu, st != TILE_NONE);
                        first = incl_mask ? (unsigned)__ffs((int)incl_mask) - 1u : 32u;
                        const unsigned need = first >= 31u ? 0xffffffffu : ((2u << first) - 1u);
                        if ((ready & need) == need) break;
                        __nanosleep(40);
                    }
                    double c0 = 0.0, c1 = 0.0;
                    if ((unsigned)lane <= first && idx >= 0) {
                        const double2 val =
                            (unsigned)lane == first ? __ldcg(in + idx) : __ldcg(ag + idx);
                        const double *ph = tab->phi[lane];
                        const double p0 = fma(ldg(ph + 0), val.x, ldg(ph + 1) * val.y);
                        const double p1 = fma(ldg(ph + 2), val.x, ldg(ph + 3) * val.y);
                        c0 = fma(w0, p0, w1 * p1);
                        c1 = fma(w2, p0, w3 * p1);
                    }
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) {
                        c0 += __shfl_xor_sync(0xffffffffu, c0, o);
                        c1 += __shfl_xor_sync(0xffffffffu, c1, o);
                    }
                    e0 += c0;
                    e1 += c1;
                    if (first < 32u) break;
                    base -= 32;
                    const double *p32 = tab->phi[32];
                    const double q0 = ldg(p32 + 0), q1 = ldg(p32 + 1), q2 = ldg(p32 + 2),
                                 q3 = ldg(p32 + 3);
                    const double n0 = fma(q0, w0, q1 * w2), n1 = fma(q0, w1, q1 * w3);
                    const double n2 = fma(q2, w0, q3 * w2), n3 = fma(q2, w1, q3 * w3);
                    w0 = n0;
                    w1 = n1;
                    w2 = n2;
                    w3 = n3;
                }
                if (lane == 0) {
                    const double *ph = tab->phi[1];
                    const double i0 = fma(ldg(ph + 0), e0, ldg(ph + 1) * e1) + a0;
                    const double i1 = fma(ldg(ph + 2), e0, ldg(ph + 3) * e1) + a1;
                    if (last) {
                        state[row * 2 + 0] = i0;
                        state[row * 2 + 1] = i1;
                    } else {
                        in[t] = make_double2(i0, i1);
                        __threadfence();
                        st_release_u32(fl + t, TILE_INCL);
                    }
                    s_in[0] = e0;
                    s_in[1] = e1;
                }
            } else if (warp == 0 && lane == 0) {
                // state-only pass, not the last tile: the aggregate is all anybody needs
                double a0 = 0.0, a1 = 0.0;
#pragma unroll
                for (int u = 0; u < SOS_NT / 32; ++u) {
                    const double t0 = fma(qm[0], a0, qm[1] * a1) + wtot[0][u][0];
                    const double t1 = fma(qm[2], a0, qm[3] * a1) + wtot[0][u][1];
                    a0 = t0;
                    a1 = t1;
                }
                ag[t] = make_double2(a0, a1);
                __threadfence();
                st_release_u32(fl + t, TILE_AGG);
            }
            if (WRITE) {
                // rest-state entering this warp, then this thread
                double cw0 = 0.0, cw1 = 0.0;
                for (int u = 0; u < warp; ++u) {
                    const double t0 = fma(qm[0], cw0, qm[1] * cw1) + wtot[0][u][0];
                    const double t1 = fma(qm[2], cw0, qm[3] * cw1) + wtot[0][u][1];
                    cw0 = t0;
                    cw1 = t1;
                }
                const double e0 = f0 + fma(lp0, cw0, lp1 * cw1);
                const double e1 = f1 + fma(lp2, cw0, lp3 * cw1);
                double s0 = __shfl_up_sync(0xffffffffu, e0, 1);
                double s1 = __shfl_up_sync(0xffffffffu, e1, 1);
                if (lane == 0) {
                    s0 = cw0;
                    s1 = cw1;
                }
                __syncthreads();                 // s_in published by warp 0
                const double i0 = s_in[0], i1 = s_in[1];
                double q0 = s0 + fma(th[0], i0, th[1] * i1);
                double q1 = s1 + fma(th[2], i0, th[3] * i1);
#pragma unroll
                for (int i = 0; i < 8; ++i) v[i] = fma(c.g0[i], q0, fma(c.g1[i], q1, v[i]));
                {
                    const double t0 = fma(c.A8[0], q0, c.A8[1] * q1) + zs0;
                    const double t1 = fma(c.A8[2], q0, c.A8[3] * q1) + zs1;
                    q0 = t0;
                    q1 = t1;
                }
#pragma unroll
                for (int i = 0; i < 8; ++i) v[8 + i] = fma(c.g0[i], q0, fma(c.g1[i], q1, v[8 + i]));
#pragma unroll
                for (int i = 0; i < T; ++i) buf[tid * LD + i] = (TIO)v[i];
                __syncthreads();
                TIO *dst = reverse ? yr - pos0 - tid : yr + pos0 + tid;
#pragma unroll
                for (int it = 0; it < T; ++it) {
                    const int e = tid + it * SOS_NT;
                    st_stream(reverse ? dst - it * SOS_NT : dst + it * SOS_NT,
                              buf[(e >> LOGT) * LD + (e & (T - 1))]);
                }
            }
        }
        __syncthreads();            // buf, wtot, s_in free; the next ticket visible
        ++k;
        job = dynamic ? s_job[k & 1] : job + G;
    }
}


// ---------------------------------------------------------------------------------------
// The same scan with the tiles moved by the TMA (float64 samples, 16-byte aligned rows).
//
// A row is presented to the TMA as a matrix of 16-sample (128-byte) lines; one tile is a box
// of 256 lines x 16 samples landed in shared memory with the 128-byte swizzle: the 16-byte
// chunk j of line r sits at chunk j ^ (r & 7).  Thread p owns line p (line 255 - p,
// backwards, for the reversed pass) and reads / writes it with eight conflict-free 128-bit
// accesses -- no transposition through shared memory, no load or store instruction touches
// global memory, and the next tile's box is in flight while this one is scanned.  The
// corrected outputs go back into the tile's own lines and leave through a TMA store.
// Tile descriptors carry no flag: a descriptor is 16 bytes written by ONE vector store, and
// "not published" is the all-ones pattern the scratch is initialised to.
// ---------------------------------------------------------------------------------------

constexpr long long TILE_EMPTY = -1LL;             // all-ones: descriptor not published

__device__ __forceinline__ double2 ld_desc(const double2 *p) {
    double2 v;
    asm volatile("ld.relaxed.gpu.global.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_desc(double2 *p, double a, double b) {
    // a payload that happens to be the all-ones NaN is stored as the canonical NaN
    if (__double_as_longlong(a) == TILE_EMPTY) a = __longlong_as_double(0x7FF8000000000000LL);
    asm volatile("st.relaxed.gpu.global.v2.f64 [%0], {%1, %2};" ::"l"(p), "d"(a), "d"(b) : "memory");
}

__device__ __forceinline__ void tma_load_3d(void *dst_smem, const CUtensorMap *map, int c0, int c1,
                                            int c2, uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
        "[%0], [%1, {%2, %3, %4}], [%5];" ::"r"(smem_u32(dst_smem)),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap *map, int c0, int c1, int c2,
                                             const void *src_smem) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%1, %2, %3}], [%4];" ::
                     "l"(reinterpret_cast<uint64_t>(map)),
                 "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(src_smem))
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() {
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// A thread's line of a staged tile (16 samples, the 16-byte chunks swizzled as the TMA left
// them) into / out of its registers, time-reversed for the backward pass.
//   float64: 128-byte lines, SWIZZLE_128B -- chunk j of line r sits at chunk j ^ (r & 7)
//   float32:  64-byte lines, SWIZZLE_64B  -- chunk j of line r sits at chunk j ^ ((r >> 1) & 3)
// (the 16-byte chunk index is XORed with address bits 7.. of the line)
template <typename TIO>
struct TileLine;
template <>
struct TileLine<double> {
    static __device__ __forceinline__ int key(int line) { return line & 7; }
    static __device__ __forceinline__ void load(const double *ln_, int sw, bool reverse, double (&v)[TILE_T]) {
        const double2 *ln = reinterpret_cast<const double2 *>(ln_);
#pragma unroll
        for (int j = 0; j < TILE_T / 2; ++j) {
            const double2 d = ln[j ^ sw];
            if (reverse) {
                v[TILE_T - 1 - 2 * j] = d.x;
                v[TILE_T - 2 - 2 * j] = d.y;
            } else {
                v[2 * j] = d.x;
                v[2 * j + 1] = d.y;
            }
        }
    }
    static __device__ __forceinline__ void store(double *ln_, int sw, bool reverse, const double (&v)[TILE_T]) {
        double2 *ln = reinterpret_cast<double2 *>(ln_);
#pragma unroll
        for (int j = 0; j < TILE_T / 2; ++j)
            ln[j ^ sw] = reverse ? make_double2(v[TILE_T - 1 - 2 * j], v[TILE_T - 2 - 2 * j])
                                 : make_double2(v[2 * j], v[2 * j + 1]);
    }
};
template <>
struct TileLine<float> {
    static __device__ __forceinline__ int key(int line) { return (line >> 1) & 3; }
    static __device__ __forceinline__ void load(const float *ln_, int sw, bool reverse, double (&v)[TILE_T]) {
        const float4 *ln = reinterpret_cast<const float4 *>(ln_);
#pragma unroll
        for (int j = 0; j < TILE_T / 4; ++j) {
            const float4 d = ln[j ^ sw];
            if (reverse) {
                v[TILE_T - 1 - 4 * j] = (double)d.x;
                v[TILE_T - 2 - 4 * j] = (double)d.y;
                v[TILE_T - 3 - 4 * j] = (double)d.z;
                v[TILE_T - 4 - 4 * j] = (double)d.w;
            } else {
                v[4 * j] = (double)d.x;
                v[4 * j + 1] = (double)d.y;
                v[4 * j + 2] = (double)d.z;
                v[4 * j + 3] = (double)d.w;
            }
        }
    }
    static __device__ __forceinline__ void store(float *ln_, int sw, bool reverse, const double (&v)[TILE_T]) {
        float4 *ln = reinterpret_cast<float4 *>(ln_);
#pragma unroll
        for (int j = 0; j < TILE_T / 4; ++j)
            ln[j ^ sw] = reverse ? make_float4((float)v[TILE_T - 1 - 4 * j], (float)v[TILE_T - 2 - 4 * j],
                                               (float)v[TILE_T - 3 - 4 * j], (float)v[TILE_T - 4 - 4 * j])
                                 : make_float4((float)v[4 * j], (float)v[4 * j + 1],
                                               (float)v[4 * j + 2], (float)v[4 * j + 3]);
    }
};

template <bool WRITE, typename TIO>
__global__ void __launch_bounds__(SOS_NT, 3)
sos_tile_tma_kernel(const __grid_constant__ SosParams1 prm1, const __grid_constant__ CUtensorMap mx,
                    const __grid_constant__ CUtensorMap my, const SosTileTab *__restrict__ tab,
                    const TIO *__restrict__ x, int64_t ldx, int rows, int64_t n_total,
                    int reverse, const double *state_in, double *state, TIO *__restrict__ y,
                    int64_t ldy, const double *__restrict__ lanepow, unsigned *__restrict__ ticket,
                    double2 *__restrict__ agg, double2 *__restrict__ incl, int ntile,
                    int use_zi, double zi0, double zi1 /* start state = zi * first sample */,
                    int dynamic /* tickets (any residency) / static round-robin (co-resident grid) */) {
    constexpr int T = TILE_T;
    constexpr int TILE_BYTES = TILE * (int)sizeof(TIO);      // 32 KB / 16 KB per stage
    extern __shared__ unsigned char tile_smem_raw[];
    __shared__ __align__(8) uint64_t full[2];
    __shared__ double wtot[2][SOS_NT / 32][2];
    __shared__ double carry[SOS_MAXSEC][2];
    __shared__ double s_in[2];
    __shared__ unsigned s_job[2];
    __shared__ double2 early[2][32];      // warp 0: descriptors polled before the scan

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // (sos_scan_block only touches sec[0 .. nsec-1] of the block it is handed)
    const SosParams &prm = reinterpret_cast<const SosParams &>(prm1);
    const SosSec &c = prm1.sec[0];
    const unsigned total = (unsigned)rows * (unsigned)ntile;
    const int64_t first_len = n_total - (int64_t)(ntile - 1) * TILE;
    // two 1024-byte aligned stages
    unsigned char *stage0 = tile_smem_raw + ((1024u - (smem_u32(tile_smem_raw) & 1023u)) & 1023u);

    double th[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) th[i] = ldg(&tab->thr[tid][i]);
    const double *lp = lanepow + lane * 4;
    const double lp0 = ldg(lp + 0), lp1 = ldg(lp + 1), lp2 = ldg(lp + 2), lp3 = ldg(lp + 3);

    // this thread's line of a tile and the swizzle of its chunks
    const int line = reverse ? SOS_NT - 1 - tid : tid;
    const int sw = TileLine<TIO>::key(line);

    constexpr int PRODUCER = 32;                      // warp 1, lane 0: issues every TMA copy
    auto tile_coord = [&](int t) { return (reverse ? ntile - 1 - t : t - 1) * SOS_NT; };
    auto issue_load = [&](unsigned jb, int stg) {     // PRODUCER only
        if (jb >= total) return;
        const int t = (int)(jb / (unsigned)rows);
        if (t == 0) return;
        const int row = (int)(jb - (unsigned)t * (unsigned)rows);
        fence_proxy_async();
        mbar_expect_tx(&full[stg], (uint32_t)TILE_BYTES);
        tma_load_3d(stage0 + (size_t)stg * TILE_BYTES, &mx, 0, tile_coord(t), row, &full[stg]);
    };

    // tickets one job ahead (see sos_tile_kernel); the counter starts at all-ones
    if (tid == 0) s_job[0] = atomicAdd(ticket, 1u) + 1u;
    if (tid == PRODUCER) {
        mbar_init(&full[0], 1);
        mbar_init(&full[1], 1);
        fence_barrier_init();
    }
    __syncthreads();
    unsigned job = s_job[0];
    const unsigned G = gridDim.x;
    if (tid == PRODUCER) issue_load(job, 0);
    unsigned phase = 0;                               // bit s: parity to wait for on stage s
    int k = 0;
    while (job < total) {
        const int stg = k & 1;
        TIO *sb = reinterpret_cast<TIO *>(stage0 + (size_t)stg * TILE_BYTES);
        const int t = (int)(job / (unsigned)rows);
        const int64_t row = (int64_t)(job - (unsigned)t * (unsigned)rows);
        if (tid == 0)                                   // the next job; read behind a barrier
            s_job[(k + 1) & 1] = dynamic ? atomicAdd(ticket, 1u) + 1u : job + G;
        double2 *ag = agg + row * ntile, *in = incl + row * ntile;
        double v[T];

        if (t == 0) {
            // ---- the short first tile: generic path from the carried state (plain loads and
            //      stores; lines unswizzled, 16 samples per thread line)
            const TIO *xr = x + row * ldx + (reverse ? n_total - 1 : 0);
            const int off = (int)(TILE - first_len);
            if (tid < 2)
                carry[0][tid] = use_zi ? (tid ? zi1 : zi0) * (double)xr[0] : state_in[row * 2 + tid];
#pragma unroll 8
            for (int e = tid; e < TILE; e += SOS_NT) {
                TIO val = (TIO)0;
                if (e >= off) {
                    const int64_t s = e - off;
                    val = ld_stream(reverse ? xr - s : xr + s);
                }
                sb[e] = val;
            }
            __syncthreads();
            if (tid == PRODUCER) {
                bulk_wait_read0();
                issue_load(s_job[(k + 1) & 1], stg ^ 1);
            }
#pragma unroll
            for (int i = 0; i < T; ++i) v[i] = (double)sb[tid * T + i];
            sos_scan_block<T, 0>(prm, v, false, off, carry, wtot, lanepow, tid, lane, warp);
            __syncthreads();                     // carry[] holds the leaving state
            if (tid == 0) {
                const double e0 = carry[0][0], e1 = carry[0][1];
                if (ntile == 1) {
                    state[row * 2 + 0] = e0;
                    state[row * 2 + 1] = e1;
                } else {
                    st_desc(in, e0, e1);
                }
            }
            if (WRITE) {
                TIO *yr = y + row * ldy + (reverse ? n_total - 1 : 0);
#pragma unroll
                for (int i = 0; i < T; ++i) sb[tid * T + i] = (TIO)v[i];
                __syncthreads();
#pragma unroll 4
                for (int e = tid; e < TILE; e += SOS_NT) {
                    const int64_t s = e - off;
                    if (e >= off) st_stream(reverse ? yr - s : yr + s, sb[e]);
                }
            }
            __syncthreads();
        } else {
            // ---- a full tile: its box was fetched while the previous tile ran
            if (warp == 0 && (WRITE || t == ntile - 1)) {
                // first poll of the look-back now, so that its round trip to L2 runs under
                // the scan: predecessors dealt a round earlier have published long ago
                const int idx = t - 1 - lane;
                double2 I = make_double2(0.0, 0.0), A = make_double2(0.0, 0.0);
                if (idx >= 0) {
                    I = ld_desc(in + idx);
                    A = ld_desc(ag + idx);
                }
                early[0][lane] = I;
                early[1][lane] = A;
            }
            mbar_wait(&full[stg], (phase >> stg) & 1u);
            phase ^= 1u << stg;
            TileLine<TIO>::load(sb + line * T, sw, reverse != 0, v);
            const double b0 = c.b0, b1 = c.b1, b2 = c.b2, na1 = -c.a1, na2 = -c.a2;
            // two independent 8-sample chains from rest
            double za0[2] = {0.0, 0.0}, za1[2] = {0.0, 0.0};
#pragma unroll
            for (int i = 0; i < 8; ++i) {
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const double xi = v[8 * j + i];
                    const double yi = fma(b0, xi, za0[j]);
                    za0[j] = fma(na1, yi, fma(b1, xi, za1[j]));
                    za1[j] = fma(na2, yi, b2 * xi);
                    v[8 * j + i] = yi;
                }
            }
            const double zs0 = za0[0], zs1 = za1[0];
            double f0 = fma(c.A8[0], zs0, c.A8[1] * zs1) + za0[1];
            double f1 = fma(c.A8[2], zs0, c.A8[3] * zs1) + za1[1];
            // warp-inclusive scan with M = A^16: M^(2^k) = {A16, P[0..3]}
#pragma unroll
            for (int kk = 0; kk < 5; ++kk) {
                const double *pm = kk == 0 ? c.A16 : c.P[kk - 1];
                const double g0 = __shfl_up_sync(0xffffffffu, f0, 1 << kk);
                const double g1 = __shfl_up_sync(0xffffffffu, f1, 1 << kk);
                if (lane >= (1 << kk)) {
                    f0 += fma(pm[0], g0, pm[1] * g1);
                    f1 += fma(pm[2], g0, pm[3] * g1);
                }
            }
            double (*wt)[2] = wtot[k & 1];
            if (lane == 31) {
                wt[warp][0] = f0;
                wt[warp][1] = f1;
            }
            __syncthreads();
            if (tid == PRODUCER) {
                // the other stage: its previous tile's store must have read it out
                bulk_wait_read0();
                issue_load(s_job[(k + 1) & 1], stg ^ 1);
            }
            const double *qm = c.P[4];           // transition over one warp: A^512
            const bool last = t == ntile - 1;
            if (warp == 0 && (WRITE || last)) {
                // aggregate of the tile = combine of the 8 warp totals
                double a0 = 0.0, a1 = 0.0;
#pragma unroll
                for (int u = 0; u < SOS_NT / 32; ++u) {
                    const double t0 = fma(qm[0], a0, qm[1] * a1) + wt[u][0];
                    const double t1 = fma(qm[2], a0, qm[3] * a1) + wt[u][1];
                    a0 = t0;
                    a1 = t1;
                }
                if (lane == 0 && !last) st_desc(ag + t, a0, a1);
                // ---- look-back: lane j inspects tile base - j
                double e0 = 0.0, e1 = 0.0;
                double w0 = 1.0, w1 = 0.0, w2 = 0.0, w3 = 1.0;      // Phi^(32 windows)
                int base = t - 1;
                bool polled = true;              // the first window was polled before the scan
                while (true) {
                    const int idx = base - lane;
                    double2 I = make_double2(0.0, 0.0), A = make_double2(0.0, 0.0);
                    unsigned first;
                    while (true) {
                        bool has_i = true, has_a = true;
                        if (idx >= 0) {
                            if (polled) {
                                I = early[0][lane];
                                A = early[1][lane];
                            } else {
                                I = ld_desc(in + idx);
                                A = ld_desc(ag + idx);
                            }
                            has_i = __double_as_longlong(I.x) != TILE_EMPTY;
                            has_a = has_i || __double_as_longlong(A.x) != TILE_EMPTY;
                        }
                        polled = false;
                        const unsigned incl_mask = __ballot_sync(0xffffffffu, has_i);
                        const unsigned ready = __ballot_sync(0xffffffffu, has_a);
                        first = incl_mask ? (unsigned)__ffs((int)incl_mask) - 1u : 32u;
                        const unsigned need = first >= 31u ? 0xffffffffu : ((2u << first) - 1u);
                        if ((ready & need) == need) break;
                        __nanosleep(20);
                    }
                    double c0 = 0.0, c1 = 0.0;
                    if ((unsigned)lane <= first && idx >= 0) {
                        const double2 val = (unsigned)lane == first ? I : A;
                        const double *ph = tab->phi[lane];
                        const double p0 = fma(ldg(ph + 0), val.x, ldg(ph + 1) * val.y);
                        const double p1 = fma(ldg(ph + 2), val.x, ldg(ph + 3) * val.y);
                        c0 = fma(w0, p0, w1 * p1);
                        c1 = fma(w2, p0, w3 * p1);
                    }
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) {
                        c0 += __shfl_xor_sync(0xffffffffu, c0, o);
                        c1 += __shfl_xor_sync(0xffffffffu, c1, o);
                    }
                    e0 += c0;
                    e1 += c1;
                    if (first < 32u) break;
                    base -= 32;
                    const double *p32 = tab->phi[32];
                    const double q0 = ldg(p32 + 0), q1 = ldg(p32 + 1), q2 = ldg(p32 + 2),
                                 q3 = ldg(p32 + 3);
                    const double n0 = fma(q0, w0, q1 * w2), n1 = fma(q0, w1, q1 * w3);
                    const double n2 = fma(q2, w0, q3 * w2), n3 = fma(q2, w1, q3 * w3);
                    w0 = n0;
                    w1 = n1;
                    w2 = n2;
                    w3 = n3;
                }
                if (lane == 0) {
                    const double *ph = tab->phi[1];
                    const double i0 = fma(ldg(ph + 0), e0, ldg(ph + 1) * e1) + a0;
                    const double i1 = fma(ldg(ph + 2), e0, ldg(ph + 3) * e1) + a1;
                    if (last) {
                        state[row * 2 + 0] = i0;
                        state[row * 2 + 1] = i1;
                    } else {
                        st_desc(in + t, i0, i1);
                    }
                    s_in[0] = e0;
                    s_in[1] = e1;
                }
            } else if (warp == 0 && lane == 0) {
                // state-only pass, not the last tile: the aggregate is all anybody needs
                double a0 = 0.0, a1 = 0.0;
#pragma unroll
                for (int u = 0; u < SOS_NT / 32; ++u) {
                    const double t0 = fma(qm[0], a0, qm[1] * a1) + wt[u][0];
                    const double t1 = fma(qm[2], a0, qm[3] * a1) + wt[u][1];
                    a0 = t0;
                    a1 = t1;
                }
                st_desc(ag + t, a0, a1);
            }
            if (WRITE) {
                // rest-state entering this warp, then this thread
                double cw0 = 0.0, cw1 = 0.0;
                for (int u = 0; u < warp; ++u) {
                    const double t0 = fma(qm[0], cw0, qm[1] * cw1) + wt[u][0];
                    const double t1 = fma(qm[2], cw0, qm[3] * cw1) + wt[u][1];
                    cw0 = t0;
                    cw1 = t1;
                }
                const double e0 = f0 + fma(lp0, cw0, lp1 * cw1);
                const double e1 = f1 + fma(lp2, cw0, lp3 * cw1);
                double s0 = __shfl_up_sync(0xffffffffu, e0, 1);
                double s1 = __shfl_up_sync(0xffffffffu, e1, 1);
                if (lane == 0) {
                    s0 = cw0;
                    s1 = cw1;
                }
                __syncthreads();                 // s_in published by warp 0
                const double i0 = s_in[0], i1 = s_in[1];
                double q0 = s0 + fma(th[0], i0, th[1] * i1);
                double q1 = s1 + fma(th[2], i0, th[3] * i1);
#pragma unroll
                for (int i = 0; i < 8; ++i) v[i] = fma(c.g0[i], q0, fma(c.g1[i], q1, v[i]));
                {
                    const double t0 = fma(c.A8[0], q0, c.A8[1] * q1) + zs0;
                    const double t1 = fma(c.A8[2], q0, c.A8[3] * q1) + zs1;
                    q0 = t0;
                    q1 = t1;
                }
#pragma unroll
                for (int i = 0; i < 8; ++i) v[8 + i] = fma(c.g0[i], q0, fma(c.g1[i], q1, v[8 + i]));
                // back into this thread's own line (nobody else touches it)
                TileLine<TIO>::store(sb + line * T, sw, reverse != 0, v);
                fence_proxy_async();
                __syncthreads();
                if (tid == PRODUCER) {
                    tma_store_3d(&my, 0, tile_coord(t), (int)row, sb);
                    bulk_commit();
                }
            }
        }
        ++k;
        job = s_job[k & 1];
    }
    if (tid == PRODUCER) bulk_wait0();
}

}  // namespace osz
