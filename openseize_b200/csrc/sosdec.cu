// Fused IIR pass + decimating FIR: the last pass of a
// forward-backward biquad filter (or a forward pass) whose full-rate output
// never leaves the SM -- it is written into a shared-memory ring and consumed
// there by the decimating filter running on the FP64 tensor cores.  Replaces,
// for the chain  IIR(dephase) -> FIR('same') -> downsample  of the reference's
// tools/pipeline.py:109-124 composition, the backward scipy.signal.sosfilt /
// lfilter call of nm.sosfiltfilt / nm.filtfilt (core/numerical.py:402,410,511,519)
// followed by the oaconvolve and resample_poly calls of the two stages after it
// (:229-269, :610-631): 8 + 8/M bytes per sample of HBM traffic instead of
// 16 (pass) + 8 + 8/M (decimator).
//
// One CTA owns one (row, time span).  Two warp groups run concurrently, coupled by two
// shared-memory counters (ring fill level, tiles finished):
//   * scan group (256 threads): the time-parallel biquad scan of sos_core.cuh
//     (16 samples per thread) over the next block, in PROCESSING order (reverse
//     time for a backward pass); outputs go to the ring, and the first / last
//     K-1 samples of the span also to a small edge buffer in global memory;
//   * FIR group (WT * KS warps): the banded-Toeplitz DMMA product of upfirdn.cu
//     over every tile of 8*S outputs whose window the scan has completed.
// Outputs whose window straddles a span or chunk boundary are computed from the
// edge buffers by sosdec_boundary_kernel.
#include <stdlib.h>

#include <vector>

#include "common.cuh"
#include "sos_core.cuh"
#include "ufd_mma.cuh"

namespace osz {

constexpr int SD_T = 16;                       // samples per scan thread
constexpr int SD_BLK = SOS_NT * SD_T;          // samples per scan block
constexpr int SD_LD = SD_T + 1;

struct SosDecGeom {
    UfdMmaGeom m;
    int nring;        // ring capacity in segments of SM samples (pitch P)
};

__device__ __forceinline__ int64_t floordiv64(int64_t a, int64_t b) {   // b > 0
    int64_t q = a / b;
    if ((a % b) != 0 && a < 0) --q;
    return q;
}

template <int WT, int KS>
__global__ void __launch_bounds__(SOS_NT + WT * KS * 32, 1)
sosdec_kernel(const __grid_constant__ SosParams prm, const __grid_constant__ SosDecGeom gd,
              const double *__restrict__ x, int64_t ldx, int64_t n_total, int reverse,
              const double *__restrict__ state_in, double *__restrict__ state,
              const double *__restrict__ lanepow, int64_t span_len, int64_t settle,
              const double *__restrict__ gpad,
              int64_t abase /* reverse: n_total-1 + A - half;  forward: half - (K-1) - A */,
              double *__restrict__ y, int64_t ldy, int64_t out_first, int64_t n_out,
              double *__restrict__ edges /* [rows][nspan][2][K-1], real-time order */) {
    constexpr int NFIR = WT * KS * 32;
    extern __shared__ __align__(16) double smem_sd[];
    __shared__ double wtot[2][SOS_NT / 32][2];
    __shared__ double carry[SOS_MAXSEC][2];
    __shared__ int64_t s_avail, s_tiles;     // ring fill level (scan -> FIR), tiles finished (FIR -> scan)

    const UfdMmaGeom &gm = gd.m;
    const int M = gm.M, SM = gm.SM, P = gm.P, S = gm.S, K = gm.K, nring = gd.nring;
    double *buf = smem_sd;                                   // SOS_NT * SD_LD
    double *ring = buf + SOS_NT * SD_LD;                     // nring * P
    double *gs = ring + (size_t)nring * P;                   // M * ldq
    double *red = gs + (size_t)M * gm.ldq;                   // 2 x KS x 8 S partial sums

    const int tid = threadIdx.x;
    const bool scan_role = tid < SOS_NT;
    const int64_t row = blockIdx.x;
    const int nsec = prm.nsec;

    // ---- span of this CTA, in logical (processing-order) sample indices
    const int64_t span = blockIdx.y;
    const int64_t a = span * span_len;
    int64_t b = a + span_len;
    if (b > n_total) b = n_total;
    const int64_t w0 = span == 0 ? a : a - settle;      // later spans warm up from rest
    const int64_t n = b - w0;
    const int64_t keep = a - w0;
    // ---- decimator geometry of the span: output j' = 0 .. nout-1 (processing order)
    //      reads ring samples u = j'*M .. j'*M + K-1,  u = (logical index) - a - doff
    int64_t jedge;          // reverse: largest j of the span; forward: smallest
    int doff;
    if (reverse) {
        jedge = floordiv64(abase - a, M);
        doff = (int)(abase - a - jedge * M);
    } else {
        jedge = -floordiv64(-(a - abase), M);            // ceil((a - abase) / M)
        doff = (int)(jedge * M + abase - a);
    }
    const int64_t room = (b - a) - K - doff;
    const int64_t nout = room >= 0 ? room / M + 1 : 0;
    // real-time bounds of the span inside the chunk (for the edge buffers)
    const int64_t ua = reverse ? n_total - b : a;
    const int64_t ub = reverse ? n_total - a : b;
    double *edge_lo = edges + ((row * gridDim.y + span) * 2 + 0) * (int64_t)(K - 1);
    double *edge_hi = edge_lo + (K - 1);

    const int64_t nblk = (n + SD_BLK - 1) / SD_BLK;
    const int64_t first_len = n - (nblk - 1) * SD_BLK;

    // ---- set-up
    for (int i = tid; i < nring * P; i += blockDim.x) ring[i] = 0.0;
    for (int i = tid; i < M * gm.ldq; i += blockDim.x) gs[i] = gpad[i];
    if (tid < nsec * 2) {
        double c0 = 0.0;
        if (span == 0) c0 = state_in[row * nsec * 2 + tid];
        carry[tid >> 1][tid & 1] = c0;
    }
    if (tid == 0) {
        s_avail = -((int64_t)1 << 60);
        s_tiles = 0;
    }
    __syncthreads();

    // scan role state
    const int lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const double *xr = x + row * ldx + (reverse ? n_total - 1 - w0 : w0);
    // FIR role state
    const int fw = warp - SOS_NT / 32;                       // FIR warp index (>= 0 for that group)
    const int wt = fw % WT, wk = fw / WT;
    const int g = lane >> 2, q = lane & 3;
    const int oa = g * S + 32 * wt + 2 * q;
    const int k_lo = (int)(((long)gm.ktotal * wk) / KS), k_hi = (int)(((long)gm.ktotal * (wk + 1)) / KS);
    const int64_t ntile = (nout + 8 * S - 1) / (8 * S);

    // ---- the two groups run decoupled: the scan publishes how far the ring is
    //      filled (s_avail, in ring coordinates) and waits only when it would
    //      overwrite segments a pending tile still reads (s_tiles: tiles finished).
    if (scan_role) {
        // the loads of block blk + 1 are in flight while block blk is scanned (`nxt`)
        double nxt[SD_T];
        auto prefetch = [&](int64_t blk) {     // full blocks only (blk >= 1)
            const int64_t pos0 = first_len + (blk - 1) * SD_BLK;
            const double *src = reverse ? xr - pos0 - tid : xr + pos0 + tid;
#pragma unroll
            for (int k = 0; k < SD_T; ++k)
                nxt[k] = ld_stream(reverse ? src - k * SOS_NT : src + k * SOS_NT);
        };
        for (int64_t blk = 0; blk < nblk; ++blk) {
            const int off = blk == 0 ? (int)(SD_BLK - first_len) : 0;
            const int64_t pos0 = blk == 0 ? 0 : first_len + (blk - 1) * SD_BLK;
            if (blk != 0) {
#pragma unroll
                for (int k = 0; k < SD_T; ++k) {
                    const int e = tid + k * SOS_NT;
                    buf[(e >> 4) * SD_LD + (e & (SD_T - 1))] = nxt[k];
                }
            } else {
#pragma unroll 4
                for (int e = tid; e < SD_BLK; e += SOS_NT) {
                    double val = 0.0;
                    if (e >= off) {
                        const int64_t s = pos0 + (e - off);
                        val = ld_stream(reverse ? xr - s : xr + s);
                    }
                    buf[(e >> 4) * SD_LD + (e & (SD_T - 1))] = val;
                }
            }
            if (blk + 1 < nblk) prefetch(blk + 1);
            named_bar_sync<1>(SOS_NT);
            double v[SD_T];
#pragma unroll
            for (int i = 0; i < SD_T; ++i) v[i] = buf[tid * SD_LD + i];
            sos_scan_block<SD_T, 1>(prm, v, blk != 0, off, carry, wtot, lanepow, tid, lane, warp);
#pragma unroll
            for (int i = 0; i < SD_T; ++i) buf[tid * SD_LD + i] = v[i];
            // ring coordinates of the block: u = s - keep - doff, s = pos0 + e - off
            const int64_t u_end = pos0 + (SD_BLK - off) - keep - doff;        // exclusive
            if (tid == 0 && u_end > 0) {
                // the slot of the newest segment written held segment sig - nring, last read
                // by tile floor((sig - nring) / 8): wait until that tile is finished
                const int64_t sig_hi = (u_end - 1) / SM;
                int64_t need_done = sig_hi - nring >= 0 ? (sig_hi - nring) / 8 + 1 : 0;
                if (need_done > ntile) need_done = ntile;
                while (*(volatile int64_t *)&s_tiles < need_done) __nanosleep(200);
                __threadfence_block();
            }
            named_bar_sync<1>(SOS_NT);
            // ---- hand the block over: ring (time-contiguous, padded segments) + edges
            // (most blocks lie wholly inside the span's interior: no edge samples, no
            //  warm-up samples, nothing before the first output's window -- plain copy)
            const int64_t tl_a = reverse ? n_total - (w0 + pos0 + SD_BLK - off) : w0 + pos0;
            const int64_t tl_b = tl_a + (SD_BLK - off);             // real-time [tl_a, tl_b)
            const bool plain = blk != 0 && pos0 >= keep && pos0 - keep - doff >= 0 &&
                               tl_a >= ua + (K - 1) && tl_b <= ub - (K - 1);
            if (plain) {
                const int64_t u0 = pos0 + (int64_t)tid - keep - doff;
                const int64_t sig = u0 / SM;
                int w = (int)(u0 - sig * SM);
                int slot = (int)(sig % nring);
#pragma unroll
                for (int k = 0; k < SD_T; ++k) {
                    const int e = tid + k * SOS_NT;
                    ring[slot * P + w] = buf[(e >> 4) * SD_LD + (e & (SD_T - 1))];
                    w += SOS_NT;
                    if (w >= SM) {
                        w -= SM;
                        if (++slot == nring) slot = 0;
                    }
                }
            } else if (pos0 + (SD_BLK - off) > keep) {
                const int64_t u0 = pos0 + (int64_t)tid - off - keep - doff;
                int64_t sig = floordiv64(u0, SM);
                int w = (int)(u0 - sig * SM);
                int slot = (int)(sig % nring);
                if (slot < 0) slot += nring;
#pragma unroll 4
                for (int k = 0; k < SD_T; ++k) {
                    const int e = tid + k * SOS_NT;
                    const int64_t s = pos0 + (e - off);
                    if (e >= off && s >= keep) {
                        const double val = buf[(e >> 4) * SD_LD + (e & (SD_T - 1))];
                        if (sig >= 0) ring[slot * P + w] = val;
                        const int64_t tl = reverse ? n_total - 1 - (w0 + s) : w0 + s;
                        const int64_t dl = tl - ua, dh = tl - (ub - (K - 1));
                        if (dl < K - 1) edge_lo[dl] = val;
                        if (dh >= 0) edge_hi[dh] = val;
                    }
                    w += SOS_NT;
                    if (w >= SM) {
                        w -= SM;
                        ++sig;
                        if (++slot == nring) slot = 0;
                    }
                }
            }
            named_bar_sync<1>(SOS_NT);             // every scan thread's ring writes are done
            if (tid == 0) {
                __threadfence_block();
                *(volatile int64_t *)&s_avail = blk + 1 == nblk ? (int64_t)1 << 60 : u_end;
            }
        }
    } else {
        for (int64_t tile = 0; tile < ntile; ++tile) {
            const int64_t need = tile * 8 * (int64_t)SM + gm.total_len;
            // one thread polls the fill level; the other FIR warps wait in the barrier
            if (tid == SOS_NT) {
                while (*(volatile int64_t *)&s_avail < need) __nanosleep(200);
                __threadfence_block();
            }
            named_bar_sync<3>(NFIR);
            // ---- banded Toeplitz product of the tile on the tensor cores
            const int segbase = (int)((tile * 8 + g) % nring);
            auto fetch = [&](int p, int n) {   // sample n of phase p in row g's window
                int seg = segbase + (n >> 5);
                if (seg >= nring) seg -= nring;
                return ring[seg * P + p + (n & 31) * M];
            };
            double c[8];
            ufd_mma_ksteps(gm, gs, k_lo, k_hi, wt, g, q, fetch, c);
            // partial sums of the k-splits to shared memory (double buffered by tile
            // parity); every FIR thread then sums and stores its share of the outputs
            double *rd = red + ((size_t)(tile & 1) * KS + wk) * 8 * S;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                rd[oa + 8 * j] = c[2 * j];
                rd[oa + 8 * j + 1] = c[2 * j + 1];
            }
            named_bar_sync<2>(NFIR);               // every FIR warp is done with the tile's window
            if (tid == SOS_NT) {
                __threadfence_block();
                *(volatile int64_t *)&s_tiles = tile + 1;
            }
            const double *r0 = red + (size_t)(tile & 1) * KS * 8 * S;
            double *yr = y + row * ldy;
            for (int o = tid - SOS_NT; o < 8 * S; o += NFIR) {
                double sum = r0[o];
#pragma unroll
                for (int k = 1; k < KS; ++k) sum += r0[(size_t)k * 8 * S + o];
                const int64_t j1 = tile * 8 * S + o;
                if (j1 < nout) {
                    const int64_t col = (reverse ? jedge - j1 : jedge + j1) - out_first;
                    if (col >= 0 && col < n_out) yr[col] = sum;
                }
            }
        }
    }
    __syncthreads();
    if (tid < nsec * 2 && span == gridDim.y - 1)
        state[row * nsec * 2 + tid] = carry[tid >> 1][tid & 1];
}

// Outputs whose K-tap window straddles a boundary between two pieces (time spans of
// a chunk, or two chunks): window sample i of boundary time T is
//   i < K-1 ? tail[i] : head[i - (K-1)]      (real-time order, T - (K-1) + i)
// tail = last K-1 samples of the piece before T, head = first K-1 of the piece after;
// either may be null (recording edge: zeros).  One warp per output.
__global__ void sosdec_boundary_kernel(const double *__restrict__ taps /* g[jj], K */, int K, int M,
                                       int half, const double *__restrict__ edges, int nspan,
                                       int reverse, const double *__restrict__ prev_tail,
                                       int64_t prev_ld, int has_end, int64_t A, int64_t n_total,
                                       int64_t span_len, double *__restrict__ y, int64_t ldy,
                                       int64_t out_first, int64_t n_out, int64_t j_min,
                                       int64_t j_max) {
    const int bnd = blockIdx.x;                 // 0 .. nspan (nspan = chunk end, only with has_end)
    const int64_t row = blockIdx.y;
    if (bnd == nspan && !has_end) return;
    // real-time piece r <-> logical span: reverse ? nspan-1-r : r; piece r starts at
    //   forward: r * span_len           reverse: n_total - min((nspan - r) * span_len, n_total)
    auto piece_start = [&](int r) -> int64_t {
        if (!reverse) return (int64_t)r * span_len;
        int64_t e = (int64_t)(nspan - r) * span_len;
        if (e > n_total) e = n_total;
        return n_total - e;
    };
    auto edge_ptr = [&](int r, int which) -> const double * {
        const int sp = reverse ? nspan - 1 - r : r;
        return edges + ((row * nspan + sp) * 2 + which) * (int64_t)(K - 1);
    };
    const int64_t T = bnd == nspan ? n_total : piece_start(bnd);
    const double *tail = bnd == 0 ? (prev_tail ? prev_tail + row * prev_ld : nullptr)
                                  : edge_ptr(bnd - 1, 1);
    const double *head = bnd == nspan ? nullptr : edge_ptr(bnd, 0);
    const int64_t Tg = A + T;
    // outputs with  j*M + half - (K-1) < Tg <= j*M + half
    int64_t j0 = -floordiv64(-(Tg - half), M);                   // ceil((Tg - half) / M)
    int64_t j1 = -floordiv64(-(Tg - half + K - 1), M) - 1;
    if (j0 < j_min) j0 = j_min;
    if (j1 > j_max) j1 = j_max;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    for (int64_t j = j0 + warp; j <= j1; j += nwarp) {
        const int64_t col = j - out_first;
        if (col < 0 || col >= n_out) continue;
        const int base = (int)(j * M + half - Tg);               // window index of tap 0
        double acc = 0.0;
        for (int jj = lane; jj < K; jj += 32) {
            const int i = base + jj;
            double v = 0.0;
            if (i < K - 1) {
                if (tail) v = tail[i];
            } else if (head) {
                v = head[i - (K - 1)];
            }
            acc = fma(taps[jj], v, acc);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) y[row * ldy + col] = acc;
    }
}

}  // namespace osz

using namespace osz;

// (defined in sos.cu / upfirdn.cu)
extern "C" int osz_sos_plan_params(const osz_sos_plan *p, SosParams *prm, const double **lanepow,
                                   int64_t *settle);
extern "C" int osz_upfirdn_plan_taps(const osz_upfirdn_plan *p, int *K, int *M, int *half,
                                     const double **d_gpad_fwd, const double **d_gpad_rev,
                                     const double **d_taps_g, int *ldq);

namespace {

// Tile geometry of the fused kernel for K taps decimated by M, or false when no
// segment length fits (the caller then runs the two kernels separately).
bool sosdec_geometry(int K, int M, int ksplit, SosDecGeom *gd, int *wt_out, size_t *smem_out) {
    int smax = 0;
    long total_steps = 0;
    const std::vector<int> ks = ufd_ksteps(K, M, &smax, &total_steps);
    for (int S : {32}) {
        const int WT = S / 32;
        const int SM = S * M;
        if (SM < SOS_NT) continue;                 // the ring writer advances one segment at most
        const int nmax = 32 * (WT - 1) + 4 * (smax + 6) + 3;
        const int row_len = (M - 1) + nmax * M + 1;
        if (row_len > 4 * SM) continue;
        const int total_len = 7 * SM + row_len;
        const int P = SM + ufd_best_pad(SM, M);
        const int nring = (2 * SD_BLK + total_len + SM - 1) / SM + 2;
        const int ldq = 7 + 4 * (smax + 1) + 4;
        const size_t smem = ((size_t)SOS_NT * SD_LD + (size_t)nring * P + (size_t)M * ldq +
                             2 * (size_t)ksplit * 8 * S) * 8;
        if (smem > 224 * 1024) continue;
        UfdMmaGeom &gm = gd->m;
        gm.K = K;
        gm.M = M;
        gm.half = (K - 1) / 2;
        gm.S = S;
        gm.logS = 5;
        gm.SM = SM;
        gm.P = P;
        gm.total_len = total_len;
        gm.ldq = ldq;
        gm.p_rem = (K - 1) % M;
        gm.ks_hi = ks[0];
        gm.ks_lo = ks[M - 1];
        gm.ktotal = (int)total_steps;
        gd->nring = nring;
        *wt_out = WT;
        *smem_out = smem;
        return true;
    }
    return false;
}

int sosdec_ksplit(int) {
    static const int forced = [] {
        const char *e = getenv("OSZ_SOSDEC_KS");
        const int v = e ? atoi(e) : 8;
        return v == 4 || v == 8 ? v : 8;
    }();
    return forced;
}

}  // namespace

extern "C" {

// Spans per row the fused kernel would use for `rows` rows of `n` samples (>= 1), or
// 0 when the pair of plans cannot run fused (no tile geometry, span shorter than the
// filter).  One CTA per SM: pick the span count that fills the last wave best, counting
// the warm-up every later span pays.
int osz_sosdec_spans(const osz_sos_plan *sos, const osz_upfirdn_plan *ufd, int64_t rows, int64_t n) {
    if (!sos || !ufd || rows <= 0 || n <= 0) return 0;
    SosParams prm;
    const double *lanepow = nullptr;
    int64_t settle = -1;
    if (osz_sos_plan_params(sos, &prm, &lanepow, &settle) != OSZ_OK) return 0;
    int K, M, half, ldq;
    const double *gf, *gr, *tg;
    if (osz_upfirdn_plan_taps(ufd, &K, &M, &half, &gf, &gr, &tg, &ldq) != OSZ_OK) return 0;
    SosDecGeom gd;
    int wt;
    size_t smem;
    if (!sosdec_geometry(K, M, sosdec_ksplit(M), &gd, &wt, &smem)) return 0;
    if (n < 4 * (int64_t)K) return 0;
    static const int forced = [] {
        const char *e = getenv("OSZ_SOSDEC_SPANS");
        return e ? atoi(e) : 0;
    }();
    const double sms = (double)sm_count();
    int64_t kmax = settle > 0 ? n / (2 * settle) : 1;
    if (kmax > n / (4 * (int64_t)K)) kmax = n / (4 * (int64_t)K);
    if (kmax > 64) kmax = 64;
    if (kmax < 1) kmax = 1;
    if (forced > 0) return (int)(forced < kmax ? forced : kmax);
    // cost of a span: FIR work ~ its length, scan work ~ (length + warm-up) * 0.15
    int best = 1;
    double best_t = 1e300;
    for (int64_t k = 1; k <= kmax; ++k) {
        const double waves = ceil((double)rows * k / sms);
        const double per = (double)n / k + 0.15 * (k > 1 ? (double)settle : 0.0);
        const double t = waves * per;
        if (t < best_t * 0.999) {
            best_t = t;
            best = (int)k;
        }
    }
    return best;
}

// x: (rows, n) input of the pass (for the backward pass of a forward-backward filter:
// the forward output F of the chunk); state: (rows, nsec, 2) state entering the pass
// (updated to the state leaving it); A: global index of the chunk's first sample;
// y: decimated outputs, column c = global output out_first + c; only outputs whose
// whole K-tap window lies inside one span are written here, the rest by
// osz_sosdec_boundary_f64 from `edges` ((rows, nspan, 2, K-1) doubles).
int osz_sosdec_exec_f64(const osz_sos_plan *sos, const osz_upfirdn_plan *ufd, const double *x,
                        int64_t ldx, int64_t rows, int64_t n, int reverse, double *state,
                        int nspan, int64_t A, double *y, int64_t ldy, int64_t out_first,
                        int64_t n_out, double *edges, void *stream) {
    if (!sos || !ufd || !x || !state || !y || !edges)
        return fail(OSZ_ERR_ARG, "osz_sosdec_exec_f64: null argument");
    if (rows <= 0 || n <= 0) return OSZ_OK;
    if (rows > 2147483647LL || nspan < 1 || nspan > 65535)
        return fail(OSZ_ERR_ARG, "osz_sosdec_exec_f64: bad rows / span count");
    cudaStream_t st = as_stream(stream);
    SosParams prm;
    const double *lanepow = nullptr;
    int64_t settle = -1;
    int rc = osz_sos_plan_params(sos, &prm, &lanepow, &settle);
    if (rc != OSZ_OK) return rc;
    int K, M, half, ldq;
    const double *gf, *gr, *tg;
    rc = osz_upfirdn_plan_taps(ufd, &K, &M, &half, &gf, &gr, &tg, &ldq);
    if (rc != OSZ_OK) return rc;
    SosDecGeom gd;
    int wt;
    size_t smem;
    const int ks = sosdec_ksplit(M);
    if (!sosdec_geometry(K, M, ks, &gd, &wt, &smem) || gd.m.ldq != ldq)
        return fail(OSZ_ERR_UNSUPPORTED, "osz_sosdec_exec_f64: no fused tile geometry");
    if (nspan > 1 && settle <= 0)
        return fail(OSZ_ERR_UNSUPPORTED, "osz_sosdec_exec_f64: time split needs a decaying filter");
    const int64_t span_len = (n + nspan - 1) / nspan;
    if (span_len < 2 * (int64_t)K || (n - (nspan - 1) * span_len) < 2 * (int64_t)K)
        return fail(OSZ_ERR_ARG, "osz_sosdec_exec_f64: spans shorter than the filter");
    const int64_t abase = reverse ? n - 1 + A - half : (int64_t)half - (K - 1) - A;
    // the last span writes the carried state while span 0 may still read it
    double *scratch = nullptr;
    const double *state_in = state;
    const int64_t n_copy = nspan > 1 ? rows * 2 * prm.nsec : 0;
    if (n_copy) {
        OSZ_CUDA(scratch_alloc((void **)&scratch, (size_t)n_copy * 8, st));
        OSZ_CUDA(cudaMemcpyAsync(scratch, state, (size_t)n_copy * 8, cudaMemcpyDeviceToDevice, st));
        state_in = scratch;
    }
    const dim3 grid((unsigned)rows, (unsigned)nspan);
    const double *gpad = reverse ? gr : gf;
#define OSZ_SD_CASE(WT, KS)                                                                      \
    if (wt == WT && ks == KS) {                                                                  \
        OSZ_CUDA(cudaFuncSetAttribute(sosdec_kernel<WT, KS>,                                     \
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));  \
        sosdec_kernel<WT, KS><<<grid, SOS_NT + WT * KS * 32, smem, st>>>(                        \
            prm, gd, x, ldx, n, reverse, state_in, state, lanepow, span_len, settle, gpad, abase, \
            y, ldy, out_first, n_out, edges);                                                    \
    }
    OSZ_SD_CASE(1, 8) OSZ_SD_CASE(1, 4)
#undef OSZ_SD_CASE
    cudaError_t err = cudaGetLastError();
    if (scratch) cudaFreeAsync(scratch, st);
    if (err != cudaSuccess)
        return fail(OSZ_ERR_CUDA, std::string("sosdec_kernel launch: ") + cudaGetErrorString(err));
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return OSZ_OK;
}

// The outputs osz_sosdec_exec_f64 left out: those whose window straddles the chunk's
// start (prev_tail: (rows, K-1) last samples of the previous chunk's pass output, row
// pitch prev_ld; null = recording start, zeros), a boundary between two spans, or --
// with has_end -- the chunk's end (recording end, zeros after it).  Only outputs
// j_min <= j <= j_max are written.
int osz_sosdec_boundary_f64(const osz_upfirdn_plan *ufd, const double *edges, int nspan,
                            int reverse, const double *prev_tail, int64_t prev_ld, int has_end,
                            int64_t rows, int64_t n, int64_t A, double *y, int64_t ldy,
                            int64_t out_first, int64_t n_out, int64_t j_min, int64_t j_max,
                            void *stream) {
    if (!ufd || !edges || !y) return fail(OSZ_ERR_ARG, "osz_sosdec_boundary_f64: null argument");
    if (rows <= 0 || n <= 0 || nspan < 1) return OSZ_OK;
    int K, M, half, ldq;
    const double *gf, *gr, *tg;
    const int rc = osz_upfirdn_plan_taps(ufd, &K, &M, &half, &gf, &gr, &tg, &ldq);
    if (rc != OSZ_OK) return rc;
    const int64_t span_len = (n + nspan - 1) / nspan;
    const dim3 grid((unsigned)(nspan + 1), (unsigned)rows);
    sosdec_boundary_kernel<<<grid, 256, 0, as_stream(stream)>>>(
        tg, K, M, half, edges, nspan, reverse, prev_tail, prev_ld, has_end, A, n, span_len, y, ldy,
        out_first, n_out, j_min, j_max);
    OSZ_LAUNCHED("sosdec_boundary_kernel");
    return OSZ_OK;
}

}  // extern "C"
