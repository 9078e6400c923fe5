// Pieces of the tensor-core (DMMA) banded-Toeplitz decimating filter shared by the
// stand-alone kernel (upfirdn.cu) and the fused backward-pass + decimator kernel
// (sosdec.cu).  See upfirdn.cu for the derivation.
#pragma once

#include <vector>

#include "common.cuh"

namespace osz {

__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

struct UfdMmaGeom {
    int K, M, half;
    int S;            // outputs per segment (32 * WT, a power of two)
    int logS;
    int SM;           // S * M: input samples per segment
    int P;            // segment pitch in shared memory (doubles)
    int total_len;    // input samples a tile touches (8 segments + reach)
    int ldq;          // doubles per phase row of the padded tap table
    int p_rem;        // phases 0 .. p_rem hold one tap more than the others
    int ks_hi, ks_lo; // k-steps of the longer / shorter phases
    int ktotal;       // k-steps of all phases
};


// One batch of NB consecutive k-steps of one phase for the FOUR 8 x 8 output tiles
// tau = 4 wt .. 4 wt + 3 of a warp.  Tile j at step s needs the A fragment of decimated
// index n = n0 + 4 s + 8 j (n0 = 32 wt + q + 4 s0), so NB steps need NB + 6 fragments:
// they and the NB tap fragments are loaded up front, unconditionally (every guard in
// this loop made the compiler sink the loads next to their MMA and pay the LDS latency
// per step), then the 4 NB MMAs run as four independent accumulator chains.
template <int NB, class Fetch>
__device__ __forceinline__ void ufd_mma_batch(Fetch fetch, const double *gp4, int n0,
                                              double (&c)[8]) {
    double f[NB + 6], tb[NB];
#pragma unroll
    for (int i = 0; i < NB + 6; ++i) f[i] = fetch(n0 + 4 * i);
#pragma unroll
    for (int i = 0; i < NB; ++i) tb[i] = gp4[4 * i];
#pragma unroll
    for (int i = 0; i < NB; ++i) {
        dmma884(c[0], c[1], f[i], tb[i]);
        dmma884(c[2], c[3], f[i + 2], tb[i]);
        dmma884(c[4], c[5], f[i + 4], tb[i]);
        dmma884(c[6], c[7], f[i + 6], tb[i]);
    }
}

// One warp's share of a tile: k-steps [k_lo, k_hi) (numbered phase by phase) of the
// banded Toeplitz product.  `fetch(p, n)` returns sample n of the decimated sequence of
// phase p in this lane's A row (segment g): window offset p + n * M.  The tensor pipe
// issues one DMMA.8x8x4 per 16 clk and SM sub-partition (profiles/r02_ncu_summary.md),
// so the loop has to stay well below 16 instructions per MMA: this form runs at ~4.
constexpr int UFD_NSB = 14;    // k-steps per batch: 20 A + 14 B fragments up front, 56 MMAs behind them
template <class Fetch>
__device__ __forceinline__ void ufd_mma_ksteps(const UfdMmaGeom &gm, const double *gs, int k_lo,
                                               int k_hi, int wt, int g, int q, Fetch fetch,
                                               double (&c)[8]) {
#pragma unroll
    for (int i = 0; i < 8; ++i) c[i] = 0.0;
    // (phase, step) of k_lo: phases 0 .. p_rem hold ks_hi steps, the others ks_lo
    const int nhi = (gm.p_rem + 1) * gm.ks_hi;
    int p, s;
    if (k_lo < nhi || gm.ks_lo == 0) {
        p = k_lo / gm.ks_hi;
        s = k_lo - p * gm.ks_hi;
    } else {
        const int rest = k_lo - nhi;
        p = gm.p_rem + 1 + rest / gm.ks_lo;
        s = rest - (rest / gm.ks_lo) * gm.ks_lo;
    }
    int kk = k_lo;
    while (kk < k_hi) {
        const int ns = p <= gm.p_rem ? gm.ks_hi : gm.ks_lo;
        int s_end = s + (k_hi - kk);
        if (s_end > ns) s_end = ns;
        const double *gp = gs + (size_t)p * gm.ldq + 7 + q - g;
        const int nbase = 32 * wt + q;
        auto fp = [&](int n) { return fetch(p, n); };
        int s0 = s;
        for (; s0 + UFD_NSB <= s_end; s0 += UFD_NSB)
            ufd_mma_batch<UFD_NSB>(fp, gp + 4 * s0, nbase + 4 * s0, c);
        switch (s_end - s0) {        // the remainder, as one unguarded batch of its own size
#define OSZ_UFD_REM(NBR) \
    case NBR: ufd_mma_batch<NBR>(fp, gp + 4 * s0, nbase + 4 * s0, c); break;
            OSZ_UFD_REM(13) OSZ_UFD_REM(12) OSZ_UFD_REM(11) OSZ_UFD_REM(10) OSZ_UFD_REM(9)
            OSZ_UFD_REM(8) OSZ_UFD_REM(7) OSZ_UFD_REM(6) OSZ_UFD_REM(5) OSZ_UFD_REM(4)
            OSZ_UFD_REM(3) OSZ_UFD_REM(2) OSZ_UFD_REM(1)
#undef OSZ_UFD_REM
            default: break;
        }
        kk += s_end - s;
        ++p;
        s = 0;
    }
}

// k-steps (groups of four taps) of every phase of a K-tap filter decimated by M:
// phase p has Q_p taps and needs ceil((Q_p + 7) / 4) steps (8 outputs share a window).
inline std::vector<int> ufd_ksteps(int K, int M, int *smax_out, long *total_out) {
    std::vector<int> ks(M);
    int smax = 0;
    long total = 0;
    for (int p = 0; p < M; ++p) {
        const int qp = p <= (K - 1) % M ? (K - 1) / M + 1 : (K - 1) / M;
        ks[p] = qp > 0 ? (qp + 7 + 3) / 4 : 0;
        if (ks[p] > smax) smax = ks[p];
        total += ks[p];
    }
    if (smax_out) *smax_out = smax;
    if (total_out) *total_out = total;
    return ks;
}

// segment pad (doubles) that spreads the 16 lanes of a half warp (g * pitch + q * M)
// over the most shared-memory banks.  Only EVEN pitches are considered: the segments are
// staged by 1-D TMA bulk copies, which need 16-byte aligned destinations.  For even M that
// leaves a 2-way conflict on the A-fragment loads (8 of 16 banks; an odd pitch would reach 16
// but has to be staged by 8-byte cp.async: 2.98 ms against 0.9 for 449 taps / M = 20).
inline int ufd_best_pad(int SM, int M) {
    int best_pad = 0, best_score = -1;
    for (int pad = (SM & 1); pad < 16; pad += 2) {
        bool seen[16] = {false};
        int score = 0;
        for (int g = 0; g < 4; ++g)
            for (int q = 0; q < 4; ++q) {
                const int b = (int)(((long)g * (SM + pad) + (long)q * M) % 16);
                if (!seen[b]) {
                    seen[b] = true;
                    ++score;
                }
            }
        if (score > best_score) {
            best_score = score;
            best_pad = pad;
        }
    }
    return best_pad;
}

// padded tap table: gpad[p][7 + v] = taps[p + v * M]  (taps in the order the kernel
// walks its window: taps[w] multiplies window sample w of output 0)
inline std::vector<double> ufd_gpad(const double *taps, int K, int M, int ldq) {
    std::vector<double> gp((size_t)M * ldq, 0.0);
    for (int j = 0; j < K; ++j) gp[(size_t)(j % M) * ldq + 7 + j / M] = taps[j];
    return gp;
}

}  // namespace osz
