// Pieces of the tensor-core (DMMA) banded-Toeplitz decimating filter shared by the
// stand-alone kernel (upfirdn.cu) and the fused backward-pass + decimator kernel
// (sosdec.cu).  See upfirdn.cu for the derivation.
#pragma once

#include <vector>

#include "common.cuh"

namespace osz {

__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

struct UfdMmaGeom {
    int K, M, half;
    int S;            // outputs per segment (32 * WT)
    int SM;           // S * M: input samples per segment
    int P;            // segment pitch in shared memory (doubles)
    int total_len;    // input samples a tile touches (8 segments + reach)
    int ldq;          // doubles per phase row of the padded tap table
    int p_rem;        // phases 0 .. p_rem hold one tap more than the others
    int ks_hi, ks_lo; // k-steps of the longer / shorter phases
    int ktotal;       // k-steps of all phases
};


// One warp's share of a tile: k-steps [k_lo, k_hi) (numbered phase by phase) of the
// banded Toeplitz product for the FOUR 8 x 8 output tiles tau = 4 wt .. 4 wt + 3.
// `fetch(r)` returns the sample at in-segment offset r of segment g (this lane's A row).
// Tile j at step s needs fragment F(8 wt + 2 j + s): one new A fragment and one tap
// fragment per k-step feed four MMAs.  The fragments of a batch of UFD_NSB steps are
// loaded up front (LDS latency paid once per batch); the four tiles are four
// independent accumulator chains.  The tensor pipe issues one DMMA.8x8x4 per 16 clk
// and SM sub-partition (profiles/r02_ncu_summary.md), so the loop must stay well below
// 16 instructions per MMA: this form runs at ~4.
constexpr int UFD_NSB = 8;
template <class Fetch>
__device__ __forceinline__ void ufd_mma_ksteps(const UfdMmaGeom &gm, const double *gs, int k_lo,
                                               int k_hi, int wt, int g, int q, Fetch fetch,
                                               double (&c)[8]) {
    const int M = gm.M;
    const int step = 4 * M;
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
        const int r0 = p + (32 * wt + q) * M;
        for (int s0 = s; s0 < s_end; s0 += UFD_NSB) {
            const int nb = s_end - s0 < UFD_NSB ? s_end - s0 : UFD_NSB;   // steps in this batch
            double f[UFD_NSB + 6], tb[UFD_NSB];
#pragma unroll
            for (int i = 0; i < UFD_NSB + 6; ++i)
                f[i] = i < nb + 6 ? fetch(r0 + (s0 + i) * step) : 0.0;
#pragma unroll
            for (int i = 0; i < UFD_NSB; ++i) tb[i] = i < nb ? gp[4 * (s0 + i)] : 0.0;
#pragma unroll
            for (int i = 0; i < UFD_NSB; ++i) {
                if (i < nb) {
                    dmma884(c[0], c[1], f[i], tb[i]);
                    dmma884(c[2], c[3], f[i + 2], tb[i]);
                    dmma884(c[4], c[5], f[i + 4], tb[i]);
                    dmma884(c[6], c[7], f[i + 6], tb[i]);
                }
            }
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
// over the most shared-memory banks
inline int ufd_best_pad(int SM, int M) {
    int best_pad = 0, best_score = -1;
    for (int pad = 0; pad < 16; ++pad) {
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
