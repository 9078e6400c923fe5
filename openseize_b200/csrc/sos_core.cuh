// Device core of the time-parallel biquad scan, shared by the stand-alone kernels
// (sos.cu) and the fused backward-pass + decimator kernel (sosdec.cu).
#pragma once

#include "common.cuh"

namespace osz {

constexpr int SOS_NT = 256;
constexpr int SOS_T = 32;
constexpr int SOS_MAXSEC = 16;

struct SosSec {
    double b0, b1, b2, a1, a2;
    double g0[SOS_T], g1[SOS_T];
    double P[5][4];     // M^(2^k), row major
    double Q[4];        // M^32 (one warp)
    double A8[4];       // A^8: state transition over one 8-sample sub-piece
    double A16[4];      // A^16: thread transition of the T = 16 kernel
};
struct SosParams {
    int nsec;
    int pad_;
    SosSec sec[SOS_MAXSEC];
};
struct SosZi {
    double zi[SOS_MAXSEC][2];
};

__device__ __forceinline__ void mat_apply(const double (&m)[4], double a0, double a1, double &o0,
                                          double &o1) {
    o0 = fma(m[0], a0, m[1] * a1);
    o1 = fma(m[2], a0, m[3] * a1);
}


// bar.sync on a NAMED barrier: the scan's cross-warp step only involves the
// SOS_NT scan threads, which are the whole CTA in sos.cu (barrier 0 ==
// __syncthreads) and one warp group of a larger CTA in sosdec.cu.
template <int BAR>
__device__ __forceinline__ void named_bar_sync(int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"n"(BAR), "r"(nthreads) : "memory");
}

__device__ __forceinline__ void named_bar_sync_id(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// One block of SOS_NT * T samples through every section of the cascade.
//   v[T]     the thread's T consecutive samples (in: section 0 input, out: cascade output)
//   full     every slot holds a real sample and the block starts on the carried
//            state (thread 0); otherwise the block is right-aligned behind `off`
//            virtual zero samples and the carry is injected at slot `off`
//   carry    [section][2] state entering the block, replaced by the state leaving it
//   wtot     [2][warps][2] scratch, double-buffered by section parity
template <int T, int BAR>
__device__ __forceinline__ void sos_scan_block(const SosParams &prm, double (&v)[T], bool full,
                                               int off, double (*carry)[2],
                                               double (*wtot)[SOS_NT / 32][2],
                                               const double *__restrict__ lanepow, int tid,
                                               int lane, int warp) {
    constexpr int LOGT = T == 32 ? 5 : 4;
    constexpr int NCH = T / 8;             // independent 8-sample chains per thread
    const int nsec = prm.nsec;
    const int pstar = off >> LOGT, ioff = off & (T - 1);
    for (int s = 0; s < nsec; ++s) {
        const SosSec &c = prm.sec[s];
        const double b0 = c.b0, b1 = c.b1, b2 = c.b2, na1 = -c.a1, na2 = -c.a2;
        double z0 = 0.0, z1 = 0.0;
        double zs0[NCH - 1], zs1[NCH - 1];     // zero-state finals of sub-pieces 0 .. NCH-2
        if (full) {
            // NCH independent 8-sample chains per thread (a single chain left
            // the FP64 pipe 2/3 idle waiting on its own results): sub-piece 0
            // starts from the thread's entering state (the carry for thread
            // 0, else zero), the others from zero and are fixed up in-thread
            // with the same zero-input tables.
            double za0[NCH], za1[NCH];
#pragma unroll
            for (int j = 0; j < NCH; ++j) za0[j] = za1[j] = 0.0;
            if (tid == 0) {
                za0[0] = carry[s][0];
                za1[0] = carry[s][1];
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) {
#pragma unroll
                for (int j = 0; j < NCH; ++j) {
                    const double xi = v[8 * j + i];
                    const double yi = fma(b0, xi, za0[j]);
                    za0[j] = fma(na1, yi, fma(b1, xi, za1[j]));
                    za1[j] = fma(na2, yi, b2 * xi);
                    v[8 * j + i] = yi;
                }
            }
            // thread-final state for a zero entering state: the sub-pieces' final
            // states chained through A^8.  Their OUTPUTS are not fixed up here: every
            // sub-piece gets one zero-input correction below, from its true entering
            // state, once the scan has delivered the thread's.
            double e0 = za0[0], e1 = za1[0];
#pragma unroll
            for (int j = 1; j < NCH; ++j) {
                const double t0 = fma(c.A8[0], e0, c.A8[1] * e1) + za0[j];
                const double t1 = fma(c.A8[2], e0, c.A8[3] * e1) + za1[j];
                e0 = t0;
                e1 = t1;
            }
            z0 = e0;
            z1 = e1;
#pragma unroll
            for (int j = 0; j < NCH - 1; ++j) {
                zs0[j] = za0[j];
                zs1[j] = za1[j];
            }
        } else {
            const bool inj = tid == pstar;
            const double c0 = carry[s][0], c1 = carry[s][1];
#pragma unroll
            for (int i = 0; i < T; ++i) {
                if (inj && i == ioff) {
                    z0 = c0;
                    z1 = c1;
                }
                const double xi = v[i];
                const double yi = fma(b0, xi, z0);
                z0 = fma(na1, yi, fma(b1, xi, z1));
                z1 = fma(na2, yi, b2 * xi);
                v[i] = yi;
            }
        }
        // ---- warp-inclusive scan of e_p = M e_{p-1} + f_p, M = A^T:
        //      M^(2^k) is P[k] for T = 32 and {A16, P[0..3]} for T = 16
        double f0 = z0, f1 = z1;
#pragma unroll
        for (int k = 0; k < 5; ++k) {
            const double *pm = T == 32 ? c.P[k] : (k == 0 ? c.A16 : c.P[k - 1]);
            const double g0 = __shfl_up_sync(0xffffffffu, f0, 1 << k);
            const double g1 = __shfl_up_sync(0xffffffffu, f1, 1 << k);
            if (lane >= (1 << k)) {
                f0 += fma(pm[0], g0, pm[1] * g1);
                f1 += fma(pm[2], g0, pm[3] * g1);
            }
        }
        double (*wt)[2] = wtot[s & 1];
        if (lane == 31) {
            wt[warp][0] = f0;
            wt[warp][1] = f1;
        }
        named_bar_sync<BAR>(SOS_NT);
        // ---- state entering this warp (transition over one warp: M^32)
        const double *qm = T == 32 ? c.Q : c.P[4];
        double cw0 = 0.0, cw1 = 0.0;
        for (int u = 0; u < warp; ++u) {
            const double t0 = fma(qm[0], cw0, qm[1] * cw1) + wt[u][0];
            const double t1 = fma(qm[2], cw0, qm[3] * cw1) + wt[u][1];
            cw0 = t0;
            cw1 = t1;
        }
        // ---- true state at the end of this thread's piece
        const double *lp = lanepow + ((size_t)s * 32 + lane) * 4;
        const double e0 = f0 + fma(ldg(lp + 0), cw0, ldg(lp + 1) * cw1);
        const double e1 = f1 + fma(ldg(lp + 2), cw0, ldg(lp + 3) * cw1);
        // ---- state entering this thread's piece
        double s0 = __shfl_up_sync(0xffffffffu, e0, 1);
        double s1 = __shfl_up_sync(0xffffffffu, e1, 1);
        if (lane == 0) {
            s0 = cw0;
            s1 = cw1;
        }
        // zero-input response of the entering state (it is exactly zero for
        // every thread up to and including the one that injected the carry)
        if (full) {
            // sub-piece j enters with E_j: E_0 = the thread's entering state,
            // E_(j+1) = A^8 E_j + (zero-state final of sub-piece j)
            double q0 = s0, q1 = s1;
#pragma unroll
            for (int j = 0; j < NCH; ++j) {
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    v[8 * j + i] = fma(c.g0[i], q0, fma(c.g1[i], q1, v[8 * j + i]));
                if (j + 1 < NCH) {
                    const double t0 = fma(c.A8[0], q0, c.A8[1] * q1) + zs0[j];
                    const double t1 = fma(c.A8[2], q0, c.A8[3] * q1) + zs1[j];
                    q0 = t0;
                    q1 = t1;
                }
            }
        } else {
#pragma unroll
            for (int i = 0; i < T; ++i) v[i] = fma(c.g0[i], s0, fma(c.g1[i], s1, v[i]));
        }
        if (tid == SOS_NT - 1) {
            carry[s][0] = e0;
            carry[s][1] = e1;
        }
        // No barrier here: the other wtot buffer takes section s+1's totals, and a
        // warp can only write this one again (section s+2) after every warp has
        // passed section s+1's barrier, i.e. has finished reading it.  carry[s] is
        // next read by thread 0 in the NEXT block, behind that block's barriers.
    }
}

}  // namespace osz
