// Transfer-function (b, a) IIR filters of any order.  Replaces
// scipy.signal.lfilter as called per chunk by nm.lfilter / nm.filtfilt
// (reference core/numerical.py:445,508,511,519) when max(len a, len b) > 3;
// second order and below ride the time-parallel biquad scan of sos.cu.
//
// scipy's lfilter is the transposed direct form II:
//     y      = b0 x + z[0]
//     z[i]   = b[i+1] x - a[i+1] y + z[i+1]        (i < S-1, S = K-1 states)
//     z[S-1] = b[S] x - a[S] y
// evaluated here in exactly that order, one thread per row walking its row in
// time (the critical path is two dependent FMAs per sample, so a row costs
// ~20 cycles per sample whatever the order; rows run concurrently).  This is
// the coverage path for the reference's `fmt='ba'` designs above second order
// -- openseize's own designs use SOS except Notch (order 2) -- not a
// bandwidth-bound kernel.
#include <vector>

#include "common.cuh"

namespace osz {

constexpr int TF_MAXS = 32;        // states (filter order) supported
constexpr int TF_BATCH = 16;       // samples loaded / stored per thread at a time

struct TfParams {
    int S;                         // number of states = max(len a, len b) - 1
    int pad_;
    double b[TF_MAXS + 1];         // normalised by a[0], zero padded
    double a[TF_MAXS + 1];
};

template <int SC /* compile-time S, 0: dynamic */, bool WRITE>
__global__ void __launch_bounds__(32)
tf_seq_kernel(const __grid_constant__ TfParams prm, const double *__restrict__ x, int64_t ldx,
              int64_t rows, int64_t n, int reverse, double *__restrict__ state,
              double *__restrict__ y, int64_t ldy) {
    const int64_t row = (int64_t)blockIdx.x * 32 + threadIdx.x;
    if (row >= rows) return;
    const int S = SC ? SC : prm.S;
    double z[SC ? SC : TF_MAXS];
    double *st = state + row * S;
    for (int i = 0; i < S; ++i) z[i] = st[i];
    const double *xr = x + row * ldx;
    double *yr = WRITE ? y + row * ldy : nullptr;
    const double b0 = prm.b[0];
    for (int64_t t0 = 0; t0 < n; t0 += TF_BATCH) {
        double xv[TF_BATCH];
        const int m = n - t0 < TF_BATCH ? (int)(n - t0) : TF_BATCH;
#pragma unroll
        for (int j = 0; j < TF_BATCH; ++j)
            if (j < m) xv[j] = ld_stream(reverse ? xr + (n - 1 - t0 - j) : xr + t0 + j);
#pragma unroll
        for (int j = 0; j < TF_BATCH; ++j) {
            if (j < m) {
                const double xi = xv[j];
                const double yi = fma(b0, xi, z[0]);
                if (SC) {
#pragma unroll
                    for (int i = 0; i < (SC ? SC : 1) - 1; ++i)
                        z[i] = fma(-prm.a[i + 1], yi, fma(prm.b[i + 1], xi, z[i + 1]));
                } else {
                    for (int i = 0; i < S - 1; ++i)
                        z[i] = fma(-prm.a[i + 1], yi, fma(prm.b[i + 1], xi, z[i + 1]));
                }
                z[S - 1] = fma(-prm.a[S], yi, prm.b[S] * xi);
                xv[j] = yi;
            }
        }
        if (WRITE) {
#pragma unroll
            for (int j = 0; j < TF_BATCH; ++j)
                if (j < m) st_stream(reverse ? yr + (n - 1 - t0 - j) : yr + t0 + j, xv[j]);
        }
    }
    for (int i = 0; i < S; ++i) st[i] = z[i];
}

__global__ void tf_state_from_sample_kernel(TfParams zi /* zi in .b[0..S) */, const double *x,
                                            int64_t ldx, int64_t rows, int64_t sample,
                                            double *state) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * zi.S) return;
    const int64_t row = i / zi.S;
    state[i] = zi.b[i % zi.S] * x[row * ldx + sample];
}

}  // namespace osz

using namespace osz;

struct osz_tf_plan {
    TfParams prm;
};

template <int SC>
static int launch_tf(const osz_tf_plan *p, const double *x, int64_t ldx, int64_t rows, int64_t n,
                     int reverse, double *state, double *y, int64_t ldy, cudaStream_t st) {
    const unsigned grid = (unsigned)((rows + 31) / 32);
    if (y)
        tf_seq_kernel<SC, true><<<grid, 32, 0, st>>>(p->prm, x, ldx, rows, n, reverse, state, y, ldy);
    else
        tf_seq_kernel<SC, false><<<grid, 32, 0, st>>>(p->prm, x, ldx, rows, n, reverse, state, y,
                                                      ldy);
    OSZ_LAUNCHED("tf_seq_kernel");
    return OSZ_OK;
}

extern "C" {

int osz_tf_plan_create(osz_tf_plan **out, const double *b, int nb, const double *a, int na) {
    if (!out || !b || !a || nb < 1 || na < 1 || a[0] == 0.0)
        return fail(OSZ_ERR_ARG, "osz_tf_plan_create: bad arguments");
    const int K = nb > na ? nb : na;
    if (K - 1 > TF_MAXS)
        return fail(OSZ_ERR_UNSUPPORTED, "osz_tf_plan_create: filter order above 32");
    if (K < 2) return fail(OSZ_ERR_ARG, "osz_tf_plan_create: a filter needs at least one state");
    osz_tf_plan *p = new osz_tf_plan();
    p->prm.S = K - 1;
    p->prm.pad_ = 0;
    for (int i = 0; i <= TF_MAXS; ++i) {
        // scipy normalises by a[0] (lfilter: "a[0] is not 1 -> both a and b are normalised")
        p->prm.b[i] = i < nb ? b[i] / a[0] : 0.0;
        p->prm.a[i] = i < na ? a[i] / a[0] : 0.0;
    }
    *out = p;
    return OSZ_OK;
}

int osz_tf_plan_destroy(osz_tf_plan *p) {
    delete p;
    return OSZ_OK;
}

int osz_tf_plan_states(const osz_tf_plan *p) { return p ? p->prm.S : 0; }

int osz_tf_exec_f64(const osz_tf_plan *p, const double *x, int64_t ldx, int64_t rows, int64_t n,
                    int reverse, double *state, double *y, int64_t ldy, void *stream) {
    if (!p || !x || !state) return fail(OSZ_ERR_ARG, "osz_tf_exec_f64: null argument");
    if (rows <= 0 || n <= 0) return OSZ_OK;
    cudaStream_t st = as_stream(stream);
    switch (p->prm.S) {
#define OSZ_TF_CASE(S) \
    case S: return launch_tf<S>(p, x, ldx, rows, n, reverse, state, y, ldy, st);
        OSZ_TF_CASE(1) OSZ_TF_CASE(2) OSZ_TF_CASE(3) OSZ_TF_CASE(4) OSZ_TF_CASE(5) OSZ_TF_CASE(6)
        OSZ_TF_CASE(7) OSZ_TF_CASE(8) OSZ_TF_CASE(9) OSZ_TF_CASE(10) OSZ_TF_CASE(11)
        OSZ_TF_CASE(12) OSZ_TF_CASE(13) OSZ_TF_CASE(14) OSZ_TF_CASE(15) OSZ_TF_CASE(16)
#undef OSZ_TF_CASE
        default: return launch_tf<0>(p, x, ldx, rows, n, reverse, state, y, ldy, st);
    }
}

int osz_tf_state_from_sample_f64(const osz_tf_plan *p, const double *zi, const double *x,
                                 int64_t ldx, int64_t rows, int64_t sample, double *state,
                                 void *stream) {
    if (!p || !zi || !x || !state)
        return fail(OSZ_ERR_ARG, "osz_tf_state_from_sample_f64: null argument");
    if (rows <= 0) return OSZ_OK;
    TfParams z = p->prm;
    for (int i = 0; i < z.S; ++i) z.b[i] = zi[i];
    const int64_t count = rows * z.S;
    tf_state_from_sample_kernel<<<(unsigned)((count + 255) / 256), 256, 0, as_stream(stream)>>>(
        z, x, ldx, rows, sample, state);
    OSZ_LAUNCHED("tf_state_from_sample_kernel");
    return OSZ_OK;
}

}  // extern "C"
