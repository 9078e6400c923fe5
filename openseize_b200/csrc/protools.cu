// Producer tools on device-resident chunks: the mask
// compaction of MaskedProducer (reference core/producer.py:427-444, np.take of
// the kept samples along the sample axis) and the reductions of
// core/protools.py -- mean (:500-543), std (:546-595), standardize (:598-668).
// All of them stream a chunk once: bandwidth-bound, no reuse, so the kernels
// are plain coalesced grid-stride loops; grids are sized in multiples of the SM
// count.
#include <math.h>

#include "common.cuh"

namespace osz {

constexpr int MOM_SLOTS = 64;   // partial sums per row (deterministic two-stage reduction)

// y[r][j] = x[r][idx[j]]  -- idx ascending (flatnonzero of the mask chunk), so
// neighbouring threads read neighbouring or nearby addresses.
__global__ void __launch_bounds__(256)
take_cols_kernel(const double *__restrict__ x, int64_t ldx, const int64_t *__restrict__ idx,
                 int64_t nkeep, double *__restrict__ y, int64_t ldy) {
    const int64_t row = blockIdx.y;
    const double *xr = x + row * ldx;
    double *yr = y + row * ldy;
    const int64_t step = (int64_t)gridDim.x * blockDim.x;
    int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; j + 3 * step < nkeep; j += 4 * step) {       // four independent loads in flight
        double v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = ld_stream(xr + __ldg(idx + j + u * step));
#pragma unroll
        for (int u = 0; u < 4; ++u) st_stream(yr + j + u * step, v[u]);
    }
    for (; j < nkeep; j += step) yr[j] = ld_stream(xr + __ldg(idx + j));
}

// Per row and slot: (sum, sum of squares, count) of the slot's share of the
// chunk; NaNs are skipped when ignore_nan (np.nanmean), else they propagate.
__global__ void __launch_bounds__(256)
row_moments_partial_kernel(const double *__restrict__ x, int64_t ldx, int64_t n, int ignore_nan,
                           double *__restrict__ part /* [row][slot][3] */) {
    __shared__ double red[3][8];
    const int64_t row = blockIdx.y;
    const double *xr = x + row * ldx;
    double s1 = 0.0, s2 = 0.0, c = 0.0;
    const int64_t step = (int64_t)gridDim.x * blockDim.x;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + 7 * step < n; i += 8 * step) {           // eight independent loads in flight
        double v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = ld_stream(xr + i + u * step);
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            if (!ignore_nan || v[u] == v[u]) {
                s1 += v[u];
                s2 = fma(v[u], v[u], s2);
                c += 1.0;
            }
        }
    }
    for (; i < n; i += step) {
        const double v = ld_stream(xr + i);
        if (!ignore_nan || v == v) {
            s1 += v;
            s2 = fma(v, v, s2);
            c += 1.0;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s1 += __shfl_xor_sync(0xffffffffu, s1, o);
        s2 += __shfl_xor_sync(0xffffffffu, s2, o);
        c += __shfl_xor_sync(0xffffffffu, c, o);
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) {
        red[0][warp] = s1;
        red[1][warp] = s2;
        red[2][warp] = c;
    }
    __syncthreads();
    if (threadIdx.x < 3) {
        double t = 0.0;
#pragma unroll
        for (int w = 0; w < 8; ++w) t += red[threadIdx.x][w];
        part[(row * gridDim.x + blockIdx.x) * 3 + threadIdx.x] = t;
    }
}

// Fold a chunk's partial sums into the running accumulators exactly as the
// reference combines chunks (protools.py:531-536, 583-590):
//   acc[r][0] += n * mean_chunk ; acc[r][1] += n * mean(chunk^2) ; acc[r][2] += n
// with mean_chunk = sum / count of the values that were not skipped.
__global__ void row_moments_fold_kernel(const double *__restrict__ part, int slots, int64_t rows,
                                        int64_t n, double *__restrict__ acc) {
    const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= rows) return;
    double s1 = 0.0, s2 = 0.0, c = 0.0;
    for (int k = 0; k < slots; ++k) {
        const double *p = part + (row * slots + k) * 3;
        s1 += p[0];
        s2 += p[1];
        c += p[2];
    }
    const double dn = (double)n;
    acc[row * 3 + 0] += dn * (s1 / c);
    acc[row * 3 + 1] += dn * (s2 / c);
    acc[row * 3 + 2] += dn;
}

// y[r][i] = (x[r][i] - mu[r]) / sd[r]   (protools.py:659-662: subtract, then divide)
__global__ void __launch_bounds__(256)
row_standardize_kernel(const double *__restrict__ x, int64_t ldx, int64_t n,
                       const double *__restrict__ mu, const double *__restrict__ sd,
                       double *__restrict__ y, int64_t ldy) {
    const int64_t row = blockIdx.y;
    const double *xr = x + row * ldx;
    double *yr = y + row * ldy;
    const double m = mu[row], s = sd[row];
    const int64_t step = (int64_t)gridDim.x * blockDim.x;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + 3 * step < n; i += 4 * step) {           // four independent loads in flight
        double v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = ld_stream(xr + i + u * step);
#pragma unroll
        for (int u = 0; u < 4; ++u)
            st_stream(yr + i + u * step, __ddiv_rn(__dsub_rn(v[u], m), s));
    }
    for (; i < n; i += step) st_stream(yr + i, __ddiv_rn(__dsub_rn(ld_stream(xr + i), m), s));
}

// Reductions ACROSS the rows of a chunk (the axis is not the production axis,
// protools.py:538-543, 592-595): per column np.(nan)mean and np.(nan)std over
// the rows, rows added in order as numpy's axis-0 reduction does.  One thread
// per column, coalesced across the warp.  y (optional): (x - mean) / std.
__global__ void __launch_bounds__(256)
col_moments_kernel(const double *__restrict__ x, int64_t ldx, int64_t rows, int64_t n,
                   int ignore_nan, double *__restrict__ mean_out, double *__restrict__ std_out,
                   double *__restrict__ y, int64_t ldy) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double s = 0.0, c = 0.0;
    for (int64_t r = 0; r < rows; ++r) {
        const double v = x[r * ldx + i];
        if (!ignore_nan || v == v) {
            s += v;
            c += 1.0;
        }
    }
    const double m = s / c;
    double q = 0.0;
    for (int64_t r = 0; r < rows; ++r) {
        const double v = x[r * ldx + i];
        if (!ignore_nan || v == v) {
            const double d = v - m;
            q = __dadd_rn(q, __dmul_rn(d, d));   // multiply, then add: as numpy rounds
        }
    }
    const double sd = sqrt(q / c);
    if (mean_out) mean_out[i] = m;
    if (std_out) std_out[i] = sd;
    if (y)
        for (int64_t r = 0; r < rows; ++r)
            y[r * ldy + i] = __ddiv_rn(__dsub_rn(x[r * ldx + i], m), sd);
}

static int grid_x(int64_t n, int64_t rows, int per_thread) {
    int64_t bx = (n + 256 * per_thread - 1) / (256 * per_thread);
    const int64_t cap = ((int64_t)sm_count() * 8 + rows - 1) / rows;
    if (bx > cap) bx = cap;
    return (int)(bx < 1 ? 1 : bx);
}

}  // namespace osz

using namespace osz;

extern "C" {

int osz_take_cols_f64(const double *x, int64_t ldx, int64_t rows, const int64_t *idx_dev,
                      int64_t nkeep, double *y, int64_t ldy, void *stream) {
    if (!x || !idx_dev || !y) return fail(OSZ_ERR_ARG, "osz_take_cols_f64: null argument");
    if (rows <= 0 || nkeep <= 0) return OSZ_OK;
    if (rows > 65535) return fail(OSZ_ERR_UNSUPPORTED, "take_cols: more than 65535 rows per call");
    take_cols_kernel<<<dim3((unsigned)grid_x(nkeep, rows, 4), (unsigned)rows), 256, 0,
                       as_stream(stream)>>>(x, ldx, idx_dev, nkeep, y, ldy);
    OSZ_LAUNCHED("take_cols_kernel");
    return OSZ_OK;
}

int osz_row_moments_slots(void) { return MOM_SLOTS; }

int osz_row_moments_f64(const double *x, int64_t ldx, int64_t rows, int64_t n, int ignore_nan,
                        double *acc, double *scratch, void *stream) {
    if (!x || !acc || !scratch) return fail(OSZ_ERR_ARG, "osz_row_moments_f64: null argument");
    if (rows <= 0 || n <= 0) return OSZ_OK;
    if (rows > 65535) return fail(OSZ_ERR_UNSUPPORTED, "row_moments: more than 65535 rows per call");
    int slots = grid_x(n, rows, 16);
    if (slots > MOM_SLOTS) slots = MOM_SLOTS;
    row_moments_partial_kernel<<<dim3((unsigned)slots, (unsigned)rows), 256, 0,
                                 as_stream(stream)>>>(x, ldx, n, ignore_nan, scratch);
    OSZ_LAUNCHED("row_moments_partial_kernel");
    row_moments_fold_kernel<<<(unsigned)((rows + 127) / 128), 128, 0, as_stream(stream)>>>(
        scratch, slots, rows, n, acc);
    OSZ_LAUNCHED("row_moments_fold_kernel");
    return OSZ_OK;
}

int osz_row_standardize_f64(const double *x, int64_t ldx, int64_t rows, int64_t n,
                            const double *mean_dev, const double *std_dev, double *y, int64_t ldy,
                            void *stream) {
    if (!x || !mean_dev || !std_dev || !y)
        return fail(OSZ_ERR_ARG, "osz_row_standardize_f64: null argument");
    if (rows <= 0 || n <= 0) return OSZ_OK;
    if (rows > 65535) return fail(OSZ_ERR_UNSUPPORTED, "standardize: more than 65535 rows per call");
    row_standardize_kernel<<<dim3((unsigned)grid_x(n, rows, 4), (unsigned)rows), 256, 0,
                             as_stream(stream)>>>(x, ldx, n, mean_dev, std_dev, y, ldy);
    OSZ_LAUNCHED("row_standardize_kernel");
    return OSZ_OK;
}

int osz_col_moments_f64(const double *x, int64_t ldx, int64_t rows, int64_t n, int ignore_nan,
                        double *mean_out, double *std_out, double *y, int64_t ldy, void *stream) {
    if (!x || (!mean_out && !std_out && !y))
        return fail(OSZ_ERR_ARG, "osz_col_moments_f64: null argument");
    if (rows <= 0 || n <= 0) return OSZ_OK;
    col_moments_kernel<<<(unsigned)((n + 255) / 256), 256, 0, as_stream(stream)>>>(
        x, ldx, rows, n, ignore_nan, mean_out, std_out, y, ldy);
    OSZ_LAUNCHED("col_moments_kernel");
    return OSZ_OK;
}

}  // extern "C"
