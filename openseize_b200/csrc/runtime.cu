// Library plumbing: error state, device info, memory / stream wrappers and the
// layout kernels (pack / unpack / widen).  No DSP here.
#include "common.cuh"

namespace osz {

static thread_local std::string t_error;
std::atomic<long long> g_launches{0};

void set_error(const std::string &msg) { t_error = msg; }

int sm_count() {
    static thread_local int cached_dev = -1;
    static thread_local int cached = 0;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (dev != cached_dev) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
            n = 148;
        cached = n;
        cached_dev = dev;
    }
    return cached;
}

// Stream-ordered scratch for a launch (cudaMallocAsync / cudaFreeAsync: no device
// synchronisation, no sharing between streams).  The default memory pool hands its
// memory back to the driver at every synchronisation point unless told otherwise,
// which makes the next allocation cost milliseconds: keep it cached (once per device).
cudaError_t scratch_alloc(void **ptr, size_t bytes, cudaStream_t st) {
    static thread_local int tuned_dev = -1;
    int dev = 0;
    cudaError_t err = cudaGetDevice(&dev);
    if (err != cudaSuccess) return err;
    if (dev != tuned_dev) {
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
            unsigned long long keep = ~0ull;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
        tuned_dev = dev;
    }
    return cudaMallocAsync(ptr, bytes, st);
}

// (outer, n, inner) -> rows (outer*inner, n): tiled transpose through shared
// memory so both sides are coalesced.
template <typename T, bool PACK>
__global__ void transpose_rows_kernel(const T *__restrict__ src, T *__restrict__ dst, int64_t outer,
                                      int64_t n, int64_t inner, int64_t ld) {
    __shared__ T tile[32][33];
    const int64_t o = blockIdx.z;
    const int64_t t0 = (int64_t)blockIdx.x * 32;
    const int64_t i0 = (int64_t)blockIdx.y * 32;
    const int tx = threadIdx.x, ty = threadIdx.y;  // 32 x 8
    if (PACK) {
        // read (t, i) with i fastest; write row (o*inner+i), t fastest
        for (int k = ty; k < 32; k += 8) {
            int64_t t = t0 + k, i = i0 + tx;
            if (t < n && i < inner) tile[k][tx] = src[(o * n + t) * inner + i];
        }
        __syncthreads();
        for (int k = ty; k < 32; k += 8) {
            int64_t i = i0 + k, t = t0 + tx;
            if (t < n && i < inner) dst[(o * inner + i) * ld + t] = tile[tx][k];
        }
    } else {
        for (int k = ty; k < 32; k += 8) {
            int64_t i = i0 + k, t = t0 + tx;
            if (t < n && i < inner) tile[k][tx] = src[(o * inner + i) * ld + t];
        }
        __syncthreads();
        for (int k = ty; k < 32; k += 8) {
            int64_t t = t0 + k, i = i0 + tx;
            if (t < n && i < inner) dst[(o * n + t) * inner + i] = tile[tx][k];
        }
    }
}

template <typename S>
__global__ void widen_kernel(const S *__restrict__ src, double *__restrict__ dst, int64_t count) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t step = (int64_t)gridDim.x * blockDim.x;
    for (; i < count; i += step) dst[i] = (double)src[i];
}

template <typename S, typename D = double>
__global__ void widen_rows_kernel(const S *__restrict__ src, int64_t ld_src,
                                  D *__restrict__ dst, int64_t ld_dst, int64_t n) {
    const int64_t row = blockIdx.y;
    const S *s = src + row * ld_src;
    D *d = dst + row * ld_dst;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t step = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += step) d[i] = (D)s[i];
}

// EDF ingest on the device.  `rec` holds whole data records exactly as they sit
// in the file: record r = [ch0: spr samples | ch1: spr samples | ...] (per_record
// int16 each).  Row c of the output is channel chan_off[c] (its offset inside a
// record), samples skip .. skip+n-1 counted from the first record:
//     dst[c][i] = rec[(skip+i) / spr][chan_off[c] + (skip+i) % spr] * slope[c] + offset[c]
// with a separate multiply and add (two roundings), so the rows are bit-identical
// to the reference's `arr * slopes; result += offsets` (file_io/edf.py:412-419).
__global__ void decode_edf_records_kernel(const int16_t *__restrict__ rec, int64_t per_record,
                                          int64_t spr, const int *__restrict__ chan_off,
                                          const double *__restrict__ slope,
                                          const double *__restrict__ offset, int64_t skip,
                                          double *__restrict__ dst, int64_t ld_dst, int64_t n) {
    const int64_t row = blockIdx.y;
    double *d = dst + row * ld_dst;
    const double a = slope[row], b = offset[row];
    const int16_t *base = rec + chan_off[row];
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t step = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += step) {
        const int64_t t = skip + i;
        const int64_t r = t / spr;
        const int16_t v = base[r * per_record + (t - r * spr)];
        d[i] = __dadd_rn(__dmul_rn((double)v, a), b);
    }
}

template <typename S, typename D = double>
static int launch_widen_rows(const S *src, int64_t ld_src, D *dst, int64_t ld_dst,
                             int64_t rows, int64_t n, void *stream) {
    if (rows <= 0 || n <= 0) return OSZ_OK;
    if (rows > 65535) return fail(OSZ_ERR_UNSUPPORTED, "widen: more than 65535 rows per call");
    int bx = (int)((n + 255) / 256);
    const int cap = (sm_count() * 16 + (int)rows - 1) / (int)rows;
    if (bx > cap) bx = cap < 1 ? 1 : cap;
    widen_rows_kernel<S, D><<<dim3((unsigned)bx, (unsigned)rows), 256, 0, as_stream(stream)>>>(
        src, ld_src, dst, ld_dst, n);
    OSZ_LAUNCHED("widen_rows");
    return OSZ_OK;
}

template <typename T, bool PACK>
static int launch_transpose(const T *src, T *dst, int64_t outer, int64_t n, int64_t inner,
                            int64_t ld, void *stream) {
    if (outer <= 0 || n <= 0 || inner <= 0) return OSZ_OK;
    if (outer > 65535 || (inner + 31) / 32 > 65535)
        return fail(OSZ_ERR_UNSUPPORTED, "pack/unpack: outer or inner extent too large");
    dim3 grid((unsigned)((n + 31) / 32), (unsigned)((inner + 31) / 32), (unsigned)outer);
    transpose_rows_kernel<T, PACK><<<grid, dim3(32, 8), 0, as_stream(stream)>>>(src, dst, outer, n,
                                                                                 inner, ld);
    OSZ_LAUNCHED("transpose_rows");
    return OSZ_OK;
}

}  // namespace osz

using namespace osz;

extern "C" {

int osz_version(void) { return 100; }

const char *osz_last_error(void) { return t_error.c_str(); }

int64_t osz_launch_count(void) { return (int64_t)g_launches.load(); }

int osz_device_info(int dev, int *sms, int *cc_major, int *cc_minor, int64_t *smem_optin,
                    int64_t *global_mem) {
    cudaDeviceProp p;
    OSZ_CUDA(cudaGetDeviceProperties(&p, dev));
    if (sms) *sms = p.multiProcessorCount;
    if (cc_major) *cc_major = p.major;
    if (cc_minor) *cc_minor = p.minor;
    if (smem_optin) *smem_optin = (int64_t)p.sharedMemPerBlockOptin;
    if (global_mem) *global_mem = (int64_t)p.totalGlobalMem;
    return OSZ_OK;
}

int osz_dev_malloc(void **p, size_t bytes) {
    if (!p) return fail(OSZ_ERR_ARG, "osz_dev_malloc: null out pointer");
    OSZ_CUDA(cudaMalloc(p, bytes));
    return OSZ_OK;
}
int osz_dev_free(void *p) {
    OSZ_CUDA(cudaFree(p));
    return OSZ_OK;
}
int osz_host_alloc(void **p, size_t bytes) {
    if (!p) return fail(OSZ_ERR_ARG, "osz_host_alloc: null out pointer");
    OSZ_CUDA(cudaHostAlloc(p, bytes, cudaHostAllocDefault));
    return OSZ_OK;
}
int osz_host_free(void *p) {
    OSZ_CUDA(cudaFreeHost(p));
    return OSZ_OK;
}
int osz_stream_create(void **s) {
    if (!s) return fail(OSZ_ERR_ARG, "osz_stream_create: null out pointer");
    cudaStream_t st;
    OSZ_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    *s = (void *)st;
    return OSZ_OK;
}
int osz_stream_destroy(void *s) {
    OSZ_CUDA(cudaStreamDestroy(as_stream(s)));
    return OSZ_OK;
}
int osz_stream_sync(void *s) {
    OSZ_CUDA(cudaStreamSynchronize(as_stream(s)));
    return OSZ_OK;
}
int osz_memcpy_h2d_async(void *dst, const void *src, size_t bytes, void *s) {
    OSZ_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, as_stream(s)));
    return OSZ_OK;
}
int osz_memcpy_d2h_async(void *dst, const void *src, size_t bytes, void *s) {
    OSZ_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, as_stream(s)));
    return OSZ_OK;
}
int osz_memcpy_d2d_async(void *dst, const void *src, size_t bytes, void *s) {
    OSZ_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, as_stream(s)));
    return OSZ_OK;
}
int osz_memcpy2d_h2d_async(void *dst, size_t dpitch, const void *src, size_t spitch, size_t width,
                           size_t height, void *s) {
    OSZ_CUDA(cudaMemcpy2DAsync(dst, dpitch, src, spitch, width, height, cudaMemcpyHostToDevice,
                               as_stream(s)));
    return OSZ_OK;
}
int osz_memcpy2d_d2h_async(void *dst, size_t dpitch, const void *src, size_t spitch, size_t width,
                           size_t height, void *s) {
    OSZ_CUDA(cudaMemcpy2DAsync(dst, dpitch, src, spitch, width, height, cudaMemcpyDeviceToHost,
                               as_stream(s)));
    return OSZ_OK;
}
int osz_memset_async(void *dst, int value, size_t bytes, void *s) {
    OSZ_CUDA(cudaMemsetAsync(dst, value, bytes, as_stream(s)));
    return OSZ_OK;
}

int osz_pack_rows_f64(const double *src, int64_t outer, int64_t n, int64_t inner, double *dst,
                      int64_t ld, void *stream) {
    return launch_transpose<double, true>(src, dst, outer, n, inner, ld, stream);
}
int osz_unpack_rows_f64(const double *src, int64_t ld, int64_t outer, int64_t n, int64_t inner,
                        double *dst, void *stream) {
    return launch_transpose<double, false>(src, dst, outer, n, inner, ld, stream);
}
int osz_unpack_rows_c128(const double *src, int64_t ld, int64_t outer, int64_t n, int64_t inner,
                         double *dst, void *stream) {
    return launch_transpose<double2, false>(reinterpret_cast<const double2 *>(src),
                                            reinterpret_cast<double2 *>(dst), outer, n, inner, ld,
                                            stream);
}
// (re, im) rows -> interleaved complex128 rows: z[r][t] = re[r][t] + i im[r][t]
__global__ void zip_complex_kernel(const double *__restrict__ re, int64_t ldre,
                                   const double *__restrict__ im, int64_t ldim, int64_t n,
                                   double2 *__restrict__ z) {
    const int64_t row = blockIdx.y;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n;
         t += (int64_t)gridDim.x * blockDim.x)
        z[row * n + t] = make_double2(ld_stream(re + row * ldre + t), ld_stream(im + row * ldim + t));
}
int osz_zip_complex_f64(const double *re, int64_t ldre, const double *im, int64_t ldim,
                        int64_t rows, int64_t n, double *z, void *stream) {
    if (!re || !im || !z) return fail(OSZ_ERR_ARG, "osz_zip_complex_f64: null argument");
    if (rows <= 0 || n <= 0) return OSZ_OK;
    if (rows > 65535) return fail(OSZ_ERR_UNSUPPORTED, "osz_zip_complex_f64: more than 65535 rows");
    int bx = (int)((n + 255) / 256);
    const int cap = (sm_count() * 16 + (int)rows - 1) / (int)rows;
    if (bx > cap) bx = cap;
    if (bx < 1) bx = 1;
    zip_complex_kernel<<<dim3((unsigned)bx, (unsigned)rows), 256, 0, as_stream(stream)>>>(
        re, ldre, im, ldim, n, reinterpret_cast<double2 *>(z));
    OSZ_LAUNCHED("zip_complex_kernel");
    return OSZ_OK;
}

int osz_widen_f32_f64(const float *src, double *dst, int64_t count, void *stream) {
    if (count <= 0) return OSZ_OK;
    int blocks = (int)((count + 255) / 256);
    if (blocks > sm_count() * 16) blocks = sm_count() * 16;
    widen_kernel<float><<<blocks, 256, 0, as_stream(stream)>>>(src, dst, count);
    OSZ_LAUNCHED("widen_f32");
    return OSZ_OK;
}
int osz_widen_i16_f64(const int16_t *src, double *dst, int64_t count, void *stream) {
    if (count <= 0) return OSZ_OK;
    int blocks = (int)((count + 255) / 256);
    if (blocks > sm_count() * 16) blocks = sm_count() * 16;
    widen_kernel<int16_t><<<blocks, 256, 0, as_stream(stream)>>>(src, dst, count);
    OSZ_LAUNCHED("widen_i16");
    return OSZ_OK;
}

int osz_widen_rows_f32_f64(const float *src, int64_t ld_src, double *dst, int64_t ld_dst,
                           int64_t rows, int64_t n, void *stream) {
    return launch_widen_rows<float>(src, ld_src, dst, ld_dst, rows, n, stream);
}
int osz_widen_rows_i16_f64(const int16_t *src, int64_t ld_src, double *dst, int64_t ld_dst,
                           int64_t rows, int64_t n, void *stream) {
    return launch_widen_rows<int16_t>(src, ld_src, dst, ld_dst, rows, n, stream);
}
// float32 I/O mode: rows between the two sample types (operators without a float32
// kernel of their own run in float64 between a widen and a narrow), int16 -> float32
int osz_narrow_rows_f64_f32(const double *src, int64_t ld_src, float *dst, int64_t ld_dst,
                            int64_t rows, int64_t n, void *stream) {
    return launch_widen_rows<double, float>(src, ld_src, dst, ld_dst, rows, n, stream);
}
int osz_widen_rows_i16_f32(const int16_t *src, int64_t ld_src, float *dst, int64_t ld_dst,
                           int64_t rows, int64_t n, void *stream) {
    return launch_widen_rows<int16_t, float>(src, ld_src, dst, ld_dst, rows, n, stream);
}
int osz_decode_edf_records_f64(const int16_t *rec, int64_t per_record, int64_t spr,
                               const int *chan_off_dev, const double *slope_dev,
                               const double *offset_dev, int64_t skip, double *dst, int64_t ld_dst,
                               int64_t rows, int64_t n, void *stream) {
    if (!rec || !chan_off_dev || !slope_dev || !offset_dev || !dst || spr < 1 || per_record < spr)
        return fail(OSZ_ERR_ARG, "osz_decode_edf_records_f64: bad arguments");
    if (rows <= 0 || n <= 0) return OSZ_OK;
    if (rows > 65535) return fail(OSZ_ERR_UNSUPPORTED, "decode: more than 65535 rows per call");
    int bx = (int)((n + 255) / 256);
    const int cap = (sm_count() * 16 + (int)rows - 1) / (int)rows;
    if (bx > cap) bx = cap < 1 ? 1 : cap;
    decode_edf_records_kernel<<<dim3((unsigned)bx, (unsigned)rows), 256, 0, as_stream(stream)>>>(
        rec, per_record, spr, chan_off_dev, slope_dev, offset_dev, skip, dst, ld_dst, n);
    OSZ_LAUNCHED("decode_edf_records");
    return OSZ_OK;
}

}  // extern "C"
