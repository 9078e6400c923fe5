// Shared helpers for the osz_b200 kernels (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>
#include <string>

#include "../../include/osz_b200.h"

namespace osz {

// ---- error plumbing ------------------------------------------------------
void set_error(const std::string &msg);
extern std::atomic<long long> g_launches;

inline int fail(osz_status code, const std::string &msg) {
    set_error(msg);
    return (int)code;
}

#define OSZ_CUDA(call)                                                          \
    do {                                                                        \
        cudaError_t err__ = (call);                                             \
        if (err__ != cudaSuccess) {                                             \
            return ::osz::fail(OSZ_ERR_CUDA, std::string(#call) + ": " +        \
                                                 cudaGetErrorString(err__));    \
        }                                                                       \
    } while (0)

// Check the launch that was just issued and count it.
#define OSZ_LAUNCHED(name)                                                      \
    do {                                                                        \
        cudaError_t err__ = cudaGetLastError();                                 \
        if (err__ != cudaSuccess) {                                             \
            return ::osz::fail(OSZ_ERR_CUDA, std::string(name) + " launch: " +  \
                                                 cudaGetErrorString(err__));    \
        }                                                                       \
        ::osz::g_launches.fetch_add(1, std::memory_order_relaxed);              \
    } while (0)

inline cudaStream_t as_stream(void *s) { return reinterpret_cast<cudaStream_t>(s); }

int sm_count();   // SMs of the current device (cached)
// stream-ordered scratch of a launch; release with cudaFreeAsync on the same stream
cudaError_t scratch_alloc(void **ptr, size_t bytes, cudaStream_t st);

// ---- device helpers --------------------------------------------------------
__device__ __forceinline__ double ldg(const double *p) { return __ldg(p); }
__device__ __forceinline__ double2 ldg(const double2 *p) { return __ldg(p); }

// streaming (evict-first) loads / stores for data touched once
__device__ __forceinline__ double ld_stream(const double *p) {
    double v;
    asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ void st_stream(double *p, double v) {
    asm volatile("st.global.L1::no_allocate.f64 [%0], %1;" ::"l"(p), "d"(v));
}

__device__ __forceinline__ float ld_stream(const float *p) {
    float v;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ void st_stream(float *p, float v) {
    asm volatile("st.global.L1::no_allocate.f32 [%0], %1;" ::"l"(p), "f"(v));
}

// ---- cp.async (LDGSTS): global -> shared without passing through registers ----------
__device__ __forceinline__ void cp_async8_zfill(uint32_t dst_smem, const void *src, bool valid) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(dst_smem), "l"(src),
                 "r"(valid ? 8 : 0)
                 : "memory");
}
__device__ __forceinline__ void cp_async_elem(uint32_t dst_smem, const double *src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst_smem), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_elem(uint32_t dst_smem, const float *src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst_smem), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int NPENDING>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(NPENDING) : "memory");
}

// ---- mbarrier + 1-D TMA bulk copy (global -> shared::cta) -------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t phase) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(phase)
        : "memory");
}
// bytes and both addresses must be multiples of 16
__device__ __forceinline__ void tma_load_1d(void *dst_smem, const void *src_gmem, uint32_t bytes,
                                            uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
            "r"(smem_u32(dst_smem)),
        "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t phase) {
    uint32_t ok;
    asm volatile(
        "{\n.reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(phase)
        : "memory");
    return ok != 0;
}

// One thread: fetch `need` (> 0) contiguous doubles starting at `src` into the
// shared buffer `dst` (16-byte aligned) with one bulk copy that completes on
// `bar` (one arrival).  Bulk copies move multiples of 16 bytes between 16-byte
// aligned addresses, so the copy starts at the aligned address at or below
// `src` -- element i of the span lands at dst[span_mis(src) + i] -- and an odd
// tail element is copied by hand before the (releasing) arrive.  Reading the
// element below an unaligned `src` is safe: allocations start 256-byte aligned,
// so it belongs to the same allocation.
__device__ __forceinline__ int span_mis(const double *src) {
    return (int)((reinterpret_cast<uintptr_t>(src) >> 3) & 1);
}
__device__ __forceinline__ void tma_fetch_span(double *dst, const double *src, int64_t need,
                                               uint64_t *bar) {
    const int mis = span_mis(src);
    int64_t cnt = need + mis;
    if (cnt & 1) {
        dst[cnt - 1] = src[need - 1];
        cnt -= 1;
    }
    mbar_expect_tx(bar, (uint32_t)cnt * 8u);
    if (cnt > 0) tma_load_1d(dst, src - mis, (uint32_t)cnt * 8u, bar);
}

// float32 samples (the opt-in float32 I/O mode): the same span fetch with four
// elements per 16 bytes; up to three tail elements are copied by hand.
__device__ __forceinline__ int span_mis(const float *src) {
    return (int)((reinterpret_cast<uintptr_t>(src) >> 2) & 3);
}
__device__ __forceinline__ void tma_fetch_span(float *dst, const float *src, int64_t need,
                                               uint64_t *bar) {
    const int mis = span_mis(src);
    int64_t cnt = need + mis;
    const int tail = (int)(cnt & 3);
    for (int i = 0; i < tail; ++i) dst[cnt - tail + i] = src[need - tail + i];
    cnt -= tail;
    mbar_expect_tx(bar, (uint32_t)cnt * 4u);
    if (cnt > 0) tma_load_1d(dst, src - mis, (uint32_t)cnt * 4u, bar);
}

}  // namespace osz
