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

#define OSZ_SQRT1_2_ 0.70710678118654752440
#define OSZ_COS_PI_8_ 0.92387953251128675613
#define OSZ_SIN_PI_8_ 0.38268343236508977173

namespace osz {
// bit-level helpers of the FP64-token tie (see SyncPingPong), per real type
__device__ __forceinline__ double tie_bits(double a, int t) {
    return __hiloint2double(__double2hiint(a) ^ t, __double2loint(a));
}
__device__ __forceinline__ float tie_bits(float a, int t) {
    return __int_as_float(__float_as_int(a) ^ t);
}
__device__ __forceinline__ int hi_bits(double a) { return __double2hiint(a); }
__device__ __forceinline__ int hi_bits(float a) { return __float_as_int(a); }
__device__ __forceinline__ float2 ldg(const float2 *p) { return __ldg(p); }
__device__ __forceinline__ float ldg(const float *p) { return __ldg(p); }
}  // namespace osz

// float64 arithmetic: namespace osz
#define OSZ_FFTNS osz
#define OSZ_R double
#define OSZ_R2 double2
#define OSZ_MK2 make_double2
#define OSZ_K(x) (x)
#include "fft_core_body.inc"
#undef OSZ_FFTNS
#undef OSZ_R
#undef OSZ_R2
#undef OSZ_MK2
#undef OSZ_K

// float32 arithmetic (the opt-in float32 compute mode): namespace oszf
namespace oszf {
using namespace osz;   // shared helpers (smem_u32, barriers, ...); overloads resolve by type
}
#define OSZ_FFTNS oszf
#define OSZ_R float
#define OSZ_R2 float2
#define OSZ_MK2 make_float2
#define OSZ_K(x) ((float)(x))
#include "fft_core_body.inc"
#undef OSZ_FFTNS
#undef OSZ_R
#undef OSZ_R2
#undef OSZ_MK2
#undef OSZ_K
