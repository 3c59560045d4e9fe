// common.cuh - shared helpers for the libcmr_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/cmr_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libcmr_b200 is written for sm_100a (B200) only"
#endif

namespace cmr {

constexpr int kWarp = 32;
constexpr unsigned kFull = 0xffffffffu;

// launch bookkeeping (bench.py reports it as gpu_launches)
inline unsigned long long g_launches = 0;
// sticky device fault word: 0 = none, 1 = index out of range, 2 = action out of range
__device__ int g_fault = 0;

inline int after_launch() {
    ++g_launches;
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? CMR_OK : (int)e;
}

inline bool aligned(const void *p, size_t a) { return (reinterpret_cast<uintptr_t>(p) & (a - 1)) == 0; }
inline size_t round_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// ---- arithmetic with the reference's evaluation order -------------------
// One output element of a k=3 bmm as torch's CPU path evaluates it.  Two regimes (measured, see
// oracle/cmr_oracle.c): products with rows*cols*k >= 400 (3 x n, n >= 45: whole clouds) run the
// FMA chain fma(a2,z, fma(a1,y, a0*x)); smaller ones (the 3x3 pose products, clouds of < 45 points)
// run the plain loop (a0*x + a1*y) + a2*z without fusion.
__device__ __forceinline__ float dot3_chain(float a0, float a1, float a2, float x, float y, float z) {
    return __fmaf_rn(a2, z, __fmaf_rn(a1, y, __fmul_rn(a0, x)));
}
__device__ __forceinline__ float dot3_plain(float a0, float a1, float a2, float x, float y, float z) {
    return __fadd_rn(__fadd_rn(__fmul_rn(a0, x), __fmul_rn(a1, y)), __fmul_rn(a2, z));
}
template <bool kChain>
__device__ __forceinline__ float dot3(float a0, float a1, float a2, float x, float y, float z) {
    return kChain ? dot3_chain(a0, a1, a2, x, y, z) : dot3_plain(a0, a1, a2, x, y, z);
}
constexpr int kBmmChainMinCols = 45;   // 3*3*n >= 400
// sum((a-b)**2, -1): unfused, left to right (pointnet_util.py:33,67)
__device__ __forceinline__ float sqdist3(float ax, float ay, float az, float bx, float by, float bz) {
    float dx = __fsub_rn(ax, bx), dy = __fsub_rn(ay, by), dz = __fsub_rn(az, bz);
    return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}

// ---- memory helpers ------------------------------------------------------------------------------
__device__ __forceinline__ float4 ldg_stream4(const float *p) {  // read-once data: do not keep in L1
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "l"(p));
    return v;
}
__device__ __forceinline__ void stg_stream4(float *p, float4 v) {  // write-once data: streaming store
    asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
                 : "memory");
}
__device__ __forceinline__ void stg_stream1(float *p, float v) {
    asm volatile("st.global.cs.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}

__device__ __forceinline__ int warp_sum(int v) { return __reduce_add_sync(kFull, v); }

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    return v;
}

}  // namespace cmr
