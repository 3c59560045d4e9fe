// common.cuh - shared helpers for the libcmr_b200 kernels (sm_100a only).
#pragma once
#include <cuda.h>
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

// ---- TMA bulk copies (cp.async.bulk, no tensor map) + mbarrier ---------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// global -> shared, completion counted in bytes on `bar` (16-byte aligned addresses and size)
__device__ __forceinline__ void bulk_g2s(void *smem_dst, const void *gsrc, unsigned bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// shared -> global
__device__ __forceinline__ void bulk_s2g(void *gdst, const void *smem_src, unsigned bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(smem_src)),
                 "r"(bytes)
                 : "memory");
}
// tiled TMA through a tensor map (3-D, tile mode): one instruction moves a whole [rows][cols] box
__device__ __forceinline__ void tma_load_3d(void *smem_dst, const CUtensorMap *map, int x, int y, int z, uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::
            "r"(smem_u32(smem_dst)),
        "l"(map), "r"(x), "r"(y), "r"(z), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap *map, int x, int y, int z, const void *smem_src) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%1, %2, %3}], [%4];" ::"l"(map), "r"(x),
                 "r"(y), "r"(z), "r"(smem_u32(smem_src))
                 : "memory");
}
__device__ __forceinline__ void tma_prefetch_map(const CUtensorMap *map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read_all() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_proxy() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// programmatic dependent launch: let the next kernel of the stream start its preamble / wait for
// the previous kernel's memory before touching what it produced
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// coherent 16-byte load that bypasses L1 (data another grid produced while this one was already resident)
__device__ __forceinline__ uint4 ld_cg_u4(const void *p) {
    uint4 v;
    asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ int ld_cg_s32(const int *p) {
    int v;
    asm volatile("ld.global.cg.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

__device__ __forceinline__ int warp_sum(int v) { return __reduce_add_sync(kFull, v); }

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    return v;
}

// ---- optional per-CTA timing instrumentation (debug builds only: -DCMR_DBG_TIMING) ---------------------
#ifdef CMR_DBG_TIMING
__device__ unsigned long long g_dbg[16 * 8192];
__device__ __forceinline__ unsigned long long dbg_now() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
#define DBG_MARK(slot)                                                                          \
    do {                                                                                        \
        if (threadIdx.x == 0) {                                                                 \
            int _id = blockIdx.y * gridDim.x + blockIdx.x;                                      \
            if (_id < 8192) g_dbg[_id * 16 + (slot)] = dbg_now();                                \
        }                                                                                       \
    } while (0)
#else
#define DBG_MARK(slot) do { } while (0)
#endif

}  // namespace cmr
