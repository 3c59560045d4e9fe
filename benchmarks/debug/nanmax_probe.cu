// probe: what max.NaN.f32 / atomicMax(unsigned) do with NaN keys on sm_100a
#include <cstdio>
#include <cuda_runtime.h>
#include <math.h>
__device__ __forceinline__ float max_nan(float a, float b) { float r; asm volatile("max.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ unsigned f2key(float v) { unsigned b = __float_as_uint(v); return (b & 0x80000000u) ? ~b : (b | 0x80000000u); }
__global__ void k(const float *in, unsigned *out, unsigned *keys) {
    float nanv = in[0] - in[1];           // inf - inf
    float a = max_nan(-INFINITY, nanv);
    float b = max_nan(nanv, 0.2f * nanv);
    float c = max_nan(b, __shfl_xor_sync(0xffffffffu, b, 1));
    if (threadIdx.x == 0) {
        out[0] = __float_as_uint(nanv); out[1] = __float_as_uint(a); out[2] = __float_as_uint(b); out[3] = __float_as_uint(c);
        out[4] = f2key(a); out[5] = f2key(-0.0f); out[6] = f2key(-INFINITY);
    }
    atomicMax(keys + (threadIdx.x & 1), f2key(threadIdx.x == 3 ? a : -INFINITY));
}
int main() {
    float h[2] = {INFINITY, INFINITY}, *d; unsigned *o, *kk, ho[8], hk[2];
    cudaMalloc(&d, 8); cudaMalloc(&o, 32); cudaMalloc(&kk, 8); cudaMemset(kk, 0, 8);
    cudaMemcpy(d, h, 8, cudaMemcpyHostToDevice);
    k<<<1, 32>>>(d, o, kk);
    cudaMemcpy(ho, o, 32, cudaMemcpyDeviceToHost); cudaMemcpy(hk, kk, 8, cudaMemcpyDeviceToHost);
    printf("nan %08x maxnan(-inf,nan) %08x lrelu(nan) %08x shfl %08x key(a) %08x key(-0) %08x key(-inf) %08x | atomics %08x %08x err %d\n",
           ho[0], ho[1], ho[2], ho[3], ho[4], ho[5], ho[6], hk[0], hk[1], (int)cudaGetLastError());
    return 0;
}
