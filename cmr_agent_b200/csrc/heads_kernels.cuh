// heads_kernels.cuh - the agent's actor-critic heads (models/CMRAgent.py:70-86, used at :106-113): three small
// MLPs (Linear - LeakyReLU - Linear - LeakyReLU - Linear) on the [B, 4f] state embedding.  At the reference's batch
// sizes every Linear is a launch-bound GEMV/GEMM of its own (17 launches, ~100 us of an iteration at B = 32, a sixth of
// an iteration at B = 1); here layer l of ALL heads is one launch of k_grouped_linear: a group = the neurons of one
// head's layer, reading that head's slice of the previous layer's output.
//   One warp per output neuron: the weight row sits in registers (K <= 256), the lanes split K, eight batch rows are
//   accumulated at a time and reduced with shuffles; explicit fmaf in a fixed order - deterministic, batch-shape
//   independent (a row's result does not depend on B).  fp32 throughout.
#pragma once
#include "common.cuh"

namespace cmr {

constexpr int kLinMaxGroups = 8;
constexpr int kLinMaxK = 256;
constexpr int kLinRows = 8;
struct LinGroup {
    int in_off, K, n0, n1;   // neurons [n0, n1) read in[b][in_off, in_off + K)
    long long w_off;         // their weight rows: W + w_off + (n - n0) * K
};
struct LinGroups {
    LinGroup g[kLinMaxGroups];
    int count;
};

__global__ void __launch_bounds__(256) k_grouped_linear(const float *__restrict__ in, int in_stride, const float *__restrict__ W,
                                                        const float *__restrict__ bias, const __grid_constant__ LinGroups groups,
                                                        int B, int N, float slope, int activate, float *__restrict__ out,
                                                        int out_stride) {
    pdl_launch_dependents();
    pdl_wait();   // `in` is the previous layer's output
    const int lane = threadIdx.x & 31;
    const int n = (int)blockIdx.x * (int)(blockDim.x >> 5) + (int)(threadIdx.x >> 5);
    if (n >= N) return;
    int gi = 0;
    while (gi + 1 < groups.count && n >= groups.g[gi].n1) ++gi;
    const LinGroup g = groups.g[gi];
    const float *wrow = W + g.w_off + (size_t)(n - g.n0) * g.K;
    float w[kLinMaxK / 32];
#pragma unroll
    for (int i = 0; i < kLinMaxK / 32; ++i) {
        const int k = lane + 32 * i;
        w[i] = k < g.K ? __ldg(wrow + k) : 0.f;
    }
    const float bn = __ldg(bias + n);
    for (int b0 = 0; b0 < B; b0 += kLinRows) {
        // every load of the eight rows is issued before the first multiply-add (written as two phases: with the loads
        // inside the accumulation the compiler chained them through one register, eight round trips per row)
        float x[kLinRows][kLinMaxK / 32];
#pragma unroll
        for (int r = 0; r < kLinRows; ++r) {
            const float *row = in + (size_t)min(b0 + r, B - 1) * in_stride + g.in_off;
#pragma unroll
            for (int i = 0; i < kLinMaxK / 32; ++i) {
                const int k = lane + 32 * i;
                x[r][i] = k < g.K ? __ldg(row + k) : 0.f;
            }
        }
        float acc[kLinRows];
#pragma unroll
        for (int r = 0; r < kLinRows; ++r) {
            acc[r] = 0.f;
#pragma unroll
            for (int i = 0; i < kLinMaxK / 32; ++i) acc[r] = __fmaf_rn(w[i], x[r][i], acc[r]);   // (w = x = 0 beyond K)
        }
#pragma unroll
        for (int r = 0; r < kLinRows; ++r)
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) acc[r] = __fadd_rn(acc[r], __shfl_xor_sync(kFull, acc[r], o));
        float mine = acc[0];
#pragma unroll
        for (int r = 1; r < kLinRows; ++r) mine = lane == r ? acc[r] : mine;
        if (lane < kLinRows && b0 + lane < B) {
            float v = __fadd_rn(mine, bn);
            if (activate && v < 0.f) v = __fmul_rn(v, slope);   // LeakyReLU
            out[(size_t)(b0 + lane) * out_stride + n] = v;
        }
    }
}

}  // namespace cmr
