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

// ---- epilogue of a convolution layer of the agent's 2-D head (models/CMRAgent.py:34-61), eval mode ---------------------
// After every 3x3 convolution the reference runs up to four elementwise launches over the whole feature map: the bias
// add, BatchNorm2d, LeakyReLU, AvgPool2d.  Folded (scale = g / sqrt(var + eps), shift = (bias - mean) * scale + beta;
// scale = 1, shift = bias without a BatchNorm) they are ONE pass:  y = pool(lrelu(x * scale[c] + shift[c])).
namespace cmr {

__device__ __forceinline__ float lrelu_affine(float x, float sc, float sh, float slope) {
    const float t = __fmaf_rn(x, sc, sh);
    return t < 0.f ? __fmul_rn(t, slope) : t;
}

// pool = 0.  One float4 per thread and step; HW % 4 == 0, so a float4 never crosses a channel plane.
__global__ void __launch_bounds__(256) k_conv_epilogue(const float *__restrict__ x, const float *__restrict__ scale,
                                                       const float *__restrict__ shift, float slope, long long n4, int HW4,
                                                       int C, float *__restrict__ y) {
    pdl_launch_dependents();
    pdl_wait();
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)((i / HW4) % C);
        const float sc = __ldg(scale + c), sh = __ldg(shift + c);
        float4 v = *reinterpret_cast<const float4 *>(x + 4 * i);
        v.x = lrelu_affine(v.x, sc, sh, slope);
        v.y = lrelu_affine(v.y, sc, sh, slope);
        v.z = lrelu_affine(v.z, sc, sh, slope);
        v.w = lrelu_affine(v.w, sc, sh, slope);
        *reinterpret_cast<float4 *>(y + 4 * i) = v;
    }
}

// pool = 1: AvgPool2d(2, 2).  A thread reads four pixels of two rows and writes two outputs; W % 4 == 0, H % 2 == 0.
__global__ void __launch_bounds__(256) k_conv_epilogue_pool2(const float *__restrict__ x, const float *__restrict__ scale,
                                                             const float *__restrict__ shift, float slope, long long n2, int H,
                                                             int W, int C, float *__restrict__ y) {
    pdl_launch_dependents();
    pdl_wait();
    const int Wo2 = W >> 2, Ho = H >> 1;   // output pairs per row, output rows
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += (long long)gridDim.x * blockDim.x) {
        const int wq = (int)(i % Wo2);
        const long long r = i / Wo2;
        const int ho = (int)(r % Ho);
        const long long plane = r / Ho;          // b * C + c
        const int c = (int)(plane % C);
        const float sc = __ldg(scale + c), sh = __ldg(shift + c);
        const float *p = x + (plane * H + 2 * ho) * W + 4 * wq;
        const float4 a = *reinterpret_cast<const float4 *>(p), b = *reinterpret_cast<const float4 *>(p + W);
        // torch's avg_pool2d: the window's values added row by row, then divided by its area
        const float s0 = __fadd_rn(__fadd_rn(__fadd_rn(lrelu_affine(a.x, sc, sh, slope), lrelu_affine(a.y, sc, sh, slope)),
                                             lrelu_affine(b.x, sc, sh, slope)), lrelu_affine(b.y, sc, sh, slope));
        const float s1 = __fadd_rn(__fadd_rn(__fadd_rn(lrelu_affine(a.z, sc, sh, slope), lrelu_affine(a.w, sc, sh, slope)),
                                             lrelu_affine(b.z, sc, sh, slope)), lrelu_affine(b.w, sc, sh, slope));
        *reinterpret_cast<float2 *>(y + (plane * Ho + ho) * (W >> 1) + 2 * wq) = make_float2(__fmul_rn(s0, 0.25f), __fmul_rn(s1, 0.25f));
    }
}

// pool = 2: AvgPool2d((H, W)) - one warp per (episode, channel) plane
__global__ void __launch_bounds__(256) k_conv_epilogue_global(const float *__restrict__ x, const float *__restrict__ scale,
                                                              const float *__restrict__ shift, float slope, int planes, int HW,
                                                              int C, float *__restrict__ y) {
    pdl_launch_dependents();
    pdl_wait();
    const int lane = threadIdx.x & 31;
    const int plane = (int)blockIdx.x * (int)(blockDim.x >> 5) + (int)(threadIdx.x >> 5);
    if (plane >= planes) return;
    const int c = plane % C;
    const float sc = __ldg(scale + c), sh = __ldg(shift + c);
    float s = 0.f;
    for (int i = lane; i < HW; i += 32) s = __fadd_rn(s, lrelu_affine(x[(size_t)plane * HW + i], sc, sh, slope));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s = __fadd_rn(s, __shfl_xor_sync(kFull, s, o));
    if (lane == 0) y[plane] = __fdiv_rn(s, (float)HW);
}

}  // namespace cmr

// ---- the same epilogues on channels-last data ([B][H][W][C] in memory): cuDNN's tensor-core convolutions are NHWC
// kernels - handed NCHW tensors they convert input and output of every layer (16 layout launches per forward, half of
// the convolutions' time at every batch size).  With the epilogue ours, the whole head can stay channels-last.
namespace cmr {

__device__ __forceinline__ float4 lrelu_affine4(float4 v, float4 sc, float4 sh, float slope) {
    return make_float4(lrelu_affine(v.x, sc.x, sh.x, slope), lrelu_affine(v.y, sc.y, sh.y, slope),
                       lrelu_affine(v.z, sc.z, sh.z, slope), lrelu_affine(v.w, sc.w, sh.w, slope));
}

__global__ void __launch_bounds__(256) k_conv_epilogue_nhwc(const float *__restrict__ x, const float *__restrict__ scale,
                                                            const float *__restrict__ shift, float slope, long long n4, int C4,
                                                            float *__restrict__ y) {
    pdl_launch_dependents();
    pdl_wait();
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        const int c4 = (int)(i % C4);
        const float4 sc = __ldg(reinterpret_cast<const float4 *>(scale) + c4), sh = __ldg(reinterpret_cast<const float4 *>(shift) + c4);
        reinterpret_cast<float4 *>(y)[i] = lrelu_affine4(reinterpret_cast<const float4 *>(x)[i], sc, sh, slope);
    }
}

// AvgPool2d(2, 2): a thread = four channels of one output pixel
__global__ void __launch_bounds__(256) k_conv_epilogue_pool2_nhwc(const float *__restrict__ x, const float *__restrict__ scale,
                                                                  const float *__restrict__ shift, float slope, long long n4,
                                                                  int H, int W, int C4, float *__restrict__ y) {
    pdl_launch_dependents();
    pdl_wait();
    const int Ho = H >> 1, Wo = W >> 1;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        const int c4 = (int)(i % C4);
        long long r = i / C4;
        const int wo = (int)(r % Wo);
        r /= Wo;
        const int ho = (int)(r % Ho);
        const long long b = r / Ho;
        const float4 sc = __ldg(reinterpret_cast<const float4 *>(scale) + c4), sh = __ldg(reinterpret_cast<const float4 *>(shift) + c4);
        const float4 *p = reinterpret_cast<const float4 *>(x) + ((b * H + 2 * ho) * W + 2 * wo) * C4 + c4;
        const float4 a = lrelu_affine4(p[0], sc, sh, slope), bq = lrelu_affine4(p[C4], sc, sh, slope);
        const float4 c = lrelu_affine4(p[(long long)W * C4], sc, sh, slope), d = lrelu_affine4(p[(long long)W * C4 + C4], sc, sh, slope);
        float4 o;   // the window's values added row by row, then divided by its area (torch's avg_pool2d)
        o.x = __fmul_rn(__fadd_rn(__fadd_rn(__fadd_rn(a.x, bq.x), c.x), d.x), 0.25f);
        o.y = __fmul_rn(__fadd_rn(__fadd_rn(__fadd_rn(a.y, bq.y), c.y), d.y), 0.25f);
        o.z = __fmul_rn(__fadd_rn(__fadd_rn(__fadd_rn(a.z, bq.z), c.z), d.z), 0.25f);
        o.w = __fmul_rn(__fadd_rn(__fadd_rn(__fadd_rn(a.w, bq.w), c.w), d.w), 0.25f);
        reinterpret_cast<float4 *>(y)[i] = o;
    }
}

// The deterministic action of models/CMRAgent.py:118-124: argmax over Categorical(logits=x).probs, for the rotation and
// the translation logits in ONE launch (torch: 2 x 11 launches - amax, abs, eq, masked_fill, sub, exp, sum, log, add, sub,
// softmax, argmax).  A half-warp per row of S logits, 9 <= S <= 16 (the reference: 11).  probs must be torch's bit for
// bit - two different logits may round to the same probability, and argmax then takes the smaller index - so every
// operation and the ORDER of both sums are torch's:
//   Categorical.__init__:  n = x - logsumexp(x)     logsumexp = log(sum(exp(x - m))) + (|m| == inf ? 0 : m), m = max x
//       the sum (ATen reduce_kernel, 9..16 inputs per output: 16 lanes along the row, halving offsets 8, 4, 2, 1; measured
//       against five other orders, benchmarks/debug/action_diag/diag.py):  a_j = x[j] + x[j + 8], b_j = a_j + a_{j+4}, ...
//   .probs = softmax(n)  (ATen softmax_warp_forward, 16 lanes per row): e = exp(n - max n), the sum by the same xor
//       butterfly 8, 4, 2, 1 (the same value in every lane, floating-point addition being commutative), p = e / sum
//   argmax: the FIRST index of the largest probability.
// tests/test_gpu_tower.py compares probs and actions with torch's own on random and near-tied rows.
__global__ void __launch_bounds__(128) k_deterministic_action(const float *__restrict__ xr, int Dr, long long sr,
                                                              const float *__restrict__ xt, int Dt, long long st_, int B, int S,
                                                              long long *__restrict__ ar, long long *__restrict__ at,
                                                              float *__restrict__ pr, float *__restrict__ pt) {
    pdl_launch_dependents();
    pdl_wait();
    const int lane = threadIdx.x & 31, l = lane & 15, half = lane & 16;
    const int row = ((int)blockIdx.x * (int)(blockDim.x >> 5) + (int)(threadIdx.x >> 5)) * 2 + (half >> 4);
    const int rows_r = B * Dr, rows = rows_r + B * Dt;
    const bool live = row < rows;                       // (a dead half-warp walks along: the shuffles are warp-wide)
    const bool is_r = row < rows_r;
    const int k = is_r ? row : row - rows_r, D = is_r ? Dr : Dt;
    const float *x = live ? (is_r ? xr + (long long)(k / D) * sr : xt + (long long)(k / D) * st_) + (long long)(k % D) * S : nullptr;
    const float ninf = __int_as_float(0xff800000);
    const float v = (live && l < S) ? __ldg(x + l) : ninf;
    float m = v;
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(kFull, m, o));
    const float mm = (fabsf(m) == __int_as_float(0x7f800000)) ? 0.f : m;
    const float e = (l < S) ? expf(__fsub_rn(v, m)) : 0.f;
    float t = e;
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) t = __fadd_rn(t, __shfl_xor_sync(kFull, t, o));
    const float lse = __fadd_rn(logf(t), mm);
    const float n = (l < S) ? __fsub_rn(v, lse) : ninf;
    float M = n;
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) M = fmaxf(M, __shfl_xor_sync(kFull, M, o));
    const float E = (l < S) ? expf(__fsub_rn(n, M)) : 0.f;
    float sum = E;
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) sum = __fadd_rn(sum, __shfl_xor_sync(kFull, sum, o));
    const float p = __fdiv_rn(E, sum);
    float best = (l < S) ? p : ninf;
    int arg = l;
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) {
        const float ob = __shfl_xor_sync(kFull, best, o);
        const int oa = __shfl_xor_sync(kFull, arg, o);
        if (ob > best || (ob == best && oa < arg)) best = ob, arg = oa;
    }
    if (!live) return;
    float *probs = is_r ? pr : pt;
    if (probs && l < S) probs[(long long)k * S + l] = p;
    if (l == 0) (is_r ? ar : at)[k] = arg;
}

// [B][C][P] -> [B][P][C] (NCHW -> torch's channels_last), P = H*W: what the head's first convolution wants of the
// observation.  A CTA turns a [32 channels][128 pixels] tile round through shared memory: 16-byte loads along the
// pixels, 16-byte stores along the channels (a pixel's 32 channels = one 128-byte line); the tile's row stride of 129
// words keeps both the scalar tile writes and the transposed reads free of bank conflicts.  Requires C % 4 == 0 and
// P % 4 == 0 (the launcher falls back to k_image_transpose otherwise).
__global__ void __launch_bounds__(256) k_to_channels_last(const float *__restrict__ x, int C, int P, float *__restrict__ y) {
    __shared__ float t[32][129];
    pdl_launch_dependents();
    pdl_wait();
    const int b = blockIdx.z, p0 = blockIdx.x * 128, c0 = blockIdx.y * 32;
    const float *src = x + (size_t)b * C * P;
    float *dst = y + (size_t)b * C * P;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int i = threadIdx.x + 256 * k, r = i >> 5, q = i & 31;   // channel row r, pixel quad q
        if (c0 + r < C && p0 + 4 * q < P) {
            const float4 v = __ldg(reinterpret_cast<const float4 *>(src + (size_t)(c0 + r) * P + p0) + q);
            t[r][4 * q + 0] = v.x, t[r][4 * q + 1] = v.y, t[r][4 * q + 2] = v.z, t[r][4 * q + 3] = v.w;
        }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int i = threadIdx.x + 256 * k, p = i >> 3, cq = i & 7;   // pixel p, channel quad cq
        if (p0 + p < P && c0 + 4 * cq < C)
            *reinterpret_cast<float4 *>(dst + (size_t)(p0 + p) * C + c0 + 4 * cq) =
                make_float4(t[4 * cq + 0][p], t[4 * cq + 1][p], t[4 * cq + 2][p], t[4 * cq + 3][p]);
    }
}

// AvgPool2d((H, W)): a CTA = 128 channels of one episode.  The additions of a channel must run in pixel order (torch's
// avg_pool2d adds the window row by row), so one thread owns a channel's sum - but the LOADS need not wait for it: four
// threads per channel bring 64 pixels at a time into shared memory (16 independent loads each, already activated), then
// the channel's thread adds the 64 values in order.  One DRAM round trip per 64 pixels instead of one per 13.
__global__ void __launch_bounds__(512) k_conv_epilogue_global_nhwc(const float *__restrict__ x, const float *__restrict__ scale,
                                                                   const float *__restrict__ shift, float slope, int B, int HW,
                                                                   int C, float *__restrict__ y) {
    __shared__ float t[64][128];
    pdl_launch_dependents();
    pdl_wait();
    const int b = blockIdx.y, c = (int)blockIdx.x * 128 + (int)(threadIdx.x & 127), q = threadIdx.x >> 7;
    const bool live = c < C;
    const float sc = live ? __ldg(scale + c) : 0.f, sh = live ? __ldg(shift + c) : 0.f;
    const float *p = x + (size_t)b * HW * C + (live ? c : 0);
    float s = 0.f;
    for (int k0 = 0; k0 < HW; k0 += 64) {
        float v[16];
#pragma unroll
        for (int jj = 0; jj < 16; ++jj) {
            const int k = k0 + q + 4 * jj;
            v[jj] = (live && k < HW) ? p[(size_t)k * C] : 0.f;
        }
#pragma unroll
        for (int jj = 0; jj < 16; ++jj) t[q + 4 * jj][threadIdx.x & 127] = lrelu_affine(v[jj], sc, sh, slope);
        __syncthreads();
        if (q == 0) {
            const int n = min(64, HW - k0);
            for (int k = 0; k < n; ++k) s = __fadd_rn(s, t[k][threadIdx.x]);
        }
        __syncthreads();
    }
    if (q == 0 && live) y[(size_t)b * C + c] = __fdiv_rn(s, (float)HW);
}

}  // namespace cmr
