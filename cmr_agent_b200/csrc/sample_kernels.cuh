// sample_kernels.cuh - image features sampled bilinearly at the points' projections ("bilinearly sample image features
// onto visible points", BASELINE.json north_star; SURVEY.md D1: an EXTRA operator - the reference has no point-side
// gather, its observation is the reverse scatter of environment/environment.py:67-83 - whose oracle is
// F.grid_sample(mode="bilinear", padding_mode="zeros", align_corners=True) on the projected pixel coordinates).
//
//   feats[b, c, n] = in_cam[b, n] * sum over the four pixels around (u, v) of w * img_geo_feat[b, c, y, x]
//   (u, v, in_cam) = the projection of environment.py:54-65 / :91-101 (project_uv below = project_point without the
//   rounding); x0 = floor(u), dx = u - x0, weights (1-dx)(1-dy), dx(1-dy), (1-dx)dy, dx dy; a neighbour outside the
//   grid counts as zero.  The four products are summed left to right, each operation rounded (no FMA contraction), so
//   the result equals the CPU restatement (oracle/sample_oracle.py) bit for bit.
//
// Layout.  The feature map arrives channel-first [C][H*W]; a point's four neighbours would be four 4-byte loads per
// channel.  k_image_transpose turns it round once per batch of images (it does not depend on the pose): [H*W][C], a
// pixel = one 256-byte row.  k_bilinear_sample: one warp per 32 consecutive points; for each point the warp loads the
// four rows (two channels per lane: 256-byte coalesced loads), blends them with the point's weights (warp-uniform, by
// shuffle) and puts the result into a point-major tile [32 points][64 channels] (XOR-swizzled as in
// cost_volume_kernels.cuh); the tile leaves channel-major - feats [B][C][N], the layout of pc_geo_feat - as full
// 128-byte lines.
#pragma once
#include "common.cuh"
#include "cost_volume_kernels.cuh"
#include "env_kernels.cuh"

namespace cmr {

// [B][C][P] -> [B][P][C], 32 x 32 tiles through shared memory
__global__ void __launch_bounds__(256) k_image_transpose(const float *__restrict__ img, int C, int P, float *__restrict__ imgT) {
    __shared__ float t[32][33];
    const int b = blockIdx.z, p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const float *src = img + (size_t)b * C * P;
    float *dst = imgT + (size_t)b * C * P;
#pragma unroll
    for (int r = ty; r < 32; r += 8)
        if (c0 + r < C && p0 + tx < P) t[r][tx] = __ldg(src + (size_t)(c0 + r) * P + p0 + tx);
    __syncthreads();
#pragma unroll
    for (int r = ty; r < 32; r += 8)
        if (p0 + r < P && c0 + tx < C) dst[(size_t)(p0 + r) * C + c0 + tx] = t[tx][r];
}

// environment.py:54-65 for one point: pixel coordinates before rounding, and the frustum test
template <bool kChain>
__device__ __forceinline__ void project_uv(const PoseK &s, float x, float y, float z, float wmax, float hmax, float &u, float &v,
                                           bool &in_cam) {
    float cx = __fsub_rn(x, s.m[0]), cy = __fsub_rn(y, s.m[1]), cz = __fsub_rn(z, s.m[2]);
    float X0 = __fadd_rn(__fadd_rn(dot3<kChain>(s.R[0], s.R[1], s.R[2], cx, cy, cz), s.m[0]), s.t[0]);
    float X1 = __fadd_rn(__fadd_rn(dot3<kChain>(s.R[3], s.R[4], s.R[5], cx, cy, cz), s.m[1]), s.t[1]);
    float X2 = __fadd_rn(__fadd_rn(dot3<kChain>(s.R[6], s.R[7], s.R[8], cx, cy, cz), s.m[2]), s.t[2]);
    float U0 = dot3<kChain>(s.K[0], s.K[1], s.K[2], X0, X1, X2);
    float U1 = dot3<kChain>(s.K[3], s.K[4], s.K[5], X0, X1, X2);
    float U2 = dot3<kChain>(s.K[6], s.K[7], s.K[8], X0, X1, X2);
    u = __fdiv_rn(U0, U2);
    v = __fdiv_rn(U1, U2);
    in_cam = (u >= 0.f) && (u <= wmax) && (v >= 0.f) && (v <= hmax) && (U2 > 0.f);
}

constexpr int kSmpThreads = 256;
constexpr int kSmpWarps = kSmpThreads / 32;
constexpr size_t kSmpSmem = (size_t)kSmpWarps * 32 * 64 * sizeof(float);   // a [32 points][64 channels] tile per warp

__global__ void __launch_bounds__(kSmpThreads, 3)
    k_bilinear_sample(const float *__restrict__ pc, const float *__restrict__ Kmat, const float *__restrict__ pose,
                      const float *__restrict__ mean, const float *__restrict__ imgT, int N, int C, int H, int W, bool vec,
                      float *__restrict__ out, uint8_t *__restrict__ in_cam_out) {
    extern __shared__ __align__(1024) unsigned char smp_smem[];
    const int b = blockIdx.y;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int n0 = (blockIdx.x * kSmpWarps + warp) * 32;
    if (n0 >= N) return;
    float *pm = reinterpret_cast<float *>(smp_smem) + (size_t)warp * 32 * 64;
    PoseK s;
    load_posek(s, pose, Kmat, mean, b, b);
    const float wmax = (float)(W - 1), hmax = (float)(H - 1);
    const int n = n0 + lane;
    float u = 0.f, v = 0.f;
    bool in_cam = false;
    if (n < N) {
        const float *px = pc + (size_t)b * 3 * N;
        const float x = __ldg(px + n), y = __ldg(px + N + n), z = __ldg(px + 2 * (size_t)N + n);
        if (N >= kBmmChainMinCols) project_uv<true>(s, x, y, z, wmax, hmax, u, v, in_cam);
        else project_uv<false>(s, x, y, z, wmax, hmax, u, v, in_cam);
        if (in_cam_out) in_cam_out[(size_t)b * N + n] = in_cam ? 1 : 0;
    }
    const unsigned cam = __ballot_sync(kFull, in_cam);
    // this lane's point: the pixel above-left of it, the fractions, which neighbours exist
    const float fx = floorf(u), fy = floorf(v);
    const int x0 = (int)fx, y0 = (int)fy;
    const float dx = __fsub_rn(u, fx), dy = __fsub_rn(v, fy);
    const float ex = __fsub_rn(1.f, dx), ey = __fsub_rn(1.f, dy);
    const float w00 = __fmul_rn(ex, ey), w01 = __fmul_rn(dx, ey), w10 = __fmul_rn(ex, dy), w11 = __fmul_rn(dx, dy);
    const int right = (x0 + 1 < W) ? 1 : 0, below = (y0 + 1 < H) ? 1 : 0;
    const int pix = in_cam ? y0 * W + x0 : 0;
    const float *imgb = imgT + (size_t)b * H * W * C;
    float *outb = out + (size_t)b * C * N;
    const int q4 = (lane & 7) * 4, g4 = lane >> 3;
    const bool whole = vec && n0 + 32 <= N;   // full 16-byte stores
    for (int c0 = 0; c0 < C; c0 += 64) {
        const bool mine = c0 + 2 * lane < C;   // this lane's channel pair exists
        const float *col = imgb + c0 + 2 * lane;
        // only the points inside the frustum are sampled (a fifth of a KITTI cloud at the ground-truth pose), four of them
        // in flight; the tile rows of the others are never written - the output pass writes their zeros from the mask
        for (unsigned todo = cam; todo;) {   // warp-uniform
            int pp[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                pp[i] = todo ? __ffs(todo) - 1 : -1;
                todo &= todo - 1;   // 0 stays 0
            }
            float2 a[4], bq[4], cq[4], dq[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int src = pp[i] & 31;
                const int ppix = __shfl_sync(kFull, pix, src), r = __shfl_sync(kFull, right, src), bl = __shfl_sync(kFull, below, src);
                const bool on = pp[i] >= 0 && mine;   // (warp-uniform apart from `mine`)
                const float *q = col + (size_t)ppix * C;
                const float2 zero = make_float2(0.f, 0.f);
                a[i] = on ? ldg_f2(q) : zero;
                bq[i] = (on && r) ? ldg_f2(q + C) : zero;
                cq[i] = (on && bl) ? ldg_f2(q + (size_t)W * C) : zero;
                dq[i] = (on && r && bl) ? ldg_f2(q + (size_t)W * C + C) : zero;
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int src = pp[i] & 31;
                const float k00 = __shfl_sync(kFull, w00, src), k01 = __shfl_sync(kFull, w01, src);
                const float k10 = __shfl_sync(kFull, w10, src), k11 = __shfl_sync(kFull, w11, src);
                if (pp[i] >= 0) {   // warp-uniform
                    float2 m;
                    m.x = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(a[i].x, k00), __fmul_rn(bq[i].x, k01)), __fmul_rn(cq[i].x, k10)),
                                    __fmul_rn(dq[i].x, k11));
                    m.y = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(a[i].y, k00), __fmul_rn(bq[i].y, k01)), __fmul_rn(cq[i].y, k10)),
                                    __fmul_rn(dq[i].y, k11));
                    const int cl = cv_col(2 * lane, pp[i]);   // an odd point: the pair's columns swap places
                    *reinterpret_cast<float2 *>(pm + pp[i] * 64 + (cl & ~1)) = (pp[i] & 1) ? make_float2(m.y, m.x) : m;
                }
            }
        }
        __syncwarp();
        // channel-major rows of 32 points
        if (whole && c0 + 64 <= C) {
            const unsigned a0 = smem_u32(pm) + (unsigned)(q4 * 64 * 4) + (unsigned)(q4 << 2);
            const unsigned t0 = a0 + 0 * 256 + ((g4 ^ 0) << 2), t1 = a0 + 1 * 256 + ((g4 ^ 1) << 2);
            const unsigned t2 = a0 + 2 * 256 + ((g4 ^ 2) << 2), t3 = a0 + 3 * 256 + ((g4 ^ 3) << 2);
            float *d = outb + (size_t)(c0 + g4) * N + n0 + q4;
            const size_t step = 4 * (size_t)N;
            const unsigned o4 = cam >> q4 & 15u;   // which of this lane's four points are inside the frustum
            if (cam) {
#pragma unroll
                for (int it = 0; it < 16; ++it) {
                    float4 val;
                    val.x = (o4 & 1u) ? lds_f32(t0 ^ (it << 4)) : 0.f;
                    val.y = (o4 & 2u) ? lds_f32(t1 ^ (it << 4)) : 0.f;
                    val.z = (o4 & 4u) ? lds_f32(t2 ^ (it << 4)) : 0.f;
                    val.w = (o4 & 8u) ? lds_f32(t3 ^ (it << 4)) : 0.f;
                    *reinterpret_cast<float4 *>(d) = val;
                    d += step;
                }
            } else {
#pragma unroll
                for (int it = 0; it < 16; ++it) {
                    *reinterpret_cast<float4 *>(d) = make_float4(0.f, 0.f, 0.f, 0.f);
                    d += step;
                }
            }
        } else {   // the cloud's last points, an odd N, a partial slab: lane = point, one channel per step
            for (int c = 0; c < 64 && c0 + c < C; ++c)
                if (n < N) outb[(size_t)(c0 + c) * N + n] = in_cam ? pm[lane * 64 + cv_col(c, lane)] : 0.f;
        }
        __syncwarp();
    }
}

}  // namespace cmr
