// dataset_kernels.cuh - the two geometric operations the reference's DATASETS run per sample on the host, in
// float64 numpy / scipy:
//   FarthestSampler.sample   dataset/KittiDataset.py:107-126 (= dataset/NuScenesDataset.py:25-44)
//   cKDTree(node).query(pc, k=1)   dataset/KittiDataset.py:359-367 (nearest node of every point)
// Both stay float64 here: the indices they produce are compared bit for bit with numpy's.
#pragma once
#include "common.cuh"

namespace cmr {

constexpr int kFps64Threads = 1024;
constexpr int kFps64MaxPpt = 16;   // points per thread: clouds of up to 16384 points per CTA

// ((p0 - p) ** 2).sum(axis=0) as numpy evaluates it for a [3, M] array: squares, then (r0 + r1) + r2, no FMA
__device__ __forceinline__ double sqdist3_f64(double ax, double ay, double az, double bx, double by, double bz) {
    const double dx = __dsub_rn(ax, bx), dy = __dsub_rn(ay, by), dz = __dsub_rn(az, bz);
    return __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
}

// FarthestSampler.sample (KittiDataset.py:115-126): one CTA per cloud.  pts [B,3,M] f64 channel-major (numpy's
// [3, M]); start [B] = init_idx of :118.  The running distances (np.minimum of :125) live in registers, the
// coordinates are re-read every round (they stay in L2; three f64 planes do not fit one SM).  np.argmax of :122
// returns the FIRST maximum: ties go to the lowest index.  out_idx [B,k] i64, out_pts [B,3,k] f64 (optional).
// kStage: x and y live in (dynamic) shared memory and z in registers for all rounds - nothing is re-read from L2;
// needs 16 M bytes of shared memory (M <= 12800 or so), otherwise the coordinates are re-read every round.
template <int PPT, bool kStage>
__global__ void __launch_bounds__(kFps64Threads) k_fps_f64(const double *__restrict__ pts, const int64_t *__restrict__ start,
                                                            int M, int k, int64_t *__restrict__ out_idx,
                                                            double *__restrict__ out_pts) {
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const double *px = pts + (size_t)b * 3 * M, *py = px + M, *pz = py + M;
    __shared__ double s_c[3];
    __shared__ double s_val[32];
    __shared__ int s_idx[32];
    __shared__ int s_win;
    extern __shared__ double s_xy[];   // kStage: [2][M]
    double dist[PPT], zreg[PPT];
    if (kStage) {
#pragma unroll
        for (int i = 0; i < PPT; ++i) {
            const int j = i * kFps64Threads + tid;
            zreg[i] = 0.0;
            if (j < M) {
                s_xy[j] = px[j];
                s_xy[M + j] = py[j];
                zreg[i] = pz[j];
            }
        }
    }
    int cur = (int)start[b];
    if (cur < 0 || cur >= M) {
        if (tid == 0) atomicExch(&g_fault, 1);
        cur = 0;
    }
    if (tid == 0) {
        s_c[0] = px[cur];
        s_c[1] = py[cur];
        s_c[2] = pz[cur];
        out_idx[(size_t)b * k] = cur;
        if (out_pts) {
            out_pts[((size_t)b * 3 + 0) * k] = s_c[0];
            out_pts[((size_t)b * 3 + 1) * k] = s_c[1];
            out_pts[((size_t)b * 3 + 2) * k] = s_c[2];
        }
    }
    __syncthreads();
    for (int r = 1; r < k; ++r) {
        const double cx = s_c[0], cy = s_c[1], cz = s_c[2];
        double best = -1.0, bx = 0.0, by = 0.0, bz = 0.0;
        int bi = 0x7fffffff;
#pragma unroll
        for (int i = 0; i < PPT; ++i) {
            const int j = i * kFps64Threads + tid;
            if (j < M) {
                const double x = kStage ? s_xy[j] : __ldg(px + j), y = kStage ? s_xy[M + j] : __ldg(py + j),
                             z = kStage ? zreg[i] : __ldg(pz + j);
                const double d = sqdist3_f64(cx, cy, cz, x, y, z);
                // round 1 initialises the distances (:120), later rounds take np.minimum (:125)
                dist[i] = (r == 1) ? d : fmin(dist[i], d);
                if (dist[i] > best) {   // ascending j inside a thread: the first maximum is kept
                    best = dist[i];
                    bi = j;
                    bx = x; by = y; bz = z;
                }
            }
        }
        // argmax over the CTA: larger distance wins, then the lower index
        double v = best;
        int vi = bi;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double ov = __shfl_xor_sync(kFull, v, o);
            const int oi = __shfl_xor_sync(kFull, vi, o);
            if (ov > v || (ov == v && oi < vi)) {
                v = ov;
                vi = oi;
            }
        }
        if (lane == 0) {
            s_val[warp] = v;
            s_idx[warp] = vi;
        }
        __syncthreads();
        if (warp == 0) {
            v = s_val[lane];
            vi = s_idx[lane];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const double ov = __shfl_xor_sync(kFull, v, o);
                const int oi = __shfl_xor_sync(kFull, vi, o);
                if (ov > v || (ov == v && oi < vi)) {
                    v = ov;
                    vi = oi;
                }
            }
            if (lane == 0) s_win = vi;
        }
        __syncthreads();
        const int win = s_win;
        if (win == bi) {   // the owner publishes the new centroid (it kept the coordinates of its best)
            s_c[0] = bx;
            s_c[1] = by;
            s_c[2] = bz;
            out_idx[(size_t)b * k + r] = win;
            if (out_pts) {
                out_pts[((size_t)b * 3 + 0) * k + r] = bx;
                out_pts[((size_t)b * 3 + 1) * k + r] = by;
                out_pts[((size_t)b * 3 + 2) * k + r] = bz;
            }
        }
        __syncthreads();
    }
}

// cKDTree(ref.T).query(query.T, k=1)[1] (KittiDataset.py:365-366): index of the nearest reference point of every
// query point, by brute force in float64 (squared Euclidean distance, lowest index on ties).  query [B,3,N],
// ref [B,3,S] channel-major; reference points go through shared memory 1024 at a time.
__global__ void __launch_bounds__(256) k_nearest_f64(const double *__restrict__ query, const double *__restrict__ ref, int N,
                                                      int S, int64_t *__restrict__ out) {
    constexpr int kChunk = 1024;
    __shared__ double s_ref[3][kChunk];
    const int b = blockIdx.y;
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const double *q = query + (size_t)b * 3 * N, *r = ref + (size_t)b * 3 * S;
    double qx = 0.0, qy = 0.0, qz = 0.0;
    if (j < N) {
        qx = q[j];
        qy = q[(size_t)N + j];
        qz = q[2 * (size_t)N + j];
    }
    double best = __longlong_as_double(0x7ff0000000000000LL);   // +inf
    int bi = 0;
    for (int s0 = 0; s0 < S; s0 += kChunk) {
        const int ns = min(kChunk, S - s0);
        __syncthreads();
        for (int i = threadIdx.x; i < ns; i += blockDim.x) {
            s_ref[0][i] = r[s0 + i];
            s_ref[1][i] = r[(size_t)S + s0 + i];
            s_ref[2][i] = r[2 * (size_t)S + s0 + i];
        }
        __syncthreads();
#pragma unroll 4
        for (int i = 0; i < ns; ++i) {
            const double d = sqdist3_f64(qx, qy, qz, s_ref[0][i], s_ref[1][i], s_ref[2][i]);
            if (d < best) {
                best = d;
                bi = s0 + i;
            }
        }
    }
    if (j < N) out[(size_t)b * N + j] = bi;
}

}  // namespace cmr
