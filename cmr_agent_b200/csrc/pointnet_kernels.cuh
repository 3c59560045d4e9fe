// pointnet_kernels.cuh - sm_100a kernels for the PointNet++ front-end
// (reference: models/pointnet_util.py; every kernel cites the lines it replaces).
#pragma once
#include <cooperative_groups.h>

#include "common.cuh"

namespace cmr {
namespace cg = cooperative_groups;

// -------------------------------------------------------------------------------------------------
// square_distance (pointnet_util.py:19-33): out[b,s,n] = (dx*dx + dy*dy) + dz*dz, written once,
// coalesced along n.  Inputs may be strided views (PointNN.py:213 passes a permuted tensor).
__global__ void __launch_bounds__(256) k_square_distance(const float *__restrict__ src, int64_t ssb, int64_t ssn,
                                                          int64_t ssc, const float *__restrict__ dst, int64_t dsb,
                                                          int64_t dsn, int64_t dsc, int S, int N,
                                                          float *__restrict__ out) {
    const int b = blockIdx.z, s = blockIdx.y;
    const float *q = src + b * ssb + s * ssn;
    const float qx = __ldg(q), qy = __ldg(q + ssc), qz = __ldg(q + 2 * ssc);
    const float *r = dst + b * dsb;
    float *o = out + ((size_t)b * S + s) * N;
    for (int n = blockIdx.x * blockDim.x + threadIdx.x; n < N; n += gridDim.x * blockDim.x) {
        const float *p = r + n * dsn;
        o[n] = sqdist3(qx, qy, qz, __ldg(p), __ldg(p + dsc), __ldg(p + 2 * dsc));
    }
}

// -------------------------------------------------------------------------------------------------
// index_points (pointnet_util.py:36-47): out[b,s,:] = points[b, idx[b,s], :], rows copied in units
// of VecT.  Out-of-range indices write zeros and raise the sticky fault flag.
template <typename VecT>
__global__ void __launch_bounds__(256) k_index_points(const VecT *__restrict__ points, const int64_t *__restrict__ idx,
                                                       int N, int S, int row_vecs, VecT *__restrict__ out,
                                                       long long total) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        long long row = i / row_vecs;
        int v = (int)(i - row * row_vecs);
        int b = (int)(row / S);
        int64_t id = __ldg(idx + row);
        VecT val{};
        if (id < 0) id += N;   // torch.gather does not wrap, but advanced indexing does; be lenient once
        if (id >= 0 && id < N)
            val = __ldg(points + ((size_t)b * N + id) * row_vecs + v);
        else
            atomicExch(&g_fault, 1);
        out[i] = val;
    }
}

// backward of index_points: grad_points[b, idx[b,s], c] += grad_out[b,s,c]
__global__ void __launch_bounds__(256) k_index_points_bwd(const float *__restrict__ go, const int64_t *__restrict__ idx,
                                                           int N, int S, int C, float *__restrict__ gp,
                                                           long long total) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        long long row = i / C;
        int c = (int)(i - row * C);
        int b = (int)(row / S);
        int64_t id = __ldg(idx + row);
        if (id < 0) id += N;
        if (id >= 0 && id < N) atomicAdd(gp + ((size_t)b * N + id) * C + c, go[i]);
    }
}

// grouping tail of sample_and_group (pointnet_util.py:120-129):
// out[b,s,k,:] = cat(xyz[b,idx] - new_xyz[b,s], points[b,idx])
__global__ void __launch_bounds__(256) k_group_points(const float *__restrict__ xyz, const float *__restrict__ points,
                                                       const float *__restrict__ new_xyz,
                                                       const int64_t *__restrict__ idx, int N, int S, int K, int D,
                                                       float *__restrict__ out, long long total) {
    const int Cw = 3 + D;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        long long row = i / Cw;   // (b, s, k)
        int c = (int)(i - row * Cw);
        long long bs = row / K;   // (b, s)
        int b = (int)(bs / S);
        int64_t id = __ldg(idx + row);
        if (id < 0) id += N;
        float v = 0.f;
        if (id >= 0 && id < N) {
            if (c < 3)
                v = __fsub_rn(__ldg(xyz + ((size_t)b * N + id) * 3 + c), __ldg(new_xyz + bs * 3 + c));
            else
                v = __ldg(points + ((size_t)b * N + id) * D + (c - 3));
        } else {
            atomicExch(&g_fault, 1);
        }
        out[i] = v;
    }
}

// -------------------------------------------------------------------------------------------------
// farthest_point_sample (pointnet_util.py:50-70).
// One thread-block CLUSTER per cloud.  Every thread keeps PPT points (x, y, z, running min distance)
// in registers for the whole kernel - the cloud is read from HBM once - and the npoint sequential
// rounds cost one __syncthreads + one cluster barrier each:
//   thread : update PPT distances, first-max (blocked index layout => lowest index)
//   warp   : redux.sync max on the distance bits, redux.sync min on the index among the maxima
//   CTA    : 16 warp records in shared memory, every warp reduces them redundantly
//   cluster: each CTA pushes {dist, index, x, y, z} of its winner into every peer's shared memory
//            (DSMEM), one barrier.cluster, every thread picks the winner of <= 16 records.
// The coordinates of a CTA's winner come from a shared-memory copy of the CTA's slice (register
// arrays cannot be indexed dynamically).  Ties resolve to the lowest index like torch.max (:69).
struct alignas(16) FpsRec {
    unsigned bits;   // distance as ordered unsigned (distances are >= 0)
    unsigned idx;
    float x, y, z;
    unsigned pad[3];
};

template <int PPT, int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) k_fps(const float *__restrict__ xyz, const int64_t *__restrict__ start,
                                                     int N, int npoint, int64_t *__restrict__ out) {
    cg::cluster_group cluster = cg::this_cluster();
    const unsigned cs = cluster.num_blocks();
    const unsigned rank = cluster.block_rank();
    const int b = blockIdx.x / cs;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NW = THREADS / 32;

    extern __shared__ float sxyz[];              // [THREADS*PPT][3] this CTA's slice, for winner look-up
    __shared__ uint2 wrec[2][NW];                // per-warp {bits, idx}, double buffered
    __shared__ FpsRec crec[2][16];               // per-CTA records of the whole cluster, double buffered

    const float *cloud = xyz + (size_t)b * N * 3;
    const int slice0 = rank * (THREADS * PPT);   // first global index of this CTA
    const int first = slice0 + tid * PPT;        // first global index of this thread (blocked layout)

    // stage the slice through shared memory (coalesced), then into registers
    for (int i = tid; i < THREADS * PPT * 3; i += THREADS) {
        int g = slice0 * 3 + i;
        sxyz[i] = g < N * 3 ? __ldg(cloud + g) : 0.f;
    }
    __syncthreads();
    float x[PPT], y[PPT], z[PPT], d[PPT];
#pragma unroll
    for (int k = 0; k < PPT; ++k) {
        x[k] = sxyz[(tid * PPT + k) * 3 + 0];
        y[k] = sxyz[(tid * PPT + k) * 3 + 1];
        z[k] = sxyz[(tid * PPT + k) * 3 + 2];
        d[k] = (first + k < N) ? 1e10f : 0.f;   // :61; padding never beats a real point (ties -> lowest index)
    }

    long long s0 = start[b];
    if (s0 < 0) s0 += N;
    unsigned cur = (unsigned)min(max(s0, 0ll), (long long)N - 1);
    float cx = __ldg(cloud + (size_t)cur * 3), cy = __ldg(cloud + (size_t)cur * 3 + 1),
          cz = __ldg(cloud + (size_t)cur * 3 + 2);
    int64_t *o = out + (size_t)b * npoint;

    for (int it = 0; it < npoint; ++it) {
        const int buf = it & 1;
        if (rank == 0 && tid == 0) o[it] = (int64_t)cur;                       // :65
        if (it == npoint - 1) break;
        float best = -1.f;
        int bk = 0;
#pragma unroll
        for (int k = 0; k < PPT; ++k) {
            float dist = sqdist3(x[k], y[k], z[k], cx, cy, cz);                // :67
            d[k] = fminf(d[k], dist);                                          // :68
            if (d[k] > best) {                                                 // :69 first maximum
                best = d[k];
                bk = k;
            }
        }
        unsigned bits = __float_as_uint(best);
        unsigned wmax = __reduce_max_sync(kFull, bits);
        unsigned widx = __reduce_min_sync(kFull, bits == wmax ? (unsigned)(first + bk) : 0xffffffffu);
        if (lane == 0) wrec[buf][warp] = make_uint2(wmax, widx);
        __syncthreads();
        uint2 r = lane < NW ? wrec[buf][lane] : make_uint2(0u, 0xffffffffu);
        unsigned cmax = __reduce_max_sync(kFull, r.x);
        unsigned cidx = __reduce_min_sync(kFull, r.x == cmax ? r.y : 0xffffffffu);
        if (cs == 1) {
            cur = cidx;
            const float *p = sxyz + (size_t)(cidx - slice0) * 3;
            cx = p[0]; cy = p[1]; cz = p[2];
        } else {
            if (warp == 0 && lane < (int)cs) {
                const float *p = sxyz + (size_t)(cidx - slice0) * 3;
                FpsRec rec;
                rec.bits = cmax; rec.idx = cidx; rec.x = p[0]; rec.y = p[1]; rec.z = p[2];
                rec.pad[0] = rec.pad[1] = rec.pad[2] = 0;
                FpsRec *peer = cluster.map_shared_rank(&crec[buf][rank], lane);
                *reinterpret_cast<uint4 *>(peer) = *reinterpret_cast<uint4 *>(&rec);
                *(reinterpret_cast<uint4 *>(peer) + 1) = *(reinterpret_cast<uint4 *>(&rec) + 1);
            }
            cluster.sync();
            unsigned gb = 0, gi = 0xffffffffu;
            float gx = 0.f, gy = 0.f, gz = 0.f;
            for (unsigned c = 0; c < cs; ++c) {
                FpsRec rc = crec[buf][c];
                if (rc.bits > gb || (rc.bits == gb && rc.idx < gi)) {
                    gb = rc.bits; gi = rc.idx; gx = rc.x; gy = rc.y; gz = rc.z;
                }
            }
            cur = gi; cx = gx; cy = gy; cz = gz;
        }
    }
    if (cs > 1) cluster.sync();   // no CTA may exit while a peer can still write its shared memory
}

// -------------------------------------------------------------------------------------------------
// Warp-wide bitonic sort of E keys per lane, ascending over position = lane*E + r.
template <int E>
__device__ __forceinline__ void warp_bitonic_sort(unsigned long long (&key)[E], int lane) {
#pragma unroll
    for (int k = 2; k <= 32 * E; k <<= 1) {
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) {
            if (j >= E) {
                const int lj = j / E;
#pragma unroll
                for (int r = 0; r < E; ++r) {
                    unsigned long long other = __shfl_xor_sync(kFull, key[r], lj);
                    const bool up = (((lane * E + r) & k) == 0);
                    const bool lower = ((lane & lj) == 0);
                    const bool take_min = (lower == up);
                    const bool other_smaller = other < key[r];
                    key[r] = (take_min == other_smaller) ? other : key[r];
                }
            } else {
#pragma unroll
                for (int r = 0; r < E; ++r) {
                    if ((r & j) == 0) {
                        const bool up = (((lane * E + r) & k) == 0);
                        unsigned long long a = key[r], c = key[r | j];
                        const bool sw = (a > c) == up;
                        key[r] = sw ? c : a;
                        key[r | j] = sw ? a : c;
                    }
                }
            }
        }
    }
}

// kNN = square_distance(query, ref).argsort()[:, :, :k] (pointnet_util.py:115-116, PointNN.py:215-216)
// in the stable order (distance, index), without materialising the [S,N] matrix.
//   CTA   : 8 warps x QPW queries of one cloud; reference points stream through a shared-memory
//           tile (SoA, padded with NaN so out-of-range points never pass a comparison)
//   warp  : each lane owns one reference point per step and evaluates it against the warp's QPW
//           queries; a point enters a query's candidate buffer when d < current k-th distance
//   flush : when a buffer could overflow, the k best so far and the <= 64 candidates are sorted
//           together by a register bitonic network on 64-bit keys (distance bits << 32 | index),
//           which is exactly the (distance, index) order; the k-th key becomes the new threshold.
// Merge a query's candidate buffer into its sorted k-best list (both in shared memory) and
// refresh the admission threshold.  Called by a whole warp.
template <int KCAP, int E>
__device__ __forceinline__ void knn_flush(unsigned long long *L, const unsigned long long *Bf, int &cnt, float &tau,
                                          int k, int lane) {
    const unsigned long long kInf = ~0ull;
    unsigned long long key[E];
#pragma unroll
    for (int r = 0; r < E; ++r) {
        int pos = lane * E + r;
        key[r] = pos < KCAP ? L[pos] : (pos - KCAP < cnt ? Bf[pos - KCAP] : kInf);
    }
    __syncwarp();
    warp_bitonic_sort<E>(key, lane);
#pragma unroll
    for (int r = 0; r < E; ++r) {
        int pos = lane * E + r;
        if (pos < KCAP) L[pos] = key[r];
    }
    __syncwarp();
    unsigned long long kth = L[k - 1];
    tau = kth == kInf ? __int_as_float(0x7f800000) : __uint_as_float((unsigned)(kth >> 32));
    cnt = 0;
}

constexpr int kKnnTile = 1024;
constexpr int kKnnBuf = 64;

template <int KCAP, int QPW>
__global__ void __launch_bounds__(256) k_knn(const float *__restrict__ query, const float *__restrict__ ref, int S,
                                              int N, int k, int64_t *__restrict__ out) {
    constexpr int T = (KCAP + kKnnBuf) <= 128 ? 128 : 256;   // keys sorted per flush
    constexpr int E = T / 32;
    __shared__ float tx[kKnnTile], ty[kKnnTile], tz[kKnnTile];
    __shared__ unsigned long long slist[8 * QPW][KCAP];
    __shared__ unsigned long long sbuf[8 * QPW][kKnnBuf];
    const int b = blockIdx.y;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int q0 = (blockIdx.x * 8 + warp) * QPW;
    const float *qb = query + (size_t)b * S * 3;
    const float *rb = ref + (size_t)b * N * 3;
    const unsigned long long kInf = ~0ull;

    float qx[QPW], qy[QPW], qz[QPW], tau[QPW];
    int cnt[QPW];
#pragma unroll
    for (int q = 0; q < QPW; ++q) {
        int s = min(q0 + q, S - 1);
        qx[q] = __ldg(qb + (size_t)s * 3);
        qy[q] = __ldg(qb + (size_t)s * 3 + 1);
        qz[q] = __ldg(qb + (size_t)s * 3 + 2);
        tau[q] = __int_as_float(0x7f800000);   // +inf
        cnt[q] = 0;
        for (int i = lane; i < KCAP; i += 32) slist[warp * QPW + q][i] = kInf;
    }
    __syncwarp();

    for (int base = 0; base < N; base += kKnnTile) {
        __syncthreads();
        for (int i = tid; i < kKnnTile; i += 256) {
            int j = base + i;
            bool ok = j < N;
            tx[i] = ok ? __ldg(rb + (size_t)j * 3) : __int_as_float(0x7fc00000);
            ty[i] = ok ? __ldg(rb + (size_t)j * 3 + 1) : 0.f;
            tz[i] = ok ? __ldg(rb + (size_t)j * 3 + 2) : 0.f;
        }
        __syncthreads();
        const int lim = min(kKnnTile, N - base);
        for (int c = 0; c < lim; c += 32) {
            const float rx = tx[c + lane], ry = ty[c + lane], rz = tz[c + lane];
            const unsigned ridx = (unsigned)(base + c + lane);
#pragma unroll
            for (int q = 0; q < QPW; ++q) {
                float dd = sqdist3(qx[q], qy[q], qz[q], rx, ry, rz);   // square_distance(new_xyz, xyz): src - dst
                bool pass = dd < tau[q];
                unsigned m = __ballot_sync(kFull, pass);
                if (m) {
                    if (pass)
                        sbuf[warp * QPW + q][cnt[q] + __popc(m & ((1u << lane) - 1))] =
                            ((unsigned long long)__float_as_uint(dd) << 32) | ridx;
                    cnt[q] += __popc(m);
                    __syncwarp();
                    if (cnt[q] > kKnnBuf - 32)
                        knn_flush<KCAP, E>(slist[warp * QPW + q], sbuf[warp * QPW + q], cnt[q], tau[q], k, lane);
                }
            }
        }
    }
#pragma unroll
    for (int q = 0; q < QPW; ++q) {
        if (cnt[q] > 0) knn_flush<KCAP, E>(slist[warp * QPW + q], sbuf[warp * QPW + q], cnt[q], tau[q], k, lane);
        if (q0 + q < S) {
            int64_t *o = out + ((size_t)b * S + q0 + q) * k;
            for (int i = lane; i < k; i += 32) o[i] = (int64_t)(unsigned)(slist[warp * QPW + q][i] & 0xffffffffull);
        }
    }
}

// -------------------------------------------------------------------------------------------------
// query_ball_point (pointnet_util.py:73-93): the first nsample indices (ascending) with
// !(d > r2), padded with the first hit, N everywhere when there is none.  A warp owns QPW queries; the
// reference points go through a shared-memory tile that the 8 * QPW queries of the CTA share (most queries
// of a KITTI-scale cloud find fewer than nsample neighbours and scan everything); ordered ballot append; the
// CTA stops as soon as all its queries are full.
template <int QPW>
__global__ void __launch_bounds__(256) k_ball_query(const float *__restrict__ query, const float *__restrict__ ref,
                                                     float r2, int nsample, int S, int N, int64_t *__restrict__ out) {
    __shared__ float tx[kKnnTile], ty[kKnnTile], tz[kKnnTile];
    const int b = blockIdx.y;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int q0 = (blockIdx.x * 8 + warp) * QPW;
    const float *qb = query + (size_t)b * S * 3;
    const float *rb = ref + (size_t)b * N * 3;
    float qx[QPW], qy[QPW], qz[QPW];
    int found[QPW];
    long long first[QPW];
#pragma unroll
    for (int q = 0; q < QPW; ++q) {
        const int s = min(q0 + q, S - 1);
        qx[q] = __ldg(qb + (size_t)s * 3);
        qy[q] = __ldg(qb + (size_t)s * 3 + 1);
        qz[q] = __ldg(qb + (size_t)s * 3 + 2);
        found[q] = q0 + q < S ? 0 : nsample;   // queries past the end count as full
        first[q] = N;
    }
    for (int base = 0; base < N; base += kKnnTile) {
        bool open = false;
#pragma unroll
        for (int q = 0; q < QPW; ++q) open |= found[q] < nsample;
        if (!__syncthreads_or(open)) break;   // (also the barrier before the tile is overwritten)
        for (int i = tid; i < kKnnTile; i += 256) {
            const int j = base + i;
            const bool ok = j < N;
            tx[i] = ok ? __ldg(rb + (size_t)j * 3) : 0.f;
            ty[i] = ok ? __ldg(rb + (size_t)j * 3 + 1) : 0.f;
            tz[i] = ok ? __ldg(rb + (size_t)j * 3 + 2) : 0.f;
        }
        __syncthreads();
        if (!open) continue;
        const int lim = min(kKnnTile, N - base);
        for (int c = 0; c < lim; c += 32) {
            const bool in = c + lane < lim;
            const float rx = tx[c + lane], ry = ty[c + lane], rz = tz[c + lane];
            const int j = base + c + lane;
#pragma unroll
            for (int q = 0; q < QPW; ++q) {
                if (found[q] >= nsample) continue;   // warp-uniform
                const float dd = sqdist3(qx[q], qy[q], qz[q], rx, ry, rz);
                const bool hit = in && !(dd > r2);                                  // :88
                const unsigned m = __ballot_sync(kFull, hit);
                if (m) {
                    if (found[q] == 0) first[q] = base + c + __ffs(m) - 1;
                    const int pos = found[q] + __popc(m & ((1u << lane) - 1));
                    if (hit && pos < nsample) out[((size_t)b * S + q0 + q) * nsample + pos] = j;
                    found[q] += __popc(m);
                }
            }
        }
    }
#pragma unroll
    for (int q = 0; q < QPW; ++q) {
        if (q0 + q >= S) continue;
        int64_t *o = out + ((size_t)b * S + q0 + q) * nsample;
        const int f = min(found[q], nsample);
        for (int i = f + lane; i < nsample; i += 32) o[i] = first[q];              // :90-92
    }
}

// -------------------------------------------------------------------------------------------------
// kNN with a uniform grid over the reference cloud (same result as k_knn, bit for bit: the same distance
// expression, the same (distance, index) order - only far fewer candidates are evaluated).
//   k_grid_setup  one CTA per cloud: bounding box, the two axes with the largest extent (a LiDAR cloud is a slab:
//                 the third axis is not binned), square cells sized for ~16 points each, at most kGridMaxCells.
//   k_grid_count / k_grid_scan / k_grid_fill   counting sort of the points by cell: sorted[cell-major] = (x, y, z, index).
//   k_knn_grid    one warp per query: rings of cells around the query's cell, nearest first; every ring is at most
//                 four contiguous spans of the sorted array.  The search stops when the k-th distance so far is
//                 STRICTLY below the squared distance to the unexplored region (minus a rounding margin): a point out
//                 there cannot enter the list, not even on a tie.  Candidates are admitted on d <= tau (points arrive
//                 in cell order, so a tie may carry a smaller index than the current k-th).
constexpr int kGridMaxCells = 4096;
constexpr int kGridTargetPerCell = 16;
struct KnnGrid {
    float o0, o1, inv0, inv1, h;
    int a0, a1, g0, g1, pad0, pad1, pad2;
};
struct KnnGridWs {
    size_t per_cloud, off_start, off_fill, off_sorted;
};
inline KnnGridWs knn_grid_ws(int N) {
    KnnGridWs w;
    w.off_start = 64;
    w.off_fill = w.off_start + sizeof(int) * (kGridMaxCells + 1 + 3);
    w.off_sorted = round_up(w.off_fill + sizeof(int) * kGridMaxCells, 256);
    w.per_cloud = round_up(w.off_sorted + sizeof(float4) * (size_t)N, 256);
    return w;
}
__device__ __forceinline__ int grid_cell_1d(float v, float o, float inv, int g) {
    const int c = (int)floorf(__fmul_rn(__fsub_rn(v, o), inv));
    return min(max(c, 0), g - 1);
}

__global__ void __launch_bounds__(256) k_grid_setup(const float *__restrict__ ref, int N, unsigned char *__restrict__ ws, size_t per_cloud,
                                                    size_t off_start, size_t off_fill) {
    const int b = blockIdx.x, tid = threadIdx.x;
    const float *rb = ref + (size_t)b * N * 3;
    float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int j = tid; j < N; j += 256) {
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const float v = __ldg(rb + (size_t)j * 3 + a);
            lo[a] = fminf(lo[a], v);
            hi[a] = fmaxf(hi[a], v);
        }
    }
    __shared__ float slo[3][8], shi[3][8];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo[a] = fminf(lo[a], __shfl_xor_sync(kFull, lo[a], o));
            hi[a] = fmaxf(hi[a], __shfl_xor_sync(kFull, hi[a], o));
        }
        if ((tid & 31) == 0) {
            slo[a][tid >> 5] = lo[a];
            shi[a][tid >> 5] = hi[a];
        }
    }
    __syncthreads();
    unsigned char *base = ws + (size_t)b * per_cloud;
    int *start = reinterpret_cast<int *>(base + off_start), *fill = reinterpret_cast<int *>(base + off_fill);
    for (int i = tid; i < kGridMaxCells + 1; i += 256) start[i] = 0;
    for (int i = tid; i < kGridMaxCells; i += 256) fill[i] = 0;
    if (tid == 0) {
        float e[3];
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            float l = slo[a][0], h = shi[a][0];
            for (int w = 1; w < 8; ++w) {
                l = fminf(l, slo[a][w]);
                h = fmaxf(h, shi[a][w]);
            }
            lo[a] = l;
            hi[a] = h;
            e[a] = h - l;
            if (!(e[a] >= 0.f) || !isfinite(e[a])) e[a] = 0.f;   // NaN / inf coordinates: everything lands in cell 0
        }
        // the two axes with the largest extent
        int a0 = 0, a1 = 1, drop = 2;
        if (e[0] <= e[1] && e[0] <= e[2]) drop = 0;
        else if (e[1] <= e[0] && e[1] <= e[2]) drop = 1;
        a0 = drop == 0 ? 1 : 0;
        a1 = drop == 2 ? 1 : 2;
        const float cells = fminf((float)kGridMaxCells, fmaxf(1.f, (float)N / kGridTargetPerCell));
        float h = sqrtf(fmaxf(e[a0] * e[a1], 1e-30f) / cells);
        if (!(h > 0.f) || !isfinite(h)) h = 1.f;
        h = fmaxf(h, fmaxf(e[a0], e[a1]) * 1e-4f);     // at most 10^4 cells along an axis before the product is capped
        int g0 = max(1, (int)ceilf(e[a0] / h)), g1 = max(1, (int)ceilf(e[a1] / h));
        while ((long long)g0 * g1 > kGridMaxCells) {
            h *= 1.1f;
            g0 = max(1, (int)ceilf(e[a0] / h));
            g1 = max(1, (int)ceilf(e[a1] / h));
        }
        KnnGrid gr;
        gr.o0 = isfinite(lo[a0]) ? lo[a0] : 0.f;
        gr.o1 = isfinite(lo[a1]) ? lo[a1] : 0.f;
        gr.h = h;
        gr.inv0 = gr.inv1 = 1.f / h;
        gr.a0 = a0; gr.a1 = a1; gr.g0 = g0; gr.g1 = g1;
        gr.pad0 = gr.pad1 = gr.pad2 = 0;
        *reinterpret_cast<KnnGrid *>(base) = gr;
    }
}

__global__ void __launch_bounds__(256) k_grid_count(const float *__restrict__ ref, int N, unsigned char *__restrict__ ws, size_t per_cloud,
                                                    size_t off_start) {
    const int b = blockIdx.y, j = blockIdx.x * 256 + threadIdx.x;
    if (j >= N) return;
    unsigned char *base = ws + (size_t)b * per_cloud;
    const KnnGrid gr = *reinterpret_cast<const KnnGrid *>(base);
    const float *p = ref + ((size_t)b * N + j) * 3;
    const int c = grid_cell_1d(__ldg(p + gr.a1), gr.o1, gr.inv1, gr.g1) * gr.g0 + grid_cell_1d(__ldg(p + gr.a0), gr.o0, gr.inv0, gr.g0);
    atomicAdd(reinterpret_cast<int *>(base + off_start) + c, 1);
}

// counts -> exclusive prefix (start[cells] = N), one CTA of 1024 threads per cloud
__global__ void __launch_bounds__(1024) k_grid_scan(unsigned char *__restrict__ ws, size_t per_cloud, size_t off_start) {
    unsigned char *base = ws + (size_t)blockIdx.x * per_cloud;
    int *start = reinterpret_cast<int *>(base + off_start);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int v[4], sum = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        v[i] = start[tid * 4 + i];
        sum += v[i];
    }
    int inc = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(kFull, inc, o);
        if (lane >= o) inc += t;
    }
    __shared__ int wtot[32];
    if (lane == 31) wtot[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        int w = wtot[lane], winc = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(kFull, winc, o);
            if (lane >= o) winc += t;
        }
        wtot[lane] = winc - w;
    }
    __syncthreads();
    int run = wtot[warp] + inc - sum;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        start[tid * 4 + i] = run;
        run += v[i];
    }
    if (tid == 1023) start[kGridMaxCells] = run;
}

__global__ void __launch_bounds__(256) k_grid_fill(const float *__restrict__ ref, int N, unsigned char *__restrict__ ws, size_t per_cloud,
                                                   size_t off_start, size_t off_fill, size_t off_sorted) {
    const int b = blockIdx.y, j = blockIdx.x * 256 + threadIdx.x;
    if (j >= N) return;
    unsigned char *base = ws + (size_t)b * per_cloud;
    const KnnGrid gr = *reinterpret_cast<const KnnGrid *>(base);
    const float *p = ref + ((size_t)b * N + j) * 3;
    const float x = __ldg(p), y = __ldg(p + 1), z = __ldg(p + 2);
    const float v0 = gr.a0 == 0 ? x : (gr.a0 == 1 ? y : z), v1 = gr.a1 == 1 ? y : (gr.a1 == 2 ? z : x);
    const int c = grid_cell_1d(v1, gr.o1, gr.inv1, gr.g1) * gr.g0 + grid_cell_1d(v0, gr.o0, gr.inv0, gr.g0);
    const int pos = reinterpret_cast<const int *>(base + off_start)[c] + atomicAdd(reinterpret_cast<int *>(base + off_fill) + c, 1);
    reinterpret_cast<float4 *>(base + off_sorted)[pos] = make_float4(x, y, z, __int_as_float(j));
}

template <int KCAP>
__global__ void __launch_bounds__(256) k_knn_grid(const float *__restrict__ query, const unsigned char *__restrict__ ws, size_t per_cloud,
                                                  size_t off_start, size_t off_sorted, int S, int N, int k, int64_t *__restrict__ out) {
    constexpr int T = (KCAP + kKnnBuf) <= 128 ? 128 : 256;
    constexpr int E = T / 32;
    __shared__ unsigned long long slist[8][KCAP];
    __shared__ unsigned long long sbuf[8][kKnnBuf];
    const int b = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int s = blockIdx.x * 8 + warp;
    if (s >= S) return;
    const unsigned char *base = ws + (size_t)b * per_cloud;
    const KnnGrid gr = *reinterpret_cast<const KnnGrid *>(base);
    const int *start = reinterpret_cast<const int *>(base + off_start);
    const float4 *sorted = reinterpret_cast<const float4 *>(base + off_sorted);
    const float *qp = query + ((size_t)b * S + s) * 3;
    const float qx = __ldg(qp), qy = __ldg(qp + 1), qz = __ldg(qp + 2);
    const float q0 = gr.a0 == 0 ? qx : (gr.a0 == 1 ? qy : qz), q1 = gr.a1 == 1 ? qy : (gr.a1 == 2 ? qz : qx);
    const int c0 = grid_cell_1d(q0, gr.o0, gr.inv0, gr.g0), c1 = grid_cell_1d(q1, gr.o1, gr.inv1, gr.g1);
    unsigned long long *L = slist[warp], *Bf = sbuf[warp];
    const unsigned long long kInf = ~0ull;
    for (int i = lane; i < KCAP; i += 32) L[i] = kInf;
    __syncwarp();
    float tau = __int_as_float(0x7f800000);
    int cnt = 0;
    auto span = [&](int cell_lo, int cell_hi) {   // cells [cell_lo, cell_hi] of one grid row: contiguous in `sorted`
        const int beg = start[cell_lo], end = start[cell_hi + 1];
        for (int i0 = beg; i0 < end; i0 += 32) {
            const int i = i0 + lane;
            bool pass = false;
            float dd = 0.f;
            unsigned ridx = 0;
            if (i < end) {
                const float4 p = __ldg(sorted + i);
                dd = sqdist3(qx, qy, qz, p.x, p.y, p.z);   // square_distance(new_xyz, xyz): src - dst, as k_knn
                ridx = (unsigned)__float_as_int(p.w);
                pass = dd <= tau;
            }
            const unsigned m = __ballot_sync(kFull, pass);
            if (m) {
                if (pass) Bf[cnt + __popc(m & ((1u << lane) - 1))] = ((unsigned long long)__float_as_uint(dd) << 32) | ridx;
                cnt += __popc(m);
                __syncwarp();
                if (cnt > kKnnBuf - 32) knn_flush<KCAP, E>(L, Bf, cnt, tau, k, lane);
            }
        }
    };
    const int rmax = max(max(c0, gr.g0 - 1 - c0), max(c1, gr.g1 - 1 - c1));
    for (int r = 0; r <= rmax; ++r) {
        const int lo0 = c0 - r, hi0 = c0 + r, lo1 = c1 - r, hi1 = c1 + r;
        const int x0 = max(lo0, 0), x1 = min(hi0, gr.g0 - 1);
        for (int y = max(lo1, 0); y <= min(hi1, gr.g1 - 1); ++y) {
            if (y == lo1 || y == hi1) {
                span(y * gr.g0 + x0, y * gr.g0 + x1);
            } else {
                if (lo0 >= 0) span(y * gr.g0 + lo0, y * gr.g0 + lo0);
                if (hi0 < gr.g0) span(y * gr.g0 + hi0, y * gr.g0 + hi0);
            }
        }
        if (cnt > 0) knn_flush<KCAP, E>(L, Bf, cnt, tau, k, lane);
        // distance from the query to the nearest face of the explored block that still has cells behind it
        float lb = __int_as_float(0x7f800000);
        if (lo0 > 0) lb = fminf(lb, q0 - (gr.o0 + lo0 * gr.h));
        if (hi0 < gr.g0 - 1) lb = fminf(lb, (gr.o0 + (hi0 + 1) * gr.h) - q0);
        if (lo1 > 0) lb = fminf(lb, q1 - (gr.o1 + lo1 * gr.h));
        if (hi1 < gr.g1 - 1) lb = fminf(lb, (gr.o1 + (hi1 + 1) * gr.h) - q1);
        lb -= 1e-3f * gr.h;                            // cell assignment and the faces are rounded: stay on the safe side
        if (lb > 0.f && tau < lb * lb * (1.f - 1e-5f)) break;
    }
    int64_t *o = out + ((size_t)b * S + s) * k;
    for (int i = lane; i < k; i += 32) o[i] = (int64_t)(unsigned)(L[i] & 0xffffffffull);
}

// query_ball_point on the same grid: only the cells that the ball's bounding square touches are read.  The hits are
// the points with !(d > r2) exactly as in k_ball_query; of those the nsample SMALLEST INDICES are wanted, ascending
// (pointnet_util.py:86-92) - the k-best machinery of the kNN with the index as the key.
template <int KCAP>
__global__ void __launch_bounds__(256) k_ball_grid(const float *__restrict__ query, const unsigned char *__restrict__ ws, size_t per_cloud,
                                                   size_t off_start, size_t off_sorted, float r2, float radius, int nsample, int S, int N,
                                                   int64_t *__restrict__ out) {
    constexpr int T = (KCAP + kKnnBuf) <= 128 ? 128 : 256;
    constexpr int E = T / 32;
    __shared__ unsigned long long slist[8][KCAP];
    __shared__ unsigned long long sbuf[8][kKnnBuf];
    const int b = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int s = blockIdx.x * 8 + warp;
    if (s >= S) return;
    const unsigned char *base = ws + (size_t)b * per_cloud;
    const KnnGrid gr = *reinterpret_cast<const KnnGrid *>(base);
    const int *start = reinterpret_cast<const int *>(base + off_start);
    const float4 *sorted = reinterpret_cast<const float4 *>(base + off_sorted);
    const float *qp = query + ((size_t)b * S + s) * 3;
    const float qx = __ldg(qp), qy = __ldg(qp + 1), qz = __ldg(qp + 2);
    const float q0 = gr.a0 == 0 ? qx : (gr.a0 == 1 ? qy : qz), q1 = gr.a1 == 1 ? qy : (gr.a1 == 2 ? qz : qx);
    unsigned long long *L = slist[warp], *Bf = sbuf[warp];
    const unsigned long long kInf = ~0ull;
    for (int i = lane; i < KCAP; i += 32) L[i] = kInf;
    __syncwarp();
    float tau = __int_as_float(0x7f800000);   // knn_flush's threshold (a float view of the key's high word): unused here
    unsigned long long kth = kInf;            // the nsample-th smallest index so far
    int cnt = 0;
    const float reach = radius * (1.f + 1e-5f) + 1e-3f * gr.h;   // cell assignment is rounded: stay on the safe side
    const bool any = r2 >= 0.f || r2 != r2;                         // a negative r2 has no hits (NaN: every !(d > r2) is true)
    if (any) {
        const int x0 = grid_cell_1d(q0 - reach, gr.o0, gr.inv0, gr.g0), x1 = grid_cell_1d(q0 + reach, gr.o0, gr.inv0, gr.g0);
        const int y0 = grid_cell_1d(q1 - reach, gr.o1, gr.inv1, gr.g1), y1 = grid_cell_1d(q1 + reach, gr.o1, gr.inv1, gr.g1);
        for (int y = y0; y <= y1; ++y) {
            const int beg = start[y * gr.g0 + x0], end = start[y * gr.g0 + x1 + 1];
            for (int i0 = beg; i0 < end; i0 += 32) {
                const int i = i0 + lane;
                bool pass = false;
                unsigned ridx = 0;
                if (i < end) {
                    const float4 p = __ldg(sorted + i);
                    const float dd = sqdist3(qx, qy, qz, p.x, p.y, p.z);
                    ridx = (unsigned)__float_as_int(p.w);
                    pass = !(dd > r2) && (unsigned long long)ridx <= kth;                  // :88
                }
                const unsigned m = __ballot_sync(kFull, pass);
                if (m) {
                    if (pass) Bf[cnt + __popc(m & ((1u << lane) - 1))] = (unsigned long long)ridx;
                    cnt += __popc(m);
                    __syncwarp();
                    if (cnt > kKnnBuf - 32) {
                        knn_flush<KCAP, E>(L, Bf, cnt, tau, nsample, lane);
                        kth = L[nsample - 1];
                    }
                }
            }
        }
        if (cnt > 0) knn_flush<KCAP, E>(L, Bf, cnt, tau, nsample, lane);
    }
    __syncwarp();
    const unsigned long long first = L[0];
    int64_t *o = out + ((size_t)b * S + s) * nsample;
    for (int i = lane; i < nsample; i += 32) {
        const unsigned long long v = L[i];
        o[i] = v != kInf ? (int64_t)v : (first != kInf ? (int64_t)first : (int64_t)N);       // :90-92
    }
}

// -------------------------------------------------------------------------------------------------
// farthest_point_sample on the grid (same indices as k_fps, bit for bit).  After s samples only the points within
// about one sample spacing of the newest centroid can still lower their distance; everything else is skipped CELL BY
// CELL: a cell whose bounding square lies farther from the centroid than the largest running distance inside the
// cell cannot change (d = min(d, dist) with dist >= d).  Work per round drops from N points to the centroid's
// neighbourhood (N ln(npoint) point updates in total instead of N npoint), and one CTA - one SM - carries a whole
// cloud, so 148 clouds run at once instead of 37 four-CTA clusters.
//   shared memory: the running distances of all points in cell order (4N bytes), per cell {max distance bits, lowest
//   original index attaining it}, the list of cells to update this round.
//   round: (1) every thread tests its cells against the centroid, (2) a warp per listed cell updates the cell's
//   points (coordinates from the cell-sorted copy in L2) and its record, (3) arg-max over the cell records.
constexpr int kFpsGridThreads = 1024;
struct FpsGridSmem {
    size_t off_d, off_cmax, off_cidx, off_list, off_cs, total;
};
inline FpsGridSmem fps_grid_smem(int N) {
    FpsGridSmem m;
    m.off_d = 0;
    m.off_cmax = round_up(sizeof(float) * (size_t)N, 16);
    m.off_cidx = m.off_cmax + sizeof(unsigned) * kGridMaxCells;
    m.off_list = m.off_cidx + sizeof(unsigned) * kGridMaxCells;
    m.off_cs = m.off_list + sizeof(unsigned short) * kGridMaxCells;        // cell start offsets, u16 pairs would not do: N > 65535 is allowed
    m.total = m.off_cs + sizeof(int) * (kGridMaxCells + 4);
    return m;
}

__global__ void __launch_bounds__(kFpsGridThreads, 1)
    k_fps_grid(const float *__restrict__ xyz, const int64_t *__restrict__ start, const unsigned char *__restrict__ ws, size_t per_cloud,
               size_t off_start, size_t off_sorted, int N, int npoint, size_t off_cmax, size_t off_cidx, size_t off_list,
               size_t off_cs, int64_t *__restrict__ out) {
    extern __shared__ __align__(16) unsigned char fps_smem[];
    float *sd = reinterpret_cast<float *>(fps_smem);
    unsigned *cmax = reinterpret_cast<unsigned *>(fps_smem + off_cmax);
    unsigned *cidx = reinterpret_cast<unsigned *>(fps_smem + off_cidx);
    unsigned short *list = reinterpret_cast<unsigned short *>(fps_smem + off_list);
    int *cs = reinterpret_cast<int *>(fps_smem + off_cs);                // the cells' start offsets, a copy in shared memory
    __shared__ int nlist;
    __shared__ uint2 wrec[32];
    __shared__ float cen[3];
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned char *base = ws + (size_t)b * per_cloud;
    const KnnGrid gr = *reinterpret_cast<const KnnGrid *>(base);
    const int *cstart = reinterpret_cast<const int *>(base + off_start);
    const float4 *sorted = reinterpret_cast<const float4 *>(base + off_sorted);
    const float *cloud = xyz + (size_t)b * N * 3;
    const int ncells = gr.g0 * gr.g1;
    for (int i = tid; i < N; i += kFpsGridThreads) sd[i] = 1e10f;                    // :61
    const unsigned kBig = __float_as_uint(1e10f);
    for (int c = tid; c <= ncells; c += kFpsGridThreads) cs[c] = cstart[c];
    // this thread's cells (at most kGridMaxCells / threads of them) and the low corners of their squares, for all rounds
    constexpr int kMyCells = kGridMaxCells / kFpsGridThreads;
    float mylo0[kMyCells], mylo1[kMyCells];
#pragma unroll
    for (int j = 0; j < kMyCells; ++j) {
        const int c = tid + j * kFpsGridThreads;
        mylo0[j] = gr.o0 + (c % gr.g0) * gr.h;
        mylo1[j] = gr.o1 + (c / gr.g0) * gr.h;
        if (c < ncells) {
            const int n = cstart[c + 1] - cstart[c];
            cmax[c] = n > 0 ? kBig : 0u;   // an empty cell never wins and never needs an update
            cidx[c] = 0xffffffffu;         // (set by the first update: every non-empty cell is updated in round 0)
        }
    }
    long long s0 = start[b];
    if (s0 < 0) s0 += N;
    unsigned cur = (unsigned)min(max(s0, 0ll), (long long)N - 1);
    if (tid == 0) {
        cen[0] = __ldg(cloud + (size_t)cur * 3);
        cen[1] = __ldg(cloud + (size_t)cur * 3 + 1);
        cen[2] = __ldg(cloud + (size_t)cur * 3 + 2);
        nlist = 0;
    }
    __syncthreads();
    int64_t *o = out + (size_t)b * npoint;
    const float slack = 1e-3f * gr.h;
    float cx = cen[0], cy = cen[1], cz = cen[2];      // the newest centroid, in every thread's registers
    for (int it = 0; it < npoint; ++it) {
        if (tid == 0) o[it] = (int64_t)cur;                                            // :65
        if (it == npoint - 1) break;
        const float c0 = gr.a0 == 0 ? cx : (gr.a0 == 1 ? cy : cz), c1 = gr.a1 == 1 ? cy : (gr.a1 == 2 ? cz : cx);
        // (1) which cells can still change
#pragma unroll
        for (int j = 0; j < kMyCells; ++j) {
            const int c = tid + j * kFpsGridThreads;
            bool active = false;
            if (c < ncells) {
                const unsigned m = cmax[c];
                if (m != 0u) {
                    const float lo0 = mylo0[j] - slack, hi0 = mylo0[j] + gr.h + slack;
                    const float lo1 = mylo1[j] - slack, hi1 = mylo1[j] + gr.h + slack;
                    const float dx = fmaxf(fmaxf(lo0 - c0, c0 - hi0), 0.f), dy = fmaxf(fmaxf(lo1 - c1, c1 - hi1), 0.f);
                    const float lb2 = (dx * dx + dy * dy) * (1.f - 1e-5f);
                    active = !(lb2 > __uint_as_float(m));
                }
            }
            const unsigned bal = __ballot_sync(kFull, active);   // (the loop is unrolled: all lanes are here)
            if (active) {
                // warp-aggregated append (lanes of a warp scan consecutive cells)
                const int leader = __ffs(bal) - 1;
                int basep = 0;
                if (lane == leader) basep = atomicAdd(&nlist, __popc(bal));
                basep = __shfl_sync(bal, basep, leader);
                list[basep + __popc(bal & ((1u << lane) - 1))] = (unsigned short)c;
            }
        }
        __syncthreads();
        // (2) a warp per listed cell
        const int nl = nlist;
        for (int li = warp; li < nl; li += kFpsGridThreads / 32) {
            const int c = list[li];
            const int beg = cs[c], end = cs[c + 1];
            unsigned bm = 0u, bi = 0xffffffffu;
            for (int i0 = beg; i0 < end; i0 += 32) {
                const int i = i0 + lane;
                unsigned bits = 0u, idx = 0xffffffffu;
                if (i < end) {
                    const float4 p = __ldg(sorted + i);
                    const float dist = sqdist3(p.x, p.y, p.z, cx, cy, cz);              // :67
                    const float d = fminf(sd[i], dist);                                 // :68
                    sd[i] = d;
                    bits = __float_as_uint(d);
                    idx = (unsigned)__float_as_int(p.w);
                }
                const unsigned wm = __reduce_max_sync(kFull, bits);
                const unsigned wi = __reduce_min_sync(kFull, bits == wm ? idx : 0xffffffffu);
                if (wm > bm || (wm == bm && wi < bi)) {
                    bm = wm;
                    bi = wi;
                }
            }
            if (lane == 0) {
                cmax[c] = bm;
                cidx[c] = bi;
            }
        }
        __syncthreads();
        // (3) arg-max over the cell records: the largest distance, the lowest original index among equals (:69)
        unsigned bm = 0u, bi = 0xffffffffu;
        for (int c = tid; c < ncells; c += kFpsGridThreads) {
            const unsigned m = cmax[c], ix = cidx[c];
            if (m > bm || (m == bm && ix < bi)) {
                bm = m;
                bi = ix;
            }
        }
        const unsigned wm = __reduce_max_sync(kFull, bm);
        const unsigned wi = __reduce_min_sync(kFull, bm == wm ? bi : 0xffffffffu);
        if (lane == 0) wrec[warp] = make_uint2(wm, wi);
        if (tid == 0) nlist = 0;           // (everybody is past step (2))
        __syncthreads();
        // every warp finishes the reduction itself: no second barrier, no broadcast through shared memory
        const uint2 r = wrec[lane];
        const unsigned gm = __reduce_max_sync(kFull, r.x);
        const unsigned gi = __reduce_min_sync(kFull, r.x == gm ? r.y : 0xffffffffu);
        const float cv = lane < 3 ? __ldg(cloud + (size_t)gi * 3 + lane) : 0.f;
        cx = __shfl_sync(kFull, cv, 0);
        cy = __shfl_sync(kFull, cv, 1);
        cz = __shfl_sync(kFull, cv, 2);
        cur = gi;
    }
}

}  // namespace cmr
