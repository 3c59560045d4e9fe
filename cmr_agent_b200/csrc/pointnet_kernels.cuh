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

}  // namespace cmr
