// scatter_kernels.cuh - the projected half of obs2d (environment/environment.py:74-86) in two kernels:
//
//   k_bin          one CTA per episode: STABLE counting sort of the episode's predicted-overlap points by
//                  32-pixel bucket of the pixel they project to (out-of-frustum points are dropped).  The
//                  result is a CSR: boff[bucket] .. boff[bucket+1] index `order`, whose entries
//                  (point << 7 | pixel % 128) are in point order inside every bucket.  ~9000 ids per
//                  episode, two passes of warp-level match/ballot ranking: a few microseconds.
//   k_tile_gather  one CTA per 128-pixel tile (or per 32-pixel bucket when a tile is dense - far points
//                  pile up on the horizon row): reads ITS entries from the CSR - no searching - and adds the
//                  feature rows (point-major, 4C contiguous bytes) to the pixels in point order: warp w owns
//                  the pixels p % 8 == w, so the per-pixel sums are sequential exactly like the reference's
//                  CPU scatter_add_ (deterministic, bit-identical, no floating-point atomics); then divides
//                  by max(count, 1) and writes obs2d[b, C + c, pixel] channel-major.
//
// Used when the grid has at most kBinMaxBuckets 32-pixel buckets (H*W <= 12288; KITTI is 5120, NuScenes
// 3200); larger grids take the search-based k_tile_scatter of env_kernels.cuh.
#pragma once
#include <cooperative_groups.h>

#include "common.cuh"

namespace cmr {

constexpr int kBinCluster = 8;          // CTAs (SMs) that share the counting sort of one episode
constexpr int kBucketPix = 32;          // pixels per bucket (power of two)
constexpr int kBinMaxBuckets = 384;     // per episode
constexpr int kBoffStride = kBinMaxBuckets + 8;
constexpr int kBinThreads = 1024;
constexpr int kGatherTile = 128;        // pixels per k_tile_gather tile = 4 buckets
constexpr int kHeavyTile = 160;         // a tile with more points than this is split into its 4 buckets
constexpr int kChunk = 1024;            // CSR entries staged per round
constexpr int kOwnCap = 256;            // private (owned) entries per warp between flushes

// One thread-block CLUSTER of kBinCluster CTAs per episode: every CTA ranks a contiguous quarter of the id
// list (the sort is instruction-bound on one SM otherwise), the per-bucket totals of the CTAs are exchanged
// through distributed shared memory, and every CTA derives its own write offsets.
template <typename PixT>
__global__ void __launch_bounds__(kBinThreads) k_bin(const PixT *pix, const int *M, int N, int ncap, int P,
                                                      unsigned *__restrict__ order, int *__restrict__ boff) {
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    pdl_launch_dependents();
    extern __shared__ int hist[];            // [32 warps][T] -> exclusive prefix over warps, then running counters
    __shared__ int tot[kBinMaxBuckets];      // points per bucket found by THIS CTA (read by the peers)
    __shared__ int cbase[kBinMaxBuckets];    // points per bucket found by the CTAs before this one
    __shared__ int goff[kBinMaxBuckets];     // exclusive prefix over buckets of the episode totals
    __shared__ int wsum[32];
    const int rank = (int)cluster.block_rank(), csz = (int)cluster.num_blocks();
    const int b = blockIdx.x / csz;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int T = (P + kBucketPix - 1) / kBucketPix;
    for (int i = tid; i < 32 * T; i += kBinThreads) hist[i] = 0;
    __syncthreads();

    pdl_wait();   // the id list is written by k_project
    const int m_total = min(ld_cg_s32(M + b), N);
    const PixT *pw = pix + (size_t)b * ncap;
    const int per_warp = ((m_total + 32 * csz - 1) / (32 * csz) + 31) / 32 * 32;   // whole steps of 32 ids
    const int beg = min((rank * 32 + warp) * per_warp, m_total), end = min(beg + per_warp, m_total);
    int *myhist = hist + warp * T;

    // pass A: per-warp histogram over buckets.  kBinBatch steps of 32 ids are loaded before the first one
    // is ranked (the loop is otherwise a chain of dependent L2 round trips).
    constexpr int kBinBatch = 8;
    for (int m0 = beg; m0 < end; m0 += 32 * kBinBatch) {
        unsigned ids[kBinBatch];
#pragma unroll
        for (int s2 = 0; s2 < kBinBatch; ++s2) {
            const int m = m0 + s2 * 32 + lane;
            ids[s2] = m < end ? (unsigned)pw[m] : 0xffffffffu;
        }
#pragma unroll
        for (int s2 = 0; s2 < kBinBatch; ++s2) {
            if (m0 + s2 * 32 >= end) break;
            const unsigned id = ids[s2];
            const unsigned key = id < (unsigned)P ? id / kBucketPix : 0xffffffffu;   // H*W (dump bin) drops out
            const unsigned same = __match_any_sync(kFull, key);
            if (key != 0xffffffffu && (__ffs(same) - 1) == lane) myhist[key] += __popc(same);
            __syncwarp();
        }
    }
    __syncthreads();
    // exclusive prefix over warps for every bucket, CTA totals
    for (int t = tid; t < T; t += kBinThreads) {
        int run = 0;
        for (int w = 0; w < 32; ++w) {
            int c = hist[w * T + t];
            hist[w * T + t] = run;
            run += c;
        }
        tot[t] = run;
    }
    cluster.sync();   // every CTA's totals are visible cluster-wide
    int v = 0;
    if (tid < T) {
        int before = 0;
        for (int c = 0; c < csz; ++c) {
            const int x = *cluster.map_shared_rank(&tot[tid], c);
            if (c < rank) before += x;
            v += x;
        }
        cbase[tid] = before;
    }
    // exclusive prefix over buckets of the episode totals (T <= 384: one value per thread of the first 12 warps)
    {
        int inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int t = __shfl_up_sync(kFull, inc, o);
            if (lane >= o) inc += t;
        }
        if (lane == 31) wsum[warp] = inc;
        __syncthreads();
        if (warp == 0) {
            int w = wsum[lane], winc = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                int t = __shfl_up_sync(kFull, winc, o);
                if (lane >= o) winc += t;
            }
            wsum[lane] = winc - w;
        }
        __syncthreads();
        const int excl = wsum[warp] + inc - v;
        if (tid < T) {
            goff[tid] = excl;
            if (rank == 0) boff[(size_t)b * kBoffStride + tid] = excl;
        }
        if (tid == T - 1 && rank == 0) boff[(size_t)b * kBoffStride + T] = excl + v;
    }
    cluster.sync();   // peers have finished reading tot[]; goff/cbase are complete
    // pass B: stable placement
    unsigned *out = order + (size_t)b * ncap;
    for (int m0 = beg; m0 < end; m0 += 32 * kBinBatch) {
        unsigned ids[kBinBatch];
#pragma unroll
        for (int s2 = 0; s2 < kBinBatch; ++s2) {
            const int m = m0 + s2 * 32 + lane;
            ids[s2] = m < end ? (unsigned)pw[m] : 0xffffffffu;
        }
#pragma unroll
        for (int s2 = 0; s2 < kBinBatch; ++s2) {
            if (m0 + s2 * 32 >= end) break;
            const int m = m0 + s2 * 32 + lane;
            const unsigned id = ids[s2];
            const unsigned key = id < (unsigned)P ? id / kBucketPix : 0xffffffffu;
            const unsigned same = __match_any_sync(kFull, key);
            int basepos = 0;
            if (key != 0xffffffffu) basepos = goff[key] + cbase[key] + myhist[key];
            __syncwarp();
            if (key != 0xffffffffu) {
                out[basepos + __popc(same & ((1u << lane) - 1))] = ((unsigned)m << 7) | (id & (kGatherTile - 1));
                if ((__ffs(same) - 1) == lane) myhist[key] += __popc(same);
            }
            __syncwarp();
        }
    }
}

// CQ2 = 64-channel slabs per feature row (a lane owns channels 2*lane, 2*lane+1 of every slab)
template <int CQ2>
__global__ void __launch_bounds__(256, CQ2 <= 1 ? 4 : 2) k_tile_gather(const unsigned *order, const int *boff,
                                                                        const float *__restrict__ featT,
                                                                        const float *__restrict__ img_feat,
                                                                        const float *__restrict__ K, int W, int N,
                                                                        int ncap, int C, int P, int tiles,
                                                                        bool copy_image, float *__restrict__ obs2d) {
    extern __shared__ __align__(16) float smem_g[];
    const int stride = C + 2;                                          // even: 8-byte aligned rows
    float *acc = smem_g;                                               // [128][C+2] sums, pixel-major
    int *cnt = reinterpret_cast<int *>(acc + kGatherTile * stride);   // [128] points per pixel
    unsigned *elist = reinterpret_cast<unsigned *>(cnt + kGatherTile); // [kChunk] CSR entries of this round
    unsigned *own = elist + kChunk;                                    // [8][kOwnCap] entries owned by warp w, in order

    // ---- which tile?  x = episode; y = 4 * rank + part, ranks walk the 128-pixel tiles outwards from the
    // horizon row v = cy (where the dense tiles are) so that the long CTAs start first
    const int b = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int part = blockIdx.y & 3;
    int t;
    {
        const float cy = __ldg(K + (size_t)b * 9 + 5);
        int row = (int)cy;
        row = row < 0 ? 0 : row;
        int c = (int)(((long long)row * W) / kGatherTile);
        c = c > tiles - 1 ? tiles - 1 : c;
        const int k = blockIdx.y >> 2, m = min(c, tiles - 1 - c);
        if (k <= 2 * m)
            t = (k & 1) ? c + (k + 1) / 2 : c - k / 2;
        else
            t = (c < tiles - 1 - c) ? k : tiles - 1 - k;
    }
    const int T = (P + kBucketPix - 1) / kBucketPix;
    const int bk0 = min(4 * t, T), bk4 = min(4 * t + 4, T);

    DBG_MARK(0);
    pdl_wait();   // the CSR is written by k_bin
    DBG_MARK(1);
    const int *bo = boff + (size_t)b * kBoffStride;
    const int tile_cnt = ld_cg_s32(bo + bk4) - ld_cg_s32(bo + bk0);
    const bool heavy = tile_cnt > kHeavyTile;
    if (!heavy && part != 0) return;                    // a light tile is one CTA
    const int first = heavy ? min(bk0 + part, T) : bk0, last = heavy ? min(bk0 + part + 1, T) : bk4;
    const int p0 = t * kGatherTile + (heavy ? part * kBucketPix : 0);      // first pixel of this CTA
    const int width = heavy ? kBucketPix : kGatherTile;
    if (p0 >= P) return;
    const int np = min(width, P - p0);
    const int poff = heavy ? part * kBucketPix : 0;     // entry pixel ids are relative to the 128-pixel tile
    const int e0 = ld_cg_s32(bo + first), e1 = ld_cg_s32(bo + last);
    const unsigned *ord = order + (size_t)b * ncap;
    const float *rows = featT + (size_t)b * N * C;
    float *out = obs2d + (size_t)b * 2 * C * P;
    {
        float4 *a4 = reinterpret_cast<float4 *>(acc);
        const int n4 = (width * stride + 3) / 4;
        for (int i = tid; i < n4; i += 256) a4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (tid < kGatherTile) cnt[tid] = 0;
    }
    __syncthreads();

    if (copy_image) {   // image half when k_project could not carry it as TMA traffic
        const float *img = img_feat + (size_t)b * C * P;
        for (int i = tid; i < C * np; i += 256) {
            int c = i / np, p = i - c * np;
            out[(size_t)c * P + p0 + p] = img[(size_t)c * P + p0 + p];
        }
    }

    DBG_MARK(2);
    unsigned *mine_list = own + warp * kOwnCap;
    constexpr int kBatch = CQ2 <= 1 ? 16 : 8;
    for (int r0 = e0; r0 < e1; r0 += kChunk) {
        const int n_round = min(kChunk, e1 - r0);
        // stage this round's entries (coalesced) and start their feature rows towards L2
        for (int i = tid; i < n_round; i += 256) {
            const unsigned e = ord[r0 + i];   // plain load: written by k_bin, complete before pdl_wait returned
            elist[i] = e;
            const float *row = rows + (size_t)(e >> 7) * C;
            for (int q = 0; q < C; q += 32) prefetch_l2(row + q);
        }
        __syncthreads();
        DBG_MARK(8);
        // every warp walks the round in point order, keeps the entries of ITS pixels, and whenever its
        // private list is (nearly) full - and at the end - adds those rows in order
        int n_own = 0;
        for (int i0 = 0; i0 < n_round + 32; i0 += 32) {   // the last pass (i0 >= n_round) only flushes
            if (i0 >= n_round || n_own > kOwnCap - 32) {
                __syncwarp();
                DBG_MARK(9);
#ifdef CMR_DBG_TIMING
                if (threadIdx.x == 0) { int _id = blockIdx.y * gridDim.x + blockIdx.x; if (_id < 8192) g_dbg[_id * 16 + 11] = n_own; }
#endif
                for (int j0 = 0; j0 < n_own; j0 += kBatch) {
                    float2 v[kBatch][CQ2];
#pragma unroll
                    for (int k = 0; k < kBatch; ++k) {
                        if (j0 + k < n_own) {
                            const float *row = rows + (size_t)(mine_list[j0 + k] >> 7) * C + 2 * lane;
#pragma unroll
                            for (int q = 0; q < CQ2; ++q)
                                if (q * 64 + 2 * lane < C) v[k][q] = __ldg(reinterpret_cast<const float2 *>(row + q * 64));
                        }
                    }
                    // add pass: consecutive entries of the same pixel keep their running sum in registers, so
                    // a hot pixel (hundreds of far points on the vanishing point) is a chain of FADDs instead
                    // of a chain of shared-memory round trips
                    int cur = -1;
                    float2 run[CQ2];
#pragma unroll
                    for (int k = 0; k < kBatch; ++k) {
                        if (j0 + k < n_own) {
                            const int pl = (int)(mine_list[j0 + k] & 127u) - poff;
                            if (pl != cur) {   // warp-uniform
                                if (cur >= 0) {
#pragma unroll
                                    for (int q = 0; q < CQ2; ++q)
                                        if (q * 64 + 2 * lane < C)
                                            *reinterpret_cast<float2 *>(acc + cur * stride + 2 * lane + q * 64) = run[q];
                                }
                                cur = pl;
#pragma unroll
                                for (int q = 0; q < CQ2; ++q)
                                    if (q * 64 + 2 * lane < C)
                                        run[q] = *reinterpret_cast<const float2 *>(acc + cur * stride + 2 * lane + q * 64);
                            }
#pragma unroll
                            for (int q = 0; q < CQ2; ++q) {
                                if (q * 64 + 2 * lane < C) {
                                    run[q].x = __fadd_rn(run[q].x, v[k][q].x);
                                    run[q].y = __fadd_rn(run[q].y, v[k][q].y);
                                }
                            }
                        }
                    }
                    if (cur >= 0) {
#pragma unroll
                        for (int q = 0; q < CQ2; ++q)
                            if (q * 64 + 2 * lane < C)
                                *reinterpret_cast<float2 *>(acc + cur * stride + 2 * lane + q * 64) = run[q];
                    }
                }
                DBG_MARK(10);
                n_own = 0;
                __syncwarp();
            }
            if (i0 < n_round) {
                const unsigned e = (i0 + lane < n_round) ? elist[i0 + lane] : 0xffffffffu;
                const bool is_mine = e != 0xffffffffu && (e & 7u) == (unsigned)warp;
                const unsigned mask = __ballot_sync(kFull, is_mine);
                if (is_mine) {
                    mine_list[n_own + __popc(mask & ((1u << lane) - 1))] = e;
                    atomicAdd(&cnt[(e & 127u) - poff], 1);
                }
                n_own += __popc(mask);
            }
        }
        __syncthreads();   // elist is rewritten by the next round
    }
    __syncthreads();

    DBG_MARK(3);
    // mean + channel-major store: obs2d[b, C + c, p0 + p].  lane <-> pixel; the divisor is per pixel, so it
    // is classified once: n <= 1 and powers of two scale exactly by a multiplication, anything else needs
    // the IEEE division.
    float *proj = out + (size_t)C * P;
    for (int k0 = 0; k0 < width; k0 += 32) {
        const int p = k0 + lane;
        int n = cnt[p];
        n = n < 1 ? 1 : n;
        const float nf = (float)n;
        const bool pow2 = (n & (n - 1)) == 0;
        const float scale = pow2 ? __fdiv_rn(1.f, nf) : 0.f;   // exact
        const bool any_hard = __any_sync(kFull, !pow2);
        const float *a = acc + p * stride;
        float *dst = proj + p0 + p;
        if (!any_hard) {
            for (int c = warp; c < C; c += 8)
                if (p < np) stg_stream1(dst + (size_t)c * P, __fmul_rn(a[c], scale));
        } else {
            for (int c = warp; c < C; c += 8) {
                const float s1 = a[c];
                const float v = pow2 ? __fmul_rn(s1, scale) : __fdiv_rn(s1, nf);
                if (p < np) stg_stream1(dst + (size_t)c * P, v);
            }
        }
    }
    DBG_MARK(4);
#ifdef CMR_DBG_TIMING
    if (threadIdx.x == 0) {
        int _id = blockIdx.y * gridDim.x + blockIdx.x;
        if (_id < 8192) {
            unsigned smid;
            asm volatile("mov.u32 %0, %smid;" : "=r"(smid));
            g_dbg[_id * 16 + 5] = smid;
            g_dbg[_id * 16 + 6] = (unsigned long long)p0;
            g_dbg[_id * 16 + 7] = (unsigned long long)(e1 - e0);
        }
    }
#endif
}

}  // namespace cmr
