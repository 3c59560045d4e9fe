// scatter_kernels.cuh - the projected half of obs2d (environment/environment.py:74-86): scatter-mean of the
// predicted-overlap points' feature rows onto the pixel grid.
//
//   k_project (env_kernels.cuh) hands every VISIBLE predicted-overlap point to the 32-pixel bucket of the
//   pixel it projects to: slot = atomicAdd(bcnt[bucket]), bbuf[bucket][slot] = point << 7 | pixel % 128; the
//   point that sees a counter cross kLightMax appends the bucket to the batch-wide queue of heavy buckets.
//   The SET of entries of a bucket is deterministic, their order in the buffer (and the queue) is not.
//
//   k_tile_gather restores the order and adds the rows; per pixel the sum is sequential in point order
//   exactly like the reference's CPU scatter_add_ (deterministic, bit-identical, no floating-point atomics),
//   then divided by max(count, 1) and written to obs2d[b, C + c, pixel] channel-major.  The unit of work is
//   a (bucket, 64-channel slab) pair: 32 pixels x 64 channels of output, two channels per lane, a feature
//   row = one 256-byte warp load.  The kernel is bound by instruction issue, not by memory latency (measured
//   with %globaltimer marks per CTA, benchmarks/debug/cta_timing.py): loops run for the entries a unit HAS.
//     * light units - ONE WARP each, four per CTA, no CTA-wide barrier on the way.  The count and the (at
//       most kLightMax = 64) entries arrive in one L2 round trip (the entry loads are speculative), two keys
//       per lane.  Keys (pixel, point) are unique: a key's place in the order is the number of smaller keys,
//       counted with n shuffles; the sorted keys go through 256 bytes of shared memory.  Sixteen rows are
//       loaded together and added to a register-carried running sum per pixel.  The 32 x 64 result tile is
//       transposed through the warp's private shared memory and stored as float4.
//     * bucket CTAs (the first kHeavyCtas * B CTAs) - buckets with more than kLightMax points: far points
//       pile up on the horizon row.  Work item i = (queue[i / slabs], slab i % slabs), strided over the bucket
//       CTAs of the whole batch (an episode looking down a road has ten times the heavy buckets of one facing
//       a wall).  The CTA bins the entries by pixel in shared memory (32 counters), orders every pixel's list
//       by point (rank = number of smaller points IN THE PIXEL), and splits the sorted list at pixel
//       boundaries into four nearly equal parts, one per warp.  A bucket that overflowed its buffer (more than
//       kBucketCap points) is rebuilt from the episode's pixel-id list in chunks of kBucketCap, in point order.
//   The last CTA of the grid to finish resets the counters and the queue for the next observe.
//
// Used when the grid has at most kBucketMaxBuckets 32-pixel buckets (H*W <= 12288; KITTI is 5120, NuScenes
// 3200); larger grids take the search-based k_tile_scatter of env_kernels.cuh.
#pragma once
#include "common.cuh"
#include "env_kernels.cuh"

namespace cmr {

constexpr int kHeavyCtas = 16;          // bucket CTAs per episode of the batch
constexpr int kGatherThreads = 128;     // 4 warps = 4 light units
constexpr int kGatherWarps = kGatherThreads / 32;
constexpr int kSlab = 64;               // channels per unit: two per lane
constexpr int kTileStride = 36;         // floats per channel row of a 32-pixel result tile: 16-byte aligned rows
constexpr int kTileFloats = kSlab * kTileStride;
static_assert(kBucketPix == 32 && kLightMax == 64, "one lane per pixel / two keys per lane");
static_assert(kBucketCap % kGatherThreads == 0, "whole entries per thread");

constexpr size_t kGatherSmemLight = kGatherWarps * (sizeof(float) * kTileFloats + sizeof(unsigned) * kLightMax);
constexpr size_t kGatherSmemHeavy = sizeof(float) * kTileFloats + sizeof(unsigned) * 3 * kBucketCap + sizeof(int) * 128;
constexpr size_t kGatherSmem = kGatherSmemLight > kGatherSmemHeavy ? kGatherSmemLight : kGatherSmemHeavy;

__device__ __forceinline__ unsigned ld_cg_u32(const unsigned *p) {
    unsigned v;
    asm volatile("ld.global.cg.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ float2 ldg_f2(const float *p) { return __ldg(reinterpret_cast<const float2 *>(p)); }

// sum / count exactly as torch's true_divide (n >= 2): powers of two scale by the exact reciprocal
__device__ __forceinline__ float2 mean2(float2 s, int n) {
    const float nf = (float)n;
    if ((n & (n - 1)) == 0) {   // warp-uniform
        const float r = __fdiv_rn(1.f, nf);
        return make_float2(__fmul_rn(s.x, r), __fmul_rn(s.y, r));
    }
    return make_float2(__fdiv_rn(s.x, nf), __fdiv_rn(s.y, nf));
}

// tile [64 channels][kTileStride] (pixel minor) -> proj[(c0 + r) * P + p0 + pixel], by NT threads (t = 0..NT-1)
template <int NT>
__device__ __forceinline__ void store_tile(const float *tile, int c0, int C, int p0, int P, bool vec,
                                           float *__restrict__ proj, int t) {
    if (vec && p0 + kBucketPix <= P) {
        const int p4 = t & 7;
        float *dst = proj + (size_t)c0 * P + p0 + 4 * p4;
#pragma unroll 4
        for (int k = 0; k < kSlab * 8 / NT; ++k) {
            const int r = (t >> 3) + k * (NT >> 3);
            if (c0 + r < C)
                stg_stream4(dst + (size_t)r * P, *reinterpret_cast<const float4 *>(tile + r * kTileStride + 4 * p4));
        }
    } else {
        const int px = t & 31;
        for (int r = t >> 5; r < kSlab && c0 + r < C; r += NT >> 5)
            if (p0 + px < P) proj[(size_t)(c0 + r) * P + p0 + px] = tile[r * kTileStride + px];
    }
}

template <int NT>
__device__ __forceinline__ void zero_tile(float *tile, int t) {
    float4 *t4 = reinterpret_cast<float4 *>(tile);
#pragma unroll
    for (int k = 0; k < (kTileFloats / 4 + NT - 1) / NT; ++k)
        if (t + k * NT < kTileFloats / 4) t4[t + k * NT] = make_float4(0.f, 0.f, 0.f, 0.f);
}

// Adds the rows of sorted keys [jb, je) (pixel << 24 | point, ordered by pixel then point) to the per-pixel
// sums in `tile`, sixteen rows in flight.  kMeans: every pixel of the range is complete (its count is the
// length of its run), so the mean is formed at once; otherwise the sums are left in the tile.
constexpr int kRowBatch = 8;   // rows a warp loads before it adds them (registers: 2 floats + an address each)
constexpr unsigned kPadKey = 32u << 24;   // pixel 32 = the padding column of the tile, point 0: a harmless entry

// One entry of the sorted list: its row v is added to the running sum of its pixel; when the pixel changes
// the finished sum (kMeans: the mean, the run is the pixel's whole list) goes to the tile.
template <bool kMeans>
__device__ __forceinline__ void add_entry(unsigned key, float2 v, int &cur, int &cnt, float2 &run, float *tile_lane) {
    const int pl = (int)(key >> 24);
    if (pl != cur) {   // warp-uniform
        if (kMeans && cnt > 1) run = mean2(run, cnt);
        tile_lane[cur] = run.x;
        tile_lane[kTileStride + cur] = run.y;
        cur = pl;
        cnt = 0;
        run = kMeans ? make_float2(0.f, 0.f) : make_float2(tile_lane[cur], tile_lane[kTileStride + cur]);
    }
    run.x = __fadd_rn(run.x, v.x);
    run.y = __fadd_rn(run.y, v.y);
    ++cnt;
}

// Adds the rows of sorted keys [jb, je) (pixel << 24 | point, ordered by pixel then point) to the per-pixel
// sums in the tile, kRowBatch rows in flight.  `rows` already points at this lane's two channels of row 0;
// tile_lane at this lane's first channel row.  kMeans: every pixel of the range is complete (its count is
// the length of its run), so the mean is formed at once; otherwise the sums are left in the tile.
template <bool kMeans>
__device__ __forceinline__ void add_rows(const unsigned *skeys, int jb, int je, const float *__restrict__ rows,
                                         unsigned C, float *tile_lane) {
    int cur = 32, cnt = 0;   // starts on the padding column: the first real entry "finishes" an empty run there
    float2 run = make_float2(0.f, 0.f);
    int j0 = jb;
    for (; j0 + kRowBatch <= je; j0 += kRowBatch) {
        unsigned k[kRowBatch];
        float2 v[kRowBatch];
#pragma unroll
        for (int i = 0; i < kRowBatch; ++i) {
            k[i] = skeys[j0 + i];
            v[i] = ldg_f2(rows + (k[i] & 0xffffffu) * C);
        }
#pragma unroll
        for (int i = 0; i < kRowBatch; ++i) add_entry<kMeans>(k[i], v[i], cur, cnt, run, tile_lane);
    }
    for (; j0 < je; ++j0) {
        const unsigned k = skeys[j0];
        add_entry<kMeans>(k, ldg_f2(rows + (k & 0xffffffu) * C), cur, cnt, run, tile_lane);
    }
    if (kMeans && cnt > 1) run = mean2(run, cnt);
    tile_lane[cur] = run.x;
    tile_lane[kTileStride + cur] = run.y;
}

__global__ void __launch_bounds__(kGatherThreads, 5)
    k_tile_gather(int *bcnt, const unsigned *bbuf, int buckets, const int *hq, const void *pix, int pix16, const int *M,
                  const float *__restrict__ featT, const float *__restrict__ img_feat, int N, int ncap, int C, int P,
                  bool copy_image, bool vec, float *__restrict__ obs2d) {
    extern __shared__ __align__(16) float smem_g[];
    const int B = (int)gridDim.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int T = buckets;
    const int slabs = (C + kSlab - 1) / kSlab;
    int *hdr = bcnt + (size_t)B * kBucketStride;   // [0] heavy-queue length, [1] ticket

    if ((int)blockIdx.y >= kHeavyCtas) {
        // ------------------------------------------------------------------ light units: one warp each
        const int b = blockIdx.x;
        const int u = ((int)blockIdx.y - kHeavyCtas) * kGatherWarps + warp;
        const int bk = u / slabs, slab = u - bk * slabs;
        float *tile = smem_g + warp * kTileFloats;
        unsigned *sk = reinterpret_cast<unsigned *>(smem_g + kGatherWarps * kTileFloats) + warp * kLightMax;
        const int p0 = bk * kBucketPix;
        const int c0 = kSlab * slab;
        // lanes beyond C read channel 0 instead (their rows of the tile are never stored)
        const float *rows = featT + (size_t)b * N * C + (c0 + 2 * lane < C ? c0 + 2 * lane : 0);
        float *out = obs2d + (size_t)b * 2 * C * P;
        float *proj = out + (size_t)C * P;
        DBG_MARK(0);
        pdl_wait();   // counters and bucket buffers are written by k_project
        DBG_MARK(1);
        if (bk < T) {
            const unsigned *src = bbuf + ((size_t)b * buckets + bk) * kBucketCap;
            const unsigned e0 = ld_cg_u32(src + lane), e1 = ld_cg_u32(src + 32 + lane);   // speculative
            const int n = ld_cg_s32(bcnt + (size_t)b * kBucketStride + bk);
#ifdef CMR_DBG_TIMING
            if (threadIdx.x == 0) g_dbg[(blockIdx.y * gridDim.x + blockIdx.x) * 16 + 7] = (unsigned long long)(n + (e0 & 0) + (e1 & 0));
            DBG_MARK(2);
#endif
            if (copy_image) {   // image half when k_project could not carry it as TMA traffic
                const float *img = img_feat + (size_t)b * C * P;
                for (int r = 0; r < kSlab && c0 + r < C; ++r)
                    if (p0 + lane < P) out[(size_t)(c0 + r) * P + p0 + lane] = img[(size_t)(c0 + r) * P + p0 + lane];
            }
            if (n == 0) {
                if (vec && p0 + kBucketPix <= P) {
                    float *dst = proj + (size_t)c0 * P + p0 + 4 * (lane & 7);
#pragma unroll
                    for (int k = 0; k < kSlab / 4; ++k) {
                        const int r = (lane >> 3) + 4 * k;
                        if (c0 + r < C) stg_stream4(dst + (size_t)r * P, make_float4(0.f, 0.f, 0.f, 0.f));
                    }
                } else {
                    for (int r = 0; r < kSlab && c0 + r < C; ++r)
                        if (p0 + lane < P) proj[(size_t)(c0 + r) * P + p0 + lane] = 0.f;
                }
            } else if (n > 0 && n <= kLightMax) {   // otherwise the bucket CTAs own the bucket
                // key = pixel % 32 << 24 | point (unique); empty slots are all-ones (never smaller than a key)
                const unsigned k0 = lane < n ? ((e0 & 31u) << 24) | (e0 >> 7) : 0xffffffffu;
                const unsigned k1 = lane + 32 < n ? ((e1 & 31u) << 24) | (e1 >> 7) : 0xffffffffu;
                int r0 = 0, r1 = 0;
                const int na = min(n, 32);
                for (int j = 0; j < na; ++j) {
                    const unsigned kj = __shfl_sync(kFull, k0, j);
                    r0 += kj < k0;
                    r1 += kj < k1;
                }
                for (int j = 32; j < n; ++j) {
                    const unsigned kj = __shfl_sync(kFull, k1, j - 32);
                    r0 += kj < k0;
                    r1 += kj < k1;
                }
                // padded to whole batches with harmless entries: they sort last and land in the padding column
                const int npad = (n + kRowBatch - 1) / kRowBatch * kRowBatch;
                if (lane < n) sk[r0] = k0;
                else if (lane < npad) sk[lane] = kPadKey;
                if (lane + 32 < n) sk[r1] = k1;
                else if (lane + 32 < npad) sk[lane + 32] = kPadKey;
                zero_tile<32>(tile, lane);
                __syncwarp();
                DBG_MARK(8);
                add_rows<true>(sk, 0, npad, rows, (unsigned)C, tile + 2 * lane * kTileStride);
                __syncwarp();
                DBG_MARK(3);
                store_tile<32>(tile, c0, C, p0, P, vec, proj, lane);
            }
        }
        DBG_MARK(4);
    } else {
        // ------------------------------------------------------------------ bucket CTA
        float *tile = smem_g;                                             // [64][kTileStride] sums, then means
        unsigned *ent = reinterpret_cast<unsigned *>(tile + kTileFloats);  // [kBucketCap] pixel << 24 | point, as they arrive
        unsigned *ulist = ent + kBucketCap;                               // [kBucketCap] points, grouped by pixel
        unsigned *slist = ulist + kBucketCap;                             // [kBucketCap] keys, sorted
        int *pcnt = reinterpret_cast<int *>(slist + kBucketCap);          // [32] entries per pixel of this chunk
        int *pstart = pcnt + 32;                                          // [33] exclusive prefix
        int *ptotal = pstart + 33;                                        // [32] entries per pixel, all chunks
        int *misc = ptotal + 32;                                          // [4]
        DBG_MARK(0);
        pdl_wait();
        DBG_MARK(1);
        const int items = min(ld_cg_s32(hdr), B * kBucketMaxBuckets) * slabs;
#ifdef CMR_DBG_TIMING
        if (threadIdx.x == 0) {
            g_dbg[(blockIdx.y * gridDim.x + blockIdx.x) * 16 + 7] = 0;
            g_dbg[(blockIdx.y * gridDim.x + blockIdx.x) * 16 + 6] = 0;
        }
#endif
        for (int item = (int)(blockIdx.y * gridDim.x + blockIdx.x); item < items; item += kHeavyCtas * B) {
            const int hr = item / slabs, slab = item - hr * slabs;
            const int qe = ld_cg_s32(hq + hr);
            const int b = qe >> 16, bk = qe & 0xffff;
            const int c = ld_cg_s32(bcnt + (size_t)b * kBucketStride + bk);
            const int p0 = bk * kBucketPix;
            const int c0 = kSlab * slab;
            const float *rows = featT + (size_t)b * N * C + (c0 + 2 * lane < C ? c0 + 2 * lane : 0);
            float *proj = obs2d + (size_t)b * 2 * C * P + (size_t)C * P;
            DBG_MARK(2);
            zero_tile<kGatherThreads>(tile, tid);
            if (tid < 32) {
                pcnt[tid] = 0;
                ptotal[tid] = 0;
            }
            __syncthreads();
            // orders the n entries of `ent` (wcnt counted in pcnt) and adds their rows to the per-pixel sums
            auto accumulate = [&](int n) {
                __syncthreads();   // ent and pcnt are complete
                if (warp == 0) {
                    const int cp = pcnt[lane];
                    int inc = cp;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const int t = __shfl_up_sync(kFull, inc, o);
                        if (lane >= o) inc += t;
                    }
                    pstart[lane] = inc - cp;
                    if (lane == 31) pstart[32] = inc;
                    ptotal[lane] += cp;
                }
                __syncthreads();
                if (tid < 32) pcnt[tid] = 0;   // reused as the fill level while placing
                __syncthreads();
                for (int i = tid; i < n; i += kGatherThreads) {
                    const unsigned e = ent[i];
                    const int pl = (int)(e >> 24);
                    ulist[pstart[pl] + atomicAdd(&pcnt[pl], 1)] = e & 0xffffffu;
                }
                __syncthreads();
                for (int i = tid; i < n; i += kGatherThreads) {
                    const unsigned e = ent[i];
                    const int pl = (int)(e >> 24);
                    const unsigned pt = e & 0xffffffu;
                    const int s0 = pstart[pl], s1 = pstart[pl + 1];
                    int r = 0;
                    for (int j = s0; j < s1; ++j) r += ulist[j] < pt;
                    slist[s0 + r] = e;
                }
                if (tid < 32) pcnt[tid] = 0;   // ready for the next chunk's counts
                __syncthreads();
                DBG_MARK(9);
                // split at pixel boundaries into four nearly equal parts: warp w takes the pixels whose first
                // entry lies in [w * n / 4, (w + 1) * n / 4)
                const int ps = pstart[lane];
                const int lo = (warp * n + 3) >> 2, hi = ((warp + 1) * n + 3) >> 2;
                const int pb = __popc(__ballot_sync(kFull, ps < lo)), pe = warp == kGatherWarps - 1 ? 32 : __popc(__ballot_sync(kFull, ps < hi));
                const int jb = pb < 32 ? pstart[pb] : n, je = pe < 32 ? pstart[pe] : n;
                add_rows<false>(slist, jb, je, rows, (unsigned)C, tile + 2 * lane * kTileStride);
                __syncthreads();
            };
            if (c >= 0 && c <= kBucketCap) {
                const unsigned *src = bbuf + ((size_t)b * buckets + bk) * kBucketCap;
                for (int i = tid; i < c; i += kGatherThreads) {
                    const unsigned e = ld_cg_u32(src + i);
                    const unsigned pl = e & 31u;
                    ent[i] = (pl << 24) | (e >> 7);
                    atomicAdd(&pcnt[pl], 1);
                }
                DBG_MARK(8);
                accumulate(c);
            } else {
                // the bucket overflowed its buffer: rebuild it from the episode's pixel-id list, kBucketCap
                // points at a time, in point order (a chunk's points all precede the next chunk's)
                const int m_total = min(ld_cg_s32(M + b), N);
                int fill = 0;
                for (int m0 = 0; m0 < m_total; m0 += kGatherThreads) {
                    const int m = m0 + tid;
                    int id = -1;
                    if (m < m_total)
                        id = pix16 ? (int)static_cast<const uint16_t *>(pix)[(size_t)b * ncap + m]
                                   : static_cast<const int32_t *>(pix)[(size_t)b * ncap + m];
                    const bool hit = id >= p0 && id < p0 + kBucketPix && id < P;
                    const unsigned bal = __ballot_sync(kFull, hit);
                    if (lane == 0) misc[warp] = __popc(bal);
                    __syncthreads();
                    int before = 0, total = 0;
                    for (int w = 0; w < kGatherWarps; ++w) {
                        if (w < warp) before += misc[w];
                        total += misc[w];
                    }
                    if (fill + total > kBucketCap) {   // uniform
                        accumulate(fill);
                        fill = 0;
                    }
                    if (hit) {
                        const unsigned pl = (unsigned)(id - p0);
                        ent[fill + before + __popc(bal & ((1u << lane) - 1))] = (pl << 24) | (unsigned)m;
                        atomicAdd(&pcnt[pl], 1);
                    }
                    fill += total;
                    __syncthreads();
                }
                if (fill > 0) accumulate(fill);
            }
            DBG_MARK(10);
            // sums -> means: warp w takes the pixels p % 4 == w
#pragma unroll 1
            for (int pl = warp; pl < 32; pl += kGatherWarps) {
                const int n = ptotal[pl];
                if (n > 1) {
                    const float2 m = mean2(make_float2(tile[(2 * lane) * kTileStride + pl], tile[(2 * lane + 1) * kTileStride + pl]), n);
                    tile[(2 * lane) * kTileStride + pl] = m.x;
                    tile[(2 * lane + 1) * kTileStride + pl] = m.y;
                }
            }
            __syncthreads();
            store_tile<kGatherThreads>(tile, c0, C, p0, P, vec, proj, tid);
            DBG_MARK(11);
#ifdef CMR_DBG_TIMING
            if (threadIdx.x == 0) {
                g_dbg[(blockIdx.y * gridDim.x + blockIdx.x) * 16 + 7] += (unsigned long long)(c < 0 ? 0 : c);
                g_dbg[(blockIdx.y * gridDim.x + blockIdx.x) * 16 + 6] += 1;
            }
#endif
            __syncthreads();
        }
        DBG_MARK(4);
    }
    // the last CTA of the grid to get here clears the counters and the queue for the next observe: every
    // CTA has read what it needed from them before it takes its ticket
    __shared__ int s_last;
    __syncthreads();
    if (tid == 0) {
        __threadfence();
        s_last = atomicAdd(hdr + 1, 1) == (int)(gridDim.x * gridDim.y) - 1;
    }
    __syncthreads();
    DBG_MARK(5);
    if (s_last) {
        for (int e = 0; e < B; ++e)
            for (int t = tid; t < T; t += kGatherThreads) bcnt[(size_t)e * kBucketStride + t] = 0;
        if (tid < 2) hdr[tid] = 0;
    }
}

}  // namespace cmr
