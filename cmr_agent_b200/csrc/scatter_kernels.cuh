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
//   a bucket: 32 pixels, done in slabs of 64 channels (two channels per lane, a feature row = one 256-byte
//   warp load; the order is established once and reused for every slab).  What bounds the kernel is the
//   latency of a warp's dependent instructions, not memory (measured with %globaltimer marks per CTA,
//   benchmarks/debug/cta_timing.py, and ncu's stall reasons): the add loop is branch-free - the sorted keys
//   carry first-of-pixel / last-of-pixel flags - and runs in batches whose loads are all in flight.
//     * light units - ONE WARP each, eight per CTA, no CTA-wide barrier on the way.  The count and the (at
//       most kLightMax = 64) entries arrive in one L2 round trip (the entry loads are speculative), two keys
//       per lane; every lane at once starts the rows of its entries towards L2.  Keys (pixel, point) are
//       unique: a key's place in the order is the number of smaller keys, counted with n shuffles; the sorted
//       keys go through 256 bytes of shared memory.  Two batches of eight rows alternate in registers and are
//       added to a register-carried running sum per pixel.  The 32 x 64 result tile is built in the warp's
//       private shared memory in the layout of TMA's 128-byte swizzle and leaves as ONE tiled TMA store.
//     * bucket CTAs (kHeavyCtas * B of them, alternating with light CTAs in launch order) - buckets with more
//       than kLightMax points: far points pile up on the horizon row.  Work item i = queue[i], strided over
//       the bucket CTAs of the whole batch (an episode looking down a road has ten times the heavy buckets of
//       one facing a wall).  The CTA bins the entries by pixel in shared memory (32 counters), orders every
//       pixel's list by point (rank = number of smaller points IN THE PIXEL), fetches the rows with cp.async
//       in sorted order into the shared memory the light units would use for their tiles (about 200 rows per
//       round trip), and splits the sorted list at pixel boundaries into eight nearly equal parts, one per
//       warp: a pixel's rows are then consecutive, the inner loop is a load and two additions per row.
//       A bucket that overflowed its buffer (more than kBucketCap points) is rebuilt from the episode's
//       pixel-id list in chunks of kBucketCap, in point order.
//     * the image half of obs2d (obs2d[b, 0:C] = img_geo_feat[b], environment.py:83) rides along: before it waits
//       for k_project, every light warp fetches the [64][32] box of its bucket's pixels into its (still idle) tile
//       with a TMA load and sends it on with a TMA store - no registers, no LSU instructions.
//     * a cost volume's four tail channels (the summed scores + padding, models/IterModel.py:343) get no slab pass
//       from the light warps: a lane adds them up for its own pixel (`tail`).
//   Counters are cleared by their readers: a light bucket's counter has one reader (its warp), a heavy bucket's
//   two (its warp, which only learns that the bucket is heavy, and the bucket CTA) - each adds kCountSeen, and
//   the one that finds it already there clears the counter; the last bucket CTA to finish clears the queue
//   length.  No memset between observes, no grid-wide pass, no word that every CTA reads.
//
// Used when the grid has at most kBucketMaxBuckets 32-pixel buckets (H*W <= 12288; KITTI is 5120, NuScenes
// 3200); larger grids take the search-based k_tile_scatter of env_kernels.cuh.
#pragma once
#include "common.cuh"
#include "env_kernels.cuh"

namespace cmr {

#ifndef CMR_GATHER_MINB
#define CMR_GATHER_MINB 3
#endif
#ifndef CMR_HEAVY_CTAS
#define CMR_HEAVY_CTAS 16
#endif
constexpr int kHeavyCtas = CMR_HEAVY_CTAS;          // bucket CTAs per episode of the batch
constexpr int kGatherThreads = 256;     // 8 warps = 8 light units
constexpr int kGatherWarps = kGatherThreads / 32;
constexpr int kSlab = 64;               // channels per pass: two per lane
constexpr int kTileFloats = kSlab * kBucketPix;   // result tile [64 channels][32 pixels]: 128-byte rows, laid out as
                                                  // TMA's 128-byte swizzle wants them (16-byte chunk ^= row % 8)
constexpr int kRowBatch = 8;            // rows per batch; a warp keeps two batches going
constexpr unsigned kKeyFirst = 1u << 31, kKeyLast = 1u << 30;   // flags of a sorted key: pixel << 24 | point
constexpr unsigned kPadKey = 0;   // no flags, point 0: loaded, added to a run that is never stored
static_assert(kBucketPix == 32 && kLightMax == 64, "one lane per pixel / two keys per lane");
static_assert(kMidMax * 16 + kMidMax * 4 <= kSlab * kBucketPix * 4, "a mid unit's tail rows and binned keys share the idle tile");
static_assert(kBucketCap % kGatherThreads == 0, "whole entries per thread");

// a warp's sorted keys + [32] points per pixel + [32] cursors: as many keys as the launch's heavy_from allows
__host__ __device__ constexpr int warp_keys(int heavy_from) { return heavy_from > kLightMax ? heavy_from + 64 : kLightMax + 32; }
constexpr int kMidPer = kMidMax / 32;   // entries per lane of a mid unit
__host__ __device__ constexpr size_t gather_smem_light(int heavy_from) {
    return kGatherWarps * (sizeof(float) * kTileFloats + sizeof(unsigned) * warp_keys(heavy_from));
}
constexpr size_t kGatherSmemLight = gather_smem_light(kLightMax);
constexpr size_t kGatherSmemHeavy = sizeof(float) * kTileFloats + sizeof(unsigned) * 3 * kBucketCap + sizeof(int) * 128;
constexpr size_t kGatherSmem = kGatherSmemLight > kGatherSmemHeavy ? kGatherSmemLight : kGatherSmemHeavy;
// what a bucket CTA does not need of the light units' tiles stages feature rows (kSlab floats per row), together
// with the two lists that are dead once the entries are sorted
constexpr size_t kGatherSmemHeavyLive = sizeof(float) * kTileFloats + sizeof(unsigned) * kBucketCap + sizeof(int) * 128;
constexpr int kStageRows = (int)((kGatherSmem - kGatherSmemHeavyLive) / (sizeof(float) * kSlab));
static_assert(kStageRows >= 64, "a bucket CTA stages at least 64 rows at a time");

__device__ __forceinline__ unsigned ld_cg_u32(const unsigned *p) {
    unsigned v;
    asm volatile("ld.global.cg.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ float2 ldg_f2(const float *p) { return __ldg(reinterpret_cast<const float2 *>(p)); }
__device__ __forceinline__ float4 ldg_f4(const float *p) { return __ldg(reinterpret_cast<const float4 *>(p)); }
__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
    asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
}

// sum / count exactly as torch's true_divide (n >= 2): powers of two scale by the exact reciprocal
__device__ __forceinline__ float2 mean2(float2 s, int n) {
    const float nf = (float)n;
    if ((n & (n - 1)) == 0) {   // warp-uniform
        const float r = __fdiv_rn(1.f, nf);
        return make_float2(__fmul_rn(s.x, r), __fmul_rn(s.y, r));
    }
    return make_float2(__fdiv_rn(s.x, nf), __fdiv_rn(s.y, nf));
}

// float offset of (channel row r, pixel px) in the swizzled tile: the 16-byte chunk px / 4 is XORed with r % 8
__device__ __forceinline__ int tile_off(int r, int px) { return r * kBucketPix + (px ^ ((r & 7) << 2)); }

// A lane's two channel rows (2 * lane, 2 * lane + 1): row 2 * lane + 1 sits 32 floats further with chunk bit 0 flipped
struct LaneRows {
    float *row;   // tile + 2 * lane * 32
    int xs;       // ((2 * lane) & 7) << 2
    __device__ __forceinline__ int off(int px) const { return px ^ xs; }
    __device__ __forceinline__ float2 get(int px) const {
        const int o = off(px);
        return make_float2(row[o], row[kBucketPix + (o ^ 4)]);
    }
    __device__ __forceinline__ void put(int px, float2 v) const {
        const int o = off(px);
        row[o] = v.x;
        row[kBucketPix + (o ^ 4)] = v.y;
    }
};
__device__ __forceinline__ LaneRows lane_rows(float *tile, int lane) { return LaneRows{tile + 2 * lane * kBucketPix, ((2 * lane) & 7) << 2}; }

// tile -> proj[(c0 + r) * P + p0 + pixel].  tma: one tiled TMA store by the calling thread (all writers have
// fenced and synchronised; returns when the tile has been read).  Otherwise NT threads copy it (t = 0..NT-1).
__device__ __forceinline__ void store_tile_tma(const float *tile, const CUtensorMap *map, int y, int p0, int b) {
    tma_store_3d(map, p0, y, b, tile);
    bulk_commit();
    bulk_wait_read_all();
}
template <int NT>
__device__ __forceinline__ void store_tile(const float *tile, int c0, int C, int p0, int P, float *__restrict__ proj, int t) {
    const int px = t & 31;
    for (int r = t >> 5; r < kSlab && c0 + r < C; r += NT >> 5)
        if (p0 + px < P) proj[(size_t)(c0 + r) * P + p0 + px] = tile[tile_off(r, px)];
}

// zeros for an empty bucket, all channels, by one warp
__device__ __forceinline__ void store_zeros(int C, int p0, int P, bool vec, float *__restrict__ proj, int lane) {
    if (vec && p0 + kBucketPix <= P) {
        float *dst = proj + p0 + 4 * (lane & 7);
#pragma unroll 4
        for (int r = lane >> 3; r < C; r += 4) stg_stream4(dst + (size_t)r * P, make_float4(0.f, 0.f, 0.f, 0.f));
    } else {
        for (int r = 0; r < C; ++r)
            if (p0 + lane < P) proj[(size_t)r * P + p0 + lane] = 0.f;
    }
}

template <int NT>
__device__ __forceinline__ void zero_tile(float *tile, int t) {
    float4 *t4 = reinterpret_cast<float4 *>(tile);
#pragma unroll
    for (int k = 0; k < kTileFloats / 4 / NT; ++k) t4[t + k * NT] = make_float4(0.f, 0.f, 0.f, 0.f);
}
static_assert(kTileFloats % (4 * kGatherThreads) == 0, "zero_tile covers the tile");

// One entry of the flagged sorted list, without a branch: a run (the entries of one pixel, in point order)
// starts from zero (kFromTile: from what earlier chunks left in the tile) and is written when it ends.
template <bool kFromTile>
__device__ __forceinline__ void add_entry(unsigned key, float2 v, float2 &run, const LaneRows &tl) {
    const int pl = (int)((key >> 24) & 31u);
    const bool first = key & kKeyFirst, last = key & kKeyLast;   // warp-uniform
    float2 base = run;
    if (first) base = kFromTile ? tl.get(pl) : make_float2(0.f, 0.f);
    run.x = __fadd_rn(base.x, v.x);
    run.y = __fadd_rn(base.y, v.y);
    if (last) tl.put(pl, run);
}

// Adds the rows of the flagged sorted keys [jb, je) to the per-pixel sums of the tile.  Two batches of
// kRowBatch rows alternate: while one is added the other is in flight (a pixel on the vanishing point holds
// hundreds of rows that one warp must add in order - what it costs is the exposed load latency per batch).
// `rows` points at this lane's two channels of row 0.
__device__ __forceinline__ void load_batch(unsigned (&k)[kRowBatch], float2 (&v)[kRowBatch], const unsigned *skeys, int j0,
                                           int je, const float *__restrict__ rows, unsigned C) {
    if (j0 + kRowBatch <= je) {   // warp-uniform
#pragma unroll
        for (int i = 0; i < kRowBatch; ++i) {
            k[i] = skeys[j0 + i];
            v[i] = ldg_f2(rows + (k[i] & 0xffffffu) * C);
        }
    } else {
#pragma unroll
        for (int i = 0; i < kRowBatch; ++i) {
            // past the end: a harmless entry (row 0, no flags: added to a run that is never stored)
            k[i] = j0 + i < je ? skeys[j0 + i] : kPadKey;
            v[i] = ldg_f2(rows + (k[i] & 0xffffffu) * C);
        }
    }
}
template <bool kFromTile>
__device__ __forceinline__ void add_batch(const unsigned (&k)[kRowBatch], const float2 (&v)[kRowBatch], float2 &run,
                                          const LaneRows &tl) {
    unsigned flags = 0;
#pragma unroll
    for (int i = 0; i < kRowBatch; ++i) flags |= k[i];
    if (!(flags & (kKeyFirst | kKeyLast))) {
        // the middle of a long run (hundreds of far points share the pixel of the vanishing point): nothing
        // but the additions, in order
#pragma unroll
        for (int i = 0; i < kRowBatch; ++i) {
            run.x = __fadd_rn(run.x, v[i].x);
            run.y = __fadd_rn(run.y, v[i].y);
        }
    } else {
#pragma unroll
        for (int i = 0; i < kRowBatch; ++i) add_entry<kFromTile>(k[i], v[i], run, tl);
    }
}
template <bool kFromTile>
__device__ __forceinline__ void add_rows(const unsigned *skeys, int jb, int je, const float *__restrict__ rows,
                                         unsigned C, const LaneRows &tl) {
    float2 run = make_float2(0.f, 0.f);
    unsigned ka[kRowBatch], kb[kRowBatch];
    float2 va[kRowBatch], vb[kRowBatch];
    if (jb < je) load_batch(ka, va, skeys, jb, je, rows, C);
    for (int j0 = jb; j0 < je; j0 += 2 * kRowBatch) {
        const bool more = j0 + kRowBatch < je;   // warp-uniform
        if (more) load_batch(kb, vb, skeys, j0 + kRowBatch, je, rows, C);
        add_batch<kFromTile>(ka, va, run, tl);
        if (j0 + 2 * kRowBatch < je) load_batch(ka, va, skeys, j0 + 2 * kRowBatch, je, rows, C);
        if (more) add_batch<kFromTile>(kb, vb, run, tl);
    }
}

// sums -> means for the pixels of `mask` (bit p: pixel p has more than one point); cnt_of_lane = points of
// pixel `lane`.  One warp.
__device__ __forceinline__ void mean_pass(unsigned mask, int cnt_of_lane, const LaneRows &tl) {
    while (mask) {   // warp-uniform
        const int pl = __ffs(mask) - 1;
        mask &= mask - 1;
        const int n = __shfl_sync(kFull, cnt_of_lane, pl);
        tl.put(pl, mean2(tl.get(pl), n));
    }
}


// kMid: the launch's warps also take buckets of kLightLimit < n <= kMidMax entries (cost volumes).  A template
// parameter, not a branch: with the mid unit compiled into the observe's kernel its light units lost 6 % (register
// allocation and scheduling of the shared code; measured: 33.5 -> 35.7 us at B = 32).
template <bool kMid>
__global__ void __launch_bounds__(kGatherThreads, CMR_GATHER_MINB)
    k_tile_gather(int *bcnt, const unsigned *bbuf, int buckets, const int *hq, const void *pix, int pix16, const int *M,
                  const float *__restrict__ featT, const float *__restrict__ img_feat, int N, int ncap, int C, int P,
                  bool copy_image, bool vec, bool tma, float *__restrict__ obs2d,
                  const __grid_constant__ CUtensorMap map_proj, int share, long long out_estride, long long proj_off,
                  int tma_y0, int mean_channels, bool img_tma, const __grid_constant__ CUtensorMap map_img, int tail) {
    // heavy_from: buckets with more entries than this were queued for the bucket CTAs by the projecting kernel
    // (kLightLimit for an observe, kMidMax for a cost volume); up to it a warp does the bucket alone.
    constexpr int heavy_from = kMid ? kHeavyFrom : kLightLimit;
    // tail (0 or 4): the last four channels (a slab of their own, all of them SUMS) are not given a pass over the
    // rows by the light warps: a lane sums them for ITS pixel, in point order (the occupancy row of a cost volume).
    // share: consecutive episodes (poses) that look at the same cloud, i.e. the same feature rows.
    // Output of episode e: obs2d + e * out_estride + proj_off, rows of P floats; the tensor map's row of channel c
    // is tma_y0 + c.  Channels >= mean_channels are SUMS, not means (the occupancy row of a cost volume).
    extern __shared__ __align__(1024) float smem_g[];   // the tiles come first: 1024-byte aligned for the swizzle
    const int B = (int)gridDim.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int T = buckets;
    const int slabs = (C + kSlab - 1) / kSlab;
    int *hdr = bcnt + (size_t)B * kBucketStride;         // [0] heavy-queue length, [1] ticket of the bucket CTAs
    pdl_launch_dependents();
    DBG_MARK(0);
    // roles in launch order (x fastest, then y): bucket CTAs and light CTAs alternate for the first
    // 2 * kHeavyCtas rows, so that the (mostly idle) bucket CTAs do not fill the first wave alone
    const int y = (int)blockIdx.y;
    const bool bucket_role = y < 2 * kHeavyCtas && !(y & 1);
    const int role_idx = y < 2 * kHeavyCtas ? y >> 1 : y - kHeavyCtas;
    // The image half of obs2d (obs2d[b, 0:C] = img_geo_feat[b], environment.py:83) does not depend on the pose:
    // every light warp fetches the [64 channels][32 pixels] box of ITS bucket into its (still unused) result tile
    // before it waits for k_project, and sends it on to obs2d once it may write there.
    __shared__ __align__(8) uint64_t img_bar[kGatherWarps];
    const bool img_here = img_tma && !bucket_role && role_idx * kGatherWarps + warp < T;
    if (img_here && lane == 0) {
        mbar_init(&img_bar[warp], 1);
        fence_async_proxy();
        mbar_arrive_expect_tx(&img_bar[warp], (unsigned)(kTileFloats * sizeof(float)));
        tma_load_3d(smem_g + warp * kTileFloats, &map_img, (role_idx * kGatherWarps + warp) * kBucketPix, 0, blockIdx.x,
                    &img_bar[warp]);
    }
    pdl_wait();   // counters, bucket buffers and queue are written by k_project
    DBG_MARK(1);

    if (!bucket_role) {
        // ------------------------------------------------------------------ light units: one warp per bucket
        const int b = blockIdx.x;
        const int bk = role_idx * kGatherWarps + warp;
        float *tile = smem_g + warp * kTileFloats;
        constexpr int wkeys = warp_keys(heavy_from);
        unsigned *sk = reinterpret_cast<unsigned *>(smem_g + kGatherWarps * kTileFloats) + warp * wkeys;
        int *pc = reinterpret_cast<int *>(sk + (kMid ? wkeys - 64 : kLightMax));   // [32] points per pixel (mid: then [32] cursors)
        const int p0 = bk * kBucketPix;
        float *out = obs2d + (size_t)b * out_estride;
        float *proj = out + proj_off;
        const int bs = b / share;
        if (bk < T) {
            const unsigned *src = bbuf + ((size_t)b * buckets + bk) * kBucketCap;
            const unsigned e0 = ld_cg_u32(src + lane), e1 = ld_cg_u32(src + 32 + lane);   // speculative
            int *cb = bcnt + (size_t)b * kBucketStride + bk;
            const int n = ld_cg_s32(cb) & (kCountSeen - 1);
            if (lane == 0 && n != 0) {
                if (n <= heavy_from) *cb = 0;   // this warp is the counter's only reader
                else if (atomicAdd(cb, kCountSeen) >= kCountSeen) *cb = 0;   // the bucket CTA has been here
            }
            if (n > 0 && n <= heavy_from) {
                // all the rows this unit will add: on their way to L2 before the first one is needed
                if (lane < n)
                    for (int q = 0; q < C; q += 32) prefetch_l2(featT + ((size_t)bs * N + (e0 >> 7)) * C + q);
                if (lane + 32 < n)
                    for (int q = 0; q < C; q += 32) prefetch_l2(featT + ((size_t)bs * N + (e1 >> 7)) * C + q);
            }
#ifdef CMR_DBG_TIMING
            if (threadIdx.x == 0) g_dbg[(blockIdx.y * gridDim.x + blockIdx.x) * 16 + 7] = (unsigned long long)(n + (e0 & 0) + (e1 & 0));
            DBG_MARK(2);
#endif
            if (img_here) {
                // slabs beyond the first are fetched and forwarded one after the other (C > 64 only)
                for (int slab = 0; slab < slabs; ++slab) {
                    if (lane == 0) {
                        if (slab) {
                            mbar_arrive_expect_tx(&img_bar[warp], (unsigned)(kTileFloats * sizeof(float)));
                            tma_load_3d(tile, &map_img, p0, kSlab * slab, b, &img_bar[warp]);
                        }
                        mbar_wait(&img_bar[warp], slab & 1);
                        tma_store_3d(&map_proj, p0, kSlab * slab, b, tile);   // rows [0, C) of obs2d[b]: the image half
                        bulk_commit();
                        bulk_wait_read_all();
                    }
                }
                __syncwarp();
            } else if (copy_image) {   // image half when neither kernel can carry it as TMA traffic
                const float *img = img_feat + (size_t)b * C * P;
                for (int r = 0; r < C; ++r)
                    if (p0 + lane < P) out[(size_t)r * P + p0 + lane] = img[(size_t)r * P + p0 + lane];
            }
            if (n == 0) {
                if (tma) {   // zeros leave through the same door: the warp's own tile, cleared
                    zero_tile<32>(tile, lane);
                    fence_async_proxy();
                    __syncwarp();
                    if (lane == 0)
                        for (int slab = 0; slab < slabs; ++slab) store_tile_tma(tile, &map_proj, tma_y0 + kSlab * slab, p0, b);
                    __syncwarp();
                } else {
                    store_zeros(C, p0, P, vec, proj, lane);
                }
            } else if (n > 0 && n <= kLightLimit) {   // otherwise the bucket CTAs own the bucket
                // the tail channels of this unit's rows: on their way while the entries are ranked
                float4 t0 = make_float4(0.f, 0.f, 0.f, 0.f), t1 = t0;
                if (tail) {
                    const float *tb = featT + (size_t)bs * N * C + (C - 4);
                    if (lane < n) t0 = ldg_f4(tb + (size_t)(e0 >> 7) * C);
                    if (lane + 32 < n) t1 = ldg_f4(tb + (size_t)(e1 >> 7) * C);
                }
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
                if (lane < n) sk[r0] = k0;
                if (lane + 32 < n) sk[r1] = k1;
                pc[lane] = 0;
                __syncwarp();
                // flags: first / last entry of its pixel; points per pixel
                {
                    const unsigned a = lane < n ? sk[lane] : 0u, a2 = lane + 32 < n ? sk[lane + 32] : 0u;
                    const unsigned ap = lane > 0 ? sk[lane - 1] : 0xffffffffu, an = lane + 1 < n ? sk[lane + 1] : 0xffffffffu;
                    const unsigned a2p = sk[lane + 31], a2n = lane + 33 < n ? sk[lane + 33] : 0xffffffffu;
                    if (lane < n) atomicAdd(&pc[a >> 24], 1);
                    if (lane + 32 < n) atomicAdd(&pc[a2 >> 24], 1);
                    __syncwarp();
                    if (lane < n)
                        sk[lane] = a | ((a >> 24) != (ap >> 24) ? kKeyFirst : 0u) | ((a >> 24) != (an >> 24) ? kKeyLast : 0u);
                    if (lane + 32 < n)
                        sk[lane + 32] = a2 | ((a2 >> 24) != (a2p >> 24) ? kKeyFirst : 0u) | ((a2 >> 24) != (a2n >> 24) ? kKeyLast : 0u);
                }
                __syncwarp();
                const int my_cnt = pc[lane];
                const unsigned multi = __ballot_sync(kFull, my_cnt > 1);
                const LaneRows tl = lane_rows(tile, lane);
                DBG_MARK(8);
                if (tail) {
                    // sorted entries are grouped by pixel, in point order: lane = pixel walks its own group
                    float4 *tv = reinterpret_cast<float4 *>(tile);   // the tile is idle until the first slab is zeroed
                    if (lane < n) tv[r0] = t0;
                    if (lane + 32 < n) tv[r1] = t1;
                    int incl = my_cnt;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const int up = __shfl_up_sync(kFull, incl, o);
                        if (lane >= o) incl += up;
                    }
                    const int start = incl - my_cnt;
                    __syncwarp();
                    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
                    for (int i = 0; i < my_cnt; ++i) {
                        const float4 v = tv[start + i];
                        acc.x = __fadd_rn(acc.x, v.x);
                        acc.y = __fadd_rn(acc.y, v.y);
                        acc.z = __fadd_rn(acc.z, v.z);
                        acc.w = __fadd_rn(acc.w, v.w);
                    }
                    __syncwarp();
                    if (p0 + lane < P) {
                        float *dst = proj + (size_t)(C - 4) * P + p0 + lane;
                        dst[0] = acc.x;
                        dst[(size_t)P] = acc.y;
                        dst[2 * (size_t)P] = acc.z;
                        dst[3 * (size_t)P] = acc.w;
                    }
                }
                for (int slab = 0; slab < (tail ? slabs - 1 : slabs); ++slab) {
                    const int c0 = kSlab * slab;
                    // lanes beyond C read channel 0 instead (their rows of the tile are never stored)
                    const float *rows = featT + (size_t)bs * N * C + (c0 + 2 * lane < C ? c0 + 2 * lane : 0);
                    zero_tile<32>(tile, lane);
                    __syncwarp();
                    add_rows<false>(sk, 0, n, rows, (unsigned)C, tl);
                    if (c0 < mean_channels) mean_pass(multi, my_cnt, tl);
                    DBG_MARK(3);
                    if (tma) {
                        fence_async_proxy();
                        __syncwarp();
                        if (lane == 0) store_tile_tma(tile, &map_proj, tma_y0 + c0, p0, b);
                    } else {
                        __syncwarp();
                        store_tile<32>(tile, c0, C, p0, P, proj, lane);
                    }
                    __syncwarp();
                }
            } else if (kMid && n > kLightLimit && n <= heavy_from) {
                // ---------------------------------------------------------- mid unit: up to kMidMax entries, still one warp
                // Entries lane + 32 i.  The order (pixel, point) is established by binning: points per pixel (shared-
                // memory atomics), an exclusive scan across the lanes, every entry dropped into its pixel's segment in
                // arrival order, then ranked against the (few) entries of its own segment only.  The flags fall out of
                // the rank.  The tile - idle until the first slab is zeroed - holds the binned keys and the tail rows.
                unsigned key[kMidPer];
                key[0] = ((e0 & 31u) << 24) | (e0 >> 7);
                key[1] = lane + 32 < n ? ((e1 & 31u) << 24) | (e1 >> 7) : 0xffffffffu;
#pragma unroll
                for (int i = 2; i < kMidPer; ++i) {
                    const unsigned e = lane + 32 * i < n ? ld_cg_u32(src + lane + 32 * i) : 0xffffffffu;
                    key[i] = e == 0xffffffffu ? e : ((e & 31u) << 24) | (e >> 7);
                }
#pragma unroll
                for (int i = 2; i < kMidPer; ++i)
                    if (lane + 32 * i < n)
                        for (int q = 0; q < C; q += 32) prefetch_l2(featT + ((size_t)bs * N + (key[i] & 0xffffffu)) * C + q);
                int *cur = pc + 32;
                unsigned *binned = reinterpret_cast<unsigned *>(tile) + kTileFloats - kMidMax;   // the tile's last kMidMax words
                pc[lane] = 0;
                cur[lane] = 0;
                __syncwarp();
#pragma unroll
                for (int i = 0; i < kMidPer; ++i)
                    if (lane + 32 * i < n) atomicAdd(&pc[key[i] >> 24], 1);
                __syncwarp();
                const int my_cnt = pc[lane];
                int incl = my_cnt;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int up = __shfl_up_sync(kFull, incl, o);
                    if (lane >= o) incl += up;
                }
                const int start = incl - my_cnt;
                int seg[kMidPer];   // where the entry's pixel starts in the sorted list
#pragma unroll
                for (int i = 0; i < kMidPer; ++i) {
                    const bool ok = lane + 32 * i < n;
                    const int px = ok ? (int)(key[i] >> 24) : 0;
                    seg[i] = __shfl_sync(kFull, start, px);
                    if (ok) binned[seg[i] + atomicAdd(&cur[px], 1)] = key[i];
                }
                __syncwarp();
                int pos[kMidPer];
#pragma unroll
                for (int i = 0; i < kMidPer; ++i) {
                    const bool ok = lane + 32 * i < n;
                    const int c = ok ? pc[key[i] >> 24] : 0;
                    int r = 0;
                    for (int j = 0; j < c; ++j) r += binned[seg[i] + j] < key[i];
                    pos[i] = seg[i] + r;
                    if (ok) sk[pos[i]] = key[i] | (r == 0 ? kKeyFirst : 0u) | (r == c - 1 ? kKeyLast : 0u);
                }
                __syncwarp();
                const unsigned multi = __ballot_sync(kFull, my_cnt > 1);
                const LaneRows tl = lane_rows(tile, lane);
                if (tail) {
                    // the four summed channels: every entry's values go to its sorted place, lane = pixel adds its own group
                    float4 *tv = reinterpret_cast<float4 *>(tile);   // n * 16 bytes <= half of the tile (binned sits in its end)
                    const float *tb = featT + (size_t)bs * N * C + (C - 4);
#pragma unroll
                    for (int h = 0; h < kMidPer; h += 4) {
                        float4 t[4];
#pragma unroll
                        for (int i = 0; i < 4; ++i)
                            if (lane + 32 * (h + i) < n) t[i] = ldg_f4(tb + (size_t)(key[h + i] & 0xffffffu) * C);
#pragma unroll
                        for (int i = 0; i < 4; ++i)
                            if (lane + 32 * (h + i) < n) tv[pos[h + i]] = t[i];
                    }
                    __syncwarp();
                    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
                    for (int i = 0; i < my_cnt; ++i) {
                        const float4 v = tv[start + i];
                        acc.x = __fadd_rn(acc.x, v.x);
                        acc.y = __fadd_rn(acc.y, v.y);
                        acc.z = __fadd_rn(acc.z, v.z);
                        acc.w = __fadd_rn(acc.w, v.w);
                    }
                    __syncwarp();
                    if (p0 + lane < P) {
                        float *dst = proj + (size_t)(C - 4) * P + p0 + lane;
                        dst[0] = acc.x;
                        dst[(size_t)P] = acc.y;
                        dst[2 * (size_t)P] = acc.z;
                        dst[3 * (size_t)P] = acc.w;
                    }
                }
                for (int slab = 0; slab < (tail ? slabs - 1 : slabs); ++slab) {
                    const int c0 = kSlab * slab;
                    const float *rows = featT + (size_t)bs * N * C + (c0 + 2 * lane < C ? c0 + 2 * lane : 0);
                    zero_tile<32>(tile, lane);
                    __syncwarp();
                    add_rows<false>(sk, 0, n, rows, (unsigned)C, tl);
                    if (c0 < mean_channels) mean_pass(multi, my_cnt, tl);
                    if (tma) {
                        fence_async_proxy();
                        __syncwarp();
                        if (lane == 0) store_tile_tma(tile, &map_proj, tma_y0 + c0, p0, b);
                    } else {
                        __syncwarp();
                        store_tile<32>(tile, c0, C, p0, P, proj, lane);
                    }
                    __syncwarp();
                }
            }
        }
        DBG_MARK(4);
        DBG_MARK(5);
    } else {
        // ------------------------------------------------------------------ bucket CTA
        float *tile = smem_g;                                             // [64][32] swizzled: sums, then means
        unsigned *slist = reinterpret_cast<unsigned *>(tile + kTileFloats);   // [kBucketCap] keys, sorted
        int *pcnt = reinterpret_cast<int *>(slist + kBucketCap);          // [32] entries per pixel of this chunk
        int *pstart = pcnt + 32;                                          // [33] exclusive prefix
        int *ptotal = pstart + 33;                                        // [32] entries per pixel, all chunks
        int *misc = ptotal + 32;                                          // [8]
        unsigned *ent = reinterpret_cast<unsigned *>(smem_g + kGatherSmemHeavyLive / sizeof(float));   // [kBucketCap] pixel << 24 | point, as they arrive
        unsigned *ulist = ent + kBucketCap;                               // [kBucketCap] points, grouped by pixel
        // [kStageRows][kSlab] staged feature rows: over ent and ulist, which are dead once slist is complete
        float *stage = reinterpret_cast<float *>(ent);
        const int first_item = (int)(role_idx * gridDim.x + blockIdx.x);
        int qe = ld_cg_s32(hq + first_item);   // speculative: valid iff first_item < items
        const int items = min(ld_cg_s32(hdr), B * kBucketMaxBuckets);
#ifdef CMR_DBG_TIMING
        if (threadIdx.x == 0) {
            g_dbg[(blockIdx.y * gridDim.x + blockIdx.x) * 16 + 7] = 0;
            g_dbg[(blockIdx.y * gridDim.x + blockIdx.x) * 16 + 6] = 0;
        }
#endif
        for (int item = first_item; item < items; item += kHeavyCtas * B) {
            if (item != first_item) qe = ld_cg_s32(hq + item);
            const int b = (int)((unsigned)qe >> 16), bk = qe & 0xffff;   // episodes up to 65535: the entry is unsigned
            int *cb = bcnt + (size_t)b * kBucketStride + bk;
            const int c = ld_cg_s32(cb) & (kCountSeen - 1);
            const int p0 = bk * kBucketPix;
            float *proj = obs2d + (size_t)b * out_estride + proj_off;
            const int bs = b / share;
            const bool chunked = c < 0 || c > kBucketCap;
            DBG_MARK(2);
            if (tid < 32) {
                pcnt[tid] = 0;
                ptotal[tid] = 0;
            }
            __syncthreads();
            // orders the n entries of `ent` (counted per pixel in pcnt): slist = keys sorted by (pixel, point),
            // pstart = where every pixel's run begins, ptotal += points per pixel
            auto order_chunk = [&](int n) {
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
                    pcnt[lane] = 0;   // reused as the fill level while placing
                }
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
#pragma unroll 8
                    for (int j = s0; j < s1; ++j) r += ulist[j] < pt;
                    slist[s0 + r] = e;
                }
                if (tid < 32) pcnt[tid] = 0;   // ready for the next chunk's counts
                __syncthreads();
            };
            // adds the rows of slist[0, n) for one slab.  The sorted list is split at pixel boundaries into eight
            // nearly equal parts, warp w takes the pixels whose first entry lies in [w * n / 8, (w + 1) * n / 8).
            // The rows come through shared memory: the whole CTA fetches kStageRows of them at a time with
            // cp.async, in sorted order - ONE round trip for a typical bucket instead of one per batch of a warp -
            // and a pixel's rows are then consecutive: the inner loop is a load and two additions per row.
            auto add_chunk = [&](int n, int c0) {
                const float *rbase = featT + (size_t)bs * N * C + c0;
                const int pieces = min(kSlab, C - c0) >> 2;   // 16-byte pieces of a row in this slab
                const int ps = pstart[lane];
                const int lo = (warp * n + kGatherWarps - 1) / kGatherWarps, hi = ((warp + 1) * n + kGatherWarps - 1) / kGatherWarps;
                const int pb = __popc(__ballot_sync(kFull, ps < lo));
                const int pe = warp == kGatherWarps - 1 ? 32 : __popc(__ballot_sync(kFull, ps < hi));
                const LaneRows tl = lane_rows(tile, lane);
                for (int r0 = 0; r0 < n; r0 += kStageRows) {
                    const int nr = min(kStageRows, n - r0);
                    for (int i = tid; i < nr * (kSlab / 4); i += kGatherThreads) {
                        const int r = i / (kSlab / 4), piece = i % (kSlab / 4);
                        if (piece < pieces)
                            cp_async16(stage + r * kSlab + 4 * piece, rbase + (size_t)(slist[r0 + r] & 0xffffffu) * C + 4 * piece);
                    }
                    cp_async_wait_all();
                    __syncthreads();
#ifdef CMR_DBG_TIMING
                    if (r0 == 0) DBG_MARK(12);
#endif
                    for (int pl = pb; pl < pe; ++pl) {   // this warp's pixels; their sums so far are in the tile
                        const int js = max(pstart[pl], r0) - r0, jn = min(pstart[pl + 1], r0 + nr) - r0;
                        if (js >= jn) continue;          // warp-uniform
                        float2 acc = tl.get(pl);
                        const float *src = stage + js * kSlab + 2 * lane;
#pragma unroll 8
                        for (int j = js; j < jn; ++j, src += kSlab) {
                            const float2 v = *reinterpret_cast<const float2 *>(src);
                            acc.x = __fadd_rn(acc.x, v.x);
                            acc.y = __fadd_rn(acc.y, v.y);
                        }
                        tl.put(pl, acc);
                    }
#ifdef CMR_DBG_TIMING
                    if (r0 == 0) DBG_MARK(13);
#endif
                    __syncthreads();   // the stage is refilled / the tile is read
#ifdef CMR_DBG_TIMING
                    if (r0 == 0) DBG_MARK(14);
                    if (threadIdx.x == 0) g_dbg[(blockIdx.y * gridDim.x + blockIdx.x) * 16 + 15] = (unsigned long long)(pstart[pe < 32 ? pe : 32] - pstart[pb < 32 ? pb : 32]);
#endif
                }
            };
            auto finish_slab = [&](int c0) {
                __syncthreads();
                // sums -> means: warp w takes the pixels p % 8 == w
                const int my_cnt = ptotal[lane];
                const unsigned multi = __ballot_sync(kFull, my_cnt > 1) & (0x01010101u << warp);
                if (c0 < mean_channels) mean_pass(multi, my_cnt, lane_rows(tile, lane));
                if (tma) {
                    fence_async_proxy();
                    __syncthreads();
                    if (tid == 0) store_tile_tma(tile, &map_proj, tma_y0 + c0, p0, b);
                } else {
                    __syncthreads();
                    store_tile<kGatherThreads>(tile, c0, C, p0, P, proj, tid);
                }
                __syncthreads();
            };
            if (!chunked) {
                const unsigned *src = bbuf + ((size_t)b * buckets + bk) * kBucketCap;
                for (int i = tid; i < c; i += kGatherThreads) {
                    const unsigned e = ld_cg_u32(src + i);
                    const unsigned pl = e & 31u;
                    for (int q = 0; q < C; q += 32) prefetch_l2(featT + ((size_t)bs * N + (e >> 7)) * C + q);
                    ent[i] = (pl << 24) | (e >> 7);
                    atomicAdd(&pcnt[pl], 1);
                }
                DBG_MARK(8);
                order_chunk(c);
                DBG_MARK(9);
                for (int slab = 0; slab < slabs; ++slab) {
                    zero_tile<kGatherThreads>(tile, tid);
                    __syncthreads();
                    add_chunk(c, kSlab * slab);
                    DBG_MARK(10);
                    finish_slab(kSlab * slab);
                }
            } else {
                // the bucket overflowed its buffer: rebuild it from the episode's pixel-id list, kBucketCap
                // points at a time, in point order (a chunk's points all precede the next chunk's); per slab
                const int m_total = min(ld_cg_s32(M + bs), N);
                for (int slab = 0; slab < slabs; ++slab) {
                    zero_tile<kGatherThreads>(tile, tid);
                    if (tid < 32) ptotal[tid] = 0;
                    __syncthreads();
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
                            order_chunk(fill);
                            add_chunk(fill, kSlab * slab);
                            fill = 0;
                            __syncthreads();
                        }
                        if (hit) {
                            const unsigned pl = (unsigned)(id - p0);
                            ent[fill + before + __popc(bal & ((1u << lane) - 1))] = (pl << 24) | (unsigned)m;
                            atomicAdd(&pcnt[pl], 1);
                        }
                        fill += total;
                        __syncthreads();
                    }
                    if (fill > 0) {
                        order_chunk(fill);
                        add_chunk(fill, kSlab * slab);
                    }
                    finish_slab(kSlab * slab);
                }
            }
            // (barriers since: every thread has read the counter) the second of its two readers clears it
            if (tid == 0 && atomicAdd(cb, kCountSeen) >= kCountSeen) *cb = 0;
            DBG_MARK(11);
#ifdef CMR_DBG_TIMING
            if (threadIdx.x == 0) {
                g_dbg[(blockIdx.y * gridDim.x + blockIdx.x) * 16 + 7] += (unsigned long long)(c < 0 ? 0 : c);
                g_dbg[(blockIdx.y * gridDim.x + blockIdx.x) * 16 + 6] += 1;
            }
#endif
        }
        DBG_MARK(4);
        DBG_MARK(5);
    }
    // the last bucket CTA to get here clears the queue length for the next observe (they are its only readers)
    if (bucket_role && tid == 0) {
        __threadfence();
        if (atomicAdd(hdr + 1, 1) == kHeavyCtas * B - 1) {
            hdr[0] = 0;
            hdr[1] = 0;
        }
    }
}

}  // namespace cmr
