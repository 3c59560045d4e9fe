// cost_volume_kernels.cuh - the pose-sampling cost volume of models/IterModel.py:272-351 for the reference's own shape
// (64 mean channels + the summed score, a grid of whole 32-pixel buckets): one counting sort per pose, one gather.
//
// The observation's kernels (k_project_masked + k_tile_gather, env_kernels.cuh / scatter_kernels.cuh) hand points to
// 32-pixel buckets through global atomics and per-bucket buffers and then restore the point order bucket by bucket.
// That is built for a view in which a bucket holds a dozen points.  A cost volume is the other regime: hundreds of
// candidate poses look at the SAME masked points (8956 of a KITTI cloud), two thirds of them land inside the image,
// 4.3 per occupied pixel, and everything a pose needs to order them fits one CTA's shared memory.
//   k_xyz_compact         once per cloud: coordinates of the masked points, in index order
//   k_cost_volume_sort    one CTA per pose.  Project (:281-304, :316-318), points per pixel -> shared-memory histogram,
//                         exclusive scan over the H*W pixels, counting-sort fill in arrival order, then every point is
//                         ranked against the (few) points of its own pixel: `sorted` = the visible points in (pixel,
//                         point) order - the order torch's CPU scatter adds them in.  A pixel with more than 256
//                         points is ordered through a bitmap of the cloud instead (rank = marked points before it).
//   k_cost_volume_gather  one warp per 32-pixel bucket: its rows are a contiguous piece of `sorted`.  A feature row is a
//                         256-byte warp load (two channels per lane, the score on lane 0); a pixel's rows are added in
//                         point order; a finished pixel leaves as its mean (scores: sum) into a pixel-major tile
//                         [32][64] (XOR-swizzled: conflict-free both ways), and the tile is written channel-major
//                         as full 128-byte lines.  (:341-343)
// No global atomics, no bucket buffers; 240 M warp instructions for the 729 poses of a KITTI cloud instead of 350 M.
// Results are bit-identical to the bucket path (same projection code, same sequential sums, same division).
// Measured on a B200: 1.03 ms (round 1) -> 0.73 ms (mid units in the bucket path) -> 0.48 ms.
#pragma once
#include "common.cuh"
#include "env_kernels.cuh"
#include "scatter_kernels.cuh"

namespace cmr {

constexpr int kCvThreads = 256;
constexpr int kCvWarps = kCvThreads / 32;
constexpr int kCvSortThreads = 512;    // k_cost_volume_sort: 18 points per thread for a KITTI cloud, two CTAs per SM
constexpr int kCvSortWarps = kCvSortThreads / 32;
constexpr int kCvMaxP = 8192;          // pixels: what phase A keeps per pixel has to fit its scratch
constexpr int kCvFeat = 64;            // mean channels (the reference's embed width); channel 64 = the summed score
constexpr int kCvLong = 256;           // a pixel with more points than this is ordered through a bitmap of the cloud
constexpr int kCvMaxLong = 65536 / kCvLong;   // ... and there are at most this many of them (ncap <= 65535)
constexpr int kCvMaxWords = 65536 / 32;       // bitmap words
constexpr int kCvBitmapBytes = kCvMaxWords * 4 + kCvMaxWords * 2;   // bitmap + its word prefix (u16)
// Dynamic shared memory: 73 KB, so that three CTAs share an SM.  Phase B: eight [32][64] fp32 tiles (64 KB).  Phase A,
// in the same bytes: histogram/cursors [P] u32 | offsets [P + 2] u16 | pixel ids [Mb] u16 | arrival-order list [Mb] u16
// when the cloud's Mb masked points leave room for the last two (a KITTI cloud's 8956 do; otherwise they live in global
// memory); the bitmap of the long pixels lies over the pixel ids, which are dead by then.
constexpr size_t kCvSmem = 74752;                                                  // k_cost_volume_sort

static_assert((size_t)kCvMaxP * 4 + ((size_t)kCvMaxP + 2) * 2 + 4 + kCvBitmapBytes <= kCvSmem, "phase A's fixed part fits");

__device__ __forceinline__ unsigned ld_cg_u16(const uint16_t *p) {
    unsigned short v;
    asm volatile("ld.global.cg.u16 %0, [%1];" : "=h"(v) : "l"(p) : "memory");
    return v;
}

// Once per cloud: coordinates of the masked points in index order (position = the row of featT the point's features
// were compacted to).  One warp = one 128-point group, as k_project_masked lists them.
__global__ void __launch_bounds__(256) k_xyz_compact(const float *__restrict__ pc, const uint8_t *__restrict__ mask,
                                                     const int *__restrict__ seg, int N, int ncap, int groups, bool vec,
                                                     float4 *__restrict__ xyzc) {
    const int bs = blockIdx.y;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = blockIdx.x * 8 + warp;
    if (g >= groups) return;
    const int j0 = g * kGroup + lane * 4;
    const unsigned flags = load_flags4(mask + (size_t)bs * N, j0, N, vec);
    const int mine = __popc(flags);
    int incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(kFull, incl, o);
        if (lane >= o) incl += t;
    }
    int pos = __ldg(seg + (size_t)bs * groups + g) + incl - mine;
    const float *px = pc + (size_t)bs * 3 * N, *py = px + N, *pz = py + N;
#pragma unroll
    for (int i = 0; i < 4; ++i)
        if (flags >> i & 1) xyzc[(size_t)bs * ncap + pos++] = make_float4(__ldg(px + j0 + i), __ldg(py + j0 + i), __ldg(pz + j0 + i), 0.f);
}

__device__ __forceinline__ float lds_f32(unsigned addr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
    return v;
}

// column of channel c in pixel px's row of the tile
__device__ __forceinline__ int cv_col(int c, int px) { return c ^ px; }

// Phase A: one CTA per pose.
__global__ void __launch_bounds__(kCvSortThreads, 2)
    k_cost_volume_sort(const float4 *__restrict__ xyzc, const int *__restrict__ M, const float *__restrict__ Kmat,
                       const float *__restrict__ poses, const float *__restrict__ zero_mean, int ncap, int H, int W, int share,
                       bool chain, uint16_t *__restrict__ pix, uint16_t *__restrict__ tmp, uint16_t *__restrict__ sorted,
                       uint16_t *__restrict__ gstart) {
    pdl_launch_dependents();
    extern __shared__ __align__(1024) unsigned char cv_smem[];   // a tile row (256 bytes) must not straddle its alignment: see the output pass
    const int P = H * W;
    __shared__ unsigned wsum[kCvSortWarps];
    __shared__ int n_long;
    __shared__ uint16_t long_px[kCvMaxLong];
    const int e = blockIdx.x, bs = e / share;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int Mb = min(__ldg(M + bs), ncap);
    const float4 *pts = xyzc + (size_t)bs * ncap;
    uint16_t *my_sorted = sorted + (size_t)e * ncap;
    uint16_t *my_start = gstart + (size_t)e * (P + 2);
    // phase A's scratch (see kCvSmem)
    unsigned *hist = reinterpret_cast<unsigned *>(cv_smem);                    // [P] counts, then cursors
    uint16_t *start = reinterpret_cast<uint16_t *>(cv_smem + (size_t)P * 4);   // [P + 1] exclusive prefix
    const size_t o_ids = ((size_t)P * 4 + ((size_t)P + 2) * 2 + 15) / 16 * 16;
    const size_t ids_bytes = max(((size_t)Mb * 2 + 15) / 16 * 16, (size_t)kCvBitmapBytes);
    const bool in_smem = o_ids + ids_bytes + (size_t)Mb * 2 <= kCvSmem;
    uint16_t *ids_s = reinterpret_cast<uint16_t *>(cv_smem + o_ids), *ids_g = pix + (size_t)e * ncap;
    uint16_t *arr = in_smem ? reinterpret_cast<uint16_t *>(cv_smem + o_ids + ids_bytes) : tmp + (size_t)e * ncap;   // (long pixels only)
    unsigned *bitmap = reinterpret_cast<unsigned *>(cv_smem + o_ids);          // [kCvMaxWords], over the ids
    uint16_t *wpre = reinterpret_cast<uint16_t *>(bitmap + kCvMaxWords);       // [kCvMaxWords]

    // ---- A1: project, count per pixel
    for (int p = tid; p < P; p += kCvSortThreads) hist[p] = 0;
    if (tid == 0) n_long = 0;
    PoseK s;
    load_posek(s, poses, Kmat, zero_mean, e, bs);
    const float wmax = (float)(W - 1), hmax = (float)(H - 1);
    __syncthreads();
    for (int k = tid; k < Mb; k += kCvSortThreads) {
        const float4 q = __ldg(pts + k);
        bool in_cam;
        const int id = chain ? project_point<true>(s, q.x, q.y, q.z, wmax, hmax, W, P, in_cam)
                             : project_point<false>(s, q.x, q.y, q.z, wmax, hmax, W, P, in_cam);
        if (in_smem) ids_s[k] = (uint16_t)id;   // P marks a point outside the frustum (:318)
        else ids_g[k] = (uint16_t)id;
        if (in_cam) atomicAdd(&hist[id], 1u);
    }
    __syncthreads();

    // ---- A2: exclusive scan over the pixels (a contiguous chunk per thread)
    {
        const int chunk = (P + kCvSortThreads - 1) / kCvSortThreads;
        const int pb = tid * chunk, pe = min(P, pb + chunk);
        unsigned local = 0;
        for (int p = pb; p < pe; ++p) local += hist[p];
        unsigned incl = local;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned t = __shfl_up_sync(kFull, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) wsum[warp] = incl;
        __syncthreads();
        unsigned base = incl - local;
        for (int w = 0; w < warp; ++w) base += wsum[w];
        for (int p = pb; p < pe; ++p) {
            const unsigned c = hist[p];
            if (c > (unsigned)kCvLong) long_px[atomicAdd(&n_long, 1)] = (uint16_t)p;
            hist[p] = base;   // from here on: the pixel's cursor
            start[p] = (uint16_t)base;
            my_start[p] = (uint16_t)base;
            base += c;
        }
        if (tid == kCvSortThreads - 1) {   // its chunk ends at P (or is empty): base is the number of visible points
            start[P] = (uint16_t)base;
            my_start[P] = (uint16_t)base;
        }
    }
    __syncthreads();

    // ---- A3: fill in arrival order, then rank inside the pixel (global copies: a thread re-reads the ids it wrote itself).
    // Instantiated per address space: through a generic pointer every access is a slower generic load.
    auto order_points = [&](const uint16_t *ids_p, uint16_t *arr_p, auto load) {
        for (int k = tid; k < Mb; k += kCvSortThreads) {
            const int id = ids_p[k];
            if (id < P) arr_p[atomicAdd(&hist[id], 1u)] = (uint16_t)k;
        }
        __syncthreads();
        for (int k = tid; k < Mb; k += kCvSortThreads) {
            const int id = ids_p[k];
            if (id < P) {
                const int sb = start[id], c = start[id + 1] - sb;
                if (c > kCvLong) continue;   // ordered below
                int r = 0, j = 0;
                for (; j + 4 <= c; j += 4)   // (independent loads: four in flight)
                    r += (load(arr_p + sb + j) < (unsigned)k) + (load(arr_p + sb + j + 1) < (unsigned)k) +
                         (load(arr_p + sb + j + 2) < (unsigned)k) + (load(arr_p + sb + j + 3) < (unsigned)k);
                for (; j < c; ++j) r += load(arr_p + sb + j) < (unsigned)k;
                my_sorted[sb + r] = (uint16_t)k;
            }
        }
    };
    if (in_smem)
        order_points(ids_s, reinterpret_cast<uint16_t *>(cv_smem + o_ids + ids_bytes), [](const uint16_t *q) { return (unsigned)*q; });
    else
        order_points(ids_g, tmp + (size_t)e * ncap, [](const uint16_t *q) { return ld_cg_u16(q); });
    // Pixels that hold more than kCvLong points (a cloud seen from far away collapses onto a few pixels): ranking every
    // point against its whole segment would be quadratic.  The CTA marks the segment's points in a bitmap of the cloud;
    // a point's rank is the number of marked points before it (word prefix + popcount).
    if (n_long > 0) __syncthreads();   // (uniform) the bitmap lies over the pixel ids: everybody is done with them
    for (int li = 0; li < n_long; ++li) {   // uniform
        const int lp = long_px[li];
        const int sb = start[lp], c = start[lp + 1] - sb;
        const int words = (Mb + 31) >> 5;
        for (int w = tid; w < words; w += kCvSortThreads) bitmap[w] = 0;
        __syncthreads();
        for (int j = tid; j < c; j += kCvSortThreads) {
            const unsigned k = in_smem ? (unsigned)arr[sb + j] : ld_cg_u16(arr + sb + j);
            atomicOr(&bitmap[k >> 5], 1u << (k & 31));
        }
        __syncthreads();
        {
            const int chunk = (words + kCvSortThreads - 1) / kCvSortThreads;
            const int wb = tid * chunk, we = min(words, wb + chunk);
            unsigned local = 0;
            for (int w = wb; w < we; ++w) local += __popc(bitmap[w]);
            unsigned incl = local;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned t = __shfl_up_sync(kFull, incl, o);
                if (lane >= o) incl += t;
            }
            if (lane == 31) wsum[warp] = incl;
            __syncthreads();
            unsigned base = incl - local;
            for (int w = 0; w < warp; ++w) base += wsum[w];
            for (int w = wb; w < we; ++w) {
                wpre[w] = (uint16_t)base;
                base += __popc(bitmap[w]);
            }
        }
        __syncthreads();
        for (int j = tid; j < c; j += kCvSortThreads) {
            const unsigned k = in_smem ? (unsigned)arr[sb + j] : ld_cg_u16(arr + sb + j);
            const int r = wpre[k >> 5] + __popc(bitmap[k >> 5] & ((1u << (k & 31)) - 1u));
            my_sorted[sb + r] = (uint16_t)k;
        }
        __syncthreads();
    }
}

// Phase B: a CTA = eight warps = eight consecutive 32-pixel buckets of one pose (blockIdx.y).
// A feature row is a 256-byte warp load (+ the score on lane 0), eight rows in flight per warp; the results leave as
// full 128-byte lines.  ncu: L1/LSU pipe 55-76 % busy, issue slots 51 %, 28 % of the warp slots.  Tried and dropped:
// two alternating batches of eight rows (364 vs 335 us), 16-pixel buckets (4 KB tiles,
// 32 warps per SM instead of 24: the same 335 us - half-line stores cost the pipe what the extra warps gained) and
// scores delivered in sorted order by k_cost_volume_sort (fewer L1 requests, but the sort's gather and the extra
// shuffle per row cost more: 0.49 -> 0.58 ms).
constexpr int kCvPerWarp = 1;           // consecutive buckets per warp of k_cost_volume_gather (4, with the next bucket's
                                       // offsets and points prefetched, measured no faster: 358 vs 335 us)
constexpr size_t kCvTileSmem = (size_t)kCvWarps * 32 * kCvFeat * sizeof(float);
__global__ void __launch_bounds__(kCvThreads, 4)   // (64 registers; shared memory admits three CTAs)
    k_cost_volume_gather(const float *__restrict__ featT, int N, int ncap, int Cx, int P, int share, const uint16_t *__restrict__ sorted,
                         const uint16_t *__restrict__ gstart, float *__restrict__ out) {
    // tile [32 pixels][64 channels] fp32 per warp; channel c of pixel px sits in column c ^ px of the pixel's 256-byte row:
    // conflict-free for the writer (a pixel's 64 means = one row) and for the transposed read of the output pass
    extern __shared__ __align__(1024) unsigned char cv_smem[];
    __shared__ float occ_s[kCvWarps][32];
    const int e = blockIdx.y, bs = e / share;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint16_t *my_sorted = sorted + (size_t)e * ncap;
    const uint16_t *my_start = gstart + (size_t)e * (P + 2);
    float *pm = reinterpret_cast<float *>(cv_smem) + (size_t)warp * 32 * kCvFeat;
    const char *rows_l = reinterpret_cast<const char *>(featT + (size_t)bs * N * Cx + 2 * lane);   // this lane's channel pair of row 0
    const unsigned row_bytes = (unsigned)Cx * 4u;
    float *oute = out + (size_t)e * Cx * P;
    const int nb = P / 32;
    const int q4 = (lane & 7) * 4, g4 = lane >> 3;   // the output pass: this lane's four pixels and its channel of four
    // A warp takes kCvPerWarp CONSECUTIVE buckets: the sorted list runs on from one bucket into the next, so while a
    // bucket is being added up the offsets and the first points of the next one are already on their way (three
    // dependent L2 round trips - offsets, points, rows - stood at the start of every bucket otherwise).
    const int bk0 = (blockIdx.x * kCvWarps + warp) * kCvPerWarp, bk1 = min(nb, bk0 + kCvPerWarp);
    pdl_wait();   // the sorted lists come from k_cost_volume_sort
    int st = 0, st_hi = 0;
    unsigned first = 0;
    if (bk0 < nb) {
        st = (int)ld_cg_u16(my_start + bk0 * 32 + lane);
        st_hi = (int)ld_cg_u16(my_start + bk0 * 32 + lane + 1);
        const int j0 = __shfl_sync(kFull, st, 0);
        if (lane < 8) first = ld_cg_u16(my_sorted + j0 + lane);   // (may run past the bucket: masked when used)
    }
    for (int bk = bk0; bk < bk1; ++bk) {
        const int p0 = bk * 32;
        const int cnt = st_hi - st;
        unsigned occ = __ballot_sync(kFull, cnt > 0);
        const unsigned occ_all = occ;
        const int jb = __shfl_sync(kFull, st, 0), je = __shfl_sync(kFull, st_hi, 31);
        // the next bucket: its offsets, and its first points (they follow this bucket's in the list)
        int nst = 0, nst_hi = 0;
        unsigned nfirst = 0;
        if (bk + 1 < bk1) {
            nst = (int)ld_cg_u16(my_start + p0 + 32 + lane);
            nst_hi = (int)ld_cg_u16(my_start + p0 + 32 + lane + 1);
            if (lane < 8) nfirst = ld_cg_u16(my_sorted + je + lane);
        }
        if (occ) {
            float2 acc = make_float2(0.f, 0.f);
            float sacc = 0.f;
            int px = __ffs(occ) - 1;
            int run = __shfl_sync(kFull, cnt, px), left = run;
            unsigned mine = (lane < 8 && jb + lane < je) ? first : 0u;
            for (int j = jb; j < je; j += 8) {
                // the next batch's points are fetched while this batch's rows are on their way
                const unsigned nxt = (lane < 8 && j + 8 + lane < je) ? ld_cg_u16(my_sorted + j + 8 + lane) : 0u;
                float2 v[8];
                float sv[8];
    #pragma unroll
                for (int i = 0; i < 8; ++i) {
                    // (past the end: row 0 - added to sums that are never stored)
                    const float *r = reinterpret_cast<const float *>(rows_l + (size_t)__shfl_sync(kFull, mine, i) * row_bytes);
                    v[i] = ldg_f2(r);
                    sv[i] = lane == 0 ? __ldg(r + kCvFeat) : 0.f;   // lane 0's pair starts the row: the score is 64 floats on
                }
                // no end-of-list test per row: once the last pixel has left, `left` never reaches zero again
    #pragma unroll
                for (int i = 0; i < 8; ++i) {
                    acc.x = __fadd_rn(acc.x, v[i].x);
                    acc.y = __fadd_rn(acc.y, v[i].y);
                    sacc = __fadd_rn(sacc, sv[i]);
                    if (--left == 0) {   // (warp-uniform) the pixel is complete: its mean (scores: its sum) goes to the tile
                        const float2 m = run > 1 ? mean2(acc, run) : acc;
                        const int col = cv_col(2 * lane, px);   // px odd: the pair's columns swap places
                        *reinterpret_cast<float2 *>(pm + px * kCvFeat + (col & ~1)) = (px & 1) ? make_float2(m.y, m.x) : m;
                        if (lane == 0) occ_s[warp][px] = sacc;
                        acc = make_float2(0.f, 0.f);
                        sacc = 0.f;
                        occ &= occ - 1;
                        px = occ ? __ffs(occ) - 1 : 0;
                        run = __shfl_sync(kFull, cnt, px);
                        left = occ ? run : 0x7fffffff;
                    }
                }
                mine = nxt;
            }
            __syncwarp();
        }
        // channel-major rows of 32 pixels.  Lane (g4, q4): channel 4 it + g4, pixels q4 .. q4 + 3 - one 16-byte store, eight
        // lanes make a channel's 128-byte line; the four tile reads hit 32 different banks.
        const unsigned o4 = occ_all >> q4 & 15u;
        float *d = oute + (size_t)g4 * P + p0 + q4;
        const size_t step = 4 * (size_t)P;
        if (occ_all) {
            // column of channel 4 it + g4 in pixel q4 + t: (4 it ^ q4) | (g4 ^ t).  A tile row is 256 bytes and 256-byte
            // aligned, so the byte address is (row | (g4 ^ t) << 2 | q4 << 2) ^ (it << 4): one XOR per load.
            const unsigned a0 = smem_u32(pm) + (unsigned)(q4 * kCvFeat * 4) + (unsigned)(q4 << 2);
            const unsigned t0 = a0 + 0 * kCvFeat * 4 + ((g4 ^ 0) << 2), t1 = a0 + 1 * kCvFeat * 4 + ((g4 ^ 1) << 2);
            const unsigned t2 = a0 + 2 * kCvFeat * 4 + ((g4 ^ 2) << 2), t3 = a0 + 3 * kCvFeat * 4 + ((g4 ^ 3) << 2);
    #pragma unroll
            for (int it = 0; it < kCvFeat / 4; ++it) {
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
            for (int it = 0; it < kCvFeat / 4; ++it) {
                *reinterpret_cast<float4 *>(d) = make_float4(0.f, 0.f, 0.f, 0.f);
                d += step;
            }
        }
        if (kCvFeat + g4 < Cx) {   // the summed scores, then the padding channels of the feature rows
            float4 val = make_float4(0.f, 0.f, 0.f, 0.f);
            if (g4 == 0) {
                val.x = (o4 & 1u) ? occ_s[warp][q4 + 0] : 0.f;
                val.y = (o4 & 2u) ? occ_s[warp][q4 + 1] : 0.f;
                val.z = (o4 & 4u) ? occ_s[warp][q4 + 2] : 0.f;
                val.w = (o4 & 8u) ? occ_s[warp][q4 + 3] : 0.f;
            }
            *reinterpret_cast<float4 *>(d) = val;
        }
        __syncwarp();   // the tile is free for the next bucket
        st = nst;
        st_hi = nst_hi;
        first = nfirst;
    }
}

}  // namespace cmr
