// cmr_b200.cu - the C ABI of libcmr_b200.so (see include/cmr_b200.h).  Single translation unit:
// argument checks + launch configuration here, kernels in env_kernels.cuh / pointnet_kernels.cuh.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "env_kernels.cuh"
#include "heads_kernels.cuh"
#include "pointnet_kernels.cuh"
#include "scatter_kernels.cuh"
#include "cost_volume_kernels.cuh"
#include "sample_kernels.cuh"
#include "dataset_kernels.cuh"
#include "tower_kernels.cuh"
#include "session.cuh"

namespace cmr {

static inline cudaStream_t S_(void *s) { return reinterpret_cast<cudaStream_t>(s); }

template <typename Kern>
static int allow_smem(Kern kern, size_t bytes) {
    // static shared memory counts against the 48 KB default too: opt in well below the limit
    if (bytes <= 32 * 1024) return CMR_OK;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    return e == cudaSuccess ? CMR_OK : (int)e;
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link against libcuda)
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}
// fp32 tensor [d2][d1][d0] (d0 contiguous), boxes of [1][b1][b0]
static bool make_map3d(CUtensorMap *m, const float *base, uint64_t d0, uint64_t d1, uint64_t d2, uint32_t b0, uint32_t b1,
                       CUtensorMapSwizzle swizzle = CU_TENSOR_MAP_SWIZZLE_NONE) {
    EncodeTiledFn fn = encode_tiled();
    if (!fn) return false;
    cuuint64_t dims[3] = {d0, d1, d2};
    cuuint64_t strides[2] = {d0 * sizeof(float), d0 * d1 * sizeof(float)};
    cuuint32_t box[3] = {b0, b1, 1};
    cuuint32_t es[3] = {1, 1, 1};
    return fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float *>(base), dims, strides, box, es,
              CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// bf16 feature planes of the 3-D tower: tensor [B][N][64] (128-byte rows), boxes of [1][rows][64], 128B swizzle
static bool make_plane_map(CUtensorMap *m, const void *base, uint64_t N, uint64_t B, uint32_t rows) {
    EncodeTiledFn fn = encode_tiled();
    if (!fn) return false;
    cuuint64_t dims[3] = {64, N, B};
    cuuint64_t strides[2] = {128, N * 128};
    cuuint32_t box[3] = {64, rows, 1};
    cuuint32_t es[3] = {1, 1, 1};
    return fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void *>(base), dims, strides, box, es,
              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

static int sm_count() {
    static int n = 0;
    if (!n) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
            n = 148;
    }
    return n;
}
}  // namespace cmr

using namespace cmr;

#define CMR_REQUIRE(cond, code) \
    do {                        \
        if (!(cond)) return (code); \
    } while (0)

template <typename PixT, int CQ>
static int launch_tile_scatter(const WsLayout &L, const char *ws, const float *img_feat, const float *K, int W, int B,
                               int N, int C, int P, bool copy_image, float *obs2d, cudaStream_t st) {
    const PixT *pix = reinterpret_cast<const PixT *>(ws + L.off_pix);
    const int *M = reinterpret_cast<const int *>(ws + L.off_m);
    const float *featT = reinterpret_cast<const float *>(ws + L.off_feat);
    size_t smem = sizeof(float) * kTilePix * (C + 1) + sizeof(int) * kTilePix + sizeof(unsigned) * 16 * kWarpList;
    int rc = allow_smem(k_tile_scatter<PixT, CQ>, smem);
    if (rc) return rc;
    const bool vec = (P % 4 == 0) && aligned(img_feat, 16) && aligned(obs2d, 16);
    // programmatic dependent launch: the preamble (accumulator clear) overlaps the tail of the preceding
    // k_project; the kernel waits (griddepcontrol.wait) before reading pixel ids or writing obs2d
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(B, ceil_div(P, kTilePix));   // x = episode, y = tile rank (heavy tiles first)
    cfg.blockDim = dim3(256);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, k_tile_scatter<PixT, CQ>, pix, M, featT, img_feat, K, W, N, L.ncap, C, P,
                                       copy_image, vec, obs2d);
    ++g_launches;
    if (e != cudaSuccess) return (int)e;
    e = cudaGetLastError();
    return e == cudaSuccess ? CMR_OK : (int)e;
}

template <typename Kern, typename... Args>
static int launch_pdl(Kern kern, dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
    static const bool chain = [] { const char *e = getenv("CMR_B200_PDL"); return !(e && e[0] == '0'); }();
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = chain ? 1 : 0;   // CMR_B200_PDL=0: plain stream order (the kernels' griddepcontrol.wait is then a no-op)
    cudaError_t e = cudaLaunchKernelEx(&cfg, kern, args...);
    ++g_launches;
    if (e != cudaSuccess) return (int)e;
    e = cudaGetLastError();
    return e == cudaSuccess ? CMR_OK : (int)e;
}

// true when the projected half of obs2d goes through the bucket buffers k_project fills (scatter_kernels.cuh)
static bool bucket_path(const WsLayout &L, int C) {
    return L.buckets > 0 && kHeavyCtas + std::max(ceil_div(L.buckets, kGatherWarps), kHeavyCtas) <= 65535;
}

// projected half of obs2d from the bucket buffers: k_tile_gather (scatter_kernels.cuh)
// share / out_rows / row0 / mean_channels: see k_tile_gather.  The observation writes the projected half of
// obs2d [B, 2C, P] (out_rows = 2C, row0 = C); a cost volume writes [E, C, P] (out_rows = C, row0 = 0).
static int launch_gather(const WsLayout &L, const char *ws, const float *img_feat, int B, int N, int C, int P,
                         bool copy_image, float *obs2d, cudaStream_t st, int share = 1, int out_rows = 0, int row0 = -1,
                         int mean_channels = 1 << 30, int heavy_from = kLightLimit) {
    if (out_rows <= 0) out_rows = 2 * C;
    if (row0 < 0) row0 = C;
    int *bcnt = reinterpret_cast<int *>(const_cast<char *>(ws) + L.off_bcnt);
    const unsigned *bbuf = reinterpret_cast<const unsigned *>(ws + L.off_bbuf);
    const void *pix = ws + L.off_pix;
    const int *M = reinterpret_cast<const int *>(ws + L.off_m);
    const float *featT = reinterpret_cast<const float *>(ws + L.off_feat);
    const int *hq = reinterpret_cast<const int *>(ws + L.off_hq);
    const bool vec = (P % 4 == 0) && aligned(obs2d, 16);
    // result tiles (32 pixels x 64 channels, 128-byte rows, 128B-swizzled in shared memory) leave as ONE tiled TMA store
    alignas(64) CUtensorMap map_proj;
    memset(&map_proj, 0, sizeof(map_proj));
    const bool tma = vec && P >= kBucketPix &&
                     make_map3d(&map_proj, obs2d, P, (uint64_t)out_rows, B, kBucketPix, kSlab, CU_TENSOR_MAP_SWIZZLE_128B);
    // image half of obs2d inside this kernel (cmr_project left it to us): boxes of the same shape, same swizzle
    alignas(64) CUtensorMap map_img;
    memset(&map_img, 0, sizeof(map_img));
    const bool img_tma = copy_image && tma && share == 1 && out_rows == 2 * C && aligned(img_feat, 16) &&
                         make_map3d(&map_img, img_feat, P, (uint64_t)C, B, kBucketPix, kSlab, CU_TENSOR_MAP_SWIZZLE_128B);
    // x = episode, y = kHeavyCtas bucket CTAs interleaved with the first light CTAs (8 buckets each), then the rest
    const int light = std::max(ceil_div(L.buckets, kGatherWarps), kHeavyCtas);
    // four SUMMED channels after whole slabs (features + occupancy of a cost volume): no slab pass of their own
    const int tail = (C > kSlab && C % kSlab == 4 && mean_channels <= C - 4) ? 4 : 0;
    if (heavy_from > kLightLimit) {   // a cost volume: its warps hold up to kMidMax keys
        const size_t smem = std::max(kGatherSmem, gather_smem_light(kHeavyFrom));
        int rc = allow_smem(k_tile_gather<true>, smem);
        if (rc) return rc;
        return launch_pdl(k_tile_gather<true>, dim3(B, kHeavyCtas + light), dim3(kGatherThreads), smem, st, bcnt, bbuf, L.buckets, hq, pix,
                          L.pix16 ? 1 : 0, M, featT, img_feat, N, L.ncap, C, P, copy_image, vec, tma, obs2d, map_proj, share,
                          (long long)out_rows * P, (long long)row0 * P, row0, mean_channels, img_tma, map_img, tail);
    }
    int rc = allow_smem(k_tile_gather<false>, kGatherSmem);
    if (rc) return rc;
    return launch_pdl(k_tile_gather<false>, dim3(B, kHeavyCtas + light), dim3(kGatherThreads), kGatherSmem, st, bcnt, bbuf,
                      L.buckets, hq, pix, L.pix16 ? 1 : 0, M, featT, img_feat, N, L.ncap, C, P, copy_image, vec, tma, obs2d, map_proj,
                      share, (long long)out_rows * P, (long long)row0 * P, row0, mean_channels, img_tma, map_img, tail);
}

// true when the image half of obs2d can travel as tiled TMA boxes inside k_project
static bool image_copy_by_tma(const float *img_feat, const float *obs2d, int B, int C, int P, CUtensorMap *map_img,
                              CUtensorMap *map_out) {
    memset(map_img, 0, sizeof(*map_img));
    memset(map_out, 0, sizeof(*map_out));
    return img_feat && obs2d && (P % 4 == 0) && P >= kProjBoxPix && aligned(img_feat, 16) && aligned(obs2d, 16) &&
           make_map3d(map_img, img_feat, P, C, B, kProjBoxPix, C) &&
           make_map3d(map_out, obs2d, P, 2 * (uint64_t)C, B, kProjBoxPix, C);
}

template <typename PixT>
static int launch_project(const WsLayout &L, char *ws, const float *pc, const uint8_t *overlap, const float *K,
                          const float *pose, const float *mean, int B, int N, int C, int H, int W, float *obs3d,
                          int32_t *pix_out, int32_t *mvis_out, bool img_tma, const CUtensorMap &map_img,
                          const CUtensorMap &map_out, bool clear_counters, cudaStream_t st, int share = 1) {
    PixT *pix = reinterpret_cast<PixT *>(ws + L.off_pix);
    const int *seg = reinterpret_cast<const int *>(ws + L.off_seg);
    const int *Mws = reinterpret_cast<const int *>(ws + L.off_m);
    const bool vec = (N % 4 == 0) && aligned(pc, 16) && aligned(obs3d, 16) && aligned(overlap, 4) &&
                     (!pix_out || aligned(pix_out, 16));
    if (mvis_out) {
        cudaError_t e = cudaMemsetAsync(mvis_out, 0, sizeof(int32_t) * B, st);
        if (e != cudaSuccess) return (int)e;
    }
    const int img_tiles = img_tma ? ceil_div(H * W, kProjBoxPix) : 0;
    const size_t smem = img_tma ? sizeof(float) * kProjBoxPix * C : 0;
    int rc = allow_smem(k_project<PixT>, smem);
    if (rc) return rc;
    int *bcnt = nullptr;
    unsigned *bbuf = nullptr;
    int *hq = nullptr;
    if (bucket_path(L, C)) {
        hq = reinterpret_cast<int *>(ws + L.off_hq);
        bcnt = reinterpret_cast<int *>(ws + L.off_bcnt);
        bbuf = reinterpret_cast<unsigned *>(ws + L.off_bbuf);
        if (clear_counters) {   // cmr_observe relies on k_tile_gather having cleared them instead
            cudaError_t e = cudaMemsetAsync(bcnt, 0, L.bcnt_bytes, st);
            if (e != cudaSuccess) return (int)e;
        }
    }
    return launch_pdl(k_project<PixT>, dim3(ceil_div(L.groups, kProjWarps), B), dim3(32 * kProjWarps), smem, st, pc, overlap, K,
                      pose, mean, seg, Mws, N, L.ncap, L.groups, H, W, vec, pix, obs3d, pix_out, mvis_out, bcnt, bbuf,
                      L.buckets, bcnt ? bcnt + (size_t)B * kBucketStride : (int *)nullptr, hq, share, img_tiles, C, map_img, map_out);
}

template <typename PixT>
static int launch_scatter(const WsLayout &L, const char *ws, const float *img_feat, const float *K, int W, int B, int N,
                          int C, int P, bool copy_image, float *obs2d, cudaStream_t st) {
    if (C <= 32) return launch_tile_scatter<PixT, 1>(L, ws, img_feat, K, W, B, N, C, P, copy_image, obs2d, st);
    if (C <= 64) return launch_tile_scatter<PixT, 2>(L, ws, img_feat, K, W, B, N, C, P, copy_image, obs2d, st);
    if (C <= 128) return launch_tile_scatter<PixT, 4>(L, ws, img_feat, K, W, B, N, C, P, copy_image, obs2d, st);
    return launch_tile_scatter<PixT, 8>(L, ws, img_feat, K, W, B, N, C, P, copy_image, obs2d, st);
}

template <typename VecT>
static int launch_index_points(const void *points, const int64_t *idx, int B, int N, int S, int row_bytes, void *out,
                               cudaStream_t st) {
    int row_vecs = row_bytes / (int)sizeof(VecT);
    long long total = (long long)B * S * row_vecs;
    int grid = (int)std::min<long long>((total + 255) / 256, (long long)sm_count() * 32);
    k_index_points<VecT><<<grid, 256, 0, st>>>(static_cast<const VecT *>(points), idx, N, S, row_vecs,
                                               static_cast<VecT *>(out), total);
    return after_launch();
}

template <int PPT, int MINB>
static int launch_fps(const float *xyz, const int64_t *start, int B, int N, int npoint, int64_t *out, int cs,
                      cudaStream_t st) {
    constexpr int THREADS = 512;
    size_t smem = sizeof(float) * 3 * THREADS * PPT;
    int rc = allow_smem(k_fps<PPT, THREADS, MINB>, smem);
    if (rc) return rc;
    if (cs > 8) {
        cudaError_t e = cudaFuncSetAttribute(k_fps<PPT, THREADS, MINB>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
        if (e != cudaSuccess) return (int)e;
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(B * cs);
    cfg.blockDim = dim3(THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cs;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, k_fps<PPT, THREADS, MINB>, xyz, start, N, npoint, out);
    ++g_launches;
    if (e != cudaSuccess) return (int)e;
    e = cudaGetLastError();
    return e == cudaSuccess ? CMR_OK : (int)e;
}

// rows of the predicted-overlap points, channel-major -> point-major: by TMA boxes when the tensor allows it
static int launch_feat_compact(const uint8_t *overlap, const float *feat, int B, int N, int C, int groups, const int *seg,
                               float *featT, cudaStream_t st) {
    alignas(64) CUtensorMap map_feat;
    memset(&map_feat, 0, sizeof(map_feat));
    if (N % 4 == 0 && N >= 32 && C >= 64 && aligned(feat, 16) &&
        make_map3d(&map_feat, feat, N, (uint64_t)C, B, 32, 64, CU_TENSOR_MAP_SWIZZLE_128B)) {
        const size_t smem = sizeof(float) * 4 * 64 * 32;
        int rc = allow_smem(k_feat_compact_tma, smem);
        if (rc) return rc;
        k_feat_compact_tma<<<dim3(groups, B), 256, smem, st>>>(overlap, N, C, groups, seg, featT, map_feat);
        return after_launch();
    }
    const size_t smem = sizeof(float) * kGroup * (C + 1);
    int rc = allow_smem(k_feat_compact<256>, smem);
    if (rc) return rc;
    k_feat_compact<256><<<dim3(groups, B), 256, smem, st>>>(overlap, feat, N, C, groups, (N % 4 == 0) && aligned(feat, 16),
                                                            seg, featT);
    return after_launch();
}

// workspace: four fp16 planes [B][N][64] (two (hi, lo) pairs, ping-pong between the blocks) + the max keys of the four blocks (64, 64, 64, 128 per episode)
struct TowerWs {
    size_t plane, off_keys, keys_bytes, total;
};
static TowerWs tower_ws(int B, int N) {
    TowerWs w;
    w.plane = round_up((size_t)B * N * 128, 1024);
    w.off_keys = 4 * w.plane;
    w.keys_bytes = (size_t)B * (64 * 3 + 128) * sizeof(unsigned);
    w.total = w.off_keys + round_up(w.keys_bytes, 1024);
    return w;
}
template <bool kLast>
static int launch_tower_mma(const void *blob, int B, int N, int tiles_per_ep, int box_rows, const CUtensorMap &in_hi, const CUtensorMap &in_lo,
                            const CUtensorMap &out_hi, const CUtensorMap &out_lo, const unsigned *prev_keys, unsigned *max_keys,
                            cudaStream_t st) {
    auto kern = k_tower_mma<kLast>;
    const size_t smem = TowerCfg<kLast>::smem_bytes;
    int rc = allow_smem(kern, smem);
    if (rc) return rc;
    const int grid = std::min(B * tiles_per_ep, sm_count());
    return launch_pdl(kern, dim3(grid), dim3(kTowerThreads), smem, st, static_cast<const unsigned char *>(blob), B, N, tiles_per_ep, box_rows,
                      in_hi, in_lo, out_hi, out_lo, prev_keys, max_keys);
}

// CMR_B200_CV=buckets: the cost volume takes the observation's bucket kernels whatever its shape (parity tests of that path)
static bool cv_force_buckets() {
    const char *e = getenv("CMR_B200_CV");
    return e && strcmp(e, "buckets") == 0;
}

extern "C" {

int cmr_abi_version(void) { return CMR_ABI_VERSION; }

const char *cmr_error_string(int code) {
    switch (code) {
        case CMR_OK: return "ok";
        case CMR_EINVAL: return "invalid argument (null pointer or non-positive size)";
        case CMR_EALIGN: return "pointer not aligned as documented";
        case CMR_ERANGE: return "size outside the supported range";
        case CMR_EUNSUPPORTED: return "unsupported configuration";
        default: return code > 0 ? cudaGetErrorString((cudaError_t)code) : "unknown error";
    }
}

unsigned long long cmr_launch_count(void) { return g_launches; }

int cmr_take_fault(void *stream) {
    int v = 0;
    cudaError_t e = cudaMemcpyFromSymbolAsync(&v, g_fault, sizeof(int), 0, cudaMemcpyDeviceToHost, S_(stream));
    if (e == cudaSuccess) e = cudaStreamSynchronize(S_(stream));
    if (e != cudaSuccess) return (int)e;
    if (v) {
        int z = 0;
        cudaMemcpyToSymbolAsync(g_fault, &z, sizeof(int), 0, cudaMemcpyHostToDevice, S_(stream));
        cudaStreamSynchronize(S_(stream));
    }
    return v;
}

// ------------------------------------------------------------------------------ environment ----

size_t cmr_workspace_bytes(int B, int N, int C, int P) {
    if (B <= 0 || N <= 0 || C <= 0 || P <= 0) return 0;
    return ws_layout(B, N, C, P).total;
}

int cmr_cloud_mean(const float *pc, int B, int N, float *mean, void *stream) {
    CMR_REQUIRE(pc && mean && B > 0 && N > 0, CMR_EINVAL);
    k_cloud_mean<<<dim3(3, B), 1024, 0, S_(stream)>>>(pc, N, (N % 4 == 0) && aligned(pc, 16), mean);
    return after_launch();
}

// The two halves of cmr_episode_prepare.  cmr_project needs only the first (the prefix of the overlap counts), so a
// host with two streams can run the second - the big one: it reads all of feat - beside the first projection.
int cmr_episode_scan(const uint8_t *overlap, int B, int N, int C, void *workspace, void *stream) {
    CMR_REQUIRE(overlap && workspace && B > 0 && N > 0 && C > 0, CMR_EINVAL);
    CMR_REQUIRE(C <= kMaxC && (C % 4) == 0 && N < (1 << 24) && B <= 65535, CMR_ERANGE);
    CMR_REQUIRE(aligned(workspace, 256), CMR_EALIGN);
    WsLayout L = ws_layout(B, N, C, 1);
    char *ws = static_cast<char *>(workspace);
    int *M = reinterpret_cast<int *>(ws + L.off_m);
    int *seg = reinterpret_cast<int *>(ws + L.off_seg);
    // bucket counters + ticket of the scatter stage start at zero; k_tile_gather leaves them at zero
    cudaError_t me = cudaMemsetAsync(ws + L.off_bcnt, 0, L.bcnt_bytes, S_(stream));
    if (me != cudaSuccess) return (int)me;
    k_overlap_scan<<<B, 1024, 0, S_(stream)>>>(overlap, N, L.groups, (N % 4 == 0) && aligned(overlap, 4), seg, M);
    return after_launch();
}

int cmr_episode_compact(const uint8_t *overlap, const float *feat, int B, int N, int C, void *workspace,
                        void *stream) {
    CMR_REQUIRE(overlap && feat && workspace && B > 0 && N > 0 && C > 0, CMR_EINVAL);
    CMR_REQUIRE(C <= kMaxC && (C % 4) == 0 && N < (1 << 24) && B <= 65535, CMR_ERANGE);
    CMR_REQUIRE(aligned(workspace, 256), CMR_EALIGN);
    WsLayout L = ws_layout(B, N, C, 1);
    char *ws = static_cast<char *>(workspace);
    const int *seg = reinterpret_cast<const int *>(ws + L.off_seg);
    float *featT = reinterpret_cast<float *>(ws + L.off_feat);
    return launch_feat_compact(overlap, feat, B, N, C, L.groups, seg, featT, S_(stream));
}

int cmr_episode_prepare(const uint8_t *overlap, const float *feat, int B, int N, int C, void *workspace,
                        void *stream) {
    CMR_REQUIRE(feat, CMR_EINVAL);
    int rc = cmr_episode_scan(overlap, B, N, C, workspace, stream);
    if (rc) return rc;
    return cmr_episode_compact(overlap, feat, B, N, C, workspace, stream);
}

static int check_observe_dims(int B, int N, int C, int H, int W) {
    CMR_REQUIRE(B > 0 && N > 0 && C > 0 && H > 0 && W > 0, CMR_EINVAL);
    CMR_REQUIRE(C <= kMaxC && (C % 4) == 0 && N < (1 << 24) && B <= 65535 && (long long)H * W < (1ll << 30), CMR_ERANGE);
    return CMR_OK;
}

static int project_impl(const float *pc, const uint8_t *overlap, const float *K, const float *pose, const float *mean,
                        void *workspace, int B, int N, int C, int H, int W, float *obs3d, int32_t *pix_out,
                        int32_t *mvis_out, const float *img_feat, float *obs2d, int *image_copied, bool clear_counters,
                        void *stream) {
    CMR_REQUIRE(pc && overlap && K && pose && mean && workspace && obs3d, CMR_EINVAL);
    int rc = check_observe_dims(B, N, C, H, W);
    if (rc) return rc;
    CMR_REQUIRE(aligned(workspace, 256) && aligned(pose, 16), CMR_EALIGN);   // pose rows are read as float4
    WsLayout L = ws_layout(B, N, C, H * W);
    char *ws = static_cast<char *>(workspace);
    alignas(64) CUtensorMap map_img, map_out;
    // who carries the image half of obs2d: k_tile_gather when the bucket path is taken (its light warps have an idle
    // tile and idle time while they wait for k_project); on grids too large for the bucket path k_project does, as
    // tiled TMA traffic beside its arithmetic
    const bool img_tma = !bucket_path(L, C) && image_copy_by_tma(img_feat, obs2d, B, C, H * W, &map_img, &map_out);
    if (!img_tma) {
        memset(&map_img, 0, sizeof(map_img));
        memset(&map_out, 0, sizeof(map_out));
    }
    if (image_copied) *image_copied = img_tma ? 1 : 0;
    if (L.pix16)
        return launch_project<uint16_t>(L, ws, pc, overlap, K, pose, mean, B, N, C, H, W, obs3d, pix_out, mvis_out, img_tma,
                                        map_img, map_out, clear_counters, S_(stream));
    return launch_project<int32_t>(L, ws, pc, overlap, K, pose, mean, B, N, C, H, W, obs3d, pix_out, mvis_out, img_tma,
                                   map_img, map_out, clear_counters, S_(stream));
}

// Stand-alone stage 1.  Unless the caller promises the project/scatter pairing, the bucket counters are
// cleared first, so that the call may be repeated; a cmr_tile_scatter consumes what ONE cmr_project left.
int cmr_project(const float *pc, const uint8_t *overlap, const float *K, const float *pose, const float *mean,
                void *workspace, int B, int N, int C, int H, int W, float *obs3d, int32_t *pix_out, int32_t *mvis_out,
                const float *img_feat, float *obs2d, int *image_copied, int flags, void *stream) {
    return project_impl(pc, overlap, K, pose, mean, workspace, B, N, C, H, W, obs3d, pix_out, mvis_out, img_feat, obs2d,
                        image_copied, !(flags & CMR_PROJECT_PAIRED), stream);
}

int cmr_tile_scatter(const float *img_feat, const float *K, void *workspace, int B, int N, int C, int H, int W,
                     int copy_image, float *obs2d, void *stream) {
    CMR_REQUIRE(K && workspace && obs2d && (img_feat || !copy_image), CMR_EINVAL);
    CMR_REQUIRE((long long)ceil_div(H * W, kTilePix) <= 65535, CMR_ERANGE);
    int rc = check_observe_dims(B, N, C, H, W);
    if (rc) return rc;
    CMR_REQUIRE(aligned(workspace, 256), CMR_EALIGN);
    WsLayout L = ws_layout(B, N, C, H * W);
    const char *ws = static_cast<const char *>(workspace);
    const int P = H * W;
    if (bucket_path(L, C))   // the visible points were bucketed by k_project: every unit reads only its own
        return launch_gather(L, ws, img_feat, B, N, C, P, copy_image != 0, obs2d, S_(stream));
    // large grids: every tile CTA searches the episode's id list itself
    if (L.pix16)
        return launch_scatter<uint16_t>(L, ws, img_feat, K, W, B, N, C, P, copy_image != 0, obs2d, S_(stream));
    return launch_scatter<int32_t>(L, ws, img_feat, K, W, B, N, C, P, copy_image != 0, obs2d, S_(stream));
}

// The hot path: stage 1 + stage 2.  The bucket counters are zero on entry (cmr_episode_prepare) and zero again
// on exit (the last CTA of k_tile_gather clears them), so no memset sits between the steps of a rollout.
int cmr_observe(const float *pc, const uint8_t *overlap, const float *img_feat, const float *K, const float *pose,
                const float *mean, void *workspace, int B, int N, int C, int H, int W, float *obs2d, float *obs3d,
                int32_t *pix_out, int32_t *mvis_out, void *stream) {
    CMR_REQUIRE(img_feat && obs2d, CMR_EINVAL);
    int copied = 0;
    int rc = check_observe_dims(B, N, C, H, W);
    if (rc) return rc;
    rc = project_impl(pc, overlap, K, pose, mean, workspace, B, N, C, H, W, obs3d, pix_out, mvis_out, img_feat, obs2d,
                      &copied, false, stream);
    const bool projected = rc == CMR_OK;
    if (!rc) rc = cmr_tile_scatter(img_feat, K, workspace, B, N, C, H, W, copied ? 0 : 1, obs2d, stream);
    if (rc && projected) {
        // k_project has filled the bucket counters and nobody will consume them: leave the workspace as the next
        // cmr_observe expects it (counters and queue header zero) instead of silently wrong results later
        WsLayout L = ws_layout(B, N, C, H * W);
        cudaMemsetAsync(static_cast<char *>(workspace) + L.off_bcnt, 0, L.bcnt_bytes, S_(stream));
    }
    return rc;
}

int cmr_to_disentangled(float *poses, const float *mean, int B, void *stream) {
    CMR_REQUIRE(poses && mean && B > 0, CMR_EINVAL);
    k_to_disentangled<<<ceil_div(B, 128), 128, 0, S_(stream)>>>(poses, mean, B);
    return after_launch();
}

int cmr_step(float *pose, const int64_t *action_r, const int64_t *action_t, const float *rot_tab, const float *t_tab,
             int nbins, int dof6, int B, void *stream) {
    CMR_REQUIRE(pose && action_r && action_t && rot_tab && t_tab && nbins > 0 && B > 0, CMR_EINVAL);
    return launch_pdl(k_step, dim3(ceil_div(B, 128)), dim3(128), 0, S_(stream), pose, action_r, action_t, rot_tab, t_tab, nbins,
                      dof6 ? 1 : 0, B);
}

int cmr_expert(const float *pose_source, const float *pose_target, const double *r_steps, const double *t_steps,
               int nbins, int dof6, int B, int64_t *action_r, int64_t *action_t, void *stream) {
    CMR_REQUIRE(pose_source && pose_target && r_steps && t_steps && action_r && action_t && nbins > 0 && B > 0, CMR_EINVAL);
    k_expert<<<ceil_div(B, 64), 64, 0, S_(stream)>>>(pose_source, pose_target, r_steps, t_steps, nbins, dof6 ? 1 : 0, B,
                                                     action_r, action_t);
    return after_launch();
}

size_t cmr_reward_scratch_bytes(int B) { return B > 0 ? (size_t)B * kRewardSlotBytes : 0; }

int cmr_reward(const float *target, const float *pc, const uint8_t *mask, const float *mean, const float *pose,
               const float *prev, int mode, int B, int N, void *scratch, float *reward, float *dist, void *stream) {
    CMR_REQUIRE(target && pc && mask && mean && scratch && reward && dist && B > 0 && N > 0, CMR_EINVAL);
    CMR_REQUIRE(mode == CMR_REWARD_SHIPPED || (mode == CMR_REWARD_INTENDED && pose), CMR_EINVAL);
    CMR_REQUIRE(aligned(scratch, 16), CMR_EALIGN);
    CMR_REQUIRE(B <= 65535, CMR_ERANGE);
    int nchunks = std::min(kRewardChunks, ceil_div(N, 2048));
    int per_chunk = (int)round_up((size_t)ceil_div(N, nchunks), 1024);
    nchunks = ceil_div(N, per_chunk);
    const bool vec = (N % 4 == 0) && aligned(mask, 4);
    return launch_pdl(k_reward, dim3(nchunks, B), dim3(256), 0, S_(stream), target, pc, mask, mean, pose, prev, mode, N,
                      per_chunk, vec, static_cast<unsigned char *>(scratch), reward, dist);
}

int cmr_reward_compare(const float *dist_cached, const float *prev, int B, float *reward, float *dist, void *stream) {
    CMR_REQUIRE(dist_cached && reward && dist && B > 0, CMR_EINVAL);
    return launch_pdl(k_reward_compare, dim3(ceil_div(B, 128)), dim3(128), 0, S_(stream), dist_cached, prev, B, reward, dist);
}

// One agent iteration in one call: step (:179-207) -> reward (:263-302) -> observation of the new pose (:25-126),
// each part optional.  Nothing here that the separate entry points do not do - it saves a host round trip per part.
int cmr_iteration(const cmr_iteration_args *a, float *pose, const int64_t *action_r, const int64_t *action_t,
                  const float *prev, float *reward, float *dist, float *obs2d, float *obs3d, void *stream) {
    CMR_REQUIRE(a && pose, CMR_EINVAL);
    int rc = CMR_OK;
    if (action_r || action_t) {
        CMR_REQUIRE(action_r && action_t, CMR_EINVAL);
        rc = cmr_step(pose, action_r, action_t, a->rot_tab, a->t_tab, a->nbins, a->dof6, a->B, stream);
        if (rc) return rc;
    }
    if (reward || dist) {
        CMR_REQUIRE(reward && dist, CMR_EINVAL);
        if (a->reward_mode == CMR_REWARD_SHIPPED && a->dist_cached)
            rc = cmr_reward_compare(a->dist_cached, prev, a->B, reward, dist, stream);
        else
            rc = cmr_reward(a->target, a->pc, a->mask, a->mean, pose, prev, a->reward_mode, a->B, a->N, a->reward_scratch, reward,
                            dist, stream);
        if (rc) return rc;
    }
    if (obs2d || obs3d) {
        CMR_REQUIRE(obs2d && obs3d, CMR_EINVAL);
        rc = cmr_observe(a->pc, a->overlap, a->img_feat, a->K, pose, a->mean, a->workspace, a->B, a->N, a->C, a->H, a->W, obs2d,
                         obs3d, nullptr, nullptr, stream);
    }
    return rc;
}

// ---------------------------------------------------------------------------- agent: heads ----

int cmr_grouped_linear(const float *in, int in_stride, const float *W, const float *bias, const int64_t *desc, int groups,
                       int B, int N, float negative_slope, int activate, float *out, int out_stride, void *stream) {
    CMR_REQUIRE(in && W && bias && desc && out && B > 0 && N > 0 && groups > 0, CMR_EINVAL);
    CMR_REQUIRE(groups <= kLinMaxGroups, CMR_ERANGE);
    LinGroups G{};
    G.count = groups;
    int next = 0;
    for (int i = 0; i < groups; ++i) {
        const int64_t *d = desc + 5 * i;
        CMR_REQUIRE(d[1] > 0 && d[1] <= kLinMaxK, CMR_ERANGE);
        // the groups tile [0, N) in order; their inputs lie inside a row of `in`
        CMR_REQUIRE(d[2] == next && d[3] > d[2] && d[3] <= N && d[0] >= 0 && d[0] + d[1] <= in_stride && d[4] >= 0, CMR_EINVAL);
        G.g[i] = LinGroup{(int)d[0], (int)d[1], (int)d[2], (int)d[3], (long long)d[4]};
        next = (int)d[3];
    }
    CMR_REQUIRE(next == N && out_stride >= N, CMR_EINVAL);
    return launch_pdl(k_grouped_linear, dim3(ceil_div(N, 8)), dim3(256), 0, S_(stream), in, in_stride, W, bias, G, B, N,
                      negative_slope, activate ? 1 : 0, out, out_stride);
}

int cmr_conv_epilogue(const float *x, const float *scale, const float *shift, float negative_slope, int pool,
                      int channels_last, int B, int C, int H, int W, float *y, void *stream) {
    CMR_REQUIRE(x && scale && shift && y && B > 0 && C > 0 && H > 0 && W > 0, CMR_EINVAL);
    CMR_REQUIRE(pool >= 0 && pool <= 2, CMR_EINVAL);
    CMR_REQUIRE(aligned(x, 16) && aligned(y, 8), CMR_EALIGN);
    const long long planes = (long long)B * C, HW = (long long)H * W;
    CMR_REQUIRE(planes < (1ll << 31) && HW < (1ll << 31), CMR_ERANGE);
    const int max_blocks = 8 * sm_count();
    auto blocks = [&](long long n) { return dim3((unsigned)std::min<long long>((n + 255) / 256, max_blocks)); };
    cudaStream_t st = S_(stream);
    if (channels_last) {   // memory [B][H][W][C]
        CMR_REQUIRE(C % 4 == 0 && aligned(scale, 16) && aligned(shift, 16) && aligned(y, 16), CMR_EUNSUPPORTED);
        if (pool == 0) {
            const long long n4 = planes * HW / 4;
            return launch_pdl(k_conv_epilogue_nhwc, blocks(n4), dim3(256), 0, st, x, scale, shift, negative_slope, n4, C / 4, y);
        }
        if (pool == 1) {
            CMR_REQUIRE(W % 2 == 0 && H % 2 == 0, CMR_EUNSUPPORTED);
            const long long n4 = planes / 4 * (H / 2) * (W / 2);
            return launch_pdl(k_conv_epilogue_pool2_nhwc, blocks(n4), dim3(256), 0, st, x, scale, shift, negative_slope, n4, H, W,
                              C / 4, y);
        }
        CMR_REQUIRE(B <= 65535, CMR_ERANGE);
        return launch_pdl(k_conv_epilogue_global_nhwc, dim3((unsigned)ceil_div(C, 128), (unsigned)B), dim3(512), 0, st, x, scale, shift,
                          negative_slope, B, (int)HW, C, y);
    }
    if (pool == 0) {
        CMR_REQUIRE(HW % 4 == 0 && aligned(y, 16), CMR_EUNSUPPORTED);
        const long long n4 = planes * HW / 4;
        return launch_pdl(k_conv_epilogue, blocks(n4), dim3(256), 0, st, x, scale, shift, negative_slope, n4, (int)(HW / 4), C, y);
    }
    if (pool == 1) {
        CMR_REQUIRE(W % 4 == 0 && H % 2 == 0, CMR_EUNSUPPORTED);
        const long long n2 = planes * (H / 2) * (W / 4);
        return launch_pdl(k_conv_epilogue_pool2, blocks(n2), dim3(256), 0, st, x, scale, shift, negative_slope, n2, H, W, C, y);
    }
    return launch_pdl(k_conv_epilogue_global, dim3((unsigned)((planes + 7) / 8)), dim3(256), 0, st, x, scale, shift,
                      negative_slope, (int)planes, (int)HW, C, y);
}

int cmr_deterministic_action(const float *r_logits, int degree_r, int64_t r_batch_stride, const float *t_logits, int degree_t,
                             int64_t t_batch_stride, int B, int steps, int64_t *action_r, int64_t *action_t, float *probs_r,
                             float *probs_t, void *stream) {
    CMR_REQUIRE(r_logits && t_logits && action_r && action_t && B > 0 && degree_r > 0 && degree_t > 0, CMR_EINVAL);
    CMR_REQUIRE(steps >= 9 && steps <= 16, CMR_EUNSUPPORTED);   // the shapes whose summation order the kernel reproduces
    CMR_REQUIRE(r_batch_stride >= (int64_t)degree_r * steps && t_batch_stride >= (int64_t)degree_t * steps, CMR_EINVAL);
    CMR_REQUIRE((long long)B * (degree_r + degree_t) < (1ll << 30), CMR_ERANGE);
    const int rows = B * (degree_r + degree_t);
    return launch_pdl(k_deterministic_action, dim3((unsigned)ceil_div(rows, 8)), dim3(128), 0, S_(stream), r_logits, degree_r,
                      (long long)r_batch_stride, t_logits, degree_t, (long long)t_batch_stride, B, steps,
                      reinterpret_cast<long long *>(action_r), reinterpret_cast<long long *>(action_t), probs_r, probs_t);
}

int cmr_to_channels_last(const float *x, int B, int C, int H, int W, float *y, void *stream) {
    CMR_REQUIRE(x && y && x != y && B > 0 && C > 0 && H > 0 && W > 0, CMR_EINVAL);
    const long long P = (long long)H * W;
    CMR_REQUIRE(B <= 65535 && P < (1ll << 31) && (long long)C * P < (1ll << 40), CMR_ERANGE);
    if (C % 4 == 0 && P % 4 == 0 && aligned(x, 16) && aligned(y, 16))
        return launch_pdl(k_to_channels_last, dim3((unsigned)ceil_div((int)P, 128), (unsigned)ceil_div(C, 32), (unsigned)B), dim3(256),
                          0, S_(stream), x, C, (int)P, y);
    CMR_REQUIRE(ceil_div(C, 32) <= 65535, CMR_ERANGE);
    k_image_transpose<<<dim3((unsigned)ceil_div((int)P, 32), (unsigned)ceil_div(C, 32), (unsigned)B), 256, 0, S_(stream)>>>(x, C, (int)P, y);
    return after_launch();
}

// ---------------------------------------------------------------------------- pointnet_util ----

int cmr_square_distance(const float *src, const int64_t src_stride[3], const float *dst, const int64_t dst_stride[3],
                        int B, int S, int N, float *out, void *stream) {
    CMR_REQUIRE(src && dst && out && src_stride && dst_stride && B > 0 && S > 0 && N > 0, CMR_EINVAL);
    CMR_REQUIRE(S <= 65535 && B <= 65535, CMR_ERANGE);
    int gx = std::min(ceil_div(N, 256), 64);
    k_square_distance<<<dim3(gx, S, B), 256, 0, S_(stream)>>>(src, src_stride[0], src_stride[1], src_stride[2], dst,
                                                             dst_stride[0], dst_stride[1], dst_stride[2], S, N, out);
    return after_launch();
}

int cmr_index_points(const void *points, const int64_t *idx, int B, int N, int S, int row_bytes, void *out,
                     void *stream) {
    CMR_REQUIRE(points && idx && out && B > 0 && N > 0 && S > 0 && row_bytes > 0, CMR_EINVAL);
    if (row_bytes % 16 == 0 && aligned(points, 16) && aligned(out, 16))
        return launch_index_points<uint4>(points, idx, B, N, S, row_bytes, out, S_(stream));
    if (row_bytes % 4 == 0 && aligned(points, 4) && aligned(out, 4))
        return launch_index_points<unsigned>(points, idx, B, N, S, row_bytes, out, S_(stream));
    return launch_index_points<unsigned char>(points, idx, B, N, S, row_bytes, out, S_(stream));
}

int cmr_index_points_backward(const float *grad_out, const int64_t *idx, int B, int N, int S, int C,
                              float *grad_points, void *stream) {
    CMR_REQUIRE(grad_out && idx && grad_points && B > 0 && N > 0 && S > 0 && C > 0, CMR_EINVAL);
    long long total = (long long)B * S * C;
    int grid = (int)std::min<long long>((total + 255) / 256, (long long)sm_count() * 32);
    k_index_points_bwd<<<grid, 256, 0, S_(stream)>>>(grad_out, idx, N, S, C, grad_points, total);
    return after_launch();
}

int cmr_group_points(const float *xyz, const float *points, const float *new_xyz, const int64_t *idx, int B, int N,
                     int S, int K, int D, float *out, void *stream) {
    CMR_REQUIRE(xyz && new_xyz && idx && out && B > 0 && N > 0 && S > 0 && K > 0 && D >= 0, CMR_EINVAL);
    CMR_REQUIRE(D == 0 || points, CMR_EINVAL);
    long long total = (long long)B * S * K * (3 + D);
    int grid = (int)std::min<long long>((total + 255) / 256, (long long)sm_count() * 32);
    k_group_points<<<grid, 256, 0, S_(stream)>>>(xyz, points, new_xyz, idx, N, S, K, D, out, total);
    return after_launch();
}

int cmr_farthest_point_sample(const float *xyz, const int64_t *start, int B, int N, int npoint, int64_t *out,
                              void *stream) {
    CMR_REQUIRE(xyz && start && out && B > 0 && N > 0 && npoint > 0, CMR_EINVAL);
    constexpr int THREADS = 512, kMaxPpt = 24;
    // smallest cluster that holds the cloud in registers, widened while the whole batch still fits on the
    // chip in one wave (fewer points per thread = shorter rounds; FPS is latency-bound).  Measured on B200
    // (config 4, 128 clouds of 40960): clusters of 4 x 20 points/thread, one CTA per SM: 7.1 ms; clusters of
    // 8 x 10 points/thread with two CTAs per SM: 9.8 ms (the 8-CTA barrier costs more than it hides).
    int cs = 1;
    while (cs < 16 && (long long)cs * THREADS * kMaxPpt < N) cs *= 2;
    CMR_REQUIRE((long long)cs * THREADS * kMaxPpt >= N, CMR_ERANGE);
    const int sms = sm_count();
    while (cs < 8 && (long long)B * cs * 2 <= sms && (long long)cs * THREADS * 4 < N) cs *= 2;
    int ppt = ceil_div(N, cs * THREADS);
    cudaStream_t st = S_(stream);
    if (ppt <= 4) return launch_fps<4, 1>(xyz, start, B, N, npoint, out, cs, st);
    if (ppt <= 8) return launch_fps<8, 1>(xyz, start, B, N, npoint, out, cs, st);
    if (ppt <= 12) return launch_fps<12, 1>(xyz, start, B, N, npoint, out, cs, st);
    if (ppt <= 16) return launch_fps<16, 1>(xyz, start, B, N, npoint, out, cs, st);
    if (ppt <= 20) return launch_fps<20, 1>(xyz, start, B, N, npoint, out, cs, st);
    return launch_fps<24, 1>(xyz, start, B, N, npoint, out, cs, st);
}

int cmr_knn(const float *query, const float *ref, int B, int S, int N, int k, int64_t *out, void *stream) {
    CMR_REQUIRE(query && ref && out && B > 0 && S > 0 && N > 0 && k > 0, CMR_EINVAL);
    CMR_REQUIRE(k <= 128 && k <= N && B <= 65535, CMR_ERANGE);
    cudaStream_t st = S_(stream);
    if (k <= 32) {
        k_knn<32, 4><<<dim3(ceil_div(S, 32), B), 256, 0, st>>>(query, ref, S, N, k, out);
    } else if (k <= 64) {
        k_knn<64, 4><<<dim3(ceil_div(S, 32), B), 256, 0, st>>>(query, ref, S, N, k, out);
    } else {
        k_knn<128, 2><<<dim3(ceil_div(S, 16), B), 256, 0, st>>>(query, ref, S, N, k, out);
    }
    return after_launch();
}

static int build_knn_grid(const float *ref, int B, int N, unsigned char *ws, const KnnGridWs &w, cudaStream_t st) {
    k_grid_setup<<<B, 256, 0, st>>>(ref, N, ws, w.per_cloud, w.off_start, w.off_fill);
    int rc = after_launch();
    if (rc) return rc;
    k_grid_count<<<dim3(ceil_div(N, 256), B), 256, 0, st>>>(ref, N, ws, w.per_cloud, w.off_start);
    rc = after_launch();
    if (rc) return rc;
    k_grid_scan<<<B, 1024, 0, st>>>(ws, w.per_cloud, w.off_start);
    rc = after_launch();
    if (rc) return rc;
    k_grid_fill<<<dim3(ceil_div(N, 256), B), 256, 0, st>>>(ref, N, ws, w.per_cloud, w.off_start, w.off_fill, w.off_sorted);
    return after_launch();
}

// farthest_point_sample with cell pruning (k_fps_grid): same indices; one SM per cloud.
int cmr_farthest_point_sample_grid(const float *xyz, const int64_t *start, int B, int N, int npoint, void *workspace, int64_t *out,
                                   void *stream) {
    CMR_REQUIRE(xyz && start && out && workspace && B > 0 && N > 0 && npoint > 0, CMR_EINVAL);
    CMR_REQUIRE(aligned(workspace, 256), CMR_EALIGN);
    const FpsGridSmem m = fps_grid_smem(N);
    CMR_REQUIRE(m.total <= 220 * 1024, CMR_ERANGE);   // the running distances of one cloud live in one SM's shared memory
    const KnnGridWs w = knn_grid_ws(N);
    unsigned char *ws = static_cast<unsigned char *>(workspace);
    cudaStream_t st = S_(stream);
    int rc = build_knn_grid(xyz, B, N, ws, w, st);
    if (rc) return rc;
    rc = allow_smem(k_fps_grid, m.total);
    if (rc) return rc;
    k_fps_grid<<<B, kFpsGridThreads, m.total, st>>>(xyz, start, ws, w.per_cloud, w.off_start, w.off_sorted, N, npoint, m.off_cmax,
                                                    m.off_cidx, m.off_list, m.off_cs, out);
    return after_launch();
}

size_t cmr_knn_grid_workspace_bytes(int B, int N) { return (B > 0 && N > 0) ? (size_t)B * knn_grid_ws(N).per_cloud : 0; }

int cmr_knn_grid(const float *query, const float *ref, int B, int S, int N, int k, void *workspace, int64_t *out, void *stream) {
    CMR_REQUIRE(query && ref && out && workspace && B > 0 && S > 0 && N > 0 && k > 0, CMR_EINVAL);
    CMR_REQUIRE(k <= 128 && k <= N && B <= 65535, CMR_ERANGE);
    CMR_REQUIRE(aligned(workspace, 256), CMR_EALIGN);
    const KnnGridWs w = knn_grid_ws(N);
    unsigned char *ws = static_cast<unsigned char *>(workspace);
    cudaStream_t st = S_(stream);
    int rc = build_knn_grid(ref, B, N, ws, w, st);
    if (rc) return rc;
    const dim3 grid(ceil_div(S, 8), B);
    if (k <= 64)
        k_knn_grid<64><<<grid, 256, 0, st>>>(query, ws, w.per_cloud, w.off_start, w.off_sorted, S, N, k, out);
    else
        k_knn_grid<128><<<grid, 256, 0, st>>>(query, ws, w.per_cloud, w.off_start, w.off_sorted, S, N, k, out);
    return after_launch();
}

int cmr_query_ball_point_grid(const float *query, const float *ref, float radius2, float radius, int nsample, int B, int S, int N,
                              void *workspace, int64_t *out, void *stream) {
    CMR_REQUIRE(query && ref && out && workspace && nsample > 0 && B > 0 && S > 0 && N > 0, CMR_EINVAL);
    CMR_REQUIRE(B <= 65535 && nsample <= 128 && radius >= 0.f, CMR_ERANGE);
    CMR_REQUIRE(aligned(workspace, 256), CMR_EALIGN);
    const KnnGridWs w = knn_grid_ws(N);
    unsigned char *ws = static_cast<unsigned char *>(workspace);
    cudaStream_t st = S_(stream);
    int rc = build_knn_grid(ref, B, N, ws, w, st);
    if (rc) return rc;
    const dim3 grid(ceil_div(S, 8), B);
    if (nsample <= 64)
        k_ball_grid<64><<<grid, 256, 0, st>>>(query, ws, w.per_cloud, w.off_start, w.off_sorted, radius2, radius, nsample, S, N, out);
    else
        k_ball_grid<128><<<grid, 256, 0, st>>>(query, ws, w.per_cloud, w.off_start, w.off_sorted, radius2, radius, nsample, S, N, out);
    return after_launch();
}

int cmr_query_ball_point(const float *query, const float *ref, float radius2, int nsample, int B, int S, int N,
                         int64_t *out, void *stream) {
    CMR_REQUIRE(query && ref && out && nsample > 0 && B > 0 && S > 0 && N > 0, CMR_EINVAL);
    CMR_REQUIRE(B <= 65535, CMR_ERANGE);
    constexpr int kQpw = 4;   // queries per warp
    k_ball_query<kQpw><<<dim3(ceil_div(S, 8 * kQpw), B), 256, 0, S_(stream)>>>(query, ref, radius2, nsample, S, N, out);
    return after_launch();
}

// ------------------------------------------------------------------------------ cost volume ----
// models/IterModel.py:272-351: the same project -> scatter kernels, K candidate poses per cloud.

size_t cmr_cost_volume_workspace_bytes(int B, int K, int N, int C, int P) {
    if (B <= 0 || K <= 0 || N <= 0 || C <= 0 || P <= 0 || (long long)B * K > 65535) return 0;
    return ws_layout(B * K, N, C, P, B).total;
}

int cmr_cost_volume_prepare(const uint8_t *mask, const float *feat, int B, int K, int N, int C, void *workspace,
                            void *stream) {
    CMR_REQUIRE(mask && feat && workspace && B > 0 && K > 0 && N > 0 && C > 0, CMR_EINVAL);
    CMR_REQUIRE(C <= kMaxC && (C % 4) == 0 && N < (1 << 24) && (long long)B * K <= 65535, CMR_ERANGE);
    CMR_REQUIRE(aligned(workspace, 256), CMR_EALIGN);
    WsLayout L = ws_layout(B * K, N, C, 1, B);
    char *ws = static_cast<char *>(workspace);
    cudaStream_t st = S_(stream);
    cudaError_t me = cudaMemsetAsync(ws + L.off_bcnt, 0, L.bcnt_bytes, st);
    if (me == cudaSuccess) me = cudaMemsetAsync(ws + L.off_zero, 0, sizeof(float) * 3 * (size_t)B * K, st);
    if (me != cudaSuccess) return (int)me;
    int *M = reinterpret_cast<int *>(ws + L.off_m);
    int *seg = reinterpret_cast<int *>(ws + L.off_seg);
    float *featT = reinterpret_cast<float *>(ws + L.off_feat);
    k_overlap_scan<<<B, 1024, 0, st>>>(mask, N, L.groups, (N % 4 == 0) && aligned(mask, 4), seg, M);
    int rc = after_launch();
    if (rc) return rc;
    return launch_feat_compact(mask, feat, B, N, C, L.groups, seg, featT, st);
}

int cmr_cost_volume_warp(const float *pc, const uint8_t *mask, const float *Kmat, const float *poses, void *workspace,
                         int B, int K, int N, int C, int H, int W, int mean_channels, float *out, void *stream) {
    CMR_REQUIRE(pc && mask && Kmat && poses && workspace && out && K > 0, CMR_EINVAL);
    CMR_REQUIRE(aligned(poses, 16), CMR_EALIGN);
    int rc = check_observe_dims(B, N, C, H, W);
    if (rc) return rc;
    CMR_REQUIRE((long long)B * K <= 65535, CMR_ERANGE);
    CMR_REQUIRE(mean_channels >= C || mean_channels % kSlab == 0, CMR_EUNSUPPORTED);   // whole slabs are means or sums
    CMR_REQUIRE(aligned(workspace, 256), CMR_EALIGN);
    const int E = B * K, P = H * W;
    WsLayout L = ws_layout(E, N, C, P, B);
    CMR_REQUIRE(bucket_path(L, C), CMR_EUNSUPPORTED);   // grids of up to 12288 pixels
    char *ws = static_cast<char *>(workspace);
    const float *zero_mean = reinterpret_cast<const float *>(ws + L.off_zero);   // X = R p + t: nothing is subtracted
    cudaStream_t st = S_(stream);
    // The reference's shape (64 mean channels + the summed score, a grid of whole 32-pixel buckets that fits the
    // shared-memory histogram): one CTA per pose does everything (cost_volume_kernels.cuh).  The bucket buffers are not
    // used on this path; its scratch arrays (compacted coordinates, pixel ids, the sorted lists) live in their place.
    {
        const size_t need = round_up(sizeof(float4) * (size_t)B * L.ncap, 256) + 3 * round_up(sizeof(uint16_t) * (size_t)E * L.ncap, 256) +
                            round_up(sizeof(uint16_t) * (size_t)E * (P + 2), 256);
        const bool fused = C == kCvFeat + 4 && mean_channels == kCvFeat && P % 32 == 0 && P <= kCvMaxP && L.ncap <= 65535 &&
                           need <= L.total - L.off_bbuf && !cv_force_buckets();
        if (fused) {
            char *o = ws + L.off_bbuf;
            float4 *xyzc = reinterpret_cast<float4 *>(o);
            o += round_up(sizeof(float4) * (size_t)B * L.ncap, 256);
            uint16_t *pix16 = reinterpret_cast<uint16_t *>(o);
            o += round_up(sizeof(uint16_t) * (size_t)E * L.ncap, 256);
            uint16_t *tmp16 = reinterpret_cast<uint16_t *>(o);
            o += round_up(sizeof(uint16_t) * (size_t)E * L.ncap, 256);
            uint16_t *sorted16 = reinterpret_cast<uint16_t *>(o);
            o += round_up(sizeof(uint16_t) * (size_t)E * L.ncap, 256);
            uint16_t *gstart = reinterpret_cast<uint16_t *>(o);
            const bool vec = (N % 4 == 0) && aligned(mask, 4);
            k_xyz_compact<<<dim3(ceil_div(L.groups, 8), B), 256, 0, st>>>(pc, mask, reinterpret_cast<const int *>(ws + L.off_seg), N, L.ncap,
                                                                         L.groups, vec, xyzc);
            rc = after_launch();
            if (rc) return rc;
            rc = allow_smem(k_cost_volume_sort, kCvSmem);
            if (rc) return rc;
            k_cost_volume_sort<<<E, kCvSortThreads, kCvSmem, st>>>(xyzc, reinterpret_cast<const int *>(ws + L.off_m), Kmat, poses, zero_mean,
                                                               L.ncap, H, W, K, N >= kBmmChainMinCols, pix16, tmp16, sorted16, gstart);
            rc = after_launch();
            if (rc) return rc;
            rc = allow_smem(k_cost_volume_gather, kCvTileSmem);
            if (rc) return rc;
            return launch_pdl(k_cost_volume_gather, dim3(ceil_div(P / 32, kCvWarps * kCvPerWarp), E), dim3(kCvThreads), kCvTileSmem, st,
                              reinterpret_cast<const float *>(ws + L.off_feat), N, L.ncap, C, P, K, (const uint16_t *)sorted16,
                              (const uint16_t *)gstart, out);
        }
    }
    // any other shape: only the masked points are projected (k_project_masked) into the observation's bucket buffers
    {
        const int *seg = reinterpret_cast<const int *>(ws + L.off_seg);
        int *bcnt = reinterpret_cast<int *>(ws + L.off_bcnt);
        unsigned *bbuf = reinterpret_cast<unsigned *>(ws + L.off_bbuf);
        int *hq = reinterpret_cast<int *>(ws + L.off_hq);
        int *hdr = bcnt + (size_t)E * kBucketStride;
        const bool vec = (N % 4 == 0) && aligned(mask, 4);
        const dim3 grid(ceil_div(L.groups, kProjWarps), E), block(32 * kProjWarps);
        if (L.pix16)
            rc = launch_pdl(k_project_masked<uint16_t>, grid, block, 0, st, pc, mask, Kmat, poses, zero_mean, seg, N, L.ncap,
                            L.groups, H, W, vec, reinterpret_cast<uint16_t *>(ws + L.off_pix), bcnt, bbuf, L.buckets, hdr, hq, K);
        else
            rc = launch_pdl(k_project_masked<int32_t>, grid, block, 0, st, pc, mask, Kmat, poses, zero_mean, seg, N, L.ncap,
                            L.groups, H, W, vec, reinterpret_cast<int32_t *>(ws + L.off_pix), bcnt, bbuf, L.buckets, hdr, hq, K);
    }
    if (rc) return rc;
    return launch_gather(L, ws, nullptr, E, N, C, P, false, out, st, K, C, 0, mean_channels, kHeavyFrom);
}

// ------------------------------------------------------------------------------ bilinear sampling ----
// image features at the points' projections (north_star; SURVEY.md D1: an extra operator, oracle F.grid_sample)

size_t cmr_sample_workspace_bytes(int B, int C, int P) {
    if (B <= 0 || C <= 0 || P <= 0) return 0;
    return round_up(sizeof(float) * (size_t)B * C * P, 256);
}

int cmr_sample_prepare(const float *img_feat, int B, int C, int P, void *workspace, void *stream) {
    CMR_REQUIRE(img_feat && workspace && B > 0 && C > 0 && P > 0, CMR_EINVAL);
    CMR_REQUIRE(B <= 65535 && ceil_div(C, 32) <= 65535, CMR_ERANGE);
    CMR_REQUIRE(aligned(workspace, 256), CMR_EALIGN);
    k_image_transpose<<<dim3(ceil_div(P, 32), ceil_div(C, 32), B), 256, 0, S_(stream)>>>(img_feat, C, P, static_cast<float *>(workspace));
    return after_launch();
}

int cmr_sample_image_features(const float *pc, const float *Kmat, const float *pose, const float *mean, const void *workspace,
                              int B, int N, int C, int H, int W, float *out, uint8_t *in_cam, void *stream) {
    CMR_REQUIRE(pc && Kmat && pose && mean && workspace && out && B > 0 && N > 0 && C > 0 && H > 0 && W > 0, CMR_EINVAL);
    CMR_REQUIRE(B <= 65535 && (long long)H * W < (1LL << 31) / 4 && (C % 2) == 0, CMR_ERANGE);
    CMR_REQUIRE(aligned(pose, 16) && aligned(workspace, 256), CMR_EALIGN);
    int rc = allow_smem(k_bilinear_sample, kSmpSmem);
    if (rc) return rc;
    const bool vec = (N % 4 == 0) && aligned(out, 16);
    k_bilinear_sample<<<dim3(ceil_div(N, 32 * kSmpWarps), B), kSmpThreads, kSmpSmem, S_(stream)>>>(
        pc, Kmat, pose, mean, static_cast<const float *>(workspace), N, C, H, W, vec, out, in_cam);
    return after_launch();
}

// ------------------------------------------------------------------------------ dataset side ----

int cmr_fps_f64(const double *pts, const int64_t *start, int B, int M, int k, int64_t *out_idx, double *out_pts,
                void *stream) {
    CMR_REQUIRE(pts && start && out_idx && B > 0 && M > 0 && k > 0, CMR_EINVAL);
    CMR_REQUIRE(k <= M && (long long)M <= (long long)kFps64Threads * kFps64MaxPpt, CMR_ERANGE);
    const int ppt = ceil_div(M, kFps64Threads);
    cudaStream_t st = S_(stream);
    const size_t smem = sizeof(double) * 2 * (size_t)M;   // x and y of the cloud, resident for all rounds
    const bool stage = smem <= 200 * 1024;
#define CMR_FPS64_LAUNCH(PPT)                                                                                    \
    do {                                                                                                         \
        if (stage) {                                                                                             \
            int rc = allow_smem(k_fps_f64<PPT, true>, smem);                                                     \
            if (rc) return rc;                                                                                   \
            k_fps_f64<PPT, true><<<B, kFps64Threads, smem, st>>>(pts, start, M, k, out_idx, out_pts);            \
        } else {                                                                                                 \
            k_fps_f64<PPT, false><<<B, kFps64Threads, 0, st>>>(pts, start, M, k, out_idx, out_pts);              \
        }                                                                                                        \
    } while (0)
    if (ppt <= 4) CMR_FPS64_LAUNCH(4);
    else if (ppt <= 8) CMR_FPS64_LAUNCH(8);
    else if (ppt <= 12) CMR_FPS64_LAUNCH(12);
    else CMR_FPS64_LAUNCH(16);
#undef CMR_FPS64_LAUNCH
    return after_launch();
}

int cmr_nearest_f64(const double *query, const double *ref, int B, int N, int S, int64_t *out, void *stream) {
    CMR_REQUIRE(query && ref && out && B > 0 && N > 0 && S > 0, CMR_EINVAL);
    CMR_REQUIRE(B <= 65535, CMR_ERANGE);
    k_nearest_f64<<<dim3(ceil_div(N, 256), B), 256, 0, S_(stream)>>>(query, ref, N, S, out);
    return after_launch();
}

// ------------------------------------------------------------------------------ rollout session ----

#define CMR_CUDA(call)                       \
    do {                                     \
        cudaError_t e_ = (call);             \
        if (e_ != cudaSuccess) return (int)e_; \
    } while (0)

static void session_free(cmr_session *s) {
    if (!s) return;
    if (s->slots) {
        for (int i = 0; i < s->depth; ++i) {
            SessionSlot &t = s->slots[i];
            void *dev[] = {t.pc, t.feat, t.img_feat, t.K, t.target_pose, t.pc_in_cam, t.overlap, t.mask_u8, t.mask_i64, t.a_r, t.a_t,
                           t.mean, t.pose, t.obs2d, t.obs3d, t.rew, t.dist, t.ws, t.scratch};
            for (void *p : dev)
                if (p) cudaFree(p);
            void *host[] = {t.h_rew, t.h_dist, t.h_pose, t.h_target};
            for (void *p : host)
                if (p) cudaFreeHost(p);
            cudaEvent_t ev[] = {t.uploaded, t.done, t.up_begin, t.up_end};
            for (cudaEvent_t e : ev)
                if (e) cudaEventDestroy(e);
        }
        delete[] s->slots;
    }
    if (s->rot_tab) cudaFree(s->rot_tab);
    if (s->t_tab) cudaFree(s->t_tab);
    if (s->copy) cudaStreamDestroy(s->copy);
    if (s->compute) cudaStreamDestroy(s->compute);
    delete s;
}

int cmr_session_create(const cmr_session_config *cfg, cmr_session **out) {
    CMR_REQUIRE(cfg && out && cfg->rot_tab && cfg->t_tab, CMR_EINVAL);
    CMR_REQUIRE(cfg->iters > 0 && cfg->depth >= 1 && cfg->depth <= 8 && cfg->nbins > 0, CMR_EINVAL);
    int rc = check_observe_dims(cfg->B, cfg->N, cfg->C, cfg->H, cfg->W);
    if (rc) return rc;
    cmr_session *s = new (std::nothrow) cmr_session();
    CMR_REQUIRE(s, CMR_EINVAL);
    s->cfg = *cfg;
    s->cfg.rot_tab = s->cfg.t_tab = nullptr;
    s->depth = cfg->depth;
    s->nbins = cfg->nbins;
    s->slots = new (std::nothrow) SessionSlot[cfg->depth];
    const size_t B = cfg->B, N = cfg->N, C = cfg->C, P = (size_t)cfg->H * cfg->W, I = cfg->iters;
    const int nr = cfg->dof6 ? 3 : 1, nt = cfg->dof6 ? 3 : 2;
    auto fail = [&](int code) {
        session_free(s);
        return code;
    };
    if (!s->slots) return fail(CMR_EINVAL);
#define CMR_S(call)                                  \
    do {                                             \
        cudaError_t e_ = (call);                     \
        if (e_ != cudaSuccess) return fail((int)e_); \
    } while (0)
    CMR_S(cudaStreamCreateWithFlags(&s->copy, cudaStreamNonBlocking));
    CMR_S(cudaStreamCreateWithFlags(&s->compute, cudaStreamNonBlocking));
    const size_t rot_bytes = sizeof(float) * 3 * (cfg->nbins + 1) * 9;
    CMR_S(cudaMalloc(&s->rot_tab, rot_bytes));
    CMR_S(cudaMalloc(&s->t_tab, sizeof(float) * cfg->nbins));
    CMR_S(cudaMemcpy(s->rot_tab, cfg->rot_tab, rot_bytes, cudaMemcpyHostToDevice));
    CMR_S(cudaMemcpy(s->t_tab, cfg->t_tab, sizeof(float) * cfg->nbins, cudaMemcpyHostToDevice));
    for (int i = 0; i < s->depth; ++i) {
        SessionSlot &t = s->slots[i];
        CMR_S(cudaMalloc(&t.pc, sizeof(float) * B * 3 * N));
        if (!cfg->features_resident) {
            CMR_S(cudaMalloc(&t.feat, sizeof(float) * B * C * N));
            CMR_S(cudaMalloc(&t.img_feat, sizeof(float) * B * C * P));
        }
        CMR_S(cudaMalloc(&t.K, sizeof(float) * B * 9));
        CMR_S(cudaMalloc(&t.target_pose, sizeof(float) * B * 16));
        CMR_S(cudaMalloc(&t.pc_in_cam, sizeof(float) * B * 3 * N));
        CMR_S(cudaMalloc(&t.overlap, B * N));
        CMR_S(cudaMalloc(&t.mask_u8, B * N));
        CMR_S(cudaMalloc(&t.mask_i64, sizeof(long long) * B * N));
        CMR_S(cudaMalloc(&t.a_r, sizeof(long long) * I * B * nr));
        CMR_S(cudaMalloc(&t.a_t, sizeof(long long) * I * B * nt));
        CMR_S(cudaMalloc(&t.mean, sizeof(float) * B * 3));
        CMR_S(cudaMalloc(&t.pose, sizeof(float) * B * 16));
        CMR_S(cudaMalloc(&t.obs2d, sizeof(float) * B * 2 * C * P));
        CMR_S(cudaMalloc(&t.obs3d, sizeof(float) * B * 5 * N));
        CMR_S(cudaMalloc(&t.rew, sizeof(float) * I * B));
        CMR_S(cudaMalloc(&t.dist, sizeof(float) * I * B));
        CMR_S(cudaMalloc(&t.ws, cmr_workspace_bytes(cfg->B, cfg->N, cfg->C, (int)P)));
        CMR_S(cudaMalloc(&t.scratch, cmr_reward_scratch_bytes(cfg->B)));
        CMR_S(cudaMemset(t.scratch, 0, cmr_reward_scratch_bytes(cfg->B)));
        CMR_S(cudaHostAlloc(&t.h_rew, sizeof(float) * I * B, cudaHostAllocDefault));
        CMR_S(cudaHostAlloc(&t.h_dist, sizeof(float) * I * B, cudaHostAllocDefault));
        CMR_S(cudaHostAlloc(&t.h_pose, sizeof(float) * B * 16, cudaHostAllocDefault));
        CMR_S(cudaHostAlloc(&t.h_target, sizeof(float) * B * 16, cudaHostAllocDefault));
        CMR_S(cudaEventCreateWithFlags(&t.uploaded, cudaEventDisableTiming));
        CMR_S(cudaEventCreateWithFlags(&t.done, cudaEventDisableTiming));
        CMR_S(cudaEventCreate(&t.up_begin));
        CMR_S(cudaEventCreate(&t.up_end));
    }
#undef CMR_S
    s->bytes_per_upload = sizeof(float) * (B * 3 * N * 2 + B * 9 + B * 16) + B * N + sizeof(long long) * (B * N + I * B * (nr + nt)) +
                          (cfg->features_resident ? 0 : sizeof(float) * (B * C * N + B * C * P));
    *out = s;
    return CMR_OK;
}

void cmr_session_destroy(cmr_session *s) {
    if (!s) return;
    if (s->compute) cudaStreamSynchronize(s->compute);
    if (s->copy) cudaStreamSynchronize(s->copy);
    session_free(s);
}

static void session_account_upload(cmr_session *s, SessionSlot &t) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, t.up_begin, t.up_end) == cudaSuccess && ms > 0.f) {
        s->h2d_bytes += (double)s->bytes_per_upload;
        s->h2d_seconds += ms * 1e-3;
        ++s->timed_uploads;
    }
}

int cmr_session_submit(cmr_session *s, const cmr_rollout_inputs *in, long long *ticket) {
    CMR_REQUIRE(s && in && ticket, CMR_EINVAL);
    CMR_REQUIRE(in->pc && in->overlap && in->feat && in->img_feat && in->K && in->P && in->pc_in_cam && in->pc_mask && in->action_r &&
                    in->action_t,
                CMR_EINVAL);
    const cmr_session_config &c = s->cfg;
    const size_t B = c.B, N = c.N, C = c.C, P = (size_t)c.H * c.W, I = c.iters;
    const int nr = c.dof6 ? 3 : 1, nt = c.dof6 ? 3 : 2;
    SessionSlot &t = s->slots[s->submitted % s->depth];
    if (t.in_flight) {   // the slot's previous rollout has not been collected: it must at least be complete
        CMR_CUDA(cudaEventSynchronize(t.done));
        session_account_upload(s, t);
        t.in_flight = false;
    }
    // ---- uploads (copy stream).  The slot's buffers are free: its previous rollout is complete (above / wait).
    CMR_CUDA(cudaEventRecord(t.up_begin, s->copy));
    const cudaMemcpyKind h2d = cudaMemcpyHostToDevice;
    CMR_CUDA(cudaMemcpyAsync(t.pc, in->pc, sizeof(float) * B * 3 * N, h2d, s->copy));
    CMR_CUDA(cudaMemcpyAsync(t.overlap, in->overlap, B * N, h2d, s->copy));
    CMR_CUDA(cudaMemcpyAsync(t.K, in->K, sizeof(float) * B * 9, h2d, s->copy));
    CMR_CUDA(cudaMemcpyAsync(t.target_pose, in->P, sizeof(float) * B * 16, h2d, s->copy));
    CMR_CUDA(cudaMemcpyAsync(t.a_r, in->action_r, sizeof(long long) * I * B * nr, h2d, s->copy));
    CMR_CUDA(cudaMemcpyAsync(t.a_t, in->action_t, sizeof(long long) * I * B * nt, h2d, s->copy));
    if (!c.features_resident) {
        CMR_CUDA(cudaMemcpyAsync(t.feat, in->feat, sizeof(float) * B * C * N, h2d, s->copy));
        CMR_CUDA(cudaMemcpyAsync(t.img_feat, in->img_feat, sizeof(float) * B * C * P, h2d, s->copy));
        t.feat_used = t.feat;
        t.img_used = t.img_feat;
    } else {
        t.feat_used = in->feat;
        t.img_used = in->img_feat;
    }
    CMR_CUDA(cudaMemcpyAsync(t.pc_in_cam, in->pc_in_cam, sizeof(float) * B * 3 * N, h2d, s->copy));
    CMR_CUDA(cudaMemcpyAsync(t.mask_i64, in->pc_mask, sizeof(long long) * B * N, h2d, s->copy));
    CMR_CUDA(cudaEventRecord(t.up_end, s->copy));
    CMR_CUDA(cudaEventRecord(t.uploaded, s->copy));
    // ---- the rollout (compute stream)
    cudaStream_t st = s->compute;
    CMR_CUDA(cudaStreamWaitEvent(st, t.uploaded, 0));
    int rc = cmr_cloud_mean(t.pc, c.B, c.N, t.mean, st);                                  // environment.py:46 - once per episode
    if (!rc) rc = cmr_episode_prepare(t.overlap, t.feat_used, c.B, c.N, c.C, t.ws, st);
    if (rc) return rc;
    k_pose_identity<<<ceil_div(c.B * 16, 256), 256, 0, st>>>(t.pose, c.B);               // env.init (:138)
    rc = after_launch();
    if (!rc) rc = cmr_to_disentangled(t.target_pose, t.mean, c.B, st);                    // Test_Agent.py:152
    if (rc) return rc;
    k_mask_to_u8<<<std::min<long long>((long long)(B * N + 255) / 256, 4096), 256, 0, st>>>(t.mask_i64, t.mask_u8, B * N);   // :268
    rc = after_launch();
    for (int it = 0; it < c.iters && !rc; ++it) {
        rc = cmr_observe(t.pc, t.overlap, t.img_used, t.K, t.pose, t.mean, t.ws, c.B, c.N, c.C, c.H, c.W, t.obs2d, t.obs3d, nullptr,
                         nullptr, st);
        if (!rc) rc = cmr_step(t.pose, (const int64_t *)t.a_r + (size_t)it * B * nr, (const int64_t *)t.a_t + (size_t)it * B * nt,
                               s->rot_tab, s->t_tab, s->nbins, c.dof6, c.B, st);
        if (!rc) {
            const float *prev = it ? t.dist + (size_t)(it - 1) * B : nullptr;
            if (c.reward_mode == CMR_REWARD_SHIPPED && it > 0)   // the shipped distance is a constant of the batch (:272-275)
                rc = cmr_reward_compare(t.dist, prev, c.B, t.rew + (size_t)it * B, t.dist + (size_t)it * B, st);
            else
                rc = cmr_reward(t.pc_in_cam, t.pc, t.mask_u8, t.mean, t.pose, prev, c.reward_mode, c.B, c.N, t.scratch,
                                t.rew + (size_t)it * B, t.dist + (size_t)it * B, st);
        }
    }
    if (rc) return rc;
    const cudaMemcpyKind d2h = cudaMemcpyDeviceToHost;
    CMR_CUDA(cudaMemcpyAsync(t.h_rew, t.rew, sizeof(float) * I * B, d2h, st));
    CMR_CUDA(cudaMemcpyAsync(t.h_dist, t.dist, sizeof(float) * I * B, d2h, st));
    CMR_CUDA(cudaMemcpyAsync(t.h_pose, t.pose, sizeof(float) * B * 16, d2h, st));
    CMR_CUDA(cudaMemcpyAsync(t.h_target, t.target_pose, sizeof(float) * B * 16, d2h, st));
    CMR_CUDA(cudaEventRecord(t.done, st));
    t.in_flight = true;
    *ticket = s->submitted++;
    return CMR_OK;
}

int cmr_session_wait(cmr_session *s, long long ticket, float *rewards, float *dists, float *poses, float *target_poses) {
    CMR_REQUIRE(s && ticket >= 0 && ticket < s->submitted && ticket >= s->submitted - s->depth, CMR_EINVAL);
    SessionSlot &t = s->slots[ticket % s->depth];
    CMR_CUDA(cudaEventSynchronize(t.done));
    if (t.in_flight) {
        session_account_upload(s, t);
        t.in_flight = false;
    }
    const size_t B = s->cfg.B, I = s->cfg.iters;
    if (rewards) memcpy(rewards, t.h_rew, sizeof(float) * I * B);
    if (dists) memcpy(dists, t.h_dist, sizeof(float) * I * B);
    if (poses) memcpy(poses, t.h_pose, sizeof(float) * B * 16);
    if (target_poses) memcpy(target_poses, t.h_target, sizeof(float) * B * 16);
    return CMR_OK;
}

int cmr_session_stats(const cmr_session *s, double *h2d_gbs, double *bytes_per_rollout) {
    CMR_REQUIRE(s, CMR_EINVAL);
    if (h2d_gbs) *h2d_gbs = s->h2d_seconds > 0 ? s->h2d_bytes / s->h2d_seconds / 1e9 : 0.0;
    if (bytes_per_rollout) *bytes_per_rollout = (double)s->bytes_per_upload;
    return CMR_OK;
}

int cmr_session_last_observation(cmr_session *s, long long ticket, float *obs2d, float *obs3d, void *stream) {
    CMR_REQUIRE(s && ticket >= 0 && ticket < s->submitted && ticket >= s->submitted - s->depth, CMR_EINVAL);
    SessionSlot &t = s->slots[ticket % s->depth];
    const size_t B = s->cfg.B, N = s->cfg.N, C = s->cfg.C, P = (size_t)s->cfg.H * s->cfg.W;
    CMR_CUDA(cudaStreamWaitEvent(S_(stream), t.done, 0));
    if (obs2d) CMR_CUDA(cudaMemcpyAsync(obs2d, t.obs2d, sizeof(float) * B * 2 * C * P, cudaMemcpyDeviceToDevice, S_(stream)));
    if (obs3d) CMR_CUDA(cudaMemcpyAsync(obs3d, t.obs3d, sizeof(float) * B * 5 * N, cudaMemcpyDeviceToDevice, S_(stream)));
    return CMR_OK;
}

// -------------------------------------------------------------------------------- 3-D tower ----
// models/CMRAgent.py:25-29,92-101 (eval mode): tower_kernels.cuh

size_t cmr_tower_blob_bytes(int kind) {
    switch (kind) {
        case CMR_TOWER_FIRST: return TowerBlobFirst::total;
        case CMR_TOWER_MID: return TowerBlobMid::total;
        case CMR_TOWER_LAST: return TowerBlobLast::total;
        default: return 0;
    }
}

int cmr_tower_pack(int kind, const float *W1, const float *b1, const float *W2, const float *b2, const float *Ws,
                   const float *bs, void *blob, void *stream) {
    CMR_REQUIRE(kind >= CMR_TOWER_FIRST && kind <= CMR_TOWER_LAST, CMR_EINVAL);
    CMR_REQUIRE(W1 && b1 && W2 && b2 && blob && (kind == CMR_TOWER_LAST || (Ws && bs)), CMR_EINVAL);
    CMR_REQUIRE(aligned(blob, 128), CMR_EALIGN);
    k_tower_pack<<<32, 256, 0, S_(stream)>>>(kind, W1, b1, W2, b2, Ws, bs, static_cast<unsigned char *>(blob));
    return after_launch();
}

size_t cmr_tower_workspace_bytes(int B, int N) { return (B > 0 && N > 0) ? tower_ws(B, N).total : 0; }

int cmr_tower_forward(const float *obs3d, const void *blob1, const void *blob2, const void *blob3, const void *blob4,
                      void *workspace, int B, int N, float *embed, void *stream) {
    CMR_REQUIRE(obs3d && blob1 && blob2 && blob3 && blob4 && workspace && embed && B > 0 && N > 0, CMR_EINVAL);
    CMR_REQUIRE(B <= 65535 && N < (1 << 24), CMR_ERANGE);
    CMR_REQUIRE(aligned(workspace, 1024) && aligned(blob2, 128) && aligned(blob3, 128) && aligned(blob4, 128) && aligned(blob1, 16),
                CMR_EALIGN);
    cudaStream_t st = S_(stream);
    const TowerWs w = tower_ws(B, N);
    char *ws = static_cast<char *>(workspace);
    char *plane[4];
    for (int i = 0; i < 4; ++i) plane[i] = ws + i * w.plane;
    unsigned *keys1 = reinterpret_cast<unsigned *>(ws + w.off_keys), *keys2 = keys1 + (size_t)B * 64, *keys3 = keys2 + (size_t)B * 64,
             *keys4 = keys3 + (size_t)B * 64;
    cudaError_t e = cudaMemsetAsync(keys1, 0, w.keys_bytes, st);
    if (e != cudaSuccess) return (int)e;
    const int tiles_per_ep = ceil_div(N, kTowerTile);
    const uint32_t rows = (uint32_t)std::min(N, kTowerTile);
    alignas(64) CUtensorMap m[4];
    for (int i = 0; i < 4; ++i)
        if (!make_plane_map(&m[i], plane[i], (uint64_t)N, (uint64_t)B, rows)) return CMR_EUNSUPPORTED;
    // block 1 (five hidden channels on the fp32 pipes, the 64 outputs as one K = 16 GEMM): obs3d -> planes 0,1
    {
        const size_t smem = FirstCfg::smem_bytes;
        int rc = allow_smem(k_tower_first, smem);
        if (rc) return rc;
        const int grid = std::min(B * tiles_per_ep, sm_count());
        rc = launch_pdl(k_tower_first, dim3(grid), dim3(kFirstThreads), smem, st, obs3d, static_cast<const unsigned char *>(blob1), B, N,
                        tiles_per_ep, m[0], m[1], keys1);
        if (rc) return rc;
    }
    // block 2: planes 0,1 -> 2,3;  block 3: planes 2,3 -> 0,1;  block 4: planes 0,1 -> keys
    int rc = launch_tower_mma<false>(blob2, B, N, tiles_per_ep, (int)rows, m[0], m[1], m[2], m[3], keys1, keys2, st);
    if (rc) return rc;
    rc = launch_tower_mma<false>(blob3, B, N, tiles_per_ep, (int)rows, m[2], m[3], m[0], m[1], keys2, keys3, st);
    if (rc) return rc;
    rc = launch_tower_mma<true>(blob4, B, N, tiles_per_ep, (int)rows, m[0], m[1], m[0], m[1], keys3, keys4, st);
    if (rc) return rc;
    return launch_pdl(k_tower_finish, dim3(ceil_div(B * 128, 256)), dim3(256), 0, st, (const unsigned *)keys4, embed, B * 128);
}

#ifdef CMR_DBG_TIMING
__attribute__((visibility("default"))) int cmr_debug_read(void *dst, size_t bytes) {
    cudaDeviceSynchronize();
    return (int)cudaMemcpyFromSymbol(dst, g_dbg, bytes);
}
#endif

}  // extern "C"
