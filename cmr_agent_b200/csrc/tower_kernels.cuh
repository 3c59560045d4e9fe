// tower_kernels.cuh - the agent's 3-D tower on tcgen05 tensor cores + TMEM.
//
// Reference: CMRAgent.forward, models/CMRAgent.py:92-101, over state_3d_embed (:25-29): four ConvBNReLURes1D
// blocks (models/PointNN.py:260-282; 5->64, 128->64, 128->64, 128->128 channels, 1x1 convolutions) with a global
// max over the N points of an episode between them.  Eval mode: every BatchNorm1d is folded into the convolution
// before it (host side, cmr_agent_b200/agent_tower.py; algebra pinned by oracle/tower_oracle.py).
//
// Algebra (oracle/tower_oracle.py):  the input of blocks 2-4 is cat([feat (64), max_prev.repeat(N) (64)]), so the
// half of every first-layer product that meets the repeated max is a PER-EPISODE BIAS  W[:, 64:] @ max_prev.
//   block 1 (k_tower_first: h on the fp32 pipes, the 64 outputs as ONE K = 16 tensor-core GEMM over (x, h, 1)):
//       h = lrelu(W1 x + b1) (5 ch);  out = lrelu(W2 h + Ws x + (b2 + bs))              -> feat1 [64]
//   blocks 2, 3 (k_tower_mma<false>):
//       h = lrelu(W1a feat + bias1_e)           (128 ch;  bias1_e = b1 + W1b max_prev)
//       out = lrelu(W2 h + Wsa feat + bias2_e)  ( 64 ch;  bias2_e = b2 + bs + Wsb max_prev)
//   block 4 (k_tower_mma<true>; identity shortcut = the concatenated input itself):
//       h as above;  out[c] = lrelu((W2 h)[c] + b2[c] + (c < 64 ? feat[c] : max_prev[c-64]))   (128 ch)
//   after every block: max over the episode's points per channel -> ordered-uint keys, atomicMax.
//
// Precision.  north_star's bar is 1e-5 (read on the output's scale, tests/test_tower_oracle.py); single-pass bf16
// or tf32 operands are 1e-3 off.  Every GEMM therefore runs as three 16-bit passes over split operands,
// x = hi + lo,  hi = fp16(x), lo = fp16(x - hi)  (22 significant bits):  hi*hi + hi*lo + lo*hi, fp32 accumulation in
// TMEM; the dropped lo*lo term and the representation error are 2^-22 of |x||w| each.
//   * measured on a B200: bf16 pieces (16-17 bits) give 5e-6 typical but 1.2e-5 worst over 300 episodes -
//     not enough; kind::f16 rejects an fp16 operand meeting a bf16 one (illegal instruction), so "bf16 hi, fp16 lo"
//     is not available; fp16 pieces cost the same three passes.
//   * range: fp16 holds |x| <= 65504.  A larger activation converts to inf, its low piece to -inf, their products
//     to NaN, and every max on the way is NaN-propagating (max.NaN): the embedding comes out NaN and
//     k_tower_finish raises the sticky fault word (cmr_take_fault() == 3) - loud, never silently wrong.
//     The tower's inputs are coordinates in metres and 0/1 flags; CMR_TOWER_FMT=0 builds the bf16 variant.
//
// Orientation: M = 128 points (TMEM lanes), N = output channels (TMEM columns), K = input channels.
//   A operand = activations: `feat` tiles arrive by TMA as two 16-bit planes [128 pts][64 ch] (128-byte rows,
//       128B swizzle: exactly the K-major UMMA layout), h is written BACK INTO TMEM by the epilogue as packed bf16
//       pairs over the accumulator it was read from, and the second GEMM takes its A operand from TMEM.
//   B operand = weights, pre-split and pre-swizzled once by k_tower_pack, resident in shared memory.
// Features travel between the blocks as two fp16 planes [B][N][64] (hi, lo) - 4 bytes per value, as fp32 would be;
// block 4 adds them back (hi + lo, 22 bits) for its identity shortcut.
//
// Warp roles of k_tower_mma (608 threads):
//   warp 0       TMA loads: the weights once, then a ring of input stages (two 16 KB planes each)
//   warp 1       TMEM allocation + MMA issue.  The whole warp runs converged and every tcgen05.mma elects its lane
//                itself (see mma_ss).  Order: C1(0), C1(1), then per tile i: C2(i), C1(i + 2) - the tensor pipe
//                executes in issue order, so the first GEMM of tile i + 2 needs no barrier against the second GEMM
//                of tile i whose H columns it overwrites.  Mid blocks: one N = 192 MMA per K-chunk covers GEMM 1
//                and the shortcut product, and tcgen05.commit hands the input stage back to the TMA warp.
//   warps 2-17   epilogue, ONE software-pipelined group: warp = (TMEM lane quarter) x (column quarter).  Per tile
//                they run E1(i + 1) (D1 -> h -> fp16 pairs back into TMEM) BEFORE E2(i) (D2 -> output, maxima), so
//                the second GEMM of tile i has a whole E1 to complete.  TMEM holds two tiles (2 x 256 columns; mid
//                blocks: [D2a 64][D1 128][D2b 64] with the D2 buffers alternating).  Running maxima stay per point
//                lane in registers and cross the warp once per episode.
//   warp 18      TMA stores (mid blocks): takes each staged output tile over through an mbarrier.
// A CTA owns a contiguous range of 128-point tiles; when the range crosses into the next episode the epilogue
// recomputes the per-episode biases (double-buffered by episode parity, E1 may be an episode ahead of E2).
// How it got here (ncu, profiles/r2_tower_ncu.txt; cycles per 128-point tile of a mid block): 4550 with MMAs issued
// from an `if (lane == 0)` region (the issuing thread was the bottleneck), 3500 with elected issue + N = 192,
// 3400 with C1 issued behind C2, 3050 with the pipelined epilogue group, register maxima and the store warp.
// The floors: tensor pipe 1920, HBM 2140, shared-memory traffic ~2060, issue slots ~1700 - all 60-65 % busy.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "common.cuh"

namespace cmr {

constexpr int kTowerF = 64;            // embed_dim (config/KittiConfig.py:63)
constexpr int kTowerTile = 128;        // points per tile = UMMA M
constexpr int kTowerThreads = 608;      // TMA-load warp + MMA warp + 16 epilogue warps + TMA-store warp
constexpr float kTowerSlope = 0.2f;    // LeakyReLU(negative_slope=0.2), PointNN.py:267,272
#ifndef CMR_TOWER_PASSES
#define CMR_TOWER_PASSES 3
#endif
#ifndef CMR_TOWER_FMT
#define CMR_TOWER_FMT 1
#endif
#ifndef CMR_TOWER_REGMAX
#define CMR_TOWER_REGMAX 0
#endif
#ifndef CMR_TOWER_PACKED
#define CMR_TOWER_PACKED 1
#endif
constexpr bool kTowerRegMax = CMR_TOWER_REGMAX != 0;   // running maxima per point lane in registers (measured SLOWER: 64 more
                                                       // registers cost occupancy / spills - off); else: one 31-step exchange per tile
constexpr bool kTowerPacked = CMR_TOWER_PACKED != 0;   // FADD2/FMUL2/FFMA2 in the epilogues
constexpr int kTowerPasses = CMR_TOWER_PASSES;   // 3: hh + hl + lh;  4: + ll
constexpr bool kTowerF16 = CMR_TOWER_FMT == 1;   // pieces are fp16 (1, default) or bf16 (0)

// ---- packed weights: byte offsets inside the blob k_tower_pack writes (one blob per block) ------------------------
// mid block (blocks 2, 3):   {Wsa [64 x 64]; W1a [128 x 64]; Wsa again} hi, the same lo: rows 0-191 and rows 64-255 are each
//                            ONE B operand (N = 192) for the first GEMM plus the shortcut product, with the shortcut's
//                            columns before or after D1's (the two D2 buffers of an epilogue group sit either side
//                            of its D1); W2 hi|lo [64 x 128] as two K-blocks
// last block (block 4):      W1a hi|lo [128 x 64], W2 hi|lo [128 x 128] as two K-blocks
// then fp32 side arrays (read from global memory when an episode's biases are set up)
struct TowerBlobMid {
    static constexpr int ws_hi = 0, w1_hi = 8192, ws2_hi = 24576, ws_lo = 32768, w1_lo = 40960, ws2_lo = 57344, w2_hi = 65536,
                         w2_lo = 81920;
    static constexpr int smem_bytes = 98304;
    static constexpr int w1bT = smem_bytes;                 // [64][128] f32: W1[:, 64+k] transposed
    static constexpr int b1 = w1bT + 64 * 128 * 4;          // [128]
    static constexpr int wsbT = b1 + 128 * 4;               // [64][64]: Ws[:, 64+k] transposed
    static constexpr int b2 = wsbT + 64 * 64 * 4;           // [64]: b2 + bs
    static constexpr int total = b2 + 64 * 4;
};
struct TowerBlobLast {
    static constexpr int w1_hi = 0, w1_lo = 16384, w2_hi = 32768, w2_lo = 65536;
    static constexpr int smem_bytes = 98304;
    static constexpr int w1bT = smem_bytes;
    static constexpr int b1 = w1bT + 64 * 128 * 4;
    static constexpr int b2 = b1 + 128 * 4;                 // [128]
    static constexpr int total = b2 + 128 * 4;
};
// block 1: W1 [5][5] and b1 [5] in fp32 (the five hidden channels are computed on the fp32 pipes), then the B operand of
// its one GEMM, hi | lo: [64 rows][K = 16] in rows of 128 bytes (the swizzled K-major layout of the other blocks; only
// the first two 16-byte chunks of a row are used).  K slots: 0-4 the point's inputs x (weights Ws), 5-9 the hidden
// channels h (weights W2), 10 the constant 1 (weight b2 + bs), 11-15 zero.
struct TowerBlobFirst {
    static constexpr int w1 = 0;                            // [5][5] f32
    static constexpr int b1 = 25 * 4;                       // [5] f32
    static constexpr int w_hi = 128, w_lo = 128 + 8192;     // [64][64] 16-bit each
    static constexpr int total = 128 + 2 * 8192;
};

// ---- order-preserving float <-> uint keys for atomicMax (0 = "no value yet") ---------------------------------------
__device__ __forceinline__ unsigned f2key(float v) {
    // in PTX on .b32 registers: written in C the compiler turns `bits | 0x80000000` into FADD(-|v|, -0), which
    // canonicalises a NaN and drops the bit (measured: benchmarks/debug/nanmax_probe.cu) - and NaN must stay the
    // LARGEST key so that it wins every atomicMax
    unsigned k;
    asm("{\n"
        ".reg .b32 t, m;\n"
        "mov.b32 t, %1;\n"
        "shr.s32 m, t, 31;\n"
        "or.b32 m, m, 0x80000000;\n"
        "xor.b32 %0, t, m;\n"
        "}\n"
        : "=r"(k)
        : "f"(v));
    return k;
}
__device__ __forceinline__ float key2f(unsigned k) {
    return __uint_as_float((k & 0x80000000u) ? (k ^ 0x80000000u) : ~k);
}
// max that PROPAGATES NaN (fmaxf drops it): an activation beyond the fp16 range turns into inf - inf = NaN in the
// split products and must reach the output instead of being silently dropped by a max
__device__ __forceinline__ float max_nan(float a, float b) {
    float r;
    asm("max.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ float lrelu(float v) { return max_nan(v, __fmul_rn(kTowerSlope, v)); }
// packed fp32 pairs (Blackwell FADD2 / FMUL2 / FFMA2): two IEEE round-to-nearest operations per instruction, the
// same results as the scalar ones
__device__ __forceinline__ float2 add2(float2 a, float2 b) {
    if (!kTowerPacked) return make_float2(__fadd_rn(a.x, b.x), __fadd_rn(a.y, b.y));
    float2 r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(reinterpret_cast<unsigned long long &>(r)) : "l"(reinterpret_cast<unsigned long long &>(a)), "l"(reinterpret_cast<unsigned long long &>(b)));
    return r;
}
__device__ __forceinline__ float2 sub2(float2 a, float2 b) {
    if (!kTowerPacked) return make_float2(__fsub_rn(a.x, b.x), __fsub_rn(a.y, b.y));
    float2 r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(reinterpret_cast<unsigned long long &>(r)) : "l"(reinterpret_cast<unsigned long long &>(a)), "l"(reinterpret_cast<unsigned long long &>(b)));
    return r;
}
__device__ __forceinline__ float2 mul2(float2 a, float2 b) {
    if (!kTowerPacked) return make_float2(__fmul_rn(a.x, b.x), __fmul_rn(a.y, b.y));
    float2 r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(reinterpret_cast<unsigned long long &>(r)) : "l"(reinterpret_cast<unsigned long long &>(a)), "l"(reinterpret_cast<unsigned long long &>(b)));
    return r;
}
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
    if (!kTowerPacked) return make_float2(__fmaf_rn(a.x, b.x, c.x), __fmaf_rn(a.y, b.y, c.y));
    float2 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(reinterpret_cast<unsigned long long &>(r)) : "l"(reinterpret_cast<unsigned long long &>(a)), "l"(reinterpret_cast<unsigned long long &>(b)), "l"(reinterpret_cast<unsigned long long &>(c)));
    return r;
}
__device__ __forceinline__ float2 lrelu2(float2 v) {
    const float2 t = mul2(v, make_float2(kTowerSlope, kTowerSlope));
    return make_float2(max_nan(v.x, t.x), max_nan(v.y, t.y));
}

// two neighbouring channels -> one packed 16-bit pair (even channel in the low half), and back to fp32
__device__ __forceinline__ unsigned pack2(float even, float odd) {
    unsigned r;
    if (kTowerF16)
        asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(odd), "f"(even));
    else
        asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(odd), "f"(even));
    return r;
}
__device__ __forceinline__ float2 unpack2(unsigned p) {
    if (kTowerF16) return __half22float2(*reinterpret_cast<const __half2 *>(&p));
    return make_float2(__uint_as_float(p << 16), __uint_as_float(p & 0xffff0000u));
}
// ---- tcgen05 / TMEM wrappers (inline PTX; SASS: UTCHMMA, LDTM/STTM, UTCBAR) -----------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t *smem_slot, uint32_t cols) {   // one full warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_slot)), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {      // the same warp
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// all tcgen05.mma issued so far have completed -> one arrival on `bar` (whole converged warp, elected lane - as the MMAs)
__device__ __forceinline__ void tc_commit(uint64_t *bar) {
    asm volatile(
        "{\n"
        ".reg .pred e;\n"
        "elect.sync _|e, 0xffffffff;\n"
        "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n"
        "}\n" ::"r"(smem_u32(bar))
        : "memory");
}
// non-blocking probe: has the phase with this parity completed?
__device__ __forceinline__ bool mbar_test(uint64_t *bar, unsigned parity) {
    unsigned ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// The MMA wrappers are called by a whole CONVERGED warp and elect the issuing lane themselves (elect.sync): issued
// from inside an `if (lane == 0)` region the compiler cannot tell that a single thread is active and wraps every
// UTCHMMA in an elect / branch loop - ~95 cycles of issue per MMA, which made the issuing thread the tower's
// bottleneck (48 MMAs x 95 = the 4550 cycles per tile ncu measured).
// D[tmem] (+)= A[smem desc] * B[smem desc]            (kind::f16: 16-bit operands, fp32 accumulate)
__device__ __forceinline__ void mma_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, bool accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p, e;\n"
        "elect.sync _|e, 0xffffffff;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"((uint32_t)accumulate)
        : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem desc]
__device__ __forceinline__ void mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, bool accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p, e;\n"
        "elect.sync _|e, 0xffffffff;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
        "}\n" ::"r"(d_tmem),
        "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"((uint32_t)accumulate)
        : "memory");
}
// shared-memory matrix descriptor: K-major, 128-byte rows, 128B swizzle, 8-row groups 1024 bytes apart
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
    return (uint64_t)((smem_addr & 0x3ffffu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
           ((uint64_t)2 << 61);
}
// instruction descriptor: 16-bit x 16-bit -> fp32, both operands K-major, M = 128, N = n; format 0 = f16, 1 = bf16
__host__ __device__ constexpr uint32_t umma_idesc_16(int n, int a_bf16, int b_bf16) {
    return (1u << 4) | ((uint32_t)a_bf16 << 7) | ((uint32_t)b_bf16 << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
// pass p of a split product: which piece of A and of B it multiplies (0 = hi, 1 = lo), and the descriptor for it
__host__ __device__ constexpr int tower_pass_a(int p) { return p >= 2 ? 1 : 0; }   // hh, hl, lh, ll
__host__ __device__ constexpr int tower_pass_b(int p) { return p & 1; }
__host__ __device__ constexpr uint32_t tower_idesc(int n, int p) {
    return umma_idesc_16(n, kTowerF16 ? 0 : 1, kTowerF16 ? 0 : 1);
}
// 32 lanes x 32 columns of fp32: lane = this thread's TMEM lane, r[i] = column i
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,"
        "%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                   "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr)
                 : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
        "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,"
        "%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
        "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]),
        "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]),
        "r"(r[31])
        : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int threads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory"); }

// 2-D tiled TMA through a 3-D tensor map [B][N][64] bf16 (boxes of [1][128][64], 128B swizzle; rows beyond N are
// zero-filled on load and clipped on store)
__device__ __forceinline__ void tma_store_3d_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_3d_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// per-channel max over the 32 points of a warp: lane l ends up with max_p v_p[l] (31 exchanges instead of 32 x 5)
__device__ __forceinline__ float warp_transpose_max(float (&v)[32], int lane) {
#pragma unroll
    for (int half = 16; half >= 1; half >>= 1) {
        const bool upper = (lane & half) != 0;
#pragma unroll
        for (int i = 0; i < half; ++i) {
            const float send = upper ? v[i] : v[i + half];
            const float keep = upper ? v[i + half] : v[i];
            v[i] = max_nan(keep, __shfl_xor_sync(kFull, send, half));
        }
    }
    return v[0];
}

// ---------------------------------------------------------------------------------------------------------------
// k_tower_pack: folded fp32 weights -> the blobs above (once per set of weights).
//   kind 0 = first block : W1 [5][5], b1 [5], W2 [64][5], b2 [64], Ws [64][5], bs [64]
//   kind 1 = mid block   : W1 [128][128], b1 [128], W2 [64][128], b2 [64], Ws [64][128], bs [64]
//   kind 2 = last block  : W1 [128][128], b1 [128], W2 [128][128], b2 [128]
// A bf16 image of W [rows][K-block of 64]: element (n, k) at n*128 + (((k>>3) ^ (n&7)) << 4) + (k&7)*2 bytes.
__device__ __forceinline__ void pack_split(unsigned char *img_hi, unsigned char *img_lo, int n, int k, float w) {
    const int off = n * 128 + ((((k & 63) >> 3) ^ (n & 7)) << 4) + (k & 7) * 2;
    const unsigned h = pack2(w, 0.f);
    const unsigned l = pack2(__fsub_rn(w, unpack2(h).x), 0.f);
    *reinterpret_cast<unsigned short *>(img_hi + off) = (unsigned short)(h & 0xffffu);
    *reinterpret_cast<unsigned short *>(img_lo + off) = (unsigned short)(l & 0xffffu);
}
__global__ void k_tower_pack(int kind, const float *__restrict__ W1, const float *__restrict__ b1, const float *__restrict__ W2,
                             const float *__restrict__ b2, const float *__restrict__ Ws, const float *__restrict__ bs,
                             unsigned char *__restrict__ blob) {
    const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
    if (kind == 0) {
        for (int i = tid; i < 64 * 16; i += nth) {
            const int c = i >> 4, k = i & 15;
            const float w = k < 5 ? Ws[c * 5 + k] : k < 10 ? W2[c * 5 + k - 5] : k == 10 ? __fadd_rn(b2[c], bs[c]) : 0.f;
            pack_split(blob + TowerBlobFirst::w_hi, blob + TowerBlobFirst::w_lo, c, k, w);
        }
        float *w1 = reinterpret_cast<float *>(blob + TowerBlobFirst::w1);
        for (int i = tid; i < 25; i += nth) w1[i] = W1[i];
        float *bb = reinterpret_cast<float *>(blob + TowerBlobFirst::b1);
        for (int i = tid; i < 5; i += nth) bb[i] = b1[i];
        return;
    }
    const bool last = kind == 2;
    const int out2 = last ? 128 : 64;
    // conv1, feature half: W1[:, :64] -> B operand [128 rows][64 K]
    for (int i = tid; i < 128 * 64; i += nth) {
        const int n = i >> 6, k = i & 63;
        pack_split(blob + (last ? TowerBlobLast::w1_hi : TowerBlobMid::w1_hi), blob + (last ? TowerBlobLast::w1_lo : TowerBlobMid::w1_lo), n, k,
                   W1[n * 128 + k]);
    }
    // conv2: W2 [out2][128] -> two K-blocks of [out2 rows][64 K]
    const int w2_hi = last ? TowerBlobLast::w2_hi : TowerBlobMid::w2_hi, w2_lo = last ? TowerBlobLast::w2_lo : TowerBlobMid::w2_lo;
    for (int i = tid; i < out2 * 128; i += nth) {
        const int n = i >> 7, k = i & 127, kb = k >> 6;
        pack_split(blob + w2_hi + kb * out2 * 128, blob + w2_lo + kb * out2 * 128, n, k, W2[n * 128 + k]);
    }
    float *w1bT = reinterpret_cast<float *>(blob + (last ? TowerBlobLast::w1bT : TowerBlobMid::w1bT));
    for (int i = tid; i < 64 * 128; i += nth) {
        const int k = i >> 7, n = i & 127;
        w1bT[i] = W1[n * 128 + 64 + k];
    }
    float *pb1 = reinterpret_cast<float *>(blob + (last ? TowerBlobLast::b1 : TowerBlobMid::b1));
    for (int i = tid; i < 128; i += nth) pb1[i] = b1[i];
    if (last) {
        float *pb2 = reinterpret_cast<float *>(blob + TowerBlobLast::b2);
        for (int i = tid; i < 128; i += nth) pb2[i] = b2[i];
    } else {
        for (int i = tid; i < 64 * 64; i += nth) {
            const int n = i >> 6, k = i & 63;
            pack_split(blob + TowerBlobMid::ws_hi, blob + TowerBlobMid::ws_lo, n, k, Ws[n * 128 + k]);
            pack_split(blob + TowerBlobMid::ws2_hi, blob + TowerBlobMid::ws2_lo, n, k, Ws[n * 128 + k]);
        }
        float *wsbT = reinterpret_cast<float *>(blob + TowerBlobMid::wsbT);
        for (int i = tid; i < 64 * 64; i += nth) {
            const int k = i >> 6, n = i & 63;
            wsbT[i] = Ws[n * 128 + 64 + k];
        }
        float *pb2 = reinterpret_cast<float *>(blob + TowerBlobMid::b2);
        for (int i = tid; i < 64; i += nth) pb2[i] = __fadd_rn(b2[i], bs[i]);
    }
}

// contiguous share of `total` tiles for CTA `cta` of `nctas`
__device__ __forceinline__ void tower_tile_range(int total, int cta, int nctas, int &t0, int &t1) {
    t0 = (int)((long long)total * cta / nctas);
    t1 = (int)((long long)total * (cta + 1) / nctas);
}

// stage one output tile: thread = point `p` of the tile, v[32] = channels 32*j .. 32*j+31 -> bf16 pieces into the
// 128B-swizzled planes (row p = 128 bytes; 16-byte chunk q of the row lives at position q ^ (p & 7))
template <int kPlanes>
__device__ __forceinline__ void tower_stage_chunk(unsigned char *planes, int plane_bytes, int p, int j, const float (&v)[32]) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        uint4 hi, lo, lo2;
        unsigned *ph = &hi.x, *pl = &lo.x, *pl2 = &lo2.x;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const float2 ab = make_float2(v[q * 8 + 2 * e], v[q * 8 + 2 * e + 1]);
            const unsigned h = pack2(ab.x, ab.y);
            const float2 res = sub2(ab, unpack2(h));
            const unsigned l = pack2(res.x, res.y);
            ph[e] = h;
            pl[e] = l;
            if (kPlanes == 3) {
                const float2 res2 = sub2(res, unpack2(l));
                pl2[e] = pack2(res2.x, res2.y);
            }
        }
        const int off = p * 128 + (((4 * j + q) ^ (p & 7)) << 4);
        *reinterpret_cast<uint4 *>(planes + off) = hi;
        *reinterpret_cast<uint4 *>(planes + plane_bytes + off) = lo;
        if (kPlanes == 3) *reinterpret_cast<uint4 *>(planes + 2 * plane_bytes + off) = lo2;
    }
}

// the same for 16 channels [16 * cq, 16 * cq + 16) of the row: two 16-byte chunks per plane
__device__ __forceinline__ void tower_stage_cols16(unsigned char *planes, int plane_bytes, int p, int cq, const float (&v)[16]) {
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        uint4 hi, lo;
        unsigned *ph = &hi.x, *pl = &lo.x;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const float2 ab = make_float2(v[q * 8 + 2 * e], v[q * 8 + 2 * e + 1]);
            const unsigned h = pack2(ab.x, ab.y);
            const float2 res = sub2(ab, unpack2(h));
            ph[e] = h;
            pl[e] = pack2(res.x, res.y);
        }
        const int off = p * 128 + (((2 * cq + q) ^ (p & 7)) << 4);
        *reinterpret_cast<uint4 *>(planes + off) = hi;
        *reinterpret_cast<uint4 *>(planes + plane_bytes + off) = lo;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// k_tower_first: block 1.  obs3d [B][5][N] -> feat1 planes (hi, lo) + max keys.
//   h = lrelu(W1 x + b1)             five channels, on the fp32 pipes (25 FMAs per point)
//   out = lrelu(W2 h + Ws x + b)     64 channels: ONE K = 16 GEMM per 128-point tile on the tensor core - the A row of a
//                                    point is (x, h, 1, 0...), the B operand (Ws | W2 | b2 + bs | 0), three fp16 passes
//                                    (the constant 1 has no low piece: the bias arrives exactly split)
// Until this version the 64 channels were 320 packed FFMA2 per point: 99 us for the KITTI batch, as much as a whole
// 128-to-64-channel block on the tensor core.
// Warps (480 threads, one CTA per SM, a contiguous range of tiles each):
//   0      the weights (one bulk copy)             1      TMEM allocation (2 x 64 columns) + MMA issue (converged, elected)
//   2-5    builders: thread = point; x (the next tile's already on its way), h, the A row's two 16-byte chunks per
//          piece into the stage (two stages)
//   6-13   epilogue: warp = (TMEM lane quarter) x (32 of the 64 columns, 16 at a time): lrelu, fp16 hi|lo pieces into
//          the output staging (two buffers), running maxima per point lane in registers
//   14     TMA stores of the staged tiles
// 86 us for 32 x 40960 points = 4.2 TB/s, 64 % of the HBM peak (276 bytes per point).  Measured and dropped: a second
// epilogue group on alternate tiles (86 us) and a second builder group as well (864 threads, spills: 102 us).
constexpr int kFirstThreads = 480;
struct FirstCfg {
    static constexpr int off_w = 0;                               // B operand: hi 8 KB | lo 8 KB
    static constexpr int off_a = 16384;                           // 2 stages x (hi 16 KB | lo 16 KB): [128 rows][128 B], first 32 B used
    static constexpr int off_out = off_a + 2 * 32768;             // 2 x (hi 16 KB | lo 16 KB): the output planes of a tile
    static constexpr int off_bars = off_out + 2 * 32768;
    static constexpr int kNumBars = 1 + 6 * 2;
    static constexpr int off_tmem_slot = off_bars + kNumBars * 8;
    static constexpr int smem_bytes = off_tmem_slot + 16;
};

__global__ void __launch_bounds__(kFirstThreads, 1)
    k_tower_first(const float *__restrict__ obs3d, const unsigned char *__restrict__ blob, int B, int N, int tiles_per_ep,
                  const __grid_constant__ CUtensorMap map_hi, const __grid_constant__ CUtensorMap map_lo, unsigned *__restrict__ max_keys) {
    extern __shared__ __align__(1024) unsigned char tower_smem[];
    unsigned char *const smem = tower_smem;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + FirstCfg::off_bars);
    uint64_t *w_full = bars, *a_full = bars + 1, *a_empty = a_full + 2, *d_full = a_empty + 2, *d_empty = d_full + 2, *o_full = d_empty + 2,
             *o_free = o_full + 2;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + FirstCfg::off_tmem_slot);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    int t0, t1;
    tower_tile_range(B * tiles_per_ep, blockIdx.x, gridDim.x, t0, t1);
    const int ntiles = t1 - t0;
    if (tid == 0) {
        mbar_init(w_full, 1);
        for (int s = 0; s < 2; ++s) {
            mbar_init(a_full + s, 128);
            mbar_init(a_empty + s, 1);
            mbar_init(d_full + s, 1);
            mbar_init(d_empty + s, 256);
            mbar_init(o_full + s, 256);
            mbar_init(o_free + s, 1);
        }
        tma_prefetch_map(&map_hi);
        tma_prefetch_map(&map_lo);
    }
    if (warp == 1) tmem_alloc(tmem_slot, 128);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    // walks the CTA's tiles without a division per tile
    int e = t0 / tiles_per_ep, n0 = (t0 - e * tiles_per_ep) * kTowerTile - kTowerTile;
    auto next_tile = [&]() {
        n0 += kTowerTile;
        if (n0 >= tiles_per_ep * kTowerTile) {
            n0 = 0;
            ++e;
        }
    };

    if (warp == 0) {
        if (lane == 0) {   // the weights do not depend on the previous kernel
            mbar_arrive_expect_tx(w_full, 16384);
            bulk_g2s(smem + FirstCfg::off_w, blob + TowerBlobFirst::w_hi, 16384, w_full);
        }
    } else if (warp == 1) {
        // ============================== MMA issuer (whole warp, converged; see mma_ss) ==============================
        const uint32_t sbase = smem_u32(smem);
        mbar_wait(w_full, 0);
        __syncwarp();
        for (int i = 0; i < ntiles; ++i) {
            const int s = i & 1;
            mbar_wait(a_full + s, (i >> 1) & 1);
            if (i >= 2) mbar_wait(d_empty + s, ((i >> 1) - 1) & 1);
            __syncwarp();
            tc_fence_after();
            const uint32_t a = sbase + FirstCfg::off_a + s * 32768, d = tmem_base + s * 64;
#pragma unroll
            for (int pass = 0; pass < kTowerPasses; ++pass)
                mma_ss(d, umma_desc_sw128(a + (tower_pass_a(pass) ? 16384 : 0)),
                       umma_desc_sw128(sbase + FirstCfg::off_w + (tower_pass_b(pass) ? 8192 : 0)), tower_idesc(64, pass), pass != 0);
            tc_commit(d_full + s);
            tc_commit(a_empty + s);
        }
    } else if (warp < 6) {
        // ============================== builders: thread = point ==============================
        const int p = tid - 64;
        float w1[25], b1[5];
#pragma unroll
        for (int i = 0; i < 25; ++i) w1[i] = __ldg(reinterpret_cast<const float *>(blob + TowerBlobFirst::w1) + i);
#pragma unroll
        for (int i = 0; i < 5; ++i) b1[i] = __ldg(reinterpret_cast<const float *>(blob + TowerBlobFirst::b1) + i);
        pdl_wait();   // obs3d comes from the kernel before
        // the next tile's inputs are loaded while this tile is built (a builder that waited for its own loads set the
        // kernel's pace: one DRAM round trip per tile)
        auto load_x = [&](float (&xx)[5]) {
            const int n = n0 + p;
#pragma unroll
            for (int c = 0; c < 5; ++c) xx[c] = n < N ? __ldg(obs3d + ((size_t)e * 5 + c) * N + n) : 0.f;
        };
        float xn[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
        if (ntiles > 0) {
            next_tile();
            load_x(xn);
        }
        for (int i = 0; i < ntiles; ++i) {
            const int s = i & 1;
            float x[5], h[5];
#pragma unroll
            for (int c = 0; c < 5; ++c) x[c] = xn[c];
            if (i + 1 < ntiles) {
                next_tile();
                load_x(xn);
            }
#pragma unroll
            for (int o = 0; o < 5; ++o) {
                float a = b1[o];
#pragma unroll
                for (int c = 0; c < 5; ++c) a = __fmaf_rn(w1[o * 5 + c], x[c], a);
                h[o] = lrelu(a);
            }
            // the A row: K slots (x0..x4, h0..h4, 1, 0 x 5) as fp16 pairs, hi and lo pieces
            const float in[16] = {x[0], x[1], x[2], x[3], x[4], h[0], h[1], h[2], h[3], h[4], 1.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            uint4 hi[2], lo[2];
            unsigned *ph = &hi[0].x, *pl = &lo[0].x;
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const float2 ab = make_float2(in[2 * q], in[2 * q + 1]);
                const unsigned hh = pack2(ab.x, ab.y);
                const float2 res = sub2(ab, unpack2(hh));
                ph[q] = hh;
                pl[q] = pack2(res.x, res.y);
            }
            if (i >= 2) mbar_wait(a_empty + s, ((i >> 1) - 1) & 1);   // the MMAs that read this stage are done
            unsigned char *st = smem + FirstCfg::off_a + s * 32768 + p * 128;
            *reinterpret_cast<uint4 *>(st + ((0 ^ (p & 7)) << 4)) = hi[0];
            *reinterpret_cast<uint4 *>(st + ((1 ^ (p & 7)) << 4)) = hi[1];
            *reinterpret_cast<uint4 *>(st + 16384 + ((0 ^ (p & 7)) << 4)) = lo[0];
            *reinterpret_cast<uint4 *>(st + 16384 + ((1 ^ (p & 7)) << 4)) = lo[1];
            fence_async_proxy();   // generic-proxy writes, read by the tensor core through the async proxy
            mbar_arrive(a_full + s);
        }
    } else if (warp < 14) {
        // ============================== epilogue ==============================
        const int ew = warp - 6;
        const int half = ew >> 2;                       // which 32 of the 64 columns
        const int wq = warp & 3;                        // TMEM lane quarter this warp may touch
        const int p = wq * 32 + lane;                   // point of the tile = TMEM lane
        const uint32_t d0 = tmem_base + ((uint32_t)(wq * 32) << 16) + 32 * half;
        float rmx[32];
#pragma unroll
        for (int q = 0; q < 32; ++q) rmx[q] = -INFINITY;
        int cur_ep = -1;
        auto flush = [&]() {
            if (cur_ep >= 0) {
                float t[32];
#pragma unroll
                for (int q = 0; q < 32; ++q) t[q] = rmx[q];
                const float m = warp_transpose_max(t, lane);
                atomicMax(max_keys + cur_ep * 64 + 32 * half + lane, f2key(m));
            }
#pragma unroll
            for (int q = 0; q < 32; ++q) rmx[q] = -INFINITY;
        };
        pdl_wait();   // the keys were cleared before this kernel
        for (int i = 0; i < ntiles; ++i) {
            next_tile();
            if (e != cur_ep) {
                flush();
                cur_ep = e;
            }
            const int grp = i & 1;                       // accumulator and staging buffer of this tile
            const uint32_t d = d0 + grp * 64;
            unsigned char *ostage = smem + FirstCfg::off_out + grp * 32768;
            const uint32_t par = (i >> 1) & 1;
            const bool valid = n0 + p < N;
            mbar_wait(d_full + grp, par);
            tc_fence_after();
            uint32_t r0[16], r1[16];
            tmem_ld16(d, r0);
            tmem_ld16(d + 16, r1);
            tc_wait_ld();
            tc_fence_before();
            mbar_arrive(d_empty + grp);                  // the accumulator may be overwritten
            if (i >= 2) mbar_wait(o_free + grp, par ^ 1);   // the store of tile i - 2 has read the staging buffer
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
                float v[16];
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const uint32_t a0 = hh ? r1[2 * q] : r0[2 * q], a1 = hh ? r1[2 * q + 1] : r0[2 * q + 1];
                    const float2 a = lrelu2(make_float2(__uint_as_float(a0), __uint_as_float(a1)));
                    v[2 * q] = a.x;
                    v[2 * q + 1] = a.y;
                }
                tower_stage_cols16(ostage, 16384, p, 2 * half + hh, v);
                if (valid) {
#pragma unroll
                    for (int q = 0; q < 16; ++q) rmx[16 * hh + q] = max_nan(rmx[16 * hh + q], v[q]);
                }
            }
            fence_async_proxy();
            mbar_arrive(o_full + grp);
        }
        flush();
    } else {
        // ============================== TMA-store warp ==============================
        if (lane == 0) {
            for (int i = 0; i < ntiles; ++i) {
                next_tile();
                const int s = i & 1;
                unsigned char *ostage = smem + FirstCfg::off_out + s * 32768;
                mbar_wait(o_full + s, (i >> 1) & 1);
                tma_store_3d(&map_hi, 0, n0, e, ostage);
                tma_store_3d(&map_lo, 0, n0, e, ostage + 16384);
                bulk_commit();
                tma_store_3d_wait_read();
                mbar_arrive(o_free + s);
            }
            tma_store_3d_wait_all();                     // the next block reads these planes
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, 128);
    pdl_launch_dependents();
}

// ---------------------------------------------------------------------------------------------------------------
// k_tower_mma<kLast>: blocks 2-4.  Features travel as two fp16 planes (hi, lo): 22 significant bits.
//   kLast = false: in planes (hi, lo) -> out planes (hi, lo); 64 max keys / episode
//   kLast = true : in planes (hi, lo); no feature output; 128 max keys / episode
// Shared memory: weights | input stages (TMA loads) | per-group output staging (mid blocks; TMA stores) | biases.
// An input stage of a mid block is handed back by the tensor core itself (tcgen05.commit after the first GEMM), and
// the outputs are staged in buffers of their own: while a stage doubled as the output staging area it stayed
// occupied for a tile's whole life (~8 us), and three stages capped the kernel at ~2.7 us per tile.
template <bool kLast>
struct TowerCfg {
    static constexpr int kInPlanes = 2;
    static constexpr int kStageBytes = 32768;                       // two planes of 16 KB
    static constexpr int kStages = kLast ? 4 : 2;
    static constexpr int kOutBytes = kLast ? 0 : 2 * 32768;         // one staging buffer per epilogue group
    static constexpr int kWeightBytes = kLast ? TowerBlobLast::smem_bytes : TowerBlobMid::smem_bytes;
    static constexpr int kN2 = kLast ? 128 : 64;                    // channels of the second GEMM
    // TMEM columns per epilogue group: last block [D1/H 128][D2 128]; mid blocks [D2a 64][D1/H 128][D2b 64] - the
    // tiles of a group alternate between D2a and D2b, so that the first GEMM of the group's NEXT tile (which also
    // writes that tile's D2) runs while the epilogue still reads this tile's D2
    static constexpr int kBufCols = 256;
    static constexpr int kD1 = kLast ? 0 : 64;                      // column of D1 inside the group's block
    static constexpr int kTmemCols = 512;
    static constexpr int off_stage = kWeightBytes;
    static constexpr int off_out = off_stage + kStages * kStageBytes;
    static constexpr int off_bias1 = off_out + kOutBytes;                  // [2][128] f32 (by episode parity)
    static constexpr int off_bias2 = off_bias1 + 1024;                     // [2][128] f32
    static constexpr int off_maxprev = off_bias2 + 1024;                   // [64] f32
    static constexpr int off_bars = off_maxprev + 256;                     // mbarriers
    static constexpr int kNumBars = 1 + 2 * kStages + 6 + 4;
    static constexpr int off_tmem_slot = off_bars + kNumBars * 8;
    static constexpr int smem_bytes = off_tmem_slot + 16;
};
static_assert(TowerCfg<false>::smem_bytes <= 232448 && TowerCfg<true>::smem_bytes <= 232448, "tower kernels: shared memory over the 227 KB limit");
static_assert(TowerCfg<false>::off_stage % 1024 == 0 && TowerCfg<true>::off_stage % 1024 == 0, "TMA stages need 1024-byte alignment (128B swizzle)");

template <bool kLast>
__global__ void __launch_bounds__(kTowerThreads, 1)
k_tower_mma(const unsigned char *__restrict__ blob, int B, int N, int tiles_per_ep, int box_rows, const __grid_constant__ CUtensorMap in_hi,
            const __grid_constant__ CUtensorMap in_lo, const __grid_constant__ CUtensorMap out_hi,
            const __grid_constant__ CUtensorMap out_lo, const unsigned *__restrict__ prev_keys, unsigned *__restrict__ max_keys) {
    using Cfg = TowerCfg<kLast>;
    using Blob = typename std::conditional<kLast, TowerBlobLast, TowerBlobMid>::type;
    extern __shared__ __align__(1024) unsigned char tower_smem[];
    unsigned char *const smem = tower_smem;
    float *bias1 = reinterpret_cast<float *>(smem + Cfg::off_bias1);
    float *bias2 = reinterpret_cast<float *>(smem + Cfg::off_bias2);
    float *maxprev = reinterpret_cast<float *>(smem + Cfg::off_maxprev);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + Cfg::off_bars);
    uint64_t *w_full = bars;
    uint64_t *x_full = bars + 1, *x_empty = x_full + Cfg::kStages;
    uint64_t *d1_full = x_empty + Cfg::kStages, *h_full = d1_full + 2, *d2_full = h_full + 2;
    uint64_t *o_full = d2_full + 2, *o_free = o_full + 2;      // output staging buffers (mid blocks)
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + Cfg::off_tmem_slot);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    int t0, t1;
    tower_tile_range(B * tiles_per_ep, blockIdx.x, gridDim.x, t0, t1);
    const int ntiles = t1 - t0;

    if (tid == 0) {
        mbar_init(w_full, 1);
        for (int s = 0; s < Cfg::kStages; ++s) {
            mbar_init(x_full + s, 1);
            mbar_init(x_empty + s, kLast ? 512 : 1);
        }
        for (int g = 0; g < 2; ++g) {
            mbar_init(d1_full + g, 1);
            mbar_init(h_full + g, 512);
            mbar_init(d2_full + g, 1);
            mbar_init(o_full + g, 512);
            mbar_init(o_free + g, 1);
        }
        tma_prefetch_map(&in_hi);
        tma_prefetch_map(&in_lo);
        if (!kLast) {
            tma_prefetch_map(&out_hi);
            tma_prefetch_map(&out_lo);
        }
    }
    if (warp == 1) tmem_alloc(tmem_slot, Cfg::kTmemCols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ============================== TMA producer ==============================
        if (lane == 0) {
            mbar_arrive_expect_tx(w_full, Cfg::kWeightBytes);          // the weights do not depend on the previous kernel
            bulk_g2s(smem, blob, Cfg::kWeightBytes, w_full);
            pdl_wait();                                                // the feature planes do
            for (int i = 0; i < ntiles; ++i) {
                const int s = i % Cfg::kStages, use = i / Cfg::kStages;
                if (use > 0) mbar_wait(x_empty + s, (use - 1) & 1);
                const int t = t0 + i, e = t / tiles_per_ep, n0 = (t - e * tiles_per_ep) * kTowerTile;
                unsigned char *st = smem + Cfg::off_stage + s * Cfg::kStageBytes;
                mbar_arrive_expect_tx(x_full + s, Cfg::kInPlanes * box_rows * 128);   // a box is min(N, 128) rows
                tma_load_3d(st, &in_hi, 0, n0, e, x_full + s);
                tma_load_3d(st + 16384, &in_lo, 0, n0, e, x_full + s);
            }
        }
    } else if (warp == 1) {
        // ============================== MMA issuer (whole warp, converged; see mma_ss) ==============================
        {
            const uint32_t sbase = smem_u32(smem);

            // second GEMM of tile i (its h is in TMEM): D2 (+)= H * W2^T, three passes over 8 K-chunks
            auto issue_c2 = [&](int i) {
                const int g = i & 1;
                tc_fence_after();
                const uint32_t d1 = tmem_base + g * Cfg::kBufCols + Cfg::kD1;
                const uint32_t d2 = kLast ? d1 + 128 : (((i >> 1) & 1) ? d1 + 128 : d1 - 64);
#pragma unroll
                for (int pass = 0; pass < kTowerPasses; ++pass) {
                    const int a_lo = tower_pass_a(pass) ? 16 : 0;               // the lo pairs sit 16 columns after the hi pairs
                    const uint32_t w = sbase + (tower_pass_b(pass) ? Blob::w2_lo : Blob::w2_hi);
                    constexpr int n2 = Cfg::kN2;
                    const uint32_t idesc = tower_idesc(n2, pass);
#pragma unroll
                    for (int kk = 0; kk < 8; ++kk) {
                        const uint32_t a = d1 + 32 * (kk >> 1) + 8 * (kk & 1) + a_lo;
                        const uint64_t b = umma_desc_sw128(w + (kk >> 2) * (Cfg::kN2 * 128) + (kk & 3) * 32);
                        mma_ts(d2, a, b, idesc, true);   // D2 already holds the shortcut product (mid) or shortcut + bias (last)
                    }
                }
                tc_commit(d2_full + g);
            };
            // first GEMM of tile i: D1 = X * W1a^T (N = 128); mid blocks also start D2 = X * Wsa^T (N = 64)
            auto issue_c1 = [&](int i) {
                const int s = i % Cfg::kStages, g = i & 1;
                tc_fence_after();
                const uint32_t xs = sbase + Cfg::off_stage + s * Cfg::kStageBytes;
                // mid blocks: Wsa and W1a are stacked into one 192-row B operand and D2's columns adjoin D1's, so one
                // MMA per K-chunk produces both X * W1a^T (128 columns) and the shortcut product X * Wsa^T (64 columns):
                // {D2a, D1} from rows {Wsa, W1a} for the group's even tiles, {D1, D2b} from rows {W1a, Wsa} for its odd ones
                constexpr int n1 = kLast ? 128 : 192;
                const bool second = ((i >> 1) & 1) != 0;
                const uint32_t dst = tmem_base + g * Cfg::kBufCols + (kLast || second ? Cfg::kD1 : 0);
#pragma unroll
                for (int pass = 0; pass < kTowerPasses; ++pass) {
                    const uint32_t xa = xs + (tower_pass_a(pass) ? 16384 : 0);
                    uint32_t w1 = sbase + (tower_pass_b(pass) ? Blob::w1_lo : Blob::w1_hi);
                    if (!kLast && !second) w1 -= 8192;                       // start at the leading copy of Wsa
                    const uint32_t idesc = tower_idesc(n1, pass);
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk)
                        mma_ss(dst, umma_desc_sw128(xa + kk * 32), umma_desc_sw128(w1 + kk * 32), idesc, (pass | kk) != 0);
                }
                tc_commit(d1_full + g);
                if (!kLast) tc_commit(x_empty + s);      // both products of x are done when this fires: refill the stage
            };
            mbar_wait(w_full, 0);
            __syncwarp();
            // Issue order: C1(0), C1(1), then for every tile i: C2(i), C1(i + 2).  The tensor pipe executes in issue
            // order, so C1(i + 2) - which overwrites the D1/H columns C2(i) reads - needs no barrier of its own; its D2
            // buffer (mid: the group's other one; last: written by the epilogue, not by C1) was released before the
            // group arrived on h_full(i).  The first GEMM of a group's next tile therefore runs while the group is still
            // in this tile's second epilogue: with C1(i + 2) held back until that epilogue had finished, each group
            // sat through C1 + E1 + C2 + E2 in series (3500 cycles per tile, ncu).
            auto wait_x = [&](int i) {
                mbar_wait(x_full + i % Cfg::kStages, (i / Cfg::kStages) & 1);
                __syncwarp();
            };
            for (int i = 0; i < 2 && i < ntiles; ++i) {
                wait_x(i);
                issue_c1(i);
            }
            for (int i = 0; i < ntiles; ++i) {
                mbar_wait(h_full + (i & 1), (i >> 1) & 1);
                __syncwarp();
                issue_c2(i);
                if (i + 2 < ntiles) {
                    wait_x(i + 2);
                    issue_c1(i + 2);
                }
            }
        }
    } else if (warp == 18) {
        // ============================== TMA-store warp (mid blocks) ==============================
        // The epilogue warps hand a staged tile over through an mbarrier and go on; with a 512-thread bar.sync before
        // the store every warp waited for the slowest one, once per tile (13 % of the epilogue's time, ncu).
        if (!kLast && lane == 0) {
            int e = t0 / tiles_per_ep, n0 = (t0 - e * tiles_per_ep) * kTowerTile - kTowerTile;
            for (int i = 0; i < ntiles; ++i) {
                n0 += kTowerTile;
                if (n0 >= tiles_per_ep * kTowerTile) {
                    n0 = 0;
                    ++e;
                }
                const int b = i & 1;
                unsigned char *ostage = smem + Cfg::off_out + b * 32768;
                mbar_wait(o_full + b, (i >> 1) & 1);
                tma_store_3d(&out_hi, 0, n0, e, ostage);
                tma_store_3d(&out_lo, 0, n0, e, ostage + 16384);
                bulk_commit();
                tma_store_3d_wait_read();
                mbar_arrive(o_free + b);
            }
            tma_store_3d_wait_all();                     // the next block reads these planes
        }
    } else {
        // ============================== epilogue (16 warps, one software-pipelined group) ==============================
        // Warp = (TMEM lane quarter wq = warp % 4) x (column quarter cq): it owns 32 of D1's 128 columns and a quarter
        // of D2's.  Per tile i the warps run E1(i + 1) BEFORE E2(i): the second GEMM of tile i has the whole of
        // E1(i + 1) to complete, so the wait for it is hidden (two alternating groups each sat in that wait 26 % of
        // the time, ncu).  The running maxima stay per point lane in registers (16 | 32 per thread) and are exchanged
        // across the warp only when an episode ends.
        const int ew = warp - 2;                        // 0..15
        const int cq = ew >> 2;                         // column quarter
        const int wq = warp & 3;                        // TMEM lane quarter this warp may touch
        const int p = wq * 32 + lane;                   // point of the tile = TMEM lane
        const int etid = tid - 64;                      // 0..511
        const uint32_t lane_addr = (uint32_t)(wq * 32) << 16;
        constexpr int kC2 = Cfg::kN2 / 4;               // D2 columns per warp: 16 (mid) | 32 (last)
        float rmx[kC2];
#pragma unroll
        for (int q = 0; q < kC2; ++q) rmx[q] = -INFINITY;
        const int e_first = t0 / tiles_per_ep;
        pdl_wait();                                      // prev_keys are the previous kernel's output

        // per-episode biases, double-buffered by episode parity (E1 may be one episode ahead of E2)
        auto setup_episode = [&](int e) {
            const int buf = (e - e_first) & 1;
            float *b1 = bias1 + buf * 128, *b2 = bias2 + buf * 128;
            // bias1[c] = b1[c] + sum_k W1[c][64 + k] * max_prev[k]: 8 channels per warp, the 64 terms in four slices of 16
            // across the lanes; bias2[c] = (b2 + bs)[c] + sum_k Ws[c][64 + k] * max_prev[k] (mid): 4 channels per warp,
            // eight slices of 8.  Every weight load is issued up front and is in flight together with the previous
            // block's maxima (a 64-step loop per thread was a chain of L2 round trips: ~5 us per episode change, ncu).
            const int c1 = 8 * ew + (lane & 7), ks1 = lane >> 3;
            const int c2 = 4 * ew + (lane & 3), ks2 = lane >> 2;
            float wv1[16], wv2[kLast ? 1 : 8];
            {
                const float *w = reinterpret_cast<const float *>(blob + Blob::w1bT) + (ks1 * 16) * 128 + c1;
#pragma unroll
                for (int k = 0; k < 16; ++k) wv1[k] = __ldg(w + k * 128);
            }
            if (!kLast) {
                const float *w = reinterpret_cast<const float *>(blob + TowerBlobMid::wsbT) + (ks2 * 8) * 64 + c2;
#pragma unroll
                for (int k = 0; k < 8; ++k) wv2[k < (kLast ? 1 : 8) ? k : 0] = __ldg(w + k * 64);
            }
            const float g1 = __ldg(reinterpret_cast<const float *>(blob + Blob::b1) + c1);
            const float g2 = __ldg(reinterpret_cast<const float *>(blob + Blob::b2) + (kLast ? (etid & 127) : c2));
            const float mp = etid < 64 ? key2f(prev_keys[e * 64 + etid]) : 0.f;
            named_bar_sync(1, 512);                      // everybody is past the E2 of the tile before last: buf is free
            if (etid < 64) maxprev[etid] = mp;
            named_bar_sync(1, 512);
            {
                float a = 0.f;
#pragma unroll
                for (int k = 0; k < 16; ++k) a = __fmaf_rn(wv1[k], maxprev[ks1 * 16 + k], a);
                a = __fadd_rn(a, __shfl_xor_sync(kFull, a, 8));
                a = __fadd_rn(a, __shfl_xor_sync(kFull, a, 16));
                if (ks1 == 0) b1[c1] = __fadd_rn(a, g1);
            }
            if (!kLast) {
                float a = 0.f;
#pragma unroll
                for (int k = 0; k < 8; ++k) a = __fmaf_rn(wv2[k < (kLast ? 1 : 8) ? k : 0], maxprev[ks2 * 8 + k], a);
                a = __fadd_rn(a, __shfl_xor_sync(kFull, a, 4));
                a = __fadd_rn(a, __shfl_xor_sync(kFull, a, 8));
                a = __fadd_rn(a, __shfl_xor_sync(kFull, a, 16));
                if (ks2 == 0) b2[c2] = __fadd_rn(a, g2);
            } else if (etid < 128) {
                b2[etid] = etid >= 64 ? __fadd_rn(g2, maxprev[etid - 64]) : g2;
            }
            named_bar_sync(1, 512);
        };
        // an episode's maxima leave the registers: lane l of the warp ends up with channel l of the warp's columns
        auto flush = [&](int e) {
            float t[32];
#pragma unroll
            for (int q = 0; q < 32; ++q) t[q] = q < kC2 ? rmx[q < kC2 ? q : 0] : -INFINITY;
            const float m = warp_transpose_max(t, lane);
            if (lane < kC2) atomicMax(max_keys + e * Cfg::kN2 + kC2 * cq + lane, f2key(kLast ? lrelu(m) : m));
#pragma unroll
            for (int q = 0; q < kC2; ++q) rmx[q] = -INFINITY;
        };

        // ---- E1(i): D1 -> h = lrelu(D1 + bias1) -> fp16 hi|lo pairs back into the same TMEM columns ----
        int e1 = e_first, n1 = (t0 - e1 * tiles_per_ep) * kTowerTile - kTowerTile, ep1 = -1;   // walked without divisions
        auto step_e1 = [&](int i) {
            n1 += kTowerTile;
            if (n1 >= tiles_per_ep * kTowerTile) {
                n1 = 0;
                ++e1;
            }
            if (e1 != ep1) {
                setup_episode(e1);
                ep1 = e1;
            }
            const int g = i & 1;
            const uint32_t par = (i >> 1) & 1;
            const float *b1 = bias1 + ((e1 - e_first) & 1) * 128 + 32 * cq;
            const uint32_t d1 = tmem_base + g * Cfg::kBufCols + Cfg::kD1 + lane_addr + 32 * cq;
            mbar_wait(d1_full + g, par);
            tc_fence_after();
            {
                uint32_t r[32];
                tmem_ld32(d1, r);
                tc_wait_ld();
                uint32_t hi[16], lo[16];
#pragma unroll
                for (int q = 0; q < 16; ++q) {
                    const float2 bb = *reinterpret_cast<const float2 *>(b1 + 2 * q);
                    const float2 ab = lrelu2(add2(make_float2(__uint_as_float(r[2 * q]), __uint_as_float(r[2 * q + 1])), bb));
                    const unsigned h = pack2(ab.x, ab.y);
                    const float2 res = sub2(ab, unpack2(h));
                    hi[q] = h;
                    lo[q] = pack2(res.x, res.y);
                }
                tmem_st16(d1, hi);
                tmem_st16(d1 + 16, lo);
            }
            if (kLast) {
                const int s = i % Cfg::kStages;
                const unsigned char *stage = smem + Cfg::off_stage + s * Cfg::kStageBytes;
                const float *b2 = bias2 + ((e1 - e_first) & 1) * 128 + 32 * cq;
                mbar_wait(x_full + s, (i / Cfg::kStages) & 1);   // observe the TMA's writes ourselves before reading them
                // D2 starts as bias + identity shortcut: feat (hi + lo) for c < 64, max_prev for c >= 64
                uint32_t r[32];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float4 b0 = *reinterpret_cast<const float4 *>(b2 + 8 * q);
                    const float4 b4 = *reinterpret_cast<const float4 *>(b2 + 8 * q + 4);
                    float o[8] = {b0.x, b0.y, b0.z, b0.w, b4.x, b4.y, b4.z, b4.w};
                    if (cq < 2) {
                        const int off = p * 128 + (((4 * cq + q) ^ (p & 7)) << 4);
                        const uint4 xh = *reinterpret_cast<const uint4 *>(stage + off);
                        const uint4 xl = *reinterpret_cast<const uint4 *>(stage + 16384 + off);
                        const unsigned *ph = &xh.x, *pl = &xl.x;
#pragma unroll
                        for (int e2 = 0; e2 < 4; ++e2) {
                            const float2 x = add2(unpack2(ph[e2]), unpack2(pl[e2]));
                            o[2 * e2] = __fadd_rn(o[2 * e2], x.x);
                            o[2 * e2 + 1] = __fadd_rn(o[2 * e2 + 1], x.y);
                        }
                    }
#pragma unroll
                    for (int e2 = 0; e2 < 8; ++e2) r[8 * q + e2] = __float_as_uint(o[e2]);
                }
                tmem_st32(d1 + 128, r);
                tc_wait_st();
                tc_fence_before();
                mbar_arrive(h_full + g);
                mbar_arrive(x_empty + s);                // this thread's reads of the input stage are done
            } else {
                tc_wait_st();
                tc_fence_before();
                mbar_arrive(h_full + g);
            }
        };

        // ---- E2(i): D2 -> out = lrelu(D2 + bias2) -> running max (+ staged 16-bit planes -> TMA store) ----
        int e2 = e_first, n2 = n1, ep2 = -1;
        auto step_e2 = [&](int i) {
            n2 += kTowerTile;
            if (n2 >= tiles_per_ep * kTowerTile) {
                n2 = 0;
                ++e2;
            }
            if (e2 != ep2) {
                if (ep2 >= 0) flush(ep2);
                ep2 = e2;
            }
            const int g = i & 1;
            const uint32_t par = (i >> 1) & 1;
            const bool valid = n2 + p < N;
            const uint32_t dg = tmem_base + g * Cfg::kBufCols + Cfg::kD1 + lane_addr;
            const uint32_t d2 = (kLast ? dg + 128 : (par ? dg + 128 : dg - 64)) + kC2 * cq;   // mid: the D2 buffers alternate
            mbar_wait(d2_full + g, par);
            tc_fence_after();
            if (kLast) {
                // only the max over the points leaves block 4, and LeakyReLU is monotonic: max(lrelu(v)) = lrelu(max(v))
                // - the activation is applied once per channel when the maxima are flushed (the bias is in D2 already)
                uint32_t r[32];
                tmem_ld32(d2, r);
                tc_wait_ld();
                if (valid) {
#pragma unroll
                    for (int q = 0; q < kC2; ++q) rmx[q] = max_nan(rmx[q], __uint_as_float(r[q < 32 ? q : 0]));
                }
            } else {
                unsigned char *ostage = smem + Cfg::off_out + (i & 1) * 32768;
                if (i >= 2) mbar_wait(o_free + (i & 1), ((i >> 1) - 1) & 1);   // the store of tile i - 2 has read the buffer
                const float *b2 = bias2 + ((e2 - e_first) & 1) * 128 + 16 * cq;
                uint32_t r[16];
                tmem_ld16(d2, r);
                tc_wait_ld();
                float v[16];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float4 bb = *reinterpret_cast<const float4 *>(b2 + 4 * q);
                    const float2 a = lrelu2(add2(make_float2(__uint_as_float(r[4 * q]), __uint_as_float(r[4 * q + 1])), make_float2(bb.x, bb.y)));
                    const float2 b = lrelu2(add2(make_float2(__uint_as_float(r[4 * q + 2]), __uint_as_float(r[4 * q + 3])), make_float2(bb.z, bb.w)));
                    v[4 * q] = a.x;
                    v[4 * q + 1] = a.y;
                    v[4 * q + 2] = b.x;
                    v[4 * q + 3] = b.y;
                }
                tower_stage_cols16(ostage, 16384, p, cq, v);
                if (valid) {
#pragma unroll
                    for (int q = 0; q < kC2; ++q) rmx[q] = max_nan(rmx[q], v[q < 16 ? q : 0]);
                }
                fence_async_proxy();
                mbar_arrive(o_full + (i & 1));           // the store warp takes it from here
            }
        };

        step_e1(0);
        for (int i = 0; i < ntiles; ++i) {
            if (i + 1 < ntiles) step_e1(i + 1);
            step_e2(i);
        }
        flush(ep2);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, Cfg::kTmemCols);
    pdl_launch_dependents();
}

// keys -> fp32 embedding [B][128]
__global__ void k_tower_finish(const unsigned *__restrict__ keys, float *__restrict__ out, int n) {
    pdl_wait();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        const float v = key2f(keys[i]);
        out[i] = v;
        if (!isfinite(v)) atomicExch(&g_fault, 3);   // an activation left the fp16 range (or the input held inf/NaN)
    }
}

}  // namespace cmr
