// conv_tc2: persistent, warp-specialised, TMA-fed tcgen05 implicit GEMM (B2J_K_CONV_TC / B2J_K_GEMM_TC):
//
//   warp 0        TMA producer: one thread issues cp.async.bulk.tensor loads straight into SWIZZLE_128B shared
//                 memory -- the weight tile Wt[n0:n0+BLOCK_N, k:k+32] (2-D tiled map) and the activation tile:
//                   A_TILED   x viewed as a dense [M, K] matrix (1x1 stride-1 convs, dot_general)
//                   A_IM2COL  NHWC x through an im2col tensor map: 128 output pixels x 32 channels of filter
//                             tap (kh, kw); padding / stride / dilation / image borders handled by the TMA unit
//   warp 1        MMA issuer: tcgen05.mma.cta_group::1.kind::tf32, M=128, N=BLOCK_N, K=8, accumulating into one
//                 of TWO TMEM accumulators so that tile i+1 is computed while tile i is drained
//   warps 2..9    epilogue: tcgen05.ld -> smem transpose -> fused per-channel / residual / ReLU steps -> coalesced
//                 128-bit stores (epilogue_chunk in gemm_tc.cuh)
//   persistent    grid = #SMs; each CTA walks tiles t = blockIdx.x + i*gridDim.x, n-tiles of one m-tile adjacent
//                 so that the activation tile is re-read from L2, not HBM.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/b2jax.h"
#include "gemm_tc.cuh"

namespace b2j {

enum { A_TILED = 0, A_IM2COL = 1, A_ROWS = 2, A_ROWS_U8 = 3 };     // A_ROWS_U8: A_ROWS over a packed uint8 source

// A_ROWS: k x k convolutions over FEW input channels (the 3-channel ResNet stem) fed from the raw NHWC input rows.
// The im2col tensor map needs C % 32 == 0, so round 1 re-laid the image out first (space-to-depth fold: a 367 MB copy written
// and read back, K = 147 padded to 256).  Here one M tile is (up to) 128 consecutive output pixels of ONE output row; the TMA
// producer stages the KH input rows the tile needs in shared memory once per tile -- a few plain 3-D boxes of 256 elements x KH
// rows, each covering the windows of seg_px consecutive pixels, zero fill outside the image (one box per 1 KB row segment was
// tried first: 21 TMA instructions per tile made the producer the bottleneck, 0.86 ms for the stem) -- four gather warps (one per TMEM lane quarter, a thread per output pixel) assemble the K-major operand -- k = (kh, kw, c),
// a run of KW*C contiguous source elements per filter row, offsets from a table built at kernel start -- and write it straight
// into TMEM (tcgen05.st) for .ts MMAs: K is only padded to the next multiple of 32 (147 -> 160), the input crosses the
// L2 -> SM fabric ~KH / stride_h times instead of once per tap, and nothing is written back to HBM.  uint8 sources are widened
// through a 256-entry table holding the fused input chain (x / 255 ...).
struct Tc2Rows {
  uint32_t seg_w;       // TMA box width in source elements (256)
  uint32_t seg_px;      // output pixels served by one box: (seg_px - 1) * pix_step + kwc + align - 1 <= seg_w, seg_px * pix_step % align == 0
  uint32_t nseg;        // boxes (seg_w elements x KH rows, overlapping by kwc - pix_step elements) per 128-pixel tile
  uint32_t pix_step;    // stride_w * C: source elements between the windows of adjacent output pixels
  uint32_t kwc;         // KW * C: the contiguous run one filter row reads
  uint32_t k;           // KH * KW * C
  uint32_t tiles_w;     // 128-pixel tiles per output row
  uint32_t src_u8;      // source elements are bytes (packed uint8 image)
  uint32_t round_in;    // single-pass mode: round the operand to nearest TF32 (the tensor core would truncate)
  uint32_t pre_n, pre_op[2], pre_imm[2];   // fused input chain applied to the widened uint8 value (b2j_relayout_params)
  uint32_t align;       // elements per 16 bytes (4: f32, 16: uint8), a power of two
};
// first source element (innermost coordinate, may be negative) of 128-pixel tile `wseg` of an output row
__device__ __forceinline__ int rows_x0(const b2j_conv_tc_params& p, const Tc2Rows& rows, uint32_t wseg) {
  return ((int)(wseg * TC_BLOCK_M * p.stride_w) - p.pad_w) * (int)p.c;
}

// The single-thread roles (TMA producer, MMA issuer) are run by their WHOLE warp with one elected lane issuing, see the MMA role.
#ifndef B2J_WARP_MMA
#define B2J_WARP_MMA 1
#endif
// The TMA producer likewise, in the single-pass kernels only: measured on ResNet-50 b256 it gains ~1 % there (the stride-2 im2col
// layer 0.138 -> 0.129 ms) but makes the 64-wide 3xTF32 layers slower (0.326 -> 0.381 ms), where the producer warp shares its
// scheduler with a splitter warp.
#ifndef B2J_WARP_TMA
#define B2J_WARP_TMA (!X3)
#endif
#ifndef B2J_PDL_DEFAULT
// Programmatic dependent launch between consecutive conv_tc2 launches.  Round 1 measured -0.01 .. -0.03 ms per step and left it
// off; with the round-2 kernels it gives 6.82 -> 6.77 ms single pass and 14.72 -> 14.66 ms 3xTF32 (A/B on one box), the full GPU
// suite passes with it: on by default, B2J_PDL=0 switches it off.
#define B2J_PDL_DEFAULT 1
#endif

// CG = 1: one CTA per 128 x BLOCK_N tile.  CG = 2: a CTA pair (cluster of 2, tcgen05 cta_group::2) per 256 x BLOCK_N tile:
// each CTA stages its own 128 activation rows and HALF of the weight tile, the pair's tensor cores share the halves,
// so the shared-memory fill per FLOP drops (the L2 -> SM fabric, not the tensor pipe, bounds single-CTA TF32 tiles).
// RES2 (single pass, 256 x 128 pair tiles, BN + residual + ReLU, K <= 256): the epilogue-bound residual layers.  Three pipeline
// stages instead of five (a tile has <= 8 k-blocks and the epilogue, not the main loop, carries the time) buy a second staging
// buffer per epilogue warp, so that the residual of BOTH 32-column chunks of a warp's tile slice is prefetched with cp.async one
// chunk of work ahead (see tc2_epilogue_role).
template <int BLOCK_N, bool X3, int CG = 1, bool ROWS = false, bool RES2 = false> struct Tc2Cfg {
  static_assert(!RES2 || (!X3 && !ROWS && CG == 2 && BLOCK_N == 128), "RES2 exists for the single-pass 256 x 128 pair kernel only");
  static constexpr int A_BYTES = TC_A_TILE_BYTES;                       // 128 x 32 floats
  static constexpr int B_ROWS = BLOCK_N / CG;                           // weight rows staged by this CTA
  static constexpr int B_BYTES = B_ROWS * TC_BLOCK_K * 4;
  static constexpr int A_SMEM_BYTES = ROWS ? 0 : A_BYTES;                // A_ROWS: the operand goes rows -> registers -> TMEM
  static constexpr int STAGE_BYTES = A_SMEM_BYTES + B_BYTES * (X3 ? 2 : 1);   // X3: raw a, b_hi, b_lo (a_hi / a_lo live in TMEM)
  // splitter (3xTF32) warps, one per TMEM lane quarter; A_ROWS: two such sets of gather warps, alternating k-blocks (one set's
  // table load -> source load -> convert -> TMEM store chain per k-block is ~1000 cycles: 0.60 ms for the stem with one set)
  static constexpr int SPLIT_WARPS = ROWS ? 8 : X3 ? 4 : 0;
  static constexpr int SPLIT_ARRIVALS = (X3 || ROWS) ? 4 : 0;              // warps that fill one stage
#ifndef B2J_TWO_CTAS_MAXN
#define B2J_TWO_CTAS_MAXN 64
#endif
  // Narrow tiles run TWO co-resident CTAs per SM (one epilogue group and 2-3 pipeline stages each instead of two groups and
  // 5-6 stages): one CTA's TMA -> MMA -> epilogue chain leaves the SM idle ~2/3 of the time at N = 64 (~165 cycles per
  // 54-cycle MMA, whatever the pipeline depth or the operand traffic: profiles/r01_patch_kernel.md); a second, independent
  // chain on the same SM fills the gaps: stem 0.474 -> 0.346 ms, stage-0 3x3 0.248 -> 0.172 ms.
  static constexpr bool TWO_CTAS = !X3 && !ROWS && BLOCK_N <= B2J_TWO_CTAS_MAXN;
  static constexpr int EPI_GROUPS = (X3 || TWO_CTAS || ROWS) ? 1 : 2;
#ifndef B2J_X3_WIDE_GROUP
#define B2J_X3_WIDE_GROUP 1
#endif
  // warps per epilogue group: 4 lane quarters x COL_PARTS column slices.  3xTF32 on 128-wide tiles uses 16 warps (4 slices of
  // 32 columns): each thread then carries 32 promotion accumulators instead of 64 (the 8-warp version spilled: 120 B of
  // stack, LDL stalls in the epilogue) and twice as many residual loads are in flight per SM.
  static constexpr int GROUP_WARPS = (X3 && BLOCK_N >= 128 && B2J_X3_WIDE_GROUP) ? 16 : 8;
  static constexpr int COL_PARTS = GROUP_WARPS / 4;
  static constexpr int COLS_PER_WARP = BLOCK_N / COL_PARTS;
  static constexpr int EPI_WARPS = GROUP_WARPS * EPI_GROUPS;
  static constexpr int THREADS = (2 + SPLIT_WARPS + EPI_WARPS) * 32;
  static constexpr int KC = 2;                                           // X3: k-blocks (of 32) per promotion chunk
  static constexpr int EPI_PITCH = 36;
  static constexpr int EPI_CHUNKS = (X3 || RES2) ? COLS_PER_WARP / 32 : 1;   // X3 stages its whole register accumulator at once; RES2: one residual buffer per chunk
  static constexpr int EPI_BYTES = EPI_WARPS * 32 * EPI_PITCH * 4 * EPI_CHUNKS;
#ifndef B2J_X3_STAGES64
#define B2J_X3_STAGES64 5
#endif
  static constexpr int STAGES = RES2 ? 3 : ROWS ? 6 : X3 ? (BLOCK_N <= 64 ? B2J_X3_STAGES64 : 4) : TWO_CTAS ? (BLOCK_N == 64 ? 3 : 2) : (BLOCK_N == 64 ? 6 : (A_BYTES + B_BYTES <= 24576 ? 5 : 4));
  // TMEM: two accumulators; 3xTF32 adds one split activation operand per pipeline stage behind them: a_hi in 32 columns
  // (128 rows x 32 K-elements, row = lane, K-element = column), a_lo in the next 32
  static constexpr int A_TMEM_COL0 = 2 * BLOCK_N;
  static constexpr int A_TMEM_COLS = (X3 ? 2 : 1) * TC_BLOCK_K;
  static constexpr int TMEM_USED = 2 * BLOCK_N + ((X3 || ROWS) ? STAGES * A_TMEM_COLS : 0);
  // A_ROWS: three staged-row buffers (one tile's KH input rows each) + the k -> offset table (1 KB) and the uint8 -> f32 table (1 KB)
  static constexpr int ROWBUF_BYTES = 28 * 1024;                          // 7 rows x 4 boxes x 1 KB: a full 128-pixel tile of a 7x7 / 2 stem
  static constexpr int NRB = 3;                                           // row buffers: tiles whose input rows are in flight / in use
  static constexpr int ROWS_BYTES = ROWS ? NRB * ROWBUF_BYTES + 2048 : 0;
  // A_ROWS keeps the whole (small) weight matrix resident in shared memory -- one slot per k-block, loaded once per CTA -- instead
  // of streaming k-blocks through the stage ring (a ring of 6 weight tiles was latency-bound: 0.45 ms for the stem); STAGES then
  // only counts the TMEM operand slots
  static constexpr int RES_SLOTS = X3 ? 5 : 8;
  static constexpr int PIPE_BYTES = (ROWS ? RES_SLOTS : STAGES) * STAGE_BYTES;
  static constexpr int ROWBUF_OFF = PIPE_BYTES, ROWTAB_OFF = PIPE_BYTES + NRB * ROWBUF_BYTES;
  static constexpr int EPI_OFF = PIPE_BYTES + ROWS_BYTES;
  static constexpr int TMEM_COLS = TMEM_USED <= 128 ? 128 : TMEM_USED <= 256 ? 256 : 512;   // allocations are powers of two
  static constexpr int OPND_BYTES = EPI_GROUPS * B2J_EPI_MAX_STEPS * BLOCK_N * 4;   // decoded per-column epilogue operands
  static constexpr int SMEM_BYTES = PIPE_BYTES + ROWS_BYTES + EPI_BYTES + OPND_BYTES + 1024 + 256;
  static_assert(SMEM_BYTES <= (TWO_CTAS ? 115712 : 232448), "exceeds the shared memory of one SM (227 KB, or 113 KB each for two co-resident CTAs)");
  static_assert(!X3 || COLS_PER_WARP <= 64, "3xTF32 keeps its accumulator slice (COLS_PER_WARP columns per thread) in registers");
  static_assert(COLS_PER_WARP % 32 == 0, "epilogue chunks are 32 columns wide");
  static_assert(TMEM_USED <= 512, "TMEM");
  static_assert(8 * (4 * STAGES + 5) <= 256, "barrier block");
  static_assert(!ROWS || (CG == 1 && BLOCK_N == 64 && STAGES >= 2 * NRB), "A_ROWS: 64-wide single-CTA tiles; barriers full[0..2*NRB) are reused for the row buffers");
};

__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_load_im2col_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c, int w, int h, int n,
                                                   uint16_t off_w, uint16_t off_h) {
  asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.im2col.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2], {%7, %8};"
               ::"r"(dst), "l"(map), "r"(bar), "r"(c), "r"(w), "r"(h), "r"(n), "h"(off_w), "h"(off_h) : "memory");
}
__device__ __forceinline__ void tma_load_tile_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
               ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
// cta_group::2 flavours: executed by both CTAs of a pair, the transaction bytes are credited to the LEADER CTA's barrier
// (same smem offset, peer bit cleared -- cute::Sm100MmaPeerBitMask)
constexpr uint32_t PEER_BIT_MASK = 0xFEFFFFFFu;
__device__ __forceinline__ void tma_load_2d_2sm(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(dst), "l"(map), "r"(bar & PEER_BIT_MASK), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_load_im2col_4d_2sm(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c, int w, int h, int n,
                                                       uint16_t off_w, uint16_t off_h) {
  asm volatile("cp.async.bulk.tensor.4d.im2col.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2], {%7, %8};"
               ::"r"(dst), "l"(map), "r"(bar & PEER_BIT_MASK), "r"(c), "r"(w), "r"(h), "r"(n), "h"(off_w), "h"(off_h) : "memory");
}
__device__ __forceinline__ void umma_tf32_2sm(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n\t}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc),
      "r"(idesc), "r"(accumulate), "r"(0u) : "memory");
}
// A operand from TMEM (.ts form): rows = TMEM lanes of the issuing CTA (of each CTA of the pair for cta_group::2), the 8
// K-elements of one instruction = 8 consecutive 32-bit columns starting at tmem_a
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc),
      "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_tf32_ts_2sm(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], [%1], %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n\t}" ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc),
      "r"(idesc), "r"(accumulate), "r"(0u) : "memory");
}
// registers -> TMEM: lane i of the warp writes its 32 values to columns taddr.col .. +31 of TMEM lane taddr.lane + i
// (a warp can only reach the lane quarter 32 * (warp id % 4))
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]),
        "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]),
        "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// MMA completion -> the barrier at this smem offset in BOTH CTAs of the pair
__device__ __forceinline__ void umma_commit_2sm(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"((uint16_t)3) : "memory");
}
template <int COLS> __device__ __forceinline__ void tmem_alloc_2sm(uint32_t dst_smem) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "n"(COLS) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <int COLS> __device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(COLS) : "memory");
}
__device__ __forceinline__ bool elect_one() {          // one lane of the (converged) warp
  uint32_t p;
  asm volatile("{\n\t.reg .pred P1;\n\telect.sync _|P1, 0xffffffff;\n\tselp.b32 %0, 1, 0, P1;\n\t}" : "=r"(p));
  return p != 0u;
}
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the barrier at the same smem offset in CTA `rank` of the cluster
// Default (.release.cta) semantics on purpose: `.release.cluster` compiles to MEMBAR.ALL.GPU + ERRBAR, which stalled every
// arriving warp for ~1000 cycles (ncu source view of the 3xTF32 kernel, round 2: 60 % of the splitter warps' samples sat on
// that ERRBAR, and the splitter is on the TMA -> split -> MMA critical path).  What the arrivals publish is ordered by
// narrower fences issued just before them: tcgen05.fence::before_thread_sync for TMEM reads / writes, fence.proxy.async for
// the splitters' shared-memory stores (read by this SM's tensor core on behalf of the leader's MMA).
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar, uint32_t rank) {
#ifdef B2J_REMOTE_ARRIVE_RELEASE_CLUSTER
  asm volatile("{\n\t.reg .b32 ra;\n\tmapa.shared::cluster.u32 ra, %0, %1;\n\t"
               "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n\t}" ::"r"(bar), "r"(rank) : "memory");
#else
  asm volatile("{\n\t.reg .b32 ra;\n\tmapa.shared::cluster.u32 ra, %0, %1;\n\t"
               "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}" ::"r"(bar), "r"(rank) : "memory");
#endif
}
__device__ __forceinline__ void tma_prefetch_l2_2d(const CUtensorMap* map, int c0, int c1) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(map), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

// ---- epilogue programs -------------------------------------------------------------------------------
// The fused epilogue is a step program (b2j_epilogue).  The programs ResNet / MLP inference actually produce are
// compiled as straight-line code (no per-step decode, no switch); anything else runs the generic interpreter.
// Both evaluate the same separately-rounded fp32 operations in the same order, so they are bit-identical.
enum {
  EPROG_GENERIC = 0,
  EPROG_BN = 1,            // (acc - mean[c]) * inv[c] + offset[c]
  EPROG_BN_RELU = 2,       // ... max imm
  EPROG_BN_ADD_RELU = 3,   // ... + residual[m, c], max imm
  EPROG_BIAS = 4,          // acc + b[c]
  EPROG_BIAS_RELU = 5      // ... max imm
};

static int classify_epilogue(const b2j_epilogue& e) {
  auto is = [&](uint32_t s, uint32_t op, uint32_t kind) {
    if (s >= e.n_steps || e.steps[s].op != op || e.steps[s].kind != kind) return false;
    return op != B2J_OP_SUB_F || !(e.steps[s].flags & B2J_STEP_SWAP);      // add / mul / max commute, sub does not
  };
  const bool bn = is(0, B2J_OP_SUB_F, B2J_EPK_CHANNEL) && is(1, B2J_OP_MUL_F, B2J_EPK_CHANNEL) && is(2, B2J_OP_ADD_F, B2J_EPK_CHANNEL);
  if (bn && e.n_steps == 3) return EPROG_BN;
  if (bn && e.n_steps == 4 && is(3, B2J_OP_MAX_F, B2J_EPK_IMM)) return EPROG_BN_RELU;
  if (bn && e.n_steps == 5 && is(3, B2J_OP_ADD_F, B2J_EPK_FULL) && is(4, B2J_OP_MAX_F, B2J_EPK_IMM)) return EPROG_BN_ADD_RELU;
  if (e.n_steps == 1 && is(0, B2J_OP_ADD_F, B2J_EPK_CHANNEL)) return EPROG_BIAS;
  if (e.n_steps == 2 && is(0, B2J_OP_ADD_F, B2J_EPK_CHANNEL) && is(1, B2J_OP_MAX_F, B2J_EPK_IMM)) return EPROG_BIAS_RELU;
  return EPROG_GENERIC;
}

// named barrier over the 8 warps of one epilogue group (immediate ids: a register id makes ptxas reserve all 16)
template <int THREADS = 256>
__device__ __forceinline__ void group_sync(int grp) {
  if (grp == 0) asm volatile("bar.sync 1, %0;" ::"n"(THREADS) : "memory");
  else asm volatile("bar.sync 2, %0;" ::"n"(THREADS) : "memory");
}
// Round to nearest TF32, ties away from zero (B2J_CT_ROUND_OUT_TF32): add half a TF32 ulp to the magnitude bits and clear the 13
// low mantissa bits.  Bit-identical to cvt.rna.tf32.f32 for every finite value, +-Inf (stays Inf; the largest finite values round
// up to Inf as they should) and quiet NaNs (stay NaN) -- only a signalling NaN whose payload sits entirely in the low 12 bits
// would turn into Inf, and arithmetic results are never signalling NaNs.  cvt.rna.tf32.f32 compiles to four instructions per
// element (FSETP + IMAD + LOP3 + select), this to two; the epilogue of every single-pass layer rounds every output element.
__device__ __forceinline__ uint32_t rna_tf32_bits(uint32_t x) { return (x + 0x1000u) & 0xFFFFE000u; }
__device__ __forceinline__ float rna_tf32_fast(float x) { return __uint_as_float(rna_tf32_bits(__float_as_uint(x))); }
__device__ __forceinline__ float4 rna4(float4 a) {
#ifdef B2J_RNA_CVT
  return make_float4(__uint_as_float(cvt_tf32(__float_as_uint(a.x))), __uint_as_float(cvt_tf32(__float_as_uint(a.y))),
                     __uint_as_float(cvt_tf32(__float_as_uint(a.z))), __uint_as_float(cvt_tf32(__float_as_uint(a.w))));
#else
  return make_float4(rna_tf32_fast(a.x), rna_tf32_fast(a.y), rna_tf32_fast(a.z), rna_tf32_fast(a.w));
#endif
}
// output store: plain, or cache-streaming (evict-first) when the output is far larger than L2 and would only push
// operands out of it (B2J_CT_STREAM_OUT, set by the host)
__device__ __forceinline__ void st_out(float* p, const float4& v, bool stream) {
  if (stream) asm volatile("st.global.cs.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
  else *reinterpret_cast<float4*>(p) = v;
}
// Output row pitches that are not a multiple of 4 floats (O % 4 != 0: reference tests/test_conv.py uses O = 33, 11, 7, 38, 2)
// cannot be accessed with 128-bit operations: bit 2 of the epilogue's `rnd_stream` flags selects element-wise, tail-guarded
// accesses for the output store and the residual load (the accumulator columns >= O are zeros: TMA zero-fills the weight rows
// beyond O).
constexpr int EPI_SCALAR_IO = 4;
__device__ __forceinline__ void st_out_any(float* row, uint32_t n, uint32_t ldo, const float4& v, int flags) {
  if (!(flags & EPI_SCALAR_IO)) { st_out(row + n, v, (flags & 2) != 0); return; }
  if (n < ldo) row[n] = v.x;
  if (n + 1 < ldo) row[n + 1] = v.y;
  if (n + 2 < ldo) row[n + 2] = v.z;
  if (n + 3 < ldo) row[n + 3] = v.w;
}
__device__ __forceinline__ float4 ld_res_any(const float* row, uint32_t n, uint32_t ldo, int flags) {
  if (!(flags & EPI_SCALAR_IO)) return ld_stream(row + n);
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (n < ldo) v.x = __ldg(row + n);
  if (n + 1 < ldo) v.y = __ldg(row + n + 1);
  if (n + 2 < ldo) v.z = __ldg(row + n + 2);
  if (n + 3 < ldo) v.w = __ldg(row + n + 3);
  return v;
}
__device__ __forceinline__ float max_nan(float a, float b) { float r; asm("max.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b)); return r; }

// Which output row (index into the [M, N] output matrix) a tile-local accumulator row belongs to; ROW_NONE = not stored.
constexpr uint32_t ROW_NONE = 0xFFFFFFFFu;
struct RowLinear {            // rows m_base .. of a linear M tile
  uint32_t m_base, M;
  __device__ __forceinline__ uint32_t operator()(int r) const { const uint32_t m = m_base + (uint32_t)r; return m < M ? m : ROW_NONE; }
};
struct RowPatch {             // conv_patch_kernel: tile row = (output row oh0 + r / P, output column r % P) of image img
  uint32_t row0, P, log2P, OW, rows_left, quarter;     // row0 = (img*OH + oh0)*OW, rows_left = OH - oh0, quarter = 32*q
  __device__ __forceinline__ uint32_t operator()(int r) const {
    const uint32_t ml = quarter + (uint32_t)r, dr = ml >> log2P, px = ml & (P - 1u);
    return (px < OW && dr < rows_left) ? row0 + dr * OW + px : ROW_NONE;
  }
};

// Generic interpreter for one 32x32 chunk.  The step program was decoded once per kernel into `ops` (4 bits per
// step: 0 add, 1 sub, 2 mul, 3 div, 4 max, 5 min, 6 reversed sub, 7 reversed div) and `full_mask` (steps that read
// a full tensor, i.e. the residual); immediates and per-channel vectors were expanded into the shared-memory
// table `opnd[step][column]`.
template <int PITCH, int BLOCK_N, typename RM>
__device__ __forceinline__ void epilogue_chunk_generic(uint32_t n_steps, uint32_t ops, uint32_t full_mask, const EpiPtrs& epi,
                                                       const float* opnd, int col, const float* stg, float* __restrict__ out,
                                                       const RM& rm, uint32_t n, uint32_t ldo, int lane, int rnd_stream) {
  const bool rnd = (rnd_stream & 1) != 0;
  const int cj = lane & 7, rr = lane >> 3;
  float4 v[8];
#pragma unroll
  for (int it = 0; it < 8; ++it) v[it] = *reinterpret_cast<const float4*>(stg + (rr + 4 * it) * PITCH + 4 * cj);
#pragma unroll 1
  for (uint32_t s = 0; s < n_steps; ++s) {
    float4 b[8];
    if ((full_mask >> s) & 1u) {
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        const uint32_t m = rm(rr + 4 * it);
        b[it] = m != ROW_NONE ? ld_res_any(epi.p[s] + (uint64_t)m * ldo, n, ldo, rnd_stream) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    } else {
      const float4 t = *reinterpret_cast<const float4*>(opnd + s * BLOCK_N + col);
#pragma unroll
      for (int it = 0; it < 8; ++it) b[it] = t;
    }
#define B2J_FAST_CASE(CODE, EXPR)                                                                       \
      case CODE:                                                                                        \
        _Pragma("unroll") for (int it = 0; it < 8; ++it) {                                              \
          float4& a = v[it]; const float4 c = b[it];                                                    \
          a.x = EXPR(a.x, c.x); a.y = EXPR(a.y, c.y); a.z = EXPR(a.z, c.z); a.w = EXPR(a.w, c.w);       \
        } break;
#define B2J_MAXF(x, y) epi_op(B2J_OP_MAX_F, x, y)
#define B2J_MINF(x, y) epi_op(B2J_OP_MIN_F, x, y)
#define B2J_RSUB(x, y) __fsub_rn(y, x)
#define B2J_RDIV(x, y) __fdiv_rn(y, x)
    switch ((ops >> (4 * s)) & 15u) {
      B2J_FAST_CASE(0, __fadd_rn)
      B2J_FAST_CASE(1, __fsub_rn)
      B2J_FAST_CASE(2, __fmul_rn)
      B2J_FAST_CASE(3, __fdiv_rn)
      B2J_FAST_CASE(4, B2J_MAXF)
      B2J_FAST_CASE(5, B2J_MINF)
      B2J_FAST_CASE(6, B2J_RSUB)
      B2J_FAST_CASE(7, B2J_RDIV)
      default: break;
    }
#undef B2J_FAST_CASE
#undef B2J_MAXF
#undef B2J_MINF
#undef B2J_RSUB
#undef B2J_RDIV
  }
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    const uint32_t m = rm(rr + 4 * it);
    if (rnd) v[it] = rna4(v[it]);
    if (m != ROW_NONE) st_out_any(out + (uint64_t)m * ldo, n, ldo, v[it], rnd_stream);
  }
}

// Straight-line epilogue of one 32x32 chunk for the programs above.  `res` holds this thread's residual values
// (8 rows x 4 channels; rows 0-3 in res_a were requested before the TMEM load, rows 4-7 in res_b right after it, so
// the loads fly while the accumulator chunk is staged through shared memory).
template <int PROG, int PITCH, int BLOCK_N, typename RM>
__device__ __forceinline__ void epilogue_chunk_spec(const float* opnd, float relu_imm, const float4 (&res_a)[4], const float4 (&res_b)[4], int col, const float* stg,
                                                    float* __restrict__ out, const RM& rm, uint32_t n, uint32_t ldo, int lane, int rnd_stream) {
  const bool rnd = (rnd_stream & 1) != 0;
  const int cj = lane & 7, rr = lane >> 3;
  constexpr bool BN = PROG == EPROG_BN || PROG == EPROG_BN_RELU || PROG == EPROG_BN_ADD_RELU;
  constexpr bool RELU = PROG == EPROG_BN_RELU || PROG == EPROG_BN_ADD_RELU || PROG == EPROG_BIAS_RELU;
  const float4 o0 = *reinterpret_cast<const float4*>(opnd + 0 * BLOCK_N + col);
  float4 o1 = o0, o2 = o0;
  if (BN) {
    o1 = *reinterpret_cast<const float4*>(opnd + 1 * BLOCK_N + col);
    o2 = *reinterpret_cast<const float4*>(opnd + 2 * BLOCK_N + col);
  }
  // two batches of 4 rows: the 4 shared-memory loads of a batch are issued back to back (one exposed LDS latency per
  // batch instead of one per row), within the 96 registers 18 warps per SM leave each thread
#pragma unroll
  for (int hb = 0; hb < 2; ++hb) {
    float4 v[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] = *reinterpret_cast<const float4*>(stg + (rr + 4 * (4 * hb + i)) * PITCH + 4 * cj);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float4 a = v[i];
      if (BN) {
        a.x = __fadd_rn(__fmul_rn(__fsub_rn(a.x, o0.x), o1.x), o2.x);
        a.y = __fadd_rn(__fmul_rn(__fsub_rn(a.y, o0.y), o1.y), o2.y);
        a.z = __fadd_rn(__fmul_rn(__fsub_rn(a.z, o0.z), o1.z), o2.z);
        a.w = __fadd_rn(__fmul_rn(__fsub_rn(a.w, o0.w), o1.w), o2.w);
      } else {
        a.x = __fadd_rn(a.x, o0.x); a.y = __fadd_rn(a.y, o0.y); a.z = __fadd_rn(a.z, o0.z); a.w = __fadd_rn(a.w, o0.w);
      }
      if (PROG == EPROG_BN_ADD_RELU) {
        const float4 r = hb == 0 ? res_a[i] : res_b[i];
        a.x = __fadd_rn(a.x, r.x); a.y = __fadd_rn(a.y, r.y); a.z = __fadd_rn(a.z, r.z); a.w = __fadd_rn(a.w, r.w);
      }
      if (RELU) { a.x = max_nan(a.x, relu_imm); a.y = max_nan(a.y, relu_imm); a.z = max_nan(a.z, relu_imm); a.w = max_nan(a.w, relu_imm); }
      if (rnd) a = rna4(a);
      const uint32_t m = rm(rr + 4 * (4 * hb + i));
      if (m != ROW_NONE) st_out_any(out + (uint64_t)m * ldo, n, ldo, a, rnd_stream);
    }
  }
}

__device__ __forceinline__ void cp_async16(uint32_t dst_smem, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst_smem), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_but_last() { asm volatile("cp.async.wait_group 1;" ::: "memory"); }   // all groups but the most recent one

// 3xTF32 residual epilogue (BN + residual + ReLU) of one 32 x 32 chunk whose RESIDUAL already sits in the warp's staging
// buffer (cp.async, requested at the start of the tile).  Pass 1, row per lane (the TMEM layout of the register accumulator):
// out = max((acc - mean[c]) * inv[c] + offset[c] + residual, imm), written back over the residual; per-column operands are
// broadcast reads of the operand table.  Pass 2, coalesced: 8 lanes per 128-byte row segment -> global.  Same fp32
// operations in the same order as epilogue_chunk_spec, so the two paths are bit-identical.
template <int PITCH, int BLOCK_N, bool BATCHED = false, typename RM>
__device__ __forceinline__ void epilogue_chunk_res_smem(const float* opnd, float relu_imm, const float (&acc)[32], int col0, float* stg,
                                                        float* __restrict__ out, const RM& rm, uint32_t n, uint32_t O, int lane, int rnd_stream) {
  const bool rnd = (rnd_stream & 1) != 0;
  const int cj = lane & 7, rr = lane >> 3;
  // all loads first, all stores last: with the store of slot j between the loads of j and j + 1 the compiler has to assume the
  // operand table aliases the staging buffer and serialises the eight iterations on their shared-memory latency (ncu source view
  // of the single-pass residual layers: 30 % short-scoreboard stalls on the first FADD of every iteration)
  // (BATCHED; the 3xTF32 kernels have 80 registers per thread and keep the interleaved loop: the batched form spills there)
  if (!BATCHED) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float4 o0 = *reinterpret_cast<const float4*>(opnd + 0 * BLOCK_N + col0 + 4 * j);
      const float4 o1 = *reinterpret_cast<const float4*>(opnd + 1 * BLOCK_N + col0 + 4 * j);
      const float4 o2 = *reinterpret_cast<const float4*>(opnd + 2 * BLOCK_N + col0 + 4 * j);
      float4* slot = reinterpret_cast<float4*>(stg + lane * PITCH + 4 * j);
      const float4 r = *slot;
      float4 a;
      a.x = __fadd_rn(__fadd_rn(__fmul_rn(__fsub_rn(acc[4 * j], o0.x), o1.x), o2.x), r.x);
      a.y = __fadd_rn(__fadd_rn(__fmul_rn(__fsub_rn(acc[4 * j + 1], o0.y), o1.y), o2.y), r.y);
      a.z = __fadd_rn(__fadd_rn(__fmul_rn(__fsub_rn(acc[4 * j + 2], o0.z), o1.z), o2.z), r.z);
      a.w = __fadd_rn(__fadd_rn(__fmul_rn(__fsub_rn(acc[4 * j + 3], o0.w), o1.w), o2.w), r.w);
      a.x = max_nan(a.x, relu_imm); a.y = max_nan(a.y, relu_imm); a.z = max_nan(a.z, relu_imm); a.w = max_nan(a.w, relu_imm);
      if (rnd) a = rna4(a);
      *slot = a;
    }
  } else {
  float4 res[8], a[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) res[j] = *reinterpret_cast<const float4*>(stg + lane * PITCH + 4 * j);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float4 o0 = *reinterpret_cast<const float4*>(opnd + 0 * BLOCK_N + col0 + 4 * j);
    const float4 o1 = *reinterpret_cast<const float4*>(opnd + 1 * BLOCK_N + col0 + 4 * j);
    const float4 o2 = *reinterpret_cast<const float4*>(opnd + 2 * BLOCK_N + col0 + 4 * j);
    a[j].x = __fadd_rn(__fadd_rn(__fmul_rn(__fsub_rn(acc[4 * j], o0.x), o1.x), o2.x), res[j].x);
    a[j].y = __fadd_rn(__fadd_rn(__fmul_rn(__fsub_rn(acc[4 * j + 1], o0.y), o1.y), o2.y), res[j].y);
    a[j].z = __fadd_rn(__fadd_rn(__fmul_rn(__fsub_rn(acc[4 * j + 2], o0.z), o1.z), o2.z), res[j].z);
    a[j].w = __fadd_rn(__fadd_rn(__fmul_rn(__fsub_rn(acc[4 * j + 3], o0.w), o1.w), o2.w), res[j].w);
    a[j].x = max_nan(a[j].x, relu_imm); a[j].y = max_nan(a[j].y, relu_imm); a[j].z = max_nan(a[j].z, relu_imm); a[j].w = max_nan(a[j].w, relu_imm);
    if (rnd) a[j] = rna4(a[j]);
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) *reinterpret_cast<float4*>(stg + lane * PITCH + 4 * j) = a[j];
  }
  __syncwarp();
  if (n < O) {
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      const uint32_t m = rm(rr + 4 * it);
      const float4 v = *reinterpret_cast<const float4*>(stg + (rr + 4 * it) * PITCH + 4 * cj);
      if (m != ROW_NONE) st_out(out + (uint64_t)m * O + n, v, (rnd_stream & 2) != 0);
    }
  }
}

// ---- TMA-store epilogue ------------------------------------------------------------------------------
// Programs without a full-tensor operand (BN, BN + ReLU, bias, bias + ReLU) are evaluated in the ROW-PER-LANE layout the
// accumulator arrives in (TMEM lane = tile row): the per-column operands are broadcast reads of the operand table, the finished
// 32 x 32 chunk goes into the warp's staging buffer once, in the 128-byte-swizzled layout of a TMA box, and ONE
// cp.async.bulk.tensor store per chunk writes it out (rows / columns beyond the tensor are clipped by the TMA unit).  Against the
// st.global path this drops the transposing second pass through shared memory, the per-row predicates and the 64-bit address
// arithmetic: ~20 % fewer epilogue instructions per chunk (the stem kernel is instruction-bound).  Same fp32 operations in the
// same order, so the result is bit-identical.
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, uint32_t src_smem, int c0, int c1, int c2, bool evict_first) {
  if (evict_first) {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group.L2::cache_hint [%0, {%2, %3, %4}], [%1], %5;"
                 ::"l"(map), "r"(src_smem), "r"(c0), "r"(c1), "r"(c2), "l"(pol) : "memory");
  } else {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
                 ::"l"(map), "r"(src_smem), "r"(c0), "r"(c1), "r"(c2) : "memory");
  }
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

template <int PROG, int BLOCK_N>
__device__ __forceinline__ void epilogue_chunk_tma(const float* opnd, float relu_imm, const uint32_t (&r)[32], int col0, uint32_t stg_s,
                                                   const CUtensorMap* tmap_out, int c0, int c1, int c2, int lane, int rnd_stream) {
  const bool rnd = (rnd_stream & 1) != 0;
  constexpr bool BN = PROG == EPROG_BN || PROG == EPROG_BN_RELU;
  constexpr bool RELU = PROG == EPROG_BN_RELU || PROG == EPROG_BIAS_RELU;
  // the previous chunk's store must have finished READING the staging buffer before it is overwritten
  if (lane == 0) tma_store_wait_read();
  __syncwarp();
  const uint32_t row_s = stg_s + (uint32_t)lane * 128u;
  const uint32_t sw = (row_s >> 7) & 7u;                      // SWIZZLE_128B: 16-byte chunk j of a row sits at chunk j ^ (address bits 7..9)
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float4 o0 = *reinterpret_cast<const float4*>(opnd + 0 * BLOCK_N + col0 + 4 * j);
    float4 a = make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]), __uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3]));
    if (BN) {
      const float4 o1 = *reinterpret_cast<const float4*>(opnd + 1 * BLOCK_N + col0 + 4 * j);
      const float4 o2 = *reinterpret_cast<const float4*>(opnd + 2 * BLOCK_N + col0 + 4 * j);
      a.x = __fadd_rn(__fmul_rn(__fsub_rn(a.x, o0.x), o1.x), o2.x);
      a.y = __fadd_rn(__fmul_rn(__fsub_rn(a.y, o0.y), o1.y), o2.y);
      a.z = __fadd_rn(__fmul_rn(__fsub_rn(a.z, o0.z), o1.z), o2.z);
      a.w = __fadd_rn(__fmul_rn(__fsub_rn(a.w, o0.w), o1.w), o2.w);
    } else {
      a.x = __fadd_rn(a.x, o0.x); a.y = __fadd_rn(a.y, o0.y); a.z = __fadd_rn(a.z, o0.z); a.w = __fadd_rn(a.w, o0.w);
    }
    if (RELU) { a.x = max_nan(a.x, relu_imm); a.y = max_nan(a.y, relu_imm); a.z = max_nan(a.z, relu_imm); a.w = max_nan(a.w, relu_imm); }
    if (rnd) a = rna4(a);
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(row_s + ((((uint32_t)j) ^ sw) << 4)), "f"(a.x), "f"(a.y), "f"(a.z), "f"(a.w) : "memory");
  }
  fence_proxy_async();                                        // generic-proxy stores -> visible to the TMA unit
  __syncwarp();
  if (lane == 0) tma_store_3d(tmap_out, stg_s, c0, c1, c2, (rnd_stream & 2) != 0);
}

struct Tc2EpiCtx {
  uint8_t* smem_gen;
  uint32_t bar_base, tmem_base;
  float* out;
  uint32_t M, num_kb, tiles_n, num_tiles;
  uint32_t first_tile, tile_step, cta_rank;     // this CTA (pair) walks tiles first_tile, first_tile + tile_step, ...
  uint32_t rows_tiles_w;                        // A_ROWS: 128-pixel tiles per output row
  const CUtensorMap* tmap_out;                  // TMA-store epilogue: [lines][pixels][channels] view of the output, or nullptr
};

// The epilogue role of conv_tc2_kernel for one epilogue program (see the kernel's header comment).
template <int BLOCK_N, bool X3, int CG, bool ROWS, int PROG, bool RES2 = false>
__device__ __forceinline__ void tc2_epilogue_role(const b2j_conv_tc_params& p, const EpiPtrs& epi, const Tc2EpiCtx& cx) {
  using Cfg = Tc2Cfg<BLOCK_N, X3, CG, ROWS, RES2>;
  constexpr int PIPE_BYTES = Cfg::EPI_OFF;          // the epilogue staging starts behind the pipeline stages (and the A_ROWS buffers)
  constexpr int COLS_PER_WARP = Cfg::COLS_PER_WARP;
  constexpr int GROUP_THREADS = Cfg::GROUP_WARPS * 32;
  constexpr bool HAS_RES = PROG == EPROG_BN_ADD_RELU;
  const uint32_t tfull0 = cx.bar_base + 8u * (4 * Cfg::STAGES), tempty0 = tfull0 + 16u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ew = warp - 2 - Cfg::SPLIT_WARPS;  // 0 .. EPI_WARPS-1
  const int grp = ew / Cfg::GROUP_WARPS;       // epilogue group = TMEM accumulator it drains (single-pass mode)
  const int q = warp & 3;                      // TMEM lane quarter accessible to this warp
  const int half = (ew % Cfg::GROUP_WARPS) >> 2;   // column slice (of COL_PARTS)
  float* stg0 = reinterpret_cast<float*>(cx.smem_gen + PIPE_BYTES) + ew * 32 * Cfg::EPI_PITCH * Cfg::EPI_CHUNKS;
  float* opnd = reinterpret_cast<float*>(cx.smem_gen + PIPE_BYTES + Cfg::EPI_BYTES) + grp * B2J_EPI_MAX_STEPS * BLOCK_N;
  const uint32_t M = cx.M, num_kb = cx.num_kb;
  float* __restrict__ out = cx.out;
  const uint32_t n_steps = p.epi.n_steps;
  uint32_t ops = 0, full_mask = 0;
  if (PROG == EPROG_GENERIC) {                 // decode the step program once for the interpreter
    for (uint32_t s = 0; s < n_steps; ++s) {
      const b2j_epi_step st = p.epi.steps[s];
      const bool sw = st.flags & B2J_STEP_SWAP;
      uint32_t code = st.op == B2J_OP_ADD_F ? 0u : st.op == B2J_OP_SUB_F ? (sw ? 6u : 1u) : st.op == B2J_OP_MUL_F ? 2u
                    : st.op == B2J_OP_DIV_F ? (sw ? 7u : 3u) : st.op == B2J_OP_MAX_F ? 4u : st.op == B2J_OP_MIN_F ? 5u : 15u;
      ops |= code << (4 * s);
      if (st.kind == B2J_EPK_FULL) full_mask |= 1u << s;
    }
  }
  const float relu_imm = n_steps ? __uint_as_float(p.epi.steps[n_steps - 1].imm) : 0.0f;
  const int rnd = (int)(p.flags & 3u) | ((p.o & 3u) ? EPI_SCALAR_IO : 0);   // bit 0: B2J_CT_ROUND_OUT_TF32, bit 1: B2J_CT_STREAM_OUT
  const float* resp = HAS_RES ? epi.p[3] : nullptr;
  const int gtid = (ew % Cfg::GROUP_WARPS) * 32 + lane;       // thread index within the group
  const int cj = lane & 7, rr = lane >> 3;
  uint32_t table_n0 = 0xFFFFFFFFu;
  uint32_t tile_i = 0, chunk = 0;
  bool res2_primed = false;
  for (uint32_t t = cx.first_tile; t < cx.num_tiles; t += cx.tile_step, ++tile_i) {
    if (Cfg::EPI_GROUPS == 2 && (tile_i & 1u) != (uint32_t)grp) continue;
    uint32_t m0, m_end, n0;
    if (ROWS) {        // tile = (image, output row, 128-pixel segment of it): rows beyond the segment are not stored
      const uint32_t wseg = t % cx.rows_tiles_w, line = t / cx.rows_tiles_w;      // line = image * OH + output row
      m0 = line * p.ow + wseg * TC_BLOCK_M;
      m_end = line * p.ow + (p.ow < (wseg + 1) * TC_BLOCK_M ? p.ow : (wseg + 1) * TC_BLOCK_M);
      n0 = 0;
    } else {
      m0 = (t / cx.tiles_n) * (TC_BLOCK_M * CG) + cx.cta_rank * TC_BLOCK_M; m_end = M; n0 = (t % cx.tiles_n) * BLOCK_N;
    }
    const RowLinear rm{m0 + (uint32_t)q * 32u, m_end};
    if (n0 != table_n0) {
      // (re)build opnd[step][column] for this column range: immediates broadcast, per-channel vectors copied
      group_sync<GROUP_THREADS>(grp);                                   // everybody is done reading the old table
      for (uint32_t idx = gtid; idx < n_steps * BLOCK_N; idx += GROUP_THREADS) {
        const uint32_t s = idx / BLOCK_N, c = idx - s * BLOCK_N;
        const b2j_epi_step st = p.epi.steps[s];
        float val = 0.0f;
        if (st.kind == B2J_EPK_IMM) val = __uint_as_float(st.imm);
        else if (st.kind == B2J_EPK_CHANNEL && n0 + c < p.o) val = __ldg(epi.p[s] + n0 + c);
        opnd[idx] = val;
      }
      group_sync<GROUP_THREADS>(grp);
      table_n0 = n0;
    }
    if constexpr (RES2 && HAS_RES) {
      // Residual tiles of the epilogue-bound layers: the residual of each of the warp's two 32 x 32 chunks lands in its own
      // staging buffer by cp.async, requested one chunk of work ahead -- chunk c of the group's NEXT tile as soon as chunk c of
      // this tile has left its buffer -- so no thread ever waits for a residual load (ncu source view of the register-load
      // version: 40 % of the epilogue warps' samples were long-scoreboard stalls on those loads).  The accumulator chunk comes
      // out of TMEM row per lane and is combined with the residual in place (epilogue_chunk_res_smem, the 3xTF32 path's
      // routine: same fp32 operations in the same order as epilogue_chunk_spec, bit-identical), then stored coalesced.
      // One cp.async group per request, also when nothing is requested, so that "all groups but the last" is always the
      // chunk about to be used.
      auto request = [&](uint32_t m0_, uint32_t n0_, int cc) {
        const RowLinear rq{m0_ + (uint32_t)q * 32u, M};
        const uint32_t n = n0_ + (uint32_t)(half * COLS_PER_WARP + 32 * cc + 4 * cj);
        float* buf = stg0 + cc * 32 * Cfg::EPI_PITCH;
        if (m0_ != ROW_NONE && n < p.o) {
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            const uint32_t m = rq(rr + 4 * it);
            if (m != ROW_NONE) cp_async16(smem_u32(buf + (rr + 4 * it) * Cfg::EPI_PITCH + 4 * cj), resp + (uint64_t)m * p.o + n);
          }
        }
        cp_async_commit();
      };
      if (!res2_primed) { request(m0, n0, 0); request(m0, n0, 1); res2_primed = true; }        // the group's first tile
      const uint32_t t_next = t + cx.tile_step * (uint32_t)Cfg::EPI_GROUPS;       // the group's next tile
      const bool has_next = t_next < cx.num_tiles;
      const uint32_t m0_next = has_next ? (t_next / cx.tiles_n) * (TC_BLOCK_M * CG) + cx.cta_rank * TC_BLOCK_M : ROW_NONE;
      const uint32_t n0_next = has_next ? (t_next % cx.tiles_n) * BLOCK_N : 0u;
      chunk = tile_i;
      mbar_wait(tfull0 + 8u * (chunk & 1u), (chunk >> 1) & 1u);
      tc_fence_after();
#pragma unroll 1
      for (int cc = 0; cc < COLS_PER_WARP / 32; ++cc) {
        const int col0 = half * COLS_PER_WARP + 32 * cc;
        uint32_t r[32];
        tmem_ld32(cx.tmem_base + ((uint32_t)(q * 32) << 16) + (chunk & 1u) * BLOCK_N + (uint32_t)col0, r);
        if (cc + 1 == COLS_PER_WARP / 32) {          // last TMEM read of this tile: hand the accumulator back
          tc_fence_before();
          __syncwarp();
          if (lane == 0) { if (CG == 2) mbar_arrive_cluster(tempty0 + 8u * (chunk & 1u), 0); else mbar_arrive(tempty0 + 8u * (chunk & 1u)); }
        }
        cp_async_wait_but_last();
        __syncwarp();
        float acc32[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) acc32[j] = __uint_as_float(r[j]);
        float* buf = stg0 + cc * 32 * Cfg::EPI_PITCH;
        epilogue_chunk_res_smem<Cfg::EPI_PITCH, BLOCK_N, true>(opnd, relu_imm, acc32, col0, buf, out, rm, n0 + (uint32_t)(col0 + 4 * cj), p.o, lane, rnd);
        __syncwarp();                                 // every lane has read its part of the buffer
        request(m0_next, n0_next, cc);
      }
      continue;
    }
    // Experiment (B2J_RES_EARLY=1, off): request the residual of the tile's first 32-column chunk before waiting for the MMAs.
    // Measured SLOWER on ResNet-50 b256 (TF32 7.02 -> 7.15 ms, 3xTF32 17.1 -> 17.5 ms): the early loads race the tile's TMA
    // L2 prefetch (issued by the MMA thread at the same moment) and the residual is then fetched from DRAM twice.
#ifndef B2J_RES_EARLY
#define B2J_RES_EARLY 0
#endif
    float4 res_a[4], res_b[4];
    if (HAS_RES && B2J_RES_EARLY) {
      const uint32_t n = n0 + (uint32_t)(half * COLS_PER_WARP + 4 * cj);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const uint32_t m = rm(rr + 4 * i);
        res_a[i] = (m != ROW_NONE && n < p.o) ? ld_res_any(resp + (uint64_t)m * p.o, n, p.o, rnd) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
      if (!X3) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const uint32_t m = rm(rr + 4 * (i + 4));
          res_b[i] = (m != ROW_NONE && n < p.o) ? ld_res_any(resp + (uint64_t)m * p.o, n, p.o, rnd) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
    }
    constexpr bool TMA_PROG = PROG == EPROG_BN || PROG == EPROG_BN_RELU || PROG == EPROG_BIAS || PROG == EPROG_BIAS_RELU;
    const bool tma_out = TMA_PROG && cx.tmap_out != nullptr;
    // 3xTF32 residual tiles: the warp's 32 x 32 residual chunk goes straight into its (idle) staging buffer with cp.async,
    // requested here, before the tile's partial sums are awaited -- no registers held, 4 KB per warp in flight, and the
    // latency hides behind the main loop.  (With register loads this epilogue spilled the residual and ran every load
    // synchronously: STL stalls on the long scoreboard in the ncu source view.)
    constexpr bool RES_SMEM = X3 && HAS_RES && COLS_PER_WARP == 32;
    const bool res_smem = RES_SMEM && !(rnd & EPI_SCALAR_IO);
    if (res_smem) {
      const uint32_t n = n0 + (uint32_t)(half * COLS_PER_WARP + 4 * cj);
      if (n < p.o) {
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          const uint32_t m = rm(rr + 4 * it);
          if (m != ROW_NONE) cp_async16(smem_u32(stg0 + (rr + 4 * it) * Cfg::EPI_PITCH + 4 * cj), resp + (uint64_t)m * p.o + n);
        }
      }
      cp_async_commit();
    }
    float acc[X3 ? COLS_PER_WARP : 1];
    if (X3) {
      // chunked promotion: add every partial sum the tensor core hands over into fp32 registers
      for (uint32_t kb0 = 0; kb0 < num_kb; kb0 += Cfg::KC, ++chunk) {
        const uint32_t ab = chunk & 1u;
        mbar_wait(tfull0 + 8u * ab, (chunk >> 1) & 1u);
        tc_fence_after();
#pragma unroll
        for (int h = 0; h < COLS_PER_WARP / 16; ++h) {                 // 16 columns at a time: acc[] already fills the registers
          uint32_t r[16];
          tmem_ld16(cx.tmem_base + ((uint32_t)(q * 32) << 16) + ab * BLOCK_N + (uint32_t)(half * COLS_PER_WARP + 16 * h), r);
          if (h == COLS_PER_WARP / 16 - 1) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) { if (CG == 2) mbar_arrive_cluster(tempty0 + 8u * ab, 0); else mbar_arrive(tempty0 + 8u * ab); }
          }
          if (kb0 == 0) {
#pragma unroll
            for (int j = 0; j < 16; ++j) acc[16 * h + j] = __uint_as_float(r[j]);
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) acc[16 * h + j] = __fadd_rn(acc[16 * h + j], __uint_as_float(r[j]));
          }
        }
      }
    } else {
      chunk = tile_i;
      mbar_wait(tfull0 + 8u * (chunk & 1u), (chunk >> 1) & 1u);
      tc_fence_after();
    }
    if (RES_SMEM && res_smem) {
      cp_async_wait_all();
      __syncwarp();
      const int col0 = half * COLS_PER_WARP;
      float acc32[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) acc32[j] = acc[X3 ? j : 0];
      epilogue_chunk_res_smem<Cfg::EPI_PITCH, BLOCK_N>(opnd, relu_imm, acc32, col0, stg0, out, rm, n0 + (uint32_t)(col0 + 4 * cj), p.o, lane, rnd);
      __syncwarp();
      continue;
    }
    if (TMA_PROG && tma_out) {
      // TMA-store epilogue (see epilogue_chunk_tma): coordinates of this warp's 32-row box in the [lines][pixels][channels] view
      int c1, c2;
      if (ROWS) { c1 = (int)((t % cx.rows_tiles_w) * TC_BLOCK_M) + q * 32; c2 = (int)(t / cx.rows_tiles_w); }
      else { c1 = (int)m0 + q * 32; c2 = 0; }
      const uint32_t stg_s = smem_u32(stg0);
#pragma unroll 1
      for (int cc = 0; cc < COLS_PER_WARP; cc += 32) {
        const int col0 = half * COLS_PER_WARP + cc;
        uint32_t r[32];
        if (X3) {
#pragma unroll
          for (int j = 0; j < 32; ++j) r[j] = __float_as_uint(acc[X3 ? cc + j : 0]);
        } else {
          const uint32_t ab = chunk & 1u;
          tmem_ld32(cx.tmem_base + ((uint32_t)(q * 32) << 16) + ab * BLOCK_N + (uint32_t)col0, r);
          if (cc + 32 >= COLS_PER_WARP) {          // last TMEM read of this tile: hand the accumulator back
            tc_fence_before();
            __syncwarp();
            if (lane == 0) { if (CG == 2) mbar_arrive_cluster(tempty0 + 8u * ab, 0); else mbar_arrive(tempty0 + 8u * ab); }
          }
        }
#ifndef B2J_DIAG_NO_EPI_TMA          // timing diagnostic only (no output is written): the epilogue drains TMEM and does nothing else
        if (n0 + (uint32_t)col0 < p.o && m0 + (uint32_t)q * 32u < m_end)
          epilogue_chunk_tma<PROG, BLOCK_N>(opnd, relu_imm, r, col0, stg_s, cx.tmap_out, (int)n0 + col0, c1, c2, lane, rnd);
#endif
      }
      continue;
    }
    if (X3) {
      // the register accumulator goes to shared memory in one go, so that it is dead before the epilogue math starts
#pragma unroll
      for (int c2 = 0; c2 < COLS_PER_WARP / 32; ++c2)
#pragma unroll
        for (int j = 0; j < 8; ++j)
          *reinterpret_cast<float4*>(stg0 + c2 * 32 * Cfg::EPI_PITCH + lane * Cfg::EPI_PITCH + 4 * j) =
              make_float4(acc[32 * c2 + 4 * j], acc[32 * c2 + 4 * j + 1], acc[32 * c2 + 4 * j + 2], acc[32 * c2 + 4 * j + 3]);
      __syncwarp();
    }
#pragma unroll 1
    for (int cc = 0; cc < COLS_PER_WARP; cc += 32) {
      float* stg = stg0 + (X3 ? (cc / 32) * 32 * Cfg::EPI_PITCH : 0);
      const int col0 = half * COLS_PER_WARP + cc;
      const int col = col0 + 4 * cj;
      const uint32_t n = n0 + col;
      if (HAS_RES && !(B2J_RES_EARLY && cc == 0)) {     // residual rows 0-3: in flight while the accumulator chunk moves TMEM -> registers -> smem
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const uint32_t m = rm(rr + 4 * i);
          res_a[i] = (m != ROW_NONE && n < p.o) ? ld_res_any(resp + (uint64_t)m * p.o, n, p.o, rnd) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
      if (X3) {
        // already staged above
      } else {
        const uint32_t ab = chunk & 1u;
        uint32_t r[32];
        tmem_ld32(cx.tmem_base + ((uint32_t)(q * 32) << 16) + ab * BLOCK_N + (uint32_t)col0, r);
        if (cc + 32 >= COLS_PER_WARP) {          // last TMEM read of this tile: hand the accumulator back
          tc_fence_before();
          __syncwarp();
          if (lane == 0) { if (CG == 2) mbar_arrive_cluster(tempty0 + 8u * ab, 0); else mbar_arrive(tempty0 + 8u * ab); }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j)
          *reinterpret_cast<uint4*>(stg + lane * Cfg::EPI_PITCH + 4 * j) = make_uint4(r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
      }
      __syncwarp();
      if (HAS_RES && !(B2J_RES_EARLY && !X3 && cc == 0)) {     // rows 4-7: requested now that the TMEM registers are free
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const uint32_t m = rm(rr + 4 * (i + 4));
          res_b[i] = (m != ROW_NONE && n < p.o) ? ld_res_any(resp + (uint64_t)m * p.o, n, p.o, rnd) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
      if (n < p.o) {
        if (PROG == EPROG_GENERIC)
          epilogue_chunk_generic<Cfg::EPI_PITCH, BLOCK_N>(n_steps, ops, full_mask, epi, opnd, col, stg, out, rm, n, p.o, lane, rnd);
        else
          epilogue_chunk_spec<PROG, Cfg::EPI_PITCH, BLOCK_N>(opnd, relu_imm, res_a, res_b, col, stg, out, rm, n, p.o, lane, rnd);
      }
      __syncwarp();
    }
  }
  // outstanding TMA stores read this CTA's shared memory: they must have completed before the CTA may exit
  if (cx.tmap_out != nullptr && lane == 0) tma_store_wait_all();
}

// ---- the kernel --------------------------------------------------------------------------------------
// X3 = false  single-pass TF32.  Warps: 0 TMA producer, 1 MMA issuer, 2..17 epilogue in TWO groups of 8: group g
//             drains TMEM accumulator g, i.e. every other tile of this CTA, so two tiles are in their epilogue at
//             once while the tensor core already works on the next one.
// X3 = true   fp32-class 3xTF32 with chunked promotion.  Warps: 0 TMA (raw fp32 activations + pre-split weights
//             Wt_hi/Wt_lo), 1 MMA, 2..5 splitters (rewrite each landed activation stage in place as hi = rna_tf32(a)
//             and write lo = a - hi next to it), 6..13 epilogue.  Per k-step the MMA warp issues a_lo*b_hi, a_hi*b_lo,
//             a_hi*b_hi.  The tensor core accumulates with truncation, which drifts ~1e-5 relative over hundreds of
//             accumulations, so every TC2_KC k-blocks the partial sum is handed to the epilogue warps (TMEM buffers
//             alternate) and added into fp32 REGISTERS with round-to-nearest; only the short in-chunk run accumulates
//             on the tensor core.
template <int BLOCK_N, int A_MODE, bool X3, int CG, bool RES2 = false>
__global__ void __launch_bounds__(Tc2Cfg<BLOCK_N, X3, CG, A_MODE >= A_ROWS, RES2>::THREADS, Tc2Cfg<BLOCK_N, X3, CG, A_MODE >= A_ROWS, RES2>::TWO_CTAS ? 2 : 1)
conv_tc2_kernel(const __grid_constant__ b2j_conv_tc_params p, const __grid_constant__ EpiPtrs epi,
                const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                const __grid_constant__ CUtensorMap tmap_b_lo, const __grid_constant__ CUtensorMap tmap_res,
                const __grid_constant__ Tc2Rows rows, const int has_res, const int epi_prog, float* __restrict__ out) {
  // has_res == 2: tmap_res is not the residual prefetch map but the OUTPUT map of the TMA-store epilogue
  constexpr bool ROWS = A_MODE >= A_ROWS;
  using Cfg = Tc2Cfg<BLOCK_N, X3, CG, ROWS, RES2>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  constexpr int PIPE_BYTES = Cfg::EPI_OFF;
  const uint32_t bar_base = smem_base + PIPE_BYTES + Cfg::EPI_BYTES + Cfg::OPND_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (Cfg::STAGES + s); };
  auto split_bar = [&](int s) { return bar_base + 8u * (2 * Cfg::STAGES + s); };
  auto bfull_bar = [&](int s) { return bar_base + 8u * (3 * Cfg::STAGES + s); };       // X3: the weight tiles of a stage
  auto tfull_bar = [&](int b) { return bar_base + 8u * (4 * Cfg::STAGES + b); };
  auto tempty_bar = [&](int b) { return bar_base + 8u * (4 * Cfg::STAGES + 2 + b); };
  // A_ROWS has no per-stage activation tile: full[0..NRB) signal "the tile's input rows have landed" for the row buffers,
  // full[NRB..2 NRB) hand them back (one arrival per gather warp)
  auto row_full = [&](uint32_t b) { return bar_base + 8u * b; };
  auto row_empty = [&](uint32_t b) { return bar_base + 8u * ((uint32_t)Cfg::NRB + b); };
  const uint32_t tmem_slot = bar_base + 8u * (4 * Cfg::STAGES + 4);
  volatile uint32_t* tmem_slot_gen =
      reinterpret_cast<volatile uint32_t*>(smem_gen + PIPE_BYTES + Cfg::EPI_BYTES + Cfg::OPND_BYTES + 8 * (4 * Cfg::STAGES + 4));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t M = p.batch * p.oh * p.ow;
  const uint32_t num_kb = p.kpad / TC_BLOCK_K;
  const uint32_t tiles_n = (p.o + BLOCK_N - 1) / BLOCK_N;
  const uint32_t tiles_m = (M + TC_BLOCK_M * CG - 1) / (TC_BLOCK_M * CG);
  const uint32_t num_tiles = ROWS ? p.batch * p.oh * rows.tiles_w : tiles_m * tiles_n;
  const uint32_t cta_rank = CG == 2 ? cluster_ctarank() : 0u;         // rank 0 of a pair leads: it arms the barriers and issues the MMAs
  const uint32_t first_tile = blockIdx.x / CG, tile_step = gridDim.x / CG;

  if (threadIdx.x == 0) {
    for (int s = 0; s < Cfg::STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
      mbar_init(split_bar(s), Cfg::SPLIT_ARRIVALS * CG);        // one arrival per splitter warp of the pair
      mbar_init(bfull_bar(s), 1);
    }
    if (ROWS) for (int b = 0; b < Cfg::NRB; ++b) mbar_init(row_empty(b), Cfg::SPLIT_WARPS);
    for (int b = 0; b < 2; ++b) { mbar_init(tfull_bar(b), 1); mbar_init(tempty_bar(b), Cfg::GROUP_WARPS * CG); }
    fence_barrier_init();
    prefetch_tmap(&tmap_a);
    prefetch_tmap(&tmap_b);
    if (X3) prefetch_tmap(&tmap_b_lo);
    if (has_res == 2) prefetch_tmap(&tmap_res);
  }
  if (ROWS) {
    // k -> source offset (bytes, relative to the pixel's window start in its staged box)
    uint32_t* offs = reinterpret_cast<uint32_t*>(smem_gen + Cfg::ROWTAB_OFF);
    for (uint32_t k = threadIdx.x; k < p.kpad; k += Cfg::THREADS) {
      uint32_t off = 0u;                   // K padding: any valid address, the gather zeroes the value
      if (k < rows.k) { const uint32_t kh = k / rows.kwc; off = (kh * rows.seg_w + (k - kh * rows.kwc)) * (A_MODE == A_ROWS_U8 ? 1u : 4u); }
      offs[k] = off;                       // byte offset from the pixel's window start
    }
    // uint8 value -> f32 through the fused input chain, every step rounded separately as the stand-alone kernels do
    if (A_MODE == A_ROWS_U8) {
      float* lut = reinterpret_cast<float*>(smem_gen + Cfg::ROWTAB_OFF + 1024);
      for (uint32_t b = threadIdx.x; b < 256; b += Cfg::THREADS) {
        float x = (float)b;
        for (uint32_t s_ = 0; s_ < rows.pre_n; ++s_) x = epi_op(rows.pre_op[s_], x, __uint_as_float(rows.pre_imm[s_]));
        if (!X3 && rows.round_in) x = __uint_as_float(cvt_tf32(__float_as_uint(x)));      // single pass: the operand rounded to nearest TF32
        lut[b] = x;
      }
    }
  }
  if (warp == 1) { if (CG == 2) tmem_alloc_2sm<Cfg::TMEM_COLS>(tmem_slot); else tmem_alloc<Cfg::TMEM_COLS>(tmem_slot); }
  tc_fence_before();
  if (CG == 2) cluster_sync_all(); else __syncthreads();      // pair: the peer's barriers must exist before anything signals them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_gen;

  // Programmatic dependent launch (B2J_PDL=1 adds the launch attribute, see launch_conv_tc2_inst): everything above --
  // barrier init, TMEM allocation, tensor-map prefetch -- touches no tensor and may run while the previous kernel of the
  // graph drains its last tiles; nothing below may start before that kernel has completed and flushed.  Both
  // instructions are no-ops for a kernel launched without the attribute / with no programmatic dependents.
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");

  if (ROWS && warp == 0) {
    // ======================================= TMA producer (A_ROWS) ==============================
    // the whole warp: per tile KH x nseg row boxes, one per lane, all crediting the row buffer's barrier; lane 0 also streams the
    // weight k-blocks through the stage ring
    const uint32_t es = A_MODE == A_ROWS_U8 ? 1u : 4u;
    const uint32_t box_bytes = p.kh * rows.seg_w * es;
    auto request_rows = [&](uint32_t t, uint32_t rt) {
      const uint32_t wseg = t % rows.tiles_w, line = t / rows.tiles_w;
      const uint32_t oh = line % p.oh, img = line / p.oh;
      // TMA wants the innermost coordinate on a 16-byte boundary (anything else is an illegal instruction: experiments/
      // tma3d_probe.cu): the staged boxes start at the aligned element at or below the tile's first source element
      const int x0 = rows_x0(p, rows, wseg) & ~(int)(rows.align - 1u);
      const int ih0 = (int)(oh * p.stride_h) - p.pad_h;
      const uint32_t rb = rt % (uint32_t)Cfg::NRB;
      if (lane == 0) {
        mbar_wait_sleepy(row_empty(rb), ((rt / (uint32_t)Cfg::NRB) & 1u) ^ 1u);
        mbar_expect_tx(row_full(rb), rows.nseg * box_bytes);
      }
      __syncwarp();
      const uint32_t dst0 = smem_base + Cfg::ROWBUF_OFF + rb * Cfg::ROWBUF_BYTES;
      if ((uint32_t)lane < rows.nseg)
        tma_load_tile_3d(dst0 + lane * box_bytes, &tmap_a, row_full(rb), x0 + (int)(lane * rows.seg_px * rows.pix_step), ih0, (int)img);
      __syncwarp();
    };
    if (lane == 0 && first_tile < num_tiles) {          // the resident weight matrix: slot kb holds k-block kb (hi, then lo for 3xTF32)
      mbar_expect_tx(bfull_bar(0), num_kb * (X3 ? 2 : 1) * Cfg::B_BYTES);
      for (uint32_t kb = 0; kb < num_kb; ++kb) {
        const uint32_t b_dst = smem_base + kb * Cfg::STAGE_BYTES;
        tma_load_2d(b_dst, &tmap_b, bfull_bar(0), (int)(kb * TC_BLOCK_K), 0);
        if (X3) tma_load_2d(b_dst + Cfg::B_BYTES, &tmap_b_lo, bfull_bar(0), (int)(kb * TC_BLOCK_K), 0);
      }
    }
    __syncwarp();
    uint32_t t_req = first_tile, rt_req = 0;
    for (; t_req < num_tiles; t_req += tile_step, ++rt_req) request_rows(t_req, rt_req);      // blocks on row_empty: NRB tiles ahead at most
  } else if (warp == 0) {
    // ======================================= TMA producer =======================================
    if (B2J_WARP_TMA || lane == 0) {
      const bool leader = B2J_WARP_TMA ? elect_one() : true;      // B2J_WARP_TMA: whole warp in the loop, one lane issues (as in the MMA role)
      // stage / phase and the filter-tap counters are carried incrementally: this single thread sits on the
      // empty -> TMA -> full -> (split) -> MMA chain, and the divisions (it % STAGES, kb / cblocks, tap / kw) of the first
      // version cost it several hundred cycles per k-block (ncu source view of the 64-wide 3xTF32 kernel, round 2)
      uint32_t st = 0, ph = 0;
      const uint32_t cblocks = A_MODE == A_IM2COL ? p.c / TC_BLOCK_K : 1;
      for (uint32_t t = first_tile; t < num_tiles; t += tile_step) {
        const uint32_t m0 = (t / tiles_n) * (TC_BLOCK_M * CG) + cta_rank * TC_BLOCK_M, n0 = (t % tiles_n) * BLOCK_N;
        const uint32_t nb0 = n0 + cta_rank * Cfg::B_ROWS;           // this CTA's slice of the weight tile
        int bw = 0, bh = 0, bn = 0;
        if (A_MODE == A_IM2COL) {
          const uint32_t ow = m0 % p.ow, t1 = m0 / p.ow;
          bw = (int)(ow * p.stride_w) - p.pad_w;
          bh = (int)((t1 % p.oh) * p.stride_h) - p.pad_h;
          bn = (int)(t1 / p.oh);
        }
        uint32_t kh = 0, kw = 0, cb = 0;
        for (uint32_t kb = 0; kb < num_kb; ++kb) {
          const int s = (int)st;
          mbar_wait_sleepy(empty_bar(s), ph ^ 1u);
          if (++st == (uint32_t)Cfg::STAGES) { st = 0; ph ^= 1u; }
          const uint32_t a_dst = smem_base + s * Cfg::STAGE_BYTES;
          const uint32_t b_dst = a_dst + Cfg::A_SMEM_BYTES;
          // single pass: everything of the stage is credited to the leader's full barrier.  3xTF32: the activation tile
          // goes to THIS CTA's full barrier (its splitter warps wait for it), the weight tiles to the leader's bfull.
          if (leader) {
          if (X3) { mbar_expect_tx(full_bar(s), Cfg::A_BYTES); if (cta_rank == 0) mbar_expect_tx(bfull_bar(s), CG * 2 * Cfg::B_BYTES); }
          else if (cta_rank == 0) mbar_expect_tx(full_bar(s), CG * (Cfg::A_BYTES + Cfg::B_BYTES));
          }
          if (A_MODE == A_IM2COL) {
            if (leader) {
            if (CG == 2 && !X3) tma_load_im2col_4d_2sm(a_dst, &tmap_a, full_bar(s), (int)(cb * TC_BLOCK_K), bw, bh, bn,
                                                (uint16_t)(kw * p.dil_w), (uint16_t)(kh * p.dil_h));
            else tma_load_im2col_4d(a_dst, &tmap_a, full_bar(s), (int)(cb * TC_BLOCK_K), bw, bh, bn,
                                    (uint16_t)(kw * p.dil_w), (uint16_t)(kh * p.dil_h));
            }
            if (++cb == cblocks) { cb = 0; if (++kw == p.kw) { kw = 0; ++kh; } }
          } else if (leader) {
            if (CG == 2 && !X3) tma_load_2d_2sm(a_dst, &tmap_a, full_bar(s), (int)(kb * TC_BLOCK_K), (int)m0);
            else tma_load_2d(a_dst, &tmap_a, full_bar(s), (int)(kb * TC_BLOCK_K), (int)m0);
          }
          const uint32_t b_bar = X3 ? bfull_bar(s) : full_bar(s);
          if (leader) {
          if (CG == 2) tma_load_2d_2sm(b_dst, &tmap_b, b_bar, (int)(kb * TC_BLOCK_K), (int)nb0);
          else tma_load_2d(b_dst, &tmap_b, b_bar, (int)(kb * TC_BLOCK_K), (int)nb0);
          if (X3) {
            if (CG == 2) tma_load_2d_2sm(b_dst + Cfg::B_BYTES, &tmap_b_lo, b_bar, (int)(kb * TC_BLOCK_K), (int)nb0);
            else tma_load_2d(b_dst + Cfg::B_BYTES, &tmap_b_lo, b_bar, (int)(kb * TC_BLOCK_K), (int)nb0);
          }
          }
          if (B2J_WARP_TMA) __syncwarp();
        }
      }
    }
  } else if (warp == 1) {
    // ======================================= MMA issuer =========================================
    // The WHOLE warp runs the loop and one elected lane issues (B2J_WARP_MMA, default): in warp-uniform control flow the compiler
    // keeps the TMEM addresses and shared-memory descriptors in uniform registers and the UTCHMMAs go out back to back.  Issued by
    // a single thread inside a divergent branch (rounds 1 - 2) every tcgen05.mma was wrapped in an ELECT / R2UR.BROADCAST /
    // BRA.U.ANY sequence costing ~90 cycles: the MMA thread of the 3xTF32 kernels spent 75 - 88 % of its time ISSUING (12 MMAs per
    // k-block; scripts/diag_mma_waits.py), which -- not the tensor pipe, not operand delivery -- was the "narrow-tile ceiling" of
    // the N <= 128 layers (experiments/umma_rate_probe.cu: the pipe itself sustains one 128 x 64 x 8 TF32 MMA per 32 cycles).
    if (cta_rank == 0 && (B2J_WARP_MMA || lane == 0)) {
      const bool leader = B2J_WARP_MMA ? elect_one() : true;
      constexpr uint32_t idesc = make_idesc_tf32(TC_BLOCK_M * CG, BLOCK_N);
      uint32_t st = 0, ph = 0, chunk = 0;     // chunk: global count of MMA -> epilogue handoffs; TMEM buffer = chunk & 1
      bool it_first = true;
#ifdef B2J_DIAG_MMA_WAITS           // developer diagnostic: where does the MMA thread wait?  (printed by CTA 0)
      long long w_acc = 0, w_opnd = 0, w_b = 0; const long long t_begin = clock64();
#define DIAG_T0 const long long dt0_ = clock64();
#define DIAG_ADD(x) x += clock64() - dt0_;
#else
#define DIAG_T0
#define DIAG_ADD(x)
#endif
      for (uint32_t t = first_tile; t < num_tiles; t += tile_step) {
        // the residual tile this output tile will add in its epilogue: pull it into L2 when the tile's MMAs start, one
        // tile ahead of the epilogue (from the TMA producer, 2-3 tiles ahead, 40 % of it was evicted again before use)
        if (has_res == 1 && leader) tma_prefetch_l2_2d(&tmap_res, (int)((t % tiles_n) * BLOCK_N), (int)((t / tiles_n) * (TC_BLOCK_M * CG)));
        const uint32_t kstep = X3 ? (uint32_t)Cfg::KC : num_kb;          // k-blocks per MMA -> epilogue handoff
        for (uint32_t kb0 = 0; kb0 < num_kb; kb0 += kstep, ++chunk) {
          const uint32_t ab = chunk & 1u;
          { DIAG_T0 mbar_wait_sleepy(tempty_bar(ab), ((chunk >> 1) & 1u) ^ 1u); DIAG_ADD(w_acc) }     // epilogue has drained this accumulator
          tc_fence_after();
          const uint32_t tmem_d = tmem_base + ab * BLOCK_N;
          const uint32_t kb1 = kb0 + kstep > num_kb ? num_kb : kb0 + kstep;
          for (uint32_t kb = kb0; kb < kb1; ++kb) {
            const int s = (int)st;
            { DIAG_T0 mbar_wait_sleepy((X3 || ROWS) ? split_bar(s) : full_bar(s), ph); DIAG_ADD(w_opnd) }
            if (ROWS) { if (it_first) { mbar_wait_sleepy(bfull_bar(0), 0u); it_first = false; } }       // resident weights: loaded once
            else if (X3) { DIAG_T0 mbar_wait_sleepy(bfull_bar(s), ph); DIAG_ADD(w_b) }
            if (++st == (uint32_t)Cfg::STAGES) { st = 0; ph ^= 1u; }
            tc_fence_after();
            const uint32_t stage = smem_base + (ROWS ? kb : (uint32_t)s) * Cfg::STAGE_BYTES;      // A_ROWS: resident weight slot of this k-block
            if (leader) {
            if (X3) {
              // activation operands from TMEM (written by the splitter warps), weights from shared memory
              const uint32_t a_hi = tmem_base + (uint32_t)(Cfg::A_TMEM_COL0 + s * Cfg::A_TMEM_COLS), a_lo = a_hi + TC_BLOCK_K;
              const uint64_t b_hi = make_smem_desc(stage + Cfg::A_SMEM_BYTES), b_lo = make_smem_desc(stage + Cfg::A_SMEM_BYTES + Cfg::B_BYTES);
#pragma unroll
              for (int k = 0; k < TC_BLOCK_K / 8; ++k) {
                const uint64_t adv = (uint64_t)(k * 2);
                const uint32_t acol = (uint32_t)(k * 8);
                if (CG == 2) {
                  umma_tf32_ts_2sm(tmem_d, a_lo + acol, b_hi + adv, idesc, (kb != kb0 || k != 0) ? 1u : 0u);
                  umma_tf32_ts_2sm(tmem_d, a_hi + acol, b_lo + adv, idesc, 1u);
                  umma_tf32_ts_2sm(tmem_d, a_hi + acol, b_hi + adv, idesc, 1u);
                } else {
                  umma_tf32_ts(tmem_d, a_lo + acol, b_hi + adv, idesc, (kb != kb0 || k != 0) ? 1u : 0u);
                  umma_tf32_ts(tmem_d, a_hi + acol, b_lo + adv, idesc, 1u);
                  umma_tf32_ts(tmem_d, a_hi + acol, b_hi + adv, idesc, 1u);
                }
              }
            } else if (ROWS) {
              // single pass, operand assembled in TMEM by the gather warps
              const uint32_t a_t = tmem_base + (uint32_t)(Cfg::A_TMEM_COL0 + s * Cfg::A_TMEM_COLS);
              const uint64_t bdesc = make_smem_desc(stage);
#pragma unroll
              for (int k = 0; k < TC_BLOCK_K / 8; ++k)
                umma_tf32_ts(tmem_d, a_t + (uint32_t)(k * 8), bdesc + (uint64_t)(k * 2), idesc, (kb | (uint32_t)k) != 0u);
            } else {
              const uint64_t adesc = make_smem_desc(stage), bdesc = make_smem_desc(stage + Cfg::A_BYTES);
#pragma unroll
              for (int k = 0; k < TC_BLOCK_K / 8; ++k)
                if (CG == 2) umma_tf32_2sm(tmem_d, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (kb | (uint32_t)k) != 0u);
                else umma_tf32(tmem_d, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (kb | (uint32_t)k) != 0u);
            }
            if (CG == 2) umma_commit_2sm(empty_bar(s)); else umma_commit(empty_bar(s));
            }
            if (B2J_WARP_MMA) __syncwarp();
          }
          if (leader) { if (CG == 2) umma_commit_2sm(tfull_bar(ab)); else umma_commit(tfull_bar(ab)); }
          if (B2J_WARP_MMA) __syncwarp();
        }
      }
#ifdef B2J_DIAG_MMA_WAITS
      if (blockIdx.x == 0 && leader) printf("[mma waits] N=%d mode=%d x3=%d cg=%d o=%u kpad=%u: total %lld cycles, accumulator %lld, operand(A / split) %lld, weights %lld\n",
                                  BLOCK_N, A_MODE, (int)X3, CG, p.o, p.kpad, clock64() - t_begin, w_acc, w_opnd, w_b);
#endif
    }
  } else if (ROWS && warp < 2 + Cfg::SPLIT_WARPS) {
    // ======================================= gather warps (A_ROWS) ===============================
    // a thread per output pixel of the tile (TMEM lane = tile row): 32 K-elements per k-block from the staged rows -> TMEM
    constexpr bool U8 = A_MODE == A_ROWS_U8;
    constexpr uint32_t ES = U8 ? 1u : 4u;
    const int q = warp & 3;
    const uint32_t a_t0 = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)Cfg::A_TMEM_COL0;
    const uint32_t* offs = reinterpret_cast<const uint32_t*>(smem_gen + Cfg::ROWTAB_OFF);
    const uint32_t lut_s = smem_base + Cfg::ROWTAB_OFF + 1024;
    // K padding (k >= KH*KW*C) exists in the last k-block only: its table entries are 0 (a valid address) and the values are
    // zeroed after the load
    const uint32_t k_last = rows.k - (num_kb - 1) * TC_BLOCK_K;      // real K-elements of the last k-block (1 .. 32)
    // This warp's set takes every other k-block of the CTA's k-block sequence: global index it = set, set + 2, ...; STAGES is even,
    // so its TMEM slot it % STAGES advances by 2, and the first own k-block of the next tile follows from where this one stopped.
    // Everything per tile is carried incrementally and the pixel's window offset is hoisted when a tile is a whole output row
    // (ncu source view, end of round 2: 17 % of the gather warps' samples sat in the per-tile divisions, another 15 % in the
    // skipped iterations of the other set's k-blocks).
    static_assert(!ROWS || Cfg::STAGES % 2 == 0, "the two gather sets alternate TMEM slots");
    const uint32_t set = (uint32_t)(warp - 2) >> 2;
    uint32_t st = set, ph = 0, rb = 0, rb_ph = 0, kb_first = set;
    // this pixel's window start inside a row buffer (bytes): the boxes begin at the 16-byte aligned element at or below their
    // first source element; rows beyond the tile's pixels are computed from the last real pixel's window and dropped by the epilogue
    auto window = [&](uint32_t wseg) -> uint32_t {
      const uint32_t valid = p.ow - wseg * TC_BLOCK_M;                 // pixels of this tile that exist (>= 128: all)
      uint32_t row = (uint32_t)(q * 32 + lane);
      row = row < valid ? row : valid - 1;
      const uint32_t seg = row / rows.seg_px;                         // the staged box this pixel reads from
      return (seg * p.kh * rows.seg_w + (row - seg * rows.seg_px) * rows.pix_step +
              (uint32_t)(rows_x0(p, rows, wseg) & (int)(rows.align - 1u))) * ES;
    };
    const bool row_tiles = rows.tiles_w == 1;                          // one tile per output row: the offset is the same for every tile
    const uint32_t win_row = window(0);
    for (uint32_t t = first_tile; t < num_tiles; t += tile_step) {
      const uint32_t win = smem_base + Cfg::ROWBUF_OFF + rb * Cfg::ROWBUF_BYTES + (row_tiles ? win_row : window(t % rows.tiles_w));
      mbar_wait(row_full(rb), rb_ph);
      uint32_t kb = kb_first;
      for (; kb < num_kb; kb += 2) {
        const int s = (int)st;
        const uint32_t ph_s = ph;
        st += 2u;
        if (st >= (uint32_t)Cfg::STAGES) { st -= (uint32_t)Cfg::STAGES; ph ^= 1u; }
        mbar_wait(empty_bar(s), ph_s ^ 1u);                          // the MMAs that read this TMEM slot have completed
        tc_fence_after();
        uint32_t v[32];
#ifdef B2J_DIAG_ROWS_NO_GATHER       // timing diagnostic only (results are garbage): no table / source loads, the operand is a constant
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = 0x3f800000u + (uint32_t)lane;
#else
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const uint4 o = *reinterpret_cast<const uint4*>(offs + kb * TC_BLOCK_K + 4 * c);   // same address in every lane: broadcast
          const uint32_t oo[4] = {o.x, o.y, o.z, o.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            uint32_t x;
            if (U8) {
              uint32_t b;
              asm volatile("ld.shared.u8 %0, [%1];" : "=r"(b) : "r"(win + oo[e]));
              asm volatile("ld.shared.u32 %0, [%1];" : "=r"(x) : "r"(lut_s + 4u * b));
            } else {
              asm volatile("ld.shared.u32 %0, [%1];" : "=r"(x) : "r"(win + oo[e]));
            }
            v[4 * c + e] = x;
          }
        }
        if (kb + 1 == num_kb && k_last < (uint32_t)TC_BLOCK_K) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = (uint32_t)j < k_last ? v[j] : 0u;
        }
#endif
        if (X3) {
          uint32_t h[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) h[j] = rna_tf32_bits(v[j]);
          tmem_st32(a_t0 + (uint32_t)(s * Cfg::A_TMEM_COLS), h);
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) - __uint_as_float(h[j]));
          tmem_st32(a_t0 + (uint32_t)(s * Cfg::A_TMEM_COLS + TC_BLOCK_K), v);
        } else {
          // Round to nearest TF32 in ONE integer add: the tensor core ignores the low 13 mantissa bits, so adding half a TF32
          // ulp to the bit pattern makes its truncation round to nearest, ties away from zero -- cvt.rna.tf32.f32, which
          // compiles to four instructions per element and dominated this loop.  (Inf stays Inf, quiet NaNs stay NaN.)  The
          // uint8 table is rounded once when it is built.
          if (!U8 && rows.round_in) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] += 0x1000u;
          }
          tmem_st32(a_t0 + (uint32_t)(s * Cfg::A_TMEM_COLS), v);
        }
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(split_bar(s));
      }
      kb_first = kb - num_kb;
      __syncwarp();
      if (lane == 0) mbar_arrive(row_empty(rb));                     // every lane's reads of the row buffer are done
      if (++rb == (uint32_t)Cfg::NRB) { rb = 0; rb_ph ^= 1u; }
    }
  } else if (X3 && warp < 2 + Cfg::SPLIT_WARPS) {
    // ======================================= splitters (3xTF32) ==================================
    // Each of the 4 warps owns the TMEM lane quarter it may access (warp id % 4); a thread converts ONE row of the landed
    // 128 x 32 activation tile: 8 swizzled 16-byte reads, hi = rna_tf32(a), lo = a - hi, two 32-column TMEM stores.
    // Nothing is written back to shared memory (round 1 rewrote the stage in place: 48 KB of shared-memory traffic per
    // k-block next to the 72 KB the three MMAs read, with the tensor pipe 41 % busy), and the MMAs no longer read A from it.
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const uint32_t a_t0 = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)Cfg::A_TMEM_COL0;
    uint32_t st = 0, ph = 0;
    for (uint32_t t = first_tile; t < num_tiles; t += tile_step) {
      for (uint32_t kb = 0; kb < num_kb; ++kb) {
        const int s = (int)st;
        mbar_wait(full_bar(s), ph);
        if (++st == (uint32_t)Cfg::STAGES) { st = 0; ph ^= 1u; }
#ifdef B2J_DIAG_NO_SPLIT             // timing diagnostic only (results are garbage): the splitters signal without converting
        tc_fence_before();
        __syncwarp();
        if (lane == 0) { if (CG == 2) mbar_arrive_cluster(split_bar(s), 0); else mbar_arrive(split_bar(s)); }
        continue;
#endif
        const uint8_t* a_row = smem_gen + s * Cfg::STAGE_BYTES + row * 128;
        uint32_t v[32], h[32];
#pragma unroll
        for (int c = 0; c < 8; ++c) {            // SWIZZLE_128B: 16-byte chunk c of row r sits at chunk c ^ (r & 7)
          const uint4 x = *reinterpret_cast<const uint4*>(a_row + ((c ^ (row & 7)) << 4));
          v[4 * c] = x.x; v[4 * c + 1] = x.y; v[4 * c + 2] = x.z; v[4 * c + 3] = x.w;
        }
#pragma unroll
        for (int j = 0; j < 32; ++j) h[j] = rna_tf32_bits(v[j]);
        tmem_st32(a_t0 + (uint32_t)(s * Cfg::A_TMEM_COLS), h);
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) - __uint_as_float(h[j]));
        tmem_st32(a_t0 + (uint32_t)(s * Cfg::A_TMEM_COLS + TC_BLOCK_K), v);
        tmem_st_wait();
        tc_fence_before();                      // the TMEM stores are complete before the arrival below is observed
        __syncwarp();
        if (lane == 0) { if (CG == 2) mbar_arrive_cluster(split_bar(s), 0); else mbar_arrive(split_bar(s)); }
      }
    }
  } else {
    // ======================================= epilogue ===========================================
    // one straight-line instantiation per epilogue program: registers are allocated per program, and only the
    // selected one ever enters the instruction cache
    Tc2EpiCtx cx;
    cx.smem_gen = smem_gen; cx.bar_base = bar_base; cx.tmem_base = tmem_base; cx.out = out;
    cx.M = M; cx.num_kb = num_kb; cx.tiles_n = tiles_n; cx.num_tiles = num_tiles;
    cx.first_tile = first_tile; cx.tile_step = tile_step; cx.cta_rank = cta_rank; cx.rows_tiles_w = ROWS ? rows.tiles_w : 1u;
    cx.tmap_out = has_res == 2 ? &tmap_res : nullptr;
    if constexpr (RES2) {          // launched for BN + residual + ReLU only (launch_conv_tc2)
      tc2_epilogue_role<BLOCK_N, X3, CG, ROWS, EPROG_BN_ADD_RELU, true>(p, epi, cx);
    } else
    switch (epi_prog) {
      case EPROG_BN:          tc2_epilogue_role<BLOCK_N, X3, CG, ROWS, EPROG_BN>(p, epi, cx); break;
      case EPROG_BN_RELU:     tc2_epilogue_role<BLOCK_N, X3, CG, ROWS, EPROG_BN_RELU>(p, epi, cx); break;
      case EPROG_BN_ADD_RELU: tc2_epilogue_role<BLOCK_N, X3, CG, ROWS, EPROG_BN_ADD_RELU>(p, epi, cx); break;
      case EPROG_BIAS:        tc2_epilogue_role<BLOCK_N, X3, CG, ROWS, EPROG_BIAS>(p, epi, cx); break;
      case EPROG_BIAS_RELU:   tc2_epilogue_role<BLOCK_N, X3, CG, ROWS, EPROG_BIAS_RELU>(p, epi, cx); break;
      default:                tc2_epilogue_role<BLOCK_N, X3, CG, ROWS, EPROG_GENERIC>(p, epi, cx); break;
    }
  }

  tc_fence_before();
  if (CG == 2) cluster_sync_all(); else __syncthreads();      // pair: no CTA may leave while its peer can still signal it
  if (warp == 1) {
    tc_fence_after();
    if (CG == 2) tmem_dealloc_2sm<Cfg::TMEM_COLS>(tmem_base); else tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
  }
}

// ---- host side: tensor maps + launch -------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
typedef CUresult (*PFN_encodeIm2col)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                     const int*, const int*, cuuint32_t, cuuint32_t, const cuuint32_t*, CUtensorMapInterleave,
                                     CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

struct TmaApi {
  PFN_encodeTiled tiled = nullptr;
  PFN_encodeIm2col im2col = nullptr;
  bool tried = false;
};
static TmaApi g_tma;

static bool tma_api_load() {
  if (g_tma.tried) return g_tma.tiled && g_tma.im2col;
  g_tma.tried = true;
  cudaDriverEntryPointQueryResult q;
  void* f = nullptr;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
    g_tma.tiled = (PFN_encodeTiled)f;
  f = nullptr;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &f, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
    g_tma.im2col = (PFN_encodeIm2col)f;
  return g_tma.tiled && g_tma.im2col;
}

static bool make_tmap_2d(CUtensorMap* map, const float* base, uint64_t inner, uint64_t outer, uint64_t row_pitch_elems, uint32_t box_inner,
                         uint32_t box_outer) {
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {row_pitch_elems * 4};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  return g_tma.tiled(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

static bool make_tmap_plain(CUtensorMap* map, const float* base, uint64_t inner, uint64_t outer, uint32_t box_inner, uint32_t box_outer) {
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {inner * 4};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  return g_tma.tiled(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

static bool make_tmap_im2col(CUtensorMap* map, const float* x, const b2j_conv_tc_params& p) {
  cuuint64_t dims[4] = {p.c, p.w, p.h, p.batch};
  cuuint64_t strides[3] = {(cuuint64_t)p.c * 4, (cuuint64_t)p.w * p.c * 4, (cuuint64_t)p.h * p.w * p.c * 4};
  // high padding implied by the output size (the reference passes low padding only: conv2d.comp PADDING)
  const int pad_w_hi = (int)((p.ow - 1) * p.stride_w + (p.kw - 1) * p.dil_w + 1) - (int)p.w - p.pad_w;
  const int pad_h_hi = (int)((p.oh - 1) * p.stride_h + (p.kh - 1) * p.dil_h + 1) - (int)p.h - p.pad_h;
  int lower[2] = {-p.pad_w, -p.pad_h};
  int upper[2] = {pad_w_hi - (int)((p.kw - 1) * p.dil_w), pad_h_hi - (int)((p.kh - 1) * p.dil_h)};
  cuuint32_t estr[4] = {1, p.stride_w, p.stride_h, 1};
  return g_tma.im2col(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void*)x, dims, strides, lower, upper, /*channelsPerPixel*/ TC_BLOCK_K,
                      /*pixelsPerColumn*/ TC_BLOCK_M, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                      CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// Output view for the TMA-store epilogue: [lines][pixels][channels] (a linear [M, O] output is one line of M pixels), boxes of
// 32 channels x 32 pixels, 128-byte swizzle.  B2J_TMA_STORE=0 keeps the st.global epilogue everywhere.
static bool make_tmap_out(CUtensorMap* map, float* out, const b2j_conv_tc_params& p, bool rows_view) {
  static int enabled = -1;
  if (enabled < 0) { const char* e = getenv("B2J_TMA_STORE"); enabled = e ? atoi(e) : 1; }
  if (!enabled || (p.o & 3u) != 0) return false;
  const uint64_t M = (uint64_t)p.batch * p.oh * p.ow;
  cuuint64_t dims[3] = {p.o, rows_view ? (cuuint64_t)p.ow : (cuuint64_t)M, rows_view ? (cuuint64_t)p.batch * p.oh : 1ull};
  cuuint64_t strides[2] = {(cuuint64_t)p.o * 4, (rows_view ? (cuuint64_t)p.ow : (cuuint64_t)M) * p.o * 4};
  cuuint32_t box[3] = {32, 32, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  return g_tma.tiled(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)out, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
static bool tma_store_program(int prog) { return prog == EPROG_BN || prog == EPROG_BN_RELU || prog == EPROG_BIAS || prog == EPROG_BIAS_RELU; }

template <int BLOCK_N, int A_MODE, bool X3, int CG, bool RES2 = false>
static int launch_conv_tc2_inst(const b2j_conv_tc_params& p, const EpiPtrs& epi, const CUtensorMap& ta, const CUtensorMap& tb,
                                const CUtensorMap& tbl, const CUtensorMap& tr, int has_res, int prog, float* out, int sm_count,
                                cudaStream_t st, const char** why, const Tc2Rows& rows = Tc2Rows{}) {
  using Cfg = Tc2Cfg<BLOCK_N, X3, CG, A_MODE >= A_ROWS, RES2>;
  static bool configured = false;
  auto kern = conv_tc2_kernel<BLOCK_N, A_MODE, X3, CG, RES2>;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
    if (e != cudaSuccess) { *why = cudaGetErrorString(e); return B2J_ECUDA; }
    configured = true;
  }
  const uint32_t M = p.batch * p.oh * p.ow;
  const uint32_t tiles = A_MODE >= A_ROWS ? p.batch * p.oh * rows.tiles_w
                                          : ((M + TC_BLOCK_M * CG - 1) / (TC_BLOCK_M * CG)) * ((p.o + BLOCK_N - 1) / BLOCK_N);
  const uint32_t slots = (uint32_t)sm_count / CG * (Cfg::TWO_CTAS ? 2 : 1);   // persistent: one CTA (pair) per SM (pair)
  const unsigned grid = (tiles < slots ? tiles : slots) * CG;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(Cfg::THREADS);
  cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  // B2J_PDL=1: programmatic stream serialization -- this kernel's CTAs may become resident (and run their set-up) as soon
  // as every CTA of the preceding kernel has passed its griddepcontrol.launch_dependents or exited; captured into the
  // CUDA graph as a programmatic edge.  The kernel's griddepcontrol.wait orders all of its memory traffic after the
  // predecessor's completion.
  static int pdl = -1;
  if (pdl < 0) { const char* e = getenv("B2J_PDL"); pdl = e ? atoi(e) : B2J_PDL_DEFAULT; }
  if (pdl) {
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.numAttrs = 2;
  }
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, p, epi, ta, tb, tbl, tr, rows, has_res, prog, out);
  if (e != cudaSuccess) { *why = cudaGetErrorString(e); return B2J_ECUDA; }
  return B2J_OK;
}

// Tile shape for the single-pass kernel, from per-layer measurements on B200 (ResNet-50 b256, profiles/r01_tile_shapes.md):
//   * 256 x 256 CTA-pair tiles win wherever N >= 256 and the main loop carries the time (K >= 512, or no residual);
//   * layers that are mostly epilogue (residual add and K <= 256) are faster with 256 x 128 pair tiles;
//   * N = 128 layers gain ~2 % from pairing, N <= 64 stays single-CTA;
//   * small problems keep the shape that still gives every SM a tile.
static void choose_tc2_tile(uint32_t M, uint32_t N, uint32_t K, bool residual, int sm_count, int* bn, int* cg) {
  static int force_cg = -1;
  if (force_cg < 0) { const char* e = getenv("B2J_TC2_CG"); force_cg = e ? atoi(e) : 0; }
  auto tiles = [&](int bn_, int cg_) { return (uint64_t)((M + 128 * cg_ - 1) / (128 * cg_)) * ((N + bn_ - 1) / bn_); };
  *bn = N <= 64 ? 64 : 128;
  *cg = 1;
  if (force_cg == 4 && N <= 64) { *cg = 2; return; }                   // experiments: pair the 64-wide tiles too
  if (force_cg == 1 || N < 128) return;
  if (force_cg == 2) { *cg = 2; return; }                              // experiments: 128-wide pairs everywhere
  if (force_cg == 3) { *cg = 2; *bn = N >= 256 ? 256 : 128; return; }  // experiments: widest pairs everywhere
  const uint64_t pairs = (uint64_t)sm_count / 2;
  // wave quantisation: with few tiles the last wave of 256 x 256 tiles leaves most pairs idle (ResNet-50 stage-3 3x3: 98 tiles on 74
  // pairs = 2 waves at 0.66 occupancy; 196 tiles of 256 x 128 = 3 waves at 0.88).  Re-measured after the MMA-issue fix: 0.099 -> 0.094 ms.
  auto wave_eff = [&](int bn_) { const uint64_t t = tiles(bn_, 2); return (double)t / (double)(((t + pairs - 1) / pairs) * pairs); };
  if (N >= 256 && !(residual && K <= 256) && tiles(256, 2) >= pairs && wave_eff(128) <= 1.25 * wave_eff(256)) { *bn = 256; *cg = 2; return; }
  if (tiles(128, 2) >= pairs) { *bn = 128; *cg = 2; return; }
}

// A_ROWS launch (B2J_CT_ROWS): x is the raw NHWC input, f32 or packed uint8.  The Python planner (plan_contraction) applies the
// same admission rule -- rows_geometry below is its C twin -- and otherwise re-lays the input out for the im2col path.
static bool rows_geometry(const b2j_conv_tc_params& p, Tc2Rows* g) {
  const uint32_t es = p.src_u8 ? 1u : 4u;
  const uint32_t tile_px = p.ow < (uint32_t)TC_BLOCK_M ? p.ow : (uint32_t)TC_BLOCK_M;
  g->align = 16u / es;
  g->seg_w = 256;
  g->pix_step = p.stride_w * p.c;
  g->kwc = p.kw * p.c;
  // pixels per 256-element box: the last pixel's window and the alignment slack must fit, and consecutive boxes must start
  // a multiple of 16 bytes apart (one alignment offset serves all of them)
  uint32_t sp = 0;
  if (g->kwc + g->align - 1 <= g->seg_w) {
    sp = (g->seg_w - g->kwc - (g->align - 1)) / g->pix_step + 1;
    while (sp > 0 && (sp * g->pix_step) % g->align != 0) --sp;
  }
  g->seg_px = sp;
  g->nseg = sp ? (tile_px + sp - 1) / sp : 0;
  g->k = p.kh * p.kw * p.c;
  g->tiles_w = (p.ow + TC_BLOCK_M - 1) / TC_BLOCK_M;
  g->src_u8 = p.src_u8 ? 1u : 0u;
  g->round_in = (p.flags & B2J_CT_ROUND_IN_TF32) ? 1u : 0u;
  g->pre_n = p.pre_n;
  for (int i = 0; i < 2; ++i) { g->pre_op[i] = p.pre_op[i]; g->pre_imm[i] = p.pre_imm[i]; }
  const uint32_t res_slots = p.precision == B2J_PREC_TF32X3 ? (uint32_t)Tc2Cfg<64, true, 1, true>::RES_SLOTS : (uint32_t)Tc2Cfg<64, false, 1, true>::RES_SLOTS;
  return sp > 0 && g->nseg <= 32 && p.kpad / TC_BLOCK_K <= res_slots && p.o <= 64 && p.dil_w == 1 && p.dil_h == 1 && p.kh <= 256 && p.kpad <= 256 && p.kpad >= g->k &&
         p.kpad % TC_BLOCK_K == 0 && p.pad_h >= 0 && p.pad_w >= 0 && ((uint64_t)p.w * p.c * es) % 16 == 0 &&
         (uint64_t)g->nseg * p.kh * g->seg_w * es <= (uint64_t)Tc2Cfg<64, false, 1, true>::ROWBUF_BYTES &&
         (p.src_u8 || p.pre_n == 0) && p.pre_n <= 2;
}

static int launch_conv_rows(const b2j_conv_tc_params& p, const EpiPtrs& epi, float* out, const void* x, const float* wt,
                            const float* wt_lo, int sm_count, cudaStream_t st, const char** why) {
  const bool x3 = p.precision == B2J_PREC_TF32X3;
  if (x3 && !wt_lo) { *why = "3xTF32 needs the wt_lo buffer"; return B2J_EINVAL; }
  Tc2Rows g;
  if (!rows_geometry(p, &g)) { *why = "B2J_CT_ROWS: geometry not supported (O <= 64, no filter dilation, Kpad <= 256, 16-byte rows, staged boxes within 32 KB)"; return B2J_EINVAL; }
  if (!tma_api_load()) { *why = "cuTensorMapEncode* not available"; return B2J_ENOTIMPL; }
  const uint32_t es = p.src_u8 ? 1u : 4u;
  CUtensorMap ta, tb, tbl;
  {
    cuuint64_t dims[3] = {(cuuint64_t)p.w * p.c, p.h, p.batch};
    cuuint64_t strides[2] = {(cuuint64_t)p.w * p.c * es, (cuuint64_t)p.h * p.w * p.c * es};
    cuuint32_t box[3] = {g.seg_w, p.kh, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    if (g_tma.tiled(&ta, p.src_u8 ? CU_TENSOR_MAP_DATA_TYPE_UINT8 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<void*>(x), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) { *why = "input-row tensor map"; return B2J_ENOTIMPL; }
  }
  if (!make_tmap_2d(&tb, wt, p.kpad, p.o, p.kpad, TC_BLOCK_K, 64)) { *why = "weight tensor map"; return B2J_ENOTIMPL; }
  tbl = tb;
  if (x3 && !make_tmap_2d(&tbl, wt_lo, p.kpad, p.o, p.kpad, TC_BLOCK_K, 64)) { *why = "weight (lo) tensor map"; return B2J_ENOTIMPL; }
  const int prog = classify_epilogue(p.epi);
  CUtensorMap tr = tb;
  const int hr = (tma_store_program(prog) && make_tmap_out(&tr, out, p, true)) ? 2 : 0;      // 2: tr is the TMA-store output map
  if (p.src_u8) {
    if (x3) return launch_conv_tc2_inst<64, A_ROWS_U8, true, 1>(p, epi, ta, tb, tbl, tr, hr, prog, out, sm_count, st, why, g);
    return launch_conv_tc2_inst<64, A_ROWS_U8, false, 1>(p, epi, ta, tb, tbl, tr, hr, prog, out, sm_count, st, why, g);
  }
  if (x3) return launch_conv_tc2_inst<64, A_ROWS, true, 1>(p, epi, ta, tb, tbl, tr, hr, prog, out, sm_count, st, why, g);
  return launch_conv_tc2_inst<64, A_ROWS, false, 1>(p, epi, ta, tb, tbl, tr, hr, prog, out, sm_count, st, why, g);
}

// Returns B2J_ENOTIMPL (why set) when the problem cannot be expressed with TMA tensor maps (the Python planner re-lays
// such activations out first, see plan_relayout in vkjax_b200/interpreter.py).
static int launch_conv_tc2(const b2j_conv_tc_params& p, const EpiPtrs& epi, float* out, const float* x, const float* wt,
                           const float* wt_lo, int sm_count, cudaStream_t st, const char** why) {
  if (p.flags & B2J_CT_ROWS) return launch_conv_rows(p, epi, out, x, wt, wt_lo, sm_count, st, why);
  const bool x3 = p.precision == B2J_PREC_TF32X3;
  if (x3 && !wt_lo) { *why = "3xTF32 needs the wt_lo buffer"; return B2J_EINVAL; }
  const bool gemm_like = p.kh == 1 && p.kw == 1 && p.stride_h == 1 && p.stride_w == 1 && p.pad_h == 0 && p.pad_w == 0 &&
                         p.oh == p.h && p.ow == p.w;
  if (gemm_like) { if (p.c % 4 != 0) { *why = "K % 4 (TMA needs 16-byte row pitch)"; return B2J_ENOTIMPL; } }
  else if (p.c % TC_BLOCK_K != 0) { *why = "im2col TMA path needs C % 32 == 0"; return B2J_ENOTIMPL; }
  if (!gemm_like && (p.kw * p.dil_w > 0xFFFFu || p.kh * p.dil_h > 0xFFFFu)) { *why = "filter offsets"; return B2J_ENOTIMPL; }
  if (!tma_api_load()) { *why = "cuTensorMapEncode* not available"; return B2J_ENOTIMPL; }
  const uint32_t M = p.batch * p.oh * p.ow;
  int bn = 64, cg = 1;
  if (x3 && p.o >= 128 && (uint64_t)((M + 255) / 256) * ((p.o + 127) / 128) >= (uint64_t)sm_count / 2) { bn = 128; cg = 2; }   // 3xTF32 pairs
  // (64-wide 3xTF32 tiles as 256 x 64 pairs were measured and dropped: 0.574 vs 0.567 ms on the stage-0 3x3 layers.  Narrow
  // tiles top out at ~345 TFLOP/s of TF32 products in every variant tried -- single pass or 3x, im2col or shared patch, one
  // CTA or a pair, with or without the splitter -- i.e. the per-instruction time of a 128 x 64 x 8 MMA, not operand delivery.)
  bool residual = false;
  for (uint32_t s = 0; s < p.epi.n_steps; ++s) residual |= p.epi.steps[s].kind == B2J_EPK_FULL;
  if (!x3) choose_tc2_tile(M, p.o, p.kpad, residual, sm_count, &bn, &cg);
  CUtensorMap ta, tb, tbl;
  if (!make_tmap_2d(&tb, wt, p.kpad, p.o, p.kpad, TC_BLOCK_K, bn / cg)) { *why = "weight tensor map"; return B2J_ENOTIMPL; }
  tbl = tb;
  if (x3 && !make_tmap_2d(&tbl, wt_lo, p.kpad, p.o, p.kpad, TC_BLOCK_K, bn / cg)) { *why = "weight (lo) tensor map"; return B2J_ENOTIMPL; }
  if (gemm_like) {
    if (!make_tmap_2d(&ta, x, p.c, M, p.c, TC_BLOCK_K, TC_BLOCK_M)) { *why = "activation tensor map"; return B2J_ENOTIMPL; }
  } else {
    if (!make_tmap_im2col(&ta, x, p)) { *why = "im2col tensor map"; return B2J_ENOTIMPL; }
  }
  // residual (first full-tensor epilogue operand): L2-prefetch map over [M, O]
  CUtensorMap tr = tb;
  int has_res = 0;
  for (uint32_t s = 0; s < p.epi.n_steps && !has_res; ++s)
    if (p.epi.steps[s].kind == B2J_EPK_FULL && epi.p[s] != nullptr)
      has_res = make_tmap_plain(&tr, epi.p[s], p.o, M, bn, TC_BLOCK_M * cg) ? 1 : 0;
  // L2 prefetch of the residual tile (B2J_NO_RES_PREFETCH=1 disables it), issued by the MMA thread when a tile's MMAs start.
  // Measured on ResNet-50 b256 (profiles/README.md): issued by the TMA producer (2-3 tiles ahead) 40 % of the prefetched
  // lines were evicted again before the epilogue used them (stage 0: 1.33 GB DRAM reads for 1.03 GB algorithmic); one tile
  // ahead the re-read is gone (1.04 GB) and the residual layers are 9 % faster (0.389 -> 0.354 ms), 12 % faster than without.
  { static int np = -1; if (np < 0) { const char* e = getenv("B2J_NO_RES_PREFETCH"); np = (e && e[0] == '1') ? 1 : 0; } if (np) has_res = 0; }
  const int prog = classify_epilogue(p.epi);
  // 3xTF32 residual tiles fetch the residual with cp.async at the start of the tile's epilogue (tc2_epilogue_role);
  // the TMA L2 prefetch above stays on for it (issued when the tile's MMAs start, i.e. one tile ahead of the epilogue in the
  // epilogue-bound layers): 15.62 -> 15.31 ms per ResNet-50 b256 step; B2J_X3_RES_PREFETCH=0 disables it
  { static int xp = -1; if (xp < 0) { const char* e = getenv("B2J_X3_RES_PREFETCH"); xp = e ? atoi(e) : 1; }
    if (x3 && prog == EPROG_BN_ADD_RELU && (p.o & 3u) == 0 && !xp) has_res = 0; }
  // programs without a residual: the epilogue writes through a TMA store (has_res = 2 hands the output map over in tr's slot)
  if (tma_store_program(prog) && make_tmap_out(&tr, out, p, false)) has_res = 2;
#define TC2_DISPATCH(BN, MODE, X3_, CG_) return launch_conv_tc2_inst<BN, MODE, X3_, CG_>(p, epi, ta, tb, tbl, tr, has_res, prog, out, sm_count, st, why)
  // epilogue-bound residual layers (1x1 expand + residual, K <= 256) on 256 x 128 pair tiles: the RES2 instantiation prefetches the
  // residual into shared memory one chunk ahead (Tc2Cfg); B2J_TF32_RES2=0 keeps the register-load epilogue (A/B runs)
  { static int r2 = -1; if (r2 < 0) { const char* e = getenv("B2J_TF32_RES2"); r2 = e ? atoi(e) : 1; }
    static int r2k = -1; if (r2k < 0) { const char* e = getenv("B2J_TF32_RES2_MAXK"); r2k = e ? atoi(e) : 128; }
    static int r2n = -1; if (r2n < 0) { const char* e = getenv("B2J_TF32_RES2_MINN"); r2n = e ? atoi(e) : 512; }
    // Measured on ResNet-50 b256 (three pipeline stages are the price of the second staging buffer): N = 512, K = 128 layers
    // 0.170 -> 0.153 ms; N = 256, K = 64 (0.333 -> 0.35) and N = 1024, K = 256 (0.098 -> 0.107) lose -- with only two column tiles
    // sharing an operand tile its loads come from DRAM, and eight k-blocks per tile need the deeper ring -- and stay on the
    // five-stage kernel.  B2J_TF32_RES2_MAXK / _MINN move the rule for A/B runs.
    if (r2 && !x3 && cg == 2 && bn == 128 && gemm_like && prog == EPROG_BN_ADD_RELU && has_res == 1 && (p.o & 3u) == 0 &&
        p.kpad <= (uint32_t)r2k && p.kpad <= 256 && p.o >= (uint32_t)r2n)
      return launch_conv_tc2_inst<128, A_TILED, false, 2, true>(p, epi, ta, tb, tbl, tr, has_res, prog, out, sm_count, st, why); }
  if (x3 && cg == 2 && bn == 128) { if (gemm_like) TC2_DISPATCH(128, A_TILED, true, 2); else TC2_DISPATCH(128, A_IM2COL, true, 2); }
  if (x3) { if (gemm_like) TC2_DISPATCH(64, A_TILED, true, 1); else TC2_DISPATCH(64, A_IM2COL, true, 1); }
  if (cg == 2 && bn == 256) { if (gemm_like) TC2_DISPATCH(256, A_TILED, false, 2); else TC2_DISPATCH(256, A_IM2COL, false, 2); }
  if (cg == 2 && bn == 64) { if (gemm_like) TC2_DISPATCH(64, A_TILED, false, 2); else TC2_DISPATCH(64, A_IM2COL, false, 2); }
  if (cg == 2) { if (gemm_like) TC2_DISPATCH(128, A_TILED, false, 2); else TC2_DISPATCH(128, A_IM2COL, false, 2); }
  if (bn == 64) { if (gemm_like) TC2_DISPATCH(64, A_TILED, false, 1); else TC2_DISPATCH(64, A_IM2COL, false, 1); }
  else          { if (gemm_like) TC2_DISPATCH(128, A_TILED, false, 1); else TC2_DISPATCH(128, A_IM2COL, false, 1); }
#undef TC2_DISPATCH
}

}  // namespace b2j
