// conv_tc2: persistent, warp-specialised, TMA-fed tcgen05 implicit GEMM (the fast path of B2J_K_CONV_TC /
// B2J_K_GEMM_TC).  Same math and epilogue as gemm_tc.cuh (v1, kept for channel counts TMA cannot address);
// what changes is how the tensor core is fed and how tiles overlap:
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

enum { A_TILED = 0, A_IM2COL = 1 };

constexpr int TC2_EPI_WARPS = 8;
constexpr int TC2_THREADS = (2 + TC2_EPI_WARPS) * 32;

template <int BLOCK_N> struct Tc2Cfg {
  static constexpr int A_BYTES = TC_A_TILE_BYTES;                       // 128 x 32 floats
  static constexpr int B_BYTES = BLOCK_N * TC_BLOCK_K * 4;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int EPI_PITCH = 36;
  static constexpr int EPI_BYTES = TC2_EPI_WARPS * 32 * EPI_PITCH * 4;   // 36 KB
  static constexpr int STAGES = BLOCK_N <= 64 ? 6 : (BLOCK_N <= 128 ? 5 : 3);
  static constexpr int TMEM_COLS = 2 * BLOCK_N;                          // two accumulators
  static constexpr int OPND_BYTES = B2J_EPI_MAX_STEPS * BLOCK_N * 4;     // decoded per-column epilogue operands
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EPI_BYTES + OPND_BYTES + 1024 + 256;
  static_assert(SMEM_BYTES <= 232448, "exceeds 227 KB of shared memory");
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
__device__ __forceinline__ void tma_prefetch_l2_2d(const CUtensorMap* map, int c0, int c1) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(map), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

// Epilogue of one 32x32 chunk for the persistent kernel.  The step program was decoded once per kernel into
// `ops` (4 bits per step: 0 add, 1 sub, 2 mul, 3 div, 4 max, 5 min, 6 reversed sub, 7 reversed div) and `full_mask`
// (which steps read a full tensor, i.e. the residual); immediates and per-channel vectors were expanded into the
// shared-memory table `opnd[step][column]`, so a step costs one LDS.128 + 32 FP ops per thread instead of
// constant-bank + global round trips.
template <int PITCH, int BLOCK_N>
__device__ __forceinline__ void epilogue_chunk_fast(uint32_t n_steps, uint32_t ops, uint32_t full_mask, const EpiPtrs& epi,
                                                    const float* opnd, int col, const float* stg, float* __restrict__ out,
                                                    uint32_t m_base, uint32_t M, uint32_t n, uint32_t ldo, int lane) {
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
        const uint32_t m = m_base + rr + 4 * it;
        b[it] = m < M ? ld_stream(epi.p[s] + (uint64_t)m * ldo + n) : make_float4(0.f, 0.f, 0.f, 0.f);
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
    const uint32_t m = m_base + rr + 4 * it;
    if (m < M) *reinterpret_cast<float4*>(out + (uint64_t)m * ldo + n) = v[it];
  }
}

template <int BLOCK_N, int A_MODE>
__global__ void __launch_bounds__(TC2_THREADS, 1)
conv_tc2_kernel(const __grid_constant__ b2j_conv_tc_params p, const __grid_constant__ EpiPtrs epi,
                const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                const __grid_constant__ CUtensorMap tmap_res, const int has_res, float* __restrict__ out) {
  using Cfg = Tc2Cfg<BLOCK_N>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t epi_base = smem_base + Cfg::STAGES * Cfg::STAGE_BYTES;
  const uint32_t bar_base = epi_base + Cfg::EPI_BYTES + Cfg::OPND_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (Cfg::STAGES + s); };
  auto tfull_bar = [&](int b) { return bar_base + 8u * (2 * Cfg::STAGES + b); };
  auto tempty_bar = [&](int b) { return bar_base + 8u * (2 * Cfg::STAGES + 2 + b); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * Cfg::STAGES + 4);
  volatile uint32_t* tmem_slot_gen =
      reinterpret_cast<volatile uint32_t*>(smem_gen + Cfg::STAGES * Cfg::STAGE_BYTES + Cfg::EPI_BYTES + Cfg::OPND_BYTES + 8 * (2 * Cfg::STAGES + 4));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t M = p.batch * p.oh * p.ow;
  const uint32_t num_kb = p.kpad / TC_BLOCK_K;
  const uint32_t tiles_n = (p.o + BLOCK_N - 1) / BLOCK_N;
  const uint32_t tiles_m = (M + TC_BLOCK_M - 1) / TC_BLOCK_M;
  const uint32_t num_tiles = tiles_m * tiles_n;

  if (threadIdx.x == 0) {
    for (int s = 0; s < Cfg::STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(tfull_bar(b), 1); mbar_init(tempty_bar(b), TC2_EPI_WARPS); }
    fence_barrier_init();
    prefetch_tmap(&tmap_a);
    prefetch_tmap(&tmap_b);
  }
  if (warp == 1) tmem_alloc<Cfg::TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_gen;

  if (warp == 0) {
    // ======================================= TMA producer =======================================
    if (lane == 0) {
      uint32_t it = 0;      // global k-block counter across tiles -> stage / phase
      const uint32_t cblocks = A_MODE == A_IM2COL ? p.c / TC_BLOCK_K : 1;
      for (uint32_t t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        const uint32_t m0 = (t / tiles_n) * TC_BLOCK_M, n0 = (t % tiles_n) * BLOCK_N;
        // the residual tile this output tile will add in its epilogue: pull it into L2 now (the producer runs
        // 1-2 tiles ahead of the epilogue), so the epilogue's loads are L2 hits instead of HBM round trips
        if (has_res) tma_prefetch_l2_2d(&tmap_res, (int)n0, (int)m0);
        int bw = 0, bh = 0, bn = 0;
        if (A_MODE == A_IM2COL) {
          const uint32_t ow = m0 % p.ow, t1 = m0 / p.ow;
          bw = (int)(ow * p.stride_w) - p.pad_w;
          bh = (int)((t1 % p.oh) * p.stride_h) - p.pad_h;
          bn = (int)(t1 / p.oh);
        }
        for (uint32_t kb = 0; kb < num_kb; ++kb, ++it) {
          const int s = it % Cfg::STAGES;
          mbar_wait_sleepy(empty_bar(s), ((it / Cfg::STAGES) & 1u) ^ 1u);
          const uint32_t a_dst = smem_base + s * Cfg::STAGE_BYTES, b_dst = a_dst + Cfg::A_BYTES;
          mbar_expect_tx(full_bar(s), Cfg::STAGE_BYTES);
          if (A_MODE == A_IM2COL) {
            const uint32_t tap = kb / cblocks, cb = kb - tap * cblocks;
            const uint32_t kh = tap / p.kw, kw = tap - kh * p.kw;
            tma_load_im2col_4d(a_dst, &tmap_a, full_bar(s), (int)(cb * TC_BLOCK_K), bw, bh, bn,
                               (uint16_t)(kw * p.dil_w), (uint16_t)(kh * p.dil_h));
          } else {
            tma_load_2d(a_dst, &tmap_a, full_bar(s), (int)(kb * TC_BLOCK_K), (int)m0);
          }
          tma_load_2d(b_dst, &tmap_b, full_bar(s), (int)(kb * TC_BLOCK_K), (int)n0);
        }
      }
    }
  } else if (warp == 1) {
    // ======================================= MMA issuer =========================================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_tf32(TC_BLOCK_M, BLOCK_N);
      uint32_t it = 0, tile_i = 0;
      for (uint32_t t = blockIdx.x; t < num_tiles; t += gridDim.x, ++tile_i) {
        const uint32_t ab = tile_i & 1u;
        mbar_wait_sleepy(tempty_bar(ab), ((tile_i >> 1) & 1u) ^ 1u);     // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + ab * BLOCK_N;
        for (uint32_t kb = 0; kb < num_kb; ++kb, ++it) {
          const int s = it % Cfg::STAGES;
          mbar_wait_sleepy(full_bar(s), (it / Cfg::STAGES) & 1u);
          tc_fence_after();
          const uint32_t stage = smem_base + s * Cfg::STAGE_BYTES;
          const uint64_t adesc = make_smem_desc(stage), bdesc = make_smem_desc(stage + Cfg::A_BYTES);
#pragma unroll
          for (int k = 0; k < TC_BLOCK_K / 8; ++k)
            umma_tf32(tmem_d, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (kb | (uint32_t)k) != 0u);
          umma_commit(empty_bar(s));
        }
        umma_commit(tfull_bar(ab));
      }
    }
  } else {
    // ======================================= epilogue ===========================================
    const int ew = warp - 2;                    // 0..7
    const int q = warp & 3;                     // TMEM lane quarter accessible to this warp
    const int half = ew >> 2;                   // column half
    float* stg = reinterpret_cast<float*>(smem_gen + Cfg::STAGES * Cfg::STAGE_BYTES) + ew * 32 * Cfg::EPI_PITCH;
    constexpr int COLS_PER_WARP = BLOCK_N / 2;
    float* opnd = reinterpret_cast<float*>(smem_gen + Cfg::STAGES * Cfg::STAGE_BYTES + Cfg::EPI_BYTES);
    // decode the step program once
    const uint32_t n_steps = p.epi.n_steps;
    uint32_t ops = 0, full_mask = 0;
    for (uint32_t s = 0; s < n_steps; ++s) {
      const b2j_epi_step st = p.epi.steps[s];
      const bool sw = st.flags & B2J_STEP_SWAP;
      uint32_t code = st.op == B2J_OP_ADD_F ? 0u : st.op == B2J_OP_SUB_F ? (sw ? 6u : 1u) : st.op == B2J_OP_MUL_F ? 2u
                    : st.op == B2J_OP_DIV_F ? (sw ? 7u : 3u) : st.op == B2J_OP_MAX_F ? 4u : st.op == B2J_OP_MIN_F ? 5u : 15u;
      ops |= code << (4 * s);
      if (st.kind == B2J_EPK_FULL) full_mask |= 1u << s;
    }
    const int etid = threadIdx.x - 64;          // 0..255 within the epilogue warps
    uint32_t table_n0 = 0xFFFFFFFFu;
    uint32_t tile_i = 0;
    for (uint32_t t = blockIdx.x; t < num_tiles; t += gridDim.x, ++tile_i) {
      const uint32_t m0 = (t / tiles_n) * TC_BLOCK_M, n0 = (t % tiles_n) * BLOCK_N;
      if (n0 != table_n0) {
        // (re)build opnd[step][column] for this column range: immediates broadcast, per-channel vectors copied
        asm volatile("bar.sync 1, 256;" ::: "memory");      // everybody is done reading the old table
        for (uint32_t idx = etid; idx < n_steps * BLOCK_N; idx += TC2_EPI_WARPS * 32) {
          const uint32_t s = idx / BLOCK_N, c = idx - s * BLOCK_N;
          const b2j_epi_step st = p.epi.steps[s];
          float val = 0.0f;
          if (st.kind == B2J_EPK_IMM) val = __uint_as_float(st.imm);
          else if (st.kind == B2J_EPK_CHANNEL && n0 + c < p.o) val = __ldg(epi.p[s] + n0 + c);
          opnd[idx] = val;
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
        table_n0 = n0;
      }
      const uint32_t ab = tile_i & 1u;
      mbar_wait_sleepy(tfull_bar(ab), (tile_i >> 1) & 1u);
      tc_fence_after();
#pragma unroll 1
      for (int cc = 0; cc < COLS_PER_WARP; cc += 32) {
        const int col0 = half * COLS_PER_WARP + cc;
        uint32_t r[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + ab * BLOCK_N + (uint32_t)col0, r);
        if (cc + 32 >= COLS_PER_WARP) {          // last TMEM read of this tile: hand the accumulator back
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(tempty_bar(ab));
        }
#pragma unroll
        for (int j = 0; j < 8; ++j)
          *reinterpret_cast<uint4*>(stg + lane * Cfg::EPI_PITCH + 4 * j) = make_uint4(r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
        __syncwarp();
        const int col = col0 + 4 * (lane & 7);
        const uint32_t n = n0 + col;
        if (n < p.o)
          epilogue_chunk_fast<Cfg::EPI_PITCH, BLOCK_N>(n_steps, ops, full_mask, epi, opnd, col, stg, out, m0 + q * 32, M, n, p.o, lane);
        __syncwarp();
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
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

template <int BLOCK_N, int A_MODE>
static int launch_conv_tc2_inst(const b2j_conv_tc_params& p, const EpiPtrs& epi, const CUtensorMap& ta, const CUtensorMap& tb,
                                const CUtensorMap& tr, int has_res, float* out, int sm_count, cudaStream_t st, const char** why) {
  using Cfg = Tc2Cfg<BLOCK_N>;
  static bool configured = false;
  auto kern = conv_tc2_kernel<BLOCK_N, A_MODE>;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
    if (e != cudaSuccess) { *why = cudaGetErrorString(e); return B2J_ECUDA; }
    configured = true;
  }
  const uint32_t M = p.batch * p.oh * p.ow;
  const uint32_t tiles = ((M + TC_BLOCK_M - 1) / TC_BLOCK_M) * ((p.o + BLOCK_N - 1) / BLOCK_N);
  const unsigned grid = tiles < (uint32_t)sm_count ? tiles : (unsigned)sm_count;     // persistent: one CTA per SM
  kern<<<grid, TC2_THREADS, Cfg::SMEM_BYTES, st>>>(p, epi, ta, tb, tr, has_res, out);
  return B2J_OK;
}

// Returns B2J_ENOTIMPL (why set) when this problem has to go to the v1 kernel.
static int launch_conv_tc2(const b2j_conv_tc_params& p, const EpiPtrs& epi, float* out, const float* x, const float* wt, int sm_count,
                           cudaStream_t st, const char** why) {
  if (p.precision != B2J_PREC_TF32) { *why = "v2 is TF32 only"; return B2J_ENOTIMPL; }
  if (p.o % 4 != 0) { *why = "O % 4"; return B2J_ENOTIMPL; }
  const bool gemm_like = p.kh == 1 && p.kw == 1 && p.stride_h == 1 && p.stride_w == 1 && p.pad_h == 0 && p.pad_w == 0 &&
                         p.oh == p.h && p.ow == p.w;
  if (gemm_like) { if (p.c % 4 != 0) { *why = "K % 4 (TMA needs 16-byte row pitch)"; return B2J_ENOTIMPL; } }
  else if (p.c % TC_BLOCK_K != 0) { *why = "im2col TMA path needs C % 32 == 0"; return B2J_ENOTIMPL; }
  if (!gemm_like && (p.kw * p.dil_w > 0xFFFFu || p.kh * p.dil_h > 0xFFFFu)) { *why = "filter offsets"; return B2J_ENOTIMPL; }
  if (!tma_api_load()) { *why = "cuTensorMapEncode* not available"; return B2J_ENOTIMPL; }
  const uint32_t M = p.batch * p.oh * p.ow;
  const int bn = p.o <= 64 ? 64 : 128;
  CUtensorMap ta, tb;
  if (!make_tmap_2d(&tb, wt, p.kpad, p.o, p.kpad, TC_BLOCK_K, bn)) { *why = "weight tensor map"; return B2J_ENOTIMPL; }
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
      has_res = make_tmap_plain(&tr, epi.p[s], p.o, M, bn, TC_BLOCK_M) ? 1 : 0;
#define TC2_DISPATCH(BN, MODE) return launch_conv_tc2_inst<BN, MODE>(p, epi, ta, tb, tr, has_res, out, sm_count, st, why)
  if (bn == 64) { if (gemm_like) TC2_DISPATCH(64, A_TILED); else TC2_DISPATCH(64, A_IM2COL); }
  else          { if (gemm_like) TC2_DISPATCH(128, A_TILED); else TC2_DISPATCH(128, A_IM2COL); }
#undef TC2_DISPATCH
}

}  // namespace b2j
