// conv_patch: stride-1 k x k convolutions with few output channels (N <= 128) on tcgen05, fed from an input PATCH that
// is fetched once per 32-channel block and shared by all filter taps.
//
// Why: conv_tc2's im2col TMA fetches the 128 x 32 activation operand once per filter tap -- 9 x for a 3x3 filter -- and
// with N <= 128 the L2 -> SM fabric (~38 B/clk/SM), not the tensor pipe, bounds the layer (ResNet-50 stage-0 3x3:
// 233 TFLOP/s, profiles/r01_ncu_full_conv_tc2.md).  Here the M tile is R full output rows of one image, each padded to a
// power-of-two pitch P (R * P = 128, P >= W + KW - 1).  One tiled 4-D TMA load brings the (R + KH - 1) x P input pixels
// x 32 channels the tile needs -- image borders arrive as zeros through TMA out-of-bounds fill -- into SWIZZLE_128B
// shared memory, pixel-major.  Because the tile's rows have the patch's own pitch, the operand of tap (kh, kw) is simply
// the 128 consecutive patch pixels starting at pixel kh * P + kw: the same buffer, read through a shared-memory
// descriptor whose start address is shifted by (kh * P + kw) * 128 bytes (the 128-byte swizzle is a function of the
// address bits, so a shifted window stays consistent with what TMA wrote).  Output columns >= OW of each row are
// computed and dropped (W / P utilisation: 56/64, 28/32).  Activation traffic per tile falls from KH*KW x 16 KB to
// (R + KH - 1) * P * 128 B per channel block: 4.5 x (56 x 56) to 6 x (28 x 28) less.
//
// Roles: warp 0 weight-tile producer (TMA, ring of NB stages, one 32-wide k-block per tap), warp 2 patch producer (TMA,
// ring of NA patches), warp 1 MMA issuer (tcgen05.mma kind::tf32 into two TMEM accumulators), warps 3.. epilogue
// (the straight-line / generic programs of conv_tc2.cuh with the RowPatch output mapping): 8 warps with one CTA per SM,
// 4 warps -- one per TMEM lane quarter -- in the default mode of two co-resident CTAs per SM (PatchCfg<BLOCK_N, true>).
#pragma once
#include "conv_tc2.cuh"

namespace b2j {

// TWO = true (the default mode, B2J_ENABLE_PATCH=2): two co-resident CTAs per SM.  One CTA's TMA -> MMA -> epilogue chain is
// latency-bound, a second, independent chain on the same SM hides it (as Tc2Cfg::TWO_CTAS does for the im2col kernel):
// 2 patch slots, 2-3 weight stages and ONE epilogue warp per TMEM lane quarter, so that a CTA stays within half an SM's
// shared memory (113 KB).
template <int BLOCK_N, bool TWO = false> struct PatchCfg {
  static constexpr int B_BYTES = BLOCK_N * TC_BLOCK_K * 4;
  static constexpr int NB = TWO ? (BLOCK_N <= 64 ? 3 : 2) : (BLOCK_N <= 64 ? 6 : 4);      // weight-tile stages
  static constexpr int A_RING_BYTES = TWO ? (BLOCK_N <= 64 ? 66 * 1024 : 52 * 1024) : (BLOCK_N <= 64 ? 128 * 1024 : 96 * 1024);
  static constexpr int EPI_WARPS = TWO ? 4 : 8;
  static constexpr int THREADS = (3 + EPI_WARPS) * 32;
  static constexpr int EPI_PITCH = 36;
  static constexpr int EPI_BYTES = EPI_WARPS * 32 * EPI_PITCH * 4;
  static constexpr int OPND_BYTES = B2J_EPI_MAX_STEPS * BLOCK_N * 4;
  static constexpr int TMEM_COLS = 2 * BLOCK_N;
  static constexpr int MAX_NA = 8;
  static constexpr int SMEM_BYTES = A_RING_BYTES + NB * B_BYTES + EPI_BYTES + OPND_BYTES + 1024 + 512;
  static_assert(SMEM_BYTES <= (TWO ? 115712 : 232448), "exceeds the shared memory of one SM (227 KB, or 113 KB each for two co-resident CTAs)");
  static_assert((TWO ? 2 : 1) * TMEM_COLS <= 512, "TMEM");
};

__device__ __forceinline__ void tma_load_tile_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c, int w, int h, int n) {
  asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
               ::"r"(dst), "l"(map), "r"(bar), "r"(c), "r"(w), "r"(h), "r"(n) : "memory");
}

struct PatchGeom {
  uint32_t P, log2P, R, rbs;       // pitch (pixels), rows per tile, row blocks per image
  uint32_t patch_bytes, slot_bytes, na;
};

__host__ __device__ inline bool patch_geometry(const b2j_conv_tc_params& p, int a_ring_bytes, PatchGeom* g) {
  uint32_t P = 8, l = 3;
  while (P < p.w + p.kw - 1 || P < p.ow + p.kw - 1) { P <<= 1; ++l; }
  if (P > 128) return false;
  g->P = P; g->log2P = l; g->R = 128 / P;
  g->rbs = (p.oh + g->R - 1) / g->R;
  g->patch_bytes = (g->R + p.kh - 1) * P * 128;
  g->slot_bytes = (g->patch_bytes + (p.kw - 1) * 128 + 1023) / 1024 * 1024;   // + the (KW-1)-pixel overrun of the last tap's window
  g->na = (uint32_t)a_ring_bytes / g->slot_bytes;
  if (g->na > 8) g->na = 8;
  return g->na >= 2 && g->R + p.kh - 1 <= 256;
}

template <int BLOCK_N, bool TWO, int PROG>
__device__ __forceinline__ void patch_epilogue_role(const b2j_conv_tc_params& p, const EpiPtrs& epi, uint8_t* smem_gen, uint32_t epi_off,
                                                    uint32_t tfull0, uint32_t tempty0, uint32_t tmem_base, float* __restrict__ out,
                                                    const PatchGeom& g, uint32_t tiles_n, uint32_t num_tiles) {
  using Cfg = PatchCfg<BLOCK_N, TWO>;
  constexpr int COLS_PER_WARP = BLOCK_N / (Cfg::EPI_WARPS / 4);     // 8 warps: two per lane quarter, half the columns each
  constexpr int EPI_THREADS = Cfg::EPI_WARPS * 32;
  constexpr bool HAS_RES = PROG == EPROG_BN_ADD_RELU;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ew = warp - 3;
  const int q = warp & 3, half = ew >> 2;
  float* stg = reinterpret_cast<float*>(smem_gen + epi_off) + ew * 32 * Cfg::EPI_PITCH;
  float* opnd = reinterpret_cast<float*>(smem_gen + epi_off + Cfg::EPI_BYTES);
  const uint32_t n_steps = p.epi.n_steps;
  uint32_t ops = 0, full_mask = 0;
  if (PROG == EPROG_GENERIC) {
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
  const int rnd = (int)(p.flags & 3u) | ((p.o & 3u) ? EPI_SCALAR_IO : 0);
  const float* resp = HAS_RES ? epi.p[3] : nullptr;
  const int etid = ew * 32 + lane;
  const int cj = lane & 7, rr = lane >> 3;
  uint32_t table_n0 = 0xFFFFFFFFu, tile_i = 0;
  for (uint32_t t = blockIdx.x; t < num_tiles; t += gridDim.x, ++tile_i) {
    const uint32_t mt = t / tiles_n, n0 = (t % tiles_n) * BLOCK_N;
    const uint32_t img = mt / g.rbs, oh0 = (mt % g.rbs) * g.R;
    const RowPatch rm{(img * p.oh + oh0) * p.ow, g.P, g.log2P, p.ow, p.oh - oh0, (uint32_t)q * 32u};
    if (n0 != table_n0) {
      if (Cfg::EPI_WARPS == 8) asm volatile("bar.sync 1, 256;" ::: "memory"); else asm volatile("bar.sync 1, 128;" ::: "memory");
      for (uint32_t idx = etid; idx < n_steps * BLOCK_N; idx += EPI_THREADS) {
        const uint32_t s = idx / BLOCK_N, c = idx - s * BLOCK_N;
        const b2j_epi_step st = p.epi.steps[s];
        float val = 0.0f;
        if (st.kind == B2J_EPK_IMM) val = __uint_as_float(st.imm);
        else if (st.kind == B2J_EPK_CHANNEL && n0 + c < p.o) val = __ldg(epi.p[s] + n0 + c);
        opnd[idx] = val;
      }
      if (Cfg::EPI_WARPS == 8) asm volatile("bar.sync 1, 256;" ::: "memory"); else asm volatile("bar.sync 1, 128;" ::: "memory");
      table_n0 = n0;
    }
    const uint32_t ab = tile_i & 1u;
    mbar_wait(tfull0 + 8u * ab, (tile_i >> 1) & 1u);
    tc_fence_after();
#pragma unroll 1
    for (int cc = 0; cc < COLS_PER_WARP; cc += 32) {
      const int col0 = half * COLS_PER_WARP + cc;
      const int col = col0 + 4 * cj;
      const uint32_t n = n0 + col;
      float4 res_a[4], res_b[4];
      if (HAS_RES) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const uint32_t m = rm(rr + 4 * i);
          res_a[i] = (m != ROW_NONE && n < p.o) ? ld_res_any(resp + (uint64_t)m * p.o, n, p.o, rnd) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
      uint32_t r[32];
      tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + ab * BLOCK_N + (uint32_t)col0, r);
      if (cc + 32 >= COLS_PER_WARP) {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tempty0 + 8u * ab);
      }
#pragma unroll
      for (int j = 0; j < 8; ++j)
        *reinterpret_cast<uint4*>(stg + lane * Cfg::EPI_PITCH + 4 * j) = make_uint4(r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
      __syncwarp();
      if (HAS_RES) {
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
}

template <int BLOCK_N, bool TWO>
__global__ void __launch_bounds__(PatchCfg<BLOCK_N, TWO>::THREADS, TWO ? 2 : 1)
conv_patch_kernel(const __grid_constant__ b2j_conv_tc_params p, const __grid_constant__ EpiPtrs epi,
                  const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                  const __grid_constant__ PatchGeom g, const int epi_prog, float* __restrict__ out) {
  using Cfg = PatchCfg<BLOCK_N, TWO>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  constexpr uint32_t B_OFF = Cfg::A_RING_BYTES, EPI_OFF = B_OFF + Cfg::NB * Cfg::B_BYTES;
  constexpr uint32_t BAR_OFF = EPI_OFF + Cfg::EPI_BYTES + Cfg::OPND_BYTES;
  const uint32_t bar_base = smem_base + BAR_OFF;
  auto a_full = [&](int s) { return bar_base + 8u * s; };
  auto a_empty = [&](int s) { return bar_base + 8u * (Cfg::MAX_NA + s); };
  auto b_full = [&](int s) { return bar_base + 8u * (2 * Cfg::MAX_NA + s); };
  auto b_empty = [&](int s) { return bar_base + 8u * (2 * Cfg::MAX_NA + Cfg::NB + s); };
  const uint32_t tfull0 = bar_base + 8u * (2 * Cfg::MAX_NA + 2 * Cfg::NB), tempty0 = tfull0 + 16u;
  const uint32_t tmem_slot = tempty0 + 16u;
  volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + BAR_OFF + 8 * (2 * Cfg::MAX_NA + 2 * Cfg::NB + 4));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t cblocks = p.c / TC_BLOCK_K, taps = p.kh * p.kw;
  const uint32_t tiles_n = (p.o + BLOCK_N - 1) / BLOCK_N;
  const uint32_t num_tiles = p.batch * g.rbs * tiles_n;

  if (threadIdx.x == 0) {
    for (int s = 0; s < Cfg::MAX_NA; ++s) { mbar_init(a_full(s), 1); mbar_init(a_empty(s), 1); }
    for (int s = 0; s < Cfg::NB; ++s) { mbar_init(b_full(s), 1); mbar_init(b_empty(s), 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(tfull0 + 8u * b, 1); mbar_init(tempty0 + 8u * b, Cfg::EPI_WARPS); }
    fence_barrier_init();
    prefetch_tmap(&tmap_a);
    prefetch_tmap(&tmap_b);
  }
  if (warp == 1) tmem_alloc<Cfg::TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_gen;

  if (warp == 2) {
    // ======================================= patch producer ======================================
    if (lane == 0) {
      uint32_t ia = 0;
      for (uint32_t t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        const uint32_t mt = t / tiles_n;
        const uint32_t img = mt / g.rbs, oh0 = (mt % g.rbs) * g.R;
        for (uint32_t cb = 0; cb < cblocks; ++cb, ++ia) {
          const uint32_t s = ia % g.na;
          mbar_wait_sleepy(a_empty(s), ((ia / g.na) & 1u) ^ 1u);
          mbar_expect_tx(a_full(s), g.patch_bytes);
          tma_load_tile_4d(smem_base + s * g.slot_bytes, &tmap_a, a_full(s), (int)(cb * TC_BLOCK_K), -p.pad_w, (int)oh0 - p.pad_h, (int)img);
        }
      }
    }
  } else if (warp == 0) {
    // ======================================= weight-tile producer ================================
    if (lane == 0) {
      uint32_t ib = 0;
      for (uint32_t t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        const uint32_t n0 = (t % tiles_n) * BLOCK_N;
        for (uint32_t cb = 0; cb < cblocks; ++cb)
          for (uint32_t tap = 0; tap < taps; ++tap, ++ib) {
            const uint32_t s = ib % Cfg::NB;
            mbar_wait_sleepy(b_empty(s), ((ib / Cfg::NB) & 1u) ^ 1u);
            mbar_expect_tx(b_full(s), Cfg::B_BYTES);
            tma_load_2d(smem_base + B_OFF + s * Cfg::B_BYTES, &tmap_b, b_full(s), (int)(tap * p.c + cb * TC_BLOCK_K), (int)n0);
          }
      }
    }
  } else if (warp == 1) {
    // ======================================= MMA issuer =========================================
    // the whole warp runs the loop, one elected lane issues (see conv_tc2.cuh: back-to-back UTCHMMAs from uniform registers)
    if (B2J_WARP_MMA || lane == 0) {
      const bool leader = B2J_WARP_MMA ? elect_one() : true;
      constexpr uint32_t idesc = make_idesc_tf32(TC_BLOCK_M, BLOCK_N);
      uint32_t ia = 0, ib = 0, tile_i = 0;
      for (uint32_t t = blockIdx.x; t < num_tiles; t += gridDim.x, ++tile_i) {
        const uint32_t ab = tile_i & 1u;
        mbar_wait_sleepy(tempty0 + 8u * ab, ((tile_i >> 1) & 1u) ^ 1u);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + ab * BLOCK_N;
        for (uint32_t cb = 0; cb < cblocks; ++cb, ++ia) {
          const uint32_t sa = ia % g.na;
          mbar_wait_sleepy(a_full(sa), (ia / g.na) & 1u);
          tc_fence_after();
          const uint32_t patch = smem_base + sa * g.slot_bytes;
          uint32_t tap = 0;
          for (uint32_t kh = 0; kh < p.kh; ++kh)
            for (uint32_t kw = 0; kw < p.kw; ++kw, ++tap, ++ib) {
              const uint32_t sb = ib % Cfg::NB;
              mbar_wait_sleepy(b_full(sb), (ib / Cfg::NB) & 1u);
              tc_fence_after();
              // the operand of this tap: 128 consecutive patch pixels starting at pixel kh*P + kw
              const uint64_t adesc = make_smem_desc(patch + ((kh << g.log2P) + kw) * 128u);
              const uint64_t bdesc = make_smem_desc(smem_base + B_OFF + sb * Cfg::B_BYTES);
              if (leader) {
#pragma unroll
                for (int k = 0; k < TC_BLOCK_K / 8; ++k)
                  umma_tf32(tmem_d, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (cb | tap | (uint32_t)k) != 0u);
                umma_commit(b_empty(sb));
              }
              if (B2J_WARP_MMA) __syncwarp();
            }
          if (leader) umma_commit(a_empty(sa));
          if (B2J_WARP_MMA) __syncwarp();
        }
        if (leader) umma_commit(tfull0 + 8u * ab);
        if (B2J_WARP_MMA) __syncwarp();
      }
    }
  } else {
    // ======================================= epilogue ===========================================
    switch (epi_prog) {
      case EPROG_BN:          patch_epilogue_role<BLOCK_N, TWO, EPROG_BN>(p, epi, smem_gen, EPI_OFF, tfull0, tempty0, tmem_base, out, g, tiles_n, num_tiles); break;
      case EPROG_BN_RELU:     patch_epilogue_role<BLOCK_N, TWO, EPROG_BN_RELU>(p, epi, smem_gen, EPI_OFF, tfull0, tempty0, tmem_base, out, g, tiles_n, num_tiles); break;
      case EPROG_BN_ADD_RELU: patch_epilogue_role<BLOCK_N, TWO, EPROG_BN_ADD_RELU>(p, epi, smem_gen, EPI_OFF, tfull0, tempty0, tmem_base, out, g, tiles_n, num_tiles); break;
      case EPROG_BIAS:        patch_epilogue_role<BLOCK_N, TWO, EPROG_BIAS>(p, epi, smem_gen, EPI_OFF, tfull0, tempty0, tmem_base, out, g, tiles_n, num_tiles); break;
      case EPROG_BIAS_RELU:   patch_epilogue_role<BLOCK_N, TWO, EPROG_BIAS_RELU>(p, epi, smem_gen, EPI_OFF, tfull0, tempty0, tmem_base, out, g, tiles_n, num_tiles); break;
      default:                patch_epilogue_role<BLOCK_N, TWO, EPROG_GENERIC>(p, epi, smem_gen, EPI_OFF, tfull0, tempty0, tmem_base, out, g, tiles_n, num_tiles); break;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
  }
}

// ---- host side -------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled4)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                     const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                     CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static bool make_tmap_patch(CUtensorMap* map, const float* x, const b2j_conv_tc_params& p, const PatchGeom& g) {
  cuuint64_t dims[4] = {p.c, p.w, p.h, p.batch};
  cuuint64_t strides[3] = {(cuuint64_t)p.c * 4, (cuuint64_t)p.w * p.c * 4, (cuuint64_t)p.h * p.w * p.c * 4};
  cuuint32_t box[4] = {TC_BLOCK_K, g.P, g.R + p.kh - 1, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  return g_tma.tiled(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void*)x, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int BLOCK_N, bool TWO>
static int launch_conv_patch_inst(const b2j_conv_tc_params& p, const EpiPtrs& epi, const CUtensorMap& ta, const CUtensorMap& tb,
                                  const PatchGeom& g, int prog, float* out, int sm_count, cudaStream_t st, const char** why) {
  using Cfg = PatchCfg<BLOCK_N, TWO>;
  static bool configured = false;
  auto kern = conv_patch_kernel<BLOCK_N, TWO>;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
    if (e != cudaSuccess) { *why = cudaGetErrorString(e); return B2J_ECUDA; }
    configured = true;
  }
  const uint32_t tiles = p.batch * g.rbs * ((p.o + BLOCK_N - 1) / BLOCK_N);
  const uint32_t slots = (uint32_t)sm_count * (TWO ? 2 : 1);
  const unsigned grid = tiles < slots ? tiles : slots;
  kern<<<grid, Cfg::THREADS, Cfg::SMEM_BYTES, st>>>(p, epi, ta, tb, g, prog, out);
  return B2J_OK;
}

// B2J_ENOTIMPL when the problem is not a patch problem (the caller then uses conv_tc2).
static int launch_conv_patch(const b2j_conv_tc_params& p, const EpiPtrs& epi, float* out, const float* x, const float* wt, int sm_count,
                             cudaStream_t st, const char** why) {
  // B2J_ENABLE_PATCH = 2 (default): two co-resident CTAs per SM; 1: one CTA per SM; 0: off (conv_tc2's im2col path).
  // Measured on B200, ResNet-50 b256 (profiles/r01_patch_kernel.md): one CTA per SM cuts the L2 -> SM traffic of the stage-0
  // 3x3 layers from 2.77 GB to 1.53 GB as designed but is SLOWER than the im2col kernel (0.300 vs 0.175 ms: a single
  // TMA -> MMA -> epilogue chain is latency-bound); two CTAs per SM are faster than it: stage-0 3x3 (N = 64) 0.175 -> 0.170 ms,
  // stage-1 3x3 (N = 128) 0.144 -> 0.119 ms.
  static int enabled = -1;
  if (enabled < 0) { const char* e = getenv("B2J_ENABLE_PATCH"); enabled = e ? ((e[0] == '1' || e[0] == '2') ? e[0] - '0' : 0) : 2; }
  if (!enabled) { *why = "patch kernel disabled"; return B2J_ENOTIMPL; }
  const bool two = enabled == 2;
  if (p.precision != B2J_PREC_TF32) { *why = "single-pass TF32 only"; return B2J_ENOTIMPL; }
  if (p.stride_h != 1 || p.stride_w != 1 || p.dil_h != 1 || p.dil_w != 1 || p.kh * p.kw < 2) { *why = "stride/dilation"; return B2J_ENOTIMPL; }
  // N <= 64 only since the MMAs are issued from warp-uniform control flow (conv_tc2.cuh): for the 128-wide stage-1 3x3 layers of
  // ResNet-50 the paired im2col kernel is now faster (0.100 vs 0.110 ms); B2J_PATCH_MAXN=128 brings the old rule back
  static int maxn = -1;
  if (maxn < 0) { const char* e = getenv("B2J_PATCH_MAXN"); maxn = e ? atoi(e) : 64; }
  if (p.c % TC_BLOCK_K != 0 || p.o > 128 || (int)p.o > maxn || p.pad_h < 0 || p.pad_w < 0) { *why = "channels"; return B2J_ENOTIMPL; }
  if (p.kpad != p.kh * p.kw * p.c) { *why = "kpad"; return B2J_ENOTIMPL; }
  const int bn = p.o <= 64 ? 64 : 128;
  PatchGeom g;
  const int ring_bytes = two ? (bn == 64 ? PatchCfg<64, true>::A_RING_BYTES : PatchCfg<128, true>::A_RING_BYTES)
                             : (bn == 64 ? PatchCfg<64>::A_RING_BYTES : PatchCfg<128>::A_RING_BYTES);
  if (!patch_geometry(p, ring_bytes, &g)) { *why = "patch geometry"; return B2J_ENOTIMPL; }
  // utilisation of the padded tile: worth it only when most of the 128 rows are real outputs
  if ((double)p.ow / g.P < 0.74 || (double)p.oh / (g.rbs * g.R) < 0.74) { *why = "tile utilisation"; return B2J_ENOTIMPL; }
  if (!tma_api_load()) { *why = "cuTensorMapEncode* not available"; return B2J_ENOTIMPL; }
  CUtensorMap ta, tb;
  if (!make_tmap_2d(&tb, wt, p.kpad, p.o, p.kpad, TC_BLOCK_K, bn)) { *why = "weight tensor map"; return B2J_ENOTIMPL; }
  if (!make_tmap_patch(&ta, x, p, g)) { *why = "patch tensor map"; return B2J_ENOTIMPL; }
  const int prog = classify_epilogue(p.epi);
  if (two) {
    if (bn == 64) return launch_conv_patch_inst<64, true>(p, epi, ta, tb, g, prog, out, sm_count, st, why);
    return launch_conv_patch_inst<128, true>(p, epi, ta, tb, g, prog, out, sm_count, st, why);
  }
  if (bn == 64) return launch_conv_patch_inst<64, false>(p, epi, ta, tb, g, prog, out, sm_count, st, why);
  return launch_conv_patch_inst<128, false>(p, epi, ta, tb, g, prog, out, sm_count, st, why);
}

}  // namespace b2j
