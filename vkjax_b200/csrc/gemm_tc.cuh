// tcgen05 implicit-GEMM convolution / GEMM for sm_100a  (B2J_K_CONV_TC, B2J_K_GEMM_TC).
//
//   out[m, n] = epilogue( sum_k A[m, k] * Wt[n, k] )      M = B*OH*OW, N = O, K = KH*KW*C
//
// A is the (virtual) im2col matrix of the NHWC activation tensor: row m is an output pixel, column
// k = (kh*KW + kw)*C + c.  Wt is the weight tensor pre-arranged [O][Kpad] (K-major, zero padded) by
// weight_prep_kernel.  Both operands are fp32 in HBM and are fed to the 5th-gen tensor cores as TF32.
//
// Blackwell mapping (what the reference's one-thread-per-output conv2d.comp becomes):
//  * CTA tile 128 (M) x BLOCK_N, K step 32 floats = one 128-byte swizzle row.
//  * 8 producer warps gather A rows (any stride / padding / dilation; zero fill) and Wt rows with
//    128-bit loads, one k-block ahead in registers, and store them into shared memory in the canonical
//    K-major SWIZZLE_128B layout (16-byte chunk j of row r lands at chunk j ^ (r & 7)); a
//    fence.proxy.async + mbarrier hands the stage to the tensor core.
//  * 1 MMA warp: a single thread issues tcgen05.mma.cta_group::1.kind::tf32 (M=128, N=BLOCK_N, K=8),
//    4 per stage (12 in 3xTF32 mode), accumulating in TMEM; tcgen05.commit releases the stage.
//  * epilogue (the 8 producer warps again): tcgen05.ld 32x32b.x32 -> registers -> shared-memory
//    transpose -> per-channel / residual / ReLU steps on float4 -> coalesced 128-bit global stores.
//  * 3xTF32 (B2J_PREC_TF32X3): A and Wt are split on the fly into hi = tf32(x) and lo = x - hi;
//    D += A_lo*B_hi + A_hi*B_lo + A_hi*B_hi  (error ~2^-22, fp32-class; the "fp32-exact variant").
//  * plain TF32: operands are rounded to nearest TF32 (cvt.rna) by the producer instead of being
//    truncated by the tensor core.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/b2jax.h"
#include "contraction_simt.cuh"

namespace b2j {

constexpr int TC_BLOCK_M = 128;
constexpr int TC_BLOCK_K = 32;                 // floats: 128 bytes = one SWIZZLE_128B row
constexpr int TC_PRODUCER_WARPS = 8;
constexpr int TC_THREADS = (TC_PRODUCER_WARPS + 1) * 32;
constexpr int TC_A_TILE_BYTES = TC_BLOCK_M * TC_BLOCK_K * 4;   // 16 KB

// ---- PTX wrappers ------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0u;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) { }
}
// Same, with a suspend-time hint: the single-thread producer / MMA roles then sleep in hardware instead of
// re-issuing try_wait every ~20 cycles and stealing issue slots from the epilogue warps of their SM sub-partition.
__device__ __forceinline__ void mbar_wait_sleepy(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(bar), "r"(parity), "r"(100000u) : "memory");
  }
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

template <int COLS> __device__ __forceinline__ void tmem_alloc(uint32_t dst_smem) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "n"(COLS) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int COLS> __device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(COLS) : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc),
      "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cvt_tf32(uint32_t x) {
  uint32_t y;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(y) : "f"(__uint_as_float(x)));
  return y;
}

// K-major SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor, sm100 version 1):
//   [0,14) start>>4 | [16,30) LBO>>4 (=1, unused for swizzled K-major) | [32,46) SBO>>4 (1024 B between
//   8-row groups) | [46,48) version=1 | [61,64) layout=2 (SWIZZLE_128B)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
// instruction descriptor (cute::UMMA::InstrDescriptor): D=F32, A=B=TF32, both K-major, N>>3, M>>4
__host__ __device__ constexpr uint32_t make_idesc_tf32(int m, int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

__device__ __forceinline__ float4 ld_stream(const float* p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}

// Epilogue for one warp-owned 32x32 accumulator chunk already staged in shared memory (row-major, pitch
// EPI_PITCH floats).  Lane (rr = lane/8, cj = lane%8) owns rows rr+4*it (it = 0..7) x channels 4*cj..4*cj+3, so
// every global access of the warp is 4 rows x 128 contiguous bytes.  Steps are the OUTER loop (not unrolled:
// code size), rows the inner one: a per-channel operand is fetched once per step, a full-tensor operand
// (residual) as 8 independent 128-bit loads in flight per thread.
template <int PITCH>
__device__ __forceinline__ void epilogue_chunk(const b2j_epilogue& e, const EpiPtrs& epi, const float* stg, float* __restrict__ out,
                                               uint32_t m_base, uint32_t M, uint32_t n, uint32_t ldo, int lane) {
  const int cj = lane & 7, rr = lane >> 3;
  float4 v[8];
#pragma unroll
  for (int it = 0; it < 8; ++it) v[it] = *reinterpret_cast<const float4*>(stg + (rr + 4 * it) * PITCH + 4 * cj);
#pragma unroll 1
  for (uint32_t s = 0; s < e.n_steps; ++s) {
    const b2j_epi_step st = e.steps[s];
    float4 b[8];
    if (st.kind == B2J_EPK_FULL) {
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        const uint32_t m = m_base + rr + 4 * it;
        b[it] = m < M ? ld_stream(epi.p[s] + (uint64_t)m * ldo + n) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    } else {
      float4 t;
      if (st.kind == B2J_EPK_IMM) { const float f = __uint_as_float(st.imm); t = make_float4(f, f, f, f); }
      else t = __ldg(reinterpret_cast<const float4*>(epi.p[s] + n));
#pragma unroll
      for (int it = 0; it < 8; ++it) b[it] = t;
    }
    // operand-on-the-left only matters for the non-commutative ops: use reversed variants instead of swapping registers
    uint32_t opc = st.op;
    if (st.flags & B2J_STEP_SWAP) opc = opc == B2J_OP_SUB_F ? 0x1001u : (opc == B2J_OP_DIV_F ? 0x1002u : opc);
#define B2J_EPI_CASE(OPC, EXPR)                                                                         \
      case OPC:                                                                                         \
        _Pragma("unroll") for (int it = 0; it < 8; ++it) {                                              \
          float4& a = v[it]; const float4 c = b[it];                                                    \
          a.x = EXPR(a.x, c.x); a.y = EXPR(a.y, c.y); a.z = EXPR(a.z, c.z); a.w = EXPR(a.w, c.w);       \
        } break;
#define B2J_MAXF(x, y) epi_op(B2J_OP_MAX_F, x, y)
#define B2J_MINF(x, y) epi_op(B2J_OP_MIN_F, x, y)
#define B2J_RSUB(x, y) __fsub_rn(y, x)
#define B2J_RDIV(x, y) __fdiv_rn(y, x)
    switch (opc) {
      B2J_EPI_CASE(B2J_OP_ADD_F, __fadd_rn)
      B2J_EPI_CASE(B2J_OP_SUB_F, __fsub_rn)
      B2J_EPI_CASE(B2J_OP_MUL_F, __fmul_rn)
      B2J_EPI_CASE(B2J_OP_DIV_F, __fdiv_rn)
      B2J_EPI_CASE(B2J_OP_MAX_F, B2J_MAXF)
      B2J_EPI_CASE(B2J_OP_MIN_F, B2J_MINF)
      B2J_EPI_CASE(0x1001u, B2J_RSUB)
      B2J_EPI_CASE(0x1002u, B2J_RDIV)
      default: break;
    }
#undef B2J_EPI_CASE
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

template <int BLOCK_N, bool X3> struct TcCfg {
  static constexpr int B_TILE_BYTES = BLOCK_N * TC_BLOCK_K * 4;
  static constexpr int STAGE_BYTES = (TC_A_TILE_BYTES + B_TILE_BYTES) * (X3 ? 2 : 1);
  static constexpr int STAGES = X3 ? (BLOCK_N <= 64 ? 4 : 3) : (BLOCK_N <= 64 ? 4 : 3);
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;
  static constexpr int B_ROWS_PER_THREAD = BLOCK_N / 32;
  static constexpr int EPI_PITCH = 36;   // floats; 144 B keeps float4 alignment and is bank-conflict free
  static_assert(TC_PRODUCER_WARPS * 32 * EPI_PITCH * 4 <= STAGES * STAGE_BYTES, "epilogue staging must fit");
};

struct TcRow {         // per-thread im2col row state (fixed for the whole K loop)
  const float* base;   // &x[n, 0, 0, 0]
  int ih0, iw0;        // oh*stride - pad, ow*stride - pad
  bool valid;
};

template <int BLOCK_N, bool X3, bool CVEC>
__global__ void __launch_bounds__(TC_THREADS, (X3 || BLOCK_N > 128) ? 1 : 2)
conv_tc_kernel(const __grid_constant__ b2j_conv_tc_params p, const __grid_constant__ EpiPtrs epi, float* __restrict__ out,
               const float* __restrict__ x, const float* __restrict__ wt_hi, const float* __restrict__ wt_lo) {
  using Cfg = TcCfg<BLOCK_N, X3>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t bar_base = smem_base + Cfg::STAGES * Cfg::STAGE_BYTES;     // full[S], empty[S], accum, tmem slot
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (Cfg::STAGES + s); };
  const uint32_t accum_bar = bar_base + 8u * (2 * Cfg::STAGES);
  const uint32_t tmem_slot = accum_bar + 8u;
  volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + Cfg::STAGES * Cfg::STAGE_BYTES + 8 * (2 * Cfg::STAGES) + 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t M = p.batch * p.oh * p.ow;
  const uint32_t K = p.kh * p.kw * p.c;
  const uint32_t num_kb = p.kpad / TC_BLOCK_K;
  const uint32_t m0 = blockIdx.y * TC_BLOCK_M;
  const uint32_t n0 = blockIdx.x * BLOCK_N;

  if (threadIdx.x == 0) {
    for (int s = 0; s < Cfg::STAGES; ++s) {
      mbar_init(full_bar(s), TC_PRODUCER_WARPS);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(accum_bar, 1);
    fence_barrier_init();
  }
  if (warp == TC_PRODUCER_WARPS) tmem_alloc<BLOCK_N>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_gen;

  if (warp < TC_PRODUCER_WARPS) {
    // =============================== producers ===============================
    const int t = threadIdx.x;            // 0..255
    const int chunk = t & 7;              // 16-byte chunk within the 128-byte k row
    const int row0 = t >> 3;              // 0..31; rows row0 + 32*i
    TcRow rows[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const uint32_t m = m0 + row0 + 32 * i;
      rows[i].valid = m < M;
      const uint32_t mm = rows[i].valid ? m : 0;
      const uint32_t ow = mm % p.ow, t1 = mm / p.ow;
      const uint32_t oh = t1 % p.oh, n = t1 / p.oh;
      rows[i].base = x + (uint64_t)n * p.h * p.w * p.c;
      rows[i].ih0 = (int)(oh * p.stride_h) - p.pad_h;
      rows[i].iw0 = (int)(ow * p.stride_w) - p.pad_w;
    }
    const float* wrow_hi[Cfg::B_ROWS_PER_THREAD];
    bool wvalid[Cfg::B_ROWS_PER_THREAD];
#pragma unroll
    for (int i = 0; i < Cfg::B_ROWS_PER_THREAD; ++i) {
      const uint32_t n = n0 + row0 + 32 * i;
      wvalid[i] = n < p.o;
      wrow_hi[i] = wt_hi + (uint64_t)(wvalid[i] ? n : 0) * p.kpad + chunk * 4;
    }
    const int64_t lo_delta = X3 ? (wt_lo - wt_hi) : 0;

    uint4 ra[4], rb[Cfg::B_ROWS_PER_THREAD], rbl[X3 ? Cfg::B_ROWS_PER_THREAD : 1];

    auto load_kb = [&](uint32_t kb) {
      const uint32_t k = kb * TC_BLOCK_K + chunk * 4;
      if (CVEC) {
        // the 4 floats of this chunk share one filter tap (C % 4 == 0)
        const uint32_t tap = k / p.c, c = k - tap * p.c;
        const uint32_t kh = tap / p.kw, kw = tap - kh * p.kw;
        const int dh = (int)(kh * p.dil_h), dw = (int)(kw * p.dil_w);
        const bool kvalid = k < K;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int ih = rows[i].ih0 + dh, iw = rows[i].iw0 + dw;
          const bool ok = kvalid && rows[i].valid && ih >= 0 && ih < (int)p.h && iw >= 0 && iw < (int)p.w;
          ra[i] = ok ? __ldg(reinterpret_cast<const uint4*>(rows[i].base + ((uint64_t)ih * p.w + iw) * p.c + c))
                     : make_uint4(0u, 0u, 0u, 0u);
        }
      } else {
        // generic channel counts (e.g. the 3-channel stem): element-wise gather
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          uint32_t v[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const uint32_t ke = k + e;
            const uint32_t tap = ke / p.c, c = ke - tap * p.c;
            const uint32_t kh = tap / p.kw, kw = tap - kh * p.kw;
            const int ih = rows[i].ih0 + (int)(kh * p.dil_h), iw = rows[i].iw0 + (int)(kw * p.dil_w);
            const bool ok = ke < K && rows[i].valid && ih >= 0 && ih < (int)p.h && iw >= 0 && iw < (int)p.w;
            v[e] = ok ? __float_as_uint(__ldg(rows[i].base + ((uint64_t)ih * p.w + iw) * p.c + c)) : 0u;
          }
          ra[i] = make_uint4(v[0], v[1], v[2], v[3]);
        }
      }
#pragma unroll
      for (int i = 0; i < Cfg::B_ROWS_PER_THREAD; ++i) {
        const float* src = wrow_hi[i] + kb * TC_BLOCK_K;
        rb[i] = wvalid[i] ? __ldg(reinterpret_cast<const uint4*>(src)) : make_uint4(0u, 0u, 0u, 0u);
        if (X3) rbl[i] = wvalid[i] ? __ldg(reinterpret_cast<const uint4*>(src + lo_delta)) : make_uint4(0u, 0u, 0u, 0u);
      }
    };

    auto split_hi = [](uint4 v) { return make_uint4(cvt_tf32(v.x), cvt_tf32(v.y), cvt_tf32(v.z), cvt_tf32(v.w)); };
    auto split_lo = [](uint4 v, uint4 h) {
      return make_uint4(__float_as_uint(__uint_as_float(v.x) - __uint_as_float(h.x)), __float_as_uint(__uint_as_float(v.y) - __uint_as_float(h.y)),
                        __float_as_uint(__uint_as_float(v.z) - __uint_as_float(h.z)), __float_as_uint(__uint_as_float(v.w) - __uint_as_float(h.w)));
    };
    auto round4 = [](uint4 v) { return make_uint4(cvt_tf32(v.x), cvt_tf32(v.y), cvt_tf32(v.z), cvt_tf32(v.w)); };

    load_kb(0);
    for (uint32_t kb = 0; kb < num_kb; ++kb) {
      const int s = kb % Cfg::STAGES;
      const uint32_t ph = (kb / Cfg::STAGES) & 1u;
      mbar_wait(empty_bar(s), ph ^ 1u);
      uint8_t* stage = smem_gen + s * Cfg::STAGE_BYTES;
      uint8_t* a_hi = stage;
      uint8_t* b_hi = stage + TC_A_TILE_BYTES;
      uint8_t* a_lo = stage + TC_A_TILE_BYTES + Cfg::B_TILE_BYTES;
      uint8_t* b_lo = a_lo + TC_A_TILE_BYTES;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int r = row0 + 32 * i;
        const uint32_t off = r * 128 + ((chunk ^ (r & 7)) << 4);
        if (X3) {
          const uint4 h = split_hi(ra[i]);
          *reinterpret_cast<uint4*>(a_hi + off) = h;
          *reinterpret_cast<uint4*>(a_lo + off) = split_lo(ra[i], h);
        } else {
          *reinterpret_cast<uint4*>(a_hi + off) = round4(ra[i]);
        }
      }
#pragma unroll
      for (int i = 0; i < Cfg::B_ROWS_PER_THREAD; ++i) {
        const int r = row0 + 32 * i;
        const uint32_t off = r * 128 + ((chunk ^ (r & 7)) << 4);
        if (X3) {
          *reinterpret_cast<uint4*>(b_hi + off) = rb[i];
          *reinterpret_cast<uint4*>(b_lo + off) = rbl[i];
        } else {
          *reinterpret_cast<uint4*>(b_hi + off) = round4(rb[i]);
        }
      }
      fence_proxy_async();                         // generic-proxy stores -> visible to the async proxy (UMMA)
      __syncwarp();
      if (lane == 0) mbar_arrive(full_bar(s));
      if (kb + 1 < num_kb) load_kb(kb + 1);       // next k-block's loads fly while the tensor core works
    }

    // =============================== epilogue ===============================
    mbar_wait(accum_bar, 0);
    tc_fence_after();
    const int q = warp & 3;                     // TMEM lane quarter this warp may access
    const int half = warp >> 2;                 // column half
    float* stg = reinterpret_cast<float*>(smem_gen) + warp * 32 * Cfg::EPI_PITCH;   // pipeline smem is idle now
    constexpr int COLS_PER_WARP = BLOCK_N / 2;
#pragma unroll 1
    for (int cc = 0; cc < COLS_PER_WARP; cc += 32) {
      const int col0 = half * COLS_PER_WARP + cc;
      uint32_t r[32];
      tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)col0, r);
      // row = lane: write 32 consecutive columns
#pragma unroll
      for (int j = 0; j < 8; ++j)
        *reinterpret_cast<uint4*>(stg + lane * Cfg::EPI_PITCH + 4 * j) = make_uint4(r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
      __syncwarp();
      const uint32_t n = n0 + col0 + 4 * (lane & 7);
      if (n < p.o) epilogue_chunk<Cfg::EPI_PITCH>(p.epi, epi, stg, out, m0 + q * 32, M, n, p.o, lane);
      __syncwarp();
    }
  } else if (lane == 0) {
    // =============================== MMA issuer (one thread) ===============================
    constexpr uint32_t idesc = make_idesc_tf32(TC_BLOCK_M, BLOCK_N);
    for (uint32_t kb = 0; kb < num_kb; ++kb) {
      const int s = kb % Cfg::STAGES;
      const uint32_t ph = (kb / Cfg::STAGES) & 1u;
      mbar_wait(full_bar(s), ph);
      tc_fence_after();
      const uint32_t stage = smem_base + s * Cfg::STAGE_BYTES;
      const uint64_t a_hi = make_smem_desc(stage);
      const uint64_t b_hi = make_smem_desc(stage + TC_A_TILE_BYTES);
      const uint64_t a_lo = make_smem_desc(stage + TC_A_TILE_BYTES + Cfg::B_TILE_BYTES);
      const uint64_t b_lo = make_smem_desc(stage + 2 * TC_A_TILE_BYTES + Cfg::B_TILE_BYTES);
#pragma unroll
      for (int k = 0; k < TC_BLOCK_K / 8; ++k) {
        const uint64_t adv = (uint64_t)(k * 2);       // 8 tf32 = 32 bytes = 2 x 16-byte units along K
        if (X3) {
          umma_tf32(tmem_base, a_lo + adv, b_hi + adv, idesc, (kb | (uint32_t)k) != 0u);
          umma_tf32(tmem_base, a_hi + adv, b_lo + adv, idesc, 1u);
          umma_tf32(tmem_base, a_hi + adv, b_hi + adv, idesc, 1u);
        } else {
          umma_tf32(tmem_base, a_hi + adv, b_hi + adv, idesc, (kb | (uint32_t)k) != 0u);
        }
      }
      umma_commit(empty_bar(s));        // implicit tcgen05.fence::before_thread_sync; frees the stage when the MMAs retire
    }
    umma_commit(accum_bar);             // accumulator complete -> epilogue
  }

  tc_fence_before();
  __syncthreads();
  if (warp == TC_PRODUCER_WARPS) {
    tc_fence_after();
    tmem_dealloc<BLOCK_N>(tmem_base);
  }
}

// ---- host-side launcher ------------------------------------------------------------------------
template <int BLOCK_N, bool X3, bool CVEC>
static int launch_conv_tc_inst(const b2j_conv_tc_params& p, const EpiPtrs& epi, float* out, const float* x, const float* wt_hi,
                               const float* wt_lo, cudaStream_t st, const char** why) {
  using Cfg = TcCfg<BLOCK_N, X3>;
  static bool configured = false;
  auto kern = conv_tc_kernel<BLOCK_N, X3, CVEC>;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
    if (e != cudaSuccess) { *why = cudaGetErrorString(e); return B2J_ECUDA; }
    configured = true;
  }
  const uint32_t M = p.batch * p.oh * p.ow;
  dim3 grid((p.o + BLOCK_N - 1) / BLOCK_N, (M + TC_BLOCK_M - 1) / TC_BLOCK_M);
  if (grid.y > 65535u) { *why = "M too large for grid.y"; return B2J_ENOTIMPL; }
  kern<<<grid, TC_THREADS, Cfg::SMEM_BYTES, st>>>(p, epi, out, x, wt_hi, wt_lo);
  return B2J_OK;
}

static int launch_conv_tc(const b2j_conv_tc_params& p, const EpiPtrs& epi, float* out, const float* x, const float* wt_hi,
                          const float* wt_lo, int sm_count, cudaStream_t st, const char** why) {
  (void)sm_count;
  if (p.o % 4 != 0) { *why = "O must be a multiple of 4"; return B2J_ENOTIMPL; }
  if (p.kpad % TC_BLOCK_K != 0 || p.kpad < p.kh * p.kw * p.c) { *why = "kpad must be K rounded up to 32"; return B2J_EINVAL; }
  const bool x3 = p.precision == B2J_PREC_TF32X3;
  if (x3 && !wt_lo) { *why = "3xTF32 needs the wt_lo buffer"; return B2J_EINVAL; }
  const bool cvec = (p.c % 4) == 0;
  const bool small_n = p.o <= 64;
#define TC_DISPATCH(BN, X3_, CV) return launch_conv_tc_inst<BN, X3_, CV>(p, epi, out, x, wt_hi, wt_lo, st, why)
  if (small_n) {
    if (x3) { if (cvec) TC_DISPATCH(64, true, true); else TC_DISPATCH(64, true, false); }
    else    { if (cvec) TC_DISPATCH(64, false, true); else TC_DISPATCH(64, false, false); }
  } else {
    if (x3) { if (cvec) TC_DISPATCH(128, true, true); else TC_DISPATCH(128, true, false); }
    else    { if (cvec) TC_DISPATCH(128, false, true); else TC_DISPATCH(128, false, false); }
  }
#undef TC_DISPATCH
}

}  // namespace b2j
