// Fused elementwise-chain kernel (B2J_K_ELTWISE).
//
// Replaces the reference's one-dispatch-per-primitive elementwise shaders (add.comp ... unary_op.comp,
// select.comp, convert_element_type.comp, integer_pow.comp, iota.comp) *and* the broadcast
// materialisation the reference performs for every mismatched operand (reference ops.py:84-86,160-184):
// broadcast operands are read through index math (MOD / DIV / STRIDED kinds) instead of being written
// out first.  HBM-bound: each thread owns 4 consecutive elements -> 128-bit loads/stores for FULL
// operands; a chain of up to 16 ops is applied in registers between one load and one store.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include "../../include/b2jax.h"

namespace b2j {

struct EltPtrs {
  uint32_t* out;
  const uint32_t* in[B2J_ELT_MAX_IN];
};

__device__ __forceinline__ float u2f(uint32_t v) { return __uint_as_float(v); }
__device__ __forceinline__ uint32_t f2u(float v) { return __float_as_uint(v); }

__device__ __forceinline__ float ipow_f(float a, int y) {
  // exponentiation by squaring; negative exponents via reciprocal (lax.integer_pow semantics)
  unsigned e = y < 0 ? (unsigned)(-(long long)y) : (unsigned)y;
  float r = 1.0f, b = a;
  while (e) { if (e & 1u) r *= b; b *= b; e >>= 1; }
  return y < 0 ? 1.0f / r : r;
}
__device__ __forceinline__ int ipow_i(int a, int y) {
  unsigned e = (unsigned)(y < 0 ? 0 : y);
  int r = 1, b = a;
  while (e) { if (e & 1u) r *= b; b *= b; e >>= 1; }
  return r;
}

// One chain step on one element.  `a` is the accumulator side, `b` the operand side (already swapped).
__device__ __forceinline__ uint32_t elt_apply(uint32_t op, uint32_t a, uint32_t b, uint32_t c, uint32_t imm) {
  const float fa = u2f(a), fb = u2f(b);
  const int ia = (int)a, ib = (int)b;
  switch (op) {
    case B2J_OP_NOP: return a;
    case B2J_OP_ADD_F: return f2u(__fadd_rn(fa, fb));
    case B2J_OP_SUB_F: return f2u(__fsub_rn(fa, fb));
    case B2J_OP_MUL_F: return f2u(__fmul_rn(fa, fb));
    case B2J_OP_DIV_F: return f2u(__fdiv_rn(fa, fb));
    case B2J_OP_MAX_F: return f2u((fa != fa || fb != fb) ? __int_as_float(0x7fc00000) : fmaxf(fa, fb));
    case B2J_OP_MIN_F: return f2u((fa != fa || fb != fb) ? __int_as_float(0x7fc00000) : fminf(fa, fb));
    case B2J_OP_POW_F: return f2u(powf(fa, fb));
    case B2J_OP_REM_F: return f2u(fmodf(fa, fb));
    case B2J_OP_NEXTAFTER_F: return f2u(nextafterf(fa, fb));
    case B2J_OP_ATAN2_F: return f2u(atan2f(fa, fb));
    case B2J_OP_ADD_I: return a + b;
    case B2J_OP_SUB_I: return a - b;
    case B2J_OP_MUL_I: return a * b;
    case B2J_OP_DIV_I: return ib == 0 ? 0xFFFFFFFFu : ((ia == INT_MIN && ib == -1) ? (uint32_t)INT_MIN : (uint32_t)(ia / ib));
    case B2J_OP_DIV_U: return b == 0 ? 0xFFFFFFFFu : a / b;
    case B2J_OP_MAX_I: return (uint32_t)max(ia, ib);
    case B2J_OP_MAX_U: return max(a, b);
    case B2J_OP_MIN_I: return (uint32_t)min(ia, ib);
    case B2J_OP_MIN_U: return min(a, b);
    case B2J_OP_REM_I: return ib == 0 ? a : ((ia == INT_MIN && ib == -1) ? 0u : (uint32_t)(ia % ib));
    case B2J_OP_REM_U: return b == 0 ? a : a % b;
    case B2J_OP_AND: return a & b;
    case B2J_OP_OR: return a | b;
    case B2J_OP_XOR: return a ^ b;
    case B2J_OP_SHL: return b >= 32u ? 0u : a << b;
    case B2J_OP_SHR_L: return b >= 32u ? 0u : a >> b;
    case B2J_OP_SHR_A: return (uint32_t)(ia >> (b >= 32u ? 31u : b));
    case B2J_OP_GT_F: return fa > fb;
    case B2J_OP_GE_F: return fa >= fb;
    case B2J_OP_LT_F: return fa < fb;
    case B2J_OP_LE_F: return fa <= fb;
    case B2J_OP_EQ_F: return fa == fb;
    case B2J_OP_NE_F: return fa != fb;
    case B2J_OP_GT_I: return ia > ib;
    case B2J_OP_GE_I: return ia >= ib;
    case B2J_OP_LT_I: return ia < ib;
    case B2J_OP_LE_I: return ia <= ib;
    case B2J_OP_EQ_I: return a == b;
    case B2J_OP_NE_I: return a != b;
    case B2J_OP_GT_U: return a > b;
    case B2J_OP_GE_U: return a >= b;
    case B2J_OP_LT_U: return a < b;
    case B2J_OP_LE_U: return a <= b;
    case B2J_OP_EXP: return f2u(expf(fa));
    case B2J_OP_LOG: return f2u(logf(fa));
    case B2J_OP_NEG_F: return a ^ 0x80000000u;
    case B2J_OP_NEG_I: return (uint32_t)(-ia);
    case B2J_OP_ABS_F: return a & 0x7fffffffu;
    case B2J_OP_ABS_I: return (uint32_t)(ia < 0 ? -ia : ia);
    case B2J_OP_RSQRT: return f2u(__fdiv_rn(1.0f, __fsqrt_rn(fa)));   // reference rsqrt.comp:11: 1.0/sqrt(x)
    case B2J_OP_SQRT: return f2u(__fsqrt_rn(fa));
    case B2J_OP_ERF: return f2u(erff(fa));
    case B2J_OP_ERF_INV: return f2u(erfinvf(fa));
    case B2J_OP_ERFC: return f2u(erfcf(fa));
    case B2J_OP_COS: return f2u(cosf(fa));
    case B2J_OP_SIN: return f2u(sinf(fa));
    case B2J_OP_TAN: return f2u(tanf(fa));
    case B2J_OP_COSH: return f2u(coshf(fa));
    case B2J_OP_SINH: return f2u(sinhf(fa));
    case B2J_OP_TANH: return f2u(tanhf(fa));
    case B2J_OP_ACOS: return f2u(acosf(fa));
    case B2J_OP_ASIN: return f2u(asinf(fa));
    case B2J_OP_ATAN: return f2u(atanf(fa));
    case B2J_OP_ACOSH: return f2u(acoshf(fa));
    case B2J_OP_ASINH: return f2u(asinhf(fa));
    case B2J_OP_ATANH: return f2u(atanhf(fa));
    case B2J_OP_CEIL: return f2u(ceilf(fa));
    case B2J_OP_FLOOR: return f2u(floorf(fa));
    case B2J_OP_ROUND: return f2u(roundf(fa));
    case B2J_OP_SIGN_F: return f2u(fa != fa ? fa : (fa > 0.0f ? 1.0f : (fa < 0.0f ? -1.0f : fa)));
    case B2J_OP_SIGN_I: return (uint32_t)((ia > 0) - (ia < 0));
    case B2J_OP_LOG1P: return f2u(log1pf(fa));
    case B2J_OP_EXPM1: return f2u(expm1f(fa));
    case B2J_OP_LOGISTIC: return f2u(__fdiv_rn(1.0f, 1.0f + expf(-fa)));
    case B2J_OP_NOT_BITS: return ~a;
    case B2J_OP_NOT_BOOL: return a == 0u;
    case B2J_OP_IPOW_F: return f2u(ipow_f(fa, (int)imm));
    case B2J_OP_IPOW_I: return (uint32_t)ipow_i(ia, (int)imm);
    case B2J_OP_CVT_F2I: return (uint32_t)(int)fa;
    case B2J_OP_CVT_F2U: return (uint32_t)fa;
    case B2J_OP_CVT_I2F: return f2u((float)ia);
    case B2J_OP_CVT_U2F: return f2u((float)a);
    case B2J_OP_CVT_TOBOOL_F: return fa != 0.0f;
    case B2J_OP_CVT_TOBOOL_I: return a != 0u;
    case B2J_OP_SELECT: return a ? b : c;     // true select (the reference blends arithmetically: quirk Q4)
    default: return a;
  }
}

__device__ __noinline__ uint32_t elt_apply_slow(uint32_t op, uint32_t a, uint32_t b, uint32_t c, uint32_t imm) {
  return elt_apply(op, a, b, c, imm);
}

// One chain step over the N elements a thread holds.  The opcode switch runs ONCE per step and thread (not once per
// element, as in the first version, whose per-element call of a 90-way switch held the kernel at 15-40 % of the HBM
// bandwidth: profiles/r02_bandwidth_kernels.md); inside a case the opcode is a compile-time constant, so the element loop is
// straight-line code.  Cheap operations are unrolled over all N elements; the libm-heavy ones loop over a shared call.
// BW = N: one operand value per element; BW = 4 ("narrow" operands: immediates, scalars, per-channel vectors whose period divides
// the distance between a thread's vectors): the same 4 operand values serve every vector of the thread.
template <int N, int BW>
__device__ __forceinline__ void elt_step(uint32_t op, bool swap, uint32_t imm, uint32_t (&acc)[N], const uint32_t (&b)[BW]) {
#define ELT_FAST(OPC)                                                                                  \
  case OPC:                                                                                            \
    _Pragma("unroll") for (int j = 0; j < N; ++j)                                                      \
      acc[j] = elt_apply(OPC, swap ? b[j % BW] : acc[j], swap ? acc[j] : b[j % BW], 0u, imm);          \
    break;
  switch (op) {
    case B2J_OP_NOP: break;
    ELT_FAST(B2J_OP_ADD_F) ELT_FAST(B2J_OP_SUB_F) ELT_FAST(B2J_OP_MUL_F) ELT_FAST(B2J_OP_DIV_F) ELT_FAST(B2J_OP_MAX_F) ELT_FAST(B2J_OP_MIN_F)
    ELT_FAST(B2J_OP_ADD_I) ELT_FAST(B2J_OP_SUB_I) ELT_FAST(B2J_OP_MUL_I) ELT_FAST(B2J_OP_MAX_I) ELT_FAST(B2J_OP_MAX_U) ELT_FAST(B2J_OP_MIN_I) ELT_FAST(B2J_OP_MIN_U)
    ELT_FAST(B2J_OP_AND) ELT_FAST(B2J_OP_OR) ELT_FAST(B2J_OP_XOR) ELT_FAST(B2J_OP_SHL) ELT_FAST(B2J_OP_SHR_L) ELT_FAST(B2J_OP_SHR_A)
    ELT_FAST(B2J_OP_GT_F) ELT_FAST(B2J_OP_GE_F) ELT_FAST(B2J_OP_LT_F) ELT_FAST(B2J_OP_LE_F) ELT_FAST(B2J_OP_EQ_F) ELT_FAST(B2J_OP_NE_F)
    ELT_FAST(B2J_OP_GT_I) ELT_FAST(B2J_OP_GE_I) ELT_FAST(B2J_OP_LT_I) ELT_FAST(B2J_OP_LE_I) ELT_FAST(B2J_OP_EQ_I) ELT_FAST(B2J_OP_NE_I)
    ELT_FAST(B2J_OP_GT_U) ELT_FAST(B2J_OP_GE_U) ELT_FAST(B2J_OP_LT_U) ELT_FAST(B2J_OP_LE_U)
    ELT_FAST(B2J_OP_NEG_F) ELT_FAST(B2J_OP_NEG_I) ELT_FAST(B2J_OP_ABS_F) ELT_FAST(B2J_OP_ABS_I) ELT_FAST(B2J_OP_RSQRT) ELT_FAST(B2J_OP_SQRT)
    ELT_FAST(B2J_OP_CEIL) ELT_FAST(B2J_OP_FLOOR) ELT_FAST(B2J_OP_SIGN_F) ELT_FAST(B2J_OP_SIGN_I) ELT_FAST(B2J_OP_NOT_BITS) ELT_FAST(B2J_OP_NOT_BOOL)
    ELT_FAST(B2J_OP_CVT_F2I) ELT_FAST(B2J_OP_CVT_F2U) ELT_FAST(B2J_OP_CVT_I2F) ELT_FAST(B2J_OP_CVT_U2F) ELT_FAST(B2J_OP_CVT_TOBOOL_F) ELT_FAST(B2J_OP_CVT_TOBOOL_I)
    ELT_FAST(B2J_OP_EXP) ELT_FAST(B2J_OP_LOG) ELT_FAST(B2J_OP_TANH) ELT_FAST(B2J_OP_LOGISTIC)
    default:      // fully unrolled as well: a rolled loop would index acc[] dynamically and push it to local memory
#pragma unroll
      for (int j = 0; j < N; ++j) acc[j] = elt_apply_slow(op, swap ? b[j % BW] : acc[j], swap ? acc[j] : b[j % BW], 0u, imm);
      break;
  }
#undef ELT_FAST
}

__device__ __forceinline__ uint64_t strided_index(uint64_t i, const b2j_elt_params& p, const uint32_t* strides) {
  uint64_t idx = 0;
#pragma unroll 1
  for (int d = (int)p.rank - 1; d >= 0; --d) {
    const uint32_t s = p.shape[d];
    const uint64_t q = i / s;
    idx += (i - q * s) * (uint64_t)strides[d];
    i = q;
  }
  return idx;
}

constexpr int ELT_VECS = 4;          // 128-bit vectors per thread and tile: 4 independent loads in flight per operand
constexpr int ELT_VECS_NARROW = 8;   // chains whose step operands are all narrow (see elt_step) keep 8 in flight: acc[32] + b[4] registers
constexpr int ELT_THREADS = 256;
// Software prefetch of the next tile's first operand: measured SLOWER on B200 (BatchNorm chain 2.8 -> 2.3 TB/s, add + max
// 5.7 -> 4.8 TB/s): its 16 extra registers cost the third resident CTA per SM, and thread-level parallelism hides the DRAM
// latency better than a deeper per-thread pipeline.  Kept behind a switch for the record.
#ifndef B2J_ELT_PREFETCH
#define B2J_ELT_PREFETCH 0
#endif

// Loads one operand for the ELT_VECS vectors of this thread: vector v covers output elements [4*(vi0 + v*ELT_THREADS), +4).
// `ok[v]`: the whole vector is in range (otherwise it is loaded element-wise with a tail guard).
template <int ELT_VECS>
__device__ __forceinline__ void elt_load(const b2j_elt_params& p, const EltPtrs& ptrs, uint32_t slot, uint32_t imm,
                                         uint64_t vi0, const bool (&ok)[ELT_VECS], uint32_t (&v)[4 * ELT_VECS]) {
  if (slot == B2J_SRC_IMM) {
#pragma unroll
    for (int j = 0; j < 4 * ELT_VECS; ++j) v[j] = imm;
    return;
  }
  if (slot == B2J_SRC_IOTA) {
    // value = coordinate of element i along dimension `imm`
    uint64_t inner = 1;
    for (int d = (int)p.rank - 1; d > (int)imm; --d) inner *= p.shape[d];
    const uint32_t s = p.shape[imm];
#pragma unroll
    for (int k = 0; k < ELT_VECS; ++k)
#pragma unroll
      for (int j = 0; j < 4; ++j) v[4 * k + j] = (uint32_t)((((vi0 + (uint64_t)k * ELT_THREADS) * 4 + j) / inner) % s);
    return;
  }
  const b2j_elt_operand& o = p.in[slot];
  const uint32_t* __restrict__ src = ptrs.in[slot];
  if (o.elem == 1) {
    // packed uint8 source (images uploaded as bytes): zero-extend on load.  FULL: the 4 elements of a vector are ONE aligned
    // 32-bit word (a warp reads 128 contiguous bytes); broadcast kinds fall back to byte loads.
    const uint8_t* __restrict__ s8 = reinterpret_cast<const uint8_t*>(src);
#pragma unroll
    for (int k = 0; k < ELT_VECS; ++k) {
      const uint64_t i0 = (vi0 + (uint64_t)k * ELT_THREADS) * 4;
      if (o.kind == B2J_OPK_FULL && ok[k]) {
        const uint32_t w = __ldg(src + (i0 >> 2));
        v[4 * k] = w & 0xFFu; v[4 * k + 1] = (w >> 8) & 0xFFu; v[4 * k + 2] = (w >> 16) & 0xFFu; v[4 * k + 3] = w >> 24;
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint64_t i = min(i0 + j, p.n - 1);
          const uint64_t idx = o.kind == B2J_OPK_FULL ? i : o.kind == B2J_OPK_SCALAR ? 0 : o.kind == B2J_OPK_MOD ? i % o.mod
                             : o.kind == B2J_OPK_DIV ? i / o.mod : strided_index(i, p, o.strides);
          v[4 * k + j] = __ldg(s8 + idx);
        }
      }
    }
    return;
  }
  const bool idx32 = p.n <= 0xFFFFFFFFull;          // 32-bit index math whenever the tensor allows it (one IDIV instead of a 64-bit division sequence)
  switch (o.kind) {
    case B2J_OPK_FULL:
#pragma unroll
      for (int k = 0; k < ELT_VECS; ++k) {
        const uint64_t i0 = (vi0 + (uint64_t)k * ELT_THREADS) * 4;
        if (ok[k]) {
          const uint4 t = __ldg(reinterpret_cast<const uint4*>(src + i0));
          v[4 * k] = t.x; v[4 * k + 1] = t.y; v[4 * k + 2] = t.z; v[4 * k + 3] = t.w;
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j) v[4 * k + j] = (i0 + j < p.n) ? __ldg(src + i0 + j) : 0u;
        }
      }
      break;
    case B2J_OPK_SCALAR: {
      const uint32_t t = __ldg(src);
#pragma unroll
      for (int j = 0; j < 4 * ELT_VECS; ++j) v[j] = t;
    } break;
    case B2J_OPK_MOD: {
      const uint32_t m = o.mod;
      const uint64_t i00 = vi0 * 4;
      const bool pow2 = (m & (m - 1u)) == 0u;
      uint32_t r = pow2 ? ((uint32_t)i00 & (m - 1u)) : (idx32 ? (uint32_t)i00 % m : (uint32_t)(i00 % m));
      const uint32_t step = pow2 ? ((4u * ELT_THREADS) & (m - 1u)) : (4u * ELT_THREADS) % m;   // residue advance from one vector of this thread to the next
      if (step == 0u && (m & 3u) == 0u) {
        // per-channel operand with m | 1024 (e.g. 64 ... 1024 channels): all ELT_VECS vectors of this thread see the SAME 4
        // operand values -- one 128-bit load instead of four
        const uint4 t = __ldg(reinterpret_cast<const uint4*>(src + r));
#pragma unroll
        for (int k = 0; k < ELT_VECS; ++k) { v[4 * k] = t.x; v[4 * k + 1] = t.y; v[4 * k + 2] = t.z; v[4 * k + 3] = t.w; }
        break;
      }
#pragma unroll
      for (int k = 0; k < ELT_VECS; ++k) {
        if ((m & 3u) == 0u) {      // i0 % 4 == 0 and m % 4 == 0  ->  r % 4 == 0 and r + 3 < m
          const uint4 t = __ldg(reinterpret_cast<const uint4*>(src + r));
          v[4 * k] = t.x; v[4 * k + 1] = t.y; v[4 * k + 2] = t.z; v[4 * k + 3] = t.w;
        } else {
          uint32_t rr = r;
#pragma unroll
          for (int j = 0; j < 4; ++j) { v[4 * k + j] = __ldg(src + rr); rr = (rr + 1 == m) ? 0u : rr + 1; }
        }
        r += step;
        if (r >= m) r -= m;
      }
    } break;
    case B2J_OPK_DIV: {
#pragma unroll
      for (int k = 0; k < ELT_VECS; ++k)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint64_t i = min((vi0 + (uint64_t)k * ELT_THREADS) * 4 + j, p.n - 1);
          v[4 * k + j] = __ldg(src + (idx32 ? (uint64_t)((uint32_t)i / o.mod) : i / o.mod));
        }
    } break;
    default: {  // STRIDED
#pragma unroll
      for (int k = 0; k < ELT_VECS; ++k)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint64_t i = min((vi0 + (uint64_t)k * ELT_THREADS) * 4 + j, p.n - 1);
          v[4 * k + j] = __ldg(src + strided_index(i, p, o.strides));
        }
    } break;
  }
}

// The 4 operand values a thread's vectors share (narrow operands only; the host checks the conditions: elt_chain_is_narrow)
__device__ __forceinline__ void elt_load_narrow(const b2j_elt_params& p, const EltPtrs& ptrs, uint32_t slot, uint32_t imm, uint64_t vi0, uint32_t (&v)[4]) {
  if (slot == B2J_SRC_IMM) { v[0] = v[1] = v[2] = v[3] = imm; return; }
  const b2j_elt_operand& o = p.in[slot];
  const uint32_t* __restrict__ src = ptrs.in[slot];
  if (o.kind == B2J_OPK_SCALAR) { v[0] = v[1] = v[2] = v[3] = __ldg(src); return; }
  const uint32_t m = o.mod;                                  // MOD with m % 4 == 0 and (4 * ELT_THREADS) % m == 0
  const uint32_t r = (m & (m - 1u)) == 0u ? ((uint32_t)(vi0 * 4) & (m - 1u)) : (uint32_t)((vi0 * 4) % m);
  const uint4 t = __ldg(reinterpret_cast<const uint4*>(src + r));
  v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}

// A tile = ELT_THREADS * ELT_VECS vectors of 4 elements; consecutive threads own consecutive vectors (coalesced 128-bit
// accesses), a thread's ELT_VECS vectors are ELT_THREADS apart.  All loads of an operand are issued before the step's
// arithmetic, and the NEXT tile's first operand is requested before the current tile is processed (software prefetch),
// so every thread keeps ELT_VECS 128-bit requests in flight through its compute / store phase as well: without it a
// 4-step BatchNorm chain spent half its time with no loads outstanding (2.8 TB/s; profiles/r02_bandwidth_kernels.md).
#ifndef B2J_ELT_MIN_CTAS
// 4 CTAs of 256 threads per SM (64 registers): the kernel is bound by the bytes it keeps in flight, not by arithmetic -- with 3
// CTAs (80 registers) the residual add + ReLU chain ran at 0.84 of the copy bandwidth, with 4 at 1.0; BatchNorm chain 0.51 -> 0.61,
// convert + div 0.59 -> 0.72 (profiles/README.md, round 2)
#define B2J_ELT_MIN_CTAS 4
#endif
template <int ELT_VECS, bool NARROW>
__global__ void __launch_bounds__(ELT_THREADS, B2J_ELT_PREFETCH ? 2 : B2J_ELT_MIN_CTAS) eltwise_kernel(const __grid_constant__ b2j_elt_params p,
                                                                const __grid_constant__ EltPtrs ptrs) {
  const uint64_t nvec = (p.n + 3) >> 2;
  const uint64_t tile_vecs = (uint64_t)ELT_THREADS * ELT_VECS;
  const uint64_t ntiles = (nvec + tile_vecs - 1) / tile_vecs;
  const bool prefetch = B2J_ELT_PREFETCH && p.init_src < B2J_ELT_MAX_IN && p.in[p.init_src < B2J_ELT_MAX_IN ? p.init_src : 0].kind == B2J_OPK_FULL &&
                        p.in[p.init_src < B2J_ELT_MAX_IN ? p.init_src : 0].elem == 0;
  auto flags = [&](uint64_t vi0, bool (&ok)[ELT_VECS], bool (&any)[ELT_VECS]) {
#pragma unroll
    for (int k = 0; k < ELT_VECS; ++k) {
      const uint64_t i0 = (vi0 + (uint64_t)k * ELT_THREADS) * 4;
      ok[k] = i0 + 3 < p.n;
      any[k] = i0 < p.n;
    }
  };
  uint32_t nxt[4 * ELT_VECS];
  bool ok[ELT_VECS], any[ELT_VECS];
  if (prefetch && blockIdx.x < ntiles) {
    flags((uint64_t)blockIdx.x * tile_vecs + threadIdx.x, ok, any);
    elt_load<ELT_VECS>(p, ptrs, p.init_src, p.init_imm, (uint64_t)blockIdx.x * tile_vecs + threadIdx.x, ok, nxt);
  }
  for (uint64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const uint64_t vi0 = tile * tile_vecs + threadIdx.x;
    flags(vi0, ok, any);
    uint32_t acc[4 * ELT_VECS], b[NARROW ? 4 : 4 * ELT_VECS];
    if (prefetch) {
#pragma unroll
      for (int j = 0; j < 4 * ELT_VECS; ++j) acc[j] = nxt[j];
      if (tile + gridDim.x < ntiles) {
        bool okn[ELT_VECS], anyn[ELT_VECS];
        flags(vi0 + (uint64_t)gridDim.x * tile_vecs, okn, anyn);
        elt_load<ELT_VECS>(p, ptrs, p.init_src, p.init_imm, vi0 + (uint64_t)gridDim.x * tile_vecs, okn, nxt);
      }
    } else {
      elt_load<ELT_VECS>(p, ptrs, p.init_src, p.init_imm, vi0, ok, acc);
    }
    if (!any[0]) continue;
#pragma unroll 1
    for (uint32_t s = 0; s < p.n_steps; ++s) {
      const b2j_elt_step st = p.steps[s];
      if constexpr (NARROW) {
        if (st.src != B2J_SRC_NONE) elt_load_narrow(p, ptrs, st.src, st.imm, vi0, b);
        elt_step<4 * ELT_VECS, 4>(st.op, (st.flags & B2J_STEP_SWAP) != 0, st.imm, acc, b);
      } else {
        if (st.op == B2J_OP_SELECT) {
          // acc is the predicate: remember it as a bit mask, take on_true, then merge on_false (no third register array)
          uint32_t mask = 0;
#pragma unroll
          for (int j = 0; j < 4 * ELT_VECS; ++j) mask |= (acc[j] != 0u ? 1u : 0u) << j;
          elt_load<ELT_VECS>(p, ptrs, st.src, st.imm, vi0, ok, acc);
          elt_load<ELT_VECS>(p, ptrs, st.src2, st.imm2, vi0, ok, b);
#pragma unroll
          for (int j = 0; j < 4 * ELT_VECS; ++j) acc[j] = ((mask >> j) & 1u) ? acc[j] : b[j];     // true select (reference blends: quirk Q4)
          continue;
        }
        if (st.src != B2J_SRC_NONE) elt_load<ELT_VECS>(p, ptrs, st.src, st.imm, vi0, ok, b);
        elt_step<4 * ELT_VECS, 4 * ELT_VECS>(st.op, (st.flags & B2J_STEP_SWAP) != 0, st.imm, acc, b);
      }
    }
#pragma unroll
    for (int k = 0; k < ELT_VECS; ++k) {
      const uint64_t i0 = (vi0 + (uint64_t)k * ELT_THREADS) * 4;
      if (ok[k]) {
        *reinterpret_cast<uint4*>(ptrs.out + i0) = make_uint4(acc[4 * k], acc[4 * k + 1], acc[4 * k + 2], acc[4 * k + 3]);
      } else if (any[k]) {
#pragma unroll
        for (int j = 0; j < 4; ++j) if (i0 + j < p.n) ptrs.out[i0 + j] = acc[4 * k + j];
      }
    }
  }
}

// Host side: all step operands narrow?  (immediates, scalars, per-channel vectors of period m with m % 4 == 0 and m | 4 * ELT_THREADS;
// no select, no iota operand, 32-bit elements)
static bool elt_chain_is_narrow(const b2j_elt_params& p) {
  for (uint32_t s = 0; s < p.n_steps; ++s) {
    const b2j_elt_step& st = p.steps[s];
    if (st.op == B2J_OP_SELECT) return false;
    if (st.src == B2J_SRC_NONE || st.src == B2J_SRC_IMM) continue;
    if (st.src >= B2J_ELT_MAX_IN) return false;                 // iota
    const b2j_elt_operand& o = p.in[st.src];
    if (o.elem != 0) return false;
    if (o.kind == B2J_OPK_SCALAR) continue;
    if (o.kind == B2J_OPK_MOD && o.mod != 0 && (o.mod & 3u) == 0 && (4u * ELT_THREADS) % o.mod == 0) continue;
    return false;
  }
  return p.n_steps > 0;
}

}  // namespace b2j
