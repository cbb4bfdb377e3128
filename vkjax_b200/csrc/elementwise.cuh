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
__device__ __noinline__ uint32_t elt_apply(uint32_t op, uint32_t a, uint32_t b, uint32_t c, uint32_t imm) {
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

// Loads the 4 operand values for output elements [i0, i0+4).  `full4` = all 4 are in range.
__device__ __forceinline__ void elt_load(const b2j_elt_params& p, const EltPtrs& ptrs, uint32_t slot, uint32_t imm,
                                         uint64_t i0, bool full4, uint32_t v[4]) {
  if (slot == B2J_SRC_IMM) { v[0] = v[1] = v[2] = v[3] = imm; return; }
  if (slot == B2J_SRC_IOTA) {
    // value = coordinate of element i along dimension `imm`
    uint64_t inner = 1;
    for (int d = (int)p.rank - 1; d > (int)imm; --d) inner *= p.shape[d];
    const uint32_t s = p.shape[imm];
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = (uint32_t)(((i0 + j) / inner) % s);
    return;
  }
  const b2j_elt_operand& o = p.in[slot];
  const uint32_t* __restrict__ src = ptrs.in[slot];
  switch (o.kind) {
    case B2J_OPK_FULL:
      if (full4) {
        const uint4 t = __ldg(reinterpret_cast<const uint4*>(src + i0));
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] = (i0 + j < p.n) ? __ldg(src + i0 + j) : 0u;
      }
      break;
    case B2J_OPK_SCALAR: {
      const uint32_t t = __ldg(src);
      v[0] = v[1] = v[2] = v[3] = t;
    } break;
    case B2J_OPK_MOD: {
      const uint32_t m = o.mod;
      const uint32_t r = (uint32_t)(i0 % m);
      if ((m & 3u) == 0u) {      // i0 % 4 == 0 and m % 4 == 0  ->  r % 4 == 0 and r + 3 < m
        const uint4 t = __ldg(reinterpret_cast<const uint4*>(src + r));
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
      } else {
        uint32_t rr = r;
#pragma unroll
        for (int j = 0; j < 4; ++j) { v[j] = __ldg(src + rr); rr = (rr + 1 == m) ? 0u : rr + 1; }
      }
    } break;
    case B2J_OPK_DIV: {
#pragma unroll
      for (int j = 0; j < 4; ++j) { const uint64_t i = min(i0 + j, p.n - 1); v[j] = __ldg(src + i / o.mod); }
    } break;
    default: {  // STRIDED
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint64_t i = min(i0 + j, p.n - 1);
        v[j] = __ldg(src + strided_index(i, p, o.strides));
      }
    } break;
  }
}

__global__ void __launch_bounds__(256) eltwise_kernel(const __grid_constant__ b2j_elt_params p,
                                                      const __grid_constant__ EltPtrs ptrs) {
  const uint64_t nvec = (p.n + 3) >> 2;
  for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < nvec; t += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t i0 = t << 2;
    const bool full4 = i0 + 3 < p.n;
    uint32_t acc[4], b[4], c[4];
    elt_load(p, ptrs, p.init_src, p.init_imm, i0, full4, acc);
    for (uint32_t s = 0; s < p.n_steps; ++s) {
      const b2j_elt_step st = p.steps[s];
      if (st.src != B2J_SRC_NONE) elt_load(p, ptrs, st.src, st.imm, i0, full4, b);
      else { b[0] = b[1] = b[2] = b[3] = 0u; }
      if (st.op == B2J_OP_SELECT) elt_load(p, ptrs, st.src2, st.imm2, i0, full4, c);
      else { c[0] = c[1] = c[2] = c[3] = 0u; }
      const bool swap = st.flags & B2J_STEP_SWAP;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint32_t x = swap ? b[j] : acc[j], y = swap ? acc[j] : b[j];
        acc[j] = elt_apply(st.op, x, y, c[j], st.imm);
      }
    }
    if (full4) {
      *reinterpret_cast<uint4*>(ptrs.out + i0) = make_uint4(acc[0], acc[1], acc[2], acc[3]);
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) if (i0 + j < p.n) ptrs.out[i0 + j] = acc[j];
    }
  }
}

}  // namespace b2j
