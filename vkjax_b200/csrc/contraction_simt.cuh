// fp32-FMA contractions: the always-available exact path.
//  * conv_direct_kernel: conv_general_dilated for ANY dimension spec / low+high padding / stride /
//    lhs+rhs dilation (≙ reference conv2d.comp:44-95, one thread per output, serial (kh,kw,c) sum).
//    Index math is hoisted: strides are resolved once per thread, taps that fall into padding or
//    between dilated input samples are skipped instead of multiplied by 0.
//  * dot_kernel: 2-D dot_general with contracting dim 0|1 per side (≙ dot_general.comp:10-34).
// Both take the fused epilogue (b2j_epilogue).  The tensor-core kernels in gemm_tc.cuh are the fast
// path for NHWC shapes; these are the reference-order kernels the parity tests pin them against.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/b2jax.h"

namespace b2j {

struct EpiPtrs {
  const float* p[B2J_EPI_MAX_STEPS];
};

__device__ __forceinline__ float epi_op(uint32_t op, float a, float b) {
  switch (op) {
    case B2J_OP_ADD_F: return __fadd_rn(a, b);
    case B2J_OP_SUB_F: return __fsub_rn(a, b);
    case B2J_OP_MUL_F: return __fmul_rn(a, b);
    case B2J_OP_DIV_F: return __fdiv_rn(a, b);
    case B2J_OP_MAX_F: { float r; asm("max.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b)); return r; }   // lax.max propagates NaN
    case B2J_OP_MIN_F: { float r; asm("min.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b)); return r; }
    default: return a;
  }
}

__device__ __forceinline__ float epi_apply(const b2j_epilogue& e, const EpiPtrs& ptrs, float acc, uint32_t channel,
                                           uint64_t out_index) {
  for (uint32_t s = 0; s < e.n_steps; ++s) {
    const b2j_epi_step st = e.steps[s];
    float b;
    if (st.kind == B2J_EPK_IMM) b = __uint_as_float(st.imm);
    else if (st.kind == B2J_EPK_CHANNEL) b = __ldg(ptrs.p[s] + channel);
    else b = __ldg(ptrs.p[s] + out_index);
    acc = (st.flags & B2J_STEP_SWAP) ? epi_op(st.op, b, acc) : epi_op(st.op, acc, b);
  }
  return acc;
}

__global__ void __launch_bounds__(256) conv_direct_kernel(const __grid_constant__ b2j_conv_direct_params p,
                                                          const __grid_constant__ EpiPtrs epi, float* __restrict__ out,
                                                          const float* __restrict__ lhs, const float* __restrict__ rhs) {
  // row-major strides of the physical layouts
  uint64_t ls[4], rs[4], os_[4];
  ls[3] = rs[3] = 1;
  for (int d = 2; d >= 0; --d) { ls[d] = ls[d + 1] * p.lhs_shape[d + 1]; rs[d] = rs[d + 1] * p.rhs_shape[d + 1]; }
  const uint64_t l_n = ls[p.lhs_spec[0]], l_c = ls[p.lhs_spec[1]], l_h = ls[p.lhs_spec[2]], l_w = ls[p.lhs_spec[3]];
  const uint64_t r_o = rs[p.rhs_spec[0]], r_i = rs[p.rhs_spec[1]], r_h = rs[p.rhs_spec[2]], r_w = rs[p.rhs_spec[3]];
  const int H = p.lhs_shape[p.lhs_spec[2]], W = p.lhs_shape[p.lhs_spec[3]], C = p.rhs_shape[p.rhs_spec[1]];
  const int KH = p.rhs_shape[p.rhs_spec[2]], KW = p.rhs_shape[p.rhs_spec[3]];
  const uint64_t n_out = (uint64_t)p.out_shape[0] * p.out_shape[1] * p.out_shape[2] * p.out_shape[3];
  (void)os_;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_out; i += (uint64_t)gridDim.x * blockDim.x) {
    uint32_t oc[4];
    uint64_t rem = i;
    for (int d = 3; d >= 0; --d) { oc[d] = (uint32_t)(rem % p.out_shape[d]); rem /= p.out_shape[d]; }
    const uint32_t n = oc[p.out_spec[0]], o = oc[p.out_spec[1]], oh = oc[p.out_spec[2]], ow = oc[p.out_spec[3]];
    const float* xb = lhs + n * l_n;
    const float* wb = rhs + o * r_o;
    float sum = 0.0f;
    for (int kh = 0; kh < KH; ++kh) {
      // position in the (lhs-dilated, padded) input
      const int jh = (int)(oh * p.stride[0]) + kh * (int)p.rhs_dil[0] - p.pad_lo[0];
      if (jh < 0 || jh % (int)p.lhs_dil[0] != 0) continue;
      const int ih = jh / (int)p.lhs_dil[0];
      if (ih >= H) continue;
      for (int kw = 0; kw < KW; ++kw) {
        const int jw = (int)(ow * p.stride[1]) + kw * (int)p.rhs_dil[1] - p.pad_lo[1];
        if (jw < 0 || jw % (int)p.lhs_dil[1] != 0) continue;
        const int iw = jw / (int)p.lhs_dil[1];
        if (iw >= W) continue;
        const float* xp = xb + ih * l_h + iw * l_w;
        const float* wp = wb + kh * r_h + kw * r_w;
        int c = 0;
        for (; c + 4 <= C; c += 4) {
          const float x0 = __ldg(xp + (c + 0) * l_c), x1 = __ldg(xp + (c + 1) * l_c);
          const float x2 = __ldg(xp + (c + 2) * l_c), x3 = __ldg(xp + (c + 3) * l_c);
          const float w0 = __ldg(wp + (c + 0) * r_i), w1 = __ldg(wp + (c + 1) * r_i);
          const float w2 = __ldg(wp + (c + 2) * r_i), w3 = __ldg(wp + (c + 3) * r_i);
          sum = fmaf(x0, w0, sum); sum = fmaf(x1, w1, sum); sum = fmaf(x2, w2, sum); sum = fmaf(x3, w3, sum);
        }
        for (; c < C; ++c) sum = fmaf(__ldg(xp + c * l_c), __ldg(wp + c * r_i), sum);
      }
    }
    out[i] = epi_apply(p.epi, epi, sum, o, i);
  }
}

__global__ void __launch_bounds__(256) dot_kernel(const __grid_constant__ b2j_dot_params p,
                                                  const __grid_constant__ EpiPtrs epi, float* __restrict__ out,
                                                  const float* __restrict__ a, const float* __restrict__ b) {
  // A is [N,C] (cdim_a==1) or [C,N] (cdim_a==0); B is [C,M] (cdim_b==0) or [M,C] (cdim_b==1)
  const uint64_t a_row = p.cdim_a ? p.c : 1, a_k = p.cdim_a ? 1 : p.n;
  const uint64_t b_col = p.cdim_b ? p.c : 1, b_k = p.cdim_b ? 1 : p.m;
  const uint64_t n_out = (uint64_t)p.n * p.m;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_out; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint32_t row = (uint32_t)(i / p.m), col = (uint32_t)(i % p.m);
    const float* ap = a + row * a_row;
    const float* bp = b + col * b_col;
    float sum = 0.0f;
    uint32_t k = 0;
    for (; k + 4 <= p.c; k += 4) {
      const float a0 = __ldg(ap + (k + 0) * a_k), a1 = __ldg(ap + (k + 1) * a_k);
      const float a2 = __ldg(ap + (k + 2) * a_k), a3 = __ldg(ap + (k + 3) * a_k);
      const float b0 = __ldg(bp + (k + 0) * b_k), b1 = __ldg(bp + (k + 1) * b_k);
      const float b2 = __ldg(bp + (k + 2) * b_k), b3 = __ldg(bp + (k + 3) * b_k);
      sum = fmaf(a0, b0, sum); sum = fmaf(a1, b1, sum); sum = fmaf(a2, b2, sum); sum = fmaf(a3, b3, sum);
    }
    for (; k < p.c; ++k) sum = fmaf(__ldg(ap + k * a_k), __ldg(bp + k * b_k), sum);
    out[i] = epi_apply(p.epi, epi, sum, col, i);
  }
}

// rhs (any spec) -> wt[O][Kpad], k = (kh*KW + kw)*I + i; optional tf32 hi/lo split for 3xTF32.
__global__ void __launch_bounds__(256) weight_prep_kernel(const __grid_constant__ b2j_weight_prep_params p,
                                                          float* __restrict__ wt_hi, const float* __restrict__ rhs,
                                                          float* __restrict__ wt_lo) {
  uint64_t rs[4];
  rs[3] = 1;
  for (int d = 2; d >= 0; --d) rs[d] = rs[d + 1] * p.rhs_shape[d + 1];
  const uint32_t O = p.rhs_shape[p.rhs_spec[0]], I = p.rhs_shape[p.rhs_spec[1]];
  const uint32_t KH = p.rhs_shape[p.rhs_spec[2]], KW = p.rhs_shape[p.rhs_spec[3]];
  const uint32_t K = KH * KW * I;
  const uint64_t n = (uint64_t)O * p.kpad;
  for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (uint64_t)gridDim.x * blockDim.x) {
    const uint32_t o = (uint32_t)(t / p.kpad), k = (uint32_t)(t % p.kpad);
    float v = 0.0f;
    if (p.cpad == 0) {
      if (k < K) {
        const uint32_t i = k % I, kw = (k / I) % KW, kh = k / (I * KW);
        v = __ldg(rhs + o * rs[p.rhs_spec[0]] + i * rs[p.rhs_spec[1]] + kh * rs[p.rhs_spec[2]] + kw * rs[p.rhs_spec[3]]);
      }
    } else {
      // K layout of re-laid-out activations (b2j_relayout_params): k = (th*taps_w + tw)*cpad + j
      const uint32_t j = k % p.cpad, tap = k / p.cpad, tw = tap % p.taps_w, th = tap / p.taps_w;
      uint32_t kh = th, kw = tw, i = j;
      bool ok = th < p.taps_h && j < I;
      if (p.n_map) {
        const b2j_fold_entry e = p.map[j < B2J_FOLD_CHANNELS ? j : 0];
        kh = p.tap_h * th + e.dh; kw = p.tap_w * tw + e.dw; i = e.c;
        ok = th < p.taps_h && j < p.n_map && e.valid;
      }
      if (ok && kh < KH && kw < KW && i < I)
        v = __ldg(rhs + o * rs[p.rhs_spec[0]] + i * rs[p.rhs_spec[1]] + kh * rs[p.rhs_spec[2]] + kw * rs[p.rhs_spec[3]]);
    }
    if (p.split == 1) {
      uint32_t r;                                  // hi = nearest TF32 (halves |lo| against truncation), lo = exact rest
      asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
      const float hi = __uint_as_float(r);
      wt_hi[t] = hi;
      wt_lo[t] = v - hi;
    } else if (p.split == 2) {
      // single-pass TF32: round to nearest here, once, so the TMA-fed kernel can hand raw tiles to the tensor core
      uint32_t r;
      asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
      wt_hi[t] = __uint_as_float(r);
    } else {
      wt_hi[t] = v;
    }
  }
}

// NHWC activations -> NHWC' with channels padded to a multiple the im2col tensor map can address, optionally with a
// stride-s window folded into the channel dimension (space-to-depth; see b2j_relayout_params).
// One CTA produces RL_TILE consecutive destination pixels of one destination row: it first stages the source pixels
// they draw from (a few source rows x a contiguous column range x all channels) in shared memory with coalesced loads,
// then every thread assembles destination float4s from shared memory and stores them coalesced.  (A direct gather --
// one thread per destination float4 reading its 4 sources from global memory -- ran at 2 TB/s: L1 wavefront bound.)
__device__ __forceinline__ float relayout_rna(float x) { uint32_t r; asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x)); return __uint_as_float(r); }
__device__ __forceinline__ float4 relayout_rna4(float4 v) { return make_float4(relayout_rna(v.x), relayout_rna(v.y), relayout_rna(v.z), relayout_rna(v.w)); }
constexpr int RL_TILE = 128;     // destination pixels per tile row
__global__ void __launch_bounds__(256) relayout_kernel(const __grid_constant__ b2j_relayout_params p, float* __restrict__ dst,
                                                       const float* __restrict__ src, const int max_dh, const int max_dw, const int tile_rows) {
  extern __shared__ float rl_smem[];
  const uint32_t tiles_w = (p.ow + RL_TILE - 1) / RL_TILE;
  const uint32_t row_blocks = (p.oh + tile_rows - 1) / tile_rows;       // a tile = tile_rows destination rows x RL_TILE pixels
  const uint32_t n_tiles = p.batch * row_blocks * tiles_w;
  const uint32_t oc4 = p.oc / 4;
  const int C = (int)p.c;
  const int rows = (int)p.fold_h * (tile_rows - 1) + max_dh + 1;          // source rows one tile can touch
  const int cols = (int)p.fold_w * (RL_TILE - 1) + max_dw + 1;          // source columns one tile can touch
  const int row_floats = cols * C;
  // Index math is hoisted out of the per-element loops (it, not memory, bounded the first version of this kernel):
  // when 256 % (oc/4) == 0 a thread always assembles the same 4 destination channels, so their shared-memory offsets
  // are resolved once per kernel.
  const bool fixed_j4 = (256u % oc4) == 0u;
  int off[4];
  bool okc[4];
  auto resolve = [&](uint32_t j4) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const uint32_t j = j4 * 4 + e;
      int dh = 0, dw = 0;
      uint32_t c = j;
      okc[e] = j < p.c;
      if (p.n_map) {
        const b2j_fold_entry f = p.map[j < B2J_FOLD_CHANNELS ? j : 0];
        dh = f.dh; dw = f.dw; c = f.c;
        okc[e] = j < p.n_map && f.valid;
      }
      off[e] = dh * row_floats + dw * C + (int)c;
    }
  };
  if (fixed_j4) resolve(threadIdx.x % oc4);
  const int px_step = fixed_j4 ? (int)(256u / oc4) : 0, px0 = fixed_j4 ? (int)(threadIdx.x / oc4) : 0;
  const int px_floats = (int)p.fold_w * C;
  for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const uint32_t tw = tile % tiles_w, t2 = tile / tiles_w;
    const uint32_t a = (t2 % row_blocks) * tile_rows, img = t2 / row_blocks;
    const int n_rows = (int)min((uint32_t)tile_rows, p.oh - a);
    const uint32_t b0 = tw * RL_TILE;
    const int h0 = (int)(p.fold_h * a) - p.pad_h, w0 = (int)(p.fold_w * b0) - p.pad_w;
    const int64_t ebase = (int64_t)((uint64_t)img * p.h * p.w * p.c) + (int64_t)w0 * C;     // element offset (f32 or packed uint8 source)
    const uint8_t* src8 = reinterpret_cast<const uint8_t*>(src);
    // valid float range of a staged row: source columns [max(0,-w0), min(cols, W-w0))
    const int f_lo = (w0 < 0 ? -w0 : 0) * C, f_hi = ((int)p.w - w0 < cols ? (int)p.w - w0 : cols) * C;
    __syncthreads();                                                     // previous tile's readers are done
    for (int r = 0; r < rows; ++r) {
      const int ih = h0 + r;
      const bool row_ok = ih >= 0 && ih < (int)p.h;
      const int64_t erow = ebase + (int64_t)((uint64_t)(row_ok ? ih : 0) * p.w * p.c);
      float* drow_s = rl_smem + r * row_floats;
      if (!p.src_u8 && p.pre_n == 0) {
        const float* srow = src + erow;
        for (int f = threadIdx.x; f < row_floats; f += 256)             // consecutive threads, consecutive floats
          drow_s[f] = (row_ok && f >= f_lo && f < f_hi) ? __ldg(srow + f) : 0.0f;
      } else if (!p.src_u8) {
        const float* srow = src + erow;
        for (int f = threadIdx.x; f < row_floats; f += 256) {
          float x = 0.0f;
          if (row_ok && f >= f_lo && f < f_hi) {
            x = __ldg(srow + f);
            for (uint32_t s_ = 0; s_ < p.pre_n; ++s_) x = epi_op(p.pre_op[s_], x, __uint_as_float(p.pre_imm[s_]));
          }
          drow_s[f] = x;
        }
      } else {
        // uint8 pixels: widen, then the fused input chain (e.g. / 255), each step rounded on its own; padding stays 0
        const uint8_t* srow = src8 + erow;
        for (int f = threadIdx.x; f < row_floats; f += 256) {
          float x = 0.0f;
          if (row_ok && f >= f_lo && f < f_hi) {
            x = (float)__ldg(srow + f);
            for (uint32_t s_ = 0; s_ < p.pre_n; ++s_) x = epi_op(p.pre_op[s_], x, __uint_as_float(p.pre_imm[s_]));
          }
          drow_s[f] = x;
        }
      }
    }
    __syncthreads();
    const int n_px = (int)min((uint32_t)RL_TILE, p.ow - b0);
    float* drow = dst + (((uint64_t)img * p.oh + a) * p.ow + b0) * p.oc;
    const int row_adv = (int)p.fold_h * row_floats;                      // one destination row further down = fold_h staged rows
    const uint64_t drow_pitch = (uint64_t)p.ow * p.oc;
    if (fixed_j4) {
      // (destination row, pixel) walked as two nested loops with constant strides: no division, no 64-bit multiply per
      // float4 (ncu on the first version: 118 instructions per stored float4, issue slots 71 % busy, DRAM 41 %;
      // 0.159 -> 0.147 ms on the ResNet-50 stem.  Streaming stores and other grid sizes did not move it further.)
      const uint32_t j4 = threadIdx.x % oc4;
      const bool any = okc[0] || okc[1] || okc[2] || okc[3];
      const bool rnd = p.round_tf32 != 0;
      const float* srow = rl_smem + px0 * px_floats;
      float* d = drow + ((uint32_t)px0 * oc4 + j4) * 4;
      const int s_step = px_step * px_floats;
      const uint32_t d_step = (uint32_t)px_step * p.oc;
      for (int dr = 0; dr < n_rows; ++dr, srow += row_adv, d += drow_pitch) {
        const float* s0 = srow;
        float* dp = d;
        for (int px = px0; px < n_px; px += px_step, s0 += s_step, dp += d_step) {
          float4 v = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
          if (any) {
            v = make_float4(okc[0] ? s0[off[0]] : 0.0f, okc[1] ? s0[off[1]] : 0.0f, okc[2] ? s0[off[2]] : 0.0f,
                            okc[3] ? s0[off[3]] : 0.0f);
            if (rnd) v = relayout_rna4(v);
          }
          *reinterpret_cast<float4*>(dp) = v;
        }
      }
    } else {
      for (uint32_t i = threadIdx.x; i < (uint32_t)(n_px * n_rows) * oc4; i += 256) {
        const uint32_t q = i / oc4, dr = q / (uint32_t)n_px, px = q - dr * (uint32_t)n_px;
        resolve(i - q * oc4);
        const float* s0 = rl_smem + dr * row_adv + px * px_floats;
        float4 v = make_float4(okc[0] ? s0[off[0]] : 0.0f, okc[1] ? s0[off[1]] : 0.0f, okc[2] ? s0[off[2]] : 0.0f,
                               okc[3] ? s0[off[3]] : 0.0f);
        if (p.round_tf32) v = relayout_rna4(v);
        *reinterpret_cast<float4*>(drow + dr * drow_pitch + ((uint64_t)px * oc4 + (i - q * oc4)) * 4) = v;
      }
    }
  }
}

}  // namespace b2j
