// Bandwidth-bound data-movement kernels: strided copy (broadcast_in_dim / slice / rev / N-D transpose),
// tiled 2-D transpose, gather, scatter-add, concatenate, threefry2x32.
// Each replaces the like-named reference shader (vkjax/shaders/*.comp); integer / index work is bit-exact.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/b2jax.h"

namespace b2j {

// ---- strided copy ------------------------------------------------------------------------------
// out[i] = in[base + sum_d coord_d(i) * stride_d].  Host side collapses dims first; when the innermost
// dim is contiguous (stride 1) consecutive threads read consecutive addresses.
__global__ void __launch_bounds__(256) strided_copy_kernel(const __grid_constant__ b2j_strided_params p,
                                                           uint32_t* __restrict__ out, const uint32_t* __restrict__ in) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < p.n; i += (uint64_t)gridDim.x * blockDim.x) {
    uint64_t rem = i;
    int64_t idx = p.base;
#pragma unroll 1
    for (int d = (int)p.rank - 1; d >= 0; --d) {
      const uint32_t s = p.shape[d];
      const uint64_t q = rem / s;
      idx += (int64_t)(rem - q * s) * p.strides[d];
      rem = q;
    }
    out[i] = __ldg(in + idx);
  }
}

// Vectorised variant: the innermost (collapsed) output dim is a multiple of 4 and is read with stride 1 (slices, rev of
// outer dims, concatenation-like copies) or stride 0 (broadcast_in_dim of a vector along new leading dims): one thread
// moves 4 consecutive outputs with one 128-bit store (and one 128-bit load, or one scalar load for stride 0), 32-bit
// index math, 4 vectors per thread in flight.  The scalar kernel above materialised a broadcast at 15 % of the HBM bandwidth.
constexpr int SC_VECS = 4;
__global__ void __launch_bounds__(256) strided_copy_vec4_kernel(const __grid_constant__ b2j_strided_params p,
                                                                uint32_t* __restrict__ out, const uint32_t* __restrict__ in) {
  const uint32_t nvec = (uint32_t)(p.n >> 2);                             // host guarantees n % 4 == 0 and n < 2^32
  const uint32_t inner4 = p.shape[p.rank - 1] >> 2;
  const bool bcast = p.strides[p.rank - 1] == 0;
  for (uint32_t tile = blockIdx.x; (uint64_t)tile * (256 * SC_VECS) < nvec; tile += gridDim.x) {
    uint4 v[SC_VECS];
    uint32_t vi[SC_VECS];
#pragma unroll
    for (int k = 0; k < SC_VECS; ++k) {
      vi[k] = tile * (256 * SC_VECS) + k * 256 + threadIdx.x;
      if (vi[k] >= nvec) continue;
      uint32_t rem = vi[k] / inner4;
      int64_t idx = p.base + (bcast ? 0 : (int64_t)((vi[k] - rem * inner4) << 2));
#pragma unroll 1
      for (int d = (int)p.rank - 2; d >= 0; --d) {
        const uint32_t s = p.shape[d];
        const uint32_t q = rem / s;
        idx += (int64_t)(rem - q * s) * p.strides[d];
        rem = q;
      }
      if (bcast) { const uint32_t t = __ldg(in + idx); v[k] = make_uint4(t, t, t, t); }
      else v[k] = __ldg(reinterpret_cast<const uint4*>(in + idx));
    }
#pragma unroll
    for (int k = 0; k < SC_VECS; ++k)
      if (vi[k] < nvec) reinterpret_cast<uint4*>(out)[vi[k]] = v[k];
  }
}

// ---- 2-D transpose, 32x32 smem tiles (+1 padding: conflict-free), coalesced both sides ----------
__global__ void __launch_bounds__(256) transpose2d_kernel(b2j_transpose_params p, uint32_t* __restrict__ out,
                                                          const uint32_t* __restrict__ in) {
  // batched: `batch` independent [rows, cols] matrices back to back (NCHW <-> NHWC is [N][C][H*W] <-> [N][H*W][C])
  __shared__ uint32_t tile[32][33];
  const uint32_t c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  const uint32_t tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
  const uint32_t nbatch = p.batch ? p.batch : 1u;
  for (uint32_t b = blockIdx.z; b < nbatch; b += gridDim.z) {
    const uint64_t off = (uint64_t)b * p.rows * p.cols;
#pragma unroll
    for (int k = 0; k < 32; k += 8) {
      const uint32_t r = r0 + ty + k, c = c0 + tx;
      if (r < p.rows && c < p.cols) tile[ty + k][tx] = __ldg(in + off + (uint64_t)r * p.cols + c);
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 32; k += 8) {
      const uint32_t c = c0 + ty + k, r = r0 + tx;    // out is [cols][rows]
      if (c < p.cols && r < p.rows) out[off + (uint64_t)c * p.rows + r] = tile[tx][ty + k];
    }
    __syncthreads();
  }
}

// Same, 64 x 64 tiles with 128-bit global accesses on BOTH sides (rows % 4 == 0 and cols % 4 == 0): 16 bytes per thread and
// request instead of 4 keep four times the bytes in flight (the 32 x 32 kernel moved 4.1 TB/s = 0.62 of the copy bandwidth:
// bench.py --config bandwidth).  The 4-byte shuffle happens in shared memory (scalar accesses, pitch 65: two-way conflicts at worst).
__global__ void __launch_bounds__(256) transpose2d_vec_kernel(b2j_transpose_params p, uint32_t* __restrict__ out,
                                                              const uint32_t* __restrict__ in) {
  __shared__ uint32_t tile[64][65];
  const uint32_t c0 = blockIdx.x * 64, r0 = blockIdx.y * 64;
  const uint32_t nbatch = p.batch ? p.batch : 1u;
  for (uint32_t b = blockIdx.z; b < nbatch; b += gridDim.z) {
    const uint64_t off = (uint64_t)b * p.rows * p.cols;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const uint32_t idx = threadIdx.x + 256u * k, rr = idx >> 4, c4 = (idx & 15u) * 4;
      const uint32_t r = r0 + rr, c = c0 + c4;
      if (r < p.rows && c < p.cols) {        // cols % 4 == 0: a vector is inside or outside as a whole
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(in + off + (uint64_t)r * p.cols + c));
        tile[c4][rr] = v.x; tile[c4 + 1][rr] = v.y; tile[c4 + 2][rr] = v.z; tile[c4 + 3][rr] = v.w;
      }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const uint32_t idx = threadIdx.x + 256u * k, cc = idx >> 4, r4 = (idx & 15u) * 4;
      const uint32_t c = c0 + cc, r = r0 + r4;           // out is [cols][rows]
      if (c < p.cols && r < p.rows)
        *reinterpret_cast<uint4*>(out + off + (uint64_t)c * p.rows + r) = make_uint4(tile[cc][r4], tile[cc][r4 + 1], tile[cc][r4 + 2], tile[cc][r4 + 3]);
    }
    __syncthreads();
  }
}

// ---- gather (XLA semantics, start indices clamped) ---------------------------------------------
__global__ void __launch_bounds__(256) gather_kernel(const __grid_constant__ b2j_gather_params p,
                                                     uint32_t* __restrict__ out, const uint32_t* __restrict__ operand,
                                                     const int32_t* __restrict__ indices) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < p.n; i += (uint64_t)gridDim.x * blockDim.x) {
    int64_t coord[B2J_MAX_RANK];
#pragma unroll
    for (int d = 0; d < B2J_MAX_RANK; ++d) coord[d] = 0;
    uint64_t rem = i;
    int64_t bidx = 0;
#pragma unroll 1
    for (int d = (int)p.out_rank - 1; d >= 0; --d) {
      const uint32_t s = p.out_shape[d];
      const uint64_t q = rem / s;
      const int64_t c = (int64_t)(rem - q * s);
      rem = q;
      const int od = p.out_dim_to_operand_dim[d];
      if (od >= 0) {
#pragma unroll
        for (int k = 0; k < B2J_MAX_RANK; ++k) if (k == od) coord[k] += c;
      } else {
        bidx += c * p.out_dim_batch_stride[d];
      }
    }
#pragma unroll 1
    for (uint32_t k = 0; k < p.idx_vec_len; ++k) {
      const uint32_t od = p.start_index_map[k];
      int64_t s = indices[bidx * p.idx_vec_len + k];
      const int64_t hi = (int64_t)p.operand_shape[od] - (int64_t)p.slice_sizes[od];
      s = s < 0 ? 0 : (s > hi ? hi : s);
#pragma unroll
      for (int q = 0; q < B2J_MAX_RANK; ++q) if (q == (int)od) coord[q] += s;
    }
    uint64_t idx = 0;
#pragma unroll
    for (int d = 0; d < B2J_MAX_RANK; ++d) if (d < (int)p.operand_rank) idx = idx * p.operand_shape[d] + (uint64_t)coord[d];
    out[i] = __ldg(operand + idx);
  }
}

// ---- scatter-add: out already holds a copy of operand; one thread per update element ------------
__global__ void __launch_bounds__(256) scatter_add_kernel(const __grid_constant__ b2j_scatter_params p,
                                                          uint32_t* __restrict__ out, const int32_t* __restrict__ indices,
                                                          const uint32_t* __restrict__ updates) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < p.n_updates; i += (uint64_t)gridDim.x * blockDim.x) {
    int64_t coord[B2J_MAX_RANK];
#pragma unroll
    for (int d = 0; d < B2J_MAX_RANK; ++d) coord[d] = 0;
    uint64_t rem = i;
    int64_t bidx = 0;
#pragma unroll 1
    for (int d = (int)p.upd_rank - 1; d >= 0; --d) {
      const uint32_t s = p.upd_shape[d];
      const uint64_t q = rem / s;
      const int64_t c = (int64_t)(rem - q * s);
      rem = q;
      const int od = p.upd_dim_to_operand_dim[d];
      if (od >= 0) {
#pragma unroll
        for (int k = 0; k < B2J_MAX_RANK; ++k) if (k == od) coord[k] += c;
      } else {
        bidx += c * p.upd_dim_batch_stride[d];
      }
    }
#pragma unroll 1
    for (uint32_t k = 0; k < p.idx_vec_len; ++k) {
      const uint32_t od = p.scatter_dims_to_operand_dims[k];
      const int64_t s = indices[bidx * p.idx_vec_len + k];
#pragma unroll
      for (int q = 0; q < B2J_MAX_RANK; ++q) if (q == (int)od) coord[q] += s;
    }
    bool ok = true;
    uint64_t idx = 0;
#pragma unroll
    for (int d = 0; d < B2J_MAX_RANK; ++d) if (d < (int)p.operand_rank) {
      ok = ok && coord[d] >= 0 && coord[d] < (int64_t)p.operand_shape[d];
      idx = idx * p.operand_shape[d] + (uint64_t)coord[d];
    }
    if (!ok) continue;     // XLA: out-of-bounds updates are dropped
    const uint32_t u = updates[i];
    if (p.dtype == B2J_F32) atomicAdd(reinterpret_cast<float*>(out) + idx, __uint_as_float(u));
    else atomicAdd(out + idx, u);
  }
}

// ---- concatenate of two operands along one axis: views [outer, ca|cb, inner] ---------------------
__global__ void __launch_bounds__(256) concat_kernel(b2j_concat_params p, uint32_t* __restrict__ out,
                                                     const uint32_t* __restrict__ a, const uint32_t* __restrict__ b) {
  const uint64_t row = (p.ca + p.cb) * p.inner;
  const uint64_t n = p.outer * row;
  const uint64_t ra = p.ca * p.inner, rb = p.cb * p.inner;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t o = i / row, r = i - o * row;
    out[i] = r < ra ? __ldg(a + o * ra + r) : __ldg(b + o * rb + (r - ra));
  }
}

// ---- threefry2x32, 20 rounds (bit-exact; Random123 KATs in tests/test_oracle_golden.py) ----------
__device__ __forceinline__ void tf_round(uint32_t& x0, uint32_t& x1, int r) {
  x0 += x1;
  x1 = __funnelshift_l(x1, x1, r);
  x1 ^= x0;
}

__global__ void __launch_bounds__(256) threefry_kernel(b2j_threefry_params p, uint32_t* __restrict__ out0,
                                                       uint32_t* __restrict__ out1, const uint32_t* __restrict__ key0,
                                                       const uint32_t* __restrict__ key1, const uint32_t* __restrict__ d0,
                                                       const uint32_t* __restrict__ d1) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < p.n; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t ki = p.key_is_scalar ? 0 : i;
    uint32_t ks[3];
    ks[0] = key0[ki];
    ks[1] = key1[ki];
    ks[2] = 0x1BD11BDAu ^ ks[0] ^ ks[1];
    uint32_t x0 = d0[i] + ks[0], x1 = d1[i] + ks[1];
    const int rot[2][4] = {{13, 15, 26, 6}, {17, 29, 16, 24}};
#pragma unroll
    for (int g = 0; g < 5; ++g) {
#pragma unroll
      for (int r = 0; r < 4; ++r) tf_round(x0, x1, rot[g & 1][r]);
      x0 += ks[(g + 1) % 3];
      x1 += ks[(g + 2) % 3] + (uint32_t)(g + 1);
    }
    out0[i] = x0;
    out1[i] = x1;
  }
}

// ---- lhs dilation (zero stuffing) for the tensor-core convolution path: an NHWC tensor is written to
//      dst[n, dil_h*h, dil_w*w, c] of a [n, (H-1)*dil_h+1, (W-1)*dil_w+1, c] tensor, zeros elsewhere, so that a
//      conv_general_dilated with lhs_dilation (transposed convolution, conv input gradients; reference conv2d.comp:32-42
//      skips the in-between taps per MAC) becomes a plain convolution the TMA im2col map can address.
//      One thread per destination element, consecutive threads along c: coalesced on both sides. -------------------
__global__ void __launch_bounds__(256) dilate_kernel(const __grid_constant__ b2j_dilate_params p, uint32_t* __restrict__ out,
                                                     const uint32_t* __restrict__ in) {
  const uint64_t n = (uint64_t)p.batch * p.oh * p.ow * p.c;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    uint64_t rem = i;
    const uint32_t c = (uint32_t)(rem % p.c); rem /= p.c;
    const uint32_t b = (uint32_t)(rem % p.ow); rem /= p.ow;
    const uint32_t a = (uint32_t)(rem % p.oh);
    const uint32_t img = (uint32_t)(rem / p.oh);
    uint32_t v = 0;
    const uint32_t h = a / p.dil_h, w = b / p.dil_w;
    if (h * p.dil_h == a && w * p.dil_w == b && h < p.h && w < p.w)
      v = __ldg(in + (((uint64_t)img * p.h + h) * p.w + w) * p.c + c);
    out[i] = v;
  }
}

}  // namespace b2j
