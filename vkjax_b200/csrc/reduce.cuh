// Reductions (B2J_K_REDUCE) and window reductions (B2J_K_REDUCE_WINDOW).
//
// Reference: reduce_sum/max/min/prod.comp, argmax/argmin.comp (one thread per output, serial loop with
// N-D unravel per element) and reduce_window_max_2d.comp.  Here:
//  * thread-per-output path: consecutive outputs map to consecutive threads, so a reduction whose
//    innermost *kept* dim is contiguous (ResNet global-average-pool [B,7,7,2048] -> [B,2048]) reads
//    coalesced; the loop keeps the reference's serial summation order (bit-equal to oracle/shader_ref.c).
//  * block-per-output path for few outputs / long reductions: 256 threads stride over the reduced
//    index, warp-shuffle + smem tree.
//  * typed (f32 / i32 / u32) instead of float-typed (reference quirk Q6); arg* tie-break = first index.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <type_traits>
#include "../../include/b2jax.h"

namespace b2j {

template <typename T> struct RedTraits;
template <> struct RedTraits<float> {
  __device__ static float lowest() { return -INFINITY; }
  __device__ static float highest() { return INFINITY; }
};
template <> struct RedTraits<int32_t> {
  __device__ static int32_t lowest() { return INT_MIN; }
  __device__ static int32_t highest() { return INT_MAX; }
};
template <> struct RedTraits<uint32_t> {
  __device__ static uint32_t lowest() { return 0u; }
  __device__ static uint32_t highest() { return 0xFFFFFFFFu; }
};

template <typename T, int KIND> struct RedAcc {
  T v;
  uint32_t idx;
  __device__ void init() {
    idx = 0;
    if (KIND == B2J_RED_SUM) v = (T)0;
    else if (KIND == B2J_RED_PROD) v = (T)1;
    else if (KIND == B2J_RED_MAX || KIND == B2J_RED_ARGMAX) v = RedTraits<T>::lowest();
    else v = RedTraits<T>::highest();
  }
  __device__ void push(T x, uint32_t i) {
    if (KIND == B2J_RED_SUM) v = v + x;
    else if (KIND == B2J_RED_PROD) v = v * x;
    else if (KIND == B2J_RED_MAX) v = (x > v || x != x) ? x : v;
    else if (KIND == B2J_RED_MIN) v = (x < v || x != x) ? x : v;
    // arg*: the first NaN wins (lax.argmax / argmin; consistent with reduce_max / min, which propagate NaN)
    else if (KIND == B2J_RED_ARGMAX) { if (i == 0 || (v == v && (x > v || x != x))) { v = x; idx = i; } }
    else { if (i == 0 || (v == v && (x < v || x != x))) { v = x; idx = i; } }
  }
  // does (x, xi) beat (y, yi)?  NaN beats every number; among equals (or two NaNs) the lower index wins
  __device__ static bool arg_better(T x, uint32_t xi, T y, uint32_t yi) {
    const bool xn = x != x, yn = y != y;
    if (xn || yn) return xn && (!yn || xi < yi);
    if (KIND == B2J_RED_ARGMAX ? x > y : x < y) return true;
    return x == y && xi < yi;
  }
  __device__ void merge(const RedAcc& o, bool o_valid) {
    if (!o_valid) return;
    if (KIND == B2J_RED_SUM) v = v + o.v;
    else if (KIND == B2J_RED_PROD) v = v * o.v;
    else if (KIND == B2J_RED_MAX) v = (o.v > v || o.v != o.v) ? o.v : v;
    else if (KIND == B2J_RED_MIN) v = (o.v < v || o.v != o.v) ? o.v : v;
    else { if (arg_better(o.v, o.idx, v, idx)) { v = o.v; idx = o.idx; } }
  }
};

__device__ __forceinline__ uint64_t red_offset(uint64_t i, uint32_t rank, const uint32_t* shape, const uint64_t* strides) {
  uint64_t off = 0;
#pragma unroll 1
  for (int d = (int)rank - 1; d >= 0; --d) {
    const uint32_t s = shape[d];
    const uint64_t q = i / s;
    off += (i - q * s) * strides[d];
    i = q;
  }
  return off;
}

template <typename T, int KIND>
__global__ void __launch_bounds__(256) reduce_thread_kernel(const __grid_constant__ b2j_reduce_params p,
                                                            uint32_t* __restrict__ out, const T* __restrict__ in) {
  for (uint64_t o = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; o < p.n_out; o += (uint64_t)gridDim.x * blockDim.x) {
    const T* base = in + red_offset(o, p.keep_rank, p.keep_shape, p.keep_strides);
    RedAcc<T, KIND> acc;
    acc.init();
    if (p.red_rank == 1) {
      const uint64_t st = p.red_strides[0];
      uint64_t r = 0;
      for (; r + 16 <= p.n_red; r += 16) {   // 16 independent loads in flight (few outputs, long columns: reduce over axis 0), serial accumulate order
        T x[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) x[j] = __ldg(base + (r + j) * st);
#pragma unroll
        for (int j = 0; j < 16; ++j) acc.push(x[j], (uint32_t)r + j);
      }
      for (; r + 4 <= p.n_red; r += 4) {     // 4 independent loads in flight, serial accumulate order
        const T x0 = __ldg(base + (r + 0) * st), x1 = __ldg(base + (r + 1) * st);
        const T x2 = __ldg(base + (r + 2) * st), x3 = __ldg(base + (r + 3) * st);
        acc.push(x0, (uint32_t)r); acc.push(x1, (uint32_t)r + 1); acc.push(x2, (uint32_t)r + 2); acc.push(x3, (uint32_t)r + 3);
      }
      for (; r < p.n_red; ++r) acc.push(__ldg(base + r * st), (uint32_t)r);
    } else {
      for (uint64_t r = 0; r < p.n_red; ++r)
        acc.push(__ldg(base + red_offset(r, p.red_rank, p.red_shape, p.red_strides)), (uint32_t)r);
    }
    if (KIND >= B2J_RED_ARGMAX) out[o] = acc.idx;
    else out[o] = *reinterpret_cast<uint32_t*>(&acc.v);
  }
}

template <typename T, int KIND>
__global__ void __launch_bounds__(256) reduce_block_kernel(const __grid_constant__ b2j_reduce_params p,
                                                           uint32_t* __restrict__ out, const T* __restrict__ in) {
  __shared__ T sv[8];
  __shared__ uint32_t si[8];
  __shared__ uint32_t sok[8];
  for (uint64_t o = blockIdx.x; o < p.n_out; o += gridDim.x) {
    const T* base = in + red_offset(o, p.keep_rank, p.keep_shape, p.keep_strides);
    RedAcc<T, KIND> acc;
    acc.init();
    bool valid = false;
    for (uint64_t r = threadIdx.x; r < p.n_red; r += blockDim.x) {
      const uint64_t off = p.red_rank == 1 ? r * p.red_strides[0] : red_offset(r, p.red_rank, p.red_shape, p.red_strides);
      const T x = __ldg(base + off);
      if (!valid) { acc.v = x; acc.idx = (uint32_t)r; valid = true;
                    if (KIND == B2J_RED_SUM || KIND == B2J_RED_PROD) { acc.init(); acc.push(x, (uint32_t)r); } }
      else if (KIND >= B2J_RED_ARGMAX) { RedAcc<T, KIND> t; t.v = x; t.idx = (uint32_t)r; acc.merge(t, true); }
      else acc.push(x, (uint32_t)r);
    }
    // warp tree
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) {
      RedAcc<T, KIND> t;
      t.v = __shfl_down_sync(0xffffffffu, acc.v, s);
      t.idx = __shfl_down_sync(0xffffffffu, acc.idx, s);
      const bool tv = __shfl_down_sync(0xffffffffu, (int)valid, s);
      if (!valid && tv) { acc = t; valid = true; }
      else acc.merge(t, tv && valid);
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) { sv[warp] = acc.v; si[warp] = acc.idx; sok[warp] = valid; }
    __syncthreads();
    if (warp == 0) {
      valid = lane < (int)(blockDim.x >> 5) ? sok[lane] : false;
      if (lane < (int)(blockDim.x >> 5)) { acc.v = sv[lane]; acc.idx = si[lane]; }
#pragma unroll
      for (int s = 4; s > 0; s >>= 1) {
        RedAcc<T, KIND> t;
        t.v = __shfl_down_sync(0xffffffffu, acc.v, s);
        t.idx = __shfl_down_sync(0xffffffffu, acc.idx, s);
        const bool tv = __shfl_down_sync(0xffffffffu, (int)valid, s);
        if (!valid && tv) { acc = t; valid = true; }
        else acc.merge(t, tv && valid);
      }
      if (lane == 0) {
        if (!valid) acc.init();
        if (KIND >= B2J_RED_ARGMAX) out[o] = acc.idx;
        else out[o] = *reinterpret_cast<uint32_t*>(&acc.v);
      }
    }
  }
}

// ---- warp-per-output path: the reduced run is contiguous in memory (reduce over the last axis: row sums, softmax
//      denominators, argmax over classes) and there are many outputs.  The thread-per-output kernel above would have each
//      thread walk its own row -- 32 different rows per warp load, nothing coalesced (0.15 of the HBM bandwidth on
//      [65536, 4096]); here the 32 lanes stride over one row with 128-bit loads and combine with warp shuffles.
template <typename T, int KIND>
__global__ void __launch_bounds__(256) reduce_warp_kernel(const __grid_constant__ b2j_reduce_params p,
                                                          uint32_t* __restrict__ out, const T* __restrict__ in) {
  const int lane = threadIdx.x & 31;
  const uint64_t warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
  for (uint64_t o = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; o < p.n_out; o += warps) {
    const T* base = in + red_offset(o, p.keep_rank, p.keep_shape, p.keep_strides);
    RedAcc<T, KIND> acc;
    acc.init();
    bool valid = false;
    auto take = [&](T x, uint32_t r) {
      if (!valid) { acc.v = x; acc.idx = r; valid = true; if (KIND == B2J_RED_SUM || KIND == B2J_RED_PROD) { acc.init(); acc.push(x, r); } }
      // arg*: a lane sees its elements in increasing index order, so a later element only wins when it is strictly better
      // (or the first NaN) -- no index tie-break needed here, that is left to the cross-lane merge below
      else if (KIND == B2J_RED_ARGMAX) { if (acc.v == acc.v && (x > acc.v || x != x)) { acc.v = x; acc.idx = r; } }
      else if (KIND == B2J_RED_ARGMIN) { if (acc.v == acc.v && (x < acc.v || x != x)) { acc.v = x; acc.idx = r; } }
      else acc.push(x, r);
    };
    const uint64_t n = p.n_red;
    uint64_t r = 0;
    if (((uintptr_t)base & 15u) == 0) {
      const uint64_t n4 = n >> 2;
      const uint4* b4 = reinterpret_cast<const uint4*>(base);
      uint64_t q = lane;
      for (; q + 32 < n4; q += 64) {                         // two independent 128-bit loads in flight per lane
        const uint4 a = __ldg(b4 + q), b = __ldg(b4 + q + 32);
        const T* xa = reinterpret_cast<const T*>(&a);
        const T* xb = reinterpret_cast<const T*>(&b);
#pragma unroll
        for (int j = 0; j < 4; ++j) take(xa[j], (uint32_t)(4 * q + j));
#pragma unroll
        for (int j = 0; j < 4; ++j) take(xb[j], (uint32_t)(4 * (q + 32) + j));
      }
      for (; q < n4; q += 32) {
        const uint4 a = __ldg(b4 + q);
        const T* xa = reinterpret_cast<const T*>(&a);
#pragma unroll
        for (int j = 0; j < 4; ++j) take(xa[j], (uint32_t)(4 * q + j));
      }
      r = n4 << 2;
    }
    for (uint64_t i = r + lane; i < n; i += 32) take(__ldg(base + i), (uint32_t)i);
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) {
      RedAcc<T, KIND> t;
      t.v = __shfl_down_sync(0xffffffffu, acc.v, s);
      t.idx = __shfl_down_sync(0xffffffffu, acc.idx, s);
      const bool tv = __shfl_down_sync(0xffffffffu, (int)valid, s);
      if (!valid && tv) { acc = t; valid = true; }
      else acc.merge(t, tv && valid);
    }
    if (lane == 0) {
      if (!valid) acc.init();
      if (KIND >= B2J_RED_ARGMAX) out[o] = acc.idx;
      else out[o] = *reinterpret_cast<uint32_t*>(&acc.v);
    }
  }
}

// Order-independent kinds (max / min / argmax / argmin) over a LONG strided run with few outputs per SM (reduce over axis 0 of
// [4096, 65536]: one thread per output left the SMs at ~20 % occupancy, 0.66 of the copy bandwidth): 8 row slices per output,
// a thread per (output, slice) with 16 loads in flight, slices combined through shared memory in index order.
template <typename T, int KIND>
__global__ void __launch_bounds__(256) reduce_sliced_kernel(const __grid_constant__ b2j_reduce_params p,
                                                            uint32_t* __restrict__ out, const T* __restrict__ in) {
  static_assert(KIND == B2J_RED_MAX || KIND == B2J_RED_MIN || KIND >= B2J_RED_ARGMAX, "sum / prod keep the serial order");
  __shared__ T sv[8][32];
  __shared__ uint32_t si[8][32];
  const uint32_t tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const uint64_t st = p.red_strides[0];
  const uint64_t per = (p.n_red + 7) / 8, r_lo = ty * per, r_hi = r_lo + per < p.n_red ? r_lo + per : p.n_red;
  for (uint64_t o0 = (uint64_t)blockIdx.x * 32; o0 < p.n_out; o0 += (uint64_t)gridDim.x * 32) {
    const uint64_t o = o0 + tx;
    RedAcc<T, KIND> acc;
    acc.init();
    bool valid = false;
    if (o < p.n_out && r_lo < r_hi) {
      const T* base = in + red_offset(o, p.keep_rank, p.keep_shape, p.keep_strides);
      acc.v = __ldg(base + r_lo * st); acc.idx = (uint32_t)r_lo; valid = true;
      auto take = [&](T x, uint32_t r) {
        if (KIND == B2J_RED_ARGMAX) { if (acc.v == acc.v && (x > acc.v || x != x)) { acc.v = x; acc.idx = r; } }
        else if (KIND == B2J_RED_ARGMIN) { if (acc.v == acc.v && (x < acc.v || x != x)) { acc.v = x; acc.idx = r; } }
        else acc.push(x, r);
      };
      uint64_t r = r_lo + 1;
      for (; r + 16 <= r_hi; r += 16) {
        T x[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) x[j] = __ldg(base + (r + j) * st);
#pragma unroll
        for (int j = 0; j < 16; ++j) take(x[j], (uint32_t)(r + j));
      }
      for (; r < r_hi; ++r) take(__ldg(base + r * st), (uint32_t)r);
    }
    sv[ty][tx] = acc.v; si[ty][tx] = valid ? acc.idx : 0xFFFFFFFFu;
    __syncthreads();
    if (ty == 0 && o < p.n_out) {
      RedAcc<T, KIND> tot;
      tot.v = sv[0][tx]; tot.idx = si[0][tx];
      bool tv = si[0][tx] != 0xFFFFFFFFu;
      for (int k = 1; k < 8; ++k) {
        if (si[k][tx] == 0xFFFFFFFFu) continue;
        RedAcc<T, KIND> t; t.v = sv[k][tx]; t.idx = si[k][tx];
        if (!tv) { tot = t; tv = true; } else tot.merge(t, true);
      }
      if (!tv) tot.init();
      if (KIND >= B2J_RED_ARGMAX) out[o] = tot.idx;
      else out[o] = *reinterpret_cast<uint32_t*>(&tot.v);
    }
    __syncthreads();
  }
}

// ---- reduce_window, 4-D.  VEC = 4 when the innermost dim is not windowed and divisible by 4:
//      each thread produces 4 consecutive innermost outputs with 128-bit loads (NHWC pooling).
//      Padded taps contribute the monoid identity (lax semantics; reference uses 0.0: quirk Q3). ------
template <typename T, int KIND, int VEC>
__global__ void __launch_bounds__(256) reduce_window_kernel(const __grid_constant__ b2j_reduce_window_params p,
                                                            T* __restrict__ out, const T* __restrict__ in) {
  const uint64_t n = (uint64_t)p.out_shape[0] * p.out_shape[1] * p.out_shape[2] * (p.out_shape[3] / VEC);
  const uint32_t c_vec = p.out_shape[3] / VEC;
  for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (uint64_t)gridDim.x * blockDim.x) {
    uint64_t rem = t;
    const uint32_t o3 = (uint32_t)(rem % c_vec) * VEC; rem /= c_vec;
    const uint32_t o2 = (uint32_t)(rem % p.out_shape[2]); rem /= p.out_shape[2];
    const uint32_t o1 = (uint32_t)(rem % p.out_shape[1]); rem /= p.out_shape[1];
    const uint32_t o0 = (uint32_t)rem;
    T acc[VEC];
#pragma unroll
    for (int j = 0; j < VEC; ++j)
      acc[j] = KIND == B2J_RW_MAX ? RedTraits<T>::lowest() : (KIND == B2J_RW_MIN ? RedTraits<T>::highest() : (T)0);
    for (uint32_t w0 = 0; w0 < p.window[0]; ++w0) {
      const int64_t i0 = (int64_t)o0 * p.strides[0] + w0 - p.pad_lo[0];
      if (i0 < 0 || i0 >= p.in_shape[0]) continue;
      for (uint32_t w1 = 0; w1 < p.window[1]; ++w1) {
        const int64_t i1 = (int64_t)o1 * p.strides[1] + w1 - p.pad_lo[1];
        if (i1 < 0 || i1 >= p.in_shape[1]) continue;
        for (uint32_t w2 = 0; w2 < p.window[2]; ++w2) {
          const int64_t i2 = (int64_t)o2 * p.strides[2] + w2 - p.pad_lo[2];
          if (i2 < 0 || i2 >= p.in_shape[2]) continue;
          const T* row = in + (((uint64_t)i0 * p.in_shape[1] + i1) * p.in_shape[2] + i2) * p.in_shape[3];
          if (VEC == 4) {
            const uint4 v = __ldg(reinterpret_cast<const uint4*>(row + o3));
            const T* x = reinterpret_cast<const T*>(&v);
#pragma unroll
            for (int j = 0; j < 4; ++j)
              acc[j] = KIND == B2J_RW_MAX ? ((x[j] > acc[j] || x[j] != x[j]) ? x[j] : acc[j])
                     : KIND == B2J_RW_MIN ? ((x[j] < acc[j] || x[j] != x[j]) ? x[j] : acc[j]) : acc[j] + x[j];
          } else {
            for (uint32_t w3 = 0; w3 < p.window[3]; ++w3) {
              const int64_t i3 = (int64_t)o3 * p.strides[3] + w3 - p.pad_lo[3];
              if (i3 < 0 || i3 >= p.in_shape[3]) continue;
              const T x = __ldg(row + i3);
              acc[0] = KIND == B2J_RW_MAX ? ((x > acc[0] || x != x) ? x : acc[0])
                     : KIND == B2J_RW_MIN ? ((x < acc[0] || x != x) ? x : acc[0]) : acc[0] + x;
            }
          }
        }
      }
    }
    T* dst = out + (((uint64_t)o0 * p.out_shape[1] + o1) * p.out_shape[2] + o2) * p.out_shape[3] + o3;
    if (VEC == 4) *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(acc);
    else dst[0] = acc[0];
  }
}

// ---- window monoid of the pooling kernels.  float max / min are ONE instruction (max.NaN / min.NaN: NaN-propagating like
//      lax.max / lax.min); the compare-compare-select form cost 3-4 instructions per element and made the 3x3 max-pool
//      issue-bound (ncu: 383 warp instructions per output float4, issue slots 58 % busy, DRAM 52 %).
template <typename T, int KIND> __device__ __forceinline__ T rw_identity() {
  return KIND == B2J_RW_MAX ? RedTraits<T>::lowest() : (KIND == B2J_RW_MIN ? RedTraits<T>::highest() : (T)0);
}
template <int KIND> __device__ __forceinline__ float rw_combine(float acc, float x) {
  float r;
  if (KIND == B2J_RW_MAX) asm("max.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(acc), "f"(x));
  else if (KIND == B2J_RW_MIN) asm("min.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(acc), "f"(x));
  else r = acc + x;
  return r;
}
template <int KIND> __device__ __forceinline__ int32_t rw_combine(int32_t acc, int32_t x) {
  return KIND == B2J_RW_MAX ? (x > acc ? x : acc) : KIND == B2J_RW_MIN ? (x < acc ? x : acc) : acc + x;
}
template <int KIND> __device__ __forceinline__ uint32_t rw_combine(uint32_t acc, uint32_t x) {
  return KIND == B2J_RW_MAX ? (x > acc ? x : acc) : KIND == B2J_RW_MIN ? (x < acc ? x : acc) : acc + x;
}

// ---- 2-D pooling fast path: window (1, KH, KW, 1), innermost dim not windowed and a multiple of 4.  One thread per
//      output float4; the KH x KW taps are fully unrolled and all loads are issued before the first combine, so each
//      thread keeps KH*KW 128-bit requests in flight (the generic kernel above walks them one by one).  Kept lean on
//      instructions, which -- not memory -- bounded the first version: IDX = uint32_t index math when the problem fits
//      (three 32-bit divisions per output instead of three 64-bit ones), tap offsets resolved once per thread, windows
//      that lie inside the image (all but the border) skip the per-tap bounds logic.  ResNet-50 stem pool ([256,112,112,64]
//      3x3 s2 max): 0.240 -> 0.158 ms = 6.5 TB/s (profiles/r01_pool_ab.txt).  A variant producing two adjacent output
//      columns per thread (15 loads per pair instead of 18) was slower at every shape (114-124 registers, half the
//      resident warps) and is gone.
template <typename T, int KIND, int KH, int KW, typename IDX>
__global__ void __launch_bounds__(256) pool2d_kernel(const __grid_constant__ b2j_reduce_window_params p, T* __restrict__ out,
                                                     const T* __restrict__ in) {
  typedef typename std::conditional<sizeof(IDX) == 4, int32_t, int64_t>::type SIDX;
  const uint32_t C = p.in_shape[3], c4 = C / 4;
  const uint32_t OW = p.out_shape[2], OH = p.out_shape[1];
  const int H = (int)p.in_shape[1], W = (int)p.in_shape[2];
  const IDX n = (IDX)p.out_shape[0] * OH * OW * c4;
  const SIDX row_pitch = (SIDX)W * (SIDX)C;
  SIDX tap_off[KH * KW];
#pragma unroll
  for (int kh = 0; kh < KH; ++kh)
#pragma unroll
    for (int kw = 0; kw < KW; ++kw) tap_off[kh * KW + kw] = (SIDX)kh * row_pitch + (SIDX)kw * (SIDX)C;
  for (IDX t = (IDX)blockIdx.x * 256u + threadIdx.x; t < n; t += (IDX)gridDim.x * 256u) {
    const uint32_t cv = (uint32_t)(t % c4);
    IDX r = t / c4;
    const uint32_t ow = (uint32_t)(r % OW); r /= OW;
    const uint32_t oh = (uint32_t)(r % OH);
    const uint32_t img = (uint32_t)(r / OH);
    const int ih0 = (int)(oh * p.strides[1]) - p.pad_lo[1], iw0 = (int)(ow * p.strides[2]) - p.pad_lo[2];
    // window origin; lies outside the tensor for padded windows and is only dereferenced at valid taps
    const T* base = in + (((SIDX)img * H + ih0) * row_pitch + (SIDX)iw0 * (SIDX)C + (SIDX)(cv * 4));
    T acc[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[j] = rw_identity<T, KIND>();
    uint4 v[KH * KW];
    if (ih0 >= 0 && ih0 + KH <= H && iw0 >= 0 && iw0 + KW <= W) {
#pragma unroll
      for (int k = 0; k < KH * KW; ++k) v[k] = __ldg(reinterpret_cast<const uint4*>(base + tap_off[k]));
#pragma unroll
      for (int k = 0; k < KH * KW; ++k) {
        const T* x = reinterpret_cast<const T*>(&v[k]);
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[j] = rw_combine<KIND>(acc[j], x[j]);
      }
    } else {
      bool ok[KH * KW];
#pragma unroll
      for (int kh = 0; kh < KH; ++kh)
#pragma unroll
        for (int kw = 0; kw < KW; ++kw) {
          const int ih = ih0 + kh, iw = iw0 + kw;
          ok[kh * KW + kw] = ih >= 0 && ih < H && iw >= 0 && iw < W;
          if (ok[kh * KW + kw]) v[kh * KW + kw] = __ldg(reinterpret_cast<const uint4*>(base + tap_off[kh * KW + kw]));
        }
#pragma unroll
      for (int k = 0; k < KH * KW; ++k) {
        if (!ok[k]) continue;
        const T* x = reinterpret_cast<const T*>(&v[k]);
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[j] = rw_combine<KIND>(acc[j], x[j]);
      }
    }
    *reinterpret_cast<uint4*>(out + (uint64_t)t * 4) = *reinterpret_cast<const uint4*>(acc);
  }
}

// ---- 2-D pooling for channel counts that are not a multiple of 4 (reference tests/test_reduce_window.py: C = 5, 17):
//      one thread per output ELEMENT, consecutive threads along (w, c), so a warp reads contiguous runs of every tap row;
//      32-bit index math, taps unrolled with all loads in flight.  The generic reduce_window_kernel (64-bit divisions,
//      rolled 4-deep window loops) ran these shapes at 9 % of the HBM bandwidth (profiles/r02_bandwidth_kernels.md).
template <typename T, int KIND, int KH, int KW>
__global__ void __launch_bounds__(256) pool2d_scalar_kernel(const __grid_constant__ b2j_reduce_window_params p, T* __restrict__ out,
                                                            const T* __restrict__ in) {
  const uint32_t C = p.in_shape[3], OW = p.out_shape[2], OH = p.out_shape[1];
  const int H = (int)p.in_shape[1], W = (int)p.in_shape[2];
  const uint32_t n = p.out_shape[0] * OH * OW * C;                     // host guarantees < 2^31
  const int row_pitch = W * (int)C;
  for (uint32_t t = blockIdx.x * 256u + threadIdx.x; t < n; t += gridDim.x * 256u) {
    const uint32_t c = t % C;
    uint32_t r = t / C;
    const uint32_t ow = r % OW; r /= OW;
    const uint32_t oh = r % OH;
    const uint32_t img = r / OH;
    const int ih0 = (int)(oh * p.strides[1]) - p.pad_lo[1], iw0 = (int)(ow * p.strides[2]) - p.pad_lo[2];
    const T* base = in + ((int)img * H + ih0) * row_pitch + iw0 * (int)C + (int)c;
    T v[KH * KW];
    bool ok[KH * KW];
#pragma unroll
    for (int kh = 0; kh < KH; ++kh)
#pragma unroll
      for (int kw = 0; kw < KW; ++kw) {
        const int ih = ih0 + kh, iw = iw0 + kw;
        ok[kh * KW + kw] = ih >= 0 && ih < H && iw >= 0 && iw < W;
        v[kh * KW + kw] = ok[kh * KW + kw] ? __ldg(base + kh * row_pitch + kw * (int)C) : rw_identity<T, KIND>();
      }
    T acc = rw_identity<T, KIND>();
#pragma unroll
    for (int k = 0; k < KH * KW; ++k) if (ok[k]) acc = rw_combine<KIND>(acc, v[k]);
    out[t] = acc;
  }
}

// ---- select_and_scatter_add (the gradient of max / min pooling; no reference handler: SURVEY.md §8 f2) --------------
//      XLA semantics: every window of `operand` selects ONE element -- scanning the window in row-major order, the
//      selected element is replaced by a candidate whenever select(selected, candidate) is false (select = ge for
//      max-pool: the first maximum wins; le for min-pool) -- and the window's `source` value is added at that position.
//      Formulated as a gather so that it needs no atomics and is deterministic: one thread per OPERAND element visits
//      the (at most ceil(k/s)^2) windows that contain it, re-runs each window's selection and adds the source value of
//      the windows that select it, in window row-major order.  Padded positions never win (they hold the identity). ----
template <bool GE>
__global__ void __launch_bounds__(256) select_and_scatter_add_kernel(const __grid_constant__ b2j_reduce_window_params p,
                                                                     float* __restrict__ out, const float* __restrict__ source,
                                                                     const float* __restrict__ operand) {
  // p.in_shape = operand shape, p.out_shape = source shape (one value per window)
  const uint64_t n = (uint64_t)p.in_shape[0] * p.in_shape[1] * p.in_shape[2] * p.in_shape[3];
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    uint64_t rem = i;
    int c[4];
    for (int d = 3; d >= 0; --d) { c[d] = (int)(rem % p.in_shape[d]); rem /= p.in_shape[d]; }
    // windows containing coordinate c[d]: o*stride - pad <= c < o*stride - pad + window
    int lo[4], hi[4];
    bool any = true;
    for (int d = 0; d < 4; ++d) {
      const int s = (int)p.strides[d], w = (int)p.window[d], x = c[d] + p.pad_lo[d];
      int l = x - w + 1;
      l = l <= 0 ? 0 : (l + s - 1) / s;
      int h = x / s;
      if (h > (int)p.out_shape[d] - 1) h = (int)p.out_shape[d] - 1;
      lo[d] = l; hi[d] = h;
      any = any && l <= h;
    }
    float acc = 0.0f;
    if (any) {
      for (int o0 = lo[0]; o0 <= hi[0]; ++o0)
        for (int o1 = lo[1]; o1 <= hi[1]; ++o1)
          for (int o2 = lo[2]; o2 <= hi[2]; ++o2)
            for (int o3 = lo[3]; o3 <= hi[3]; ++o3) {
              const int o[4] = {o0, o1, o2, o3};
              // re-run the window's selection
              bool have = false;
              float best = 0.0f;
              uint64_t best_idx = 0;
              for (uint32_t w0 = 0; w0 < p.window[0]; ++w0)
                for (uint32_t w1 = 0; w1 < p.window[1]; ++w1)
                  for (uint32_t w2 = 0; w2 < p.window[2]; ++w2)
                    for (uint32_t w3 = 0; w3 < p.window[3]; ++w3) {
                      const int q0 = o[0] * (int)p.strides[0] + (int)w0 - p.pad_lo[0], q1 = o[1] * (int)p.strides[1] + (int)w1 - p.pad_lo[1];
                      const int q2 = o[2] * (int)p.strides[2] + (int)w2 - p.pad_lo[2], q3 = o[3] * (int)p.strides[3] + (int)w3 - p.pad_lo[3];
                      if (q0 < 0 || q1 < 0 || q2 < 0 || q3 < 0 || q0 >= (int)p.in_shape[0] || q1 >= (int)p.in_shape[1] ||
                          q2 >= (int)p.in_shape[2] || q3 >= (int)p.in_shape[3]) continue;
                      const uint64_t qi = (((uint64_t)q0 * p.in_shape[1] + q1) * p.in_shape[2] + q2) * p.in_shape[3] + q3;
                      const float v = __ldg(operand + qi);
                      const bool keep = have && (GE ? best >= v : best <= v);
                      if (!keep) { best = v; best_idx = qi; have = true; }
                    }
              if (have && best_idx == i)
                acc += __ldg(source + (((uint64_t)o0 * p.out_shape[1] + o1) * p.out_shape[2] + o2) * p.out_shape[3] + o3);
            }
    }
    out[i] = acc;
  }
}

}  // namespace b2j
