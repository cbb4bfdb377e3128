/* CPU ORACLE -- TEST INFRASTRUCTURE ONLY (see oracle/eval_jaxpr.py header).
 *
 * Plain-C restatement of the reference's hot-path shaders, keeping their arithmetic: one output
 * element at a time, float32 accumulator, terms added in the shader's loop order.
 *   conv2d_ref            <- reference vkjax/shaders/conv2d.comp:44-95        (any dimension spec,
 *                            low padding, strides, lhs/rhs dilation; OOB / dilation holes add 0)
 *   dot_general_ref       <- reference vkjax/shaders/dot_general.comp:10-34   (contracting dim 0|1)
 *   reduce_window_max_ref <- reference vkjax/shaders/reduce_window_max_2d.comp:29-52; `q3` = 1 keeps the
 *                            shader's padded-tap value of 0.0 (quirk Q3), 0 uses -inf (lax semantics)
 *   reduce_sum_ref        <- reference vkjax/shaders/reduce_sum.comp:39-52    ([outer, red, inner] view)
 * `fma_mode` = 1 contracts a*b+sum into fmaf (what nvcc does in the CUDA kernels), 0 rounds the
 * product first (GLSL without contraction).  Each function computes outputs [begin, end): the Python
 * wrapper (oracle/shader_ref.py) splits that range over host threads -- per-output order stays serial.
 */
#include <math.h>
#include <stdint.h>
#include <stddef.h>

static void unravel(uint64_t idx, const uint32_t* shape, int n, int64_t* c) {
  for (int i = n - 1; i >= 0; --i) { c[i] = (int64_t)(idx % shape[i]); idx /= shape[i]; }
}
static uint64_t ravel(const int64_t* c, const uint32_t* shape, int n) {
  uint64_t idx = 0, stride = 1;
  for (int i = n - 1; i >= 0; --i) { idx += (uint64_t)c[i] * stride; stride *= shape[i]; }
  return idx;
}

void conv2d_ref(float* out, const float* a, const float* b, const uint32_t* shape_a, const uint32_t* shape_b,
                const uint32_t* shape_out, const uint32_t* spec_lhs, const uint32_t* spec_rhs, const uint32_t* spec_out,
                const int32_t* padding, const uint32_t* strides, const uint32_t* dil_lhs, const uint32_t* dil_rhs,
                int fma_mode, int64_t begin, int64_t end) {
  const int64_t dH = (int64_t)shape_a[spec_lhs[2]] * dil_lhs[0], dW = (int64_t)shape_a[spec_lhs[3]] * dil_lhs[1];
  const int64_t KHd = (int64_t)shape_b[spec_rhs[2]] * dil_rhs[0], KWd = (int64_t)shape_b[spec_rhs[3]] * dil_rhs[1];
  const int64_t C = shape_b[spec_rhs[1]];
  for (int64_t index = begin; index < end; ++index) {
    int64_t co[4];
    unravel((uint64_t)index, shape_out, 4, co);
    float sum = 0.0f;
    for (int64_t i0 = 0; i0 < KHd; i0 += dil_rhs[0]) {
      const int64_t j0 = co[spec_out[2]] * strides[0] + i0 - padding[0];
      for (int64_t i1 = 0; i1 < KWd; i1 += dil_rhs[1]) {
        const int64_t j1 = co[spec_out[3]] * strides[1] + i1 - padding[1];
        for (int64_t c = 0; c < C; ++c) {
          /* is_out_of_bounds on the dilated input, coords_in_dilation for the holes */
          if (j0 < 0 || j0 >= dH || j1 < 0 || j1 >= dW) continue;
          if ((j0 % dil_lhs[0]) > 0 || (j1 % dil_lhs[1]) > 0) continue;
          int64_t ca[4], cb[4];
          ca[spec_lhs[0]] = co[spec_out[0]]; ca[spec_lhs[1]] = c;
          ca[spec_lhs[2]] = j0 / dil_lhs[0]; ca[spec_lhs[3]] = j1 / dil_lhs[1];
          cb[spec_rhs[0]] = co[spec_out[1]]; cb[spec_rhs[1]] = c;
          cb[spec_rhs[2]] = i0 / dil_rhs[0]; cb[spec_rhs[3]] = i1 / dil_rhs[1];
          const float x = a[ravel(ca, shape_a, 4)], w = b[ravel(cb, shape_b, 4)];
          if (fma_mode) sum = fmaf(x, w, sum);
          else { volatile float prod = x * w; sum += prod; }
        }
      }
    }
    out[index] = sum;
  }
}

void dot_general_ref(float* out, const float* a, const float* b, uint32_t N, uint32_t C, uint32_t M, uint32_t cdim_a,
                     uint32_t cdim_b, int fma_mode, int64_t begin, int64_t end) {
  const uint64_t stride_a = cdim_a ? 1 : N, stride_b = cdim_b ? 1 : M;
  for (int64_t index = begin; index < end; ++index) {
    const uint64_t row = (uint64_t)index / M, col = (uint64_t)index % M;
    const uint64_t off_a = cdim_a ? row * C : row, off_b = cdim_b ? col * C : col;
    float sum = 0.0f;
    for (uint32_t i = 0; i < C; ++i) {
      const float x = a[off_a + i * stride_a], w = b[off_b + i * stride_b];
      if (fma_mode) sum = fmaf(x, w, sum);
      else { volatile float prod = x * w; sum += prod; }
    }
    out[index] = sum;
  }
}

void reduce_window_max_ref(float* out, const float* a, const uint32_t* shape_a, const uint32_t* shape_out,
                           const uint32_t* padding, const uint32_t* strides, const uint32_t* window, int q3, int64_t begin, int64_t end) {
  const uint64_t wsize = (uint64_t)window[0] * window[1] * window[2] * window[3];
  for (int64_t index = begin; index < end; ++index) {
    int64_t co[4], cw[4], ca[4];
    unravel((uint64_t)index, shape_out, 4, co);
    float acc = -INFINITY;
    for (uint64_t i = 0; i < wsize; ++i) {
      unravel(i, window, 4, cw);
      int oob = 0;
      for (int d = 0; d < 4; ++d) {
        ca[d] = co[d] * strides[d] + cw[d] - (int64_t)padding[d];
        oob |= (ca[d] < 0 || ca[d] >= (int64_t)shape_a[d]);
      }
      if (oob) { if (q3) acc = fmaxf(acc, 0.0f); }
      else acc = fmaxf(acc, a[ravel(ca, shape_a, 4)]);
    }
    out[index] = acc;
  }
}

void reduce_sum_ref(float* out, const float* a, uint64_t outer, uint64_t red, uint64_t inner, int64_t begin, int64_t end) {
  (void)outer;
  for (int64_t index = begin; index < end; ++index) {
    const uint64_t o = (uint64_t)index / inner, i = (uint64_t)index % inner;
    float acc = 0.0f;
    for (uint64_t r = 0; r < red; ++r) acc += a[(o * red + r) * inner + i];
    out[index] = acc;
  }
}
