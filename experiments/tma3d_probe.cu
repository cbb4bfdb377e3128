// Probe: which 3-D tiled, non-swizzled TMA loads does the hardware accept?  (B2J_CT_ROWS staging; round 2)
// nvcc -gencode arch=compute_100a,code=sm_100a -o tma3d_probe tma3d_probe.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>

__global__ void probe(const __grid_constant__ CUtensorMap map, int c0, int c1, int c2, uint32_t bytes, int lanes, uint32_t lane_stride_bytes, int elems_per_lane, float* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint32_t sbase = (uint32_t)__cvta_generic_to_shared(smem);
  sbase = (sbase + 1023u) & ~1023u;
  uint8_t* gen = smem + (sbase - (uint32_t)__cvta_generic_to_shared(smem));
  const uint32_t bar = sbase + 65536;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes * lanes) : "memory");
  __syncwarp();
  if ((int)threadIdx.x < lanes) {
    const uint32_t dst = sbase + threadIdx.x * lane_stride_bytes;
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(dst), "l"(&map), "r"(bar), "r"(c0 + (int)threadIdx.x * elems_per_lane), "r"(c1), "r"(c2) : "memory");
  }
  uint32_t ok = 0;
  while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar) : "memory");
  for (uint32_t i = threadIdx.x; i < bytes * lanes / 4; i += blockDim.x) out[i] = reinterpret_cast<const float*>(gen)[i];
}

int main() {
  const int W = 672, H = 32, N = 4;
  std::vector<float> h((size_t)W * H * N);
  for (size_t i = 0; i < h.size(); ++i) h[i] = (float)(i % 100000);
  float *d, *o;
  cudaMalloc(&d, h.size() * 4); cudaMalloc(&o, 1 << 20);
  cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
  typedef CUresult (*PFN)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                          CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  void* f = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q);
  PFN enc = (PFN)f;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  const int boxes[] = {32, 64, 128, 256};
  for (int bi = 0; bi < 4; ++bi)
    for (int lanes = 1; lanes <= 3; lanes += 2)
      for (int c0 = 0; c0 >= -12; c0 -= 4) {
        const int bw = boxes[bi];
        CUtensorMap m;
        cuuint64_t dims[3] = {W, H, N}; cuuint64_t strides[2] = {W * 4ull, (cuuint64_t)W * H * 4};
        cuuint32_t box[3] = {(cuuint32_t)bw, 1, 1}, es[3] = {1, 1, 1};
        CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("box %d: encode failed %d\n", bw, (int)r); continue; }
        probe<<<1, 64, 100 * 1024>>>(m, c0, c0 == -12 ? -1 : 1, 1, bw * 4, lanes, bw * 4, bw, o);
        cudaError_t e = cudaDeviceSynchronize();
        float v[4] = {0, 0, 0, 0};
        if (e == cudaSuccess) cudaMemcpy(v, o, 16, cudaMemcpyDeviceToHost);
        printf("box %3d floats, lanes %d, c0 %3d: %s  first = %.0f %.0f (expect %.0f)\n", bw, lanes, c0, cudaGetErrorString(e), v[0], v[1],
               c0 == -12 ? 0.0f : c0 < 0 ? 0.0f : h[(size_t)(1 * H + 1) * W]);
        if (e != cudaSuccess) { printf("sticky error, stopping\n"); return 1; }
      }
  return 0;
}
