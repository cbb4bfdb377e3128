// Micro-benchmark for round 2 (written without a GPU): does the ACCESS PATTERN of the conv epilogue explain why write-heavy /
// residual 1x1 layers reach only 0.79-0.84 of the copy bandwidth while read-heavy ones reach 0.95?
//
// Writes (and optionally reads a residual of) an [M, N] fp32 matrix the way conv_tc2_kernel's epilogue does -- persistent CTAs
// walking 128 x 128 tiles, 16 warps, each warp owning 32 rows x 64 columns of a tile as two 32 x 32 chunks, one store
// instruction = 4 rows x 128 bytes at a row pitch of N * 4 bytes -- against the same traffic issued as 2 rows x 256 bytes,
// 1 row x 512 bytes per instruction, and against a plain contiguous grid-stride copy.  No tensor work: pure LSU traffic.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/store_pattern experiments/store_pattern.cu && /tmp/store_pattern
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)

__device__ __forceinline__ float4 ld_stream4(const float* p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}

// ROWS_PER_INSTR in {4, 2, 1}: one warp store instruction covers ROWS_PER_INSTR rows x (512 / ROWS_PER_INSTR) bytes.
// A warp owns 32 rows x (BLOCK_N / 2) columns of each tile of its group (two groups alternate tiles, as in the kernel).
template <int ROWS_PER_INSTR, bool RES>
__global__ void __launch_bounds__(512) epilogue_pattern(float* __restrict__ out, const float* __restrict__ res, uint32_t M, uint32_t N) {
  constexpr int BLOCK_N = 128;
  constexpr int LANES_PER_ROW = 32 / ROWS_PER_INSTR;            // lanes along the columns
  constexpr int COLS_PER_INSTR = LANES_PER_ROW * 4;              // floats
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int grp = warp >> 3, q = warp & 3, half = (warp & 7) >> 2;
  const uint32_t tiles_n = N / BLOCK_N, num_tiles = ((M + 127) / 128) * tiles_n;
  const int cl = lane % LANES_PER_ROW, rl = lane / LANES_PER_ROW;
  uint32_t tile_i = 0;
  for (uint32_t t = blockIdx.x; t < num_tiles; t += gridDim.x, ++tile_i) {
    if ((tile_i & 1u) != (uint32_t)grp) continue;
    const uint32_t m0 = (t / tiles_n) * 128 + q * 32, n0 = (t % tiles_n) * BLOCK_N + half * 64;
    // 32 rows x 64 columns per warp
    for (int c0 = 0; c0 < 64; c0 += COLS_PER_INSTR > 64 ? 64 : COLS_PER_INSTR) {
#pragma unroll 8
      for (int r0 = 0; r0 < 32; r0 += ROWS_PER_INSTR) {
        const uint32_t m = m0 + r0 + rl, n = n0 + c0 + cl * 4;
        if (m < M && (cl * 4 + c0) < 64) {
          float4 v = make_float4(1.f, 2.f, 3.f, (float)m);
          if (RES) { const float4 r = ld_stream4(res + (uint64_t)m * N + n); v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w; }
          *reinterpret_cast<float4*>(out + (uint64_t)m * N + n) = v;
        }
      }
    }
  }
}

template <bool RES>
__global__ void __launch_bounds__(256) contiguous(float* __restrict__ out, const float* __restrict__ res, uint64_t n4) {
  for (uint64_t i = (uint64_t)blockIdx.x * 256 + threadIdx.x; i < n4; i += (uint64_t)gridDim.x * 256) {
    float4 v = make_float4(1.f, 2.f, 3.f, (float)i);
    if (RES) { const float4 r = ld_stream4(res + i * 4); v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w; }
    reinterpret_cast<float4*>(out)[i] = v;
  }
}

template <typename F> static float time_ms(F launch, int reps = 20) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int i = 0; i < 3; ++i) launch();
  cudaEventRecord(e0);
  for (int i = 0; i < reps; ++i) launch();
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  return ms / reps;
}

int main() {
  const uint32_t M = 802816;                       // ResNet-50 b256 stage 0: 256 x 56 x 56 output pixels
  int sms = 0;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  for (uint32_t N : {256u, 512u, 1024u}) {
    const uint32_t Mn = M * 256 / N;               // same number of bytes for every N
    const uint64_t elems = (uint64_t)Mn * N;
    float *out = nullptr, *res = nullptr;
    CK(cudaMalloc(&out, elems * 4));
    CK(cudaMalloc(&res, elems * 4));
    CK(cudaMemset(res, 0, elems * 4));
    const double gb_w = elems * 4 / 1e9, gb_rw = 2 * gb_w;
    printf("N = %u (row pitch %u B), M = %u, %.0f MB per matrix, %d SMs\n", N, N * 4, Mn, gb_w * 1e3, sms);
#define RUN(NAME, EXPR, GB) { float ms = time_ms([&] { EXPR; }); printf("  %-46s %.4f ms  %6.0f GB/s\n", NAME, ms, (GB) / ms * 1e3); }
    RUN("write, contiguous grid-stride", (contiguous<false><<<sms * 16, 256>>>(out, res, elems / 4)), gb_w);
    RUN("write, epilogue pattern 4 rows x 128 B", (epilogue_pattern<4, false><<<sms, 512>>>(out, res, Mn, N)), gb_w);
    RUN("write, epilogue pattern 2 rows x 256 B", (epilogue_pattern<2, false><<<sms, 512>>>(out, res, Mn, N)), gb_w);
    RUN("write, epilogue pattern 1 row x 256 B (x2)", (epilogue_pattern<1, false><<<sms, 512>>>(out, res, Mn, N)), gb_w);
    RUN("residual + write, contiguous", (contiguous<true><<<sms * 16, 256>>>(out, res, elems / 4)), gb_rw);
    RUN("residual + write, epilogue pattern 4 x 128 B", (epilogue_pattern<4, true><<<sms, 512>>>(out, res, Mn, N)), gb_rw);
    RUN("residual + write, epilogue pattern 2 x 256 B", (epilogue_pattern<2, true><<<sms, 512>>>(out, res, Mn, N)), gb_rw);
    RUN("residual + write, epilogue pattern 1 x 256 B", (epilogue_pattern<1, true><<<sms, 512>>>(out, res, Mn, N)), gb_rw);
    RUN("residual + write, 4 x 128 B, 2 CTAs per SM", (epilogue_pattern<4, true><<<sms * 2, 512>>>(out, res, Mn, N)), gb_rw);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    cudaFree(out); cudaFree(res);
  }
  return 0;
}
