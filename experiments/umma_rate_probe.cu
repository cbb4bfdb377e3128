// Probe: issue rate of tcgen05.mma kind::tf32 as a function of the tile width N, the A-operand source (shared memory / TMEM) and the
// number of co-resident CTAs per SM -- to settle whether the "narrow-tile ceiling" of the N = 64 layers is the
// tensor core's own rate.  One thread per CTA issues a long chain of M = 128 MMAs on zero operands (no loads, no epilogue) and the
// kernel reports cycles per MMA; TFLOP/s = 148 SMs x ctas x 2 * 128 * N * 8 / cycles * clock.
//   nvcc -gencode arch=compute_100a,code=sm_100a -o umma_rate_probe umma_rate_probe.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {      // K-major SWIZZLE_128B, SBO 1024 B (as vkjax_b200/csrc/gemm_tc.cuh)
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
__host__ __device__ constexpr uint32_t idesc_tf32(int m, int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

template <int N, bool A_TMEM, int COLS, bool WARP>
__global__ void __launch_bounds__(128) probe(int iters, int n_acc, long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  // A tile 128 x 32 floats (16 KB) at base, B tile N x 32 floats behind it, barrier + tmem slot at the end
  const uint32_t a_s = base, b_s = base + 16384, bar = base + 16384 + N * 128, slot = bar + 8;
  for (uint32_t i = threadIdx.x; i < (16384u + N * 128u) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(gen)[i] = 0u;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot), "n"(COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(gen + (slot - base));
  if (WARP ? threadIdx.x < 32 : threadIdx.x == 0) {
    // WARP: the whole warp runs the loop (warp-uniform control flow: operands can live in uniform registers) and one elected lane
    // issues; otherwise a single thread in a divergent branch does everything (how libb2jax.so issued its MMAs until round 2)
    const uint64_t ad = make_desc(a_s), bd = make_desc(b_s);
    constexpr uint32_t id = idesc_tf32(128, N);
    const uint32_t a_t = tmem + (uint32_t)(COLS - 32);           // TMEM A operand: the last 32 columns (contents irrelevant)
    uint32_t leader = 1;
    if (WARP) asm volatile("{\n\t.reg .pred P1;\n\telect.sync _|P1, 0xffffffff;\n\tselp.b32 %0, 1, 0, P1;\n\t}" : "=r"(leader));
    const long long t0 = clock64();
#pragma unroll 4
    for (int i = 0; i < iters; ++i) {
      const uint32_t d = tmem + (uint32_t)((i & (n_acc - 1)) * N);  // n_acc (1 or 2) accumulators in turn (1: one dependent chain); no division in the loop
      const uint32_t k = (uint32_t)(i & 3);
      if (leader) {
        if (A_TMEM)
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
                       ::"r"(d), "r"(a_t + 8u * k), "l"(bd + 2ull * k), "r"(id), "r"(1u) : "memory");
        else
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                       ::"r"(d), "l"(ad + 2ull * k), "l"(bd + 2ull * k), "r"(id), "r"(1u) : "memory");
      }
      if (WARP) __syncwarp();
    }
    if (leader) {
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
      uint32_t ok = 0;
      while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar) : "memory");
      const long long t1 = clock64();
      if (blockIdx.x == 0) cycles[0] = t1 - t0;
    }
    if (WARP) __syncwarp();
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(COLS) : "memory");
}

template <int N, bool A_TMEM, int COLS, bool WARP = false>
static void run(const char* name, int ctas_per_sm, int n_acc, int sms, double ghz) {
  long long* d; cudaMalloc(&d, 8);
  const int smem = 16384 + N * 128 + 1024 + 64;
  cudaFuncSetAttribute(probe<N, A_TMEM, COLS, WARP>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int iters = 1 << 15;
  for (int rep = 0; rep < 2; ++rep) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    probe<N, A_TMEM, COLS, WARP><<<sms * ctas_per_sm, 128, smem>>>(iters, n_acc, d);
    cudaEventRecord(e1);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); exit(1); }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long c; cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
    if (rep == 1) {
      const double flops = 2.0 * 128 * N * 8 * iters * (double)sms * ctas_per_sm;
      printf("%-44s ctas/SM %d acc %d: %7.1f cycles/MMA (clock64), %8.1f TFLOP/s by event time (%.3f ms)\n", name, ctas_per_sm, n_acc,
             (double)c / iters, flops / (ms * 1e-3) / 1e12, ms);
    }
  }
  cudaFree(d);
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  const int sms = p.multiProcessorCount; const double ghz = p.clockRate / 1e6;
  printf("%s, %d SMs, %.2f GHz\n", p.name, sms, ghz);
  printf("--- one thread in a divergent branch issues\n");
  run<64, false, 128>("tf32 M=128 N=64  A from shared memory", 1, 1, sms, ghz);
  run<64, false, 128>("tf32 M=128 N=64  A from shared memory", 2, 1, sms, ghz);
  run<64, true, 256>("tf32 M=128 N=64  A from TMEM", 1, 1, sms, ghz);
  run<128, false, 256>("tf32 M=128 N=128 A from shared memory", 1, 1, sms, ghz);
  run<128, false, 256>("tf32 M=128 N=128 A from shared memory", 2, 1, sms, ghz);
  run<128, true, 512>("tf32 M=128 N=128 A from TMEM", 1, 1, sms, ghz);
  run<256, false, 512>("tf32 M=128 N=256 A from shared memory", 1, 1, sms, ghz);
  run<256, true, 512>("tf32 M=128 N=256 A from TMEM", 1, 1, sms, ghz);
  printf("--- the whole warp runs the loop, one elected lane issues\n");
  run<64, false, 128, true>("tf32 M=128 N=64  A from shared memory", 1, 1, sms, ghz);
  run<64, false, 128, true>("tf32 M=128 N=64  A from shared memory", 1, 2, sms, ghz);
  run<64, false, 128, true>("tf32 M=128 N=64  A from shared memory", 2, 1, sms, ghz);
  run<64, true, 256, true>("tf32 M=128 N=64  A from TMEM", 1, 1, sms, ghz);
  run<64, true, 256, true>("tf32 M=128 N=64  A from TMEM", 2, 1, sms, ghz);
  run<128, false, 256, true>("tf32 M=128 N=128 A from shared memory", 1, 1, sms, ghz);
  run<128, false, 256, true>("tf32 M=128 N=128 A from shared memory", 1, 2, sms, ghz);
  run<128, false, 256, true>("tf32 M=128 N=128 A from shared memory", 2, 1, sms, ghz);
  run<128, true, 512, true>("tf32 M=128 N=128 A from TMEM", 1, 1, sms, ghz);
  run<256, false, 512, true>("tf32 M=128 N=256 A from shared memory", 1, 1, sms, ghz);
  run<256, false, 512, true>("tf32 M=128 N=256 A from shared memory", 1, 2, sms, ghz);
  run<256, true, 512, true>("tf32 M=128 N=256 A from TMEM", 1, 1, sms, ghz);
  return 0;
}
