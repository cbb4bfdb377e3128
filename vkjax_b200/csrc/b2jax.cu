// libb2jax.so -- runtime + kernel dispatch behind include/b2jax.h.
//
// Plays the role Vulkan Kompute (`kp`, C++/pybind11, un-vendored) plays for the reference
// (SURVEY.md §2.1): device context, device memory, a recorded-once / replayed-per-call program.
// B200-native choices: one CUDA stream per context; memory from a stream-ordered pool
// (cudaMallocAsync); the recorded program is captured once into a CUDA Graph and replayed with a
// single cudaGraphLaunch (≙ kp.Sequence.eval()); profiling mode replays op by op with CUDA events
// (≙ Vulkan timestamp queries); outputs of N ranks are all-gathered with NCCL over NVLink.
#include <cuda_runtime.h>
#include <cuda.h>
#include <dlfcn.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/b2jax.h"
#include "elementwise.cuh"
#include "data_movement.cuh"
#include "reduce.cuh"
#include "contraction_simt.cuh"
#include "gemm_tc.cuh"
#include "conv_tc2.cuh"
#include "conv_patch.cuh"


using namespace b2j;

// ------------------------------------------------------------------------------------------------
static thread_local std::string g_err;

struct b2j_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaDeviceProp prop{};
  std::string err;
  void* nccl_comm = nullptr;
  int nranks = 1, rank = 0;
  void* flush_buf = nullptr;
  size_t flush_bytes = 0;
  cudaStream_t copy_stream = nullptr;                       // copy lanes (b2j_lane_*)
  cudaStream_t down_stream = nullptr;                       // result downloads behind the context stream (b2j_lane_download)
  cudaEvent_t down_after = nullptr;
  cudaEvent_t lane_ready[B2J_COPY_LANES] = {}, lane_consumed[B2J_COPY_LANES] = {};
};

struct SeqOp {
  uint32_t kid = 0;
  std::vector<b2j_buf> bufs;
  std::vector<uint8_t> params;
  size_t bytes = 0;  // allgather
};

struct b2j_seq {
  b2j_ctx* ctx = nullptr;
  bool profiling = false;
  bool finalized = false;
  std::vector<SeqOp> ops;
  cudaGraph_t graph = nullptr;
  cudaGraphExec_t exec = nullptr;
  std::vector<cudaEvent_t> evs;  // profiling: ops.size()+1 events
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  int n_launches = 0;
  size_t eager_tail = 0;   // number of trailing collective ops issued eagerly after the graph
};

static int fail(b2j_ctx* ctx, int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_err = buf;
  if (ctx) ctx->err = buf;
  return code;
}

#define CU_CHECK(ctx, expr)                                                                          \
  do {                                                                                               \
    cudaError_t e_ = (expr);                                                                         \
    if (e_ != cudaSuccess) {                                                                         \
      int code_ = (e_ == cudaErrorMemoryAllocation) ? B2J_ENOMEM : B2J_ECUDA;                        \
      return fail(ctx, code_, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, __LINE__); \
    }                                                                                                \
  } while (0)

// ---- NCCL, loaded lazily so that single-GPU use has no NCCL dependency ---------------------------
typedef struct { char internal[128]; } nccl_uid;
struct NcclApi {
  void* lib = nullptr;
  int (*GetUniqueId)(nccl_uid*) = nullptr;
  int (*CommInitRank)(void**, int, nccl_uid, int) = nullptr;
  int (*CommDestroy)(void*) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr;
  int (*Broadcast)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};
static NcclApi g_nccl;

static int nccl_load(b2j_ctx* ctx) {
  if (g_nccl.lib) return B2J_OK;
  const char* names[] = {getenv("B2J_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
  for (const char* n : names) {
    if (!n) continue;
    g_nccl.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (g_nccl.lib) break;
  }
  if (!g_nccl.lib) return fail(ctx, B2J_ECUDA, "cannot dlopen libnccl.so.2 (set B2J_NCCL_LIB): %s", dlerror());
#define NCCL_SYM(field, name)                                                    \
  *(void**)(&g_nccl.field) = dlsym(g_nccl.lib, name);                            \
  if (!g_nccl.field) return fail(ctx, B2J_ECUDA, "libnccl: missing symbol %s", name);
  NCCL_SYM(GetUniqueId, "ncclGetUniqueId");
  NCCL_SYM(CommInitRank, "ncclCommInitRank");
  NCCL_SYM(CommDestroy, "ncclCommDestroy");
  NCCL_SYM(AllGather, "ncclAllGather");
  NCCL_SYM(Broadcast, "ncclBroadcast");
  NCCL_SYM(GetErrorString, "ncclGetErrorString");
#undef NCCL_SYM
  return B2J_OK;
}
#define NCCL_CHECK(ctx, expr)                                                                       \
  do {                                                                                              \
    int e_ = (expr);                                                                                \
    if (e_ != 0) return fail(ctx, B2J_ECUDA, "%s failed: %s", #expr, g_nccl.GetErrorString(e_));   \
  } while (0)

// ------------------------------------------------------------------------------------------------
extern "C" {

int b2j_abi_version(void) { return B2J_ABI_VERSION; }

const char* b2j_last_error(b2j_ctx* ctx) { return ctx ? ctx->err.c_str() : g_err.c_str(); }

int b2j_device_count(int* n) {
  CU_CHECK(nullptr, cudaGetDeviceCount(n));
  return B2J_OK;
}

int b2j_ctx_create(int device, b2j_ctx** out) {
  int n = 0;
  CU_CHECK(nullptr, cudaGetDeviceCount(&n));
  if (device < 0 || device >= n) return fail(nullptr, B2J_EINVAL, "device %d out of range (%d devices)", device, n);
  CU_CHECK(nullptr, cudaSetDevice(device));
  b2j_ctx* ctx = new b2j_ctx();
  ctx->device = device;
  cudaError_t e = cudaGetDeviceProperties(&ctx->prop, device);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
  if (e != cudaSuccess) {
    delete ctx;
    return fail(nullptr, B2J_ECUDA, "context creation failed: %s", cudaGetErrorString(e));
  }
  if (ctx->prop.major != 10) {
    // kernels are built for sm_100a only; refuse loudly rather than fall back
    fail(nullptr, B2J_ENOTIMPL, "device %d is sm_%d%d; libb2jax is built for sm_100a (B200) only", device,
         ctx->prop.major, ctx->prop.minor);
    cudaStreamDestroy(ctx->stream);
    delete ctx;
    return B2J_ENOTIMPL;
  }
  // keep freed blocks in the pool: sequences allocate/free arenas repeatedly
  cudaMemPool_t pool;
  if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
    uint64_t thr = UINT64_MAX;
    cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
  }
  *out = ctx;
  return B2J_OK;
}

int b2j_ctx_destroy(b2j_ctx* ctx) {
  if (!ctx) return B2J_OK;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  if (ctx->nccl_comm && g_nccl.CommDestroy) g_nccl.CommDestroy(ctx->nccl_comm);
  if (ctx->flush_buf) cudaFree(ctx->flush_buf);
  if (ctx->copy_stream) {
    cudaStreamSynchronize(ctx->copy_stream);
    for (int i = 0; i < B2J_COPY_LANES; ++i) { cudaEventDestroy(ctx->lane_ready[i]); cudaEventDestroy(ctx->lane_consumed[i]); }
    cudaStreamDestroy(ctx->copy_stream);
  }
  if (ctx->down_stream) {
    cudaStreamSynchronize(ctx->down_stream);
    cudaEventDestroy(ctx->down_after);
    cudaStreamDestroy(ctx->down_stream);
  }
  cudaStreamDestroy(ctx->stream);
  delete ctx;
  return B2J_OK;
}

int b2j_device_props(b2j_ctx* ctx, b2j_props* out) {
  memset(out, 0, sizeof *out);
  snprintf(out->name, sizeof out->name, "%s", ctx->prop.name);
  out->cc_major = ctx->prop.major;
  out->cc_minor = ctx->prop.minor;
  out->sm_count = ctx->prop.multiProcessorCount;
  out->max_threads_per_block = ctx->prop.maxThreadsPerBlock;
  out->max_block_dim_x = ctx->prop.maxThreadsDim[0];
  out->shared_mem_per_block_optin = ctx->prop.sharedMemPerBlockOptin;
  out->l2_bytes = ctx->prop.l2CacheSize;
  CU_CHECK(ctx, cudaSetDevice(ctx->device));
  CU_CHECK(ctx, cudaMemGetInfo(&out->free_mem, &out->total_mem));
  return B2J_OK;
}

int b2j_ctx_sync(b2j_ctx* ctx) {
  CU_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
  if (ctx->down_stream) CU_CHECK(ctx, cudaStreamSynchronize(ctx->down_stream));
  return B2J_OK;
}

// ---- memory ------------------------------------------------------------------------------------
int b2j_mem_alloc(b2j_ctx* ctx, size_t bytes, b2j_buf* out) {
  void* p = nullptr;
  CU_CHECK(ctx, cudaSetDevice(ctx->device));
  if (bytes == 0) bytes = 256;
  CU_CHECK(ctx, cudaMallocAsync(&p, bytes, ctx->stream));
  *out = (b2j_buf)(uintptr_t)p;
  return B2J_OK;
}

int b2j_mem_free(b2j_ctx* ctx, b2j_buf buf) {
  if (!buf) return B2J_OK;
  CU_CHECK(ctx, cudaFreeAsync((void*)(uintptr_t)buf, ctx->stream));
  return B2J_OK;
}

int b2j_mem_set(b2j_ctx* ctx, b2j_buf buf, int byte, size_t bytes) {
  CU_CHECK(ctx, cudaMemsetAsync((void*)(uintptr_t)buf, byte, bytes, ctx->stream));
  return B2J_OK;
}

int b2j_host_alloc(b2j_ctx* ctx, size_t bytes, void** out) {
  CU_CHECK(ctx, cudaSetDevice(ctx->device));
  CU_CHECK(ctx, cudaHostAlloc(out, bytes ? bytes : 64, cudaHostAllocDefault));
  return B2J_OK;
}

int b2j_host_free(b2j_ctx* ctx, void* p) {
  if (p) CU_CHECK(ctx, cudaFreeHost(p));
  return B2J_OK;
}

int b2j_upload_async(b2j_ctx* ctx, b2j_buf dst, const void* host, size_t bytes) {
  if (bytes) CU_CHECK(ctx, cudaMemcpyAsync((void*)(uintptr_t)dst, host, bytes, cudaMemcpyHostToDevice, ctx->stream));
  return B2J_OK;
}

int b2j_download_async(b2j_ctx* ctx, b2j_buf src, void* host, size_t bytes) {
  if (bytes) CU_CHECK(ctx, cudaMemcpyAsync(host, (const void*)(uintptr_t)src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  return B2J_OK;
}

int b2j_upload(b2j_ctx* ctx, b2j_buf dst, const void* host, size_t bytes) {
  int rc = b2j_upload_async(ctx, dst, host, bytes);
  if (rc) return rc;
  CU_CHECK(ctx, cudaStreamSynchronize(ctx->stream));   // pageable source: the copy must finish before we return
  return B2J_OK;
}

int b2j_download(b2j_ctx* ctx, b2j_buf src, void* host, size_t bytes) {
  int rc = b2j_download_async(ctx, src, host, bytes);
  if (rc) return rc;
  CU_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
  return B2J_OK;
}

int b2j_copy_async(b2j_ctx* ctx, b2j_buf dst, b2j_buf src, size_t bytes) {
  if (bytes)
    CU_CHECK(ctx, cudaMemcpyAsync((void*)(uintptr_t)dst, (const void*)(uintptr_t)src, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
  return B2J_OK;
}

// ---- copy lanes: uploads on a second stream, ordered against the context stream by events ------------
static int lanes_init(b2j_ctx* ctx) {
  if (ctx->copy_stream) return B2J_OK;
  CU_CHECK(ctx, cudaSetDevice(ctx->device));
  CU_CHECK(ctx, cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
  for (int i = 0; i < B2J_COPY_LANES; ++i) {
    CU_CHECK(ctx, cudaEventCreateWithFlags(&ctx->lane_ready[i], cudaEventDisableTiming));
    CU_CHECK(ctx, cudaEventCreateWithFlags(&ctx->lane_consumed[i], cudaEventDisableTiming));
  }
  return B2J_OK;
}

int b2j_lane_upload(b2j_ctx* ctx, int lane, b2j_buf dst, const void* pinned, size_t bytes) {
  if (lane < 0 || lane >= B2J_COPY_LANES) return fail(ctx, B2J_EINVAL, "copy lane %d out of range", lane);
  int rc = lanes_init(ctx);
  if (rc) return rc;
  CU_CHECK(ctx, cudaStreamWaitEvent(ctx->copy_stream, ctx->lane_consumed[lane], 0));   // never-recorded event: no wait
  if (bytes) CU_CHECK(ctx, cudaMemcpyAsync((void*)(uintptr_t)dst, pinned, bytes, cudaMemcpyHostToDevice, ctx->copy_stream));
  CU_CHECK(ctx, cudaEventRecord(ctx->lane_ready[lane], ctx->copy_stream));
  return B2J_OK;
}

int b2j_lane_acquire(b2j_ctx* ctx, int lane) {
  if (lane < 0 || lane >= B2J_COPY_LANES || !ctx->copy_stream) return fail(ctx, B2J_EINVAL, "copy lane %d: nothing was uploaded", lane);
  CU_CHECK(ctx, cudaStreamWaitEvent(ctx->stream, ctx->lane_ready[lane], 0));
  return B2J_OK;
}

int b2j_lane_release(b2j_ctx* ctx, int lane) {
  if (lane < 0 || lane >= B2J_COPY_LANES || !ctx->copy_stream) return fail(ctx, B2J_EINVAL, "copy lane %d: nothing was uploaded", lane);
  CU_CHECK(ctx, cudaEventRecord(ctx->lane_consumed[lane], ctx->stream));
  return B2J_OK;
}

// Device -> host copy on a third stream, ordered after everything enqueued on the context stream so far: the context stream
// does not wait for it (a D2H copy on the context stream itself delays the next replay by its own duration -- 8 MB of gathered
// logits cost 0.35 ms per step on the root rank at N = 8).  The caller must keep `src` unchanged until the copy has completed
// (b2j_lane_download_record + b2j_event_sync); JaxprInterpreter.run_many copies the outputs into a ring of device staging
// buffers first.
int b2j_lane_download(b2j_ctx* ctx, void* pinned, b2j_buf src, size_t bytes) {
  if (!ctx->down_stream) {
    CU_CHECK(ctx, cudaSetDevice(ctx->device));
    CU_CHECK(ctx, cudaStreamCreateWithFlags(&ctx->down_stream, cudaStreamNonBlocking));
    CU_CHECK(ctx, cudaEventCreateWithFlags(&ctx->down_after, cudaEventDisableTiming));
  }
  CU_CHECK(ctx, cudaEventRecord(ctx->down_after, ctx->stream));
  CU_CHECK(ctx, cudaStreamWaitEvent(ctx->down_stream, ctx->down_after, 0));
  if (bytes) CU_CHECK(ctx, cudaMemcpyAsync(pinned, (const void*)(uintptr_t)src, bytes, cudaMemcpyDeviceToHost, ctx->down_stream));
  return B2J_OK;
}
// Records `ev` (b2j_event_create) behind the downloads issued so far.
int b2j_lane_download_record(b2j_ctx* ctx, void* ev) {
  if (!ctx->down_stream) return fail(ctx, B2J_EINVAL, "b2j_lane_download_record: nothing was downloaded");
  CU_CHECK(ctx, cudaEventRecord((cudaEvent_t)ev, ctx->down_stream));
  return B2J_OK;
}

int b2j_lane_sync(b2j_ctx* ctx, int lane) {
  if (lane < 0 || lane >= B2J_COPY_LANES) return fail(ctx, B2J_EINVAL, "copy lane %d out of range", lane);
  if (ctx->copy_stream) CU_CHECK(ctx, cudaEventSynchronize(ctx->lane_ready[lane]));
  return B2J_OK;
}

// ---- events / L2 flush ---------------------------------------------------------------------------
int b2j_event_create(b2j_ctx* ctx, void** ev) {
  cudaEvent_t e;
  CU_CHECK(ctx, cudaEventCreate(&e));
  *ev = e;
  return B2J_OK;
}
int b2j_event_record(b2j_ctx* ctx, void* ev) {
  CU_CHECK(ctx, cudaEventRecord((cudaEvent_t)ev, ctx->stream));
  return B2J_OK;
}
int b2j_event_elapsed_ms(b2j_ctx* ctx, void* start, void* stop, float* ms) {
  CU_CHECK(ctx, cudaEventSynchronize((cudaEvent_t)stop));
  CU_CHECK(ctx, cudaEventElapsedTime(ms, (cudaEvent_t)start, (cudaEvent_t)stop));
  return B2J_OK;
}
int b2j_event_sync(b2j_ctx* ctx, void* ev) {
  CU_CHECK(ctx, cudaEventSynchronize((cudaEvent_t)ev));
  return B2J_OK;
}
int b2j_event_destroy(b2j_ctx* ctx, void* ev) {
  CU_CHECK(ctx, cudaEventDestroy((cudaEvent_t)ev));
  return B2J_OK;
}
int b2j_flush_l2(b2j_ctx* ctx) {
  if (!ctx->flush_buf) {
    ctx->flush_bytes = (size_t)ctx->prop.l2CacheSize * 2;
    if (ctx->flush_bytes < (256u << 20)) ctx->flush_bytes = 256u << 20;
    CU_CHECK(ctx, cudaMalloc(&ctx->flush_buf, ctx->flush_bytes));
  }
  CU_CHECK(ctx, cudaMemsetAsync(ctx->flush_buf, 0, ctx->flush_bytes, ctx->stream));
  return B2J_OK;
}

// ---- multi-GPU -------------------------------------------------------------------------------------
int b2j_nccl_unique_id(void* out128) {
  int rc = nccl_load(nullptr);
  if (rc) return rc;
  nccl_uid id;
  NCCL_CHECK(nullptr, g_nccl.GetUniqueId(&id));
  memcpy(out128, &id, 128);
  return B2J_OK;
}

int b2j_comm_init(b2j_ctx* ctx, int nranks, int rank, const void* id128) {
  int rc = nccl_load(ctx);
  if (rc) return rc;
  CU_CHECK(ctx, cudaSetDevice(ctx->device));
  nccl_uid id;
  memcpy(&id, id128, 128);
  NCCL_CHECK(ctx, g_nccl.CommInitRank(&ctx->nccl_comm, nranks, id, rank));
  ctx->nranks = nranks;
  ctx->rank = rank;
  return B2J_OK;
}

int b2j_comm_destroy(b2j_ctx* ctx) {
  if (ctx->nccl_comm) {
    NCCL_CHECK(ctx, g_nccl.CommDestroy(ctx->nccl_comm));
    ctx->nccl_comm = nullptr;
  }
  return B2J_OK;
}

static int allgather_on(b2j_ctx* ctx, b2j_buf send, b2j_buf recv, size_t bytes, cudaStream_t st) {
  if (!ctx->nccl_comm) {
    if (ctx->nranks == 1) {   // single rank: gather == copy
      CU_CHECK(ctx, cudaMemcpyAsync((void*)(uintptr_t)recv, (const void*)(uintptr_t)send, bytes, cudaMemcpyDeviceToDevice, st));
      return B2J_OK;
    }
    return fail(ctx, B2J_EINVAL, "b2j_allgather: communicator not initialised");
  }
  NCCL_CHECK(ctx, g_nccl.AllGather((const void*)(uintptr_t)send, (void*)(uintptr_t)recv, bytes, /*ncclInt8*/ 0, ctx->nccl_comm, st));
  return B2J_OK;
}

int b2j_allgather(b2j_ctx* ctx, b2j_buf send, b2j_buf recv, size_t bytes_per_rank) {
  return allgather_on(ctx, send, recv, bytes_per_rank, ctx->stream);
}

int b2j_broadcast(b2j_ctx* ctx, b2j_buf buf, size_t bytes, int root) {
  if (!ctx->nccl_comm) return ctx->nranks == 1 ? B2J_OK : fail(ctx, B2J_EINVAL, "b2j_broadcast: communicator not initialised");
  NCCL_CHECK(ctx, g_nccl.Broadcast((const void*)(uintptr_t)buf, (void*)(uintptr_t)buf, bytes, 0, root, ctx->nccl_comm, ctx->stream));
  return B2J_OK;
}

// ---- kernel dispatch -------------------------------------------------------------------------------
size_t b2j_param_size(uint32_t kid) {
  switch (kid) {
    case B2J_K_ELTWISE: return sizeof(b2j_elt_params);
    case B2J_K_STRIDED_COPY: return sizeof(b2j_strided_params);
    case B2J_K_TRANSPOSE2D: return sizeof(b2j_transpose_params);
    case B2J_K_REDUCE: return sizeof(b2j_reduce_params);
    case B2J_K_REDUCE_WINDOW: return sizeof(b2j_reduce_window_params);
    case B2J_K_CONV_DIRECT: return sizeof(b2j_conv_direct_params);
    case B2J_K_DOT: return sizeof(b2j_dot_params);
    case B2J_K_CONV_TC: return sizeof(b2j_conv_tc_params);
    case B2J_K_WEIGHT_PREP: return sizeof(b2j_weight_prep_params);
    case B2J_K_GATHER: return sizeof(b2j_gather_params);
    case B2J_K_SCATTER_ADD: return sizeof(b2j_scatter_params);
    case B2J_K_CONCAT: return sizeof(b2j_concat_params);
    case B2J_K_THREEFRY: return sizeof(b2j_threefry_params);
    case B2J_K_GEMM_TC: return sizeof(b2j_gemm_tc_params);
    case B2J_K_RELAYOUT: return sizeof(b2j_relayout_params);
    case B2J_K_DILATE: return sizeof(b2j_dilate_params);
    case B2J_K_SELECT_SCATTER_ADD: return sizeof(b2j_reduce_window_params);
    default: return 0;
  }
}

}  // extern "C"

static inline unsigned grid_for(uint64_t work_items, int block, const b2j_ctx* ctx, int waves = 16) {
  // enough CTAs to cover the work, capped at a multiple of the SM count (grid-stride loops inside)
  uint64_t need = (work_items + block - 1) / block;
  uint64_t cap = (uint64_t)ctx->prop.multiProcessorCount * waves;
  if (need < 1) need = 1;
  return (unsigned)(need < cap ? need : cap);
}

template <typename T> static T* P(b2j_buf b) { return reinterpret_cast<T*>((uintptr_t)b); }

static int fill_epi(b2j_ctx* ctx, const b2j_epilogue& e, const SeqOp& op, EpiPtrs* out) {
  if (e.n_steps > B2J_EPI_MAX_STEPS) return fail(ctx, B2J_EINVAL, "epilogue: too many steps");
  for (uint32_t s = 0; s < B2J_EPI_MAX_STEPS; ++s) out->p[s] = nullptr;
  for (uint32_t s = 0; s < e.n_steps; ++s) {
    const b2j_epi_step& st = e.steps[s];
    if (st.kind == B2J_EPK_IMM) continue;
    if (st.buf >= op.bufs.size()) return fail(ctx, B2J_EINVAL, "epilogue: operand buffer index out of range");
    out->p[s] = P<const float>(op.bufs[st.buf]);
  }
  return B2J_OK;
}

template <typename T, int KIND>
static void launch_reduce_t(const b2j_reduce_params& p, const SeqOp& op, b2j_ctx* ctx, cudaStream_t st) {
  const bool block_path = p.n_red >= 1024 && p.n_out <= (uint64_t)ctx->prop.multiProcessorCount * 64;
  // the reduced run is contiguous in memory and there are many outputs: one warp per output (coalesced 128-bit loads)
  const bool warp_path = !block_path && p.red_rank == 1 && p.red_strides[0] == 1 && p.n_red >= 64;
  if (warp_path) {
    const uint64_t blocks = (p.n_out + 7) / 8, cap = (uint64_t)ctx->prop.multiProcessorCount * 32;
    reduce_warp_kernel<T, KIND><<<(unsigned)(blocks < cap ? blocks : cap), 256, 0, st>>>(p, P<uint32_t>(op.bufs[0]), P<const T>(op.bufs[1]));
  } else if (block_path) {
    unsigned grid = (unsigned)(p.n_out < 65535 ? p.n_out : 65535);
    reduce_block_kernel<T, KIND><<<grid, 256, 0, st>>>(p, P<uint32_t>(op.bufs[0]), P<const T>(op.bufs[1]));
  } else if constexpr (KIND == B2J_RED_MAX || KIND == B2J_RED_MIN || KIND >= B2J_RED_ARGMAX) {
    // a long strided run and too few outputs to fill the SMs with one thread each: 8 row slices per output
    if (p.red_rank == 1 && p.n_red >= 512 && p.n_out < (uint64_t)ctx->prop.multiProcessorCount * 1024) {
      const uint64_t blocks = (p.n_out + 31) / 32, cap = (uint64_t)ctx->prop.multiProcessorCount * 16;
      reduce_sliced_kernel<T, KIND><<<(unsigned)(blocks < cap ? blocks : cap), 256, 0, st>>>(p, P<uint32_t>(op.bufs[0]), P<const T>(op.bufs[1]));
    } else {
      reduce_thread_kernel<T, KIND><<<grid_for(p.n_out, 256, ctx), 256, 0, st>>>(p, P<uint32_t>(op.bufs[0]), P<const T>(op.bufs[1]));
    }
  } else {
    reduce_thread_kernel<T, KIND><<<grid_for(p.n_out, 256, ctx), 256, 0, st>>>(p, P<uint32_t>(op.bufs[0]), P<const T>(op.bufs[1]));
  }
}

template <typename T>
static int launch_reduce(const b2j_reduce_params& p, const SeqOp& op, b2j_ctx* ctx, cudaStream_t st) {
  switch (p.kind) {
    case B2J_RED_SUM: launch_reduce_t<T, B2J_RED_SUM>(p, op, ctx, st); break;
    case B2J_RED_MAX: launch_reduce_t<T, B2J_RED_MAX>(p, op, ctx, st); break;
    case B2J_RED_MIN: launch_reduce_t<T, B2J_RED_MIN>(p, op, ctx, st); break;
    case B2J_RED_PROD: launch_reduce_t<T, B2J_RED_PROD>(p, op, ctx, st); break;
    case B2J_RED_ARGMAX: launch_reduce_t<T, B2J_RED_ARGMAX>(p, op, ctx, st); break;
    case B2J_RED_ARGMIN: launch_reduce_t<T, B2J_RED_ARGMIN>(p, op, ctx, st); break;
    default: return fail(ctx, B2J_ENOTIMPL, "reduce kind %u", p.kind);
  }
  return B2J_OK;
}

template <typename T>
static int launch_reduce_window(const b2j_reduce_window_params& p, const SeqOp& op, b2j_ctx* ctx, cudaStream_t st) {
  const bool vec = p.window[3] == 1 && p.strides[3] == 1 && p.pad_lo[3] == 0 && p.in_shape[3] == p.out_shape[3] &&
                   (p.out_shape[3] % 4 == 0);
  const uint64_t n = (uint64_t)p.out_shape[0] * p.out_shape[1] * p.out_shape[2] * (p.out_shape[3] / (vec ? 4 : 1));
  const unsigned grid = grid_for(n, 256, ctx, 64);
  T* out = P<T>(op.bufs[0]);
  const T* in = P<const T>(op.bufs[1]);
  // NHWC pooling: window (1, kh, kw, 1) with kh, kw in {2, 3} -> unrolled kernel
  const bool pool = vec && p.window[0] == 1 && p.strides[0] == 1 && p.pad_lo[0] == 0 && p.in_shape[0] == p.out_shape[0];
  const int kk = (int)(p.window[1] * 10 + p.window[2]);
  // 32-bit index math in the pooling kernel: element offsets (signed) and the grid-stride counter must not wrap;
  // B2J_POOL_IDX64=1 forces the 64-bit instantiation (tests)
  static int idx64 = -1;
  if (idx64 < 0) { const char* e = getenv("B2J_POOL_IDX64"); idx64 = (e && e[0] == '1') ? 1 : 0; }
  const uint64_t in_elems = (uint64_t)p.in_shape[0] * p.in_shape[1] * p.in_shape[2] * p.in_shape[3];
  const bool idx32 = !idx64 && in_elems + 4ull * p.in_shape[2] * p.in_shape[3] < (1ull << 31) && n + 256ull * 148 * 64 < (1ull << 32);
  // channel counts that are not a multiple of 4: element-per-thread pooling kernel (window (1, kh, kw, 1), 32-bit index space)
  const bool spool = !vec && p.window[0] == 1 && p.strides[0] == 1 && p.pad_lo[0] == 0 && p.in_shape[0] == p.out_shape[0] &&
                     p.window[3] == 1 && p.strides[3] == 1 && p.pad_lo[3] == 0 && p.in_shape[3] == p.out_shape[3] &&
                     in_elems + 4ull * p.in_shape[2] * p.in_shape[3] < (1ull << 31) && n + 256ull * 148 * 64 < (1ull << 31);
#define RW_LAUNCH(KIND)                                                                                          \
  if (spool && kk == 33) pool2d_scalar_kernel<T, KIND, 3, 3><<<grid, 256, 0, st>>>(p, out, in);                  \
  else if (spool && kk == 22) pool2d_scalar_kernel<T, KIND, 2, 2><<<grid, 256, 0, st>>>(p, out, in);             \
  else if (pool && kk == 33 && idx32) pool2d_kernel<T, KIND, 3, 3, uint32_t><<<grid, 256, 0, st>>>(p, out, in);       \
  else if (pool && kk == 22 && idx32) pool2d_kernel<T, KIND, 2, 2, uint32_t><<<grid, 256, 0, st>>>(p, out, in);  \
  else if (pool && kk == 33) pool2d_kernel<T, KIND, 3, 3, uint64_t><<<grid, 256, 0, st>>>(p, out, in);           \
  else if (pool && kk == 22) pool2d_kernel<T, KIND, 2, 2, uint64_t><<<grid, 256, 0, st>>>(p, out, in);           \
  else if (vec) reduce_window_kernel<T, KIND, 4><<<grid, 256, 0, st>>>(p, out, in);          \
  else reduce_window_kernel<T, KIND, 1><<<grid, 256, 0, st>>>(p, out, in);
  switch (p.kind) {
    case B2J_RW_MAX: RW_LAUNCH(B2J_RW_MAX); break;
    case B2J_RW_MIN: RW_LAUNCH(B2J_RW_MIN); break;
    case B2J_RW_SUM: RW_LAUNCH(B2J_RW_SUM); break;
    default: return fail(ctx, B2J_ENOTIMPL, "reduce_window kind %u", p.kind);
  }
#undef RW_LAUNCH
  return B2J_OK;
}

#define NEED_BUFS(n_)                                                                                       \
  if ((int)op.bufs.size() < (n_)) return fail(ctx, B2J_EINVAL, "kernel %u needs >= %d buffers, got %zu", op.kid, (n_), op.bufs.size())

static int launch_op(b2j_ctx* ctx, const SeqOp& op, cudaStream_t st, int* launches) {
  switch (op.kid) {
    case B2J_K_ELTWISE: {
      const b2j_elt_params& p = *reinterpret_cast<const b2j_elt_params*>(op.params.data());
      if (p.n_in > B2J_ELT_MAX_IN || p.n_steps > B2J_ELT_MAX_STEPS || p.rank > B2J_MAX_RANK)
        return fail(ctx, B2J_EINVAL, "eltwise: n_in/n_steps/rank out of range");
      NEED_BUFS(1 + (int)p.n_in);
      if (p.n == 0) return B2J_OK;
      EltPtrs ptrs{};
      ptrs.out = P<uint32_t>(op.bufs[0]);
      for (uint32_t i = 0; i < p.n_in; ++i) ptrs.in[i] = P<const uint32_t>(op.bufs[1 + i]);
      {
        // chains whose step operands are all narrow (immediates, scalars, per-channel vectors) run with 8 vectors per thread
        const bool narrow = elt_chain_is_narrow(p);
        const int vecs = narrow ? ELT_VECS_NARROW : ELT_VECS;
        const uint64_t tiles = ((p.n + 3) / 4 + (uint64_t)ELT_THREADS * vecs - 1) / ((uint64_t)ELT_THREADS * vecs);
        const uint64_t cap = (uint64_t)ctx->prop.multiProcessorCount * 16;
        const unsigned grid = (unsigned)(tiles < cap ? (tiles ? tiles : 1) : cap);
        if (narrow) eltwise_kernel<ELT_VECS_NARROW, true><<<grid, ELT_THREADS, 0, st>>>(p, ptrs);
        else eltwise_kernel<ELT_VECS, false><<<grid, ELT_THREADS, 0, st>>>(p, ptrs);
      }
      ++*launches;
    } break;
    case B2J_K_STRIDED_COPY: {
      const b2j_strided_params& p = *reinterpret_cast<const b2j_strided_params*>(op.params.data());
      NEED_BUFS(2);
      if (p.n == 0) return B2J_OK;
      {
        // vectorised path: innermost dim a multiple of 4 read with stride 1 (all other strides, the base and both
        // pointers 16-byte aligned) or stride 0; 32-bit index space
        const int64_t s_in = p.rank ? p.strides[p.rank - 1] : 1;
        bool vec = p.rank >= 1 && p.rank <= B2J_MAX_RANK && (p.shape[p.rank - 1] & 3u) == 0 && (s_in == 0 || s_in == 1) && p.n < (1ull << 32) &&
                   (op.bufs[0] & 15u) == 0;
        if (vec && s_in == 1) {
          vec = (op.bufs[1] & 15u) == 0 && (p.base & 3) == 0;
          for (uint32_t d = 0; d + 1 < p.rank; ++d) vec = vec && (p.strides[d] & 3) == 0;
        }
        if (vec) {
          const uint64_t tiles = ((p.n >> 2) + 256 * SC_VECS - 1) / (256 * SC_VECS);
          const uint64_t cap = (uint64_t)ctx->prop.multiProcessorCount * 16;
          strided_copy_vec4_kernel<<<(unsigned)(tiles < cap ? tiles : cap), 256, 0, st>>>(p, P<uint32_t>(op.bufs[0]), P<const uint32_t>(op.bufs[1]));
        } else {
          strided_copy_kernel<<<grid_for(p.n, 256, ctx, 32), 256, 0, st>>>(p, P<uint32_t>(op.bufs[0]), P<const uint32_t>(op.bufs[1]));
        }
      }
      ++*launches;
    } break;
    case B2J_K_TRANSPOSE2D: {
      const b2j_transpose_params& p = *reinterpret_cast<const b2j_transpose_params*>(op.params.data());
      NEED_BUFS(2);
      if (p.rows == 0 || p.cols == 0) return B2J_OK;
      const uint32_t nb = p.batch ? p.batch : 1u;
      const bool vec = (p.rows & 3u) == 0 && (p.cols & 3u) == 0 && ((op.bufs[0] | op.bufs[1]) & 15u) == 0;      // 128-bit accesses on both sides
      const uint32_t t = vec ? 64u : 32u;
      dim3 grid((p.cols + t - 1) / t, (p.rows + t - 1) / t, nb < 65535u ? nb : 65535u);
      if (grid.y > 65535u) return fail(ctx, B2J_ENOTIMPL, "transpose2d: more than 65535 row tiles");
      if (vec) transpose2d_vec_kernel<<<grid, 256, 0, st>>>(p, P<uint32_t>(op.bufs[0]), P<const uint32_t>(op.bufs[1]));
      else transpose2d_kernel<<<grid, 256, 0, st>>>(p, P<uint32_t>(op.bufs[0]), P<const uint32_t>(op.bufs[1]));
      ++*launches;
    } break;
    case B2J_K_REDUCE: {
      const b2j_reduce_params& p = *reinterpret_cast<const b2j_reduce_params*>(op.params.data());
      NEED_BUFS(2);
      if (p.n_out == 0) return B2J_OK;
      int rc;
      if (p.dtype == B2J_F32) rc = launch_reduce<float>(p, op, ctx, st);
      else if (p.dtype == B2J_I32) rc = launch_reduce<int32_t>(p, op, ctx, st);
      else if (p.dtype == B2J_U32 || p.dtype == B2J_BOOL) rc = launch_reduce<uint32_t>(p, op, ctx, st);
      else return fail(ctx, B2J_ENOTIMPL, "reduce dtype %u", p.dtype);
      if (rc) return rc;
      ++*launches;
    } break;
    case B2J_K_REDUCE_WINDOW: {
      const b2j_reduce_window_params& p = *reinterpret_cast<const b2j_reduce_window_params*>(op.params.data());
      NEED_BUFS(2);
      int rc;
      if (p.dtype == B2J_F32) rc = launch_reduce_window<float>(p, op, ctx, st);
      else if (p.dtype == B2J_I32) rc = launch_reduce_window<int32_t>(p, op, ctx, st);
      else if (p.dtype == B2J_U32) rc = launch_reduce_window<uint32_t>(p, op, ctx, st);
      else return fail(ctx, B2J_ENOTIMPL, "reduce_window dtype %u", p.dtype);
      if (rc) return rc;
      ++*launches;
    } break;
    case B2J_K_CONV_DIRECT: {
      const b2j_conv_direct_params& p = *reinterpret_cast<const b2j_conv_direct_params*>(op.params.data());
      NEED_BUFS(3);
      EpiPtrs epi;
      int rc = fill_epi(ctx, p.epi, op, &epi);
      if (rc) return rc;
      const uint64_t n = (uint64_t)p.out_shape[0] * p.out_shape[1] * p.out_shape[2] * p.out_shape[3];
      if (n == 0) return B2J_OK;
      conv_direct_kernel<<<grid_for(n, 256, ctx, 64), 256, 0, st>>>(p, epi, P<float>(op.bufs[0]), P<const float>(op.bufs[1]),
                                                                    P<const float>(op.bufs[2]));
      ++*launches;
    } break;
    case B2J_K_DOT: {
      const b2j_dot_params& p = *reinterpret_cast<const b2j_dot_params*>(op.params.data());
      NEED_BUFS(3);
      EpiPtrs epi;
      int rc = fill_epi(ctx, p.epi, op, &epi);
      if (rc) return rc;
      const uint64_t n = (uint64_t)p.n * p.m;
      if (n == 0) return B2J_OK;
      dot_kernel<<<grid_for(n, 256, ctx, 64), 256, 0, st>>>(p, epi, P<float>(op.bufs[0]), P<const float>(op.bufs[1]),
                                                            P<const float>(op.bufs[2]));
      ++*launches;
    } break;
    case B2J_K_WEIGHT_PREP: {
      const b2j_weight_prep_params& p = *reinterpret_cast<const b2j_weight_prep_params*>(op.params.data());
      NEED_BUFS(p.split == 1 ? 3 : 2);
      const uint64_t n = (uint64_t)p.rhs_shape[p.rhs_spec[0]] * p.kpad;
      if (p.cpad && (p.kpad % p.cpad != 0 || p.taps_w == 0 || (p.n_map && p.cpad > B2J_FOLD_CHANNELS)))
        return fail(ctx, B2J_EINVAL, "weight_prep: inconsistent folded layout");
      weight_prep_kernel<<<grid_for(n, 256, ctx, 32), 256, 0, st>>>(p, P<float>(op.bufs[0]), P<const float>(op.bufs[1]),
                                                                    p.split == 1 ? P<float>(op.bufs[2]) : nullptr);
      ++*launches;
    } break;
    case B2J_K_RELAYOUT: {
      const b2j_relayout_params& p = *reinterpret_cast<const b2j_relayout_params*>(op.params.data());
      NEED_BUFS(2);
      if (p.oc % 4 != 0 || (p.n_map && p.oc > B2J_FOLD_CHANNELS)) return fail(ctx, B2J_EINVAL, "relayout: bad destination channel count");
      const uint64_t n = (uint64_t)p.batch * p.oh * p.ow * (p.oc / 4);
      if (n == 0) return B2J_OK;
      int max_dh = 0, max_dw = 0;
      for (uint32_t j = 0; j < p.n_map && j < B2J_FOLD_CHANNELS; ++j) {
        if (p.map[j].dh > max_dh) max_dh = p.map[j].dh;
        if (p.map[j].dw > max_dw) max_dw = p.map[j].dw;
      }
      // destination rows per tile: as many as keep the staged source rows within ~32 KB (amortises the two block barriers)
      const size_t row_bytes = (size_t)(p.fold_w * (RL_TILE - 1) + max_dw + 1) * p.c * sizeof(float);
      int tile_rows = 8;
      while (tile_rows > 1 && (size_t)(p.fold_h * (tile_rows - 1) + max_dh + 1) * row_bytes > 32 * 1024) --tile_rows;
      const size_t smem = (size_t)(p.fold_h * (tile_rows - 1) + max_dh + 1) * row_bytes;
      if (smem > 200 * 1024) return fail(ctx, B2J_ENOTIMPL, "relayout: %zu bytes of staging per tile (too many channels)", smem);
      static size_t configured = 48 * 1024;
      if (smem > configured) {
        CU_CHECK(ctx, cudaFuncSetAttribute(relayout_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
      }
      const uint64_t tiles = (uint64_t)p.batch * ((p.oh + tile_rows - 1) / tile_rows) * ((p.ow + RL_TILE - 1) / RL_TILE);
      relayout_kernel<<<grid_for(tiles * 256, 256, ctx, 16), 256, smem, st>>>(p, P<float>(op.bufs[0]), P<const float>(op.bufs[1]), max_dh, max_dw, tile_rows);
      ++*launches;
    } break;
    case B2J_K_CONV_TC: {
      const b2j_conv_tc_params& p = *reinterpret_cast<const b2j_conv_tc_params*>(op.params.data());
      NEED_BUFS(4);
      EpiPtrs epi;
      int rc = fill_epi(ctx, p.epi, op, &epi);
      if (rc) return rc;
      const char* why = nullptr;
      b2j_conv_tc_params ps = p;
      // outputs of >= 256 MB (twice the L2) are written with evict-first stores: they cannot stay cached until their reader
      // runs and would only evict the residual / weight lines the running kernel still needs (-3 % on the stage-0 residual layers)
      { static int so = -1; if (so < 0) { const char* e = getenv("B2J_STREAM_OUT_MB"); so = e ? atoi(e) : 256; }
        if (so > 0 && (uint64_t)p.batch * p.oh * p.ow * p.o * 4 > (uint64_t)so << 20) ps.flags |= B2J_CT_STREAM_OUT; }
      // stride-1 k x k with N <= 128: patch kernel (one activation fetch serves all filter taps); else the im2col kernel
      rc = launch_conv_patch(ps, epi, P<float>(op.bufs[0]), P<const float>(op.bufs[1]), P<const float>(op.bufs[2]),
                             ctx->prop.multiProcessorCount, st, &why);
      if (rc == B2J_ENOTIMPL)
        rc = launch_conv_tc2(ps, epi, P<float>(op.bufs[0]), P<const float>(op.bufs[1]), P<const float>(op.bufs[2]),
                             P<const float>(op.bufs[3]), ctx->prop.multiProcessorCount, st, &why);
      if (rc) return fail(ctx, rc, "conv_tc: %s", why ? why : "launch failed");
      ++*launches;
    } break;
    case B2J_K_GEMM_TC: {
      const b2j_gemm_tc_params& p = *reinterpret_cast<const b2j_gemm_tc_params*>(op.params.data());
      NEED_BUFS(4);
      EpiPtrs epi;
      int rc = fill_epi(ctx, p.epi, op, &epi);
      if (rc) return rc;
      b2j_conv_tc_params c{};   // a GEMM is a 1x1 convolution over M "pixels"
      c.batch = 1; c.h = 1; c.w = p.m; c.c = p.k; c.kh = c.kw = 1; c.o = p.n; c.oh = 1; c.ow = p.m;
      c.stride_h = c.stride_w = c.dil_h = c.dil_w = 1; c.kpad = p.kpad; c.precision = p.precision; c.flags = p.flags; c.epi = p.epi;
      const char* why = nullptr;
      rc = launch_conv_tc2(c, epi, P<float>(op.bufs[0]), P<const float>(op.bufs[1]), P<const float>(op.bufs[2]),
                           P<const float>(op.bufs[3]), ctx->prop.multiProcessorCount, st, &why);
      if (rc) return fail(ctx, rc, "gemm_tc: %s", why ? why : "launch failed");
      ++*launches;
    } break;
    case B2J_K_GATHER: {
      const b2j_gather_params& p = *reinterpret_cast<const b2j_gather_params*>(op.params.data());
      NEED_BUFS(3);
      if (p.n == 0) return B2J_OK;
      gather_kernel<<<grid_for(p.n, 256, ctx, 32), 256, 0, st>>>(p, P<uint32_t>(op.bufs[0]), P<const uint32_t>(op.bufs[1]),
                                                                 P<const int32_t>(op.bufs[2]));
      ++*launches;
    } break;
    case B2J_K_SCATTER_ADD: {
      const b2j_scatter_params& p = *reinterpret_cast<const b2j_scatter_params*>(op.params.data());
      NEED_BUFS(4);
      if (op.bufs[0] != op.bufs[1])
        CU_CHECK(ctx, cudaMemcpyAsync(P<void>(op.bufs[0]), P<const void>(op.bufs[1]), p.n_operand * 4, cudaMemcpyDeviceToDevice, st));
      if (p.n_updates) {
        scatter_add_kernel<<<grid_for(p.n_updates, 256, ctx, 32), 256, 0, st>>>(p, P<uint32_t>(op.bufs[0]), P<const int32_t>(op.bufs[2]),
                                                                                P<const uint32_t>(op.bufs[3]));
        ++*launches;
      }
    } break;
    case B2J_K_CONCAT: {
      const b2j_concat_params& p = *reinterpret_cast<const b2j_concat_params*>(op.params.data());
      NEED_BUFS(3);
      const uint64_t n = p.outer * (p.ca + p.cb) * p.inner;
      if (n == 0) return B2J_OK;
      concat_kernel<<<grid_for(n, 256, ctx, 32), 256, 0, st>>>(p, P<uint32_t>(op.bufs[0]), P<const uint32_t>(op.bufs[1]),
                                                               P<const uint32_t>(op.bufs[2]));
      ++*launches;
    } break;
    case B2J_K_THREEFRY: {
      const b2j_threefry_params& p = *reinterpret_cast<const b2j_threefry_params*>(op.params.data());
      NEED_BUFS(6);
      if (p.n == 0) return B2J_OK;
      threefry_kernel<<<grid_for(p.n, 256, ctx, 32), 256, 0, st>>>(p, P<uint32_t>(op.bufs[0]), P<uint32_t>(op.bufs[1]),
                                                                   P<const uint32_t>(op.bufs[2]), P<const uint32_t>(op.bufs[3]),
                                                                   P<const uint32_t>(op.bufs[4]), P<const uint32_t>(op.bufs[5]));
      ++*launches;
    } break;
    case B2J_K_DILATE: {
      const b2j_dilate_params& p = *reinterpret_cast<const b2j_dilate_params*>(op.params.data());
      NEED_BUFS(2);
      const uint64_t n = (uint64_t)p.batch * p.oh * p.ow * p.c;
      if (n == 0) return B2J_OK;
      if (p.dil_h == 0 || p.dil_w == 0) return fail(ctx, B2J_EINVAL, "dilate: zero dilation");
      dilate_kernel<<<grid_for(n, 256, ctx, 32), 256, 0, st>>>(p, P<uint32_t>(op.bufs[0]), P<const uint32_t>(op.bufs[1]));
      ++*launches;
    } break;
    case B2J_K_SELECT_SCATTER_ADD: {
      const b2j_reduce_window_params& p = *reinterpret_cast<const b2j_reduce_window_params*>(op.params.data());
      NEED_BUFS(3);
      if (p.dtype != B2J_F32) return fail(ctx, B2J_ENOTIMPL, "select_and_scatter_add dtype %u", p.dtype);
      const uint64_t n = (uint64_t)p.in_shape[0] * p.in_shape[1] * p.in_shape[2] * p.in_shape[3];
      if (n == 0) return B2J_OK;
      for (int d = 0; d < 4; ++d) if (p.strides[d] == 0 || p.window[d] == 0) return fail(ctx, B2J_EINVAL, "select_and_scatter_add: empty window / zero stride");
      if (p.kind == B2J_RW_MAX)
        select_and_scatter_add_kernel<true><<<grid_for(n, 256, ctx, 32), 256, 0, st>>>(p, P<float>(op.bufs[0]), P<const float>(op.bufs[1]), P<const float>(op.bufs[2]));
      else if (p.kind == B2J_RW_MIN)
        select_and_scatter_add_kernel<false><<<grid_for(n, 256, ctx, 32), 256, 0, st>>>(p, P<float>(op.bufs[0]), P<const float>(op.bufs[1]), P<const float>(op.bufs[2]));
      else return fail(ctx, B2J_ENOTIMPL, "select_and_scatter_add: select must be ge or le");
      ++*launches;
    } break;
    case 0xA11u: {  // all-gather pseudo-op
      int rc = allgather_on(ctx, op.bufs[0], op.bufs[1], op.bytes, st);
      if (rc) return rc;
    } break;
    default:
      return fail(ctx, B2J_ENOTIMPL, "no sm_100a kernel for kernel_id %u", op.kid);
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(ctx, B2J_ECUDA, "launch of kernel_id %u failed: %s", op.kid, cudaGetErrorString(e));
  return B2J_OK;
}

// ---- sequences ---------------------------------------------------------------------------------------
extern "C" {

int b2j_seq_create(b2j_ctx* ctx, int profiling, b2j_seq** out) {
  b2j_seq* s = new b2j_seq();
  s->ctx = ctx;
  s->profiling = profiling != 0;
  *out = s;
  return B2J_OK;
}

int b2j_seq_destroy(b2j_seq* seq) {
  if (!seq) return B2J_OK;
  cudaSetDevice(seq->ctx->device);
  cudaStreamSynchronize(seq->ctx->stream);
  if (seq->exec) cudaGraphExecDestroy(seq->exec);
  if (seq->graph) cudaGraphDestroy(seq->graph);
  for (cudaEvent_t e : seq->evs) cudaEventDestroy(e);
  if (seq->ev0) cudaEventDestroy(seq->ev0);
  if (seq->ev1) cudaEventDestroy(seq->ev1);
  delete seq;
  return B2J_OK;
}

int b2j_seq_record(b2j_seq* seq, uint32_t kernel_id, const b2j_buf* bufs, int nbufs, const void* params, size_t params_bytes) {
  b2j_ctx* ctx = seq->ctx;
  if (seq->finalized) return fail(ctx, B2J_EINVAL, "sequence already finalized");
  const size_t want = b2j_param_size(kernel_id);
  if (want == 0) return fail(ctx, B2J_ENOTIMPL, "no sm_100a kernel for kernel_id %u", kernel_id);
  if (want != params_bytes)
    return fail(ctx, B2J_EINVAL, "kernel_id %u: params are %zu bytes, expected %zu (binding out of date?)", kernel_id, params_bytes, want);
  SeqOp op;
  op.kid = kernel_id;
  op.bufs.assign(bufs, bufs + nbufs);
  op.params.assign((const uint8_t*)params, (const uint8_t*)params + params_bytes);
  seq->ops.push_back(std::move(op));
  return B2J_OK;
}

int b2j_seq_record_allgather(b2j_seq* seq, b2j_buf send, b2j_buf recv, size_t bytes_per_rank) {
  if (seq->finalized) return fail(seq->ctx, B2J_EINVAL, "sequence already finalized");
  SeqOp op;
  op.kid = 0xA11u;
  op.bufs = {send, recv};
  op.bytes = bytes_per_rank;
  seq->ops.push_back(std::move(op));
  return B2J_OK;
}

int b2j_seq_finalize(b2j_seq* seq) {
  b2j_ctx* ctx = seq->ctx;
  if (seq->finalized) return B2J_OK;
  CU_CHECK(ctx, cudaSetDevice(ctx->device));
  CU_CHECK(ctx, cudaEventCreate(&seq->ev0));
  CU_CHECK(ctx, cudaEventCreate(&seq->ev1));
  if (seq->profiling) {
    seq->evs.resize(seq->ops.size() + 1);
    for (auto& e : seq->evs) CU_CHECK(ctx, cudaEventCreate(&e));
    // validate once, eagerly
    int launches = 0;
    for (const SeqOp& op : seq->ops) {
      int rc = launch_op(ctx, op, ctx->stream, &launches);
      if (rc) return rc;
    }
    CU_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    seq->n_launches = launches;
    seq->finalized = true;
    return B2J_OK;
  }
  if (seq->ops.empty()) {
    seq->finalized = true;
    return B2J_OK;
  }
  // Trailing collectives (the output all-gather) can be kept out of the graph and issued right behind it on the
  // same stream (B2J_EAGER_COLLECTIVES=1): NCCL kernels captured into a graph were measured at ~2.7 ms per replay
  // for a 1 MB all-gather on this pool (profiles/README.md), against tens of microseconds when launched eagerly.
  {
    const char* e = getenv("B2J_EAGER_COLLECTIVES");
    seq->eager_tail = 0;
    if (!e || e[0] != '0')
      while (seq->eager_tail < seq->ops.size() && seq->ops[seq->ops.size() - 1 - seq->eager_tail].kid == 0xA11u) ++seq->eager_tail;
  }
  CU_CHECK(ctx, cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal));
  int launches = 0, rc = B2J_OK;
  for (size_t i = 0; i + seq->eager_tail < seq->ops.size(); ++i) {
    rc = launch_op(ctx, seq->ops[i], ctx->stream, &launches);
    if (rc) break;
  }
  cudaGraph_t graph = nullptr;
  cudaError_t e = cudaStreamEndCapture(ctx->stream, &graph);
  if (rc) {
    if (graph) cudaGraphDestroy(graph);
    return rc;
  }
  if (e != cudaSuccess) return fail(ctx, B2J_ECUDA, "graph capture failed: %s", cudaGetErrorString(e));
  seq->graph = graph;
  CU_CHECK(ctx, cudaGraphInstantiate(&seq->exec, graph, 0));
  seq->n_launches = launches;
  seq->finalized = true;
  return B2J_OK;
}

int b2j_seq_launch(b2j_seq* seq) {
  b2j_ctx* ctx = seq->ctx;
  if (!seq->finalized) {
    int rc = b2j_seq_finalize(seq);
    if (rc) return rc;
  }
  CU_CHECK(ctx, cudaEventRecord(seq->ev0, ctx->stream));
  if (seq->profiling) {
    int launches = 0;
    CU_CHECK(ctx, cudaEventRecord(seq->evs[0], ctx->stream));
    for (size_t i = 0; i < seq->ops.size(); ++i) {
      int rc = launch_op(ctx, seq->ops[i], ctx->stream, &launches);
      if (rc) return rc;
      CU_CHECK(ctx, cudaEventRecord(seq->evs[i + 1], ctx->stream));
    }
  } else {
    if (seq->exec) CU_CHECK(ctx, cudaGraphLaunch(seq->exec, ctx->stream));
    int launches = 0;
    for (size_t i = seq->ops.size() - seq->eager_tail; i < seq->ops.size(); ++i) {
      int rc = launch_op(ctx, seq->ops[i], ctx->stream, &launches);
      if (rc) return rc;
    }
  }
  CU_CHECK(ctx, cudaEventRecord(seq->ev1, ctx->stream));
  return B2J_OK;
}

int b2j_seq_eval(b2j_seq* seq) {
  int rc = b2j_seq_launch(seq);
  if (rc) return rc;
  CU_CHECK(seq->ctx, cudaStreamSynchronize(seq->ctx->stream));
  return B2J_OK;
}

int b2j_seq_num_ops(b2j_seq* seq, int* n) {
  *n = (int)seq->ops.size();
  return B2J_OK;
}

int b2j_seq_num_launches(b2j_seq* seq, int* n) {
  *n = seq->n_launches;
  return B2J_OK;
}

int b2j_seq_timestamps(b2j_seq* seq, float* ms, int n) {
  b2j_ctx* ctx = seq->ctx;
  if (!seq->profiling) return fail(ctx, B2J_EINVAL, "sequence was not created with profiling=1");
  if (n != (int)seq->ops.size()) return fail(ctx, B2J_EINVAL, "expected room for %zu timestamps", seq->ops.size());
  CU_CHECK(ctx, cudaEventSynchronize(seq->evs.back()));
  for (int i = 0; i < n; ++i) CU_CHECK(ctx, cudaEventElapsedTime(&ms[i], seq->evs[i], seq->evs[i + 1]));
  return B2J_OK;
}

int b2j_seq_last_elapsed_ms(b2j_seq* seq, float* ms) {
  CU_CHECK(seq->ctx, cudaEventSynchronize(seq->ev1));
  CU_CHECK(seq->ctx, cudaEventElapsedTime(ms, seq->ev0, seq->ev1));
  return B2J_OK;
}

}  // extern "C"
