/* b2jax.h -- C ABI of libb2jax.so, the B200 (sm_100a) runtime behind the vkJAX API.
 *
 * This is the drop-in boundary.  The reference reaches its device through the
 * pybind11 module `kp` (Vulkan Kompute 0.7.0, un-vendored) and `pyshaderc`; every entry point
 * below names the reference call site it replaces.  All functions return 0 on success or a
 * B2J_E* status; `b2j_last_error()` gives the message.  Plain pointers and sizes only: no
 * torch / numpy / C++ types cross this boundary.  The binding a maintainer would add on the
 * reference side is shown in INTEGRATION.md (ctypes, mirrored by vkjax_b200/runtime.py).
 *
 * Threading: a context is not thread-safe; one CUDA stream per context; a sequence belongs to
 * its context.  Host pointers are borrowed for the duration of a call only, except pointers
 * obtained from b2j_host_alloc (pinned; owned by the library until b2j_host_free).
 */
#ifndef B2JAX_H
#define B2JAX_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B2J_ABI_VERSION 1

/* status codes → Python exceptions (vkjax_b200/runtime.py): RuntimeError, NotImplementedError, MemoryError */
enum {
  B2J_OK = 0,
  B2J_ECUDA = 1,     /* a CUDA / driver / NCCL call failed                       → RuntimeError        */
  B2J_ENOTIMPL = 2,  /* no sm_100a kernel for this kernel_id / dtype / shape     → NotImplementedError */
  B2J_ENOMEM = 3,    /* device or pinned allocation failed                       → MemoryError         */
  B2J_EINVAL = 4     /* malformed params (size mismatch, bad rank, ...)          → ValueError          */
};

typedef struct b2j_ctx b2j_ctx; /* ≙ kp.Manager(device)      reference kompute_jaxpr_interpreter.py:20 */
typedef struct b2j_seq b2j_seq; /* ≙ kp.Sequence             reference kompute_jaxpr_interpreter.py:49 */
typedef uint64_t b2j_buf;       /* device address; ≙ kp.Tensor reference buffers.py:202               */

typedef struct {
  char name[128];
  int cc_major, cc_minor;
  int sm_count;
  int max_threads_per_block; /* ≙ max_work_group_invocations, reference kompute_jaxpr_interpreter.py:100-102 */
  int max_block_dim_x;       /* ≙ max_work_group_size[0] */
  size_t shared_mem_per_block_optin;
  size_t total_mem, free_mem;
  int l2_bytes;
} b2j_props;

/* ---- context ------------------------------------------------------------------------------ */
int b2j_abi_version(void);
int b2j_device_count(int* n);
int b2j_ctx_create(int device, b2j_ctx** out);   /* device ≙ env VKJAX_DEVICE, reference :19 */
int b2j_ctx_destroy(b2j_ctx* ctx);
int b2j_device_props(b2j_ctx* ctx, b2j_props* out); /* ≙ mgr.get_device_properties() reference :101 */
const char* b2j_last_error(b2j_ctx* ctx);            /* ctx may be NULL: last error of the calling thread */
int b2j_ctx_sync(b2j_ctx* ctx);

/* ---- memory (≙ mgr.tensor(...) + OpTensorSyncDevice/Local, reference buffers.py:189-204) --- */
int b2j_mem_alloc(b2j_ctx* ctx, size_t bytes, b2j_buf* out); /* stream-ordered pool (cudaMallocAsync) */
int b2j_mem_free(b2j_ctx* ctx, b2j_buf buf);
int b2j_mem_set(b2j_ctx* ctx, b2j_buf buf, int byte, size_t bytes);
int b2j_host_alloc(b2j_ctx* ctx, size_t bytes, void** out);  /* pinned host staging */
int b2j_host_free(b2j_ctx* ctx, void* p);
int b2j_upload(b2j_ctx* ctx, b2j_buf dst, const void* host, size_t bytes);   /* blocking */
int b2j_download(b2j_ctx* ctx, b2j_buf src, void* host, size_t bytes);       /* blocking */
int b2j_upload_async(b2j_ctx* ctx, b2j_buf dst, const void* pinned, size_t bytes);
int b2j_download_async(b2j_ctx* ctx, b2j_buf src, void* pinned, size_t bytes);
int b2j_copy_async(b2j_ctx* ctx, b2j_buf dst, b2j_buf src, size_t bytes);
/* Copy lanes: host->device uploads on a second (copy-engine) stream so that the upload of batch i+1 overlaps the
 * replay of batch i (the reference uploads, evaluates and downloads strictly in sequence, kompute_jaxpr_interpreter.py:72-81).
 *   upload : copy stream waits until the lane's previous contents were consumed, copies pinned -> dst, marks the lane ready
 *   acquire: the context stream waits for the lane to be ready
 *   release: the context stream marks the lane consumed (call after the last kernel/copy that reads dst) */
#define B2J_COPY_LANES 4
int b2j_lane_upload(b2j_ctx* ctx, int lane, b2j_buf dst, const void* pinned, size_t bytes);
int b2j_lane_acquire(b2j_ctx* ctx, int lane);
int b2j_lane_release(b2j_ctx* ctx, int lane);
int b2j_lane_sync(b2j_ctx* ctx, int lane);     /* host waits until the lane's last upload has left the pinned source */
/* Result downloads that do not hold up the next replay (the reference downloads synchronously after eval(),
 * kompute_jaxpr_interpreter.py:79-81): device -> pinned host copy on a separate stream, ordered after everything enqueued on the
 * context stream so far; b2j_lane_download_record records an event (b2j_event_create) behind the copies issued so far.
 * `src` must stay unchanged until that event has completed. */
int b2j_lane_download(b2j_ctx* ctx, void* pinned, b2j_buf src, size_t bytes);
int b2j_lane_download_record(b2j_ctx* ctx, void* ev);

/* ---- recorded sequence (≙ sequence.record(OpAlgoDispatch(mgr.algorithm(tensors, spirv, wg)))
 *      reference kompute_jaxpr_interpreter.py:55-60; replay ≙ sequence.eval() :77) ---------- */
int b2j_seq_create(b2j_ctx* ctx, int profiling, b2j_seq** out); /* profiling ≙ total_timestamps>0, :48-49 */
int b2j_seq_destroy(b2j_seq* seq);
/* `params` is the POD struct for `kernel_id` (below); it replaces the constants the reference
 * bakes into GLSL source (reference ops.py:35-41).  bufs[] order is documented per struct. */
int b2j_seq_record(b2j_seq* seq, uint32_t kernel_id, const b2j_buf* bufs, int nbufs,
                   const void* params, size_t params_bytes);
int b2j_seq_record_allgather(b2j_seq* seq, b2j_buf send, b2j_buf recv, size_t bytes_per_rank);
int b2j_seq_finalize(b2j_seq* seq);            /* capture into a CUDA graph and instantiate it */
int b2j_seq_launch(b2j_seq* seq);              /* enqueue one replay on the context stream      */
int b2j_seq_eval(b2j_seq* seq);                /* launch + wait (≙ sequence.eval())             */
int b2j_seq_num_ops(b2j_seq* seq, int* n);
int b2j_seq_num_launches(b2j_seq* seq, int* n); /* CUDA kernels per replay                      */
/* per-op device time of the last profiled eval, ms (≙ sequence.get_timestamps(), reference :92) */
int b2j_seq_timestamps(b2j_seq* seq, float* ms, int n);
/* device time between the start and end of the last b2j_seq_launch (CUDA events), ms */
int b2j_seq_last_elapsed_ms(b2j_seq* seq, float* ms);

/* ---- timing helpers for bench.py (events on the context stream) ---------------------------- */
int b2j_event_create(b2j_ctx* ctx, void** ev);
int b2j_event_record(b2j_ctx* ctx, void* ev);
int b2j_event_elapsed_ms(b2j_ctx* ctx, void* start, void* stop, float* ms);
int b2j_event_sync(b2j_ctx* ctx, void* ev);      /* host waits for the event */
int b2j_event_destroy(b2j_ctx* ctx, void* ev);
int b2j_flush_l2(b2j_ctx* ctx);                /* write a >L2-sized scratch buffer */

/* ---- multi-GPU: one process per GPU; outputs all-gathered over NVLink ------ */
int b2j_nccl_unique_id(void* out128);                          /* rank 0; 128 bytes */
int b2j_comm_init(b2j_ctx* ctx, int nranks, int rank, const void* id128);
int b2j_comm_destroy(b2j_ctx* ctx);
int b2j_allgather(b2j_ctx* ctx, b2j_buf send, b2j_buf recv, size_t bytes_per_rank);
int b2j_broadcast(b2j_ctx* ctx, b2j_buf buf, size_t bytes, int root);

/* ============================================================================================
 * Kernel ids and parameter structs.  All tensors are dense row-major, 32-bit elements
 * (f32 / i32 / u32; bool is stored as u32 0/1, as in the reference: buffers.py:31-34,70-72).
 * ============================================================================================ */
#define B2J_MAX_RANK 8

enum {
  B2J_K_ELTWISE = 1,       /* fused elementwise chain; replaces every binary/unary .comp + select,
                              convert_element_type, integer_pow, iota (reference ops.py:73-155,300,338,527,563) */
  B2J_K_STRIDED_COPY = 2,  /* broadcast_in_dim / slice / rev / N-D transpose (ops.py:187,534,491,435) */
  B2J_K_TRANSPOSE2D = 3,   /* smem-tiled 2-D transpose (transpose.comp) */
  B2J_K_REDUCE = 4,        /* reduce_sum/max/min/prod, argmax/argmin (ops.py:305-335) */
  B2J_K_REDUCE_WINDOW = 5, /* reduce_window_max (+min/sum), 4-D (ops.py:505-524) */
  B2J_K_CONV_DIRECT = 6,   /* conv_general_dilated, any spec/pad/stride/dilation, fp32 FMA (ops.py:463-487) */
  B2J_K_DOT = 7,           /* dot_general 2-D, contracting dim 0/1 each side, fp32 FMA (ops.py:277-297) */
  B2J_K_CONV_TC = 8,       /* NHWC x OHWI implicit GEMM on tcgen05 (TF32 or 3xTF32), fused epilogue */
  B2J_K_WEIGHT_PREP = 9,   /* rhs (any spec) -> [O][Kpad] K-major, zero padded; optional hi/lo split */
  B2J_K_GATHER = 10,       /* XLA gather (ops.py:372-401) */
  B2J_K_SCATTER_ADD = 11,  /* XLA scatter-add (ops.py:404-433) */
  B2J_K_CONCAT = 12,       /* concatenate of 2 operands along any axis (ops.py:349-369) */
  B2J_K_THREEFRY = 13,     /* threefry2x32 (ops.py:550-560) */
  B2J_K_GEMM_TC = 14,      /* dense [M,K]x[N,K]^T on tcgen05 with TMA-fed operands, fused epilogue */
  B2J_K_RELAYOUT = 15,     /* NHWC activations -> channel-padded / space-to-depth folded NHWC' that TMA can address */
  B2J_K_DILATE = 16,       /* NHWC zero stuffing: lhs_dilation of conv_general_dilated for the tensor-core path (conv2d.comp:32-42) */
  B2J_K_SELECT_SCATTER_ADD = 17, /* select_and_scatter_add, 4-D, select = ge | le (max / min pool gradient; no reference handler) */
  B2J_K_MAX = 18
};

/* dtype tags */
enum { B2J_F32 = 0, B2J_I32 = 1, B2J_U32 = 2, B2J_BOOL = 3 };

/* ---- elementwise chain ----------------------------------------------------------------------
 * acc = init;  for s in steps: acc = op_s(acc, operand_s)   (or op_s(operand_s, acc) if SWAP)
 * bufs = [out, in0, in1, ...].  One thread handles 4 consecutive output elements.            */
enum { /* operand kinds */
  B2J_OPK_FULL = 0,    /* same shape as out: index = i                          */
  B2J_OPK_SCALAR = 1,  /* one element                                           */
  B2J_OPK_MOD = 2,     /* index = i % mod  (trailing-dims broadcast, e.g. per-channel) */
  B2J_OPK_STRIDED = 3, /* index = sum coord_d(i) * stride_d (general broadcast) */
  B2J_OPK_DIV = 4      /* index = i / mod  (leading-dims broadcast, e.g. (B,1) against (B,N)) */
};
#define B2J_SRC_NONE 0xFF
#define B2J_SRC_IMM 0xFE
#define B2J_SRC_IOTA 0xFD
#define B2J_STEP_SWAP 1u

enum { /* chain opcodes; _F float32, _I int32, _U uint32; results of compares are u32 0/1 */
  B2J_OP_NOP = 0,
  B2J_OP_ADD_F, B2J_OP_SUB_F, B2J_OP_MUL_F, B2J_OP_DIV_F, B2J_OP_MAX_F, B2J_OP_MIN_F, B2J_OP_POW_F,
  B2J_OP_REM_F, B2J_OP_NEXTAFTER_F, B2J_OP_ATAN2_F,
  B2J_OP_ADD_I, B2J_OP_SUB_I, B2J_OP_MUL_I, B2J_OP_DIV_I, B2J_OP_DIV_U, B2J_OP_MAX_I, B2J_OP_MAX_U,
  B2J_OP_MIN_I, B2J_OP_MIN_U, B2J_OP_REM_I, B2J_OP_REM_U,
  B2J_OP_AND, B2J_OP_OR, B2J_OP_XOR, B2J_OP_SHL, B2J_OP_SHR_L, B2J_OP_SHR_A,
  B2J_OP_GT_F, B2J_OP_GE_F, B2J_OP_LT_F, B2J_OP_LE_F, B2J_OP_EQ_F, B2J_OP_NE_F,
  B2J_OP_GT_I, B2J_OP_GE_I, B2J_OP_LT_I, B2J_OP_LE_I, B2J_OP_EQ_I, B2J_OP_NE_I,
  B2J_OP_GT_U, B2J_OP_GE_U, B2J_OP_LT_U, B2J_OP_LE_U,
  /* unary (operand ignored) */
  B2J_OP_EXP, B2J_OP_LOG, B2J_OP_NEG_F, B2J_OP_NEG_I, B2J_OP_ABS_F, B2J_OP_ABS_I, B2J_OP_RSQRT, B2J_OP_SQRT,
  B2J_OP_ERF, B2J_OP_ERF_INV, B2J_OP_ERFC, B2J_OP_COS, B2J_OP_SIN, B2J_OP_TAN, B2J_OP_COSH, B2J_OP_SINH,
  B2J_OP_TANH, B2J_OP_ACOS, B2J_OP_ASIN, B2J_OP_ATAN, B2J_OP_ACOSH, B2J_OP_ASINH, B2J_OP_ATANH,
  B2J_OP_CEIL, B2J_OP_FLOOR, B2J_OP_ROUND, B2J_OP_SIGN_F, B2J_OP_SIGN_I, B2J_OP_LOG1P, B2J_OP_EXPM1,
  B2J_OP_LOGISTIC, B2J_OP_NOT_BITS, B2J_OP_NOT_BOOL,
  B2J_OP_IPOW_F, B2J_OP_IPOW_I,            /* exponent in imm (int32, may be negative for _F) */
  B2J_OP_CVT_F2I, B2J_OP_CVT_F2U, B2J_OP_CVT_I2F, B2J_OP_CVT_U2F, B2J_OP_CVT_TOBOOL_F, B2J_OP_CVT_TOBOOL_I,
  /* ternary: acc is the predicate; src = on_true, src2 = on_false */
  B2J_OP_SELECT,
  B2J_OP_COUNT
};

typedef struct {
  uint32_t kind;
  uint32_t mod;
  uint32_t strides[B2J_MAX_RANK];
  uint32_t elem;    /* 0: 32-bit elements; 1: packed uint8 (input-only extension), zero-extended to u32 on load */
} b2j_elt_operand;

typedef struct {
  uint16_t op;
  uint8_t src;    /* input slot 0..B2J_ELT_MAX_IN-1, B2J_SRC_IMM or B2J_SRC_NONE */
  uint8_t flags;  /* B2J_STEP_SWAP */
  uint32_t imm;   /* bit pattern when src == IMM; exponent for IPOW */
  uint8_t src2;   /* SELECT only */
  uint8_t pad[3];
  uint32_t imm2;
} b2j_elt_step;

#define B2J_ELT_MAX_IN 6
#define B2J_ELT_MAX_STEPS 16

typedef struct {
  uint64_t n;                       /* output elements */
  uint32_t rank;
  uint32_t shape[B2J_MAX_RANK];     /* output shape (for STRIDED operands / IOTA) */
  uint32_t n_in;
  b2j_elt_operand in[B2J_ELT_MAX_IN];
  uint32_t init_src;                /* input slot, B2J_SRC_IMM or B2J_SRC_IOTA */
  uint32_t init_imm;                /* bit pattern (IMM) or dimension (IOTA) */
  uint32_t n_steps;
  b2j_elt_step steps[B2J_ELT_MAX_STEPS];
} b2j_elt_params;

/* ---- epilogue fused into the contraction kernels (conv / dot / gemm) ------------------------
 * same step model as the elementwise chain, operands restricted to: immediate, a per-output-
 * channel vector (length = N of the GEMM), or a full tensor of the output's shape (residual).
 * Operand buffers follow the kernel's fixed bufs.                                             */
#define B2J_EPI_MAX_STEPS 8
enum { B2J_EPK_IMM = 0, B2J_EPK_CHANNEL = 1, B2J_EPK_FULL = 2 };
typedef struct {
  uint16_t op;      /* B2J_OP_{ADD,SUB,MUL,DIV,MAX,MIN}_F */
  uint8_t kind;     /* B2J_EPK_* */
  uint8_t flags;    /* B2J_STEP_SWAP */
  uint32_t imm;
  uint32_t buf;     /* index into bufs[] */
} b2j_epi_step;
typedef struct {
  uint32_t n_steps;
  b2j_epi_step steps[B2J_EPI_MAX_STEPS];
} b2j_epilogue;

/* ---- strided copy: out[i] = in[base + sum coord_d(i)*stride_d]; bufs = [out, in] ----------- */
typedef struct {
  uint64_t n;
  uint32_t rank;
  uint32_t shape[B2J_MAX_RANK];
  int64_t base;
  int64_t strides[B2J_MAX_RANK];
} b2j_strided_params;

/* ---- (batched) 2-D transpose: out[b][c][r] = in[b][r][c]; bufs = [out, in] ------------------- */
typedef struct { uint32_t rows, cols, batch; } b2j_transpose_params;   /* batch (0 = 1) independent matrices back to back */

/* ---- reduce: bufs = [out, in].  The input is addressed as
 *      in[ sum_k ocoord_k*keep_stride_k + sum_r rcoord_r*red_stride_r ] ---------------------- */
enum { B2J_RED_SUM = 0, B2J_RED_MAX = 1, B2J_RED_MIN = 2, B2J_RED_PROD = 3, B2J_RED_ARGMAX = 4, B2J_RED_ARGMIN = 5 };
typedef struct {
  uint32_t kind, dtype;
  uint64_t n_out, n_red;
  uint32_t keep_rank; uint32_t keep_shape[B2J_MAX_RANK]; uint64_t keep_strides[B2J_MAX_RANK];
  uint32_t red_rank;  uint32_t red_shape[B2J_MAX_RANK];  uint64_t red_strides[B2J_MAX_RANK];
} b2j_reduce_params;

/* ---- reduce_window, 4-D; bufs = [out, in] -------------------------------------------------- */
enum { B2J_RW_MAX = 0, B2J_RW_MIN = 1, B2J_RW_SUM = 2 };
typedef struct {
  uint32_t kind, dtype;
  uint32_t in_shape[4], out_shape[4], window[4], strides[4];
  int32_t pad_lo[4];
} b2j_reduce_window_params;

/* ---- direct convolution (fp32 FMA, sequential (kh,kw,c) sum like conv2d.comp:58-94);
 *      bufs = [out, lhs, rhs, epilogue operands...] ------------------------------------------ */
typedef struct {
  uint32_t lhs_shape[4], rhs_shape[4], out_shape[4];
  uint32_t lhs_spec[4], rhs_spec[4], out_spec[4];   /* (batch, feature, sp0, sp1) positions */
  int32_t pad_lo[2];
  uint32_t stride[2], lhs_dil[2], rhs_dil[2];
  b2j_epilogue epi;
} b2j_conv_direct_params;

/* ---- dot_general 2-D (fp32 FMA): out[N,M] = sum_c A.B; bufs = [out, a, b, epilogue...] ------ */
typedef struct {
  uint32_t n, m, c;
  uint32_t cdim_a, cdim_b;       /* contracting dim (0 or 1) of each operand */
  b2j_epilogue epi;
} b2j_dot_params;

/* ---- weight prep: rhs (any rhs_spec) -> wt[O][Kpad], k = (kh*KW + kw)*I + i, zero padded;
 *      bufs = [wt_hi, rhs] or [wt_hi, rhs, wt_lo] when split != 0 ----------------------------- */
typedef struct {
  uint8_t dh, dw, c, valid;        /* folded channel j reads source pixel (fold*a + dh - pad, fold*b + dw - pad), channel c */
} b2j_fold_entry;
#define B2J_FOLD_CHANNELS 32
typedef struct {
  uint32_t rhs_shape[4], rhs_spec[4];
  uint32_t kpad;
  uint32_t split;  /* 0: copy; 1: wt_hi = nearest TF32, wt_lo = w - wt_hi (3xTF32); 2: round to nearest TF32 */
  /* re-laid-out activations (B2J_K_RELAYOUT): k = (th*taps_w + tw)*cpad + j with
   *   cpad != 0, n_map == 0 : channel padding only   -> w[kh = th, kw = tw, i = j]          (0 for j >= I)
   *   n_map != 0            : folded                  -> w[kh = tap_h*th + dh_j, kw = tap_w*tw + dw_j, i = c_j]
   * entries outside the filter are 0.  cpad == 0: plain layout above. */
  uint32_t cpad, taps_h, taps_w, tap_h, tap_w, n_map;
  b2j_fold_entry map[B2J_FOLD_CHANNELS];
} b2j_weight_prep_params;

/* ---- activation re-layout for the TMA-fed tensor-core kernels; bufs = [dst, src] ---------------
 *   dst[n, a, b, j] = src[n, fold_h*a + dh_j - pad_h, fold_w*b + dw_j - pad_w, c_j]   (0 outside the source / !valid_j)
 * n_map == 0: channel padding only (dh = dw = 0, c_j = j, valid for j < c).  Used for channel counts the im2col
 * tensor map cannot address (C % 32 != 0, e.g. the 3-channel ResNet stem, which is space-to-depth folded:
 * 7x7 stride 2 over 3 channels becomes 4x2 taps (dilation 1x2) over 24 of 32 folded channels). */
typedef struct {
  uint32_t batch, h, w, c;          /* source NHWC */
  uint32_t oh, ow, oc;              /* destination [batch, oh, ow, oc], oc % 4 == 0 */
  uint32_t fold_h, fold_w;
  int32_t pad_h, pad_w;
  uint32_t n_map;
  uint32_t round_tf32;              /* 1: store values rounded to nearest TF32 (the consumer is a single-pass TF32 contraction) */
  /* fused input chain (uint8 images + convert_element_type + div 255 folded into the stem's operand load):
   * src_u8 = 1: src holds packed uint8 elements, widened to f32; then pre_n <= 2 steps x = x (op) imm with
   * op in B2J_OP_{ADD,SUB,MUL,DIV}_F, each rounded separately as the stand-alone elementwise kernel would. */
  uint32_t src_u8, pre_n, pre_op[2], pre_imm[2];
  b2j_fold_entry map[B2J_FOLD_CHANNELS];
} b2j_relayout_params;

/* ---- lhs dilation: dst[n, dil_h*h, dil_w*w, c] = src[n, h, w, c], zeros elsewhere; bufs = [dst, src] ------------------ */
typedef struct {
  uint32_t batch, h, w, c;          /* source NHWC */
  uint32_t oh, ow;                  /* destination extents: (h-1)*dil_h + 1, (w-1)*dil_w + 1 */
  uint32_t dil_h, dil_w;
} b2j_dilate_params;

/* ---- select_and_scatter_add uses b2j_reduce_window_params: in_shape = operand, out_shape = source, kind = B2J_RW_MAX
 *      (select = ge) or B2J_RW_MIN (select = le); bufs = [out, source, operand], f32 only -------------------------- */

/* ---- tcgen05 implicit-GEMM convolution, NHWC activations x [O][Kpad] weights -> NHWC --------
 *      M = B*OH*OW, N = O, K = KH*KW*C.  bufs = [out, x, wt_hi, wt_lo|0, epilogue operands...] */
enum { B2J_PREC_TF32 = 0, B2J_PREC_TF32X3 = 1 };
/* Store the result rounded to the nearest TF32 value.  Set (single-pass TF32 mode only) when every reader of the output is
 * a tensor-core contraction that would otherwise have its operand TRUNCATED to TF32 by the tensor core: rounding where the
 * value is produced removes the systematic shrink (-2^-11 per layer, ~1 % over ResNet-50's 53 layers) at no cost. */
#define B2J_CT_ROUND_OUT_TF32 1u
/* Write the output with cache-streaming (evict-first) stores: set by the library itself for outputs much larger than L2. */
#define B2J_CT_STREAM_OUT 2u
/* x is the RAW NHWC input (f32, or packed uint8 with src_u8): few-channel k x k convolutions (the 3-channel ResNet stem) are fed
 * from staged input rows instead of a re-laid-out copy -- no B2J_K_RELAYOUT launch, K = KH*KW*C padded to a multiple of 32 only.
 * Admission rule: O <= 64, dil_w == 1, Kpad <= 256, 16-byte input rows, KH staged rows within 32 KB, padding >= 0
 * (vkjax_b200/interpreter.py:rows_mode_ok; anything else goes through B2J_K_RELAYOUT + the im2col tensor map). */
#define B2J_CT_ROWS 4u
/* B2J_CT_ROWS, single-pass TF32: round the assembled operand to nearest TF32 (what B2J_K_RELAYOUT's round_tf32 does) */
#define B2J_CT_ROUND_IN_TF32 8u
typedef struct {
  uint32_t batch, h, w, c;
  uint32_t kh, kw, o, oh, ow;
  int32_t pad_h, pad_w;
  uint32_t stride_h, stride_w, dil_h, dil_w;
  uint32_t kpad;
  uint32_t precision;
  uint32_t flags;     /* B2J_CT_* */
  /* B2J_CT_ROWS only: fused input chain as in b2j_relayout_params (src_u8 = 1: x holds packed uint8 elements, widened to f32,
   * then pre_n <= 2 steps x = x (op) imm); the chain must map 0 to 0 (padding is zero-filled before it). */
  uint32_t src_u8, pre_n, pre_op[2], pre_imm[2];
  b2j_epilogue epi;
} b2j_conv_tc_params;

/* ---- tcgen05 GEMM: out[M,N] = A[M,K] . Wt[N,K]^T ; bufs = [out, a, wt_hi, wt_lo|0, epilogue...] */
typedef struct {
  uint32_t m, n, k, kpad;
  uint32_t precision;
  uint32_t flags;     /* B2J_CT_* */
  b2j_epilogue epi;
} b2j_gemm_tc_params;

/* ---- gather: bufs = [out, operand, indices(i32)] -------------------------------------------
 * out index -> (batch coords, offset coords); see vkjax_b200/ops.py:gather for the mapping.  */
typedef struct {
  uint64_t n;
  uint32_t out_rank; uint32_t out_shape[B2J_MAX_RANK];
  uint32_t operand_rank; uint32_t operand_shape[B2J_MAX_RANK];
  uint32_t idx_vec_len;                         /* last dim of indices */
  uint32_t start_index_map[B2J_MAX_RANK];
  uint32_t slice_sizes[B2J_MAX_RANK];
  int32_t out_dim_to_operand_dim[B2J_MAX_RANK]; /* for offset dims: operand dim; -1 for batch dims */
  int64_t out_dim_batch_stride[B2J_MAX_RANK];   /* for batch dims: stride (in index vectors) into indices */
} b2j_gather_params;

/* ---- scatter-add: out = operand; out[scatter(idx)] += updates; bufs = [out, operand, idx, updates] */
typedef struct {
  uint64_t n_operand, n_updates;
  uint32_t operand_rank; uint32_t operand_shape[B2J_MAX_RANK];
  uint32_t upd_rank; uint32_t upd_shape[B2J_MAX_RANK];
  uint32_t idx_vec_len;
  uint32_t scatter_dims_to_operand_dims[B2J_MAX_RANK];
  int32_t upd_dim_to_operand_dim[B2J_MAX_RANK]; /* window dims: operand dim; -1 for scatter dims */
  int64_t upd_dim_batch_stride[B2J_MAX_RANK];   /* scatter dims: stride (in index vectors) into idx */
  uint32_t dtype;
} b2j_scatter_params;

/* ---- concatenate 2 operands: bufs = [out, a, b]; shapes viewed as [outer, ca|cb, inner] ----- */
typedef struct { uint64_t outer, ca, cb, inner; } b2j_concat_params;

/* ---- threefry2x32: bufs = [out0, out1, key0, key1, data0, data1] --------------------------- */
typedef struct { uint64_t n; uint32_t key_is_scalar; } b2j_threefry_params;

size_t b2j_param_size(uint32_t kernel_id); /* sizeof the struct for kernel_id (binding self-check) */

#ifdef __cplusplus
}
#endif
#endif /* B2JAX_H */
