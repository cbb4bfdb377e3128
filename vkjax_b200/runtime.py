"""ctypes binding of libb2jax.so (include/b2jax.h) -- the only place Python touches the device.

≙ the role of the `kp` pybind11 module in the reference (kp.Manager / kp.Tensor / kp.Sequence,
call sites listed in SURVEY.md §2.1).  There is deliberately no fallback: if the shared library
or a GPU is missing, creating a `Context` raises.
"""
import ctypes as C
import os
import threading
import weakref

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# B2J_LIB: developer switch, loads another build of the same library (A/B runs of compile-time variants on one GPU box)
LIB_PATH = os.environ.get('B2J_LIB') or os.path.join(_HERE, 'csrc', 'libb2jax.so')

MAX_RANK = 8
ELT_MAX_IN = 6
ELT_MAX_STEPS = 16
EPI_MAX_STEPS = 8

# kernel ids (enum in b2jax.h)
K_ELTWISE, K_STRIDED_COPY, K_TRANSPOSE2D, K_REDUCE, K_REDUCE_WINDOW, K_CONV_DIRECT, K_DOT, K_CONV_TC, \
    K_WEIGHT_PREP, K_GATHER, K_SCATTER_ADD, K_CONCAT, K_THREEFRY, K_GEMM_TC, K_RELAYOUT, K_DILATE, K_SELECT_SCATTER_ADD = range(1, 18)
KERNEL_NAMES = {1: 'eltwise', 2: 'strided_copy', 3: 'transpose2d', 4: 'reduce', 5: 'reduce_window', 6: 'conv_direct',
                7: 'dot', 8: 'conv_tc', 9: 'weight_prep', 10: 'gather', 11: 'scatter_add', 12: 'concat',
                13: 'threefry', 14: 'gemm_tc', 15: 'relayout', 16: 'dilate', 17: 'select_and_scatter_add'}

F32, I32, U32, BOOL = 0, 1, 2, 3
DTYPE_TAGS = {np.dtype('float32'): F32, np.dtype('int32'): I32, np.dtype('uint32'): U32, np.dtype('bool'): BOOL}

OPK_FULL, OPK_SCALAR, OPK_MOD, OPK_STRIDED, OPK_DIV = range(5)
SRC_NONE, SRC_IMM, SRC_IOTA = 0xFF, 0xFE, 0xFD
STEP_SWAP = 1
EPK_IMM, EPK_CHANNEL, EPK_FULL = 0, 1, 2
RED_SUM, RED_MAX, RED_MIN, RED_PROD, RED_ARGMAX, RED_ARGMIN = range(6)
RW_MAX, RW_MIN, RW_SUM = range(3)
PREC_TF32, PREC_TF32X3 = 0, 1
CT_ROUND_OUT_TF32 = 1
CT_ROWS, CT_ROUND_IN_TF32 = 4, 8

_OP_NAMES = '''NOP
ADD_F SUB_F MUL_F DIV_F MAX_F MIN_F POW_F REM_F NEXTAFTER_F ATAN2_F
ADD_I SUB_I MUL_I DIV_I DIV_U MAX_I MAX_U MIN_I MIN_U REM_I REM_U
AND OR XOR SHL SHR_L SHR_A
GT_F GE_F LT_F LE_F EQ_F NE_F
GT_I GE_I LT_I LE_I EQ_I NE_I
GT_U GE_U LT_U LE_U
EXP LOG NEG_F NEG_I ABS_F ABS_I RSQRT SQRT
ERF ERF_INV ERFC COS SIN TAN COSH SINH
TANH ACOS ASIN ATAN ACOSH ASINH ATANH
CEIL FLOOR ROUND SIGN_F SIGN_I LOG1P EXPM1
LOGISTIC NOT_BITS NOT_BOOL
IPOW_F IPOW_I
CVT_F2I CVT_F2U CVT_I2F CVT_U2F CVT_TOBOOL_F CVT_TOBOOL_I
SELECT
COUNT'''.split()
OP = {name: i for i, name in enumerate(_OP_NAMES)}     # must match the B2J_OP_* enum order in b2jax.h


class Props(C.Structure):
    _fields_ = [('name', C.c_char * 128), ('cc_major', C.c_int), ('cc_minor', C.c_int), ('sm_count', C.c_int),
                ('max_threads_per_block', C.c_int), ('max_block_dim_x', C.c_int),
                ('shared_mem_per_block_optin', C.c_size_t), ('total_mem', C.c_size_t), ('free_mem', C.c_size_t),
                ('l2_bytes', C.c_int)]


class EltOperand(C.Structure):
    _fields_ = [('kind', C.c_uint32), ('mod', C.c_uint32), ('strides', C.c_uint32 * MAX_RANK), ('elem', C.c_uint32)]


class EltStep(C.Structure):
    _fields_ = [('op', C.c_uint16), ('src', C.c_uint8), ('flags', C.c_uint8), ('imm', C.c_uint32),
                ('src2', C.c_uint8), ('pad', C.c_uint8 * 3), ('imm2', C.c_uint32)]


class EltParams(C.Structure):
    _fields_ = [('n', C.c_uint64), ('rank', C.c_uint32), ('shape', C.c_uint32 * MAX_RANK), ('n_in', C.c_uint32),
                ('in_', EltOperand * ELT_MAX_IN), ('init_src', C.c_uint32), ('init_imm', C.c_uint32),
                ('n_steps', C.c_uint32), ('steps', EltStep * ELT_MAX_STEPS)]


class EpiStep(C.Structure):
    _fields_ = [('op', C.c_uint16), ('kind', C.c_uint8), ('flags', C.c_uint8), ('imm', C.c_uint32), ('buf', C.c_uint32)]


class Epilogue(C.Structure):
    _fields_ = [('n_steps', C.c_uint32), ('steps', EpiStep * EPI_MAX_STEPS)]


class StridedParams(C.Structure):
    _fields_ = [('n', C.c_uint64), ('rank', C.c_uint32), ('shape', C.c_uint32 * MAX_RANK), ('base', C.c_int64),
                ('strides', C.c_int64 * MAX_RANK)]


class TransposeParams(C.Structure):
    _fields_ = [('rows', C.c_uint32), ('cols', C.c_uint32), ('batch', C.c_uint32)]


class ReduceParams(C.Structure):
    _fields_ = [('kind', C.c_uint32), ('dtype', C.c_uint32), ('n_out', C.c_uint64), ('n_red', C.c_uint64),
                ('keep_rank', C.c_uint32), ('keep_shape', C.c_uint32 * MAX_RANK), ('keep_strides', C.c_uint64 * MAX_RANK),
                ('red_rank', C.c_uint32), ('red_shape', C.c_uint32 * MAX_RANK), ('red_strides', C.c_uint64 * MAX_RANK)]


class ReduceWindowParams(C.Structure):
    _fields_ = [('kind', C.c_uint32), ('dtype', C.c_uint32), ('in_shape', C.c_uint32 * 4), ('out_shape', C.c_uint32 * 4),
                ('window', C.c_uint32 * 4), ('strides', C.c_uint32 * 4), ('pad_lo', C.c_int32 * 4)]


class ConvDirectParams(C.Structure):
    _fields_ = [('lhs_shape', C.c_uint32 * 4), ('rhs_shape', C.c_uint32 * 4), ('out_shape', C.c_uint32 * 4),
                ('lhs_spec', C.c_uint32 * 4), ('rhs_spec', C.c_uint32 * 4), ('out_spec', C.c_uint32 * 4),
                ('pad_lo', C.c_int32 * 2), ('stride', C.c_uint32 * 2), ('lhs_dil', C.c_uint32 * 2),
                ('rhs_dil', C.c_uint32 * 2), ('epi', Epilogue)]


class DotParams(C.Structure):
    _fields_ = [('n', C.c_uint32), ('m', C.c_uint32), ('c', C.c_uint32), ('cdim_a', C.c_uint32), ('cdim_b', C.c_uint32),
                ('epi', Epilogue)]


FOLD_CHANNELS = 32


class FoldEntry(C.Structure):
    _fields_ = [('dh', C.c_uint8), ('dw', C.c_uint8), ('c', C.c_uint8), ('valid', C.c_uint8)]


class WeightPrepParams(C.Structure):
    _fields_ = [('rhs_shape', C.c_uint32 * 4), ('rhs_spec', C.c_uint32 * 4), ('kpad', C.c_uint32), ('split', C.c_uint32),
                ('cpad', C.c_uint32), ('taps_h', C.c_uint32), ('taps_w', C.c_uint32), ('tap_h', C.c_uint32),
                ('tap_w', C.c_uint32), ('n_map', C.c_uint32), ('map', FoldEntry * FOLD_CHANNELS)]


class RelayoutParams(C.Structure):
    _fields_ = [('batch', C.c_uint32), ('h', C.c_uint32), ('w', C.c_uint32), ('c', C.c_uint32), ('oh', C.c_uint32),
                ('ow', C.c_uint32), ('oc', C.c_uint32), ('fold_h', C.c_uint32), ('fold_w', C.c_uint32),
                ('pad_h', C.c_int32), ('pad_w', C.c_int32), ('n_map', C.c_uint32), ('round_tf32', C.c_uint32),
                ('src_u8', C.c_uint32), ('pre_n', C.c_uint32), ('pre_op', C.c_uint32 * 2), ('pre_imm', C.c_uint32 * 2),
                ('map', FoldEntry * FOLD_CHANNELS)]


class DilateParams(C.Structure):
    _fields_ = [('batch', C.c_uint32), ('h', C.c_uint32), ('w', C.c_uint32), ('c', C.c_uint32), ('oh', C.c_uint32),
                ('ow', C.c_uint32), ('dil_h', C.c_uint32), ('dil_w', C.c_uint32)]


class ConvTcParams(C.Structure):
    _fields_ = [('batch', C.c_uint32), ('h', C.c_uint32), ('w', C.c_uint32), ('c', C.c_uint32), ('kh', C.c_uint32),
                ('kw', C.c_uint32), ('o', C.c_uint32), ('oh', C.c_uint32), ('ow', C.c_uint32), ('pad_h', C.c_int32),
                ('pad_w', C.c_int32), ('stride_h', C.c_uint32), ('stride_w', C.c_uint32), ('dil_h', C.c_uint32),
                ('dil_w', C.c_uint32), ('kpad', C.c_uint32), ('precision', C.c_uint32), ('flags', C.c_uint32),
                ('src_u8', C.c_uint32), ('pre_n', C.c_uint32), ('pre_op', C.c_uint32 * 2), ('pre_imm', C.c_uint32 * 2), ('epi', Epilogue)]


class GemmTcParams(C.Structure):
    _fields_ = [('m', C.c_uint32), ('n', C.c_uint32), ('k', C.c_uint32), ('kpad', C.c_uint32), ('precision', C.c_uint32),
                ('flags', C.c_uint32), ('epi', Epilogue)]


class GatherParams(C.Structure):
    _fields_ = [('n', C.c_uint64), ('out_rank', C.c_uint32), ('out_shape', C.c_uint32 * MAX_RANK),
                ('operand_rank', C.c_uint32), ('operand_shape', C.c_uint32 * MAX_RANK), ('idx_vec_len', C.c_uint32),
                ('start_index_map', C.c_uint32 * MAX_RANK), ('slice_sizes', C.c_uint32 * MAX_RANK),
                ('out_dim_to_operand_dim', C.c_int32 * MAX_RANK), ('out_dim_batch_stride', C.c_int64 * MAX_RANK)]


class ScatterParams(C.Structure):
    _fields_ = [('n_operand', C.c_uint64), ('n_updates', C.c_uint64), ('operand_rank', C.c_uint32),
                ('operand_shape', C.c_uint32 * MAX_RANK), ('upd_rank', C.c_uint32), ('upd_shape', C.c_uint32 * MAX_RANK),
                ('idx_vec_len', C.c_uint32), ('scatter_dims_to_operand_dims', C.c_uint32 * MAX_RANK),
                ('upd_dim_to_operand_dim', C.c_int32 * MAX_RANK), ('upd_dim_batch_stride', C.c_int64 * MAX_RANK),
                ('dtype', C.c_uint32)]


class ConcatParams(C.Structure):
    _fields_ = [('outer', C.c_uint64), ('ca', C.c_uint64), ('cb', C.c_uint64), ('inner', C.c_uint64)]


class ThreefryParams(C.Structure):
    _fields_ = [('n', C.c_uint64), ('key_is_scalar', C.c_uint32)]


PARAM_STRUCTS = {K_ELTWISE: EltParams, K_STRIDED_COPY: StridedParams, K_TRANSPOSE2D: TransposeParams,
                 K_REDUCE: ReduceParams, K_REDUCE_WINDOW: ReduceWindowParams, K_CONV_DIRECT: ConvDirectParams,
                 K_DOT: DotParams, K_CONV_TC: ConvTcParams, K_WEIGHT_PREP: WeightPrepParams, K_GATHER: GatherParams,
                 K_SCATTER_ADD: ScatterParams, K_CONCAT: ConcatParams, K_THREEFRY: ThreefryParams, K_GEMM_TC: GemmTcParams,
                 K_RELAYOUT: RelayoutParams, K_DILATE: DilateParams, K_SELECT_SCATTER_ADD: ReduceWindowParams}

# every symbol include/b2jax.h declares (tests/test_cabi.py checks the library exports exactly these)
EXPORTS = '''b2j_abi_version b2j_device_count b2j_ctx_create b2j_ctx_destroy b2j_device_props b2j_last_error b2j_ctx_sync
b2j_mem_alloc b2j_mem_free b2j_mem_set b2j_host_alloc b2j_host_free b2j_upload b2j_download b2j_upload_async
b2j_download_async b2j_copy_async b2j_lane_upload b2j_lane_acquire b2j_lane_release b2j_lane_sync b2j_lane_download b2j_lane_download_record b2j_seq_create b2j_seq_destroy b2j_seq_record b2j_seq_record_allgather
b2j_seq_finalize b2j_seq_launch b2j_seq_eval b2j_seq_num_ops b2j_seq_num_launches b2j_seq_timestamps
b2j_seq_last_elapsed_ms b2j_event_create b2j_event_record b2j_event_elapsed_ms b2j_event_sync b2j_event_destroy b2j_flush_l2
b2j_nccl_unique_id b2j_comm_init b2j_comm_destroy b2j_allgather b2j_broadcast b2j_param_size'''.split()

_lib = None
_lib_lock = threading.Lock()


def load_library():
    """dlopen libb2jax.so and check the binding against it.  Raises if it is missing (no fallback)."""
    global _lib
    with _lib_lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f'{LIB_PATH} not found: build it with `python -c "import __graft_entry__ as g; g.build()"` '
                               f'(or `make -C vkjax_b200/csrc`). There is no CPU fallback.')
        # NCCL is dlopen'ed lazily by the library; point it at the torch-bundled copy when there is one
        if 'B2J_NCCL_LIB' not in os.environ:
            try:
                import importlib.util
                spec = importlib.util.find_spec('nvidia.nccl')
                if spec and spec.submodule_search_locations:
                    cand = os.path.join(list(spec.submodule_search_locations)[0], 'lib', 'libnccl.so.2')
                    if os.path.exists(cand):
                        os.environ['B2J_NCCL_LIB'] = cand
            except Exception:
                pass
        lib = C.CDLL(LIB_PATH)
        lib.b2j_last_error.restype = C.c_char_p
        lib.b2j_last_error.argtypes = [C.c_void_p]
        lib.b2j_param_size.restype = C.c_size_t
        lib.b2j_param_size.argtypes = [C.c_uint32]
        if lib.b2j_abi_version() != 1:
            raise RuntimeError('libb2jax.so ABI version mismatch')
        for kid, st in PARAM_STRUCTS.items():
            want = lib.b2j_param_size(kid)
            if want != C.sizeof(st):
                raise RuntimeError(f'binding out of date: {st.__name__} is {C.sizeof(st)} bytes, library says {want}')
        vp, u64, sz = C.c_void_p, C.c_uint64, C.c_size_t
        sigs = {
            'b2j_device_count': [C.POINTER(C.c_int)],
            'b2j_ctx_create': [C.c_int, C.POINTER(vp)], 'b2j_ctx_destroy': [vp], 'b2j_device_props': [vp, C.POINTER(Props)],
            'b2j_ctx_sync': [vp],
            'b2j_mem_alloc': [vp, sz, C.POINTER(u64)], 'b2j_mem_free': [vp, u64], 'b2j_mem_set': [vp, u64, C.c_int, sz],
            'b2j_host_alloc': [vp, sz, C.POINTER(vp)], 'b2j_host_free': [vp, vp],
            'b2j_upload': [vp, u64, vp, sz], 'b2j_download': [vp, u64, vp, sz],
            'b2j_upload_async': [vp, u64, vp, sz], 'b2j_download_async': [vp, u64, vp, sz],
            'b2j_copy_async': [vp, u64, u64, sz],
            'b2j_lane_upload': [vp, C.c_int, u64, vp, sz], 'b2j_lane_acquire': [vp, C.c_int], 'b2j_lane_release': [vp, C.c_int], 'b2j_lane_sync': [vp, C.c_int],
            'b2j_lane_download': [vp, vp, u64, sz], 'b2j_lane_download_record': [vp, vp],
            'b2j_seq_create': [vp, C.c_int, C.POINTER(vp)], 'b2j_seq_destroy': [vp],
            'b2j_seq_record': [vp, C.c_uint32, C.POINTER(u64), C.c_int, vp, sz],
            'b2j_seq_record_allgather': [vp, u64, u64, sz],
            'b2j_seq_finalize': [vp], 'b2j_seq_launch': [vp], 'b2j_seq_eval': [vp],
            'b2j_seq_num_ops': [vp, C.POINTER(C.c_int)], 'b2j_seq_num_launches': [vp, C.POINTER(C.c_int)],
            'b2j_seq_timestamps': [vp, C.POINTER(C.c_float), C.c_int], 'b2j_seq_last_elapsed_ms': [vp, C.POINTER(C.c_float)],
            'b2j_event_create': [vp, C.POINTER(vp)], 'b2j_event_record': [vp, vp],
            'b2j_event_elapsed_ms': [vp, vp, vp, C.POINTER(C.c_float)], 'b2j_event_destroy': [vp, vp], 'b2j_event_sync': [vp, vp], 'b2j_flush_l2': [vp],
            'b2j_nccl_unique_id': [vp], 'b2j_comm_init': [vp, C.c_int, C.c_int, vp], 'b2j_comm_destroy': [vp],
            'b2j_allgather': [vp, u64, u64, sz], 'b2j_broadcast': [vp, u64, sz, C.c_int],
        }
        for name, argtypes in sigs.items():
            fn = getattr(lib, name)
            fn.argtypes = argtypes
            fn.restype = C.c_int
        _lib = lib
        return lib


_ERRORS = {1: RuntimeError, 2: NotImplementedError, 3: MemoryError, 4: ValueError}


def _check(rc, ctx_handle=None):
    if rc != 0:
        msg = load_library().b2j_last_error(ctx_handle)
        raise _ERRORS.get(rc, RuntimeError)((msg or b'unknown error').decode())


class HostBuffer:
    """Pinned host memory exposed as a numpy array (≙ kp.Tensor.data(), the mapped staging view,
    reference buffers.py:24)."""
    def __init__(self, ctx, nbytes):
        self.ctx, self.nbytes = ctx, int(nbytes)
        p = C.c_void_p()
        _check(ctx.lib.b2j_host_alloc(ctx.handle, self.nbytes, C.byref(p)), ctx.handle)
        self.ptr = p.value
        self.array = np.ctypeslib.as_array((C.c_uint8 * max(self.nbytes, 1)).from_address(self.ptr))[:self.nbytes]
        self._fin = weakref.finalize(self, HostBuffer._free, ctx, self.ptr)

    @staticmethod
    def _free(ctx, ptr):
        if ctx.handle:
            ctx.lib.b2j_host_free(ctx.handle, ptr)


class Context:
    """≙ kp.Manager(device).  One per (process, device), shared by all interpreters -- the
    reference creates one Vulkan context per traced signature (kompute_jaxpr_interpreter.py:20)."""
    _instances = {}

    @classmethod
    def get(cls, device=None):
        if device is None:
            device = int(os.environ.get('VKJAX_DEVICE', os.environ.get('LOCAL_RANK', 0)))
        if device not in cls._instances:
            cls._instances[device] = Context(device)
        return cls._instances[device]

    def __init__(self, device=0):
        self.lib = load_library()
        self.device = device
        h = C.c_void_p()
        _check(self.lib.b2j_ctx_create(device, C.byref(h)))
        self.handle = h.value
        self._pinned_ranges = []          # [(start, end, HostBuffer)] for arrays made by pinned_empty()
        self.nranks, self.rank = 1, 0

    def props(self) -> Props:
        p = Props()
        _check(self.lib.b2j_device_props(self.handle, C.byref(p)), self.handle)
        return p

    def sync(self):
        _check(self.lib.b2j_ctx_sync(self.handle), self.handle)

    # memory --------------------------------------------------------------------------------
    def alloc(self, nbytes) -> int:
        out = C.c_uint64()
        _check(self.lib.b2j_mem_alloc(self.handle, int(nbytes), C.byref(out)), self.handle)
        return out.value

    def free(self, buf):
        if self.handle:
            self.lib.b2j_mem_free(self.handle, buf)

    def memset(self, buf, byte, nbytes):
        _check(self.lib.b2j_mem_set(self.handle, buf, byte, int(nbytes)), self.handle)

    def pinned_empty(self, shape, dtype=np.float32) -> np.ndarray:
        """A numpy array in pinned host memory: passing it to a wrapped function lets the H2D copy
        run straight from it (no staging memcpy)."""
        dtype = np.dtype(dtype)
        n = int(np.prod(shape, dtype=np.int64)) * dtype.itemsize
        hb = HostBuffer(self, n)
        arr = hb.array.view(dtype).reshape(shape)
        self._pinned_ranges.append((hb.ptr, hb.ptr + n, hb))
        return arr

    def is_pinned(self, arr: np.ndarray) -> bool:
        if not arr.flags['C_CONTIGUOUS']:
            return False
        a = arr.ctypes.data
        return any(s <= a and a + arr.nbytes <= e for s, e, _ in self._pinned_ranges)

    def upload(self, buf, arr: np.ndarray):
        arr = np.ascontiguousarray(arr)
        _check(self.lib.b2j_upload(self.handle, buf, arr.ctypes.data, arr.nbytes), self.handle)

    def upload_async(self, buf, ptr, nbytes):
        _check(self.lib.b2j_upload_async(self.handle, buf, ptr, int(nbytes)), self.handle)

    def download(self, buf, arr: np.ndarray):
        assert arr.flags['C_CONTIGUOUS']
        _check(self.lib.b2j_download(self.handle, buf, arr.ctypes.data, arr.nbytes), self.handle)

    def download_async(self, buf, ptr, nbytes):
        _check(self.lib.b2j_download_async(self.handle, buf, ptr, int(nbytes)), self.handle)

    def copy_async(self, dst, src, nbytes):
        _check(self.lib.b2j_copy_async(self.handle, dst, src, int(nbytes)), self.handle)

    # copy lanes (uploads overlapped with compute) ------------------------------------------------
    def lane_upload(self, lane, buf, ptr, nbytes):
        _check(self.lib.b2j_lane_upload(self.handle, lane, buf, ptr, int(nbytes)), self.handle)

    def lane_acquire(self, lane):
        _check(self.lib.b2j_lane_acquire(self.handle, lane), self.handle)

    def lane_release(self, lane):
        _check(self.lib.b2j_lane_release(self.handle, lane), self.handle)

    def lane_sync(self, lane):
        _check(self.lib.b2j_lane_sync(self.handle, lane), self.handle)

    def lane_download(self, ptr, buf, nbytes):
        _check(self.lib.b2j_lane_download(self.handle, ptr, buf, int(nbytes)), self.handle)

    def lane_download_record(self, ev):
        _check(self.lib.b2j_lane_download_record(self.handle, ev), self.handle)

    # events ----------------------------------------------------------------------------------
    def event(self):
        e = C.c_void_p()
        _check(self.lib.b2j_event_create(self.handle, C.byref(e)), self.handle)
        return e.value

    def record(self, ev):
        _check(self.lib.b2j_event_record(self.handle, ev), self.handle)

    def elapsed_ms(self, start, stop) -> float:
        ms = C.c_float()
        _check(self.lib.b2j_event_elapsed_ms(self.handle, start, stop, C.byref(ms)), self.handle)
        return ms.value

    def event_sync(self, ev):
        _check(self.lib.b2j_event_sync(self.handle, ev), self.handle)

    def flush_l2(self):
        _check(self.lib.b2j_flush_l2(self.handle), self.handle)

    # multi-GPU ---------------------------------------------------------------------------------
    def nccl_unique_id(self) -> bytes:
        buf = C.create_string_buffer(128)
        _check(self.lib.b2j_nccl_unique_id(buf))
        return buf.raw

    def comm_init(self, nranks, rank, uid: bytes):
        _check(self.lib.b2j_comm_init(self.handle, nranks, rank, C.create_string_buffer(uid, 128)), self.handle)
        self.nranks, self.rank = nranks, rank

    def allgather(self, send, recv, bytes_per_rank):
        _check(self.lib.b2j_allgather(self.handle, send, recv, int(bytes_per_rank)), self.handle)

    def broadcast(self, buf, nbytes, root=0):
        _check(self.lib.b2j_broadcast(self.handle, buf, int(nbytes), root), self.handle)


class Sequence:
    """≙ kp.Sequence: record once, replay per call (reference kompute_jaxpr_interpreter.py:48-63,77)."""
    def __init__(self, ctx: Context, profiling=False):
        self.ctx = ctx
        h = C.c_void_p()
        _check(ctx.lib.b2j_seq_create(ctx.handle, int(bool(profiling)), C.byref(h)), ctx.handle)
        self.handle = h.value
        self._fin = weakref.finalize(self, Sequence._destroy, ctx, self.handle)

    @staticmethod
    def _destroy(ctx, handle):
        if ctx.handle:
            ctx.lib.b2j_seq_destroy(handle)

    def record(self, kernel_id, bufs, params):
        arr = (C.c_uint64 * len(bufs))(*bufs)
        _check(self.ctx.lib.b2j_seq_record(self.handle, kernel_id, arr, len(bufs), C.byref(params), C.sizeof(params)),
               self.ctx.handle)

    def record_allgather(self, send, recv, bytes_per_rank):
        _check(self.ctx.lib.b2j_seq_record_allgather(self.handle, send, recv, int(bytes_per_rank)), self.ctx.handle)

    def finalize(self):
        _check(self.ctx.lib.b2j_seq_finalize(self.handle), self.ctx.handle)

    def launch(self):
        _check(self.ctx.lib.b2j_seq_launch(self.handle), self.ctx.handle)

    def eval(self):
        _check(self.ctx.lib.b2j_seq_eval(self.handle), self.ctx.handle)

    def num_launches(self) -> int:
        n = C.c_int()
        _check(self.ctx.lib.b2j_seq_num_launches(self.handle, C.byref(n)), self.ctx.handle)
        return n.value

    def num_ops(self) -> int:
        n = C.c_int()
        _check(self.ctx.lib.b2j_seq_num_ops(self.handle, C.byref(n)), self.ctx.handle)
        return n.value

    def timestamps(self):
        n = self.num_ops()
        ms = (C.c_float * max(n, 1))()
        _check(self.ctx.lib.b2j_seq_timestamps(self.handle, ms, n), self.ctx.handle)
        return [ms[i] for i in range(n)]

    def last_elapsed_ms(self) -> float:
        ms = C.c_float()
        _check(self.ctx.lib.b2j_seq_last_elapsed_ms(self.handle, C.byref(ms)), self.ctx.handle)
        return ms.value
