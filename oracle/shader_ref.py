"""ctypes access to oracle/shader_ref.c (CPU ORACLE -- TEST INFRASTRUCTURE ONLY)."""
import ctypes as C
import os
import subprocess
from concurrent.futures import ThreadPoolExecutor

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, '_build', 'libshader_ref.so')
_lib = None


def build():
    subprocess.check_call(['make', '-C', _HERE], stdout=subprocess.DEVNULL)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB):
            build()
        _lib = C.CDLL(_LIB)
    return _lib


def _run(fn, n, *args):
    """fn(*args, begin, end) over [0, n) split across host threads (ctypes drops the GIL)."""
    threads = max(1, min(os.cpu_count() or 1, n // 4096 or 1))
    bounds = [n * i // threads for i in range(threads + 1)]
    if threads == 1:
        fn(*args, C.c_int64(0), C.c_int64(n))
        return
    with ThreadPoolExecutor(threads) as ex:
        list(ex.map(lambda i: fn(*args, C.c_int64(bounds[i]), C.c_int64(bounds[i + 1])), range(threads)))


def _u32(x):
    return np.ascontiguousarray(np.asarray(x, np.uint32))


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def conv2d(lhs, rhs, out_shape, dn, padding_lo, strides, lhs_dil=(1, 1), rhs_dil=(1, 1), fma=True):
    lhs = np.ascontiguousarray(lhs, np.float32)
    rhs = np.ascontiguousarray(rhs, np.float32)
    out = np.empty(out_shape, np.float32)
    args = [_u32(lhs.shape), _u32(rhs.shape), _u32(out_shape), _u32(dn.lhs_spec), _u32(dn.rhs_spec), _u32(dn.out_spec),
            np.ascontiguousarray(np.asarray(padding_lo, np.int32)), _u32(strides), _u32(lhs_dil), _u32(rhs_dil)]
    _run(lib().conv2d_ref, out.size, _p(out), _p(lhs), _p(rhs), *[_p(a) for a in args], C.c_int(int(fma)))
    return out


def dot_general(a, b, cdim_a, cdim_b, fma=True):
    a = np.ascontiguousarray(a, np.float32)
    b = np.ascontiguousarray(b, np.float32)
    n = a.shape[1 - cdim_a]
    m = b.shape[1 - cdim_b]
    c = a.shape[cdim_a]
    out = np.empty((n, m), np.float32)
    _run(lib().dot_general_ref, out.size, _p(out), _p(a), _p(b), C.c_uint32(n), C.c_uint32(c), C.c_uint32(m),
         C.c_uint32(cdim_a), C.c_uint32(cdim_b), C.c_int(int(fma)))
    return out


def reduce_window_max(x, out_shape, padding_lo, strides, window, q3=False):
    x = np.ascontiguousarray(x, np.float32)
    out = np.empty(out_shape, np.float32)
    keep = [_u32(x.shape), _u32(out_shape), _u32(padding_lo), _u32(strides), _u32(window)]
    _run(lib().reduce_window_max_ref, out.size, _p(out), _p(x), *[_p(k) for k in keep], C.c_int(int(q3)))
    return out


def reduce_sum(x, outer, red, inner):
    x = np.ascontiguousarray(x, np.float32)
    out = np.empty((outer, inner), np.float32)
    _run(lib().reduce_sum_ref, out.size, _p(out), _p(x), C.c_uint64(outer), C.c_uint64(red), C.c_uint64(inner))
    return out
