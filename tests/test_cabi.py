"""The C-ABI library loads and exports every symbol include/b2jax.h declares; the ctypes structs match the
library's sizeof; no compute call is made (runs without a GPU)."""
import ctypes
import os
import re

import pytest

from vkjax_b200 import runtime

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    hdr = open(os.path.join(ROOT, 'include', 'b2jax.h')).read()
    hdr = re.sub(r'/\*.*?\*/', '', hdr, flags=re.S)
    return sorted(set(re.findall(r'\b(b2j_[a-z0-9_]+)\s*\(', hdr)))


def test_library_exports_every_declared_symbol():
    lib = runtime.load_library()
    syms = declared_symbols()
    assert len(syms) >= 35
    for s in syms:
        assert hasattr(lib, s), f'{s} declared in include/b2jax.h but not exported by libb2jax.so'
    assert sorted(runtime.EXPORTS) == syms, set(runtime.EXPORTS) ^ set(syms)


def test_param_struct_sizes_match():
    lib = runtime.load_library()
    for kid, st in runtime.PARAM_STRUCTS.items():
        assert lib.b2j_param_size(kid) == ctypes.sizeof(st), runtime.KERNEL_NAMES[kid]
    assert lib.b2j_param_size(999) == 0


def test_opcode_table_matches_header():
    hdr = open(os.path.join(ROOT, 'include', 'b2jax.h')).read()
    body = hdr[hdr.index('B2J_OP_NOP = 0'):hdr.index('B2J_OP_COUNT')]
    body = re.sub(r'/\*.*?\*/', '', body, flags=re.S)
    names = re.findall(r'B2J_OP_([A-Z0-9_]+)', body)
    assert names == runtime._OP_NAMES[:len(names)]
    assert runtime.OP['COUNT'] == len(names)


def test_no_device_means_loud_failure():
    """No CPU fallback: without a usable GPU, creating a context raises."""
    lib = runtime.load_library()
    n = ctypes.c_int()
    if lib.b2j_device_count(ctypes.byref(n)) == 0 and n.value > 0:
        pytest.skip('a GPU is present')
    with pytest.raises((RuntimeError, NotImplementedError, ValueError)):
        runtime.Context(0)
