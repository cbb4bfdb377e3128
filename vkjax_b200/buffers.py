"""Device buffers and the liveness-based buffer pool.

≙ reference vkjax/buffers.py (Buffer :13-51, BufferPool :90-204).  Same planning rule -- a buffer
may reuse an already created tensor that is at least as large and whose access interval is
disjoint (buffers.py:144-187), largest buffers first -- so the reference's white-box tests
(tests/test_bufferpool.py) hold verbatim.  What changes is what a "tensor" is: not a Kompute
tensor with its own host-mapped staging copy, but a 256-byte aligned slot in ONE stream-ordered
CUDA arena (constants and I/O get their own allocations so that they can be bound/replaced
independently).  Everything is 32 bits per element, bool included, as in the reference
(buffers.py:31-34,70-72).
"""
import typing as tp

import numpy as np

from . import core

ALIGN = 256
INF = float('inf')


def canonicalize_host(x: np.ndarray, dtype=None) -> np.ndarray:
    """Host → device dtype coercion (≙ view_as_float32, reference buffers.py:68-77): everything becomes
    one of float32 / int32 / uint32 words; bool is widened to uint32 0/1."""
    x = np.asarray(x)
    if dtype is not None:
        x = x.astype(dtype, copy=False)
    if x.dtype == np.bool_:
        return x.astype(np.uint32)
    if x.dtype == np.float64:
        return x.astype(np.float32)
    if x.dtype == np.int64:
        return x.astype(np.int32)
    if x.dtype == np.uint64:
        return x.astype(np.uint32)
    if x.dtype == np.uint8:
        return x                        # packed bytes (input-only extension: only convert_element_type reads them)
    if x.dtype.type not in (np.float32, np.int32, np.uint32):
        raise NotImplementedError(f'{x.dtype} data types currently not supported')
    return x


def from_device_words(words: np.ndarray, dtype, shape) -> np.ndarray:
    """≙ view_or_convert_from_32bit + reshape (reference buffers.py:23-29,54-59)."""
    dtype = np.dtype(dtype)
    n = int(np.prod(shape, dtype=np.int64))
    if dtype == np.uint8:
        return words.view(np.uint8)[:n].reshape(shape).copy()
    words = words[:n]
    if dtype == np.bool_:
        return (words.view(np.uint32) > 0).reshape(shape)
    return words.view(dtype).reshape(shape).copy()


class Tensor:
    """A device allocation (own) or a slot of the arena."""
    __slots__ = ('nbytes', 'offset', 'addr', 'own', 'initial_value', 'name')

    def __init__(self, nbytes, own=False, initial_value=None, name=''):
        self.nbytes = int(nbytes)
        self.offset = None
        self.addr = None
        self.own = own
        self.initial_value = initial_value
        self.name = name

    def __repr__(self):
        return f'Tensor({self.name}, {self.nbytes}B, own={self.own}, addr={self.addr})'


class TensorPlaceholder:
    def __init__(self, t=None):
        self.t = t


class Buffer:
    """dtype/shape view over a (shared) tensor placeholder + the list of op indices that touch it."""
    def __init__(self, tensor, dtype, shape):
        self._ph = tensor if isinstance(tensor, TensorPlaceholder) else TensorPlaceholder(tensor)
        self.dtype = np.dtype(dtype)
        self.shape = tuple(int(s) for s in shape)
        self.accesses: tp.List[float] = []

    @property
    def tensor(self):
        return self._ph.t

    @tensor.setter
    def tensor(self, value):
        self._ph.t = value

    @property
    def size(self):
        return int(np.prod(self.shape, dtype=np.int64))

    def nbytes(self):
        if self.dtype == np.uint8:
            return (self.size + 3) // 4 * 4          # packed bytes, padded to whole words
        if self.dtype.type not in (np.bool_, np.float32, np.int32, np.uint32):
            raise NotImplementedError(self.dtype)
        return self.size * 4

    def view(self, new_dtype, new_shape):
        new_b = Buffer(self._ph, new_dtype, new_shape)
        new_b.accesses = self.accesses          # views share liveness (reference buffers.py:48-51)
        return new_b

    def same_storage(self, other: 'Buffer'):
        return self._ph is other._ph

    @property
    def addr(self):
        return self.tensor.addr

    def is_constant(self):
        return len(self.accesses) > 0 and max(self.accesses) == INF

    def __repr__(self):
        return f'Buffer({self.dtype.name}{list(self.shape)})'


class BufferPool:
    def __init__(self, ctx=None, workgroup_size: int = 1, reuse_tensors: bool = True):
        self.ctx = ctx                        # runtime.Context or None (planning only, e.g. CPU tests)
        self.workgroup_size = workgroup_size  # kept for API parity; kernels guard their tails (quirk Q11)
        self.buffers: tp.Dict[tp.Any, tp.Optional[Buffer]] = dict()
        self.op_counter = 0
        self.reuse_tensors = reuse_tensors
        self._temp_count = 0
        self.arena_addr = None
        self.arena_bytes = 0
        self.tensors: tp.List[Tensor] = []

    # -- reference API ----------------------------------------------------------------------
    def get_buffer(self, var, increment_op_counter: bool = False) -> tp.Optional[Buffer]:
        varhash = core.hashable(var)
        if varhash not in self.buffers:
            if core.is_unit(var):
                return None
            self.buffers[varhash] = Buffer(None, var.aval.dtype, var.aval.shape)
        b = self.buffers[varhash]
        if b is None:
            return None
        b.accesses.append(self.op_counter)
        if core.is_literal(var) and b.tensor is None:
            self.mark_buffer_as_constant(b, var, var.val)
        if increment_op_counter:
            self.op_counter += 1
        return b

    def mark_buffer_as_constant(self, b: Buffer, var, value):
        if b.tensor is None:
            init = None
            if value is not None:
                init = canonicalize_host(np.asarray(value).astype(np.dtype(var.aval.dtype)))
                assert init.size == b.size, (init.shape, b.shape)
            b.tensor = Tensor(b.nbytes(), own=True, initial_value=init, name=str(var))
        b.accesses += [0, INF]

    def set_buffer(self, var, b: tp.Optional[Buffer]):
        if b is None:
            self.buffers[core.hashable(var)] = None
            return
        self.buffers[core.hashable(var)] = b
        b.accesses.append(self.op_counter)

    # -- additions ----------------------------------------------------------------------------
    def new_temp(self, shape, dtype, name='tmp') -> Buffer:
        """Workspace that is not a jaxpr variable (≙ the `_broadcast` vars of reference ops.py:163)."""
        key = ('__temp__', self._temp_count, name)
        self._temp_count += 1
        b = Buffer(None, dtype, shape)
        b.accesses.append(self.op_counter)
        self.buffers[key] = b
        return b

    def recompute_accesses(self, ops, live_out=()):
        """After the fusion pass the op list differs from what the handlers saw: rebuild liveness from it.
        `live_out` are the buffers of the jaxpr's outvars: they are read after the last op (the download), so they get
        an access at len(ops) -- the reference keeps exactly that access through pool.get_buffer(outvar) at the end of
        analyze_closed_jaxpr (kompute_jaxpr_interpreter.py:45).  Without it an output that is a call result
        (nn.relu -> custom_jvp_call_jaxpr) or a reshape view of an intermediate would hand its arena slot to a later op."""
        seen = set()
        for b in self.buffers.values():
            if b is None or id(b.accesses) in seen:
                continue
            seen.add(id(b.accesses))
            const = b.is_constant()
            del b.accesses[:]
            if const:
                b.accesses += [0, INF]
        for i, op in enumerate(ops):
            for b in op.all_buffers():
                b.accesses.append(i)
        for b in live_out:
            if b is not None and not b.is_constant():
                b.accesses.append(len(ops))

    def create_tensors(self):
        tensor_access_map: tp.Dict[int, tp.List[float]] = dict()
        keys = list(self.buffers.keys())
        buffers = [self.buffers[k] for k in keys]
        order = sorted((i for i, b in enumerate(buffers) if b is not None), key=lambda i: buffers[i].nbytes())
        for i in reversed(order):                       # largest first (easiest to be re-used)
            b = buffers[i]
            if b.tensor is not None:
                continue                                # constant / already planned through a view
            if not b.accesses:
                continue                                # fused away: never materialised
            t = self._search_for_free_tensor(b, tensor_access_map) if self.reuse_tensors else None
            if t is None:
                t = Tensor(b.nbytes(), own=False, name=str(keys[i]))
                tensor_access_map[id(t)] = list(b.accesses)
                self.tensors.append(t)
            b.tensor = t
        self._layout_and_allocate()

    def _search_for_free_tensor(self, b: Buffer, tensor_access_map):
        for other_v, other_b in self.buffers.items():
            if other_b is None:
                continue
            t = other_b.tensor
            if t is None or t.own or isinstance(other_v, int) or other_b.is_constant():
                continue
            if t.nbytes < b.nbytes():
                continue
            t_accesses = tensor_access_map[id(t)]
            if max(b.accesses) < min(t_accesses) or min(b.accesses) > max(t_accesses):
                tensor_access_map[id(t)] = t_accesses + list(b.accesses)
                return t
        return None

    def unique_tensors(self):
        return set(id(b.tensor) for b in self.buffers.values() if b is not None and b.tensor is not None)

    def _layout_and_allocate(self):
        off = 0
        for t in self.tensors:
            t.offset = off
            off += (t.nbytes + ALIGN - 1) // ALIGN * ALIGN
        self.arena_bytes = off
        if self.ctx is None:
            return
        if off:
            self.arena_addr = self.ctx.alloc(off)
        for t in self.tensors:
            t.addr = self.arena_addr + t.offset
        seen = set()
        for b in self.buffers.values():
            if b is None or b.tensor is None or not b.tensor.own or id(b.tensor) in seen:
                continue
            t = b.tensor
            seen.add(id(t))
            t.addr = self.ctx.alloc(max(t.nbytes, 4))
            if t.initial_value is not None:
                self.ctx.upload(t.addr, np.ascontiguousarray(t.initial_value).reshape(-1))
            else:
                self.ctx.memset(t.addr, 0, max(t.nbytes, 4))

    def release(self):
        if self.ctx is None:
            return
        seen = set()
        for b in self.buffers.values():
            if b is None or b.tensor is None or not b.tensor.own or id(b.tensor) in seen or b.tensor.addr is None:
                continue
            seen.add(id(b.tensor))
            self.ctx.free(b.tensor.addr)
            b.tensor.addr = None
        if self.arena_addr:
            self.ctx.free(self.arena_addr)
            self.arena_addr = None
