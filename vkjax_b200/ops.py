"""Per-primitive dispatch: jaxpr equation -> device ops.

≙ reference vkjax/ops.py.  The dispatch rule is the reference's (ops.py:52-65): the handler is the
module-level callable named like the primitive ('-' -> '_'); equations whose outputs are all
dropped are skipped; an unknown primitive raises NotImplementedError(eq).  Handlers keep the
reference's parameter validation (bare asserts, ValueError for bad broadcasts, NotImplementedError
for unsupported dtypes / dimension numbers -- SURVEY.md §8b "Error conventions") but instead of
baking shapes into GLSL and compiling SPIR-V per op (ops.py:35-41) they emit records for the
ahead-of-time compiled sm_100a kernels of libb2jax.so:

  ChainOp        fused elementwise chain            (B2J_K_ELTWISE)
  ContractionOp  conv_general_dilated / dot_general (direct fp32 | tcgen05 TF32 / 3xTF32) + fused epilogue
  KernelOp       everything else, params struct filled here

Implicit broadcasts of binary operands are never materialised (the reference does, ops.py:84-86,
FIXME at :171): the elementwise kernel reads them through index math.
"""
import builtins
import typing as tp

import numpy as np

from . import core
from . import runtime as rt
from .buffers import Buffer, BufferPool

OP = rt.OP

# contraction precision modes (JaxprInterpreter(precision=...))
PRECISIONS = ('fp32', 'tf32', 'simt')
# below this many FLOPs a contraction is not worth the weight-prep + tensor-core tile machinery
TC_MIN_FLOPS = 1 << 20


def dtype_tag(dt) -> int:
    dt = np.dtype(dt)
    if dt not in rt.DTYPE_TAGS:
        raise NotImplementedError(f'{dt} data types currently not supported')
    return rt.DTYPE_TAGS[dt]


def _imm_bits(val, dtype) -> int:
    dtype = np.dtype(dtype)
    if dtype == np.bool_:
        return int(bool(val))
    return int(np.asarray(val).astype(dtype).reshape(()).view(np.uint32))


# =================================================================================================
# op records
class Operand(tp.NamedTuple):
    kind: str                       # 'buf' | 'imm' | 'iota'
    buf: tp.Optional[Buffer] = None
    imm: int = 0


class Step(tp.NamedTuple):
    op: int
    operand: tp.Optional[Operand] = None     # None for unary ops
    swap: bool = False                       # operand is the left-hand side
    operand2: tp.Optional[Operand] = None    # SELECT: on_false
    imm: int = 0                             # IPOW exponent


class ChainOp:
    """acc = init; for step: acc = op(acc, operand).  One launch of the fused elementwise kernel."""
    kind = 'chain'

    def __init__(self, out: Buffer, init: Operand, steps: tp.List[Step], equation=None):
        self.out, self.init, self.steps = out, init, list(steps)
        self.equations = [equation]

    @property
    def equation(self):
        return self.equations[-1]

    def inputs(self) -> tp.List[Buffer]:
        bufs = []
        for o in [self.init] + [s.operand for s in self.steps] + [s.operand2 for s in self.steps]:
            if o is not None and o.kind == 'buf':
                bufs.append(o.buf)
        return bufs

    def outputs(self):
        return [self.out]

    def all_buffers(self):
        return [self.out] + self.inputs()

    def label(self):
        return '+'.join(str(getattr(getattr(e, 'primitive', None), 'name', e)) for e in self.equations)


class ContractionOp:
    """conv_general_dilated or dot_general with a fused epilogue (list of Step whose operands are
    immediates, per-output-channel vectors or full-shape tensors)."""
    kind = 'contraction'

    def __init__(self, what, out, lhs, rhs, attrs, equation):
        self.what = what              # 'conv' | 'dot'
        self.out, self.lhs, self.rhs = out, lhs, rhs
        self.attrs = attrs
        self.epilogue: tp.List[Step] = []
        self.temps: tp.List[Buffer] = []     # workspace (prepared weights, transposed lhs)
        self.equations = [equation]
        self.path = None

    @property
    def equation(self):
        return self.equations[-1]

    def inputs(self):
        return [self.lhs, self.rhs] + [s.operand.buf for s in self.epilogue if s.operand is not None and s.operand.kind == 'buf']

    def outputs(self):
        return [self.out]

    def all_buffers(self):
        return [self.out] + self.inputs() + self.temps

    def label(self):
        return '+'.join(str(getattr(getattr(e, 'primitive', None), 'name', e)) for e in self.equations)

    def work(self):
        """(M, N, K, algorithmic FLOPs, algorithmic bytes) -- SURVEY.md §8d: bytes = 4 * (lhs + rhs + out), where only the lhs
        elements some filter tap actually reads are counted: a strided 1x1 projection (ResNet's stride-2 shortcuts) touches
        one input pixel in stride_h * stride_w, and charging the whole input overstated its bandwidth."""
        a = self.attrs
        if self.what == 'dot':
            m, n, k = a['n'] * a.get('batch', 1), a['m'], a['c']
            lhs_elems = self.lhs.size
        else:
            rs, sp = a['rhs_shape'], a['rhs_spec']
            n = rs[sp[0]]
            k = rs[sp[1]] * rs[sp[2]] * rs[sp[3]]
            m = int(np.prod(a['out_shape'], dtype=np.int64)) // n
            L, sl, O_, so = a['lhs_shape'], a['lhs_spec'], a['out_shape'], a['out_spec']
            touched = []
            for d in range(2):
                size, out = L[sl[2 + d]], O_[so[2 + d]]
                taps, stride, rdil, ldil, pad = rs[sp[2 + d]], a['stride'][d], a['rhs_dil'][d], a['lhs_dil'][d], a['pad_lo'][d]
                pos = (np.arange(out)[:, None] * stride + np.arange(taps)[None, :] * rdil - pad).reshape(-1)
                pos = pos[(pos >= 0) & (pos % ldil == 0)] // ldil
                touched.append(len(np.unique(pos[pos < size])))
            lhs_elems = L[sl[0]] * L[sl[1]] * touched[0] * touched[1]
        nbytes = 4 * (lhs_elems + self.rhs.size + self.out.size)
        return m, n, k, 2 * m * n * k, nbytes


class KernelOp:
    kind = 'kernel'

    def __init__(self, kernel_id, outs, ins, params, equation):
        self.kernel_id, self.outs, self.ins, self.params = kernel_id, list(outs), list(ins), params
        self.equations = [equation]

    @property
    def equation(self):
        return self.equations[-1]

    def inputs(self):
        return self.ins

    def outputs(self):
        return self.outs

    def all_buffers(self):
        return self.outs + self.ins

    def label(self):
        return str(getattr(getattr(self.equation, 'primitive', None), 'name', self.equation))


Op = tp.Union[ChainOp, ContractionOp, KernelOp]


# =================================================================================================
def analyze_jaxpr(bufferpool: BufferPool, jaxpr) -> tp.List[Op]:
    """Analyzes (possibly inner) jaxprs (≙ reference ops.py:52-65)."""
    all_ops = []
    for eq in jaxpr.eqns:
        if all(core.is_dropvar(v) for v in eq.outvars):
            continue                                    # seems like a redundant operation
        opname = eq.primitive.name.replace('-', '_')
        method = globals().get(opname, None)
        if method is None or not callable(method) or opname.startswith('_') or opname not in PRIMITIVES:
            raise NotImplementedError(eq)
        all_ops += method(bufferpool, eq)
    return all_ops


PRIMITIVES = set()


def primitive(*names):
    def deco(fn):
        for n in names:
            globals()[n] = fn
            PRIMITIVES.add(n)
        return fn
    return deco


# =================================================================================================
# elementwise
def _operand(bufferpool, var) -> Operand:
    """Scalar literals become immediates (no device tensor, unlike reference buffers.py:112-114)."""
    if core.is_literal(var) and np.shape(var.val) == ():
        return Operand('imm', None, _imm_bits(var.val, var.aval.dtype))
    return Operand('buf', bufferpool.get_buffer(var))


def check_broadcastable(in_shape, out_shape):
    """≙ the shape rule of the reference's `broadcast` helper (ops.py:165-169): trailing-aligned, a dim
    must match or be 1; raises ValueError otherwise."""
    if len(in_shape) > len(out_shape):
        raise ValueError(f'Cannot broadcast from {in_shape} to {out_shape}')
    for olddim, newdim in zip(in_shape[::-1], out_shape[::-1]):
        if olddim != newdim and olddim != 1:
            raise ValueError(f'Cannot broadcast from {in_shape} to {out_shape}')


_BINARY_TABLE = {
    # name: (float op, int op, uint op)
    'add': ('ADD_F', 'ADD_I', 'ADD_I'), 'add_any': ('ADD_F', 'ADD_I', 'ADD_I'),
    'sub': ('SUB_F', 'SUB_I', 'SUB_I'), 'mul': ('MUL_F', 'MUL_I', 'MUL_I'),
    'div': ('DIV_F', 'DIV_I', 'DIV_U'), 'max': ('MAX_F', 'MAX_I', 'MAX_U'), 'min': ('MIN_F', 'MIN_I', 'MIN_U'),
    'rem': ('REM_F', 'REM_I', 'REM_U'), 'pow': ('POW_F', None, None), 'nextafter': ('NEXTAFTER_F', None, None),
    'atan2': ('ATAN2_F', None, None),
    'gt': ('GT_F', 'GT_I', 'GT_U'), 'ge': ('GE_F', 'GE_I', 'GE_U'), 'lt': ('LT_F', 'LT_I', 'LT_U'),
    'le': ('LE_F', 'LE_I', 'LE_U'), 'eq': ('EQ_F', 'EQ_I', 'EQ_I'), 'ne': ('NE_F', 'NE_I', 'NE_I'),
    'and': (None, 'AND', 'AND'), 'or': (None, 'OR', 'OR'), 'xor': (None, 'XOR', 'XOR'),
    'shift_left': (None, 'SHL', 'SHL'), 'shift_right_logical': (None, 'SHR_L', 'SHR_L'),
    'shift_right_arithmetic': (None, 'SHR_A', 'SHR_A'),
}


def _typed(table_row, dtype, eq):
    kind = np.dtype(dtype).kind
    name = table_row[{'f': 0, 'i': 1, 'u': 2, 'b': 2}[kind]]
    if name is None:
        raise NotImplementedError(eq)
    return OP[name]


@primitive(*_BINARY_TABLE.keys())
def element_wise_binary_op(bufferpool, equation):
    """≙ reference ops.py:73-114."""
    assert equation.params == {}
    assert len(equation.invars) == 2
    assert len(equation.outvars) == 1
    outvar = equation.outvars[0]
    for invar in equation.invars:
        dtype_tag(invar.aval.dtype)
        check_broadcastable(tuple(invar.aval.shape), tuple(outvar.aval.shape))
    name = equation.primitive.name.replace('-', '_')
    opcode = _typed(_BINARY_TABLE[name], equation.invars[0].aval.dtype, equation)
    a = _operand(bufferpool, equation.invars[0])
    b = _operand(bufferpool, equation.invars[1])
    outbuf = bufferpool.get_buffer(outvar, increment_op_counter=True)
    return [ChainOp(outbuf, a, [Step(opcode, b)], equation)]


_UNARY_TABLE = {
    'exp': 'EXP', 'log': 'LOG', 'rsqrt': 'RSQRT', 'sqrt': 'SQRT', 'erf': 'ERF', 'erf_inv': 'ERF_INV', 'erfc': 'ERFC',
    'cos': 'COS', 'sin': 'SIN', 'tan': 'TAN', 'cosh': 'COSH', 'sinh': 'SINH', 'tanh': 'TANH', 'acos': 'ACOS',
    'asin': 'ASIN', 'atan': 'ATAN', 'acosh': 'ACOSH', 'asinh': 'ASINH', 'atanh': 'ATANH', 'ceil': 'CEIL',
    'floor': 'FLOOR', 'round': 'ROUND', 'log1p': 'LOG1P', 'expm1': 'EXPM1', 'logistic': 'LOGISTIC',
}


@primitive(*_UNARY_TABLE.keys())
def element_wise_unary_op(bufferpool, equation):
    """≙ reference ops.py:119-155 (float32 only, same shape)."""
    assert len(equation.invars) == 1
    assert len(equation.outvars) == 1
    outvar, invar = equation.outvars[0], equation.invars[0]
    assert outvar.aval.shape == invar.aval.shape
    assert outvar.aval.dtype == invar.aval.dtype == np.float32
    a = _operand(bufferpool, invar)
    outbuf = bufferpool.get_buffer(outvar, increment_op_counter=True)
    return [ChainOp(outbuf, a, [Step(OP[_UNARY_TABLE[equation.primitive.name]])], equation)]


@primitive('neg', 'abs', 'sign', 'not')
def typed_unary_op(bufferpool, equation):
    """neg/abs/sign on float and integer operands; `not` on bool / integers (reference: float only)."""
    outvar, invar = equation.outvars[0], equation.invars[0]
    assert outvar.aval.shape == invar.aval.shape
    kind = np.dtype(invar.aval.dtype).kind
    name = equation.primitive.name
    table = {'neg': {'f': 'NEG_F', 'i': 'NEG_I', 'u': 'NEG_I'}, 'abs': {'f': 'ABS_F', 'i': 'ABS_I', 'u': 'NOP'},
             'sign': {'f': 'SIGN_F', 'i': 'SIGN_I', 'u': 'CVT_TOBOOL_I'},
             'not': {'b': 'NOT_BOOL', 'i': 'NOT_BITS', 'u': 'NOT_BITS'}}[name]
    if kind not in table:
        raise NotImplementedError(equation)
    a = _operand(bufferpool, invar)
    outbuf = bufferpool.get_buffer(outvar, increment_op_counter=True)
    return [ChainOp(outbuf, a, [Step(OP[table[kind]])], equation)]


@primitive('integer_pow')
def integer_pow(bufferpool, equation):
    """≙ reference ops.py:527-531 (negative exponents handled: quirk Q7)."""
    invar, outvar = equation.invars[0], equation.outvars[0]
    kind = np.dtype(invar.aval.dtype).kind
    y = int(equation.params['y'])
    if kind != 'f' and y < 0:
        raise NotImplementedError(equation)
    a = _operand(bufferpool, invar)
    outbuf = bufferpool.get_buffer(outvar, increment_op_counter=True)
    return [ChainOp(outbuf, a, [Step(OP['IPOW_F' if kind == 'f' else 'IPOW_I'], imm=y & 0xFFFFFFFF)], equation)]


def _convert_opcode(src, dst, equation):
    s, d = np.dtype(src).kind, np.dtype(dst).kind
    if np.dtype(src) != np.uint8:           # packed uint8 is accepted as a SOURCE only: the kernel widens the bytes to u32 on load
        dtype_tag(src)
    dtype_tag(dst)
    if s == d:
        return OP['NOP']
    table = {('f', 'i'): 'CVT_F2I', ('f', 'u'): 'CVT_F2U', ('f', 'b'): 'CVT_TOBOOL_F',
             ('i', 'f'): 'CVT_I2F', ('i', 'u'): 'NOP', ('i', 'b'): 'CVT_TOBOOL_I',
             ('u', 'f'): 'CVT_U2F', ('u', 'i'): 'NOP', ('u', 'b'): 'CVT_TOBOOL_I',
             ('b', 'f'): 'CVT_U2F', ('b', 'i'): 'NOP', ('b', 'u'): 'NOP'}
    if (s, d) not in table:
        raise NotImplementedError(equation)
    return OP[table[(s, d)]]


@primitive('convert_element_type')
def convert_element_type(bufferpool, equation):
    """≙ reference ops.py:563-566; target dtype is the outvar's."""
    invar, outvar = equation.invars[0], equation.outvars[0]
    opcode = _convert_opcode(invar.aval.dtype, outvar.aval.dtype, equation)
    a = _operand(bufferpool, invar)
    outbuf = bufferpool.get_buffer(outvar, increment_op_counter=True)
    return [ChainOp(outbuf, a, [Step(opcode)], equation)]


@primitive('select')
def select(bufferpool, equation):
    """≙ reference ops.py:338-347; a true select, not the reference's arithmetic blend (quirk Q4)."""
    assert equation.invars[0].aval.shape == equation.invars[1].aval.shape \
        == equation.invars[2].aval.shape == equation.outvars[0].aval.shape
    pred, on_true, on_false = [_operand(bufferpool, v) for v in equation.invars]
    outbuf = bufferpool.get_buffer(equation.outvars[0], increment_op_counter=True)
    return [ChainOp(outbuf, pred, [Step(OP['SELECT'], on_true, False, on_false)], equation)]


@primitive('select_n')
def select_n(bufferpool, equation):
    """Modern JAX spelling: select_n(pred, on_false, on_true)."""
    assert len(equation.invars) == 3
    pred, on_false, on_true = [_operand(bufferpool, v) for v in equation.invars]
    outvar = equation.outvars[0]
    for v in equation.invars:
        check_broadcastable(tuple(v.aval.shape), tuple(outvar.aval.shape))
    outbuf = bufferpool.get_buffer(outvar, increment_op_counter=True)
    steps = []
    if np.dtype(equation.invars[0].aval.dtype) != np.bool_:
        steps.append(Step(OP['CVT_TOBOOL_I']))
    steps.append(Step(OP['SELECT'], on_true, False, on_false))
    return [ChainOp(outbuf, pred, steps, equation)]


@primitive('iota')
def iota(bufferpool, equation):
    """≙ reference ops.py:300-303; any `dimension` / rank (the reference asserts dimension == 0 and is
    only right for 1-D: quirk Q7)."""
    outvar = equation.outvars[0]
    dim = int(equation.params['dimension'])
    assert 0 <= dim < builtins.max(len(outvar.aval.shape), 1)
    kind = np.dtype(outvar.aval.dtype).kind
    dtype_tag(outvar.aval.dtype)
    outbuf = bufferpool.get_buffer(outvar, increment_op_counter=True)
    steps = [Step(OP['CVT_U2F'])] if kind == 'f' else []
    return [ChainOp(outbuf, Operand('iota', None, dim), steps, equation)]


# =================================================================================================
# structural (zero-kernel) handlers
@primitive('stop_gradient', 'squeeze', 'bitcast_convert_type', 'copy', 'copy_p', 'expand_dims', 'optimization_barrier',
           'random_wrap', 'random_unwrap')       # typed keys are stored as their uint32[..., 2] threefry key data
def noop(bufferpool, equation):
    """does not perform any operations, simply re-uses the input buffer (≙ reference ops.py:445-458)"""
    assert len(equation.invars) == len(equation.outvars) == 1
    invar, outvar = equation.invars[0], equation.outvars[0]
    if core.is_literal(invar) and np.shape(invar.val) == ():
        # a literal flowing straight to an output: needs real storage
        outbuf = bufferpool.get_buffer(outvar, increment_op_counter=True)
        return [ChainOp(outbuf, _operand(bufferpool, invar), [], equation)]
    inbuf = bufferpool.get_buffer(invar)
    outbuf = inbuf.view(outvar.aval.dtype, outvar.aval.shape)
    bufferpool.set_buffer(outvar, outbuf)
    return []


@primitive('reshape')
def reshape(bufferpool, equation):
    """≙ reference ops.py:263-275."""
    assert equation.params.get('dimensions') is None
    assert len(equation.outvars) == 1
    assert len(equation.invars) == 1
    invar, outvar = equation.invars[0], equation.outvars[0]
    assert invar.aval.dtype == outvar.aval.dtype
    buffer = bufferpool.get_buffer(invar)
    outbuf = buffer.view(outvar.aval.dtype, outvar.aval.shape)
    bufferpool.set_buffer(outvar, outbuf)
    return []


def _inline_call(bufferpool, equation, jaxpr, consts=()):
    assert len(equation.invars) == len(jaxpr.invars)
    assert len(equation.outvars) == len(jaxpr.outvars)
    for cv, c in zip(getattr(jaxpr, 'constvars', []), consts):
        b = bufferpool.get_buffer(cv)
        bufferpool.mark_buffer_as_constant(b, cv, value=c)
    for eq_var, jaxpr_var in zip(equation.invars, jaxpr.invars):
        if core.is_unit(eq_var):
            continue
        if core.is_literal(eq_var) and np.shape(eq_var.val) == ():
            # connect a literal argument: give the inner variable a constant buffer
            b = bufferpool.get_buffer(jaxpr_var)
            bufferpool.mark_buffer_as_constant(b, jaxpr_var, value=np.asarray(eq_var.val))
            continue
        bufferpool.set_buffer(jaxpr_var, bufferpool.get_buffer(eq_var))
    all_ops = analyze_jaxpr(bufferpool, jaxpr)
    for eq_var, jaxpr_var in zip(equation.outvars, jaxpr.outvars):
        if core.is_dropvar(eq_var):
            continue
        if core.is_unit(jaxpr_var):
            bufferpool.set_buffer(eq_var, None)
            continue
        bufferpool.set_buffer(eq_var, bufferpool.get_buffer(jaxpr_var))
    return all_ops


@primitive('xla_call')
def xla_call(bufferpool, equation):
    """≙ reference ops.py:218-244."""
    assert equation.params['device'] is None
    assert equation.params['backend'] is None
    return _inline_call(bufferpool, equation, equation.params['call_jaxpr'])


@primitive('custom_jvp_call_jaxpr')
def custom_jvp_call_jaxpr(bufferpool, equation):
    """≙ reference ops.py:246-260."""
    assert equation.params['num_consts'] == 0
    closed = equation.params['fun_jaxpr']
    return _inline_call(bufferpool, equation, closed.jaxpr, closed.consts)


@primitive('pjit', 'closed_call', 'core_call', 'custom_jvp_call', 'custom_vjp_call', 'remat', 'checkpoint')
def modern_call(bufferpool, equation):
    """Today's spellings of the two call primitives above."""
    inner = equation.params.get('jaxpr', equation.params.get('call_jaxpr', equation.params.get('fun_jaxpr')))
    consts = getattr(inner, 'consts', ())
    inner = getattr(inner, 'jaxpr', inner)
    return _inline_call(bufferpool, equation, inner, consts)


# =================================================================================================
# data movement
def _row_major_strides(shape):
    strides, acc = [0] * len(shape), 1
    for d in range(len(shape) - 1, -1, -1):
        strides[d] = acc
        acc *= shape[d]
    return strides


def _collapse(shape, strides):
    """Drop size-1 dims and merge dims that are contiguous w.r.t. each other."""
    dims = [(s, st) for s, st in zip(shape, strides) if s != 1]
    out = []
    for s, st in dims:
        if out and out[-1][1] == st * s:
            out[-1] = (out[-1][0] * s, st)
        else:
            out.append((s, st))
    if not out:
        out = [(1, 0)]
    return [d[0] for d in out], [d[1] for d in out]


def as_batched_transpose(shape, strides, base):
    """(collapsed) strided copy -> (batch, rows, cols) when it is out[b][c][r] = in[b][r][c] over dense matrices -- e.g. the
    NCHW <-> NHWC permutations -- so that the smem-tiled transpose kernel (coalesced on both sides) can run it; else None."""
    if base != 0:
        return None
    if len(shape) == 2:
        shape, strides = [1] + list(shape), [shape[0] * shape[1]] + list(strides)
    if len(shape) != 3:
        return None
    b, x, y = shape                       # output [b][x][y] reads in[b*sb + x*sx + y*sy]; in is [b][y][x] row-major
    sb, sx, sy = strides
    if sx == 1 and sy == x and (b == 1 or sb == x * y) and x > 1 and y > 1:
        return b, y, x                    # rows = y, cols = x of the source matrices
    return None


def strided_copy_op(outbuf, inbuf, out_shape, in_strides, base, equation):
    shape, strides = _collapse(list(out_shape), list(in_strides))
    bt = as_batched_transpose(shape, strides, base)
    if bt is not None and bt[1] * bt[2] >= 1024:
        return KernelOp(rt.K_TRANSPOSE2D, [outbuf], [inbuf], rt.TransposeParams(rows=bt[1], cols=bt[2], batch=bt[0]), equation)
    if len(shape) > rt.MAX_RANK:
        raise NotImplementedError(equation)
    p = rt.StridedParams()
    p.n = int(np.prod(out_shape, dtype=np.int64))
    p.rank = len(shape)
    for d, (s, st) in enumerate(zip(shape, strides)):
        p.shape[d], p.strides[d] = s, st
    p.base = base
    return KernelOp(rt.K_STRIDED_COPY, [outbuf], [inbuf], p, equation)


@primitive('broadcast_in_dim')
def broadcast_in_dim(bufferpool, equation):
    """≙ reference ops.py:187-215: same element count => pure view; otherwise one copy kernel."""
    bdims = tuple(equation.params['broadcast_dimensions'])
    assert np.all(np.diff(bdims) > 0)
    invar, outvar = equation.invars[0], equation.outvars[0]
    if core.is_literal(invar) and np.shape(invar.val) == ():
        outbuf = bufferpool.get_buffer(outvar, increment_op_counter=True)
        return [ChainOp(outbuf, _operand(bufferpool, invar), [], equation)]
    inbuf = bufferpool.get_buffer(invar)
    if np.prod(inbuf.shape, dtype=np.int64) == np.prod(outvar.aval.shape, dtype=np.int64):
        outbuf = inbuf.view(outvar.aval.dtype, outvar.aval.shape)
        bufferpool.set_buffer(outvar, outbuf)
        return []
    outbuf = bufferpool.get_buffer(outvar, increment_op_counter=True)
    in_strides_src = _row_major_strides(inbuf.shape)
    strides = [0] * len(outbuf.shape)
    for i, d in enumerate(bdims):
        if inbuf.shape[i] == outbuf.shape[d]:
            strides[d] = in_strides_src[i]
        else:
            assert inbuf.shape[i] == 1
    op = strided_copy_op(outbuf, inbuf, outbuf.shape, strides, 0, equation)
    # lets the fusion pass skip the copy: readers can index `inbuf` viewed with this rank-aligned shape
    view_shape = [1] * len(outbuf.shape)
    for i, d in enumerate(bdims):
        view_shape[d] = inbuf.shape[i]
    op.bcast_src = (inbuf, tuple(view_shape))
    return [op]


@primitive('slice')
def slice(bufferpool, equation):
    """≙ reference ops.py:534-547."""
    inbuf = bufferpool.get_buffer(equation.invars[0])
    outbuf = bufferpool.get_buffer(equation.outvars[0], increment_op_counter=True)
    n = len(inbuf.shape)
    start = tuple(equation.params['start_indices'])
    strides = tuple(equation.params['strides'] or (1,) * n)
    src = _row_major_strides(inbuf.shape)
    base = sum(s * st for s, st in zip(start, src))
    return [strided_copy_op(outbuf, inbuf, outbuf.shape, [st * k for st, k in zip(src, strides)], base, equation)]


@primitive('rev')
def rev(bufferpool, equation):
    """≙ reference ops.py:491-502."""
    inbuf = bufferpool.get_buffer(equation.invars[0])
    outbuf = bufferpool.get_buffer(equation.outvars[0], increment_op_counter=True)
    dims = set(int(d) for d in equation.params['dimensions'])
    src = _row_major_strides(inbuf.shape)
    base = sum((inbuf.shape[d] - 1) * src[d] for d in dims)
    strides = [-src[d] if d in dims else src[d] for d in range(len(inbuf.shape))]
    return [strided_copy_op(outbuf, inbuf, outbuf.shape, strides, base, equation)]


@primitive('transpose')
def transpose(bufferpool, equation):
    """≙ reference ops.py:435-442, generalised from permutation == (1,0) to any N-D permutation."""
    perm = tuple(int(p) for p in equation.params['permutation'])
    inbuf = bufferpool.get_buffer(equation.invars[0])
    outbuf = bufferpool.get_buffer(equation.outvars[0], increment_op_counter=True)
    assert sorted(perm) == list(range(len(inbuf.shape)))
    if perm == (1, 0):
        p = rt.TransposeParams(rows=inbuf.shape[0], cols=inbuf.shape[1], batch=1)
        return [KernelOp(rt.K_TRANSPOSE2D, [outbuf], [inbuf], p, equation)]
    src = _row_major_strides(inbuf.shape)
    return [strided_copy_op(outbuf, inbuf, outbuf.shape, [src[p] for p in perm], 0, equation)]


@primitive('concatenate')
def concatenate(bufferpool, equation):
    """≙ reference ops.py:349-369 (2 operands, last axis), extended to any axis and N operands (pairwise)."""
    ndims = len(equation.outvars[0].aval.shape)
    dim = int(equation.params['dimension']) % ndims
    inbufs = [bufferpool.get_buffer(v) for v in equation.invars]
    out_shape = tuple(equation.outvars[0].aval.shape)
    for b in inbufs:
        assert b.shape[:dim] == out_shape[:dim] and b.shape[dim + 1:] == out_shape[dim + 1:]
    outer = int(np.prod(out_shape[:dim], dtype=np.int64))
    inner = int(np.prod(out_shape[dim + 1:], dtype=np.int64))
    ops = []
    if len(inbufs) == 1:
        outbuf = inbufs[0].view(equation.outvars[0].aval.dtype, out_shape)
        bufferpool.set_buffer(equation.outvars[0], outbuf)
        return []
    acc, acc_c = inbufs[0], inbufs[0].shape[dim]
    for i, b in enumerate(inbufs[1:]):
        last = i == len(inbufs) - 2
        cb = b.shape[dim]
        if last:
            dst = bufferpool.get_buffer(equation.outvars[0], increment_op_counter=True)
        else:
            dst = bufferpool.new_temp(out_shape[:dim] + (acc_c + cb,) + out_shape[dim + 1:], b.dtype, 'concat')
            bufferpool.op_counter += 1
        p = rt.ConcatParams(outer=outer, ca=acc_c, cb=cb, inner=inner)
        ops.append(KernelOp(rt.K_CONCAT, [dst], [acc, b], p, equation))
        acc, acc_c = dst, acc_c + cb
    return ops


@primitive('gather')
def gather(bufferpool, equation):
    """≙ reference ops.py:372-401.  Full XLA gather semantics with index_vector_dim = last dim of the
    indices; start indices are clamped as XLA does (the reference does not clamp)."""
    params = equation.params
    dn = params['dimension_numbers']
    operand, indices = [bufferpool.get_buffer(v) for v in equation.invars]
    outbuf = bufferpool.get_buffer(equation.outvars[0], increment_op_counter=True)
    if np.dtype(indices.dtype).kind not in 'iu':
        raise NotImplementedError(equation)
    offset_dims = tuple(dn.offset_dims)
    collapsed = tuple(dn.collapsed_slice_dims)
    sim = tuple(dn.start_index_map)
    slice_sizes = tuple(int(s) for s in params['slice_sizes'])
    if builtins.max(len(outbuf.shape), len(operand.shape)) > rt.MAX_RANK:
        raise NotImplementedError(equation)
    idx_shape = indices.shape if len(indices.shape) > 0 else (1,)
    assert idx_shape[-1] == len(sim)
    batch_shape = idx_shape[:-1]
    noncollapsed = [d for d in range(len(operand.shape)) if d not in collapsed]
    p = rt.GatherParams()
    p.n = outbuf.size
    p.out_rank = len(outbuf.shape)
    p.operand_rank = len(operand.shape)
    p.idx_vec_len = len(sim)
    for d, s in enumerate(outbuf.shape):
        p.out_shape[d] = s
    for d, s in enumerate(operand.shape):
        p.operand_shape[d] = s
        p.slice_sizes[d] = slice_sizes[d]
    for k, od in enumerate(sim):
        p.start_index_map[k] = od
    batch_strides = _row_major_strides(batch_shape)
    oi = bi = 0
    for d in range(len(outbuf.shape)):
        if d in offset_dims:
            p.out_dim_to_operand_dim[d] = noncollapsed[oi]
            oi += 1
        else:
            p.out_dim_to_operand_dim[d] = -1
            p.out_dim_batch_stride[d] = batch_strides[bi]
            bi += 1
    return [KernelOp(rt.K_GATHER, [outbuf], [operand, indices], p, equation)]


@primitive('scatter_add')
def scatter_add(bufferpool, equation):
    """≙ reference ops.py:404-433, which hard-codes two ScatterDimensionNumbers (scatter0/1.comp); this is
    the general XLA scatter-add (index_vector_dim = last), atomics on a copy of the operand."""
    dn = equation.params['dimension_numbers']
    operand, indices, updates = [bufferpool.get_buffer(v) for v in equation.invars]
    outbuf = bufferpool.get_buffer(equation.outvars[0], increment_op_counter=True)
    uwd = tuple(dn.update_window_dims)
    iwd = tuple(dn.inserted_window_dims)
    sd2od = tuple(dn.scatter_dims_to_operand_dims)
    assert operand.shape == outbuf.shape
    idx_shape = indices.shape if len(indices.shape) > 0 else (1,)
    if idx_shape[-1] != len(sd2od):
        raise NotImplementedError(equation)
    scatter_dims = [d for d in range(len(updates.shape)) if d not in uwd]
    if len(scatter_dims) != len(idx_shape) - 1:
        raise NotImplementedError(equation)
    window_operand_dims = [d for d in range(len(operand.shape)) if d not in iwd]
    batch_strides = _row_major_strides(idx_shape[:-1])
    p = rt.ScatterParams()
    p.n_operand, p.n_updates = operand.size, updates.size
    p.operand_rank, p.upd_rank = len(operand.shape), len(updates.shape)
    p.idx_vec_len = len(sd2od)
    p.dtype = dtype_tag(operand.dtype)
    for d, s in enumerate(operand.shape):
        p.operand_shape[d] = s
    for d, s in enumerate(updates.shape):
        p.upd_shape[d] = s
    for k, od in enumerate(sd2od):
        p.scatter_dims_to_operand_dims[k] = od
    wi = si = 0
    for d in range(len(updates.shape)):
        if d in uwd:
            p.upd_dim_to_operand_dim[d] = window_operand_dims[wi]
            wi += 1
        else:
            p.upd_dim_to_operand_dim[d] = -1
            p.upd_dim_batch_stride[d] = batch_strides[si]
            si += 1
    return [KernelOp(rt.K_SCATTER_ADD, [outbuf], [operand, indices, updates], p, equation)]


@primitive('threefry2x32')
def threefry2x32(bufferpool, equation):
    """≙ reference ops.py:550-560."""
    inbufs = [bufferpool.get_buffer(v) for v in equation.invars]
    outbufs = [bufferpool.get_buffer(v) for v in equation.outvars]
    bufferpool.op_counter += 1          # the reference forgets this (quirk Q8)
    assert inbufs[0].shape == inbufs[1].shape
    assert inbufs[2].shape == inbufs[3].shape
    assert outbufs[0].shape == outbufs[1].shape
    p = rt.ThreefryParams(n=outbufs[0].size, key_is_scalar=int(inbufs[0].size == 1))
    return [KernelOp(rt.K_THREEFRY, outbufs, inbufs, p, equation)]


# ---- typed PRNG keys (today's jax.random: `key<fry>[...]` values) ----------------------------
# A typed key array key<fry>[dims] is stored as its threefry key data, uint32[dims + (2,)] (jaxpr_text.KeyAval; a real
# jax.core aval of a key array exposes the same through its dtype's impl).  random_wrap / random_unwrap are therefore views
# (registered with the structural handlers above); the other primitives expand into the primitive chain jax.random used
# to trace to before keys were typed -- threefry2x32 over iota counters -- which is what the device already runs for the
# reference's tests/test_random.py.  This is JAX's ORIGINAL bit layout (`jax_threefry_partitionable=False`, the default up
# to JAX 0.4.x and the one the documented values in tests/golden/jax_random.json come from).
def _expand_traced(bufferpool, equation, fun):
    """Replace `equation` by the primitives `fun` traces to on arguments shaped like the equation's operands."""
    from .frontend.tracing import make_jaxpr
    args = [np.zeros(v.aval.shape, v.aval.dtype) for v in equation.invars]
    closed = make_jaxpr(fun)(*args)
    return _inline_call(bufferpool, equation, closed.jaxpr, closed.consts)


def _require_threefry(equation):
    impl = str(equation.params.get('impl', 'fry'))
    if 'fry' not in impl:
        raise NotImplementedError(f'PRNG implementation {impl!r}: only threefry2x32 keys are supported')


def _key_from_seed(seed):
    """jax._src.prng.threefry_seed for a 32-bit seed: key data = [0, seed as uint32] (the high word of a 32-bit seed is 0)"""
    from .frontend import lax
    from .frontend.tracing import abstractify
    a = abstractify(seed)
    shape = tuple(a.shape)
    lo = seed if np.dtype(a.dtype) == np.uint32 else lax.bitcast_convert_type(seed, np.uint32)
    lo = lax.reshape(lo, shape + (1,))
    return lax.concatenate([np.zeros(shape + (1,), np.uint32), lo], len(shape))


@primitive('random_seed')
def random_seed(bufferpool, equation):
    """seed (int32 / uint32, any shape) -> key<fry>[shape]"""
    _require_threefry(equation)
    dt = np.dtype(equation.invars[0].aval.dtype)
    if dt.itemsize != 4 or dt.kind not in 'iu':
        raise NotImplementedError(f'random_seed: {dt} seeds (64-bit seeds need jax_enable_x64, which the executor does not model)')
    return _expand_traced(bufferpool, equation, _key_from_seed)


@primitive('random_bits')
def random_bits(bufferpool, equation):
    """key<fry>[] -> uint32[shape]: threefry2x32 over iota(size), counters split into halves (odd sizes padded by one)"""
    from .frontend import random as frandom
    if int(equation.params.get('bit_width', 32)) != 32:
        raise NotImplementedError('random_bits: bit_width 8 / 16 / 64')
    if tuple(equation.invars[0].aval.shape) != (2,):
        raise NotImplementedError('random_bits on a key array (vmapped keys)')
    shape = tuple(int(d) for d in equation.params['shape'])
    return _expand_traced(bufferpool, equation, lambda key: frandom._random_bits(key, shape))


@primitive('random_split')
def random_split(bufferpool, equation):
    """key<fry>[] -> key<fry>[shape]"""
    from .frontend import random as frandom, lax
    if tuple(equation.invars[0].aval.shape) != (2,):
        raise NotImplementedError('random_split of a key array')
    shape = tuple(int(d) for d in equation.params['shape'])
    n = int(np.prod(shape, dtype=np.int64))
    return _expand_traced(bufferpool, equation, lambda key: lax.reshape(frandom.split(key, n), shape + (2,)))


@primitive('random_fold_in')
def random_fold_in(bufferpool, equation):
    """(key<fry>[], uint32[]) -> key<fry>[]: threefry2x32(key, seed(data))"""
    from .frontend import random as frandom
    if tuple(equation.invars[0].aval.shape) != (2,) or tuple(equation.invars[1].aval.shape) != ():
        raise NotImplementedError('random_fold_in on arrays')
    return _expand_traced(bufferpool, equation, lambda key, data: frandom.threefry_2x32(key, _key_from_seed(data)))


# =================================================================================================
# reductions
_REDUCE_KINDS = {'reduce_sum': rt.RED_SUM, 'reduce_max': rt.RED_MAX, 'reduce_min': rt.RED_MIN,
                 'reduce_prod': rt.RED_PROD, 'argmax': rt.RED_ARGMAX, 'argmin': rt.RED_ARGMIN}


@primitive(*_REDUCE_KINDS.keys())
def reduce_op(bufferpool, equation):
    """≙ reference ops.py:305-335.  Does not mutate the aval of 1-D inputs (quirk Q9); typed (quirk Q6)."""
    axes = tuple(int(a) for a in equation.params['axes'])
    if axes == ():
        return noop(bufferpool, equation)       # strange but can happen -> noop
    invar, outvar = equation.invars[0], equation.outvars[0]
    inbuf = bufferpool.get_buffer(invar)
    outbuf = bufferpool.get_buffer(outvar, increment_op_counter=True)
    src = _row_major_strides(inbuf.shape)
    keep = [(s, st) for d, (s, st) in enumerate(zip(inbuf.shape, src)) if d not in axes]
    red = [(s, st) for d, (s, st) in enumerate(zip(inbuf.shape, src)) if d in axes]
    kshape, kstrides = _collapse([k[0] for k in keep], [k[1] for k in keep])
    rshape, rstrides = _collapse([r[0] for r in red], [r[1] for r in red])
    if builtins.max(len(kshape), len(rshape)) > rt.MAX_RANK:
        raise NotImplementedError(equation)
    p = rt.ReduceParams()
    p.kind = _REDUCE_KINDS[equation.primitive.name]
    p.dtype = dtype_tag(inbuf.dtype)
    p.n_out = outbuf.size
    p.n_red = int(np.prod([r[0] for r in red], dtype=np.int64))
    p.keep_rank, p.red_rank = len(kshape), len(rshape)
    for d in range(len(kshape)):
        p.keep_shape[d], p.keep_strides[d] = kshape[d], kstrides[d]
    for d in range(len(rshape)):
        p.red_shape[d], p.red_strides[d] = rshape[d], rstrides[d]
    return [KernelOp(rt.K_REDUCE, [outbuf], [inbuf], p, equation)]


def _reduce_window(bufferpool, equation, kind):
    assert len(equation.outvars[0].aval.shape) == 4, NotImplemented       # only 2D implemented
    params = equation.params
    assert tuple(params['base_dilation']) == (1, 1, 1, 1), NotImplemented
    assert tuple(params['window_dilation']) == (1, 1, 1, 1), NotImplemented
    inbuf = bufferpool.get_buffer(equation.invars[0])
    outbuf = bufferpool.get_buffer(equation.outvars[0], increment_op_counter=True)
    p = rt.ReduceWindowParams()
    p.kind, p.dtype = kind, dtype_tag(inbuf.dtype)
    if p.dtype == rt.BOOL:
        raise NotImplementedError(equation)
    for d in range(4):
        p.in_shape[d], p.out_shape[d] = inbuf.shape[d], outbuf.shape[d]
        p.window[d] = params['window_dimensions'][d]
        p.strides[d] = params['window_strides'][d]
        p.pad_lo[d] = params['padding'][d][0]
    return [KernelOp(rt.K_REDUCE_WINDOW, [outbuf], [inbuf], p, equation)]


@primitive('reduce_window_max')
def reduce_window_max(bufferpool, equation):
    """≙ reference ops.py:505-524; padding contributes -inf (lax), not 0.0 (reference quirk Q3)."""
    return _reduce_window(bufferpool, equation, rt.RW_MAX)


@primitive('reduce_window_min')
def reduce_window_min(bufferpool, equation):
    return _reduce_window(bufferpool, equation, rt.RW_MIN)


@primitive('reduce_window_sum')
def reduce_window_sum(bufferpool, equation):
    """avg-pool building block; no reference handler exists."""
    return _reduce_window(bufferpool, equation, rt.RW_SUM)


@primitive('select_and_scatter_add')
def select_and_scatter_add(bufferpool, equation):
    """The transpose of reduce_window_max / _min (max-pool gradient): invars (source, operand); every window of `operand`
    selects one element with `select_prim` (ge: the first maximum; le: the first minimum) and adds its `source` value
    there.  No reference handler exists (jax.grad of a max-pool raises NotImplementedError there)."""
    params = equation.params
    source, operand = [bufferpool.get_buffer(v) for v in equation.invars]
    outbuf = bufferpool.get_buffer(equation.outvars[0], increment_op_counter=True)
    assert len(operand.shape) == 4, NotImplemented
    assert operand.shape == outbuf.shape
    if operand.dtype != np.float32:
        raise NotImplementedError(equation)
    sel = getattr(params['select_prim'], 'name', params['select_prim'])
    if sel not in ('ge', 'le'):
        raise NotImplementedError(equation)
    p = rt.ReduceWindowParams()
    p.kind, p.dtype = (rt.RW_MAX if sel == 'ge' else rt.RW_MIN), rt.F32
    for d in range(4):
        p.in_shape[d], p.out_shape[d] = operand.shape[d], source.shape[d]
        p.window[d] = params['window_dimensions'][d]
        p.strides[d] = params['window_strides'][d]
        p.pad_lo[d] = params['padding'][d][0]
    return [KernelOp(rt.K_SELECT_SCATTER_ADD, [outbuf], [source, operand], p, equation)]


# =================================================================================================
# contractions
@primitive('dot_general')
def dot_general(bufferpool, equation):
    """≙ reference ops.py:277-297 (2-D operands, one contracting dim each, no batch dims).  Extension (SURVEY.md §8 f1;
    the reference asserts batch dims away at ops.py:280): leading batch dimensions shared by both operands, i.e.
    [B..., n, c] x [B..., c, m] in any of the four contracting-dim combinations -- what jnp.matmul / einsum('bij,bjk')
    emit.  Each batch element is one GEMM on the same kernels (records with per-batch address offsets)."""
    assert equation.params['precision'] is None
    dim_numbers = equation.params['dimension_numbers']
    (lc, rc), (lb, rb) = [tuple(map(tuple, d)) for d in dim_numbers]
    assert len(equation.invars) == 2
    assert len(equation.outvars) == 1
    assert all(v.aval.dtype == np.float32 for v in list(equation.invars) + list(equation.outvars))
    nb = len(lb)
    if lb != tuple(range(nb)) or rb != tuple(range(nb)):
        raise NotImplementedError(equation)              # batch dims must lead both operands, in order
    assert tuple(d - nb for d in lc) in [(0,), (1,)]
    assert tuple(d - nb for d in rc) in [(0,), (1,)]
    assert all(len(v.aval.shape) == nb + 2 for v in equation.invars)
    inbufs = [bufferpool.get_buffer(v) for v in equation.invars]
    outbuf = bufferpool.get_buffer(equation.outvars[0], increment_op_counter=True)
    cdim_a, cdim_b = lc[0] - nb, rc[0] - nb
    batch = int(np.prod(outbuf.shape[:nb], dtype=np.int64))
    attrs = dict(n=outbuf.shape[nb], m=outbuf.shape[nb + 1], c=inbufs[0].shape[nb + cdim_a], cdim_a=cdim_a, cdim_b=cdim_b, batch=batch)
    return [ContractionOp('dot', outbuf, inbufs[0], inbufs[1], attrs, equation)]


@primitive('conv_general_dilated')
def conv_general_dilated(bufferpool, equation):
    """≙ reference ops.py:463-487."""
    params = equation.params
    assert params['precision'] is None
    assert params['batch_group_count'] == 1
    assert params['feature_group_count'] == 1
    assert len(equation.outvars[0].aval.shape) == 4       # 2D conv
    assert all(v.aval.dtype == np.float32 for v in list(equation.invars) + list(equation.outvars))
    dn = params['dimension_numbers']
    attrs = dict(lhs_shape=tuple(equation.invars[0].aval.shape), rhs_shape=tuple(equation.invars[1].aval.shape),
                 out_shape=tuple(equation.outvars[0].aval.shape),
                 lhs_spec=tuple(dn.lhs_spec), rhs_spec=tuple(dn.rhs_spec), out_spec=tuple(dn.out_spec),
                 pad_lo=(int(params['padding'][0][0]), int(params['padding'][1][0])),
                 stride=tuple(int(s) for s in params['window_strides']),
                 lhs_dil=tuple(int(s) for s in (params.get('lhs_dilation') or (1, 1))),
                 rhs_dil=tuple(int(s) for s in (params.get('rhs_dilation') or (1, 1))))
    inbufs = [bufferpool.get_buffer(v) for v in equation.invars]
    outbuf = bufferpool.get_buffer(equation.outvars[0], increment_op_counter=True)
    return [ContractionOp('conv', outbuf, inbufs[0], inbufs[1], attrs, equation)]
