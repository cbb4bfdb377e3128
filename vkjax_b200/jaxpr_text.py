"""Reader for the text form of a jaxpr -- what `print(jax.make_jaxpr(f)(*args))` shows in today's JAX:

    { lambda ; a:f32[8,128] b:f32[128,16]. let
        c:f32[8,16] = dot_general[
          dimension_numbers=(([1], [0]), ([], []))
          preferred_element_type=float32
        ] a b
        d:f32[8,16] = pjit[name=relu jaxpr={ lambda ; e:f32[8,16]. let f:f32[8,16] = max e 0.0 in (f,) }] c
      in (d,) }

Why it exists: `vkjax.wrap` accepts real `jax.core.ClosedJaxpr` objects, and ops.py
carries the modern spellings (`pjit`, `custom_jvp_call`, `select_n`, `logistic`, `sqrt`, `reduce_window_sum`,
`random_bits` / typed keys, ...), but JAX cannot be installed in this image -- a checked-in dump
(tests/golden/modern_jax_jaxpr.txt) parsed by this reader is the only way to run those handlers on a jaxpr that the
repo's own tracer did NOT emit.  The result is a vkjax_b200.core.ClosedJaxpr: feed it to JaxprInterpreter or to the oracle.

Grammar handled: `{ lambda CONSTVARS ; INVARS . let EQN* in (OUTVARS) }`; variables `name:dtype[dims]`; equations
`OUTS = prim[PARAMS] OPERANDS`; operands are variable names or literals (`0.0`, `1:i32[]`, `True`, `-inf`); parameter values
are Python-literal-like (ints, floats, tuples, lists, None, True/False, dtype names, bare identifiers as strings,
`ConvDimensionNumbers(...)`, `GatherDimensionNumbers(...)`, `ScatterDimensionNumbers(...)`, `GatherScatterMode.X`) or nested
jaxprs in braces; `<function ...>` / `<object ...>` values are kept as opaque strings.  Constants of a closed jaxpr are not
part of the text: pass them with `consts=`.
"""
import re
import typing as tp

import numpy as np

from . import core

_DTYPES = {'f32': np.float32, 'i32': np.int32, 'u32': np.uint32, 'bool': np.bool_, 'f64': np.float64, 'i64': np.int64,
           'u64': np.uint64, 'u8': np.uint8, 'i8': np.int8, 'f16': np.float16, 'bf16': 'bfloat16'}
_DTYPE_NAMES = {'float32': np.float32, 'int32': np.int32, 'uint32': np.uint32, 'bool': np.bool_, 'float64': np.float64,
                'int64': np.int64, 'uint64': np.uint64, 'uint8': np.uint8, 'bool_': np.bool_}


class KeyAval(core.ShapedArray):
    """`key<fry>[dims]`: a typed PRNG key array; stored as uint32[dims + (2,)] (threefry key data)"""
    def __init__(self, shape):
        super().__init__(tuple(shape) + (2,), np.uint32)
        self.key_shape = tuple(shape)


def _parse_aval(text: str):
    m = re.fullmatch(r'(key<\w+>|\w+)\[([\d,\s]*)\]', text.strip())
    if not m:
        raise ValueError(f'cannot parse type {text!r}')
    dims = tuple(int(d) for d in m.group(2).replace(' ', '').split(',') if d)
    if m.group(1).startswith('key<'):
        return KeyAval(dims)
    if m.group(1) not in _DTYPES:
        raise NotImplementedError(f'{m.group(1)} data types currently not supported')
    return core.ShapedArray(dims, _DTYPES[m.group(1)])


def _match_close(text, i, open_ch, close_ch):
    depth = 0
    for j in range(i, len(text)):
        if text[j] == open_ch:
            depth += 1
        elif text[j] == close_ch:
            depth -= 1
            if depth == 0:
                return j
    raise ValueError(f'unbalanced {open_ch}{close_ch}')


class _Scope:
    def __init__(self, counter):
        self.vars: tp.Dict[str, core.Var] = {}
        self.counter = counter

    def define(self, name, aval):
        if name == '_':
            return core.DropVar(aval)
        v = core.Var(self.counter[0], '', aval)
        self.counter[0] += 1
        self.vars[name] = v
        return v


def _parse_literal(tok):
    val, _, ty = tok.partition(':')
    table = {'True': True, 'False': False, 'inf': np.inf, '-inf': -np.inf, 'nan': np.nan}
    if val in table:
        v = table[val]
    else:
        v = float(val) if any(c in val for c in '.e') and not val.lstrip('-').isdigit() else int(val)
    if ty:
        aval = _parse_aval(ty)
        arr = np.asarray(v, aval.dtype)
    else:
        arr = np.asarray(v)
        arr = arr.astype({'f': np.float32, 'i': np.int32, 'u': np.uint32, 'b': np.bool_}[arr.dtype.kind])
    return core.Literal(arr[()], core.ShapedArray((), arr.dtype, weak_type=not ty))


def _split_params(text):
    """'a=1 b=(1, 2) c={ lambda ... }' -> [('a', '1'), ('b', '(1, 2)'), ('c', '{ lambda ... }')]"""
    out, i, n = [], 0, len(text)
    while i < n:
        m = re.compile(r'\s*(\w+)=').match(text, i)
        if not m:
            if text[i:].strip():
                raise ValueError(f'cannot parse parameters at {text[i:i + 40]!r}')
            break
        key, j = m.group(1), m.end()
        depth, k = 0, j
        while k < n:
            ch = text[k]
            if ch in '([{<':
                depth += 1
            elif ch in ')]}>':
                depth -= 1
            elif ch.isspace() and depth == 0:
                nxt = re.compile(r'\s*\w+=').match(text, k)
                if nxt:
                    break
            k += 1
        out.append((key, text[j:k].strip()))
        i = k
    return out


class _NS(dict):
    """evaluation namespace of parameter values: unknown bare identifiers evaluate to their own name"""
    def __missing__(self, key):
        return key


def _param_value(text, scope_counter):
    text = text.strip()
    if text.startswith('{'):
        return _parse_jaxpr(text, scope_counter, closed=True)
    if text.startswith('<'):
        return text
    ns = _NS(__builtins__={}, inf=np.inf, nan=np.nan,
             ConvDimensionNumbers=core.ConvDimensionNumbers,
             GatherDimensionNumbers=lambda offset_dims, collapsed_slice_dims, start_index_map, **kw: core.GatherDimensionNumbers(
                 tuple(offset_dims), tuple(collapsed_slice_dims), tuple(start_index_map)),
             ScatterDimensionNumbers=lambda update_window_dims, inserted_window_dims, scatter_dims_to_operand_dims, **kw: core.ScatterDimensionNumbers(
                 tuple(update_window_dims), tuple(inserted_window_dims), tuple(scatter_dims_to_operand_dims)),
             **{k: np.dtype(v) for k, v in _DTYPE_NAMES.items()})
    try:
        return eval(re.sub(r'\b(\w+)\.(\w+)\b', r'"\1.\2"', text) if re.search(r'[A-Za-z_]\w*\.[A-Za-z_]', text) else text, ns)   # noqa: S307
    except Exception:                                                        # noqa: BLE001 - opaque value
        return text


_MULTI = {'pjit', 'xla_call', 'custom_jvp_call', 'custom_jvp_call_jaxpr', 'custom_vjp_call', 'closed_call', 'core_call', 'remat',
          'checkpoint', 'threefry2x32', 'while', 'cond', 'scan'}


def _parse_jaxpr(text, counter, closed):
    text = text.strip()
    assert text.startswith('{') and text.endswith('}'), text[:40]
    body = text[1:-1].strip()
    m = re.match(r'lambda\s*(.*?);\s*(.*?)\.\s*let\b', body, re.S)
    if not m:
        raise ValueError('expected `{ lambda CONSTS ; ARGS . let ... in (...) }`')
    scope = _Scope(counter)

    def binders(s):
        return [scope.define(n, _parse_aval(t)) for n, t in re.findall(r'(\w+):((?:key<\w+>|\w+)\[[\d,\s]*\])', s)]
    constvars, invars = binders(m.group(1)), binders(m.group(2))
    rest = body[m.end():]
    # the final ` in (...)` at nesting depth 0
    depth, k_in = 0, None
    for i, ch in enumerate(rest):
        if ch in '([{':
            depth += 1
        elif ch in ')]}':
            depth -= 1
        elif depth == 0 and rest.startswith('in', i) and (i == 0 or rest[i - 1].isspace()) and re.match(r'in\s*\(', rest[i:]):
            k_in = i
    if k_in is None:
        raise ValueError('missing `in (...)`')
    eqn_text, out_text = rest[:k_in], rest[k_in + 2:].strip()
    eqns = []
    binder = re.compile(r'([A-Za-z_]\w*):((?:key<\w+>|\w+)\[[\d,\s]*\])|(_)(?=\s)')
    ws = re.compile(r'\s*')
    pos, n = ws.match(eqn_text, 0).end(), len(eqn_text)
    while pos < n:
        outvars = []
        while True:
            mb = binder.match(eqn_text, pos)
            if not mb:
                break
            outvars.append((mb.group(1) or '_', mb.group(2)))
            pos = ws.match(eqn_text, mb.end()).end()
        if not outvars or eqn_text[pos] != '=':
            raise ValueError(f'cannot parse equation at {eqn_text[pos:pos + 60]!r}')
        pos = ws.match(eqn_text, pos + 1).end()
        mp = re.compile(r'[\w\-]+').match(eqn_text, pos)
        prim_name, pos = mp.group(0), mp.end()
        params = {}
        if pos < n and eqn_text[pos] == '[':
            e = _match_close(eqn_text, pos, '[', ']')
            params = {k: _param_value(v, counter) for k, v in _split_params(eqn_text[pos + 1:e])}
            pos = e + 1
        invs = []
        while True:
            pos = ws.match(eqn_text, pos).end()
            if pos >= n or binder.match(eqn_text, pos):
                break
            mt = re.compile(r'\S+').match(eqn_text, pos)
            tok, pos = mt.group(0), mt.end()
            invs.append(scope.vars[tok] if tok in scope.vars else _parse_literal(tok))
        outs_v = [scope.define(name, _parse_aval(ty) if ty else core.ShapedArray((), np.float32)) for name, ty in outvars]
        prim = core.Primitive(prim_name, multiple_results=prim_name in _MULTI or len(outs_v) > 1)
        eqns.append(core.JaxprEqn(invs, outs_v, prim, params))
    outs = []
    for tok in out_text.strip().lstrip('(').rstrip(')').replace(',', ' ').split():
        name = tok.split(':')[0]
        outs.append(scope.vars[name] if name in scope.vars else _parse_literal(tok))
    jaxpr = core.Jaxpr(constvars, invars, outs, eqns)
    return core.ClosedJaxpr(jaxpr, []) if closed else jaxpr


def parse_jaxpr(text: str, consts: tp.Sequence = ()) -> core.ClosedJaxpr:
    """text of one jaxpr -> ClosedJaxpr (constvars bound to `consts`)"""
    start = text.index('{')
    end = _match_close(text, start, '{', '}')
    jaxpr = _parse_jaxpr(text[start:end + 1], [0], closed=False)
    assert len(jaxpr.constvars) == len(consts), f'{len(jaxpr.constvars)} constvars, {len(consts)} consts given'
    return core.ClosedJaxpr(jaxpr, [np.asarray(c) for c in consts])


def parse_file(path: str) -> tp.Dict[str, core.ClosedJaxpr]:
    """A fixture file holds several dumps, each introduced by a line `### name`."""
    out, name, buf = {}, None, []
    for line in open(path):
        if line.startswith('###'):
            if name is not None:
                out[name] = ''.join(buf)
            name, buf = line[3:].strip(), []
        elif not line.startswith('#'):
            buf.append(line)
    if name is not None:
        out[name] = ''.join(buf)
    return {k: parse_jaxpr(v) for k, v in out.items()}
