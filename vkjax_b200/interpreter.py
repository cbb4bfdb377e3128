"""JaxprInterpreter: analyse a ClosedJaxpr once, record it into a CUDA graph, replay it per call.

≙ reference vkjax/kompute_jaxpr_interpreter.py.  Same life cycle -- __init__ analyses the jaxpr into
ops, plans buffers and records a sequence (:27-63); run() uploads inputs, evaluates the sequence and
reads the outputs back (:66-89); get_profiling_info() returns per-op device times (:91-95) -- with the
Kompute manager/sequence replaced by the C-ABI library (vkjax_b200/runtime.py -> libb2jax.so).

Differences, all motivated in SURVEY.md Appendix E:
  * one device context per process/device instead of one per interpreter (Q2);
  * inputs given as `DeviceArray` stay resident and are not re-uploaded; pass-through outputs are
    returned without a device round trip (Q2);
  * `fuse=True` (default) runs the fusion pass of vkjax_b200/fusion.py; `precision` selects the
    contraction path: 'fp32' (3xTF32 tensor cores, fp32-class accuracy, default), 'tf32' (single pass,
    rtol 2e-3), 'simt' (fp32 FMA kernels in reference summation order).
"""
import ctypes as C
import os
import typing as tp
import weakref

import numpy as np

from . import core, ops, fusion, tree_util
from . import runtime as rt
from .buffers import BufferPool, Buffer, canonicalize_host, from_device_words
from .ops import ChainOp, ContractionOp, KernelOp, OP


class DeviceArray:
    """An immutable array resident in device memory (see `device_put`).  Passing it to a wrapped
    function skips the per-call host->device copy the reference performs for every input
    (kompute_jaxpr_interpreter.py:72-75)."""
    def __init__(self, ctx: rt.Context, value: np.ndarray):
        value = np.asarray(value)
        self.shape, self.dtype = value.shape, value.dtype
        words = np.ascontiguousarray(canonicalize_host(value)).reshape(-1)
        self.dtype = np.dtype(bool) if value.dtype == np.bool_ else words.dtype
        self.ctx = ctx
        self.nbytes = max((words.nbytes + 3) // 4 * 4, 4)
        self.addr = ctx.alloc(self.nbytes)
        ctx.upload(self.addr, words)

    def numpy(self) -> np.ndarray:
        n = int(np.prod(self.shape, dtype=np.int64))
        words = np.empty(max(n if self.dtype != np.uint8 else (n + 3) // 4, 1), np.uint32)
        self.ctx.download(self.addr, words)
        return from_device_words(words, self.dtype, self.shape)

    __array__ = lambda self, dtype=None, copy=None: self.numpy() if dtype is None else self.numpy().astype(dtype)

    def __del__(self):
        try:
            self.ctx.free(self.addr)
        except Exception:
            pass


def device_put(x, device=None):
    """Pytree of arrays -> pytree of DeviceArray on the context's GPU."""
    ctx = rt.Context.get(device)
    return tree_util.tree_map(lambda a: a if isinstance(a, DeviceArray) else DeviceArray(ctx, np.asarray(a)), x)


def leaf_shape_dtype(x):
    if isinstance(x, DeviceArray):
        return tuple(x.shape), np.dtype(x.dtype)
    a = np.asarray(x)
    return tuple(a.shape), a.dtype


# =================================================================================================
# lowering: op records -> b2j_seq_record calls
def _fill_epilogue(epi: rt.Epilogue, op: ContractionOp, bufs: list, full_offset: int = 0):
    """`full_offset`: byte offset of this launch's slice within full-shape operands (batched dot_general)."""
    epi.n_steps = len(op.epilogue)
    for i, s in enumerate(op.epilogue):
        kind = fusion.epilogue_operand_kind(op, s.operand)
        st = epi.steps[i]
        st.op = s.op
        st.flags = rt.STEP_SWAP if s.swap else 0
        if kind == 'imm':
            st.kind, st.imm = rt.EPK_IMM, s.operand.imm
        else:
            st.kind = rt.EPK_CHANNEL if kind == 'channel' else rt.EPK_FULL
            st.buf = len(bufs)
            bufs.append(s.operand.buf.addr + (full_offset if kind == 'full' else 0))


def lower_chain(op: ChainOp):
    out = op.out
    if len(out.shape) > rt.MAX_RANK:
        raise NotImplementedError(op.equation)
    p = rt.EltParams()
    p.n = out.size
    p.rank = len(out.shape)
    for d, s in enumerate(out.shape):
        p.shape[d] = s
    slots: tp.List[Buffer] = []

    def slot(o: ops.Operand):
        for i, b in enumerate(slots):
            if b.same_storage(o.buf) and b.shape == o.buf.shape:
                return i
        if len(slots) == rt.ELT_MAX_IN:
            raise NotImplementedError('elementwise chain with more than %d inputs' % rt.ELT_MAX_IN)
        i = len(slots)
        slots.append(o.buf)
        kind, mod, strides = fusion.classify_operand(o.buf.shape, out.shape)
        p.in_[i].kind, p.in_[i].mod = kind, mod
        p.in_[i].elem = 1 if o.buf.dtype == np.uint8 else 0
        if strides is not None:
            for d, st in enumerate(strides):
                p.in_[i].strides[d] = st
        return i

    def src(o: tp.Optional[ops.Operand]):
        if o is None:
            return rt.SRC_NONE, 0
        if o.kind == 'imm':
            return rt.SRC_IMM, o.imm
        if o.kind == 'iota':
            return rt.SRC_IOTA, o.imm
        return slot(o), 0

    p.init_src, p.init_imm = src(op.init)
    p.n_steps = len(op.steps)
    for i, s in enumerate(op.steps):
        st = p.steps[i]
        st.op = s.op
        st.src, imm = src(s.operand)
        st.imm = s.imm if s.operand is None else imm
        st.flags = rt.STEP_SWAP if s.swap else 0
        st.src2, st.imm2 = src(s.operand2) if s.operand2 is not None else (rt.SRC_NONE, 0)
    p.n_in = len(slots)
    return [(rt.K_ELTWISE, [out.addr] + [b.addr for b in slots], p, op.label(), False)]


NHWC_SPEC = (0, 3, 1, 2)          # (batch, feature, spatial0, spatial1) positions of an NHWC array


def plan_contraction(op: ContractionOp, pool: BufferPool, precision: str):
    """Chooses the kernel path and reserves workspace.  Called after fusion, before buffer planning.

    Every conv_general_dilated the reference's handler accepts (ops.py:463-487: any dimension specs, low padding, strides,
    lhs / rhs dilation) is brought to the one form the tcgen05 kernels compute -- NHWC activations x [O][Kpad] filters ->
    NHWC -- by cheap data-movement launches around the tensor-core kernel:
      * lhs in another layout (e.g. NCHW, reference tests/test_conv.py:25-28)  -> strided copy to NHWC
      * lhs dilation (transposed convolution / conv input gradients)            -> zero stuffing (B2J_K_DILATE)
      * channel counts TMA cannot address (C % 32, C % 4 for GEMM-like)          -> channel padding / space-to-depth fold
      * any rhs layout (HWIO, OIHW, ...)                                         -> weight_prep reads it through rhs_spec
      * output in another layout                                                 -> NHWC temp + strided copy
      * output channel counts that are not a multiple of 4                       -> tail-guarded scalar stores in the epilogue
    Only negative padding and full-tensor epilogue operands of a non-NHWC output stay on the fp32 FMA kernel."""
    a = op.attrs
    if op.what == 'dot':
        flops = 2 * a['n'] * a['m'] * a['c']
        tc_ok = a.get('batch', 1) <= 64               # one weight-prep + GEMM launch pair per batch element
    else:
        o = a['rhs_shape'][a['rhs_spec'][0]]
        kk = a['rhs_shape'][a['rhs_spec'][1]] * a['rhs_shape'][a['rhs_spec'][2]] * a['rhs_shape'][a['rhs_spec'][3]]
        flops = 2 * int(np.prod(a['out_shape'], dtype=np.int64)) * kk
        full_epi = any(fusion.epilogue_operand_kind(op, s_.operand) == 'full' for s_ in op.epilogue)
        tc_ok = (a['pad_lo'][0] >= 0 and a['pad_lo'][1] >= 0 and (tuple(a['out_spec']) == NHWC_SPEC or not full_epi)
                 and all(d > 0 for d in a['lhs_shape']) and all(d > 0 for d in a['out_shape']))
    if precision == 'simt' or not tc_ok or flops < ops.TC_MIN_FLOPS:
        op.path = 'direct'
        return
    op.path = 'tc'
    x3 = precision == 'fp32'
    a['relayout'] = None
    f32 = np.float32
    if op.what == 'dot':
        k, n = a['c'], a['m']
        if k % 4 != 0:                      # TMA needs a 16-byte row pitch: pad the contraction dim
            a['relayout'] = plan_relayout(1, 1, a['n'], k, 1, 1, (1, 1), (0, 0), (1, 1), 1, a['n'], gemm_like=True)
    else:
        k, n = kk, o
        L, R, O_ = a['lhs_shape'], a['rhs_shape'], a['out_shape']
        sl, sp, so = tuple(a['lhs_spec']), tuple(a['rhs_spec']), tuple(a['out_spec'])
        nb, c, h, w = L[sl[0]], L[sl[1]], L[sl[2]], L[sl[3]]
        dh, dw = a['lhs_dil']
        hd, wd = (h - 1) * dh + 1, (w - 1) * dw + 1
        a['x_nhwc'] = ls = (nb, hd, wd, c)                                    # what the conv kernel reads
        a['y_nhwc'] = os_ = (O_[so[0]], O_[so[2]], O_[so[3]], O_[so[1]])      # what it writes
        a['lhs_nhwc_t'] = pool.new_temp((nb, h, w, c), f32, 'lhs_nhwc') if sl != NHWC_SPEC else None
        a['lhs_dil_t'] = pool.new_temp(ls, f32, 'lhs_dilated') if (dh, dw) != (1, 1) else None
        a['out_nhwc_t'] = pool.new_temp(os_, f32, 'out_nhwc') if so != NHWC_SPEC else None
        kh, kw = R[sp[2]], R[sp[3]]
        gemm_like = (kh == 1 and kw == 1 and a['stride'] == (1, 1) and tuple(a['pad_lo']) == (0, 0)
                     and os_[1] == ls[1] and os_[2] == ls[2])
        a['rows'] = (ls[3] % 32 != 0 and not gemm_like and a['lhs_dil_t'] is None
                     and rows_mode_ok(ls, kh, kw, a['stride'], a['pad_lo'], a['rhs_dil'], os_, kk, x3))
        if a['rows']:
            pass                                    # fed from the raw input rows (B2J_CT_ROWS): no re-layout, K padded to 32 only
        elif ls[3] % (4 if gemm_like else 32) != 0:
            a['relayout'] = plan_relayout(ls[0], ls[1], ls[2], ls[3], kh, kw, a['stride'], a['pad_lo'], a['rhs_dil'],
                                          os_[1], os_[2], gemm_like)
    if a['relayout']:
        k = a['relayout']['k']
    kpad = (k + 31) // 32 * 32
    a['kpad'], a['x3'] = kpad, x3
    nbatch = a.get('batch', 1) if op.what == 'dot' else 1
    op.temps = [pool.new_temp((nbatch * n, kpad), np.float32, 'wt_hi')]
    if x3:
        op.temps.append(pool.new_temp((nbatch * n, kpad), np.float32, 'wt_lo'))
    if op.what == 'dot' and a['cdim_a'] == 0:
        a['lhs_t'] = pool.new_temp((nbatch * a['n'], a['c']), np.float32, 'lhs_t')
        op.temps.append(a['lhs_t'])
    if op.what == 'conv':
        op.temps += [t for t in (a['lhs_nhwc_t'], a['lhs_dil_t'], a['out_nhwc_t']) if t is not None]
    if a['relayout']:
        a['xprime'] = pool.new_temp((nbatch,) + tuple(a['relayout']['dst_shape']), np.float32, 'x_relayout')
        op.temps.append(a['xprime'])


def rows_mode_ok(ls, kh, kw, stride, pad_lo, dil, os_, k, x3=False):
    """Admission rule of the B2J_CT_ROWS path of the tcgen05 kernel (rows_geometry in vkjax_b200/csrc/conv_tc2.cuh is its C twin):
    few-channel k x k convolutions -- the 3-channel ResNet stem -- are fed from staged raw input rows instead of a re-laid-out
    copy.  B2J_ENABLE_ROWS=0 switches the path off (the space-to-depth fold + im2col path of round 1 is used instead)."""
    if os.environ.get('B2J_ENABLE_ROWS', '1') == '0':
        return False
    n, h, w, c = ls
    ow, o = os_[2], os_[3]
    tile_px = min(ow, 128)
    kpad = (k + 31) // 32 * 32
    # kpad: the weight matrix stays resident in shared memory (8 k-blocks single pass, 5 split hi / lo)
    if not (o <= 64 and tuple(dil) == (1, 1) and kpad <= (160 if x3 else 256) and kh <= 256 and pad_lo[0] >= 0 and pad_lo[1] >= 0 and (w * c) % 16 == 0):
        return False
    for es in (4, 1):                 # float32 source, and packed uint8 should an input chain be folded in later
        align, step, kwc = 16 // es, stride[1] * c, kw * c
        if kwc + align - 1 > 256:
            return False
        sp = (256 - kwc - (align - 1)) // step + 1
        while sp > 0 and (sp * step) % align:
            sp -= 1
        if sp == 0 or -(-tile_px // sp) > 32 or -(-tile_px // sp) * kh * 256 * es > 28 * 1024:
            return False
    return True


def _row_major(shape):
    st, acc = [0] * len(shape), 1
    for d in range(len(shape) - 1, -1, -1):
        st[d] = acc
        acc *= shape[d]
    return st


def _strided_record(dst_addr, src_addr, out_shape, in_strides, label):
    """out[i] = in[sum coord_d(i) * in_strides[d]] as a B2J_K_STRIDED_COPY record (layout changes around the conv kernel)."""
    shape, strides = ops._collapse(list(out_shape), list(in_strides))
    bt = ops.as_batched_transpose(shape, strides, 0)
    if bt is not None and bt[1] * bt[2] >= 1024:      # NCHW <-> NHWC and the like: smem-tiled batched transpose
        return (rt.K_TRANSPOSE2D, [dst_addr, src_addr], rt.TransposeParams(rows=bt[1], cols=bt[2], batch=bt[0]), label, False)
    if len(shape) > rt.MAX_RANK:
        raise NotImplementedError(label)
    p = rt.StridedParams()
    p.n = int(np.prod(out_shape, dtype=np.int64))
    p.rank = len(shape)
    for d, (s_, st) in enumerate(zip(shape, strides)):
        p.shape[d], p.strides[d] = s_, st
    p.base = 0
    return (rt.K_STRIDED_COPY, [dst_addr, src_addr], p, label, False)


_PRE_OPS = None


def fuse_input_chains(all_ops, keep_ids) -> int:
    """Host/wire side of the call: images arrive as packed uint8 and the model starts with
    `x.astype(float32) / 255` (reference tests/test_elegy_mlp.py:21).  When that elementwise chain -- convert_element_type
    plus at most two steps with an immediate operand -- feeds ONLY the re-layout of a tensor-core convolution (the 3-channel
    stem is always re-laid out), it is folded into the re-layout's source load: the f32 image tensor (4 bytes per pixel
    value written and read again) never exists.  Returns the number of chains folded."""
    global _PRE_OPS
    if _PRE_OPS is None:
        _PRE_OPS = {OP[n] for n in ('ADD_F', 'SUB_F', 'MUL_F', 'DIV_F')}
    readers = {}
    for op in all_ops:
        for b in op.inputs():
            readers[id(b._ph)] = readers.get(id(b._ph), 0) + 1
    producer = {id(b._ph): op for op in all_ops for b in op.outputs()}
    n = 0
    for op in list(all_ops):
        a = getattr(op, 'attrs', None)
        if not isinstance(op, ContractionOp) or op.what != 'conv' or op.path != 'tc' or not (a.get('relayout') or a.get('rows')):
            continue
        if a.get('lhs_nhwc_t') is not None or a.get('lhs_dil_t') is not None:
            continue
        sid = id(op.lhs._ph)
        prod = producer.get(sid)
        if not isinstance(prod, ChainOp) or readers.get(sid, 0) != 1 or sid in keep_ids:
            continue
        if prod.init.kind != 'buf' or tuple(prod.init.buf.shape) != tuple(op.lhs.shape):
            continue
        steps = list(prod.steps)
        u8 = prod.init.buf.dtype == np.uint8
        if u8:
            if not steps or steps[0].op != OP['CVT_U2F']:
                continue
            steps = steps[1:]
        elif prod.init.buf.dtype != np.float32:
            continue
        if len(steps) > 2 or (not u8 and not steps):
            continue
        if not all(s_.op in _PRE_OPS and s_.operand is not None and s_.operand.kind == 'imm' and not s_.swap for s_ in steps):
            continue
        if a.get('rows'):
            # the rows path widens uint8 through a table and zero-fills padding BEFORE the chain: uint8 sources only, and the
            # chain has to map 0 to 0 (x / 255 does); anything else keeps its stand-alone elementwise launch
            if not u8 or _chain_of_zero(steps) != 0.0:
                continue
        a['lhs_pre'] = dict(u8=u8, steps=[(s_.op, s_.operand.imm) for s_ in steps])
        op.lhs = prod.init.buf
        op.equations = prod.equations + op.equations
        all_ops.remove(prod)
        n += 1
    return n


def _chain_of_zero(steps):
    """The fused input chain (ADD/SUB/MUL/DIV with immediates) applied to 0.0 in float32."""
    x = np.float32(0.0)
    with np.errstate(all='ignore'):
        for s_ in steps:
            imm = np.array(s_.operand.imm, np.uint32).view(np.float32)
            name = {OP['ADD_F']: 'add', OP['SUB_F']: 'subtract', OP['MUL_F']: 'multiply', OP['DIV_F']: 'divide'}[s_.op]
            x = np.float32(getattr(np, name)(x, imm))
    return float(x)


def plan_tf32_rounding(all_ops, keep_ids):
    """Single-pass TF32 mode: the tensor core TRUNCATES fp32 operands to TF32, a systematic -2^-11 shrink per layer
    (1.1 % of the logits' norm over ResNet-50).  Round-to-nearest has to happen before the operand reaches shared memory;
    where every reader of a tensor-core contraction's output is itself a tensor-core contraction (as lhs or residual) or a
    max-pool feeding such contractions (max commutes with rounding), the producer stores the value already rounded
    (B2J_CT_ROUND_OUT_TF32) -- the rounding its readers would need, done once, for free in the epilogue."""
    readers = {}
    for op in all_ops:
        for b in op.inputs():
            readers.setdefault(id(b._ph), []).append(op)

    def is_max_pool(op):
        return isinstance(op, KernelOp) and op.kernel_id == rt.K_REDUCE_WINDOW and op.params.kind == rt.RW_MAX

    def feeds_only_tc(sid, depth=0):
        if sid in keep_ids or depth > 2:
            return False
        rs = readers.get(sid, [])
        if not rs:
            return False
        for r in rs:
            if isinstance(r, ContractionOp) and r.path == 'tc' and not r.attrs.get('x3') and \
                    (id(r.lhs._ph) == sid or any(s_.operand is not None and s_.operand.kind == 'buf' and id(s_.operand.buf._ph) == sid
                                                 and tuple(s_.operand.buf.shape) == tuple(r.out.shape) for s_ in r.epilogue)) \
                    and id(r.rhs._ph) != sid:
                continue
            if is_max_pool(r) and feeds_only_tc(id(r.outs[0]._ph), depth + 1):
                continue
            return False
        return True

    n = 0
    for op in all_ops:
        if isinstance(op, ContractionOp) and op.path == 'tc' and not op.attrs.get('x3') and feeds_only_tc(id(op.out._ph)):
            op.attrs['round_out'] = True
            n += 1
    return n


def plan_relayout(batch, h, w, c, kh, kw, stride, pad_lo, dil, oh, ow, gemm_like):
    """Activation layout for channel counts the TMA tensor maps cannot address (C % 32 != 0 for k x k, C % 4 != 0 for
    GEMM-like problems).  Two schemes (b2j_relayout_params in include/b2jax.h):

      fold  equal strides s, no filter dilation, s*s*C <= 32: the s x s stride window and F horizontally adjacent
            window cells go into the channel dimension (space-to-depth), so e.g. the ResNet stem (7x7/2 over 3 channels,
            K = 147) becomes 4x2 taps with dilation (1, 2) over 32 channels (24 used, K' = 256) instead of 49 taps
            over 32 padded channels (K' = 1568).  Padding is materialised by the re-layout, the new conv has none.
      pad   otherwise: channels zero-padded to a multiple of 32, geometry unchanged.
    """
    s = stride[0]
    if not gemm_like and stride[0] == stride[1] and tuple(dil) == (1, 1) and s * s * c <= 32 and (kh > 1 or kw > 1):
        taps_h, cells_w = -(-kh // s), -(-kw // s)
        f = min(cells_w, 32 // (s * s * c))
        taps_w = -(-cells_w // f)
        fmap = [(di, s * bp + dj, ch, 1) for bp in range(f) for di in range(s) for dj in range(s) for ch in range(c)]
        dst = (batch, oh + taps_h - 1, ow + (taps_w - 1) * f, 32)
        return dict(mode='fold', dst_shape=dst, fold=(s, s), pad=(int(pad_lo[0]), int(pad_lo[1])), map=fmap,
                    conv=dict(h=dst[1], w=dst[2], c=32, kh=taps_h, kw=taps_w, stride=(1, 1), pad=(0, 0), dil=(1, f)),
                    wprep=dict(cpad=32, taps_h=taps_h, taps_w=taps_w, tap_h=s, tap_w=s * f), k=taps_h * taps_w * 32)
    cp = (c + 31) // 32 * 32
    return dict(mode='pad', dst_shape=(batch, h, w, cp), fold=(1, 1), pad=(0, 0), map=[],
                conv=dict(h=h, w=w, c=cp, kh=kh, kw=kw, stride=tuple(stride), pad=(int(pad_lo[0]), int(pad_lo[1])), dil=tuple(dil)),
                wprep=dict(cpad=cp, taps_h=kh, taps_w=kw, tap_h=1, tap_w=1), k=kh * kw * cp)


def _relayout_records(op, src_addr, src_dims, dst_offset=0):
    """(re-layout launch record, fields to set on the weight-prep params)."""
    r = op.attrs['relayout']
    n, h, w, c = src_dims
    p = rt.RelayoutParams(batch=n, h=h, w=w, c=c, oh=r['dst_shape'][1], ow=r['dst_shape'][2], oc=r['dst_shape'][3],
                          fold_h=r['fold'][0], fold_w=r['fold'][1], pad_h=r['pad'][0], pad_w=r['pad'][1], n_map=len(r['map']),
                          round_tf32=0 if op.attrs['x3'] else 1)
    for j, (dh, dw, ch, valid) in enumerate(r['map']):
        p.map[j].dh, p.map[j].dw, p.map[j].c, p.map[j].valid = dh, dw, ch, valid
    pre = op.attrs.get('lhs_pre')
    if pre:                                            # fused input chain: uint8 -> f32 (+ up to two immediate steps, e.g. / 255)
        p.src_u8 = 1 if pre['u8'] else 0
        p.pre_n = len(pre['steps'])
        for j, (opc, imm) in enumerate(pre['steps']):
            p.pre_op[j], p.pre_imm[j] = opc, imm
    return (rt.K_RELAYOUT, [op.attrs['xprime'].addr + dst_offset, src_addr], p, op.label() + ':relayout', False)


def _fill_wprep_fold(wp, r):
    for key, val in r['wprep'].items():
        setattr(wp, key, val)
    wp.n_map = len(r['map'])
    for j, (dh, dw, ch, valid) in enumerate(r['map']):
        wp.map[j].dh, wp.map[j].dw, wp.map[j].c, wp.map[j].valid = dh, dw, ch, valid


def lower_contraction(op: ContractionOp):
    a = op.attrs
    recs = []
    if op.path == 'direct':
        bufs = [op.out.addr, op.lhs.addr, op.rhs.addr]
        if op.what == 'dot':
            for b in range(a.get('batch', 1)):          # batched dot_general: one launch per batch element
                bufs = [op.out.addr + 4 * b * a['n'] * a['m'], op.lhs.addr + 4 * b * a['n'] * a['c'], op.rhs.addr + 4 * b * a['c'] * a['m']]
                p = rt.DotParams(n=a['n'], m=a['m'], c=a['c'], cdim_a=a['cdim_a'], cdim_b=a['cdim_b'])
                _fill_epilogue(p.epi, op, bufs, 4 * b * a['n'] * a['m'])
                recs.append((rt.K_DOT, bufs, p, op.label(), False))
            return recs
        p = rt.ConvDirectParams()
        for d in range(4):
            p.lhs_shape[d], p.rhs_shape[d], p.out_shape[d] = a['lhs_shape'][d], a['rhs_shape'][d], a['out_shape'][d]
            p.lhs_spec[d], p.rhs_spec[d], p.out_spec[d] = a['lhs_spec'][d], a['rhs_spec'][d], a['out_spec'][d]
        for d in range(2):
            p.pad_lo[d], p.stride[d], p.lhs_dil[d], p.rhs_dil[d] = a['pad_lo'][d], a['stride'][d], a['lhs_dil'][d], a['rhs_dil'][d]
        _fill_epilogue(p.epi, op, bufs)
        return [(rt.K_CONV_DIRECT, bufs, p, op.label(), False)]

    # tensor-core path: weight prep (+ optional lhs transpose) then the tcgen05 kernel
    x3 = a['x3']
    wt_hi = op.temps[0]
    wt_lo = op.temps[1] if x3 else None
    prec = rt.PREC_TF32X3 if x3 else rt.PREC_TF32
    rl = a.get('relayout')
    if op.what == 'dot':
        # view rhs as a 1x1 HWIO (cdim_b == 0: [C, M]) or OHWI-like (cdim_b == 1: [M, C]) filter
        shape4 = (1, 1) + tuple(op.rhs.shape[-2:])
        spec = (3, 2, 0, 1) if a['cdim_b'] == 0 else (2, 3, 0, 1)
        n_, m_, c_ = a['n'], a['m'], a['c']
        for b in range(a.get('batch', 1)):               # batched dot_general: one weight-prep + GEMM per batch element
            wp = rt.WeightPrepParams(kpad=a['kpad'], split=1 if x3 else 2)
            for d in range(4):
                wp.rhs_shape[d], wp.rhs_spec[d] = shape4[d], spec[d]
            if rl:
                _fill_wprep_fold(wp, rl)
            w_off = 4 * b * m_ * a['kpad']
            recs.append((rt.K_WEIGHT_PREP, [wt_hi.addr + w_off, op.rhs.addr + 4 * b * c_ * m_] + ([wt_lo.addr + w_off] if x3 else []), wp,
                         op.label() + ':weight_prep', getattr(op, 'prep_hoisted', False)))
            lhs_addr = op.lhs.addr + 4 * b * n_ * c_
            if a['cdim_a'] == 0:
                lhs_t = a['lhs_t']
                recs.append((rt.K_TRANSPOSE2D, [lhs_t.addr + 4 * b * n_ * c_, lhs_addr], rt.TransposeParams(rows=c_, cols=n_, batch=1),
                             op.label() + ':lhs_transpose', False))
                lhs_addr = lhs_t.addr + 4 * b * n_ * c_
            k = c_
            if rl:
                x_off = 4 * b * int(np.prod(rl['dst_shape'], dtype=np.int64))
                recs.append(_relayout_records(op, lhs_addr, (1, 1, n_, c_), x_off))
                lhs_addr, k = a['xprime'].addr + x_off, rl['k']
            bufs = [op.out.addr + 4 * b * n_ * m_, lhs_addr, wt_hi.addr + w_off, (wt_lo.addr + w_off) if x3 else 0]
            p = rt.GemmTcParams(m=n_, n=m_, k=k, kpad=a['kpad'], precision=prec,
                                flags=rt.CT_ROUND_OUT_TF32 if a.get('round_out') else 0)
            _fill_epilogue(p.epi, op, bufs, 4 * b * n_ * m_)
            recs.append((rt.K_GEMM_TC, bufs, p, op.label(), False))
        return recs
    wp = rt.WeightPrepParams(kpad=a['kpad'], split=1 if x3 else 2)
    shape4, spec = a['rhs_shape'], a['rhs_spec']
    for d in range(4):
        wp.rhs_shape[d], wp.rhs_spec[d] = shape4[d], spec[d]
    if rl:
        _fill_wprep_fold(wp, rl)
    recs.append((rt.K_WEIGHT_PREP, [wt_hi.addr, op.rhs.addr] + ([wt_lo.addr] if x3 else []), wp, op.label() + ':weight_prep',
                 getattr(op, 'prep_hoisted', False)))
    rs = a['rhs_shape']
    ls, os_ = a['x_nhwc'], a['y_nhwc']
    lhs_addr = op.lhs.addr
    if a['lhs_nhwc_t'] is not None:                       # any lhs layout -> NHWC
        L, sl = a['lhs_shape'], a['lhs_spec']
        st = _row_major(L)
        t = a['lhs_nhwc_t']
        recs.append(_strided_record(t.addr, lhs_addr, t.shape, [st[sl[0]], st[sl[2]], st[sl[3]], st[sl[1]]], op.label() + ':lhs_to_nhwc'))
        lhs_addr = t.addr
    if a['lhs_dil_t'] is not None:                        # lhs dilation -> zero stuffing
        t = a['lhs_dil_t']
        dh, dw = a['lhs_dil']
        p = rt.DilateParams(batch=ls[0], h=(ls[1] - 1) // dh + 1, w=(ls[2] - 1) // dw + 1, c=ls[3], oh=ls[1], ow=ls[2], dil_h=dh, dil_w=dw)
        recs.append((rt.K_DILATE, [t.addr, lhs_addr], p, op.label() + ':lhs_dilate', False))
        lhs_addr = t.addr
    g = dict(h=ls[1], w=ls[2], c=ls[3], kh=rs[a['rhs_spec'][2]], kw=rs[a['rhs_spec'][3]], stride=a['stride'],
             pad=a['pad_lo'], dil=a['rhs_dil'])
    if rl:
        recs.append(_relayout_records(op, lhs_addr, ls))
        lhs_addr, g = a['xprime'].addr, rl['conv']
    p = rt.ConvTcParams(batch=ls[0], h=g['h'], w=g['w'], c=g['c'], kh=g['kh'], kw=g['kw'],
                        o=os_[3], oh=os_[1], ow=os_[2], pad_h=g['pad'][0], pad_w=g['pad'][1],
                        stride_h=g['stride'][0], stride_w=g['stride'][1], dil_h=g['dil'][0], dil_w=g['dil'][1],
                        kpad=a['kpad'], precision=prec, flags=rt.CT_ROUND_OUT_TF32 if a.get('round_out') else 0)
    if a.get('rows'):                                     # raw input rows in, optional fused uint8 -> f32 (/ 255) chain
        p.flags |= rt.CT_ROWS | (0 if x3 else rt.CT_ROUND_IN_TF32)
        pre = a.get('lhs_pre')
        if pre:
            p.src_u8 = 1 if pre['u8'] else 0
            p.pre_n = len(pre['steps'])
            for j, (opc, imm) in enumerate(pre['steps']):
                p.pre_op[j], p.pre_imm[j] = opc, imm
    out_t = a['out_nhwc_t']
    bufs = [op.out.addr if out_t is None else out_t.addr, lhs_addr, wt_hi.addr, wt_lo.addr if x3 else 0]
    _fill_epilogue(p.epi, op, bufs)
    recs.append((rt.K_CONV_TC, bufs, p, op.label(), False))
    if out_t is not None:                                 # NHWC -> the requested output layout
        so = a['out_spec']
        nhwc_st = _row_major(os_)                          # strides of (n, h, w, c) in the temp
        role = {so[0]: nhwc_st[0], so[2]: nhwc_st[1], so[3]: nhwc_st[2], so[1]: nhwc_st[3]}
        recs.append(_strided_record(op.out.addr, out_t.addr, a['out_shape'], [role[d] for d in range(4)], op.label() + ':nhwc_to_out'))
    return recs


def lower(op):
    """op record -> [(kernel_id, bufs, params, label, hoisted)]; `hoisted` records go to the prologue sequence."""
    if isinstance(op, ChainOp):
        recs = lower_chain(op)
    elif isinstance(op, ContractionOp):
        recs = lower_contraction(op)
    else:
        recs = [(op.kernel_id, [b.addr for b in op.outs] + [b.addr for b in op.ins], op.params, op.label(), False)]
    if getattr(op, 'hoisted', False):
        recs = [r[:4] + (True,) for r in recs]
    return recs


# =================================================================================================
class JaxprInterpreter:
    def __init__(self, jaxpr, static_argnums: tp.Tuple[int] = (), profiling: bool = False, reuse_buffers: bool = True,
                 fuse: bool = True, precision: str = 'fp32', device: tp.Optional[int] = None, allgather_outputs: bool = False,
                 dry_run: bool = False, resident_inputs: tp.Optional[tp.Sequence[bool]] = None, hoist: bool = True):
        if precision not in ops.PRECISIONS:
            raise ValueError(f'precision must be one of {ops.PRECISIONS}')
        self.jaxpr = jaxpr
        self.static_argnums = tuple(static_argnums)
        self.profiling = profiling
        self.fuse = fuse
        self.precision = precision
        self.allgather_outputs = allgather_outputs
        # which (flattened, non-static) inputs arrive as DeviceArray: work that depends only on those and on constants
        # (filter re-layout, BatchNorm parameter folding) is hoisted into a prologue that is replayed only when such
        # an input is rebound -- the reference recomputes it, and re-uploads every weight, on every call (quirk Q2)
        self.resident_inputs = tuple(resident_inputs) if (resident_inputs is not None and hoist and fuse) else None
        # dry_run: analyse + plan only, no device (host-logic tests on machines without a GPU)
        self.ctx = None if dry_run else rt.Context.get(device)
        self.workgroup_size = 1 if dry_run else get_maximum_workgroup_size(self.ctx)
        self.bufferpool = BufferPool(self.ctx, self.workgroup_size, reuse_buffers)
        self._res = _DeviceResources(self.ctx, self.bufferpool)
        self._finalizer = weakref.finalize(self, self._res.release)
        self.analyze_closed_jaxpr(jaxpr)

    def analyze_closed_jaxpr(self, jaxpr):
        """Starts the analysis of the top level jaxpr.  Records operations into a sequence and creates
        required buffers (≙ reference kompute_jaxpr_interpreter.py:27-63)."""
        pool = self.bufferpool
        assert len(jaxpr.consts) == len(jaxpr.jaxpr.constvars)
        for constvar, constval in zip(jaxpr.jaxpr.constvars, jaxpr.consts):
            b = pool.get_buffer(constvar)
            pool.mark_buffer_as_constant(b, constvar, value=np.asarray(constval))
        for v in list(jaxpr.jaxpr.invars) + list(jaxpr.jaxpr.outvars):
            b = pool.get_buffer(v)
            if b is not None:
                # I/O buffers get their own allocation (≙ reference :36-41)
                pool.mark_buffer_as_constant(b, v, None)

        self.input_buffers = [pool.get_buffer(v) for v in jaxpr.jaxpr.invars]
        self.all_ops = ops.analyze_jaxpr(pool, jaxpr.jaxpr)
        self.output_buffers = [pool.get_buffer(v) for v in jaxpr.jaxpr.outvars]
        self.unfused_ops = len(self.all_ops)

        if self.fuse:
            keep = {id(b._ph) for b in self.output_buffers if b is not None}
            self.all_ops = fusion.fuse(self.all_ops, keep)
        for op in self.all_ops:
            if isinstance(op, ContractionOp):
                plan_contraction(op, pool, self.precision)
        self.n_input_chains_fused = fuse_input_chains(self.all_ops, {id(b._ph) for b in self.output_buffers if b is not None})
        self.n_rounded = 0
        if self.precision == 'tf32':
            self.n_rounded = plan_tf32_rounding(self.all_ops, {id(b._ph) for b in self.output_buffers if b is not None})
        self.n_hoisted = self._plan_hoisting() if self.resident_inputs and any(self.resident_inputs) else 0
        if self.fuse or any(isinstance(op, ContractionOp) and op.temps for op in self.all_ops):
            pool.recompute_accesses(self.all_ops, live_out=self.output_buffers)
        else:
            for b in self.output_buffers:        # unfused: the end-of-program read of every output (≙ reference :45)
                if b is not None and not b.is_constant():
                    b.accesses.append(max(pool.op_counter, len(self.all_ops)))
        pool.create_tensors()

        # pass-through outputs (an outvar that is an invar): returned without a device round trip
        invar_pos = {core.hashable(v): i for i, v in enumerate(jaxpr.jaxpr.invars)}
        self.passthrough = [invar_pos.get(core.hashable(v)) if not core.is_literal(v) else None
                            for v in jaxpr.jaxpr.outvars]

        if self.ctx is None:
            self.sequence = None
            self.labels = [op.label() for op in self.all_ops]
            return

        self.sequence = rt.Sequence(self.ctx, self.profiling)
        self.prologue = rt.Sequence(self.ctx, self.profiling) if self.n_hoisted else None
        self._prologue_done = False
        self.labels = []
        self.prologue_labels = []
        self.label_ops = []          # op record behind each recorded kernel (bench.py's per-layer table)
        self._param_keepalive = []
        for op in self.all_ops:
            for kid, bufs, params, label, hoisted in lower(op):
                self._param_keepalive.append(params)
                if hoisted and self.prologue is not None:
                    self.prologue.record(kid, bufs, params)
                    self.prologue_labels.append(label)
                    continue
                self.sequence.record(kid, bufs, params)
                self.labels.append(label)
                self.label_ops.append(op)
        if self.prologue is not None:
            self.prologue.finalize()

        # multi-GPU: batch-sharded ranks all-gather their outputs over NVLink behind the same graph.  Pass-through
        # outputs (replicated state handed back unchanged) are not gathered.
        self.gather_buffers = None
        if self.allgather_outputs and self.ctx.nranks > 1:
            self.gather_buffers = []
            for k, b in enumerate(self.output_buffers):
                if self.passthrough[k] is not None:
                    self.gather_buffers.append(None)
                    continue
                nbytes = b.nbytes()
                addr = self.ctx.alloc(max(nbytes, 4) * self.ctx.nranks)
                self.sequence.record_allgather(b.addr, addr, nbytes)
                self.labels.append('all_gather')
                self.label_ops.append(None)
                self.gather_buffers.append(addr)
                self._res.gather.append(addr)
        self.sequence.finalize()

        # pinned staging for inputs / outputs (≙ the host-mapped side of kp.Tensor)
        self._in_stage = [rt.HostBuffer(self.ctx, max(b.nbytes(), 4)) if b is not None else None for b in self.input_buffers]
        mult = self.ctx.nranks if self.gather_buffers is not None else 1
        self._out_stage = [rt.HostBuffer(self.ctx, max(b.nbytes(), 4) * mult) for b in self.output_buffers]
        self._resident = [None] * len(self.input_buffers)      # DeviceArray currently bound to each input slot
        self.h2d_bytes = 0
        self.d2h_bytes = 0

    def _plan_hoisting(self) -> int:
        """Marks the ops (and the weight-prep half of tensor-core contractions) whose inputs are all call-invariant --
        resident DeviceArray inputs, constants, outputs of already hoisted ops -- and makes their outputs persistent."""
        pool = self.bufferpool
        inv = set()
        for b, r in zip(self.input_buffers, self.resident_inputs):
            if b is not None and r:
                inv.add(id(b._ph))
        for b in pool.buffers.values():
            if b is not None and b.tensor is not None and b.tensor.initial_value is not None:
                inv.add(id(b._ph))
        out_ids = {id(b._ph) for b in self.output_buffers if b is not None}

        def persist(b):
            if not b.is_constant():
                b.accesses += [0, float('inf')]

        n = 0
        for op in self.all_ops:
            if isinstance(op, ContractionOp):
                if op.path == 'tc' and id(op.rhs._ph) in inv:
                    op.prep_hoisted = True
                    for t in op.temps[:2 if op.attrs['x3'] else 1]:
                        persist(t)
                    n += 1
                continue
            if all(id(b._ph) in inv for b in op.inputs()) and not any(id(b._ph) in out_ids for b in op.outputs()):
                op.hoisted = True
                for b in op.outputs():
                    inv.add(id(b._ph))
                    persist(b)
                n += 1
        return n

    # ---------------------------------------------------------------------------------------------
    def _flatten_args(self, X):
        X = tree_util.tree_leaves([x for i, x in enumerate(X) if i not in self.static_argnums])
        if len(self.input_buffers) != len(X):
            raise TypeError(f'Expected {len(self.input_buffers)} input arguments, received {len(X)}')
        return X

    def upload_inputs(self, X, only_resident=False):
        """≙ reference :72-75, minus the copies that are not needed.  Returns True when a resident input was rebound
        (the prologue has to be replayed)."""
        self.h2d_bytes = 0
        rebound = False
        for i, (buf, x, var) in enumerate(zip(self.input_buffers, X, self.jaxpr.jaxpr.invars)):
            if buf is None or (only_resident and not isinstance(x, DeviceArray)):
                continue
            if isinstance(x, DeviceArray):
                if self._resident[i] is not x:
                    if x.nbytes < buf.nbytes():
                        raise TypeError(f'input {i}: DeviceArray too small')
                    self.ctx.copy_async(buf.addr, x.addr, buf.nbytes())
                    self._resident[i] = x
                    rebound = True
                continue
            self._resident[i] = None
            arr = np.asarray(x)
            words = canonicalize_host(arr if arr.dtype == var.aval.dtype else arr.astype(var.aval.dtype))
            n = words.nbytes                    # == buf.nbytes() for 32-bit dtypes; packed uint8 buffers are padded to whole words
            if n == 0:
                continue
            if self.ctx.is_pinned(words):
                self.ctx.upload_async(buf.addr, words.ctypes.data, n)           # straight from user pinned memory
            else:
                stage = self._in_stage[i].array[:n].view(words.dtype)
                np.copyto(stage, words.reshape(-1), casting='no')
                self.ctx.upload_async(buf.addr, self._in_stage[i].ptr, n)
            self.h2d_bytes += n
        if rebound:
            self._prologue_done = False
        return rebound

    def download_outputs(self, X=None):
        outs = []
        self.d2h_bytes = 0
        pending = []
        for k, (buf, var) in enumerate(zip(self.output_buffers, self.jaxpr.jaxpr.outvars)):
            src_pos = self.passthrough[k]
            if X is not None and src_pos is not None:
                outs.append(X[src_pos])
                continue
            gathered = self._downloads_gathered(k)
            mult = self.ctx.nranks if gathered else 1
            n = buf.nbytes() * mult
            addr = self.gather_buffers[k] if gathered else buf.addr
            if n:
                self.ctx.download_async(addr, self._out_stage[k].ptr, n)
            self.d2h_bytes += n
            outs.append(None)
            pending.append((k, mult))
        self.ctx.sync()
        for k, mult in pending:
            buf = self.output_buffers[k]
            shape = buf.shape
            if mult > 1:
                shape = (shape[0] * mult,) + tuple(shape[1:]) if len(shape) else (mult,)
            n_words = int(np.prod(shape, dtype=np.int64))
            words = self._out_stage[k].array[:n_words * 4].view(np.uint32)
            outs[k] = from_device_words(words, buf.dtype, shape)
        return tuple(outs)

    def _downloads_gathered(self, k) -> bool:
        """Does THIS rank read back the all-gathered copy of output k?  allgather_outputs=True: every rank (SPMD: each
        process returns the whole batch); 'root': the all-gather still runs on every rank over NVLink, but only rank 0
        downloads the gathered tensor -- the others read back just their own shard (8 x less device->host traffic at 8 GPUs)."""
        if self.gather_buffers is None or self.gather_buffers[k] is None:
            return False
        return self.allgather_outputs != 'root' or self.ctx.rank == 0

    def run(self, *X, return_all=False):
        """Executes a previously recorded sequence with actual data (≙ reference :66-89)."""
        return self.run_leaves(self._flatten_args(X), return_all=return_all)

    def run_leaves(self, X, return_all=False):
        """run() for already flattened, static-free inputs (Function flattens them for its cache key anyway)."""
        if len(self.input_buffers) != len(X):
            raise TypeError(f'Expected {len(self.input_buffers)} input arguments, received {len(X)}')
        self.upload_inputs(X)
        self.launch()
        output_values = self.download_outputs(X)
        if not return_all:
            return output_values
        all_arrays = {}
        for var, buf in self.bufferpool.buffers.items():
            if buf is None or buf.tensor is None or buf.tensor.addr is None:
                all_arrays[var] = None
                continue
            words = np.empty(max(buf.size, 1), np.uint32)
            self.ctx.download(buf.addr, words)
            all_arrays[var] = from_device_words(words, buf.dtype, buf.shape)
        return output_values, all_arrays

    def launch(self):
        """Enqueues the prologue (only if a resident input was rebound since it last ran) and the per-call sequence."""
        if self.prologue is not None and not self._prologue_done:
            self.prologue.launch()
            self._prologue_done = True
        self.sequence.launch()

    def run_many(self, arg_batches: tp.Sequence[tp.Sequence], lanes: int = 2):
        """Runs the recorded sequence once per argument tuple with the host->device copy of batch i+1 overlapping the
        replay of batch i (copy lanes on a second stream; b2j_lane_* in include/b2jax.h).  The reference's run() is strictly
        upload -> eval -> download (kompute_jaxpr_interpreter.py:72-81); this is the same per-batch work, pipelined.
        Returns the list of output tuples."""
        return self.run_many_leaves([self._flatten_args(a) for a in arg_batches], lanes)

    def run_many_leaves(self, Xs: tp.Sequence[tp.Sequence], lanes: int = 2):
        n = len(Xs)
        if n == 0:
            return []
        ctx = self.ctx
        lanes = max(1, min(lanes, 4, n))
        host_idx = [i for i, (buf, x) in enumerate(zip(self.input_buffers, Xs[0])) if buf is not None and not isinstance(x, DeviceArray)]
        # every batch must bind the same resident (DeviceArray) inputs: they are uploaded / the prologue is replayed once
        for k in range(1, n):
            for i, (x0, xk) in enumerate(zip(Xs[0], Xs[k])):
                if (isinstance(x0, DeviceArray) or isinstance(xk, DeviceArray)) and x0 is not xk:
                    raise TypeError(f'run_many / Function.map: input {i} of batch {k} is a different DeviceArray than in batch 0; '
                                    'device-resident inputs (weights, state) must be the same objects in every batch')
        if len(self._res.lane_dev) < lanes:
            for lane in self._res.lane_dev:
                for addr in lane.values():
                    ctx.free(addr)
            self._res.lane_dev = [{i: ctx.alloc(max(self.input_buffers[i].nbytes(), 4)) for i in host_idx} for _ in range(lanes)]
            # stream-ordered allocations (cudaMallocAsync on the context stream) are first written from the COPY stream:
            # the pool may hand out a block whose cudaFreeAsync is still pending behind running kernels, so the allocation
            # has to be complete on the context stream before the copy stream may touch it
            ctx.sync()
            self._lane_host = [{i: None for i in host_idx} for _ in range(lanes)]
        self._lane_dev = self._res.lane_dev
        h2d = 0

        def stage(k):
            nonlocal h2d
            lane = k % lanes
            for i in host_idx:
                buf, var = self.input_buffers[i], self.jaxpr.jaxpr.invars[i]
                arr = np.asarray(Xs[k][i])
                words = canonicalize_host(arr if arr.dtype == var.aval.dtype else arr.astype(var.aval.dtype))
                nb = words.nbytes
                if nb == 0:
                    continue
                if ctx.is_pinned(words):
                    src = words.ctypes.data
                else:
                    if self._lane_host[lane][i] is None:
                        self._lane_host[lane][i] = rt.HostBuffer(ctx, buf.nbytes())
                    ctx.lane_sync(lane)                                  # the previous upload from this staging area is done
                    hb = self._lane_host[lane][i]
                    np.copyto(hb.array[:nb].view(words.dtype), words.reshape(-1), casting='no')
                    src = hb.ptr
                ctx.lane_upload(lane, self._lane_dev[lane][i], src, nb)
                h2d += nb

        # resident inputs are bound once (taken from the first batch)
        self.upload_inputs([x if isinstance(x, DeviceArray) else None for x in Xs[0]], only_resident=True)
        out_bufs = [(k, b) for k, b in enumerate(self.output_buffers) if self.passthrough[k] is None]
        mult = ctx.nranks if self.gather_buffers is not None else 1
        # pinned result staging: a ring of lanes + 2 slots; slot contents are converted to numpy as soon as their
        # device->host copy has completed (event), while later batches are still running
        ring = lanes + 2
        if not hasattr(self, '_out_ring') or len(self._out_ring) < ring:
            self._out_ring = [([rt.HostBuffer(ctx, max(b.nbytes(), 4) * (mult if self._downloads_gathered(k) else 1))
                                for k, b in out_bufs], ctx.event()) for _ in range(ring)]
            # Device-side staging for the result downloads: the outputs of batch k are copied (device to device, microseconds) into
            # ring slot k, and the device -> host copy runs from there on its own stream (b2j_lane_download), so the context stream
            # goes straight on with batch k+1.  With the copy on the context stream itself every replay waited for the previous
            # download: 0.35 ms per step for the 8 MB of gathered logits on the root rank at N = 8 (e2e scaling 0.91).
            for addr in getattr(self._res, 'out_stage', []):
                ctx.free(addr)
            self._res.out_stage = [ctx.alloc(max(b.nbytes(), 4) * (mult if self._downloads_gathered(k) else 1))
                                   for _ in range(ring) for k, b in out_bufs]
            ctx.sync()
        n_out = len(out_bufs)
        results = [None] * n

        def collect(j):
            stage, ev = self._out_ring[j % ring]
            ctx.event_sync(ev)
            outs = [None] * len(self.output_buffers)
            for (ko, b), hb in zip(out_bufs, stage):
                gathered = self._downloads_gathered(ko)
                shape = b.shape
                if gathered:
                    shape = (shape[0] * mult,) + tuple(shape[1:]) if len(shape) else (mult,)
                n_words = int(np.prod(shape, dtype=np.int64))
                outs[ko] = from_device_words(hb.array[:n_words * 4].view(np.uint32), b.dtype, shape)
            for ko, src_pos in enumerate(self.passthrough):
                if src_pos is not None:
                    outs[ko] = Xs[j][src_pos]
            results[j] = tuple(outs)

        for k in range(min(lanes, n)):
            stage(k)
        d2h = 0
        for k in range(n):
            lane = k % lanes
            if k >= ring:
                collect(k - ring)                                        # frees ring slot k % ring
            if host_idx:
                ctx.lane_acquire(lane)
                for i in host_idx:
                    ctx.copy_async(self.input_buffers[i].addr, self._lane_dev[lane][i], self.input_buffers[i].nbytes())
                    self._resident[i] = None
                ctx.lane_release(lane)
            self.launch()
            stage_k, ev = self._out_ring[k % ring]
            for j, ((ko, b), hb) in enumerate(zip(out_bufs, stage_k)):
                gathered = self._downloads_gathered(ko)
                nb = b.nbytes() * (mult if gathered else 1)
                if nb:
                    dev = self._res.out_stage[(k % ring) * n_out + j]
                    ctx.copy_async(dev, self.gather_buffers[ko] if gathered else b.addr, nb)
                    ctx.lane_download(hb.ptr, dev, nb)
                d2h += nb
            if any(b.nbytes() for _, b in out_bufs):
                ctx.lane_download_record(ev)
            else:
                ctx.record(ev)
            if k + lanes < n:
                stage(k + lanes)
        for j in range(max(0, n - ring), n):
            collect(j)
        self.h2d_bytes, self.d2h_bytes = h2d // n, d2h // n
        return results

    def get_profiling_info(self):
        """[(label, milliseconds)] per recorded kernel of the last run (≙ reference :91-95; the labels
        'data2device'/'data2host' of the reference correspond to copies that now happen outside the
        recorded sequence)."""
        return list(zip(self.labels, self.sequence.timestamps()))

    def close(self):
        """Frees every device allocation of this interpreter (arena, own tensors, gather / lane buffers).  Also runs when
        the interpreter is garbage collected (weakref.finalize), so dropping a Function does not leak device memory."""
        self._finalizer()


class _DeviceResources:
    """Device allocations owned by one JaxprInterpreter, released by its finalizer (must not reference the interpreter)."""
    def __init__(self, ctx, pool):
        self.ctx, self.pool = ctx, pool
        self.lane_dev = []
        self.gather = []
        self.out_stage = []

    def release(self):
        ctx = self.ctx
        if ctx is None or not ctx.handle:
            return
        for lane in self.lane_dev:
            for addr in lane.values():
                ctx.free(addr)
        self.lane_dev = []
        for addr in self.gather:
            if addr:
                ctx.free(addr)
        self.gather = []
        for addr in self.out_stage:
            ctx.free(addr)
        self.out_stage = []
        self.pool.release()


def get_maximum_workgroup_size(ctx: rt.Context):
    """≙ reference kompute_jaxpr_interpreter.py:100-102."""
    props = ctx.props()
    return min(props.max_threads_per_block, props.max_block_dim_x)
