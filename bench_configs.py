"""bench.py --config c1|c2|c3|bandwidth: the BASELINE.json configs besides the metric's own workload (c4, bench.py), and the
stand-alone bandwidth kernels ("throughput / roofline for configs C1, C2, C3, and a stand-alone
elementwise / transpose / broadcast bandwidth figure").

Every row: the function goes through vkjax.wrap (trace -> fuse -> plan -> CUDA graph) with device-resident inputs, the
graph replay is timed with CUDA events on the library's stream (L2 flushed between iterations when the working set fits
in L2), and each recorded kernel is timed by an op-by-op profiled replay.  Fractions are against MEASURED_PEAKS.json
(HBM copy bandwidth) and the TF32 peak measured in this run.  Prints ONE JSON line.
"""
import json
import os
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))


def _peaks():
    import bench
    return bench.measured_peaks()


class Harness:
    def __init__(self, args):
        import vkjax_b200 as vkjax
        from vkjax_b200 import runtime as rt
        self.vkjax, self.rt, self.args = vkjax, rt, args
        self.ctx = rt.Context.get(0)
        self.l2 = self.ctx.props().l2_bytes
        self.peaks = _peaks()
        self.hbm = self.peaks['hbm_gbs']
        import bench
        try:
            self.tf32_burst, self.tf32 = bench.measure_tf32_peak(0)
            self.tf32_src = 'measured in this run (cuBLAS fp32 8192^3, TF32 allowed): sustained; burst beside it'
        except Exception as exc:
            self.tf32_burst, self.tf32 = self.peaks['bf16_tflops'] / 2, self.peaks['bf16_tflops_sustained'] / 2
            self.tf32_src = f'MEASURED_PEAKS bf16 / 2 ({exc!r})'

    def run(self, name, f, args, precision='tf32', static_argnums=(), resident=()):
        """-> dict(name, ms (graph replay), rows (per recorded kernel), e2e_ms (host inputs through the public call)).
        `resident`: positions of the arguments that are model state (weights): bound as DeviceArrays, so that their filter
        re-layout is hoisted into the prologue as in a model.  Every other argument is a per-call input: it is uploaded by the
        first call and stays in HBM for the timed replays, but nothing computed from it may be hoisted out of the graph
        (with every argument a DeviceArray the interpreter rightly treats the whole function as call-invariant and the
        replayed graph shrinks to its last launch)."""
        from vkjax_b200 import tree_util
        from vkjax_b200.interpreter import JaxprInterpreter, device_put
        from vkjax_b200.ops import ContractionOp
        ctx, steps, warmup = self.ctx, self.args.steps, max(self.args.warmup, 3)
        fn = self.vkjax.wrap(f, precision=precision, static_argnums=static_argnums)
        dev = [device_put(a) if i in resident else a for i, a in enumerate(args)]
        fn(*dev)
        interp = list(fn._jaxpr_interpreters.values())[0]
        seq = interp.sequence
        total_bytes = sum(t.nbytes for t in interp.bufferpool.tensors) + sum(np.asarray(a).nbytes for a in tree_util.tree_leaves(list(args)))
        flush = total_bytes < 2 * self.l2
        for _ in range(warmup):
            seq.launch()
        ctx.sync()
        ev0, ev1 = ctx.event(), ctx.event()
        ms = 0.0
        for _ in range(steps):
            if flush:
                ctx.flush_l2()
            ctx.record(ev0)
            seq.launch()
            ctx.record(ev1)
            ms += ctx.elapsed_ms(ev0, ev1)
        ms /= steps
        # end to end through the public call: host (pageable numpy) inputs up, results down, every call
        n_e2e = max(3, min(steps, 10))
        host_fn = self.vkjax.wrap(f, precision=precision, static_argnums=static_argnums)
        host_fn(*args)
        t0 = time.perf_counter()
        for _ in range(n_e2e):
            host_fn(*args)
        e2e_ms = (time.perf_counter() - t0) / n_e2e * 1e3
        # per-kernel times
        prof = JaxprInterpreter(interp.jaxpr, profiling=True, precision=precision, device=0)
        leaves = tree_util.tree_leaves(list(args))
        prof.upload_inputs(leaves)
        for _ in range(3):
            prof.sequence.launch()
        ctx.sync()
        acc = None
        for _ in range(5):
            if flush:
                ctx.flush_l2()
            prof.sequence.launch()
            ts = np.array(prof.sequence.timestamps())
            acc = ts if acc is None else acc + ts
        per_op = acc / 5
        rows = []
        for label, op, t in zip(prof.labels, prof.label_ops, per_op):
            if op is None:
                continue
            row = {'kernel': label, 'ms': float(t)}
            if isinstance(op, ContractionOp) and ':' not in label:
                m, n, k, fl, by = op.work()
                res = sum(4 * s_.operand.buf.size for s_ in op.epilogue if s_.operand is not None and s_.operand.kind == 'buf'
                          and tuple(s_.operand.buf.shape) == tuple(op.out.shape))
                by += res
                ideal = max(fl / (self.tf32 * 1e12), by / (self.hbm * 1e9)) * 1e3
                row.update(path=op.path, M=m, N=n, K=k, gflop=fl / 1e9, mbytes=by / 1e6, tflops=fl / (t * 1e-3) / 1e12 if t > 0 else None,
                           gbs=by / 1e6 / t if t > 0 else None, bound='tensor' if fl / by >= self.tf32 * 1e3 / self.hbm else 'hbm',
                           roofline_ms=ideal, frac=ideal / t if t > 0 else None)
            elif isinstance(op, ContractionOp):
                row.update(path='layout / weight prep launch of the contraction above')
            else:
                by = sum(b.nbytes() for b in op.all_buffers())
                row.update(mbytes=by / 1e6, gbs=by / 1e6 / t if t > 0 else None, bound='hbm', frac=(by / 1e6 / t / self.hbm) if t > 0 else None)
            rows.append(row)
        prof.close()
        for fobj in (fn, host_fn):
            for it in fobj._jaxpr_interpreters.values():
                it.close()
        del prof, dev
        out = {'name': name, 'precision': precision, 'graph_ms': ms, 'e2e_ms_host_inputs': e2e_ms, 'l2_flushed_between_iterations': bool(flush),
               'launches': seq.num_launches(), 'kernels': rows}
        return out


# =================================================================================================
def config_c1(h):
    """README minimal example: jnp.dot(x[8,128], W[128,16]) + b (reference README.md:7-22)"""
    from vkjax_b200.frontend import jnp
    rng = np.random.default_rng(0)
    x, W, b = rng.random((8, 128)), rng.random((128, 16)), rng.random(16)
    r = h.run('c1 README dot x[8,128] W[128,16] + b', lambda x, W, b: jnp.dot(x, W) + b, [x, W, b], precision='fp32')
    return {'metric': 'c1_readme_dot_calls_per_sec', 'value': 1e3 / r['e2e_ms_host_inputs'], 'unit': 'calls/s',
            'ms_per_step': r['graph_ms'], 'rows': [r],
            'note': '32.8 kFLOP / 12.9 kB: launch-latency bound; value = calls/s through vkjax.wrap with host (float64 numpy) inputs'}


def config_c2(h):
    """Elegy MLP (reference tests/test_elegy_mlp.py:14-33) forward, batch 4096"""
    from vkjax_b200 import nets
    m = nets.MLP()
    st = m.init(3)
    x = np.random.default_rng(1).integers(0, 256, (4096, 32, 32, 3)).astype(np.float32)
    rows = [h.run(f'c2 MLP b4096 forward [{p}]', lambda x, s: m.apply(s, x), [x, st], precision=p, resident=(1,)) for p in ('tf32', 'fp32')]
    xu8 = x.astype(np.uint8)
    try:
        rows.append(h.run('c2 MLP b4096 forward, uint8 pixels (convert + /255 on the device) [tf32]', lambda x, s: m.apply(s, x), [xu8, st], precision='tf32', resident=(1,)))
    except NotImplementedError as exc:
        rows.append({'name': 'uint8 input', 'error': repr(exc)})
    return {'metric': 'c2_mlp_b4096_images_per_sec', 'value': 4096 / (rows[0]['graph_ms'] * 1e-3), 'unit': 'images/s',
            'ms_per_step': rows[0]['graph_ms'], 'fp32_exact_images_per_s': 4096 / (rows[1]['graph_ms'] * 1e-3), 'rows': rows,
            'note': '7.80 GFLOP, ~67 MB algorithmic; the div-255 elementwise launch is the stand-alone elementwise figure of this config'}


def config_c3(h):
    """conv_general_dilated + reduce_window sweep: reference tests/test_conv.py:66-85 and tests/test_reduce_window.py:14-21
    shapes with the batch dimension -> 256, one roofline row per case"""
    import sys
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    from vkjax_b200.frontend import lax, jnp
    from vkjax_b200.core import ConvDimensionNumbers
    NHWC = ConvDimensionNumbers((0, 3, 1, 2), (3, 2, 0, 1), (0, 3, 1, 2))
    NCHW = ConvDimensionNumbers((0, 1, 2, 3), (0, 1, 2, 3), (0, 1, 2, 3))
    convs = [
        ('conv0 1x1 VALID C5->O33', (256, 100, 100, 5), (1, 1, 5, 33), (1, 1), 'VALID', None, None, NHWC),
        ('conv1 1x1 VALID C33->O11', (256, 100, 100, 33), (1, 1, 33, 11), (1, 1), 'VALID', None, None, NHWC),
        ('conv1 3x3 VALID', (256, 65, 33, 5), (3, 3, 5, 7), (1, 1), 'VALID', None, None, NHWC),
        ('conv1a 3x3 NCHW OIHW', (256, 8, 65, 35), (39, 8, 3, 3), (1, 1), 'VALID', None, None, NCHW),
        ('conv2 1x1 SAME', (256, 17, 9, 12), (1, 1, 12, 11), (1, 1), 'SAME', None, None, NHWC),
        ('conv2 3x3 SAME', (256, 44, 19, 7), (3, 3, 7, 38), (1, 1), 'SAME', None, None, NHWC),
        ('conv2 7x7 SAME', (256, 12, 19, 3), (7, 7, 3, 4), (1, 1), 'SAME', None, None, NHWC),
        ('conv3 uneven pad', (256, 67, 42, 11), (3, 3, 11, 38), (1, 1), [(2, 0), (0, 3)], None, None, NHWC),
        ('conv4 1x1 s2 SAME', (256, 67, 42, 3), (1, 1, 3, 2), (2, 2), 'SAME', None, None, NHWC),
        ('conv4 3x3 s2 SAME', (256, 67, 42, 11), (3, 3, 11, 7), (2, 2), 'SAME', None, None, NHWC),
        ('conv5 s2 rhs_dil 2', (256, 67, 42, 11), (3, 3, 11, 7), (2, 2), 'VALID', (2, 2), None, NHWC),
        ('conv6a lhs_dil 2 pad', (256, 67, 42, 11), (3, 3, 11, 7), (1, 1), [(2, 2), (3, 3)], None, (2, 2), NHWC),
        ('conv6b lhs_dil 2 s2', (256, 67, 42, 11), (3, 3, 11, 7), (2, 2), [(0, 0), (0, 0)], None, (2, 2), NHWC),
    ]
    rs = np.random.RandomState(0)
    rows = []
    for desc, xs, ws, stride, pad, dil, ldil, dn in convs:
        x, w = rs.random_sample(xs).astype(np.float32), rs.random_sample(ws).astype(np.float32)
        f = (lambda stride, pad, dil, ldil, dn: lambda x, w: lax.conv_general_dilated(x, w, stride, pad, lhs_dilation=ldil, rhs_dilation=dil,
                                                                                    dimension_numbers=dn))(stride, pad, dil, ldil, dn)
        for p in ('tf32', 'fp32'):
            rows.append(h.run(f'c3 {desc} [{p}]', f, [x, w], precision=p, resident=(1,)))
    pools = [('max 2x2 s1 VALID', (256, 100, 111, 5), (1, 2, 2, 1), (1, 1, 1, 1), 'VALID'),
             ('max 3x3 s2 SAME', (256, 10, 99, 17), (1, 3, 3, 1), (1, 2, 2, 1), 'SAME'),
             ('max 3x3 s2 SAME ResNet stem pool', (256, 112, 112, 64), (1, 3, 3, 1), (1, 2, 2, 1), 'SAME')]
    for desc, xs, win, st, pad in pools:
        x = rs.random_sample(xs).astype(np.float32)
        rows.append(h.run(f'c3 pool {desc}', (lambda win, st, pad: lambda x: lax.reduce_window(x, -jnp.inf, lax.max, win, st, pad))(win, st, pad), [x]))
        rows.append(h.run(f'c3 avg-pool {desc[4:]}', (lambda win, st, pad: lambda x: lax.reduce_window(x, 0.0, lax.add, win, st, pad) / float(win[1] * win[2]))(win, st, pad), [x]))
    total_ms = sum(r['graph_ms'] for r in rows if '[tf32]' in r['name'] or 'pool' in r['name'])
    return {'metric': 'c3_conv_pool_sweep_b256_total_ms', 'value': total_ms, 'unit': 'ms', 'higher_is_better': False, 'ms_per_step': total_ms, 'rows': rows}


def config_bandwidth(h):
    """stand-alone bandwidth kernels at sizes beyond L2: the north star asks for >= 70 % of HBM peak in elementwise / pooling"""
    from vkjax_b200.frontend import lax, jnp, nn
    rs = np.random.RandomState(0)
    act = rs.random_sample((256, 112, 112, 64)).astype(np.float32)            # 822 MB: ResNet-50 stem output at batch 256
    ch = lambda: rs.random_sample((1, 1, 1, 64)).astype(np.float32)
    rows = []
    rows.append(h.run('BatchNorm + ReLU chain (sub, mul, add, max) on [256,112,112,64], unfused from its conv',
                      lambda x, m, i, o: nn.relu((x - m) * i + o), [act, ch(), ch(), ch()]))
    rows.append(h.run('residual add + ReLU on two [256,56,56,256] tensors', lambda a, b: nn.relu(a + b),
                      [rs.random_sample((256, 56, 56, 256)).astype(np.float32), rs.random_sample((256, 56, 56, 256)).astype(np.float32)]))
    rows.append(h.run('C2 input chain: convert_element_type(int32 -> f32) + div 255 on [16384,32,32,3]',
                      lambda x: x.astype(jnp.float32) / 255.0, [rs.randint(0, 256, (16384, 32, 32, 3)).astype(np.int32)]))
    rows.append(h.run('broadcast_in_dim (1,1,1,64) -> (256,112,112,64), materialised',
                      lambda v: lax.broadcast_in_dim(v, (256, 112, 112, 64), (0, 1, 2, 3)), [ch()]))
    rows.append(h.run('transpose 2-D [8192,16384]', lambda x: jnp.transpose(x), [rs.random_sample((8192, 16384)).astype(np.float32)]))
    rows.append(h.run('transpose NCHW -> NHWC [256,64,56,56]', lambda x: lax.transpose(x, (0, 2, 3, 1)), [rs.random_sample((256, 64, 56, 56)).astype(np.float32)]))
    rows.append(h.run('transpose NHWC -> NCHW [256,56,56,64]', lambda x: lax.transpose(x, (0, 3, 1, 2)), [rs.random_sample((256, 56, 56, 64)).astype(np.float32)]))
    rows.append(h.run('max-pool 3x3 s2 SAME on [256,112,112,64]', lambda x: lax.reduce_window(x, -jnp.inf, lax.max, (1, 3, 3, 1), (1, 2, 2, 1), 'SAME'), [act]))
    rows.append(h.run('global average pool: reduce_sum over (1,2) of [2048,7,7,2048] / 49', lambda x: jnp.mean(x, axis=(1, 2)),
                      [rs.random_sample((2048, 7, 7, 2048)).astype(np.float32)]))
    rows.append(h.run('reduce_sum over the last axis of [65536,4096]', lambda x: jnp.sum(x, axis=1), [rs.random_sample((65536, 4096)).astype(np.float32)]))
    rows.append(h.run('reduce_max over axis 0 of [4096,65536]', lambda x: jnp.max(x, axis=0), [rs.random_sample((4096, 65536)).astype(np.float32)]))
    rows.append(h.run('argmax over the last axis of [262144,1000]', lambda x: jnp.argmax(x, axis=1), [rs.random_sample((262144, 1000)).astype(np.float32)]))
    # the figure of merit covers the launches that stream more than twice the L2 through HBM; a launch whose operands fit the L2
    # (the [2048,2048] `div` behind the average pool: 33 MB, 12 us) is launch-latency bound and listed beside it
    kernels = [k for r in rows for k in r['kernels'] if k.get('frac')]
    big = [k for k in kernels if k['mbytes'] * 1e6 >= 2 * h.l2]
    small = [{'kernel': k['kernel'], 'mbytes': k['mbytes'], 'ms': k['ms'], 'frac': k['frac']} for k in kernels if k['mbytes'] * 1e6 < 2 * h.l2]
    return {'metric': 'bandwidth_kernels_min_frac_of_hbm_peak', 'value': min(k['frac'] for k in big), 'unit': 'fraction of measured HBM copy bandwidth',
            'ms_per_step': sum(r['graph_ms'] for r in rows), 'launches_counted': len(big), 'launches_within_l2_not_counted': small, 'rows': rows}


def run(args):
    h = Harness(args)
    res = {'c1': config_c1, 'c2': config_c2, 'c3': config_c3, 'bandwidth': config_bandwidth}[args.config](h)
    res.setdefault('higher_is_better', True)
    res.update(n_gpus=1, steps=args.steps, warmup=max(args.warmup, 3), scaling='weak', vs_baseline=None, data='synthetic', dtype='tf32 / f32(3xtf32) per row',
               config={'workload': args.config, 'timing': 'CUDA events on the library stream around each graph replay; L2 flushed between iterations when the '
                                                          'working set is below 2x L2 (see l2_flushed_between_iterations per row)'},
               peaks={'hbm_gbs': h.hbm, 'hbm_source': h.peaks['source'], 'tf32_tflops_sustained': h.tf32, 'tf32_tflops_burst': h.tf32_burst, 'tf32_source': h.tf32_src})
    print(json.dumps(res))
    return res
