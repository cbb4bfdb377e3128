"""Diagnostic: where does the pipelined end-to-end step time go?  Times Function.map over 20 batches with pieces of the
pipeline stubbed out (no host->device copy / no result download / neither), against the bare graph replay."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from vkjax_b200 import nets, runtime as rt
from vkjax_b200.elegy import vkModel
ctx = rt.Context.get(int(os.environ.get("LOCAL_RANK", 0)))
B, N = 256, 20
m = vkModel(nets.ResNet50(), precision='tf32'); m.init(seed=0)
x = ctx.pinned_empty((4 * B, 224, 224, 3), np.float32); x[...] = 0.5
y = m.predict_on_batch(x[:B])
f = m.call_pred_step_jit
interp = list(f._jaxpr_interpreters.values())[0]
for _ in range(5): interp.sequence.launch()
ctx.sync(); t0 = time.perf_counter()
for _ in range(N): interp.sequence.launch()
ctx.sync(); print('graph only: %.3f ms/step' % ((time.perf_counter() - t0) / N * 1e3))
batches = [(x[(i % 4) * B:(i % 4 + 1) * B], m.states, False, False) for i in range(N)]


def timed(label):
    f.map(batches[:4])
    ctx.sync(); t0 = time.perf_counter(); f.map(batches); dt = time.perf_counter() - t0
    print('%-46s %.3f ms/step' % (label, dt / N * 1e3))


timed('map, everything')
real = {k: getattr(ctx, k) for k in ('lane_upload', 'lane_download', 'copy_async', 'lane_acquire', 'lane_release')}
ctx.lane_upload = lambda *a: None
timed('map, no host->device copy')
ctx.lane_download = lambda *a: None
timed('map, no H2D, no D2H')
ctx.copy_async = lambda *a: None
timed('map, no H2D, no D2H, no device staging copies')
ctx.lane_acquire = lambda *a: None; ctx.lane_release = lambda *a: None
timed('map, ... and no lane events')
for k, v in real.items():
    setattr(ctx, k, v)
ctx.lane_download = lambda *a: None
timed('map, H2D only (no D2H)')
ctx.lane_download = real['lane_download']
# how long does one replay take when an upload runs beside it?  (HBM / power contention)
dev = ctx.alloc(x[:B].nbytes)
ev0, ev1 = ctx.event(), ctx.event()
for with_copy in (0, 1):
    ctx.sync()
    ctx.record(ev0)
    for k in range(N):
        if with_copy:
            ctx.lane_upload(k % 2, dev, x[:B].ctypes.data, x[:B].nbytes)
        interp.sequence.launch()
    ctx.record(ev1); ctx.sync()
    print('replay x%d %s: %.3f ms/step (device events)' % (N, 'with a concurrent 154 MB upload per step' if with_copy else 'alone', ctx.elapsed_ms(ev0, ev1) / N))
import cProfile, pstats
pr = cProfile.Profile(); pr.enable(); f.map(batches); pr.disable()
pstats.Stats(pr).sort_stats('cumulative').print_stats(12)
