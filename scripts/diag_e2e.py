"""Diagnostic: where does the pipelined end-to-end step time go?"""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from vkjax_b200 import nets, runtime as rt
from vkjax_b200.elegy import vkModel
ctx = rt.Context.get(int(os.environ.get("LOCAL_RANK", 0)))
B = 256
m = vkModel(nets.ResNet50(), precision='tf32'); m.init(seed=0)
x = ctx.pinned_empty((4 * B, 224, 224, 3), np.float32); x[...] = 0.5
y = m.predict_on_batch(x[:B])
interp = list(m.call_pred_step_jit._jaxpr_interpreters.values())[0]
# raw H2D bandwidth
dev = ctx.alloc(x[:B].nbytes)
ctx.sync(); t0 = time.perf_counter()
for _ in range(10): ctx.upload_async(dev, x[:B].ctypes.data, x[:B].nbytes)
ctx.sync(); dt = (time.perf_counter() - t0) / 10
print('H2D 154MB: %.2f ms  %.1f GB/s' % (dt * 1e3, x[:B].nbytes / dt / 1e9))
for _ in range(5): interp.sequence.launch()
ctx.sync(); t0 = time.perf_counter()
for _ in range(20): interp.sequence.launch()
ctx.sync(); print('graph only: %.2f ms/step' % ((time.perf_counter() - t0) / 20 * 1e3))
batches = [(x[(i % 4) * B:(i % 4 + 1) * B], m.states, False, False) for i in range(20)]
f = m.call_pred_step_jit
for lanes in (1, 2, 3, 4):
    f.map(batches[:4], lanes=lanes)
    t0 = time.perf_counter(); f.map(batches, lanes=lanes); dt = time.perf_counter() - t0
    print('map lanes=%d: %.2f ms/step' % (lanes, dt / 20 * 1e3))
# host-side cost of issuing only (no sync inside): time the python loop by stubbing sync
import cProfile, pstats
pr = cProfile.Profile(); pr.enable(); f.map(batches, lanes=2); pr.disable()
pstats.Stats(pr).sort_stats('cumulative').print_stats(14)
