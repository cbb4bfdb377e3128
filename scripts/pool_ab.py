"""Micro-benchmark of the NHWC pooling kernel (pool2d_kernel) on the ResNet-50 stem pool ([256,112,112,64] 3x3 s2 SAME max)
and avg-pool shapes, as an A/B over one environment switch (default B2J_POOL_IDX64=0|1: 32- vs 64-bit index math;
`pool_ab.py VAR` compares VAR=0 against VAR=1).  CUDA events on the library's stream around graph replays; the 822 MB
input is far larger than L2, so every replay streams from HBM."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def one():
    import numpy as np
    import vkjax_b200 as vkjax
    from vkjax_b200 import runtime as rt
    from vkjax_b200.frontend import lax, jnp
    ctx = rt.Context.get(0)
    cases = [('max 3x3 s2 SAME [256,112,112,64]', (256, 112, 112, 64),
              lambda x: lax.reduce_window(x, -jnp.inf, lax.max, (1, 3, 3, 1), (1, 2, 2, 1), 'SAME')),
             ('sum 3x3 s2 SAME [256,56,56,256]', (256, 56, 56, 256),
              lambda x: lax.reduce_window(x, 0.0, lax.add, (1, 3, 3, 1), (1, 2, 2, 1), 'SAME')),
             ('sum 2x2 s2 VALID [256,56,56,256]', (256, 56, 56, 256),
              lambda x: lax.reduce_window(x, 0.0, lax.add, (1, 2, 2, 1), (1, 2, 2, 1), 'VALID')),
             ('max 3x3 s1 SAME [256,56,56,64]', (256, 56, 56, 64),
              lambda x: lax.reduce_window(x, -jnp.inf, lax.max, (1, 3, 3, 1), (1, 1, 1, 1), 'SAME'))]
    res = {}
    for name, shape, f in cases:
        x = np.random.default_rng(0).random(shape, np.float32) - 0.5
        dx = vkjax.device_put(x)
        vk = vkjax.wrap(f)
        y = vk(dx)
        seq = list(vk._jaxpr_interpreters.values())[0].sequence
        n = int(os.environ.get('POOL_AB_REPS', '30'))
        for _ in range(5 if n > 1 else 0):
            seq.launch()
        ctx.sync()
        e0, e1 = ctx.event(), ctx.event()
        ctx.record(e0)
        for _ in range(n):
            seq.launch()
        ctx.record(e1)
        ctx.sync()
        ms = ctx.elapsed_ms(e0, e1) / n
        gb = (x.nbytes + np.asarray(y).nbytes) / 1e9
        res[name] = {'ms': ms, 'GBps': gb / ms * 1e3, 'checksum': float(np.asarray(y, np.float64).sum())}
    print(json.dumps(res))


if __name__ == '__main__':
    if len(sys.argv) > 1 and sys.argv[1] == 'one':
        one()
    else:
        out = {}
        var = sys.argv[1] if len(sys.argv) > 1 else 'B2J_POOL_IDX64'
        for mode in ('0', '1'):
            env = dict(os.environ, **{var: mode})
            r = subprocess.run([sys.executable, __file__, 'one'], env=env, capture_output=True, text=True)
            line = [l for l in r.stdout.splitlines() if l.startswith('{')]
            out[mode] = json.loads(line[-1]) if line else {'error': r.stderr[-2000:]}
        for name in out['0']:
            if name == 'error':
                continue
            a, b = out['0'][name], out['1'].get(name, {})
            print(f"{name}: {var}=0 {a['ms']:.4f} ms {a['GBps']:.0f} GB/s | {var}=1 {b.get('ms', float('nan')):.4f} ms "
                  f"{b.get('GBps', float('nan')):.0f} GB/s | same result: {a['checksum'] == b.get('checksum')}")
        print(json.dumps(out))
