"""Diagnostic (torchrun): host->device bandwidth per rank, alone and concurrently, with and without NUMA-local CPU affinity."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch, torch.distributed as dist
from vkjax_b200 import runtime as rt
rank, world = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1))
local = int(os.environ.get('LOCAL_RANK', 0))
dist.init_process_group('gloo')
aff = None
try:
    import pynvml
    pynvml.nvmlInit()
    h = pynvml.nvmlDeviceGetHandleByIndex(local)
    n = (os.cpu_count() + 63) // 64
    mask = pynvml.nvmlDeviceGetCpuAffinity(h, n)
    aff = [i * 64 + b for i, m in enumerate(mask) for b in range(64) if (m >> b) & 1]
except Exception as e:
    aff = 'err %r' % (e,)
print(f'rank {rank}: cpus {os.cpu_count()}, current affinity {len(os.sched_getaffinity(0))} cpus, nvml ideal {str(aff)[:120]}', flush=True)
def bw(ctx, tag):
    nbytes = 154140672
    host = ctx.pinned_empty((nbytes // 4,), np.float32); host[...] = 1.0
    dev = ctx.alloc(nbytes)
    for mode in ('alone', 'together'):
        for r in range(world if mode == 'alone' else 1):
            dist.barrier()
            if mode == 'together' or r == rank:
                ctx.sync(); t0 = time.perf_counter()
                for _ in range(10): ctx.upload_async(dev, host.ctypes.data, nbytes)
                ctx.sync(); dt = (time.perf_counter() - t0) / 10
                print(f'rank {rank} {tag} {mode}: {nbytes / dt / 1e9:.1f} GB/s', flush=True)
            dist.barrier()
ctx = rt.Context.get(local)
bw(ctx, 'default-affinity')
if isinstance(aff, list) and aff:
    os.sched_setaffinity(0, aff)
    bw(ctx, 'nvml-affinity')
dist.destroy_process_group()
