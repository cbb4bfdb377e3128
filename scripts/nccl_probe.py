"""2-rank probe: latency of a 1 MB all-gather through torch's NCCL vs through libb2jax's communicator."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from vkjax_b200 import runtime as rt, dist as vdist

rank, local, world = vdist.env()
ctx = rt.Context.get(local)
vdist.init(ctx, 'nccl')
n = 256 * 1000
x = torch.full((n,), float(rank), device='cuda')
out = torch.empty((world * n,), device='cuda')
for _ in range(5): dist.all_gather_into_tensor(out, x)
torch.cuda.synchronize(); dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(50): dist.all_gather_into_tensor(out, x)
e1.record(); torch.cuda.synchronize()
t_torch = e0.elapsed_time(e1) / 50
send = ctx.alloc(n * 4); recv = ctx.alloc(n * 4 * world)
for _ in range(5): ctx.allgather(send, recv, n * 4)
ctx.sync(); dist.barrier()
a, b = ctx.event(), ctx.event()
ctx.record(a)
for _ in range(50): ctx.allgather(send, recv, n * 4)
ctx.record(b)
t_mine = ctx.elapsed_ms(a, b) / 50
# with compute in between (a 10 ms spin of memset work) to mimic the bench
big = ctx.alloc(1 << 30)
ctx.sync(); dist.barrier()
ctx.record(a)
for _ in range(20):
    for _ in range(4): ctx.memset(big, 0, 1 << 30)
    ctx.allgather(send, recv, n * 4)
ctx.record(b)
t_mix = ctx.elapsed_ms(a, b) / 20
ctx.record(a)
for _ in range(20):
    for _ in range(4): ctx.memset(big, 0, 1 << 30)
ctx.record(b)
t_base = ctx.elapsed_ms(a, b) / 20
print(f'rank {rank}: torch all_gather {t_torch*1e3:.1f} us, b2j allgather {t_mine*1e3:.1f} us, memset+allgather {t_mix:.3f} ms vs memset only {t_base:.3f} ms', flush=True)
dist.destroy_process_group()
