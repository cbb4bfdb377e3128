"""Multi-GPU plumbing: one process per GPU, batch-sharded inference.

The reference is single-device (kompute_jaxpr_interpreter.py:19-20).  Inference jaxprs of the configs are
per-sample, so the batch shards across ranks with replicated weights and NO data-path collective except
the one the north star asks for: an all-gather of the outputs, recorded into the same CUDA graph
(`JaxprInterpreter(allgather_outputs=True)` -> b2j_seq_record_allgather -> ncclAllGather over NVLink).
torch.distributed is used for rendezvous only (exchange of the 128-byte NCCL unique id, barriers).
"""
import os

import numpy as np


def env():
    return int(os.environ.get('RANK', 0)), int(os.environ.get('LOCAL_RANK', 0)), int(os.environ.get('WORLD_SIZE', 1))


def shard_bounds(n, rank, world):
    """Contiguous, equal shards of the leading dimension (n must divide evenly: one graph per shape)."""
    if n % world:
        raise ValueError(f'batch {n} does not divide over {world} ranks')
    per = n // world
    return rank * per, (rank + 1) * per


def shard_batch(x, rank, world):
    lo, hi = shard_bounds(np.shape(x)[0], rank, world)
    return x[lo:hi]


def exchange_bytes(payload, src=0):
    """Broadcast a bytes object from `src` through the already initialised torch.distributed group."""
    import torch.distributed as dist
    box = [payload if dist.get_rank() == src else None]
    dist.broadcast_object_list(box, src=src)
    return box[0]


def init(ctx=None, backend=None):
    """Initialise torch.distributed from the torchrun environment and, when `ctx` (runtime.Context) is given,
    the library's NCCL communicator for the in-graph all-gather.  Returns (rank, world)."""
    rank, local_rank, world = env()
    if world == 1:
        return rank, world
    import torch
    import torch.distributed as dist
    if backend is None:
        backend = 'nccl' if (ctx is not None and torch.cuda.is_available()) else 'gloo'
    if not dist.is_initialized():
        if backend == 'nccl':
            torch.cuda.set_device(local_rank)
            dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
        else:
            dist.init_process_group('gloo')
    if ctx is not None:
        uid = exchange_bytes(ctx.nccl_unique_id() if rank == 0 else None)
        ctx.comm_init(world, rank, uid)
    return rank, world


def all_gather_host(x):
    """Host-side all-gather along the leading dim through torch.distributed (gloo in the CPU tests): the
    semantics the in-graph NCCL all-gather has to reproduce."""
    import torch
    import torch.distributed as dist
    t = torch.from_numpy(np.ascontiguousarray(x))
    out = [torch.empty_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(out, t)
    return np.concatenate([o.numpy() for o in out], axis=0)
