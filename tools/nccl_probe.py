"""Minimal NCCL sanity probe (run under torchrun): init, barrier, all_reduce, all_gather."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from synt_isic_b200.dist import init_from_env, sharded_eval

t0 = time.time()
rank, world, local = init_from_env("nccl")
print(f"[{rank}] init {time.time() - t0:.1f}s", flush=True)
x = torch.ones(4, device=f"cuda:{local}") * (rank + 1)
dist.all_reduce(x)
torch.cuda.synchronize()
print(f"[{rank}] all_reduce {x.tolist()} {time.time() - t0:.1f}s", flush=True)
dist.barrier()
print(f"[{rank}] barrier {time.time() - t0:.1f}s", flush=True)
items = torch.arange(10, device=f"cuda:{local}", dtype=torch.float32).view(10, 1, 1, 1)
out = sharded_eval(lambda t: t.view(-1, 1) * 2, items, dist.group.WORLD)
print(f"[{rank}] sharded_eval {out.view(-1).tolist()}", flush=True)
dist.destroy_process_group()
