#!/usr/bin/env python
"""Multi-GPU check of the host-pointer sharded calls (needs >= 2 GPUs; tests/ run on one):
   torchrun --nproc-per-node 2 --master-addr 127.0.0.1 tools/gpu_sharded_check.py
rank 0 calls mincob_optimize_sharded (reads back every rank's coefficients), the others
mincob_optimize_sharded_local (their own shard); rank 0's gathered array must contain every rank's local result
bit for bit, in rank-major order, and equal the single-rank result of the same problems."""
import os, sys
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from allocnet_b200 import api, sharded, synth
from allocnet_b200.params import default_params

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("gloo")
B, N, K, S = 512, 8, 16, 3
pb = synth.make_problems(B, N=N, K=K, S=S, first=rank * B)
mb = api.MincoBatch(default_params(S), device=local)
mb.comm_init(world, rank, sharded.exchange_unique_id(mb.nccl_unique_id, rank, world, dist))
mb.set_problems(pb)
n, cnt = pb.nvars, B * N * 3 * 2 * S
x = pb.x0().copy(); f = np.zeros(B); T = np.zeros((B, N))
status = np.zeros(B, np.int32); iters = np.zeros(B, np.int32); evals = np.zeros(B, np.int32)
if rank == 0:
    call = np.zeros(world * cnt)
    mb.optimize_sharded_host_buffers(x, f, status, iters, evals, call, T)
    mine = call[:cnt].copy()
else:
    mine = np.zeros(cnt)
    mb.optimize_sharded_local_host_buffers(x, f, status, iters, evals, mine, T)
ptr, count = mb.gathered_device()
assert ptr and count == world * cnt, (ptr, count)
# the same shard through the single-rank path must give the same coefficients (the kernels are deterministic)
mb1 = api.MincoBatch(default_params(S), device=local)
mb1.set_problems(pb)
ref = mb1.optimize(pb.x0())
assert np.array_equal(ref["coeffs"].ravel(), mine), "sharded call differs from mincob_optimize on the same shard"
parts = [torch.zeros(cnt, dtype=torch.float64) for _ in range(world)] if rank == 0 else None
dist.gather(torch.from_numpy(mine), parts, dst=0)
if rank == 0:
    for r in range(world):
        assert np.array_equal(parts[r].numpy(), call[r * cnt:(r + 1) * cnt]), f"rank {r} slice of the gathered array"
    print(f"sharded check ok: {world} ranks x {B} problems, gathered array == rank-major local results, status ok "
          f"{(status >= 0).mean():.3f}")
mb.close(); mb1.close()
dist.destroy_process_group()
