"""Multi-GPU host logic: one process per GPU, problems block-partitioned, ONE all-gather of the solved
coefficients (BASELINE.json north_star / configs[4]).  No collective touches the data
path before that: every problem is independent.

`torch.distributed` is only the rendezvous (it carries the 128-byte NCCL unique id and the barriers);
the all-gather itself is `mincob_allgather_device` (NCCL on the handle's stream, straight out of the
buffer the optimize kernel wrote).  The pure-host pieces -- partition, id exchange, rank-major layout --
are what tests/test_sharding_gloo.py runs with world_size 2 on CPU."""
from __future__ import annotations

import numpy as np


def shard_range(total: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous block partition: rank r owns problems [lo, hi).  The all-gather needs equal counts, so
    `total` must divide by `world` (bench.py uses a fixed batch per GPU: weak scaling)."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad world/rank")
    if total % world != 0:
        raise ValueError(f"{total} problems do not split evenly over {world} ranks")
    per = total // world
    return rank * per, (rank + 1) * per


def exchange_unique_id(make_id, rank: int, world: int, dist=None) -> bytes:
    """Rank 0 calls make_id() (mincob_nccl_unique_id) and every rank returns the same 128 bytes."""
    if world == 1:
        return make_id()
    box = [make_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    uid = box[0]
    if not isinstance(uid, (bytes, bytearray)) or len(uid) != 128:
        raise RuntimeError("unique id exchange failed")
    return bytes(uid)


def gathered_view(flat_all, world: int, per_rank: int, N: int, S: int):
    """[world * per_rank * N*3*2S] rank-major buffer -> [world*per_rank][N][3][2S]: with the block partition of
    shard_range this IS the unsharded batch order."""
    return flat_all.reshape(world * per_rank, N, 3, 2 * S)


class ShardedMinco:
    """A MincoBatch per rank plus the communicator.  `dist` is torch.distributed (already initialised)."""

    def __init__(self, params, local_device: int, rank: int, world: int, dist=None, stream: int | None = None):
        from . import api
        self.rank, self.world = rank, world
        self.mb = api.MincoBatch(params, device=local_device)
        if stream:
            self.mb.set_stream(stream)
        if world > 1:
            uid = exchange_unique_id(self.mb.nccl_unique_id, rank, world, dist)
            self.mb.comm_init(world, rank, uid)

    def optimize_and_gather_device(self, x, f, status, iters, evals, coeffs_local, T, coeffs_all, count: int):
        """optimize this rank's shard (device tensors), then all-gather the coefficients (rank-major)."""
        self.mb.optimize_device(x, f, status, iters, evals, coeffs_local, T)
        if self.world > 1:
            self.mb.allgather_device(coeffs_local, coeffs_all, count)

    def close(self):
        self.mb.close()
