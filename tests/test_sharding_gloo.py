"""World-size-2 `gloo` test (CPU) of the multi-GPU host logic in allocnet_b200/sharded.py: block partition,
unique-id exchange, and the invariant the design rests on -- problems are independent and seeded by
index, so "each rank solves its block, all-gather rank-major" reproduces the unsharded batch exactly.
The per-shard solve is the CPU oracle here; on the GPU box the same partition/gather code wraps the
CUDA path (bench.py --gpus N)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, total, N, K, outdir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from allocnet_b200 import sharded, synth
        from allocnet_b200.params import default_params
        from oracle.pyoracle import Oracle
        lo, hi = sharded.shard_range(total, world, rank)
        uid = sharded.exchange_unique_id(lambda: bytes(range(128)), rank, world, dist)
        assert uid == bytes(range(128))
        prm = default_params(3, max_iterations=20)
        pb = synth.make_problems(hi - lo, N=N, K=K, S=3, first=lo)          # rank-local generation, no scatter
        res = Oracle().optimize_batch(prm, pb)
        local = torch.from_numpy(res["coeffs"].reshape(-1).copy())
        allc = torch.empty(world * local.numel(), dtype=torch.float64)
        dist.all_gather_into_tensor(allc, local)                            # the ONE collective of the path
        got = sharded.gathered_view(allc.numpy(), world, hi - lo, N, 3)
        np.save(os.path.join(outdir, f"gathered_{rank}.npy"), got)
        f_all = [None] * world
        dist.all_gather_object(f_all, res["f"])
        if rank == 0:
            np.save(os.path.join(outdir, "f_all.npy"), np.concatenate(f_all))
    finally:
        dist.destroy_process_group()


def test_shard_range_partition():
    sys.path.insert(0, ROOT)
    from allocnet_b200 import sharded
    for total, world in ((524288, 8), (65536, 1), (12, 4)):
        spans = [sharded.shard_range(total, world, r) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == total
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        assert len({hi - lo for lo, hi in spans}) == 1
    with pytest.raises(ValueError):
        sharded.shard_range(10, 4, 0)
    with pytest.raises(ValueError):
        sharded.shard_range(8, 2, 2)


def test_two_rank_gather_equals_unsharded(tmp_path, oracle):
    world, total, N, K = 2, 24, 8, 16
    mp.spawn(_worker, args=(world, _free_port(), total, N, K, str(tmp_path)), nprocs=world, join=True)
    from allocnet_b200 import synth
    from allocnet_b200.params import default_params
    ref = oracle.optimize_batch(default_params(3, max_iterations=20), synth.make_problems(total, N=N, K=K, S=3))
    g0 = np.load(tmp_path / "gathered_0.npy"); g1 = np.load(tmp_path / "gathered_1.npy")
    np.testing.assert_array_equal(g0, g1)                    # every rank holds the whole result
    np.testing.assert_array_equal(g0, ref["coeffs"])         # rank-major gather == unsharded order, bit for bit
    np.testing.assert_array_equal(np.load(tmp_path / "f_all.npy"), ref["f"])
