"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: sharding, gradient all-reduce, packed metrics.
The GPU path itself is covered on one device by test_shard_invariance_bit_exact (units are independent)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from hands_b200.distributed import PackedMetrics, allreduce_gradients, shard, shard_bounds


def test_shard_bounds_cover_and_balance():
    for n in (0, 1, 7, 8, 65536, 65537):
        for world in (1, 2, 4, 8):
            spans = [shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_bounds(4, 2, 2)
    t = torch.arange(12).reshape(6, 2)          # 3 samples x 2 crops
    assert shard(t, 1, 2, units_per_row=2).tolist() == [[8, 9], [10, 11]]
    assert shard(t, 0, 2, units_per_row=2).shape[0] == 4


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)
        # HaMeR-light read-out shaped parameters (hamer_light/mano_head.py:38-40): 1024 -> 96 / 10 / 3
        lin = torch.nn.ModuleList([torch.nn.Linear(1024, 96), torch.nn.Linear(1024, 10), torch.nn.Linear(1024, 3)])
        unused = torch.nn.Parameter(torch.zeros(5))             # never receives a gradient (find_unused_parameters)
        full = torch.randn(8, 1024, generator=torch.Generator().manual_seed(1))
        mine = shard(full, rank, world)
        loss = sum(l(mine).pow(2).sum() for l in lin) / full.shape[0]
        loss.backward()
        n_calls = allreduce_gradients(list(lin.parameters()) + [unused], bucket_bytes=256 << 10)
        # single-process reference on the whole batch: mean over ranks of per-shard grads * world == full-batch grad
        ref = torch.nn.ModuleList([torch.nn.Linear(1024, 96), torch.nn.Linear(1024, 10), torch.nn.Linear(1024, 3)])
        ref.load_state_dict(lin.state_dict())
        (sum(l(full).pow(2).sum() for l in ref) / full.shape[0]).backward()
        err = max(float((p.grad * world - q.grad).abs().max()) for p, q in zip(lin.parameters(), ref.parameters()))
        m = PackedMetrics(["loss", "mpjpe", "empty"], "cpu")
        m.add("loss", float(loss) * mine.shape[0], mine.shape[0])
        m.add("mpjpe", float(rank + 1) * 3.0, 3)
        out = m.reduce()
        ret[rank] = dict(err=err, calls=n_calls, unused=float(unused.grad.abs().sum()), metrics=out)
    finally:
        dist.destroy_process_group()


def test_gradient_allreduce_and_packed_metrics_world2():
    world = 2
    port = _free_port()
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
        r0, r1 = ret[0], ret[1]
    assert r0["err"] < 1e-5 and r1["err"] < 1e-5
    assert r0["calls"] >= 2 and r0["calls"] == r1["calls"]     # 393 KB + small ones with a 256 KB bucket cap
    assert r0["unused"] == 0.0
    assert r0["metrics"]["mpjpe"] == pytest.approx(1.5) and r0["metrics"] == pytest.approx(r1["metrics"], nan_ok=True)
    assert r0["metrics"]["empty"] != r0["metrics"]["empty"]     # nan: no samples
