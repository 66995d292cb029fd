"""world_size-2 gloo tests (CPU) of the N>1 host logic: sharding bounds and the single all-gather."""
import importlib
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_total, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    par = importlib.import_module("neural-photo-editor_b200.parallel")
    g = torch.Generator().manual_seed(0)
    x = torch.rand(n_total, 3, 64, 64, generator=g)               # same full batch on every rank
    calls = []

    def fake_reconstruct(shard):                                  # stands in for the per-GPU CUDA call
        calls.append(shard.shape[0])
        return shard * 2.0 - 1.0

    out = par.sharded_reconstruct(fake_reconstruct, x)
    ok = torch.equal(out, x * 2.0 - 1.0)
    lo, hi = par.shard_bounds(n_total, rank, world)
    q.put((rank, bool(ok), calls == [hi - lo]))
    dist.destroy_process_group()


@pytest.mark.parametrize("n_total", [8, 7])
def test_sharded_reconstruct_world2_gloo(n_total):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_total, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(r[0] for r in res) == [0, 1]
    assert all(r[1] and r[2] for r in res), res


def test_shard_bounds_cover_batch():
    par = importlib.import_module("neural-photo-editor_b200.parallel")
    for n in (0, 1, 7, 256, 4096):
        for world in (1, 2, 3, 8):
            b = [par.shard_bounds(n, r, world) for r in range(world)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        par.shard_bounds(8, 2, 2)
