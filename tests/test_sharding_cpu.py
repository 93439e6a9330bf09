"""Multi-rank host logic on CPU: contiguous prompt/seed shards (reference rules) and the result gather over a
world_size-2 gloo group."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from consolver_b200 import sharding


def test_sd_shards_cover_everything_and_last_rank_takes_the_remainder():
    for n, w in [(5000, 8), (10, 3), (7, 8), (0, 4), (64, 1)]:
        spans = [sharding.shard_bounds_sd(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
        assert all(e - s == n // w for s, e in spans[:-1])


def test_flux_shards_are_ceil_chunks():
    for n, w in [(1026, 8), (3, 8), (16, 4)]:
        spans = [sharding.shard_bounds_flux(n, r, w) for r in range(w)]
        got = [i for s, e in spans for i in range(s, e)]
        assert got == list(range(n))
        assert max(e - s for s, e in spans) == -(-n // w)


def test_batches_seed_rule():
    out = list(sharding.batches(list(range(70)), 32, seed=43))
    assert [(b, len(it), s) for b, it, s in out] == [(0, 32, 43), (1, 32, 44), (2, 6, 45)]


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_items, q):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    r, w, _ = sharding.init_from_env("gloo")
    s, e = sharding.shard_bounds_sd(n_items, r, w)
    # every rank "samples" its shard: the unit of work is independent, the checksum stands in for the latents
    done, checksum = 0, 0.0
    for b, items, seed in sharding.batches(list(range(s, e)), 4, seed=100):
        g = torch.Generator().manual_seed(seed)
        checksum += float(torch.randn(len(items), generator=g).sum()) + sum(items)
        done += len(items)
    stats = sharding.gather_job_stats(done, 0.5 + rank, checksum)
    dist.barrier()
    q.put((rank, stats))
    dist.destroy_process_group()


def test_two_rank_gloo_job_gathers_counts_and_checksums():
    world, n_items = 2, 11
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_items, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0]["total"] == n_items and res[0]["per_rank"] == [5, 6]
    assert res[0]["max_elapsed_s"] == pytest.approx(1.5)
    assert res[0]["checksum"] == pytest.approx(res[1]["checksum"])
    assert res[0]["world"] == 2
