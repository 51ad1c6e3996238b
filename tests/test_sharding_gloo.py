"""Multi-rank host logic on CPU (gloo, world_size 2): contiguous replica sharding and the ragged all-gather that stands
in for the reference's `glob("*_mean.csv")` join (parallel_postprocessing.py:8-29)."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from cylinder_b200.replicas import RecordExchange, gather_ragged, parameter_grid, pipeline_slices, shard_range


def test_shard_range_partitions():
    for n in (1, 7, 64, 65536, 4097):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_pipeline_slices_partition_and_taper():
    """step_host's slices: a contiguous partition of the batch in order, whatever the sizes; with enough replicas the first and
    last slices are the short ones (a quarter of a middle slice), so the pipeline's fill and drain are short."""
    for n in (1, 3, 7, 10, 100, 607, 8192, 65536):
        for nch in (1, 2, 4, 5, 16, 64):
            b = pipeline_slices(n, nch)
            assert b[0][0] == 0 and b[-1][1] == n and len(b) == max(1, min(nch, n))
            assert all(b[i][1] == b[i + 1][0] for i in range(len(b) - 1)) and all(hi >= lo for lo, hi in b)
    sizes = [hi - lo for lo, hi in pipeline_slices(8192, 16)]
    mid = max(sizes)
    assert sizes[0] <= mid // 3 and sizes[-1] <= mid // 3 and sizes[1] <= mid // 2 + 1 and sizes[-2] <= mid // 2 + 1
    assert min(sizes[2:-2]) >= mid - 1


def test_parameter_grid_order():
    g = parameter_grid(alpha=[-1, 0], C=[1, 2, 3])
    assert g["alpha"].tolist() == [-1, -1, -1, 0, 0, 0] and g["C"].tolist() == [1, 2, 3, 1, 2, 3]


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n_total = 11
    lo, hi = shard_range(n_total, rank, world)
    width = 5 + rank                                     # records may differ in width (different Zmax per rank)
    rec = torch.zeros(hi - lo, width, dtype=torch.float64)
    rec[:, 0] = torch.arange(lo, hi, dtype=torch.float64)
    rec[:, 1:] = rank + 1
    out = gather_ragged(rec)
    # the repeated exchange with shapes fixed at construction: two rounds through the same buffers
    ex = RecordExchange(hi - lo, 3, 4 + rank, "cpu")
    rounds = []
    for it in range(2):
        r64 = torch.full((hi - lo, 3), float(10 * it + rank), dtype=torch.float64)
        r64[:, 0] = torch.arange(lo, hi, dtype=torch.float64)
        r32 = torch.full((hi - lo, 4 + rank), float(it) + .5, dtype=torch.float32)
        ex.start(r64, r32)
        a, b = ex.finish()
        rounds.append((a.numpy().copy(), b.numpy().copy()))
    q.put((rank, (out.numpy(), rounds)))
    dist.destroy_process_group()


def test_gather_ragged_two_ranks():
    world, port = 2, 29517
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for r in range(world):
        out, rounds = res[r]
        for it, (a, b) in enumerate(rounds):
            assert a.shape == (11, 3) and b.shape == (11, 5)
            assert a[:, 0].tolist() == list(range(11))
            lo0, hi0 = shard_range(11, 0, world)
            assert np.all(a[lo0:hi0, 1:] == 10 * it) and np.all(a[hi0:, 1:] == 10 * it + 1)
            assert np.all(b[lo0:hi0, :4] == it + .5) and np.all(b[hi0:, :5] == it + .5)
        assert out.shape == (11, 6)
        assert out[:, 0].tolist() == list(range(11))            # replica order preserved on every rank
        lo, hi = shard_range(11, 0, world)
        assert np.all(out[lo:hi, 1:5] == 1) and np.all(out[hi:, 1:6] == 2)
