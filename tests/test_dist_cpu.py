"""CPU: host-side multi-GPU logic with world_size=2 over gloo (views sharding, gradient all-reduce,
tile-row band assembly).  The render itself is faked by the oracle here - the GPU path is covered by the
`-m gpu` tests; this checks the partitioning and the collectives' bookkeeping."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from b200gs.dist import GradBucket, allreduce_gradients, render_tile_row_sharded, shard_tile_rows, shard_views
from b200gs.run import GradientReducer, ShardedRandomSampler


def test_shard_views_round_robin():
    assert shard_views(10, 0, 4) == [0, 4, 8] and shard_views(10, 3, 4) == [3, 7]
    got = sorted(v for r in range(8) for v in shard_views(21, r, 8))
    assert got == list(range(21))


@pytest.mark.parametrize("rows,world", [(135, 8), (68, 8), (53, 4), (3, 8), (1, 1), (16, 5)])
def test_shard_tile_rows_even(rows, world):
    bands = shard_tile_rows(rows, world)
    assert len(bands) == world and bands[0][0] == 0 and bands[-1][1] == rows
    assert all(b[1] == n[0] for b, n in zip(bands, bands[1:]))
    sizes = [e - b for b, e in bands]
    assert max(sizes) - min(sizes) <= 1


def test_shard_tile_rows_weighted():
    w = [1] * 10 + [100] * 2 + [1] * 10
    bands = shard_tile_rows(22, 2, w)
    assert bands[0][0] == 0 and bands[-1][1] == 22 and bands[0][1] == bands[1][0]
    load = [sum(w[b:e]) for b, e in bands]
    assert abs(load[0] - load[1]) <= 100
    assert shard_tile_rows(22, 2, [0] * 22) == shard_tile_rows(22, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # data-parallel gradients: each rank differentiates its own views, all-reduce sums them
        p = torch.nn.Parameter(torch.arange(6.0).reshape(2, 3))
        q = torch.nn.Parameter(torch.ones(4))
        views = shard_views(5, rank, world)
        loss = sum(((v + 1) * p).sum() for v in views)
        loss.backward()                     # q gets no gradient on purpose
        allreduce_gradients([p, q])
        out[f"p{rank}"] = p.grad.clone()
        out[f"q{rank}"] = q.grad.clone()
        # tile-row bands: a fake renderer paints its band with rank+1; the assembled frame has every row set
        H, W, rows = 40, 8, 3

        def fake_render(tile_rows):
            img = torch.zeros(H, W, 3)
            b, e = tile_rows
            img[b * 16:min(e * 16, H)] = float(rank + 1)
            return img
        out[f"img{rank}"] = render_tile_row_sharded(fake_render, rows)
        # a generator argument with average=True (consumed once), and a non-contiguous gradient
        a = torch.nn.Parameter(torch.zeros(3, 4))
        a.grad = torch.full((4, 3), float(rank + 1)).t()
        allreduce_gradients((x for x in [a]), average=True)
        out[f"avg{rank}"] = a.grad.clone()
        # one flat bucket, one collective: same sums as the per-tensor all-reduce; p.grad becomes a view of the bucket
        p2 = torch.nn.Parameter(torch.arange(6.0).reshape(2, 3))
        q2 = torch.nn.Parameter(torch.ones(5))
        bucket = GradBucket([p2, q2])
        for it in range(2):
            bucket.zero_grad()
            sum(((v + 1 + it) * p2).sum() for v in views).backward()
            bucket.allreduce()
            assert p2.grad.data_ptr() == bucket.views[0].data_ptr()
            out[f"bp{rank}_{it}"] = p2.grad.clone()
            out[f"bq{rank}_{it}"] = q2.grad.clone()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n,world", [(6_000_000, 8), (1000, 3), (50, 8), (0, 2), (33, 1)])
def test_route_plan_slices_cover_the_gaussians_once(n, world):
    """Routed tile-row frames: slices are aligned, disjoint, cover [0, n) in rank order (the segment order the stable
    depth sort relies on) and fit their segments; band rows are the contiguous cuts of shard_tile_rows."""
    from b200gs.dist import route_plan, shard_tile_rows
    n_rows = 135
    for weights in (None, [1.0] + [0.0] * (n_rows - 1), [float(i % 7) for i in range(n_rows)]):
        bands = shard_tile_rows(n_rows, world, weights)
        per, slices, band_row = route_plan(n, world, bands, n_rows)
        assert per % 32 == 0 and per * world >= n
        assert slices[0][0] == 0 and slices[-1][1] == n
        for r in range(world):
            lo, hi = slices[r]
            assert lo % 32 == 0 or lo == n
            assert 0 <= hi - lo <= per
            if r:
                assert lo == slices[r - 1][1]
        assert band_row[0] == 0 and band_row[-1] == n_rows and len(band_row) == world + 1
        assert all(band_row[q] <= band_row[q + 1] for q in range(world))
        for q, (b, e) in enumerate(bands):
            assert (band_row[q], band_row[q + 1]) == (b, e) or e <= b
    with pytest.raises(ValueError):
        route_plan(10, 2, [(0, 5)], 10)
    with pytest.raises(ValueError):
        route_plan(10, 2, [(5, 10), (0, 5)], 10)


def test_two_rank_gloo_allreduce_and_bands():
    world, port = 2, _free_port()
    with mp.get_context("spawn").Manager() as mgr:        # no fork() of this multi-threaded process
        out = mgr.dict()
        mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
        out = dict(out)
    expect = torch.full((2, 3), float(sum(v + 1 for v in range(5))))
    for r in range(world):
        assert torch.equal(out[f"p{r}"], expect)
        assert torch.equal(out[f"q{r}"], torch.zeros(4))
        img = out[f"img{r}"]
        assert torch.equal(img[:32], torch.full((32, 8, 3), 1.0))       # rows 0-1 -> rank 0 (2 of 3 tile rows)
        assert torch.equal(img[32:], torch.full((8, 8, 3), 2.0))        # row 2 -> rank 1
        assert torch.equal(out[f"avg{r}"], torch.full((3, 4), 1.5))      # (1 + 2) / 2, generator consumed once
        for it in range(2):
            assert torch.equal(out[f"bp{r}_{it}"], torch.full((2, 3), float(sum(v + 1 + it for v in range(5)))))
            assert torch.equal(out[f"bq{r}_{it}"], torch.zeros(5))


@pytest.mark.parametrize("n,world", [(12, 2), (12, 8), (7, 2), (5, 8), (1, 2)])
def test_dp_launcher_sampler_shards_one_shared_permutation(n, world):
    """`b200gs.run --dp`: p ranks x batch 1 visit, per iteration, the views 1 rank x batch p visits."""
    per_rank = [list(iter(ShardedRandomSampler(n, r, world, seed=3))) for r in range(world)]
    assert len({len(x) for x in per_rank}) == 1 == len({len(ShardedRandomSampler(n, r, world, 3)) for r in range(world)})
    single = list(iter(ShardedRandomSampler(n, 0, 1, seed=3)))
    assert sorted(single) == list(range(n))
    padded = (single * (world + 1))[:len(per_rank[0]) * world]
    for it in range(len(per_rank[0])):
        assert [per_rank[r][it] for r in range(world)] == padded[it * world:(it + 1) * world]
    # a second epoch draws a new permutation, the same on every rank
    s0, s1 = ShardedRandomSampler(n, 0, world, 3), ShardedRandomSampler(n, min(1, world - 1), world, 3)
    e1 = [list(iter(s0)), list(iter(s1))]
    e2 = [list(iter(s0)), list(iter(s1))]
    if n > 3:
        assert e1 != e2
    assert len(e2[0]) == len(e2[1])


def _dp_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)
        p = torch.nn.Parameter(torch.randn(4, 3))
        q = torch.nn.Parameter(torch.randn(5))
        reducer = GradientReducer()
        Adam = reducer.wrap_optimizer(torch.optim.Adam)
        clip = reducer.wrap_clip(torch.nn.utils.clip_grad_norm_)
        opt = Adam([{"params": [p], "lr": 0.1}, {"params": [q], "lr": 0.01}], eps=1e-15)
        for it in range(3):
            opt.zero_grad()
            x = torch.full((4, 3), float(rank + 1 + it))
            ((p * x).sum() + (q * (rank + 2)).sum()).backward()        # this rank's "view"
            clip(p, max_norm=1.0)                                      # train.py:536 - triggers the reduction
            if it == 0:
                out[f"g{rank}"] = (p.grad.clone(), q.grad.clone())
            opt.step()                                                 # must not reduce a second time
        out[f"p{rank}"], out[f"q{rank}"] = p.detach().clone(), q.detach().clone()
        # the script re-creates the optimizer after densification: new parameters are tracked
        p2 = torch.nn.Parameter(torch.ones(2))
        opt2 = Adam([{"params": [p2], "lr": 0.1}], eps=1e-15)
        opt2.zero_grad()
        (p2 * float(rank + 1)).sum().backward()
        opt2.step()
        out[f"g2_{rank}"] = p2.grad.clone()
    finally:
        dist.destroy_process_group()


def test_dp_launcher_reduces_once_per_iteration_and_keeps_replicas_identical():
    world, port = 2, _free_port()
    with mp.get_context("spawn").Manager() as mgr:        # no fork() of this multi-threaded process
        out = mgr.dict()
        mp.spawn(_dp_worker, args=(world, port, out), nprocs=world, join=True)
        out = dict(out)
    # iteration 0: mean over the ranks of x (1 and 2) = 1.5, clipped to norm 1 by the script's own call; q: mean(2, 3)
    gp, gq = out["g0"]
    assert torch.allclose(gp, torch.full((4, 3), 1.5) / (1.5 * (12 ** 0.5) + 1e-6), atol=1e-6)
    assert torch.equal(gq, torch.full((5,), 2.5))
    assert torch.equal(out["g0"][0], out["g1"][0]) and torch.equal(out["g0"][1], out["g1"][1])
    assert torch.equal(out["p0"], out["p1"]) and torch.equal(out["q0"], out["q1"])       # replicas never diverge
    assert torch.equal(out["g2_0"], torch.full((2,), 1.5)) and torch.equal(out["g2_0"], out["g2_1"])
