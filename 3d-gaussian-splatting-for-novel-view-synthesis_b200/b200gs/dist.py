"""Multi-GPU partitioning of the render path on one NVLink/NVSwitch box (one process per GPU).

The reference has no distributed code at all (`num_gpus` is only printed, scripts/train.py:285-291);
the two natural shards of the path are:

  * training views are independent -> data parallel over views; the only exchange is one SUM all-reduce
    of the six per-Gaussian gradient tensors (59 floats per Gaussian) before the optimizer step
    (scripts/train.py:530-538);
  * tiles of one frame are independent -> a large frame is split into bands of 16-pixel tile rows, the
    Gaussian set is replicated, and the bands are gathered at the end.

Orbit frames (render_trained.py) are independent too: frame i -> rank i mod world, no collective.
The host-side logic here is backend-agnostic (NCCL on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

from typing import Iterable, List, Sequence, Tuple

import torch
import torch.distributed as dist

PARAM_ORDER = ("pos", "opacity_raw", "f_dc", "f_rest", "scale_raw", "q_raw")


def shard_views(n_views: int, rank: int, world: int) -> List[int]:
    """Round-robin assignment of view indices to ranks (view v -> rank v mod world)."""
    return list(range(rank, n_views, world))


def shard_tile_rows(n_rows: int, world: int, weights: Sequence[float] | None = None) -> List[Tuple[int, int]]:
    """Contiguous bands [begin, end) of tile rows, one per rank.

    Without weights the rows are split as evenly as possible (the first n_rows % world ranks get one
    extra row).  With per-row weights (e.g. the number of tile intersections in each row from a previous
    frame) the cut points balance the cumulative weight instead."""
    if world <= 0:
        raise ValueError("world must be positive")
    if weights is None:
        base, extra = divmod(n_rows, world)
        out, start = [], 0
        for r in range(world):
            size = base + (1 if r < extra else 0)
            out.append((start, start + size))
            start += size
        return out
    if len(weights) != n_rows:
        raise ValueError("need one weight per tile row")
    total = float(sum(weights))
    if total <= 0:
        return shard_tile_rows(n_rows, world)
    cuts, acc, r = [0], 0.0, 1
    for i, w in enumerate(weights):
        acc += float(w)
        while r < world and acc >= total * r / world:
            cuts.append(i + 1)
            r += 1
    while len(cuts) < world:
        cuts.append(n_rows)
    cuts.append(n_rows)
    return [(cuts[i], max(cuts[i], cuts[i + 1])) for i in range(world)]


def allreduce_gradients(params: Iterable[torch.Tensor], group=None, average: bool = False) -> None:
    """SUM all-reduce of `.grad` of every parameter (in place).  Ranks that have no gradient for a
    parameter contribute zeros, so densify/prune decisions taken from the reduced gradients stay
    identical on every rank (scripts/train.py:544-557)."""
    if not dist.is_available() or not dist.is_initialized():
        return
    world = dist.get_world_size(group)
    if world == 1:
        return
    handles = []
    for p in params:
        if p.grad is None:
            p.grad = torch.zeros_like(p)
        handles.append(dist.all_reduce(p.grad, op=dist.ReduceOp.SUM, group=group, async_op=True))
    for h in handles:
        h.wait()
    if average:
        for p in params:
            p.grad.div_(world)


def render_tile_row_sharded(render_fn, n_tile_rows: int, group=None, weights=None) -> torch.Tensor:
    """Every rank renders its band of tile rows with `render_fn(tile_rows=(begin, end))` (pixels outside
    the band are zero), then the bands are combined with a SUM all-reduce: bands are disjoint, so the sum
    is the full frame on every rank."""
    if not dist.is_available() or not dist.is_initialized():
        return render_fn(tile_rows=(0, n_tile_rows))
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    begin, end = shard_tile_rows(n_tile_rows, world, weights)[rank]
    if end <= begin:            # more ranks than rows: an empty band is (n_rows, n_rows); (0, 0) means "all rows"
        begin = end = n_tile_rows
    image = render_fn(tile_rows=(begin, end))
    if world > 1:
        dist.all_reduce(image, op=dist.ReduceOp.SUM, group=group)
    return image
