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


def route_plan(n: int, world: int, bands: Sequence[Tuple[int, int]], n_rows: int):
    """Host arithmetic of a routed tile-row frame (csrc/route.cu): returns (per, slices, band_row).

    `per` = slice length = capacity of one (source rank, band) segment: ceil(n / world) rounded up to 32 entries, so
    that every slice starts 16-byte aligned in all six parameter arrays; slices[r] = [begin, end) of the Gaussians rank
    r projects (trailing ranks may get empty slices); band_row[q] = first tile row of band q, band_row[world] = n_rows
    (the bands of `shard_tile_rows` are contiguous, an empty band is [r, r))."""
    if world <= 0 or n < 0:
        raise ValueError("world must be positive and n non-negative")
    if len(bands) != world:
        raise ValueError("need one band per rank")
    per = max(32, -(-((n + world - 1) // world) // 32) * 32)
    slices = [(min(n, r * per), min(n, (r + 1) * per)) for r in range(world)]
    band_row = [min(int(b), n_rows) for b, _e in bands] + [n_rows]
    if any(band_row[q + 1] < band_row[q] for q in range(world)) or band_row[0] != 0:
        raise ValueError("bands must be contiguous, in rank order and start at row 0")
    return per, slices, band_row


def allreduce_gradients(params: Iterable[torch.Tensor], group=None, average: bool = False) -> None:
    """SUM all-reduce of `.grad` of every parameter (in place).  Ranks that have no gradient for a
    parameter contribute zeros, so densify/prune decisions taken from the reduced gradients stay
    identical on every rank (scripts/train.py:544-557).  `params` may be any iterable (a generator is
    consumed once)."""
    params = list(params)
    if not dist.is_available() or not dist.is_initialized():
        return
    world = dist.get_world_size(group)
    if world == 1:
        return
    handles = []
    for p in params:
        if p.grad is None:
            p.grad = torch.zeros_like(p)
        elif not p.grad.is_contiguous():
            p.grad = p.grad.contiguous()
        handles.append(dist.all_reduce(p.grad, op=dist.ReduceOp.SUM, group=group, async_op=True))
    for h in handles:
        h.wait()
    if average:
        for p in params:
            p.grad.div_(world)


class GradBucket:
    """One flat gradient buffer for a parameter set, reduced with ONE collective per iteration.

    Six `all_reduce` calls (f_rest 180 MB + five tensors of <= 16 MB) cost a launch and a ring/tree set-up each; here
    the gradients of `b200gs.render`'s backward are written straight into views of one flat fp32 buffer
    (`ops.register_grad_sink`: autograd adopts the view as `.grad`, no copy), and `allreduce()` sums that buffer over
    the ranks in one call.  Gradients that did not come through a sink (another producer, a second backward into the
    same leaf) are copied into their slot first.  After `allreduce()` every `p.grad` is a view of the reduced buffer.
    Backend-agnostic (NCCL on GPUs; the CPU tests run it over gloo, where the sinks are simply never handed out)."""

    def __init__(self, params: Iterable[torch.Tensor], group=None):
        self.params = list(params)
        if not self.params:
            raise ValueError("GradBucket: no parameters")
        self.group = group
        dev, dt = self.params[0].device, self.params[0].dtype
        offs, total = [], 0
        for p in self.params:
            if p.device != dev or p.dtype != dt:
                raise ValueError("GradBucket: all parameters must share a device and a dtype")
            offs.append(total)
            total += (p.numel() + 31) // 32 * 32          # 128-byte aligned slots
        self.flat = torch.zeros(total, dtype=dt, device=dev)
        self.views = [self.flat[o:o + p.numel()].view(p.shape) for o, p in zip(offs, self.params)]
        if dev.type == "cuda":
            from . import ops
            for p, v in zip(self.params, self.views):
                ops.register_grad_sink(p, v)

    def allreduce(self, average: bool = False) -> None:
        for p, v in zip(self.params, self.views):
            g = p.grad
            if g is None:
                v.zero_()
            elif g.data_ptr() != v.data_ptr():
                v.copy_(g)
            p.grad = v
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group)
            if average:
                self.flat.div_(dist.get_world_size(self.group))
        if self.flat.is_cuda:
            from . import ops
            ops.release_grad_sinks(self.params)

    def zero_grad(self) -> None:
        """`optimizer.zero_grad(set_to_none=True)` for the bucket's parameters (the sinks are handed out again)."""
        for p in self.params:
            p.grad = None


def render_tile_row_sharded(render_fn, n_tile_rows: int, group=None, weights=None) -> torch.Tensor:
    """Every rank renders its band of tile rows with `render_fn(tile_rows=(begin, end))` (pixels outside
    the band are zero) and the bands are gathered into the full frame on every rank.

    Bands are contiguous row ranges of the [H,W,3] image, so the exchange is an all-gather of the band slices
    (each rank sends only its own pixels), not a reduction of whole frames.  Backend-agnostic (gloo in the CPU
    tests).  On an NVLink box `TileRowRenderer` skips the collective altogether."""
    if not dist.is_available() or not dist.is_initialized():
        return render_fn(tile_rows=(0, n_tile_rows))
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    bands = shard_tile_rows(n_tile_rows, world, weights)
    begin, end = bands[rank]
    if end <= begin:            # more ranks than rows: an empty band is (n_rows, n_rows); (0, 0) means "all rows"
        image = render_fn(tile_rows=(n_tile_rows, n_tile_rows))
    else:
        image = render_fn(tile_rows=(begin, end))
    if world == 1:
        return image
    H = image.shape[0]
    rows_of = [(min(b * 16, H), min(e * 16, H)) for b, e in bands]
    pad = max(r1 - r0 for r0, r1 in rows_of)
    if pad == 0:
        return image
    mine = image.new_zeros((pad,) + tuple(image.shape[1:]))
    r0, r1 = rows_of[rank]
    mine[:r1 - r0] = image[r0:r1]
    gathered = image.new_empty((world * pad,) + tuple(image.shape[1:]))
    dist.all_gather_into_tensor(gathered, mine, group=group)
    for q, (q0, q1) in enumerate(rows_of):
        if q != rank and q1 > q0:
            image[q0:q1] = gathered[q * pad:q * pad + (q1 - q0)]
    return image


class TileRowRenderer:
    """A large single frame split into bands of 16-pixel tile rows, one band per GPU of an NVLink box, with NO
    collective on the data path (BASELINE.json configs[4]; tiles are independent in the reference's loop,
    render.py:325-399).

    Every rank holds the whole Gaussian set, culls it to its band BEFORE evaluating colours (the band variant of the
    preprocess kernel), depth-sorts only the band's survivors and blends its tile rows; the blend kernel stores the
    band's pixels straight into the frame buffer of the `root` rank over peer-mapped memory (plain st.global over
    NVLink), so when the closing flag barrier completes the full frame sits in root's buffer:

        tr = b200gs.dist.TileRowRenderer(H, W, device)
        img = tr.render(pos, color, opacity_raw, sigma, c2w, fx, fy, cx, cy)    # complete on root, in stream order

    `weights` (one number per tile row, e.g. intersections per row of an earlier frame - `row_weights()`) balances the
    bands.  One process (no process group) renders the whole frame locally.  Forward only.

    By default (`routed`, world > 1) the per-Gaussian work is divided over the ranks as well: every rank projects its
    1/world slice of the Gaussians and writes the survivors' splat records into the workspaces of the bands they touch
    over peer memory (csrc/route.cu); `routed=False` keeps the Gaussians replicated.

    `render(..., defer_check=True)` queues the frame without waiting for its intersection counters (the only host
    synchronisation of a frame); the check - and the rare re-rasterization of a band whose lists outgrew their buffers -
    then happens at the start of the next `render()` or in `finish()`, before any input of that frame is overwritten;
    `finish()` / the counter `redone` say whether it happened (the frame buffer was incomplete until then)."""

    def __init__(self, H: int, W: int, device, group=None, root: int = 0, weights=None, routed=None):
        from . import _lib, peer
        self.H, self.W, self.device, self.group, self.root = int(H), int(W), torch.device(device), group, int(root)
        self.n_rows = (self.H + 15) // 16
        multi = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
        self.world = dist.get_world_size(group) if multi else 1
        self.rank = dist.get_rank(group) if multi else 0
        self.set_weights(weights)
        # the frame lives in the parameter half of a peer area (the control block carries the barrier flags)
        self.area = peer.PeerArea([self.H * self.W * 3], self.device, group=group, multicast=False)
        self.image = self.area.view(self.area.flat_params, 0, (self.H, self.W, 3))
        off = self.image.data_ptr() - self.area.buf.data_ptr()
        self.root_ptr = int(self.area.c_group.area[self.root]) + off
        self._buffers = [None, None]      # frame / intersection workspaces, reused from frame to frame
        # routed ("sort-middle") bands: the per-Gaussian work is divided over the ranks as well, see _render_routed
        if routed is None:
            routed = _lib.env("B200GS_TILE_ROWS_ROUTED", "1") != "0"
        self.routed = bool(routed) and self.world > 1
        self._route = None                # (b200gs_route, N it was sized for, workspace area, band workspace view)
        # opt-in (B200GS_ROUTE_SPLIT=1): the 48-byte splat records cross NVLink on a second stream, beside the
        # destinations' depth sort, and only the 16 bytes per entry the binning needs stay on the critical path.
        # Measured at 8 ranks: 0.555 against 0.540 ms per frame - a radix-pass CTA holds a whole SM's registers, so
        # the two do not share SMs, and the scheme pays a fourth barrier; off by default.
        self.split_records = self.routed and _lib.env("B200GS_ROUTE_SPLIT", "0") == "1"
        self._side = None
        self._slice_ws = [None]
        self._pending = None              # frame whose counters have not been looked at yet (defer_check)
        self.redone = 0                   # deferred frames that had to be rasterized again (finish())

    def set_weights(self, weights=None):
        self.bands = shard_tile_rows(self.n_rows, self.world, weights)

    def render(self, pos, color, opacity_raw, sigma, c2w, fx, fy, cx, cy, **kw):
        from . import api, ops
        begin, end = self.bands[self.rank]
        if end <= begin:
            begin = end = self.n_rows
        defer = bool(kw.get("defer_check", False))
        self.finish()
        with torch.no_grad():
            args, _ = api._resolve(pos, color, opacity_raw, sigma, c2w, self.H, self.W, fx, fy, cx, cy,
                                   kw.get("near", 0.01), kw.get("far", 100.0), kw.get("pix_guard", 32), kw.get("T", 16),
                                   kw.get("min_conis", 1e-6), kw.get("chi_square_clip", 6.25), kw.get("alpha_max", 0.99),
                                   kw.get("alpha_cutoff", 1 / 128.), (begin, end) if self.world > 1 else None)
            cfg = args[-1]
            cfg.out = self.image
            if self.routed:
                return self._render_routed(args, defer)
            if self.world > 1:
                cfg.keep_outside_band = True
                cfg.out_ptr = self.root_ptr
                self.area.barrier()            # root has consumed the previous frame: its buffer may be overwritten
            image, frame = ops.launch_frame(*args, buffers=self._buffers)
            if not defer:
                frame.finish()
            if self.world > 1:
                self.area.barrier()            # every band has landed in root's buffer
        self.last_frame = frame
        self._pending = frame if defer else None
        return image

    def finish(self) -> bool:
        """Looks at the counters of a frame queued with `defer_check=True`.  Returns True when this rank's band did not
        fit its intersection lists and was rasterized again - from inputs that stay untouched until this rank enters
        the next frame's first barrier; whatever consumed the frame buffer before that saw an incomplete band (the lists
        are sized 1.25 x the largest frame seen so far, so this needs a jump in the view)."""
        frame, self._pending = self._pending, None
        redone = bool(frame.finish()) if frame is not None else False
        self.redone += int(redone)
        return redone

    # -- routed bands ----------------------------------------------------------------------------------------------
    def _ensure_route(self, n: int):
        """Band workspaces in peer-visible memory, sized for `n` Gaussians (collective when it has to allocate)."""
        from . import _lib, ops, peer
        if self._route is not None and self._route[1] >= n:
            return self._route
        self._route = None
        per, _, _ = route_plan(n, self.world, self.bands, self.n_rows)         # slice length: a multiple of 32 entries
        lib = _lib.load()
        ws_bytes, _ = ops._sizes(lib, self.world * per, self.H, self.W, 0)
        area = peer.PeerArea([(ws_bytes + 3) // 4], self.device, group=self.group, multicast=False)
        ws = area.flat_params.view(torch.uint8)[:ws_bytes]
        off = ws.data_ptr() - area.buf.data_ptr()
        route = _lib.Route(world=self.world, rank=self.rank, seg_capacity=per, band_ws_bytes=ws_bytes)
        for q in range(self.world):
            route.band_ws[q] = int(area.c_group.area[q]) + off
        self._route = (route, self.world * per, area, ws, per)
        self._buffers[0] = ws             # the band's frame workspace IS the routed-to workspace
        return self._route

    def _render_routed(self, args, defer=False):
        """Sort-middle: this rank projects Gaussians [rank * N/p, (rank+1) * N/p) with the full frame's camera and
        writes every survivor's splat record into the workspace of each band its tile rect meets (peer stores, index
        order); after a flag barrier it depth-sorts, bins and blends what was routed to its own band.  Same per-tile
        lists as the one-GPU frame, hence the same pixels."""
        import copy
        from . import ops
        pos, cfg, c2w = args[0], args[-1], args[-2]
        n = int(pos.shape[0])
        route, _, area, ws, per = self._ensure_route(n)
        _, slices, band_row = route_plan(n, self.world, self.bands, self.n_rows)
        for q, row in enumerate(band_row):                # shard_tile_rows: contiguous, band q = [cut q, cut q+1)
            route.band_row[q] = row
        full = copy.copy(cfg)
        full.tile_row_begin = full.tile_row_end = 0
        full.out, full.out_ptr, full.keep_outside_band = None, 0, False
        cfg.keep_outside_band = True
        cfg.out_ptr = self.root_ptr
        lo, hi = slices[self.rank]
        # root has consumed the previous frame and every rank has finished its previous band: frame buffer and band
        # workspaces may be overwritten
        self.area.barrier()
        from . import _lib
        route.flags = _lib.ROUTE_RECORDS_LATER if self.split_records else 0
        keep = ops.route_project_slice(*args[:8], c2w, full, route, lo, hi, self._slice_ws)
        hook = None
        if self.split_records:
            if self._side is None:
                self._side = (torch.cuda.Stream(self.device), torch.cuda.Event(), torch.cuda.Event())
            side, ev_meta, ev_rec = self._side
            main = torch.cuda.current_stream(self.device)
            ev_meta.record(main)
            side.wait_event(ev_meta)
            with torch.cuda.stream(side):
                ops.route_records(hi - lo, c2w, full, route, self._slice_ws)
                ev_rec.record(side)

            def hook():                    # between the band's depth sort and its rasterize phase
                main.wait_event(ev_rec)
                self.area.barrier()        # every record of every band has landed
        self.area.barrier()                # every segment's keys / rects of every band have landed
        with torch.cuda.device(self.device):
            frame = ops.RoutedFrame(route, cfg, c2w, self.device)
            frame.keep = keep
            frame.after_project = hook
            image = frame.launch("speculative", self._buffers)
        if not defer:
            frame.finish()
        self.area.barrier()                # every band has landed in root's buffer
        self.last_frame = frame
        self._pending = frame if defer else None
        return image

    def row_weights(self, pos, color, opacity_raw, sigma, c2w, fx, fy, cx, cy, **kw):
        """Intersections per tile row of this view, measured by rendering the whole frame once on this GPU; feed the
        result to `set_weights` (every rank computes the same numbers from the same scene)."""
        from . import api, ops
        with torch.no_grad():
            args, _ = api._resolve(pos, color, opacity_raw, sigma, c2w, self.H, self.W, fx, fy, cx, cy,
                                   kw.get("near", 0.01), kw.get("far", 100.0), kw.get("pix_guard", 32), kw.get("T", 16),
                                   kw.get("min_conis", 1e-6), kw.get("chi_square_clip", 6.25), kw.get("alpha_max", 0.99),
                                   kw.get("alpha_cutoff", 1 / 128.), None)
            g, keep = ops._gaussians(*args[:8])
            with torch.cuda.device(self.device):
                frame = ops.Frame(g, keep, args[-1], args[-2], self.device)
                frame.render("sync")
                ranges = frame.export()["ranges"].to(torch.int64)
        tiles_x = (self.W + 15) // 16
        per_tile = (ranges[:, 1] - ranges[:, 0]).view(self.n_rows, tiles_x)
        return (per_tile.sum(dim=1) + per_tile.sum() // (8 * self.n_rows) + 1).tolist()
