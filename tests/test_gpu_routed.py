"""Routed ("sort-middle") tile-row bands on ONE GPU: the C-ABI calls a p-rank `TileRowRenderer` makes
(b200gs_route_project_slice on every rank's slice, then b200gs_render_project_routed + rasterize on every band), with
the p band workspaces in local memory instead of peer-mapped memory - the kernels cannot tell the difference.

The assembled frame must equal the one-GPU frame bit for bit, for every band layout (even, weighted, empty bands,
more ranks than slices); per band V / I must add up to the full frame's counters where bands do not share entries.
Reference: tiles are independent in render.py:325-399, Gaussians in render.py:104-258.
"""
import copy

import pytest
import torch

pytestmark = pytest.mark.gpu


def _routed_frame(b200gs, args, H, W, bands, dev, split=False):
    """Emulates len(bands) ranks in this process; returns (image, per band (V, I))."""
    from b200gs import _lib, ops
    lib = _lib.load()
    world = len(bands)
    n_rows = (H + 15) // 16
    n = int(args[0].shape[0])
    per = max(32, -(-((n + world - 1) // world) // 32) * 32)
    ws_bytes, _ = ops._sizes(lib, world * per, H, W, 0)
    band_ws = [torch.empty(ws_bytes, dtype=torch.uint8, device=dev) for _ in range(world)]
    image = torch.full((H, W, 3), -1.0, dtype=torch.float32, device=dev)
    cfg, c2w = args[-1], args[-2]
    full = copy.copy(cfg)
    full.tile_row_begin = full.tile_row_end = 0
    routes = []
    for r in range(world):
        route = _lib.Route(world=world, rank=r, seg_capacity=per, band_ws_bytes=ws_bytes,
                           flags=_lib.ROUTE_RECORDS_LATER if split else 0)
        for q in range(world):
            route.band_ws[q] = band_ws[q].data_ptr()
            route.band_row[q] = min(bands[q][0], n_rows)
        route.band_row[world] = n_rows
        routes.append(route)
    keep, slice_ws = [], [None]
    for r in range(world):                                   # source role of every rank
        lo, hi = min(n, r * per), min(n, (r + 1) * per)
        keep.append(ops.route_project_slice(*args[:8], c2w, full, routes[r], lo, hi, slice_ws))
        if split:                                            # ... records in a second pass (a side stream in real life)
            ops.route_records(hi - lo, c2w, full, routes[r], slice_ws)
    counts = []
    for r in range(world):                                   # destination role of every rank
        band = copy.copy(cfg)
        b, e = bands[r]
        band.tile_row_begin, band.tile_row_end = (b, e) if e > b else (n_rows, n_rows)
        band.keep_outside_band = True
        band.out = image
        with torch.cuda.device(dev):
            fr = ops.RoutedFrame(routes[r], band, c2w, dev)
            fr.launch("speculative", [band_ws[r], None])
            fr.finish()
        counts.append((fr.n_visible, fr.n_isect))
    torch.cuda.synchronize()
    return image, counts


@pytest.mark.parametrize("n,H,W,log_scale", [(20000, 200, 320, -3.6), (3000, 97, 71, -3.0), (50, 64, 64, -2.5)])
def test_routed_bands_equal_the_one_gpu_frame(n, H, W, log_scale):
    import b200gs
    from b200gs import api, ops
    from b200gs.dist import shard_tile_rows
    from oracle import gs_oracle as O
    dev = torch.device("cuda", 0)
    sc = {k: v.to(dev) for k, v in O.make_scene(n, seed=11, log_scale=log_scale).items()}
    cam = O.make_camera(W, H, view=1, n_views=6)
    c2w = cam["c2w"].to(dev)
    n_rows = (H + 15) // 16
    with torch.no_grad():
        sigma = b200gs.build_sigma_from_params(sc["scale_raw"], sc["q_raw"])
        col = b200gs.evaluate_sh(sc["f_dc"], sc["f_rest"], sc["pos"], c2w)
        layouts = [shard_tile_rows(n_rows, 2), shard_tile_rows(n_rows, 3), shard_tile_rows(n_rows, 8),
                   shard_tile_rows(n_rows, 4, [1.0] + [0.0] * (n_rows - 1))]          # the last one: three empty bands
        for fused in (True, False):
            s_in, c_in = (sigma, col) if fused else (api._real(sigma).clone(), api._real(col).clone())
            full = b200gs.render(sc["pos"], c_in, sc["opacity_raw"], s_in, c2w, H, W, cam["fx"], cam["fy"], cam["cx"], cam["cy"])
            for bands in layouts:
                args, _ = api._resolve(sc["pos"], c_in, sc["opacity_raw"], s_in, c2w, H, W, cam["fx"], cam["fy"], cam["cx"],
                                       cam["cy"], 0.01, 100.0, 32, 16, 1e-6, 6.25, 0.99, 1 / 128., None)
                assert (args[2] is not None) == fused
                img, counts = _routed_frame(b200gs, args, H, W, bands, dev)
                assert torch.equal(img, full), (fused, bands, float((img - full).abs().max()))
                img2, counts2 = _routed_frame(b200gs, args, H, W, bands, dev, split=True)
                assert torch.equal(img2, full) and counts2 == counts, (fused, bands)
                # every intersection belongs to exactly one band
                frame = ops.Frame(*ops._gaussians(*args[:8]), args[-1], c2w, dev)
                with torch.cuda.device(dev):
                    frame.render("sync")
                assert sum(i for _, i in counts) == frame.n_isect, (counts, frame.n_isect)
                assert sum(v for v, _ in counts) >= frame.n_visible


def test_route_argument_checks():
    from b200gs import _lib, ops
    from oracle import gs_oracle as O
    dev = torch.device("cuda", 0)
    sc = {k: v.to(dev) for k, v in O.make_scene(256, seed=3, log_scale=-3.0).items()}
    cam = O.make_camera(64, 64, view=0, n_views=4)
    cfg = ops.RenderConfig(H=64, W=64, fx=cam["fx"], fy=cam["fy"], cx=cam["cx"], cy=cam["cy"])
    c2w = cam["c2w"].to(dev)
    lib = _lib.load()
    ws_bytes, _ = ops._sizes(lib, 2 * 128, 64, 64, 0)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    route = _lib.Route(world=2, rank=0, seg_capacity=128, band_ws_bytes=ws_bytes)
    route.band_ws[0] = ws.data_ptr()                # band 1 unmapped
    route.band_row[0], route.band_row[1], route.band_row[2] = 0, 2, 4
    args = (sc["pos"], sc["opacity_raw"], sc["scale_raw"], sc["q_raw"], None, sc["f_dc"], sc["f_rest"], None)
    with pytest.raises(_lib.B200GSError):
        ops.route_project_slice(*args, c2w, cfg, route, 0, 128, [None])
    route.band_ws[1] = ws.data_ptr()
    with pytest.raises(_lib.B200GSError):           # slice larger than a segment
        ops.route_project_slice(*args, c2w, cfg, route, 0, 256, [None])
    band = copy.copy(cfg)
    band.tile_row_begin, band.tile_row_end = 0, 2
    with pytest.raises(_lib.B200GSError):           # source role needs the full frame's camera
        ops.route_project_slice(*args, c2w, band, route, 0, 128, [None])
    fr = ops.RoutedFrame(route, band, c2w, dev)
    with pytest.raises(_lib.B200GSError):
        fr.backward(None, None)
