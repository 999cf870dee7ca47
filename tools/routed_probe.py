"""Per-phase times of a routed ("sort-middle") tile-row frame, emulated on ONE GPU (band workspaces in local memory):
source role of every rank's slice, destination role of every band, per kernel group (CUDA events inside the library).

    python tools/routed_probe.py [world=8] [n=6000000] [W=3840] [H=2160]

max over ranks of (source role) + max over bands of (destination role) + three flag barriers predicts the p-GPU frame
(peer stores instead of local ones: the routed records are ~60 B per survivor and band, a few MB per rank).
"""
import copy
import ctypes
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "3d-gaussian-splatting-for-novel-view-synthesis_b200")):
    sys.path.insert(0, p)


def main():
    import b200gs
    from b200gs import _lib, api, ops
    from b200gs.dist import TileRowRenderer, shard_tile_rows
    from oracle import gs_oracle as O
    world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 6_000_000
    W = int(sys.argv[3]) if len(sys.argv) > 3 else 3840
    H = int(sys.argv[4]) if len(sys.argv) > 4 else 2160
    dev = torch.device("cuda", 0)
    lib = _lib.load()
    sc = {k: v.to(dev) for k, v in O.make_scene(n, seed=0, log_scale=-6.0).items()}
    cam = O.make_camera(W, H, view=0, n_views=16)
    c2w = cam["c2w"].to(dev)
    n_rows = (H + 15) // 16
    nreg = 32
    ms_buf, call_buf = (ctypes.c_float * nreg)(), (ctypes.c_int32 * nreg)()

    def collect():
        k = lib.b200gs_profile_collect(ms_buf, call_buf, nreg)
        return {lib.b200gs_profile_region_name(r).decode(): round(1e3 * ms_buf[r] / max(1, call_buf[r]), 1)
                for r in range(k) if call_buf[r]}

    with torch.no_grad():
        sigma = b200gs.build_sigma_from_params(sc["scale_raw"], sc["q_raw"])
        col = b200gs.evaluate_sh(sc["f_dc"], sc["f_rest"], sc["pos"], c2w)
        tr = TileRowRenderer(H, W, dev)
        weights = tr.row_weights(sc["pos"], col, sc["opacity_raw"], sigma, c2w, cam["fx"], cam["fy"], cam["cx"], cam["cy"])
        bands = shard_tile_rows(n_rows, world, weights)
        args, _ = api._resolve(sc["pos"], col, sc["opacity_raw"], sigma, c2w, H, W, cam["fx"], cam["fy"], cam["cx"], cam["cy"],
                               0.01, 100.0, 32, 16, 1e-6, 6.25, 0.99, 1 / 128., None)
        cfg = args[-1]
        per = max(32, -(-((n + world - 1) // world) // 32) * 32)
        ws_bytes, _ = ops._sizes(lib, world * per, H, W, 0)
        band_ws = [torch.empty(ws_bytes, dtype=torch.uint8, device=dev) for _ in range(world)]
        image = torch.zeros((H, W, 3), dtype=torch.float32, device=dev)
        routes = []
        for r in range(world):
            route = _lib.Route(world=world, rank=r, seg_capacity=per, band_ws_bytes=ws_bytes)
            for q in range(world):
                route.band_ws[q] = band_ws[q].data_ptr()
                route.band_row[q] = min(bands[q][0], n_rows)
            route.band_row[world] = n_rows
            routes.append(route)
        slice_ws, isect = [None], [[band_ws[r], None] for r in range(world)]
        ev = lambda: torch.cuda.Event(enable_timing=True)
        out = {"world": world, "n": n, "W": W, "H": H, "bands": bands, "src_us": [], "dst_us": [], "V": [], "I": []}
        for rep in range(3):
            prof = rep == 2
            if prof:
                lib.b200gs_profile_enable(1)
            src, dst, host_src, host_dst = [], [], [], []
            for r in range(world):
                lo, hi = min(n, r * per), min(n, (r + 1) * per)
                a, b = ev(), ev()
                a.record()
                t0 = time.perf_counter()
                keep = ops.route_project_slice(*args[:8], c2w, cfg, routes[r], lo, hi, slice_ws)
                host_src.append(round(1e6 * (time.perf_counter() - t0), 1))
                b.record()
                torch.cuda.synchronize()
                src.append(round(1e3 * a.elapsed_time(b), 1))
            if prof:
                out["src_regions_last_rank"] = collect()
            frames = []
            for r in range(world):
                band = copy.copy(cfg)
                band.tile_row_begin, band.tile_row_end = bands[r]
                band.keep_outside_band = True
                band.out = image
                a, b = ev(), ev()
                a.record()
                t0 = time.perf_counter()
                fr = ops.RoutedFrame(routes[r], band, c2w, dev)
                fr.launch("speculative", isect[r])
                host_dst.append(round(1e6 * (time.perf_counter() - t0), 1))
                b.record()
                fr.finish()
                torch.cuda.synchronize()
                dst.append(round(1e3 * a.elapsed_time(b), 1))
                frames.append(fr)
                if prof and r in (0, world // 2):
                    out[f"dst_regions_band{r}"] = collect()
            out["src_us"], out["dst_us"] = src, dst
            out["host_src_us"], out["host_dst_us"] = host_src, host_dst
            out["V"], out["I"] = [f.n_visible for f in frames], [f.n_isect for f in frames]
        lib.b200gs_profile_enable(0)
        full = b200gs.render(sc["pos"], col, sc["opacity_raw"], sigma, c2w, H, W, cam["fx"], cam["fy"], cam["cx"], cam["cy"])
        out["equal_full_frame"] = bool(torch.equal(full, image))
        out["predicted_us"] = max(out["src_us"]) + max(out["dst_us"])
    print(json.dumps(out))


if __name__ == "__main__":
    main()
