"""Runs the BASELINE.json configurations C2-C5 at full size on one GPU (synthetic seeded scenes) and
prints one JSON line per configuration: timings (CUDA events), scene statistics and sanity invariants.
C5's tile-row bands are rendered one after the other on this GPU and must add up to the full frame bit for bit.

    python tools/configs_run.py [c2] [c3] [c4] [c5]
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "3d-gaussian-splatting-for-novel-view-synthesis_b200"))
import torch  # noqa: E402
import b200gs  # noqa: E402
from b200gs import ops  # noqa: E402
from b200gs.dist import shard_tile_rows  # noqa: E402
from oracle import gs_oracle as O  # noqa: E402  (scene generator only)

os.environ.setdefault("B200GS_CAPACITY_MODE", "speculative")
CONFIGS = {
    "c2": dict(n=100_000, W=1920, H=1080, ls=-5.0, what="fwd+bwd"),
    "c3": dict(n=1_000_000, W=1297, H=840, ls=-5.5, what="train"),
    "c4": dict(n=3_000_000, W=1920, H=1080, ls=-5.5, what="orbit"),
    "c5": dict(n=6_000_000, W=3840, H=2160, ls=-6.0, what="bands"),
}


def timed(fn, reps, warm=3):
    for i in range(warm):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        fn(warm + i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def run(name):
    cfg = CONFIGS[name]
    n, W, H = cfg["n"], cfg["W"], cfg["H"]
    sc = {k: v.cuda() for k, v in O.make_scene(n, seed=0, log_scale=cfg["ls"]).items()}
    cams = [O.make_camera(W, H, view=v, n_views=16) for v in range(16)]
    c2ws = [c["c2w"].cuda() for c in cams]
    K = cams[0]
    out = {"config": name, "N": n, "W": W, "H": H, "log_scale": cfg["ls"]}

    def fwd(i, p=sc, **kw):
        c2w = c2ws[i % 16]
        sg = b200gs.build_sigma_from_params(p["scale_raw"], p["q_raw"])
        col = b200gs.evaluate_sh(p["f_dc"], p["f_rest"], p["pos"], c2w)
        return b200gs.render(p["pos"], col, p["opacity_raw"], sg, c2w, H, W, K["fx"], K["fy"], K["cx"], K["cy"], **kw)
    with torch.no_grad():
        g, keep = ops._gaussians(sc["pos"], sc["opacity_raw"], sc["scale_raw"], sc["q_raw"], None, sc["f_dc"], sc["f_rest"], None)
        fr = ops.Frame(g, keep, ops.RenderConfig(H=H, W=W, fx=K["fx"], fy=K["fy"], cx=K["cx"], cy=K["cy"]), c2ws[0], sc["pos"].device)
        img0 = fr.render("sync")
        fr.refresh_stats()
        out.update(V=fr.n_visible, I=fr.n_isect, super_pairs=fr.n_super, image_mean=float(img0.mean()),
                   finite=bool(torch.isfinite(img0).all()), in_range=bool((img0 >= 0).all() and (img0 <= 1).all()))
        out["fwd_ms"] = timed(lambda i: fwd(i), 20)
        out["fwd_fps"] = 1e3 / out["fwd_ms"]
        if cfg["what"] == "orbit":      # 120-frame orbit through the frame pipeline (render_trained.py:333-358)
            pipe = b200gs.RenderPipeline()
            sigma = b200gs.build_sigma_from_params(sc["scale_raw"], sc["q_raw"])
            pend = []

            def orbit(i):
                c2w = c2ws[i % 16]
                col = b200gs.evaluate_sh(sc["f_dc"], sc["f_rest"], sc["pos"], c2w)
                pend.append(pipe.submit(sc["pos"], col, sc["opacity_raw"], sigma, c2w, H, W, K["fx"], K["fy"], K["cx"], K["cy"]))
                if len(pend) > 1:
                    pipe.result(pend.pop(0))
            for i in range(5):
                orbit(i)
            while pend:
                pipe.result(pend.pop(0))
            pipe.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            pipe.wait_event(e0)
            for i in range(120):
                orbit(i)
            while pend:
                pipe.result(pend.pop(0))
            torch.cuda.current_stream().wait_stream(pipe.blend_stream)
            e1.record()
            torch.cuda.synchronize()
            out["orbit_120_frames_ms"] = e0.elapsed_time(e1)
            out["orbit_fps"] = 120e3 / out["orbit_120_frames_ms"]
        if cfg["what"] == "bands":      # 8 tile-row bands (one per rank on an 8-GPU box), here one after the other
            rows = (H + 15) // 16
            bands = shard_tile_rows(rows, 8)
            full = fwd(0)
            acc = torch.zeros_like(full)
            band_ms = []
            for b in bands:
                acc += fwd(0, tile_rows=b)
                band_ms.append(timed(lambda i, b=b: fwd(0, tile_rows=b), 5, warm=1))
            out["bands"] = bands
            out["band_ms"] = [round(x, 3) for x in band_ms]
            out["bands_sum_equals_full_frame"] = bool(torch.equal(acc, full))
            out["slowest_band_ms"] = max(band_ms)
    if cfg["what"] in ("fwd+bwd", "train"):
        leaves = {k: v.clone().requires_grad_(True) for k, v in sc.items()}
        target = torch.rand(H, W, 3, device="cuda")
        lrs = {"pos": 1.6e-6, "opacity_raw": 0.05, "f_dc": 2.5e-3, "f_rest": 1.25e-4, "scale_raw": 5e-3, "q_raw": 1e-3}
        opt = b200gs.FusedAdam([{"params": [leaves[k]], "lr": lrs[k]} for k in leaves], lr=1e-3, eps=1e-15)

        def fb(i):
            for p in leaves.values():
                p.grad = None
            loss, _ = b200gs.compute_loss_tensors(fwd(i, leaves), target)
            loss.backward()

        def it(i):
            fb(i)
            b200gs.clip_grad_norm_(leaves["pos"], 1.0)
            opt.step()
        out["fwd_bwd_loss_ms"] = timed(fb, 10)
        out["grads_finite"] = all(bool(torch.isfinite(p.grad).all()) for p in leaves.values())
        out["full_iteration_ms"] = timed(it, 10)
        out["train_it_per_s"] = 1e3 / out["full_iteration_ms"]
    out["peak_mem_GB"] = torch.cuda.max_memory_allocated() / 1e9
    print(json.dumps(out), flush=True)
    del sc
    torch.cuda.empty_cache()
    torch.cuda.reset_peak_memory_stats()


if __name__ == "__main__":
    for name in (sys.argv[1:] or list(CONFIGS)):
        run(name)
