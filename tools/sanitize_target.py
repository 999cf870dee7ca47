"""Small end-to-end program for compute-sanitizer (memcheck / racecheck / initcheck / synccheck): forward + backward
through both routes, the loss, one Adam step, clip, frame sink, a tile-row band and the frame pipeline, on scenes
small enough for the sanitizer's slowdown."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "3d-gaussian-splatting-for-novel-view-synthesis_b200"))
import torch  # noqa: E402
import b200gs  # noqa: E402
from oracle import gs_oracle as O  # noqa: E402

for n, W, H, ls in ((3000, 97, 71, -3.0), (20000, 200, 136, -3.6)):
    sc = O.make_scene(n, seed=n, log_scale=ls)
    cam = O.make_camera(W, H, view=1, n_views=4)
    c2w = cam["c2w"].cuda()
    target = torch.rand(H, W, 3, device="cuda")
    for fused in (True, False):
        p = {k: v.cuda().requires_grad_(True) for k, v in sc.items()}
        sigma = b200gs.build_sigma_from_params(p["scale_raw"], p["q_raw"])
        color = b200gs.evaluate_sh(p["f_dc"], p["f_rest"], p["pos"], c2w)
        if not fused:
            sigma, color = sigma * 1.0, color * 1.0           # materialise: render consumes sigma / color as given
        img = b200gs.render(p["pos"], color, p["opacity_raw"], sigma, c2w, H, W, cam["fx"], cam["fy"], cam["cx"], cam["cy"])
        loss, d = b200gs.compute_loss(img, target)
        loss.backward()
        b200gs.clip_grad_norm_(p["pos"], 1.0)
        opt = b200gs.FusedAdam([{"params": [t], "lr": 1e-3} for t in p.values()], eps=1e-15)
        opt.step()
        u8 = b200gs.to_uint8(img)
    with torch.no_grad():
        q = {k: v.cuda() for k, v in sc.items()}
        sigma = b200gs.build_sigma_from_params(q["scale_raw"], q["q_raw"])
        color = b200gs.evaluate_sh(q["f_dc"], q["f_rest"], q["pos"], c2w)
        band = b200gs.render(q["pos"], color, q["opacity_raw"], sigma, c2w, H, W, cam["fx"], cam["fy"], cam["cx"], cam["cy"],
                             tile_rows=(1, 3))
        # narrow bands (select + project + compact route, <= 35 % of the rows), fused and unfused, incl. the last ragged row
        rows = (H + 15) // 16
        full = b200gs.render(q["pos"], color, q["opacity_raw"], sigma, c2w, H, W, cam["fx"], cam["fy"], cam["cx"], cam["cy"])
        acc = torch.zeros_like(full)
        for b in range(rows):
            color = b200gs.evaluate_sh(q["f_dc"], q["f_rest"], q["pos"], c2w)
            acc += b200gs.render(q["pos"], color, q["opacity_raw"], sigma, c2w, H, W, cam["fx"], cam["fy"], cam["cx"], cam["cy"],
                                 tile_rows=(b, b + 1))
        assert torch.equal(acc, full), "bands of one tile row do not add up to the full frame"
        one = b200gs.render(q["pos"], color * 1.0, q["opacity_raw"], sigma * 1.0, c2w, H, W, cam["fx"], cam["fy"], cam["cx"],
                            cam["cy"], tile_rows=(rows - 1, rows))
        from b200gs.dist import TileRowRenderer
        tr = TileRowRenderer(H, W, q["pos"].device)
        assert torch.equal(tr.render(q["pos"], color, q["opacity_raw"], sigma, c2w, cam["fx"], cam["fy"], cam["cx"], cam["cy"]), full)
        os.environ["B200GS_CAPACITY_MODE"] = "speculative"
        pipe = b200gs.RenderPipeline()
        ts = []
        for i in range(5):
            color = b200gs.evaluate_sh(q["f_dc"], q["f_rest"], q["pos"], c2w)
            ts.append(pipe.submit(q["pos"], color, q["opacity_raw"], sigma, c2w, H, W, cam["fx"], cam["fy"], cam["cx"], cam["cy"]))
            if len(ts) > 2:
                pipe.result(ts[-3])
        imgs = [pipe.result(t) for t in ts]
        pipe.synchronize()
        os.environ["B200GS_CAPACITY_MODE"] = "sync"
    torch.cuda.synchronize()
    print("ok", n, float(img.mean()), d["total"], float(imgs[-1].mean()))
