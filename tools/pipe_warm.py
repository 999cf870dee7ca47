"""How long does the frame pipeline take to reach its steady state?  Frames per second over consecutive blocks of 25."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "3d-gaussian-splatting-for-novel-view-synthesis_b200"))
import torch  # noqa: E402
import b200gs  # noqa: E402
from oracle import gs_oracle as O  # noqa: E402

n, W, H = 1_000_000, 1920, 1080
sc = {k: v.cuda() for k, v in O.make_scene(n, seed=0, log_scale=-5.5).items()}
cams = [O.make_camera(W, H, view=v, n_views=16) for v in range(16)]
c2ws = [c["c2w"].cuda() for c in cams]
K = cams[0]
os.environ["B200GS_CAPACITY_MODE"] = "speculative"
with torch.no_grad():
    sigma = b200gs.build_sigma_from_params(sc["scale_raw"], sc["q_raw"])
    pipe = b200gs.RenderPipeline()
    pend = []

    def step(i):
        c2w = c2ws[i % 16]
        col = b200gs.evaluate_sh(sc["f_dc"], sc["f_rest"], sc["pos"], c2w)
        pend.append(pipe.submit(sc["pos"], col, sc["opacity_raw"], sigma, c2w, H, W, K["fx"], K["fy"], K["cx"], K["cy"]))
        if len(pend) > 1:
            pipe.result(pend.pop(0))
    out = []
    for blk in range(12):
        t0 = time.perf_counter()
        for i in range(25):
            step(blk * 25 + i)
        pipe.blend_stream.synchronize()
        out.append(round(25 / (time.perf_counter() - t0)))
    print("fps per block of 25 frames:", out, "reserved GB", round(torch.cuda.memory_reserved() / 1e9, 2),
          "mallocs", torch.cuda.memory_stats()["num_device_alloc"])
