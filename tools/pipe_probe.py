"""Where does a pipelined frame go?  CPU time per queued frame vs GPU time per frame (headline workload)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "3d-gaussian-splatting-for-novel-view-synthesis_b200"))
import torch  # noqa: E402
import b200gs  # noqa: E402
from b200gs import ops  # noqa: E402
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

    def step(i):
        c2w = c2ws[i % 16]
        col = b200gs.evaluate_sh(sc["f_dc"], sc["f_rest"], sc["pos"], c2w)
        return pipe.render(sc["pos"], col, sc["opacity_raw"], sigma, c2w, H, W, K["fx"], K["fy"], K["cx"], K["cy"])
    for i in range(5):
        step(i)
    pipe.synchronize()
    N = 60
    t0 = time.perf_counter()
    for i in range(N):
        step(i)
    t1 = time.perf_counter()
    pipe.synchronize()
    t2 = time.perf_counter()
    print(f"cpu loop {1e6 * (t1 - t0) / N:.0f} us/frame (includes the wait for each frame's statistics event), "
          f"total {1e6 * (t2 - t0) / N:.0f} us/frame")
    # the same without the per-frame wait: pretend the statistics are known (sync mode off, capacity fixed)
    import cProfile
    import pstats
    pr = cProfile.Profile()
    pr.enable()
    for i in range(N):
        step(i)
    pr.disable()
    pipe.synchronize()
    pstats.Stats(pr).sort_stats("cumulative").print_stats(14)
