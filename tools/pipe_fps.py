"""Headline workload: frames/s one frame at a time against the software-pipelined RenderPipeline (B200GS_FRONT_STREAMS, LAG)."""
import os, sys, time
ROOT = os.getcwd()
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "3d-gaussian-splatting-for-novel-view-synthesis_b200"))
import torch, b200gs
from oracle import gs_oracle as O
n, W, H = 1_000_000, 1920, 1080
sc = {k: v.cuda() for k, v in O.make_scene(n, seed=0, log_scale=-5.5).items()}
cams = [O.make_camera(W, H, view=v, n_views=16) for v in range(16)]
c2ws = [c["c2w"].cuda() for c in cams]; K = cams[0]
os.environ["B200GS_CAPACITY_MODE"] = "speculative"
with torch.no_grad():
    sigma = b200gs.build_sigma_from_params(sc["scale_raw"], sc["q_raw"])
    def single(i):
        col = b200gs.evaluate_sh(sc["f_dc"], sc["f_rest"], sc["pos"], c2ws[i % 16])
        b200gs.render(sc["pos"], col, sc["opacity_raw"], sigma, c2ws[i % 16], H, W, K["fx"], K["fy"], K["cx"], K["cy"])
    for i in range(5): single(i)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for i in range(60): single(i)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    pipe = b200gs.RenderPipeline(); pend = []; LAG = int(os.environ.get('LAG', pipe.lag))
    def step(i):
        col = b200gs.evaluate_sh(sc["f_dc"], sc["f_rest"], sc["pos"], c2ws[i % 16])
        pend.append(pipe.submit(sc["pos"], col, sc["opacity_raw"], sigma, c2ws[i % 16], H, W, K["fx"], K["fy"], K["cx"], K["cy"]))
        if len(pend) > LAG: pipe.result(pend.pop(0))
    for i in range(6): step(i)
    pipe.synchronize(); pend.clear()
    res = []
    for rep in range(5):
        t2 = time.perf_counter()
        for i in range(60): step(i)
        while pend: pipe.result(pend.pop(0))
        pipe.synchronize(); t3 = time.perf_counter()
        res.append(round(60 / (t3 - t2)))
    print("fronts", len(pipe.front_streams), "lag", LAG, "single", round(60 / (t1 - t0)), "fps  pipelined", res, "fps",
          "reserved GB", round(torch.cuda.memory_reserved() / 1e9, 2))
