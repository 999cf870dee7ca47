"""GPU timeline (CUDA event timestamps) of the pipelined e2e loop: front (project+binning) / blend / D2H copy per frame."""
import os
import sys

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
E = lambda: torch.cuda.Event(enable_timing=True)
with torch.no_grad():
    sigma = b200gs.build_sigma_from_params(sc["scale_raw"], sc["q_raw"])
    pipe = b200gs.RenderPipeline()
    copy_stream = torch.cuda.Stream()
    ring = 4
    pins = [torch.empty((H, W, 3), dtype=torch.float32).pin_memory() for _ in range(ring)]
    done = [torch.cuda.Event() for _ in range(ring)]
    for e in done:
        e.record(copy_stream)
    rows, pend, cnt = [], [], [0]
    base = E()

    def deliver(t, rec):
        img = pipe.result(t)
        k = cnt[0] % ring
        cnt[0] += 1
        done[k].synchronize()
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(pipe.done_event(t))
            img.record_stream(copy_stream)
            rec["c0"] = E(); rec["c0"].record(copy_stream)
            pins[k].copy_(img, non_blocking=True)
            rec["c1"] = E(); rec["c1"].record(copy_stream)
            done[k].record(copy_stream)

    def step(i):
        rec = {}
        c2w = c2ws[i % 16]
        rec["f0"] = E(); rec["f0"].record(pipe.front_stream)
        rec["b0"] = E(); rec["b0"].record(pipe.blend_stream)       # when the blend stream got free for this frame
        col = b200gs.evaluate_sh(sc["f_dc"], sc["f_rest"], sc["pos"], c2w)
        t = pipe.submit(sc["pos"], col, sc["opacity_raw"], sigma, c2w, H, W, K["fx"], K["fy"], K["cx"], K["cy"])
        rec["f1"] = E(); rec["f1"].record(pipe.front_stream)
        rec["b1"] = E(); rec["b1"].record(pipe.blend_stream)
        pend.append((t, rec))
        rows.append(rec)
        if len(pend) > 1:
            deliver(*pend.pop(0))
    for i in range(20):
        step(i)
    base.record(pipe.front_stream)
    rows.clear()
    for i in range(14):
        step(i)
    while pend:
        deliver(*pend.pop(0))
    pipe.synchronize(); copy_stream.synchronize()
    for i, r in enumerate(rows):
        ts = {k: base.elapsed_time(v) * 1e3 for k, v in r.items()}
        print(f"frame {i:2d}: front {ts['f0']:7.0f}-{ts['f1']:7.0f}  blend-stream free {ts['b0']:7.0f} blend end {ts['b1']:7.0f}  "
              f"copy {ts.get('c0', 0):7.0f}-{ts.get('c1', 0):7.0f}")
