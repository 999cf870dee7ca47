"""Why is the pipelined e2e frame rate below both the pipelined render rate and the D2H rate?  Variants of the
delivery loop on the headline workload."""
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
dummy = torch.rand(H, W, 3, device="cuda")
dummy2 = torch.empty_like(dummy)
with torch.no_grad():
    sigma = b200gs.build_sigma_from_params(sc["scale_raw"], sc["q_raw"])
    pipe = b200gs.RenderPipeline()
    copy_stream = torch.cuda.Stream()
    ring = int(os.environ.get('RING', '3'))
    LAG = int(os.environ.get('LAG', '1'))
    pins = [torch.empty((H, W, 3), dtype=torch.float32).pin_memory() for _ in range(ring)]
    done = [torch.cuda.Event() for _ in range(ring)]

    spans = []

    def run(mode, N=300):
        pend, cnt = [], [0]
        spans.clear()
        for e in done:
            e.record(copy_stream)

        def deliver(t):
            img = pipe.result(t)
            if mode == "none":
                return
            k = cnt[0] % ring
            cnt[0] += 1
            done[k].synchronize()
            with torch.cuda.stream(copy_stream):
                if mode == "real":
                    copy_stream.wait_event(pipe.done_event(t))
                    img.record_stream(copy_stream)
                    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    a.record(copy_stream)
                    pins[k].copy_(img, non_blocking=True)
                    b.record(copy_stream)
                    spans.append((a, b))
                elif mode == "dummy":                 # same bytes, no dependency on the frame
                    pins[k].copy_(dummy, non_blocking=True)
                elif mode == "d2d":                   # same bytes, copy engine, no PCIe
                    dummy2.copy_(dummy, non_blocking=True)
                elif mode == "h2d":                   # PCIe in the other direction
                    dummy2.copy_(pins[k], non_blocking=True)
                elif mode == "dummy_dep":             # dependency, fixed source
                    copy_stream.wait_event(pipe.done_event(t))
                    pins[k].copy_(dummy, non_blocking=True)
                done[k].record(copy_stream)

        def step(i):
            c2w = c2ws[i % 16]
            col = b200gs.evaluate_sh(sc["f_dc"], sc["f_rest"], sc["pos"], c2w)
            pend.append(pipe.submit(sc["pos"], col, sc["opacity_raw"], sigma, c2w, H, W, K["fx"], K["fy"], K["cx"], K["cy"]))
            if len(pend) > LAG:
                deliver(pend.pop(0))
        for i in range(12):
            step(i)
        while pend:
            deliver(pend.pop(0))
        pipe.synchronize(); copy_stream.synchronize()
        t0 = time.perf_counter()
        for i in range(N):
            step(i)
        while pend:
            deliver(pend.pop(0))
        pipe.synchronize(); copy_stream.synchronize()
        return N / (time.perf_counter() - t0)
    run("real", 60)
    for mode in ("none", "dummy", "real", "none"):
        fps = run(mode)
        extra = ""
        if spans:
            torch.cuda.synchronize()
            d = sorted(a.elapsed_time(b) for a, b in spans[-200:])
            gaps = sorted(spans[i][1].elapsed_time(spans[i + 1][0]) for i in range(len(spans) - 201, len(spans) - 1))
            extra = f" copy median {d[len(d)//2]*1e3:.0f} us max {d[-1]*1e3:.0f} us; idle between copies median {gaps[len(gaps)//2]*1e3:.0f} us"
        print(mode, round(fps), "fps" + extra)
