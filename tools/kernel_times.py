"""Per-kernel CUDA-event times of the headline frame (and train step) through the library's own region
profiler.  Usage: python tools/kernel_times.py [frames] [train_steps]"""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "3d-gaussian-splatting-for-novel-view-synthesis_b200"))
import torch  # noqa: E402
import b200gs  # noqa: E402
from oracle import gs_oracle as O  # noqa: E402  (scene generator only)

frames = int(sys.argv[1]) if len(sys.argv) > 1 else 20
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 0
os.environ.setdefault("B200GS_CAPACITY_MODE", "speculative")
n, W, H = 1_000_000, 1920, 1080
lib = b200gs.load_library()
sc = {k: v.cuda() for k, v in O.make_scene(n, seed=0, log_scale=-5.5).items()}
cams = [O.make_camera(W, H, view=v, n_views=16) for v in range(16)]
c2ws = [c["c2w"].cuda() for c in cams]
K = cams[0]


def frame(i, p=sc):
    c2w = c2ws[i % 16]
    sg = b200gs.build_sigma_from_params(p["scale_raw"], p["q_raw"])
    col = b200gs.evaluate_sh(p["f_dc"], p["f_rest"], p["pos"], c2w)
    return b200gs.render(p["pos"], col, p["opacity_raw"], sg, c2w, H, W, K["fx"], K["fy"], K["cx"], K["cy"])


def collect():
    ms, calls = (ctypes.c_float * 32)(), (ctypes.c_int32 * 32)()
    nreg = lib.b200gs_profile_collect(ms, calls, 32)
    return {lib.b200gs_profile_region_name(r).decode(): round(ms[r] / calls[r] * 1e3, 1) for r in range(nreg) if calls[r]}


with torch.no_grad():
    for i in range(3):
        frame(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(frames):
        frame(i)
    e1.record()
    torch.cuda.synchronize()
    print("frame us", round(e0.elapsed_time(e1) / frames * 1e3, 1))
    lib.b200gs_profile_enable(1)
    for i in range(frames):
        frame(i)
    print("fwd", collect())
if steps:
    leaves = {k: v.clone().requires_grad_(True) for k, v in sc.items()}
    w = torch.rand(H, W, 3, device="cuda")
    for i in range(steps):
        for p in leaves.values():
            p.grad = None
        (frame(i, leaves) * w).sum().backward()
    print("train", collect())
lib.b200gs_profile_enable(0)
