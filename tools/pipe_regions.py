"""Per-kernel-group CUDA-event times in the PIPELINED frame loop (binning of frame i+1 beside the blend of frame i)
next to the same groups one frame at a time: shows which groups stretch when they share the SMs."""
import ctypes, os, sys, time
ROOT = os.getcwd()
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "3d-gaussian-splatting-for-novel-view-synthesis_b200"))
import torch, b200gs
from b200gs import _lib
from oracle import gs_oracle as O
lib = _lib.load()
n, W, H = 1_000_000, 1920, 1080
sc = {k: v.cuda() for k, v in O.make_scene(n, seed=0, log_scale=-5.5).items()}
cams = [O.make_camera(W, H, view=v, n_views=16) for v in range(16)]
c2ws = [c["c2w"].cuda() for c in cams]; K = cams[0]
nreg = 32
ms_buf, call_buf = (ctypes.c_float * nreg)(), (ctypes.c_int32 * nreg)()
def collect():
    k = lib.b200gs_profile_collect(ms_buf, call_buf, nreg)
    return {lib.b200gs_profile_region_name(r).decode(): round(1e3 * ms_buf[r] / max(1, call_buf[r]), 1) for r in range(k) if call_buf[r]}
with torch.no_grad():
    sigma = b200gs.build_sigma_from_params(sc["scale_raw"], sc["q_raw"])
    def single(i):
        col = b200gs.evaluate_sh(sc["f_dc"], sc["f_rest"], sc["pos"], c2ws[i % 16])
        b200gs.render(sc["pos"], col, sc["opacity_raw"], sigma, c2ws[i % 16], H, W, K["fx"], K["fy"], K["cx"], K["cy"])
    for i in range(5): single(i)
    torch.cuda.synchronize()
    lib.b200gs_profile_enable(1)
    for i in range(40): single(i)
    torch.cuda.synchronize()
    one = collect()
    lib.b200gs_profile_enable(0)
    pipe = b200gs.RenderPipeline(); pend = []
    def step(i):
        col = b200gs.evaluate_sh(sc["f_dc"], sc["f_rest"], sc["pos"], c2ws[i % 16])
        pend.append(pipe.submit(sc["pos"], col, sc["opacity_raw"], sigma, c2ws[i % 16], H, W, K["fx"], K["fy"], K["cx"], K["cy"]))
        if len(pend) > 1: pipe.result(pend.pop(0))
    for i in range(6): step(i)
    pipe.synchronize(); pend.clear()
    lib.b200gs_profile_enable(1)
    t0 = time.perf_counter()
    for i in range(60): step(i)
    while pend: pipe.result(pend.pop(0))
    pipe.synchronize(); t1 = time.perf_counter()
    two = collect()
    lib.b200gs_profile_enable(0)
    print("pipelined frame", round(1e6 * (t1 - t0) / 60, 1), "us (with the event records)")
    print(f"{'group':18s} {'alone us':>9s} {'pipelined us':>13s}")
    for k in one:
        print(f"{k:18s} {one[k]:9.1f} {two.get(k, float('nan')):13.1f}")
    print(f"{'sum':18s} {sum(one.values()):9.1f} {sum(two.values()):13.1f}")
