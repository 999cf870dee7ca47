"""Host-side overhead of one render call: tiny scene, so GPU work is negligible."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "3d-gaussian-splatting-for-novel-view-synthesis_b200"))
import torch, b200gs
from oracle import gs_oracle as O
os.environ.setdefault("B200GS_CAPACITY_MODE", sys.argv[1] if len(sys.argv) > 1 else "speculative")
sc = {k: v.cuda() for k, v in O.make_scene(2000, seed=0, log_scale=-3.0).items()}
cam = O.make_camera(64, 64); c2w = cam["c2w"].cuda()
with torch.no_grad():
    sigma = b200gs.build_sigma_from_params(sc["scale_raw"], sc["q_raw"])
    def step():
        col = b200gs.evaluate_sh(sc["f_dc"], sc["f_rest"], sc["pos"], c2w)
        return b200gs.render(sc["pos"], col, sc["opacity_raw"], sigma, c2w, 64, 64, cam["fx"], cam["fy"], cam["cx"], cam["cy"])
    for _ in range(20): step()
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(200): step()
    torch.cuda.synchronize(); dt = (time.perf_counter() - t) / 200
print(f"mode={os.environ['B200GS_CAPACITY_MODE']} per-call {dt*1e6:.1f} us")
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
with torch.no_grad():
    for _ in range(200): step()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
