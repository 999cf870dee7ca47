"""adam_step kernel time (CUDA events around opt.step()) right after a real forward + backward of the headline frame."""
import os, sys
ROOT = os.getcwd()
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "3d-gaussian-splatting-for-novel-view-synthesis_b200"))
import torch, b200gs
from oracle import gs_oracle as O
n, W, H = 1_000_000, 1920, 1080
PARAMS = ("pos", "opacity_raw", "f_dc", "f_rest", "scale_raw", "q_raw")
sc = {k: v.cuda() for k, v in O.make_scene(n, seed=0, log_scale=-5.5).items()}
cams = [O.make_camera(W, H, view=v, n_views=16) for v in range(16)]
c2ws = [c["c2w"].cuda() for c in cams]; K = cams[0]
leaves = {k: sc[k].clone().requires_grad_(True) for k in PARAMS}
target = torch.rand(H, W, 3, device="cuda")
opt = b200gs.FusedAdam([{"params": [leaves[k]], "lr": 1e-4} for k in PARAMS], lr=1e-3, eps=1e-15)
big = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")

def fwd_bwd(i):
    opt.zero_grad(set_to_none=True)
    c2w = c2ws[i % 16]
    sg = b200gs.build_sigma_from_params(leaves["scale_raw"], leaves["q_raw"])
    col = b200gs.evaluate_sh(leaves["f_dc"], leaves["f_rest"], leaves["pos"], c2w)
    img = b200gs.render(leaves["pos"], col, leaves["opacity_raw"], sg, c2w, H, W, K["fx"], K["fy"], K["cx"], K["cy"])
    loss, _ = b200gs.compute_loss_tensors(img, target)
    loss.backward()
    b200gs.clip_grad_norm_(leaves["pos"], max_norm=1.0)

def timed(extra, reps=12):
    ms = []
    for i in range(reps):
        fwd_bwd(i); extra()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); opt.step(); b.record()
        torch.cuda.synchronize()
        ms.append(a.elapsed_time(b))
    ms.sort()
    return round(ms[len(ms) // 2] * 1e3, 1)
for i in range(3):
    fwd_bwd(i); opt.step()
print("mode", os.environ.get("B200GS_ADAM_MODE", "0"), "after fwd+bwd:", timed(lambda: None), "us;  after fwd+bwd+sync:",
      timed(torch.cuda.synchronize), "us;  after fwd+bwd+L2 flush:", timed(lambda: big.zero_()), "us")
