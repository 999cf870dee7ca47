"""Small fixed program for ncu: a few forward frames and train steps of the headline workload
(1M Gaussians, 1920x1080, SH3).  Usage: python tools/profile_target.py [frames] [train_steps]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "3d-gaussian-splatting-for-novel-view-synthesis_b200"))
import torch  # noqa: E402
import b200gs  # noqa: E402
from oracle import gs_oracle as O  # noqa: E402  (scene generator only)

frames = int(sys.argv[1]) if len(sys.argv) > 1 else 3
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
n, W, H = 1_000_000, 1920, 1080
sc = {k: v.cuda() for k, v in O.make_scene(n, seed=0, log_scale=-5.5).items()}
cams = [O.make_camera(W, H, view=v, n_views=16) for v in range(16)]
with torch.no_grad():
    sigma = b200gs.build_sigma_from_params(sc["scale_raw"], sc["q_raw"])
    for i in range(frames):
        c2w = cams[i]["c2w"].cuda()
        col = b200gs.evaluate_sh(sc["f_dc"], sc["f_rest"], sc["pos"], c2w)
        img = b200gs.render(sc["pos"], col, sc["opacity_raw"], sigma, c2w, H, W, cams[i]["fx"], cams[i]["fy"], cams[i]["cx"],
                            cams[i]["cy"])
# full training iterations (scripts/train.py:463-538): render + L1/SSIM loss + backward + clip + Adam
leaves = {k: v.clone().requires_grad_(True) for k, v in sc.items()}
target = torch.rand(H, W, 3, device="cuda")
lrs = {"pos": 1.6e-6, "opacity_raw": 0.05, "f_dc": 2.5e-3, "f_rest": 1.25e-4, "scale_raw": 5e-3, "q_raw": 1e-3}
opt = b200gs.FusedAdam([{"params": [leaves[k]], "lr": lrs[k]} for k in leaves], lr=1e-3, eps=1e-15)
for i in range(steps):
    c2w = cams[i]["c2w"].cuda()
    opt.zero_grad(set_to_none=True)
    sg = b200gs.build_sigma_from_params(leaves["scale_raw"], leaves["q_raw"])
    col = b200gs.evaluate_sh(leaves["f_dc"], leaves["f_rest"], leaves["pos"], c2w)
    img = b200gs.render(leaves["pos"], col, leaves["opacity_raw"], sg, c2w, H, W, cams[i]["fx"], cams[i]["fy"], cams[i]["cx"],
                        cams[i]["cy"])
    loss, _ = b200gs.compute_loss_tensors(img, target)
    loss.backward()
    b200gs.clip_grad_norm_(leaves["pos"], 1.0)
    opt.step()
torch.cuda.synchronize()
print("ok", float(img.mean()))
