"""Per-kernel times of one tile-row band of the C5 frame (6M Gaussians, 3840x2160; a central and an edge band of 8, the
first band of 2) vs the whole frame, all on one GPU."""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "3d-gaussian-splatting-for-novel-view-synthesis_b200"))
import torch  # noqa: E402
import b200gs  # noqa: E402
from b200gs.dist import shard_tile_rows  # noqa: E402
from oracle import gs_oracle as O  # noqa: E402

os.environ.setdefault("B200GS_CAPACITY_MODE", "speculative")
n, W, H = 6_000_000, 3840, 2160
lib = b200gs.load_library()
sc = {k: v.cuda() for k, v in O.make_scene(n, seed=0, log_scale=-6.0).items()}
cam = O.make_camera(W, H)
c2w = cam["c2w"].cuda()
bands = shard_tile_rows((H + 15) // 16, 8)


def collect():
    ms, calls = (ctypes.c_float * 32)(), (ctypes.c_int32 * 32)()
    nreg = lib.b200gs_profile_collect(ms, calls, 32)
    return {lib.b200gs_profile_region_name(r).decode(): round(ms[r] / calls[r] * 1e3, 1) for r in range(nreg) if calls[r]}


with torch.no_grad():
    sigma = b200gs.build_sigma_from_params(sc["scale_raw"], sc["q_raw"])

    def frame(rows=None):
        col = b200gs.evaluate_sh(sc["f_dc"], sc["f_rest"], sc["pos"], c2w)
        return b200gs.render(sc["pos"], col, sc["opacity_raw"], sigma, c2w, H, W, cam["fx"], cam["fy"], cam["cx"], cam["cy"],
                             tile_rows=rows)
    bands2 = shard_tile_rows((H + 15) // 16, 2)
    for rows in (None, bands[3], bands[0], bands2[0]):
        for _ in range(3):
            frame(rows)
        torch.cuda.synchronize()
        lib.b200gs_profile_enable(1)
        for _ in range(5):
            frame(rows)
        t = collect()
        lib.b200gs_profile_enable(0)
        print("rows", rows, "sum", round(sum(t.values()), 1), t)
