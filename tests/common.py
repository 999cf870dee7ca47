"""Helpers shared by the CPU and GPU tests."""
import os

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
PARAMS = ("pos", "scale_raw", "q_raw", "opacity_raw", "f_dc", "f_rest")
GOLDEN_CASES = ("c1_10k_sh0_256", "sh3_4k_200x136_rot", "edge_1500_97x71", "dense_600_48x40")
GRAD_CASES = ("sh3_4k_200x136_rot", "edge_1500_97x71", "dense_600_48x40")


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz")))


def golden_inputs(G, device="cpu", dtype=torch.float32):
    sc = {k: torch.from_numpy(G["in_" + k]).to(device=device, dtype=dtype) for k in PARAMS}
    H, W, fx, fy, cx, cy = G["cam"]
    cam = dict(c2w=torch.from_numpy(G["c2w"]).to(device=device, dtype=dtype), H=int(H), W=int(W), fx=float(fx),
               fy=float(fy), cx=float(cx), cy=float(cy))
    return sc, cam


def canonical_lists(list_tile, list_id, z_of_id):
    """Sort every tile's list by (depth, id): removes the freedom the reference's unstable argsort leaves
    among equal depths (hard part 2)."""
    list_tile = np.asarray(list_tile, np.int64)
    list_id = np.asarray(list_id, np.int64)
    order = np.lexsort((list_id, z_of_id[list_id], list_tile))
    return list_tile[order], list_id[order]


def integer_stages_from_splats(u, v, radius, z, vis, H, W, T=16):
    """numpy restatement of render.py S9-S14 given per-Gaussian (u, v, radius, z): tile rects, counts,
    depth-sorted per-tile lists (ties by id) and dense per-tile ranges."""
    n = u.shape[0]
    r = radius.astype(np.float32)
    umin = np.floor(u - r).astype(np.int64); umax = np.floor(u + r).astype(np.int64)
    vmin = np.floor(v - r).astype(np.int64); vmax = np.floor(v + r).astype(np.int64)
    on = vis & (umax >= 0) & (umin < W) & (vmax >= 0) & (vmin < H)
    umin, umax = np.clip(umin, 0, W - 1), np.clip(umax, 0, W - 1)
    vmin, vmax = np.clip(vmin, 0, H - 1), np.clip(vmax, 0, H - 1)
    rect = np.stack([umin // T, umax // T, vmin // T, vmax // T], 1)
    rect[~on] = 0
    cnt = (rect[:, 1] - rect[:, 0] + 1) * (rect[:, 3] - rect[:, 2] + 1)
    cnt[~on] = 0
    tiles_x = (W + T - 1) // T
    tiles_y = (H + T - 1) // T
    ids = np.repeat(np.arange(n), cnt)
    first = np.cumsum(cnt) - cnt
    local = np.arange(int(cnt.sum())) - first[ids]
    nu = rect[:, 1] - rect[:, 0] + 1
    tile = (rect[ids, 2] + local // nu[ids]) * tiles_x + rect[ids, 0] + local % nu[ids]
    order = np.lexsort((ids, z[ids], tile))
    tile, ids = tile[order], ids[order]
    ranges = np.zeros((tiles_x * tiles_y, 2), np.int64)
    if tile.size:
        starts = np.flatnonzero(np.r_[True, tile[1:] != tile[:-1]])
        ends = np.r_[starts[1:], tile.size]
        ranges[tile[starts], 0] = starts
        ranges[tile[starts], 1] = ends
    return on, rect, cnt, tile, ids, ranges


def image_report(mine, ref32, ref64=None, tol=1e-4):
    mine = np.asarray(mine, np.float64)
    d32 = np.abs(mine - ref32)
    rep = dict(max_vs_ref32=float(d32.max()), n_bad_vs_ref32=int((d32 > tol).sum()))
    if ref64 is not None:
        d64 = np.abs(mine - ref64)
        dref = np.abs(np.asarray(ref32, np.float64) - ref64)
        rep.update(max_vs_ref64=float(d64.max()), n_bad_min=int((np.minimum(d32, d64) > tol).sum()),
                   n_bad_ref32_vs_ref64=int((dref > tol).sum()), max_ref32_vs_ref64=float(dref.max()))
    return rep


def grad_relerr(mine, ref):
    mine, ref = np.asarray(mine, np.float64), np.asarray(ref, np.float64)
    return float(np.abs(mine - ref).max() / max(np.abs(ref).max(), 1e-30))
