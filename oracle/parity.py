"""Frame-level comparison of a b200gs frame with the oracle's stages.  TEST INFRASTRUCTURE ONLY.

Used by tests/ (the `-m gpu` parity tests at BASELINE.json config sizes) and by bench.py's `cpu_baseline` leg, where
the oracle's full headline frame is compared with the GPU frame of the same view and the result is published as the
`parity` object of the JSON line.  Nothing in the product imports this.

What "equal" means (BASELINE.json north_star; reference lines in oracle/gs_oracle.py):
  * survivor set, V, I, per-Gaussian depth, radius, tile rect (render.py:104-258)          - exact
  * per-tile depth-sorted index lists (render.py:260-303), ties canonicalised by (depth, id) - exact
  * image (render.py:317-410)                                                              - |diff| <= tol (1e-4)
"""
from __future__ import annotations

from typing import Dict

import numpy as np


def canonical_lists(list_tile, list_id, z_of_id):
    """Every tile's list sorted by (depth, id): the reference's argsort leaves ties between equal depths unordered."""
    list_tile = np.asarray(list_tile, np.int64)
    list_id = np.asarray(list_id, np.int64)
    order = np.lexsort((list_id, z_of_id[list_id], list_tile))
    return list_tile[order], list_id[order]


def compare_frame(ex: Dict[str, np.ndarray], n_isect: int, n_visible: int, image, proj, bins, image_ref,
                  tol: float = 1e-4, image_ref64=None) -> dict:
    """`ex` = Frame.export() as numpy arrays; `proj`, `bins` = the oracle's Projected / Binned; images [H,W,3]."""
    n = ex["depth"].shape[0]
    np_ = lambda t: t.detach().numpy()                   # the stages may sit on an autograd graph (backward tests)
    gid = np_(proj.ids)
    vis = ex["tiles_touched"] >= 0
    vis_ref = np.zeros(n, bool)
    vis_ref[gid] = True
    rep = {"N": int(n), "V_ref": int(gid.shape[0]), "V": int(n_visible), "I_ref": int(bins.tile_ids.shape[0]),
           "I": int(n_isect)}
    rep["V_equal"] = bool(rep["V"] == rep["V_ref"])
    rep["I_equal"] = bool(rep["I"] == rep["I_ref"])
    rep["survivors_differ"] = int((vis != vis_ref).sum())
    both = vis & vis_ref
    idx = gid[both[gid]]                               # reference survivors that the GPU also kept, in depth order
    z_ref = np_(proj.z)[both[gid]]
    rep["depth_bit_equal"] = bool(np.array_equal(ex["depth"][idx], z_ref))
    rad_bad = ex["radius"][idx] != np_(proj.radius)[both[gid]]
    rect_bad = (ex["rect"][idx] != np_(proj.rect)[both[gid]]).any(1)
    rep["radius_mismatches"] = int(rad_bad.sum())
    rep["rect_mismatches"] = int(rect_bad.sum())
    rep["max_abs_uv"] = float(max(np.abs(ex["xy"][idx, 0] - np_(proj.u)[both[gid]]).max(initial=0.0),
                                  np.abs(ex["xy"][idx, 1] - np_(proj.v)[both[gid]]).max(initial=0.0)))
    # per-tile lists, canonicalised; Gaussians whose rect differs (if any) are taken out of BOTH sides and counted
    z_of = np.full(n, np.inf, np.float32)
    z_of[gid] = np_(proj.z)
    ref_tile, ref_id = np_(bins.tile_ids), gid[np_(bins.ranks)]
    flipped = np.zeros(n, bool)
    flipped[idx[rect_bad]] = True
    flipped |= vis != vis_ref
    if flipped.any():
        keep_r, keep_m = ~flipped[ref_id], ~flipped[ex["list_id"]]
        ref_tile, ref_id = ref_tile[keep_r], ref_id[keep_r]
        my_tile, my_id = ex["list_tile"][keep_m], ex["list_id"][keep_m]
    else:
        my_tile, my_id = ex["list_tile"], ex["list_id"]
    t_ref, i_ref = canonical_lists(ref_tile, ref_id, z_of)
    t_me, i_me = canonical_lists(my_tile, my_id, z_of)
    rep["lists_equal"] = bool(np.array_equal(t_ref, t_me) and np.array_equal(i_ref, i_me))
    rep["lists_compared_excluding"] = int(flipped.sum())
    d = np.abs(np.asarray(image, np.float64) - np.asarray(image_ref, np.float64))
    rep.update(tol=tol, max_abs=float(d.max()), n_gt_tol=int((d > tol).sum()), n_values=int(d.size),
               mean_abs=float(d.mean()))
    if image_ref64 is not None:
        d64 = np.abs(np.asarray(image, np.float64) - image_ref64)
        dref = np.abs(np.asarray(image_ref, np.float64) - image_ref64)
        rep.update(n_gt_tol_vs_ref64=int((d64 > tol).sum()), n_gt_tol_min=int((np.minimum(d, d64) > tol).sum()),
                   ref32_vs_ref64_n_gt_tol=int((dref > tol).sum()), ref32_vs_ref64_max_abs=float(dref.max()))
    return rep


def grad_relerr(mine, ref) -> float:
    """Max-norm relative error of a gradient tensor (the tolerance north_star states: 1e-3)."""
    mine, ref = np.asarray(mine, np.float64), np.asarray(ref, np.float64)
    return float(np.abs(mine - ref).max() / max(np.abs(ref).max(), 1e-30))
