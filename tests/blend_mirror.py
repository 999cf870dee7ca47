"""Python mirror (tests only) of the blend forward/backward formulas that csrc/blend.cu implements.

Given per-Gaussian splat values (u, v, conic, opacity, rgb) and the per-tile depth-sorted lists, it
evaluates the image and the nine per-splat gradients (u, v, A11, A12, A22, op, r, g, b) with the
analytic formulas of blend_bwd_kernel, vectorised per tile in float64.  Feeding those through the host
harness' per-Gaussian chain must reproduce the reference's autograd gradients (tests/golden).
"""
from __future__ import annotations

import numpy as np


def blend_and_grads(H, W, tiles_x, uniq_tiles, start, end, list_id, u, v, conic, op, rgb, grad_img=None,
                    chi2=6.25, alpha_max=0.99, alpha_cutoff=1 / 128.0, T=16):
    n = u.shape[0]
    img = np.zeros((H, W, 3), np.float64)
    sg = np.zeros((n, 9), np.float64) if grad_img is not None else None
    u, v, conic, op, rgb = (np.asarray(a, np.float64) for a in (u, v, conic, op, rgb))
    for t, s0, s1 in zip(uniq_tiles.tolist(), start.tolist(), end.tolist()):
        ids = list_id[s0:s1]
        tx, ty = t % tiles_x, t // tiles_x
        x0, y0 = tx * T, ty * T
        x1, y1 = min(x0 + T, W), min(y0 + T, H)
        xs, ys = np.meshgrid(np.arange(x0, x1, dtype=np.float64), np.arange(y0, y1, dtype=np.float64), indexing="xy")
        pu, pv = xs.reshape(-1), ys.reshape(-1)
        du = pu[None, :] - u[ids][:, None]
        dv = pv[None, :] - v[ids][:, None]
        A11, A12, A22 = conic[ids, 0][:, None], conic[ids, 1][:, None], conic[ids, 2][:, None]
        q = A11 * du * du + 2 * A12 * du * dv + A22 * dv * dv
        inside = q <= chi2
        g = np.where(inside, np.exp(-0.5 * np.minimum(q, chi2)), 0.0)
        araw = op[ids][:, None] * g
        a = np.minimum(araw, alpha_max)
        a = np.where(a >= alpha_cutoff, a, 0.0)
        Tr = np.cumprod(1 - a, axis=0)
        Tr = np.concatenate([np.ones((1, a.shape[1])), Tr[:-1]], axis=0)
        alive = (Tr > 5e-5).astype(np.float64)
        w = a * Tr * alive
        col = rgb[ids]                                   # [n,3]
        C = np.einsum("np,nc->pc", w, col)               # [P,3]
        py, px = pv.astype(np.int64), pu.astype(np.int64)
        img[py, px] = np.clip(C, 0, 1)
        if grad_img is None:
            continue
        gp = grad_img[py, px].astype(np.float64) * ((C >= 0) & (C <= 1))          # [P,3]
        # dL/dcolor
        g_col = np.einsum("np,pc->nc", w, gp)
        # dL/dalpha_i = sum_ch g_ch (T_i alive_i c_i - S_i / (1 - a_i)),  S_i = sum_{j>i} w_j c_j
        wc = w[:, :, None] * col[:, None, :]                                    # [n,P,3]
        suffix = np.cumsum(wc[::-1], axis=0)[::-1] - wc                         # exclusive suffix sums
        dalpha = ((Tr * alive)[:, :, None] * col[:, None, :] - suffix / (1 - a)[:, :, None]) * gp[None, :, :]
        dalpha = dalpha.sum(axis=2)
        dalpha = np.where(a > 0, dalpha, 0.0)             # alpha cutoff / chi2 gates
        draw = np.where(araw <= alpha_max, dalpha, 0.0)
        g_op = (draw * g).sum(axis=1)
        dq = -0.5 * araw * draw
        g_u = (-dq * (2 * A11 * du + 2 * A12 * dv)).sum(axis=1)
        g_v = (-dq * (2 * A22 * dv + 2 * A12 * du)).sum(axis=1)
        g_a11 = (dq * du * du).sum(axis=1)
        g_a12 = (dq * 2 * du * dv).sum(axis=1)
        g_a22 = (dq * dv * dv).sum(axis=1)
        np.add.at(sg, ids, np.stack([g_u, g_v, g_a11, g_a12, g_a22, g_op, g_col[:, 0], g_col[:, 1], g_col[:, 2]], 1))
    return img, sg
