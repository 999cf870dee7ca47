"""CPU oracle for the Gaussian-splat render path.  TEST INFRASTRUCTURE ONLY.

This file is a CPU restatement (torch CPU ops, fp32 or fp64) of the reference's
algorithm for the hot path `gaussian_splatting.render.render` and the two per-Gaussian
functions its callers run right before it.  It exists so that tests can look at the
intermediate results the reference keeps local (survivor mask, projected centres, radii,
tile rectangles, per-tile depth-sorted index lists, tile ranges) and so that a CPU
baseline can be timed on a GPU box where the reference checkout does not exist.

Who may import this: `tests/`, `__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline`
/ `--impl reference` legs, as the checker or the thing timed as "CPU baseline" - never
the product path (`b200gs` never imports it and fails loudly without its CUDA library).

Parity pin: the reference ships no tests, golden vectors or fixtures, so the
pin is the reference itself: `oracle/make_golden.py` imports the unmodified reference from
/root/reference in the build container, asserts this restatement reproduces its image
BIT-FOR-BIT (fp32, CPU) and its autograd gradients, and commits the resulting vectors
under `tests/golden/`.  `tests/test_oracle_golden.py` re-checks the oracle against those
vectors everywhere.

Reference lines followed (paths relative to the reference checkout):
  gaussian_splatting/gaussian.py:24-68, 71-127          quaternion -> R, Sigma = R S S^T R^T
  gaussian_splatting/spherical_harmonics.py:50-67,70-166 SH basis, channel-major f_rest, sigmoid
  gaussian_splatting/utils.py:10-34, 37-96, 152-191      w2c transform, frustum test, inv2x2
  gaussian_splatting/render.py:104-410                   S1..S17 of SURVEY.md section 3.1
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import torch

TILE = 16

# --- SH normalisation constants (spherical_harmonics.py:50-67) -------------------------------
SH_C0 = 0.28209479177387814
SH_C1 = 0.4886025119029199
SH_C2 = (1.0925484305920792, 1.0925484305920792, 0.31539156525252005,
         1.0925484305920792, 0.5462742152960396)            # xy, yz, zz, xz, xx-yy
SH_C3 = (0.5900435899266435, 2.890611442640554, 0.4570457994644658, 0.3731763325901154,
         0.4570457994644658, 1.445305721320277, 0.5900435899266435)


# =============================================================================================
# Per-Gaussian functions
# =============================================================================================
def quat_to_rotmat(q: torch.Tensor) -> torch.Tensor:
    """gaussian.py:24-68 - (x, y, z, w) quaternion to a row-major 3x3 rotation."""
    x, y, z, w = q.unbind(dim=-1)
    xx, yy, zz = x * x, y * y, z * z
    xy, xz, yz = x * y, x * z, y * z
    xw, yw, zw = x * w, y * w, z * w
    rows = [1 - 2 * (yy + zz), 2 * (xy - zw), 2 * (xz + yw),
            2 * (xy + zw), 1 - 2 * (xx + zz), 2 * (yz - xw),
            2 * (xz - yw), 2 * (yz + xw), 1 - 2 * (xx + yy)]
    return torch.stack(rows, dim=-1).reshape(q.shape[:-1] + (3, 3))


def build_sigma_from_params(scale_raw: torch.Tensor, q_raw: torch.Tensor) -> torch.Tensor:
    """gaussian.py:71-127 - s = max(exp(scale_raw), 1e-6); q normalised with +1e-9; R S S R^T."""
    s = torch.exp(scale_raw).clamp_min(1e-6)
    qn = q_raw / (q_raw.norm(dim=-1, keepdim=True) + 1e-9)
    R = quat_to_rotmat(qn)
    S = torch.diag_embed(s)
    return R @ S @ S @ R.transpose(1, 2)


def sh_basis(d: torch.Tensor) -> torch.Tensor:
    """spherical_harmonics.py:136-163 - the reference's 16 basis values (its own sign choices)."""
    x, y, z = d[:, 0], d[:, 1], d[:, 2]
    xx, yy, zz = x * x, y * y, z * z
    xy, xz, yz = x * y, x * z, y * z
    Y = [torch.full_like(x, SH_C0),
         -SH_C1 * y, SH_C1 * z, -SH_C1 * x,
         SH_C2[0] * xy, SH_C2[1] * yz, SH_C2[2] * (3 * zz - 1), SH_C2[3] * xz, SH_C2[4] * (xx - yy),
         SH_C3[0] * y * (3 * xx - yy), SH_C3[1] * x * y * z, SH_C3[2] * y * (4 * zz - xx - yy),
         SH_C3[3] * z * (2 * zz - 3 * xx - 3 * yy), SH_C3[4] * x * (4 * zz - xx - yy),
         SH_C3[5] * z * (xx - yy), SH_C3[6] * x * (xx - 3 * yy)]
    return torch.stack(Y, dim=1)


def evaluate_sh(f_dc, f_rest, points, c2w) -> torch.Tensor:
    """spherical_harmonics.py:70-166 - colour = sigmoid(sum_k sh[n,k,c] * Y_k(dir))."""
    n = points.shape[0]
    coeff = torch.empty((n, 16, 3), device=points.device, dtype=points.dtype)
    coeff[:, 0] = f_dc
    for c in range(3):                       # channel-major f_rest (:125-127)
        coeff[:, 1:, c] = f_rest[:, 15 * c:15 * (c + 1)]
    d = points - c2w[:3, 3].unsqueeze(0)
    d = d / (d.norm(dim=-1, keepdim=True) + 1e-8)
    return torch.sigmoid((coeff * sh_basis(d).unsqueeze(2)).sum(dim=1))


def world_to_camera(pc, c2w):
    """utils.py:10-34 - homogeneous 4x4 product, same op order as the reference."""
    w2c = torch.eye(4, device=pc.device, dtype=pc.dtype)
    R = c2w[:3, :3]
    w2c[:3, :3] = R.t()
    w2c[:3, 3] = -R.t() @ c2w[:3, 3]
    hom = torch.concatenate([pc, torch.ones_like(pc[:, :1])], dim=1)
    cam = ((w2c @ hom.t()).t())[:, :3]
    return cam[:, 0], cam[:, 1], cam[:, 2]


def in_frustum(x, y, z, fx, fy, cx, cy, H, W, near, far, guard):
    """utils.py:37-96 - strict inequalities, guard band in pixels, no division."""
    ok = (z > 0) & (z > near) & (z < far)
    fxx, fyy = fx * x, fy * y
    ok = ok & (fxx > z * (-guard - cx)) & (fxx < z * (W + guard - cx))
    ok = ok & (fyy > z * (-guard - cy)) & (fyy < z * (H + guard - cy))
    return ok


def invert_2x2(M, eps=1e-12):
    """utils.py:152-191 - adjugate over clamp(det, min=eps)."""
    a, b, c, d = M[:, 0, 0], M[:, 0, 1], M[:, 1, 0], M[:, 1, 1]
    det = torch.clamp(a * d - b * c, min=eps)
    out = torch.empty_like(M)
    out[:, 0, 0] = d / det
    out[:, 0, 1] = -b / det
    out[:, 1, 0] = -c / det
    out[:, 1, 1] = a / det
    return out


# =============================================================================================
# Stage records
# =============================================================================================
@dataclass
class Projected:
    """Everything S1-S11 + S15 produce.  Row i of every tensor is depth rank i (front first)."""
    ids: torch.Tensor            # [V] int64 index into the ORIGINAL input arrays
    u: torch.Tensor              # [V]
    v: torch.Tensor              # [V]
    z: torch.Tensor              # [V] camera depth, ascending
    opacity: torch.Tensor        # [V] sigmoid().clamp(0, .999)
    color: torch.Tensor          # [V,3]
    cov2d: torch.Tensor          # [V,2,2] after the eigen clamp
    lam_max: torch.Tensor        # [V]
    radius: torch.Tensor         # [V] int64
    conic: torch.Tensor          # [V,2,2] inverse with clamped diagonal
    rect: torch.Tensor           # [V,4] int64 tile rect (tu0, tu1, tv0, tv1) inclusive
    tiles_touched: torch.Tensor  # [V] int64
    H: int = 0
    W: int = 0
    stage_counts: Dict[str, int] = field(default_factory=dict)


@dataclass
class Binned:
    """S12-S14: per-tile depth-sorted lists."""
    tile_ids: torch.Tensor       # [I] int64, non-decreasing
    ranks: torch.Tensor          # [I] int64 index into Projected rows
    uniq_tiles: torch.Tensor     # [n_nonempty]
    start: torch.Tensor          # [n_nonempty]
    end: torch.Tensor            # [n_nonempty]
    tiles_x: int = 0
    tiles_y: int = 0

    def per_tile_ids(self, proj: Projected) -> Dict[int, torch.Tensor]:
        return {int(t): proj.ids[self.ranks[s:e]]
                for t, s, e in zip(self.uniq_tiles.tolist(), self.start.tolist(), self.end.tolist())}


class NothingVisible(Exception):
    """Raised internally when S1/S3/S7 leave no Gaussian (reference returns a zero image)."""


# =============================================================================================
# S1-S11, S15  (render.py:104-258, 305-315)
# =============================================================================================
def project(pos, color, opacity_raw, sigma, c2w, H, W, fx, fy, cx, cy,
            near=0.01, far=100.0, pix_guard=32, T=TILE, min_conis=1e-6,
            alpha_cutoff=1 / 128.) -> Projected:
    H, W = int(H), int(W)
    counts = {"N": pos.shape[0]}
    # S1 opacity pre-cull (:106-117)
    pre = torch.sigmoid(opacity_raw).clamp(0, 0.999)
    keep0 = pre >= alpha_cutoff * 0.5
    if not keep0.any():
        raise NothingVisible("opacity")
    orig = torch.nonzero(keep0, as_tuple=False).squeeze(1)
    pos, color, opacity_raw, sigma = pos[keep0], color[keep0], opacity_raw[keep0], sigma[keep0]
    counts["after_opacity"] = pos.shape[0]
    # S2/S3 (:122-136)
    x, y, z = world_to_camera(pos, c2w)
    vis = in_frustum(x, y, z, fx, fy, cx, cy, H, W, near, far, pix_guard)
    pos, color, opacity_raw, sigma = pos[vis], color[vis], opacity_raw[vis], sigma[vis]
    x, y, z, orig = x[vis], y[vis], z[vis], orig[vis]
    counts["after_frustum"] = pos.shape[0]
    if pos.shape[0] == 0:
        raise NothingVisible("frustum")
    # S4 (:146-148)
    uv = torch.stack([fx * x / z + cx, fy * y / z + cy], dim=-1)
    opacity = torch.sigmoid(opacity_raw).clamp(0, 0.999)
    # S5 (:156-175)
    Rwc = c2w[:3, :3].t()
    cam_cov = Rwc.unsqueeze(0) @ sigma @ Rwc.t().unsqueeze(0)
    invz = 1 / z.clamp_min(1e-6)
    invz2 = invz * invz
    J = torch.zeros((pos.shape[0], 2, 3), device=pos.device, dtype=pos.dtype)
    J[:, 0, 0] = fx * invz
    J[:, 1, 1] = fy * invz
    J[:, 0, 2] = -fx * x * invz2
    J[:, 1, 2] = -fy * y * invz2
    cov2d = J @ cam_cov @ J.transpose(1, 2)
    cov2d = 0.5 * (cov2d + cov2d.transpose(1, 2))
    # S6 (:177-179)
    lam, vec = torch.linalg.eigh(cov2d)
    lam = torch.clamp(lam, min=1e-6, max=1e4)
    cov2d = vec @ torch.diag_embed(lam) @ vec.transpose(1, 2)
    # S7 (:187-201)
    fin = torch.isfinite(cov2d.reshape(cov2d.shape[0], -1)).all(dim=-1)
    if not fin.any():
        raise NothingVisible("finite")
    uv, color, opacity, z, cov2d, orig, lam = uv[fin], color[fin], opacity[fin], z[fin], cov2d[fin], orig[fin], lam[fin]
    counts["after_finite"] = uv.shape[0]
    # S8 (:211-219)
    order = torch.argsort(z, descending=False)
    uv, color, opacity, cov2d, lam, orig, z = uv[order], color[order], opacity[order], cov2d[order], lam[order], orig[order], z[order]
    u, v = uv[:, 0], uv[:, 1]
    # S9 (:227-233)
    lam_max = lam[:, 1].clamp_min(1e-12).clamp_max(1e4)
    radius = torch.ceil(2.5 * torch.sqrt(lam_max)).to(torch.int64)
    umin = torch.floor(u - radius).to(torch.int64)
    umax = torch.floor(u + radius).to(torch.int64)
    vmin = torch.floor(v - radius).to(torch.int64)
    vmax = torch.floor(v + radius).to(torch.int64)
    # S10 (:234-247)
    on = (umax >= 0) & (umin < W) & (vmax >= 0) & (vmin < H)
    if not on.any():
        raise Exception("All projected points are off-screen")
    u, v, z, color, opacity, cov2d, orig = u[on], v[on], z[on], color[on], opacity[on], cov2d[on], orig[on]
    lam_max, radius = lam_max[on], radius[on]
    umin, umax = umin[on].clamp(0, W - 1), umax[on].clamp(0, W - 1)
    vmin, vmax = vmin[on].clamp(0, H - 1), vmax[on].clamp(0, H - 1)
    counts["visible"] = u.shape[0]
    # S11 (:251-258)
    rect = torch.stack([umin // T, umax // T, vmin // T, vmax // T], dim=1)
    touched = (rect[:, 1] - rect[:, 0] + 1) * (rect[:, 3] - rect[:, 2] + 1)
    counts["intersections"] = int(touched.sum())
    # S15 (:307-315)
    conic = invert_2x2(cov2d)
    d00 = torch.clamp(conic[:, 0, 0], min=min_conis)
    d11 = torch.clamp(conic[:, 1, 1], min=min_conis)
    conic = conic.clone()
    conic[:, 0, 0] = d00
    conic[:, 1, 1] = d11
    return Projected(ids=orig, u=u, v=v, z=z, opacity=opacity, color=color, cov2d=cov2d,
                     lam_max=lam_max, radius=radius, conic=conic, rect=rect,
                     tiles_touched=touched, H=H, W=W, stage_counts=counts)


# =============================================================================================
# S12-S14 (render.py:260-303): expansion, (tile, depth-rank) sort, ranges
# =============================================================================================
def bin_tiles(proj: Projected, T=TILE) -> Binned:
    """Same result as the reference's dense-mask expansion + composite-key sort, built without
    the [V,max_u,max_v] mask: the composite keys tile*(V+1)+rank are unique, so any enumeration
    order of the pairs sorts to the same sequence."""
    V = proj.u.shape[0]
    tiles_x = (proj.W + T - 1) // T
    tiles_y = (proj.H + T - 1) // T
    tu0, tu1, tv0, tv1 = proj.rect.unbind(dim=1)
    nu = tu1 - tu0 + 1
    cnt = proj.tiles_touched
    rank = torch.repeat_interleave(torch.arange(V, dtype=torch.int64), cnt)
    first = torch.cumsum(cnt, 0) - cnt
    local = torch.arange(int(cnt.sum()), dtype=torch.int64) - first[rank]
    tile = (tv0[rank] + local // nu[rank]) * tiles_x + (tu0[rank] + local % nu[rank])
    M = V + 1
    comp, _ = torch.sort(tile * M + rank)
    tile_sorted = torch.div(comp, M, rounding_mode='floor')
    rank_sorted = comp - tile_sorted * M
    uniq, per_tile = torch.unique_consecutive(tile_sorted, return_counts=True)
    start = torch.zeros_like(uniq)
    start[1:] = torch.cumsum(per_tile[:-1], dim=0)
    return Binned(tile_ids=tile_sorted, ranks=rank_sorted, uniq_tiles=uniq, start=start,
                  end=start + per_tile, tiles_x=tiles_x, tiles_y=tiles_y)


# =============================================================================================
# S16-S17 (render.py:317-410): per-tile front-to-back blend
# =============================================================================================
def blend_tile(u, v, color, opacity, conic, px_u, px_v,
               chi_square_clip=6.25, alpha_max=0.99, alpha_cutoff=1 / 128., want_aux=False):
    """One tile: [n] Gaussians (front first) x [P] pixels (render.py:351-395)."""
    du = px_u.unsqueeze(0) - u.unsqueeze(-1)
    dv = px_v.unsqueeze(0) - v.unsqueeze(-1)
    A11 = conic[:, 0, 0].unsqueeze(-1)
    A12 = conic[:, 0, 1].unsqueeze(-1)
    A22 = conic[:, 1, 1].unsqueeze(-1)
    q = A11 * du * du + 2 * A12 * du * dv + A22 * dv * dv
    inside = q <= chi_square_clip
    g = torch.exp(-0.5 * torch.clamp(q, max=chi_square_clip))
    g = torch.where(inside, g, torch.zeros_like(g))
    alpha = (opacity.unsqueeze(-1) * g).clamp_max(alpha_max)
    alpha = torch.where(alpha >= alpha_cutoff, alpha, torch.zeros_like(alpha))
    trans = torch.cumprod(1 - alpha, dim=0)
    trans = torch.concatenate([torch.ones((1, alpha.shape[-1]), device=u.device, dtype=u.dtype),
                               trans[:-1]], dim=0)
    alive = (trans > 5e-5).to(u.dtype)
    w = alpha * trans * alive
    out = (w.unsqueeze(-1) * color.unsqueeze(1)).sum(dim=0)
    if want_aux:
        return out, {"q": q, "alpha": alpha, "T": trans, "alive": alive}
    return out


def blend(proj: Projected, bins: Binned, T=TILE, chi_square_clip=6.25, alpha_max=0.99,
          alpha_cutoff=1 / 128., tile_stride: int = 1) -> torch.Tensor:
    """`tile_stride` > 1 blends only every tile_stride-th non-empty tile (bounded CPU-baseline samples in
    bench.py); 1 is the reference's behaviour."""
    H, W = proj.H, proj.W
    dt, dev = proj.u.dtype, proj.u.device
    image = torch.zeros((H * W, 3), device=dev, dtype=dt)
    parts, where = [], []
    work = list(zip(bins.uniq_tiles.tolist(), bins.start.tolist(), bins.end.tolist()))[::max(1, int(tile_stride))]
    for tile, s0, s1 in work:
        sel = bins.ranks[s0:s1]
        tx, ty = tile % bins.tiles_x, tile // bins.tiles_x
        x0, y0 = tx * T, ty * T
        x1, y1 = min(x0 + T, W), min(y0 + T, H)
        if x0 >= x1 or y0 >= y1:
            continue
        gu, gv = torch.meshgrid(torch.arange(x0, x1, device=dev, dtype=dt),
                                torch.arange(y0, y1, device=dev, dtype=dt), indexing='xy')
        px_u, px_v = gu.reshape(-1), gv.reshape(-1)
        parts.append(blend_tile(proj.u[sel], proj.v[sel], proj.color[sel], proj.opacity[sel],
                                proj.conic[sel], px_u, px_v, chi_square_clip, alpha_max, alpha_cutoff))
        where.append((px_v * W + px_u).to(torch.int64))
    if parts:
        image = image.scatter_add(0, torch.cat(where).unsqueeze(-1).expand(-1, 3), torch.cat(parts))
    return image.reshape(H, W, 3).clamp(0, 1)


# =============================================================================================
# Whole path
# =============================================================================================
def render(pos, color, opacity_raw, sigma, c2w, H, W, fx, fy, cx, cy,
           near=0.01, far=100.0, pix_guard=32, T=TILE, min_conis=1e-6,
           chi_square_clip=6.25, alpha_max=0.99, alpha_cutoff=1 / 128., return_stages=False):
    """Oracle for render.py:62-410 (same signature; `return_stages` also hands back the
    Projected / Binned records)."""
    H, W = int(H), int(W)
    try:
        proj = project(pos, color, opacity_raw, sigma, c2w, H, W, fx, fy, cx, cy,
                       near, far, pix_guard, T, min_conis, alpha_cutoff)
    except NothingVisible:
        img = (color.sum() * 0.0).expand(H * W * 3).reshape(H, W, 3)   # graph-connected zeros
        return (img, None, None) if return_stages else img
    bins = bin_tiles(proj, T)
    img = blend(proj, bins, T, chi_square_clip, alpha_max, alpha_cutoff)
    return (img, proj, bins) if return_stages else img


def render_from_params(pos, scale_raw, q_raw, opacity_raw, f_dc, f_rest, c2w, H, W, fx, fy, cx, cy,
                       **kw):
    """The triple every reference script calls (scripts/train.py:463,502,505-508)."""
    sigma = build_sigma_from_params(scale_raw, q_raw)
    color = evaluate_sh(f_dc, f_rest, pos, c2w)
    return render(pos, color, opacity_raw, sigma, c2w, H, W, fx, fy, cx, cy, **kw)


# =============================================================================================
# Seeded synthetic scenes - shared by tests and bench for the CPU legs
# =============================================================================================
def make_scene(n: int, seed: int = 0, log_scale: float = -3.5, sh_degree: int = 3,
               unique_depth: bool = False, dtype=torch.float32) -> Dict[str, torch.Tensor]:
    g = torch.Generator().manual_seed(seed)
    pos = torch.rand(n, 3, generator=g) * 2 - 1
    scale_raw = torch.randn(n, 3, generator=g) * 0.5 + log_scale
    q_raw = torch.randn(n, 4, generator=g)
    opacity_raw = torch.randn(n, generator=g) * 2
    f_dc = torch.randn(n, 3, generator=g)
    f_rest = torch.randn(n, 45, generator=g) * 0.1
    if sh_degree == 0:
        f_rest = torch.zeros(n, 45)
    if unique_depth:
        pos[:, 2] += 1e-6 * torch.arange(n, dtype=torch.float32)
    out = dict(pos=pos, scale_raw=scale_raw, q_raw=q_raw, opacity_raw=opacity_raw, f_dc=f_dc, f_rest=f_rest)
    return {k: t.to(dtype).contiguous() for k, t in out.items()}


def make_camera(W: int, H: int, view: int = 0, n_views: int = 1, radius: float = 3.0,
                dtype=torch.float32) -> Dict[str, object]:
    """Camera on a circle of `radius` around the world Y axis looking at the origin down +z."""
    th = 2.0 * math.pi * view / max(n_views, 1)
    c, s = math.cos(th), math.sin(th)
    c2w = torch.tensor([[c, 0.0, s, -radius * s],
                        [0.0, 1.0, 0.0, 0.0],
                        [-s, 0.0, c, -radius * c],
                        [0.0, 0.0, 0.0, 1.0]], dtype=dtype)
    return dict(c2w=c2w, H=H, W=W, fx=0.9 * W, fy=0.9 * W, cx=W / 2.0, cy=H / 2.0)
