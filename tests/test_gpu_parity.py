"""GPU parity tests: the CUDA path (through the C ABI, via the b200gs host package) against the golden
vectors of the reference and against the oracle.  Tolerances (BASELINE.json north_star):
  - integer stages (survivors, radii, tile rects/counts, per-tile sorted lists, ranges): bit-exact;
  - image: <= 1e-4 absolute (fp32), threshold-flip pixels counted against the reference's own
    fp32-vs-fp64 flips;
  - gradients: <= 1e-3 relative (max-norm per tensor).
"""
import ctypes

import numpy as np
import pytest
import torch

from common import (GOLDEN_CASES, GRAD_CASES, PARAMS, canonical_lists, golden_inputs, grad_relerr, image_report,
                    integer_stages_from_splats, load_golden)

pytestmark = pytest.mark.gpu
IMG_TOL = 1e-4
GRAD_TOL = 1e-3
# radius / tile-rect flips allowed per golden case (measured on a B200: see profiles/PARITY_r02.json); cases not
# listed fall back to 1 in 2000 survivors
RADIUS_FLIP_LIMIT = {"c1_10k_sh0_256": 0, "sh3_4k_200x136_rot": 0, "edge_1500_97x71": 0, "dense_600_48x40": 0}


@pytest.fixture(scope="module")
def gs():
    import b200gs
    b200gs.load_library()
    return b200gs


def _render_frame(gs, sc, cam, fused=True, tile_rows=None, mode=None):
    """Runs project+rasterize through the ops layer and returns (image, Frame)."""
    from b200gs import ops
    dev = sc["pos"].device
    cfg = ops.RenderConfig(H=cam["H"], W=cam["W"], fx=cam["fx"], fy=cam["fy"], cx=cam["cx"], cy=cam["cy"])
    if tile_rows:
        cfg.tile_row_begin, cfg.tile_row_end = tile_rows
    if fused:
        g, keep = ops._gaussians(sc["pos"], sc["opacity_raw"], sc["scale_raw"], sc["q_raw"], None, sc["f_dc"],
                                 sc["f_rest"], None)
    else:
        sigma = gs.build_sigma_from_params(sc["scale_raw"], sc["q_raw"])
        color = gs.evaluate_sh(sc["f_dc"], sc["f_rest"], sc["pos"], cam["c2w"])
        g, keep = ops._gaussians(sc["pos"], sc["opacity_raw"], None, None, sigma, None, None, color)
    frame = ops.Frame(g, keep, cfg, cam["c2w"].contiguous(), dev)
    img = frame.render(mode)
    torch.cuda.synchronize()
    return img, frame


# ---------------------------------------------------------------------------------------------------
# integer primitives
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n", [1, 31, 4095, 4096, 4097, 100_003, 1_500_001])
def test_exclusive_scan(gs, n):
    from b200gs import _lib
    lib = _lib.load()
    rng = np.random.default_rng(n)
    x = rng.integers(0, 40, size=n, dtype=np.uint32)
    xin = torch.from_numpy(x.view(np.int32)).cuda()
    out = torch.empty_like(xin)
    total = torch.zeros(1, dtype=torch.int32, device="cuda")
    nb = lib.b200gs_scan_scratch_bytes(n)
    scratch = torch.empty(nb, dtype=torch.uint8, device="cuda")
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    for _ in range(2):   # twice: scratch reuse must be safe
        _lib.check(lib.b200gs_exclusive_scan_u32(xin.data_ptr(), out.data_ptr(), n, total.data_ptr(), scratch.data_ptr(),
                                                 nb, st))
    torch.cuda.synchronize()
    ref = np.cumsum(x, dtype=np.uint64) - x
    assert np.array_equal(out.cpu().numpy().view(np.uint32), ref.astype(np.uint32))
    assert int(total.item()) == int(x.sum())


@pytest.mark.parametrize("n,bits", [(1, (0, 32)), (257, (0, 32)), (4096, (0, 8)), (4097, (0, 16)), (50_000, (0, 13)),
                                    (1_000_003, (0, 32)), (300_001, (3, 11)), (2_000_000, (0, 15))])
def test_radix_sort_pairs_is_a_stable_sort(gs, n, bits):
    from b200gs import _lib
    lib = _lib.load()
    rng = np.random.default_rng(n + bits[1])
    keys = rng.integers(0, 2 ** 32, size=n, dtype=np.uint64).astype(np.uint32)
    if n > 1000:
        keys[: n // 3] &= 0xFF00FFFF                   # skew: many equal digits
    vals = np.arange(n, dtype=np.uint32)
    k_in = torch.from_numpy(keys.view(np.int32)).cuda()
    v_in = torch.from_numpy(vals.view(np.int32)).cuda()
    k_out, v_out = torch.empty_like(k_in), torch.empty_like(v_in)
    nb = lib.b200gs_sort_scratch_bytes(n)
    scratch = torch.empty(nb, dtype=torch.uint8, device="cuda")
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    _lib.check(lib.b200gs_radix_sort_pairs(k_in.data_ptr(), v_in.data_ptr(), k_out.data_ptr(), v_out.data_ptr(), n,
                                           bits[0], bits[1], scratch.data_ptr(), nb, st))
    torch.cuda.synchronize()
    mask = np.uint32(((1 << (bits[1] - bits[0])) - 1) if bits[1] - bits[0] < 32 else 0xFFFFFFFF)
    digit = (keys >> np.uint32(bits[0])) & mask
    order = np.argsort(digit, kind="stable")
    assert np.array_equal(v_out.cpu().numpy().view(np.uint32), vals[order])
    assert np.array_equal(k_out.cpu().numpy().view(np.uint32), keys[order])


# ---------------------------------------------------------------------------------------------------
# per-Gaussian functions
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["sh3_4k_200x136_rot", "edge_1500_97x71"])
def test_build_sigma_and_evaluate_sh(gs, name):
    G = load_golden(name)
    sc, cam = golden_inputs(G, "cuda")
    leaves = {k: sc[k].clone().requires_grad_(True) for k in PARAMS}
    sigma = gs.build_sigma_from_params(leaves["scale_raw"], leaves["q_raw"])
    color = gs.evaluate_sh(leaves["f_dc"], leaves["f_rest"], leaves["pos"], cam["c2w"])
    ref_s, ref_c = G["sigma"], G["color"]
    assert np.abs(sigma.detach().cpu().numpy() - ref_s).max() <= 2e-6 * max(1.0, np.abs(ref_s).max())
    assert np.abs(color.detach().cpu().numpy() - ref_c).max() <= 1e-6
    # backward of the two stand-alone functions against the oracle's autograd
    from oracle import gs_oracle as O
    cpu = {k: torch.from_numpy(G["in_" + k]).clone().requires_grad_(True) for k in PARAMS}
    ws = torch.randn(ref_s.shape, generator=torch.Generator().manual_seed(1))
    wc = torch.randn(ref_c.shape, generator=torch.Generator().manual_seed(2))
    (O.build_sigma_from_params(cpu["scale_raw"], cpu["q_raw"]) * ws).sum().backward()
    (O.evaluate_sh(cpu["f_dc"], cpu["f_rest"], cpu["pos"], torch.from_numpy(G["c2w"])) * wc).sum().backward()
    ((sigma * ws.cuda()).sum() + (color * wc.cuda()).sum()).backward()
    for k in ("scale_raw", "q_raw", "f_dc", "f_rest", "pos"):
        assert grad_relerr(leaves[k].grad.cpu().numpy(), cpu[k].grad.numpy()) <= 1e-4, k


# ---------------------------------------------------------------------------------------------------
# forward: integer stages + image
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("fused", [True, False])
@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_forward_against_golden(gs, name, fused, parity_log):
    G = load_golden(name)
    sc, cam = golden_inputs(G, "cuda")
    img, frame = _render_frame(gs, sc, cam, fused=fused)
    ex = {k: v.numpy() for k, v in frame.export().items()}
    H, W = cam["H"], cam["W"]
    n = sc["pos"].shape[0]
    vis = ex["tiles_touched"] >= 0
    # (a) integer stages are bit-exact GIVEN the kernel's own (u, v, radius, z)
    on, rect, cnt, tile, ids, ranges = integer_stages_from_splats(ex["xy"][:, 0], ex["xy"][:, 1], ex["radius"],
                                                                  ex["depth"], vis, H, W)
    assert np.array_equal(on, vis)
    assert np.array_equal(rect[vis], ex["rect"][vis])
    assert np.array_equal(cnt[vis], ex["tiles_touched"][vis])
    assert frame.n_isect == int(cnt.sum()) == ex["list_id"].shape[0]
    # the lists are stored grouped by supertile: a STABLE regrouping by tile id must give exactly the
    # (tile, depth, id)-ordered list, i.e. every tile's list is bit-exact and in depth order
    regroup = np.argsort(ex["list_tile"], kind="stable")
    assert np.array_equal(ex["list_tile"][regroup], tile) and np.array_equal(ex["list_id"][regroup], ids)
    r = ex["ranges"].astype(np.int64)
    assert np.array_equal(r[:, 1] - r[:, 0], ranges[:, 1] - ranges[:, 0])
    for t in np.flatnonzero(r[:, 1] > r[:, 0])[:: max(1, r.shape[0] // 64)]:
        assert (ex["list_tile"][r[t, 0]:r[t, 1]] == t).all()
    order = ex["depth_order"][: int(vis.sum())]
    assert np.array_equal(order, np.lexsort((np.arange(n), np.where(vis, ex["depth"], np.inf)))[: int(vis.sum())])
    # (b) against the reference: survivors, depth, radii, rects, counts, per-tile lists
    gid = G["ids"]
    vis_ref = np.zeros(n, bool)
    vis_ref[gid] = True
    assert np.array_equal(vis, vis_ref), f"survivor sets differ in {(vis != vis_ref).sum()} Gaussians"
    assert frame.n_visible == gid.shape[0]
    assert np.array_equal(ex["depth"][gid], G["z"])
    rad_bad = ex["radius"][gid] != G["radius"]
    rect_bad = (ex["rect"][gid] != G["rect"]).any(1)
    n_rad, n_rect = int(rad_bad.sum()), int(rect_bad.sum())
    # Radius = ceil(2.5 sqrt(lambda_max)) and the rect's floor() are step functions of fp32 values that the GPU forms
    # with FMAs and its own exp/sqrt: a value within an ulp of an integer may land on the other side, exactly as it
    # does between the reference's own fp32 and fp64 runs.  The count is bounded (1 in 2000) AND recorded
    # (profiles/PARITY_r02.json); the list comparison below never depends on it being zero.
    limit = RADIUS_FLIP_LIMIT.get(name, max(1, gid.shape[0] // 2000))
    assert n_rad <= limit and n_rect <= limit, (n_rad, n_rect)
    # per-tile lists against the REFERENCE, always: Gaussians whose rect flipped (none on most cases) are taken out
    # of both sides, every other (tile, id) entry must match in order
    z_of = np.full(n, np.inf, np.float32)
    z_of[gid] = G["z"]
    flipped = np.zeros(n, bool)
    flipped[gid[rect_bad]] = True
    keep_r, keep_m = ~flipped[G["list_id"]], ~flipped[ex["list_id"]]
    t_ref, i_ref = canonical_lists(G["list_tile"][keep_r], G["list_id"][keep_r], z_of)
    t_me, i_me = canonical_lists(ex["list_tile"][keep_m], ex["list_id"][keep_m], z_of)
    assert np.array_equal(t_ref, t_me) and np.array_equal(i_ref, i_me)
    if n_rect == 0:
        assert frame.n_isect == G["list_id"].shape[0]
    assert np.abs(ex["xy"][gid, 0] - G["u"]).max() <= 1e-4 and np.abs(ex["xy"][gid, 1] - G["v"]).max() <= 1e-4
    assert np.abs(ex["opacity"][gid] - G["opacity"]).max() <= 1e-6
    assert np.abs(ex["color"][gid] - G["color"][gid]).max() <= 2e-6
    # image
    rep = image_report(img.cpu().numpy(), G["image"], G["image64"], IMG_TOL)
    parity_log[f"golden/{name}/{'fused' if fused else 'unfused'}"] = dict(
        rep, V=int(gid.shape[0]), I=int(frame.n_isect), radius_mismatches=n_rad, rect_mismatches=n_rect,
        survivors_equal=True, depth_bit_equal=True, lists_equal=True, lists_compared_excluding=int(flipped.sum()),
        n_values=int(G["image"].size))
    # Image: <= 1e-4 absolute, except pixels where a threshold test (q <= 6.25, alpha >= 1/128, T > 5e-5) flips for a
    # pair within a few ulp of the threshold - which also happens between the reference's own fp32 and fp64 runs, so
    # the number of such pixels is bounded by the reference's own fp32-vs-fp64 count (and recorded).
    assert rep["n_bad_min"] <= 2 * rep["n_bad_ref32_vs_ref64"] + 2, rep
    assert rep["max_vs_ref32"] <= max(0.012, 2 * rep["max_ref32_vs_ref64"]), rep
    if name != "edge_1500_97x71":
        assert rep["n_bad_vs_ref32"] <= 3, rep


def test_public_api_matches_reference_calling_convention(gs):
    """scripts/train.py:463,502,505-508: build_sigma -> evaluate_sh -> render, H/W as 0-dim tensors."""
    G = load_golden("sh3_4k_200x136_rot")
    sc, cam = golden_inputs(G, "cuda")
    with torch.no_grad():
        sigma = gs.build_sigma_from_params(sc["scale_raw"], sc["q_raw"])
        color = gs.evaluate_sh(sc["f_dc"], sc["f_rest"], sc["pos"], cam["c2w"])
        img = gs.render(sc["pos"], color, sc["opacity_raw"], sigma, cam["c2w"], torch.tensor(cam["H"]),
                        torch.tensor(cam["W"]), cam["fx"], cam["fy"], cam["cx"], cam["cy"])
        img_kw = gs.render(sc["pos"], color.clone(), sc["opacity_raw"], sigma.clone(), cam["c2w"].cpu(), cam["H"],
                           cam["W"], cam["fx"], cam["fy"], cam["cx"], cam["cy"], pix_guard=32, chi_square_clip=6.25,
                           alpha_cutoff=1 / 128.)
    assert img.shape == (cam["H"], cam["W"], 3) and img.dtype == torch.float32 and img.is_cuda
    assert float(img.min()) >= 0.0 and float(img.max()) <= 1.0
    assert np.abs(img.cpu().numpy() - G["image"]).max() <= IMG_TOL
    assert np.abs(img_kw.cpu().numpy() - G["image"]).max() <= IMG_TOL     # untagged clones -> unfused path
    with pytest.raises(NotImplementedError):
        gs.render(sc["pos"], color, sc["opacity_raw"], sigma, cam["c2w"], 64, 64, 50., 50., 32., 32., T=8)


def test_empty_and_offscreen_behaviour(gs):
    from oracle import gs_oracle as O
    sc = {k: v.cuda() for k, v in O.make_scene(50, seed=2).items()}
    cam = O.make_camera(64, 64)
    c2w = cam["c2w"].cuda()
    sigma = gs.build_sigma_from_params(sc["scale_raw"], sc["q_raw"])
    color = gs.evaluate_sh(sc["f_dc"], sc["f_rest"], sc["pos"], c2w)
    op = torch.full((50,), -20.0, device="cuda", requires_grad=True)
    img = gs.render(sc["pos"], color, op, sigma, c2w, 64, 64, 57.6, 57.6, 32., 32.)     # render.py:109-112
    assert img.shape == (64, 64, 3) and float(img.abs().max()) == 0.0
    img.sum().backward()
    assert float(op.grad.abs().max()) == 0.0
    G = load_golden("offscreen_case")
    sig2 = gs.build_sigma_from_params(torch.from_numpy(G["scale_raw"]).cuda(), torch.from_numpy(G["q_raw"]).cuda())
    with pytest.raises(Exception, match="All projected points are off-screen"):         # render.py:235-236
        gs.render(torch.from_numpy(G["pos"]).cuda(), color, torch.full((50,), 2.0, device="cuda"), sig2,
                  torch.from_numpy(G["c2w"]).cuda(), 64, 64, 57.6, 57.6, 32., 32.)
    # n = 0
    z3 = torch.zeros(0, 3, device="cuda")
    img0 = gs.render(z3, z3, torch.zeros(0, device="cuda"), torch.zeros(0, 3, 3, device="cuda"), c2w, 32, 48, 40., 40.,
                     24., 16.)
    assert img0.shape == (32, 48, 3) and float(img0.abs().max()) == 0.0


# ---------------------------------------------------------------------------------------------------
# backward
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("fused", [True, False])
@pytest.mark.parametrize("name", GRAD_CASES)
def test_gradients_against_reference_autograd(gs, name, fused, monkeypatch, parity_log):
    monkeypatch.setenv("B200GS_FUSE", "1" if fused else "0")
    G = load_golden(name)
    sc, cam = golden_inputs(G, "cuda")
    leaves = {k: sc[k].clone().requires_grad_(True) for k in PARAMS}
    sigma = gs.build_sigma_from_params(leaves["scale_raw"], leaves["q_raw"])
    color = gs.evaluate_sh(leaves["f_dc"], leaves["f_rest"], leaves["pos"], cam["c2w"])
    img = gs.render(leaves["pos"], color, leaves["opacity_raw"], sigma, cam["c2w"], cam["H"], cam["W"], cam["fx"],
                    cam["fy"], cam["cx"], cam["cy"])
    (img * torch.from_numpy(G["loss_w"]).cuda()).sum().backward()
    rec = {}
    for k in PARAMS:
        mine = leaves[k].grad.cpu().numpy()
        assert np.isfinite(mine).all(), k
        e32, e64 = grad_relerr(mine, G["grad_" + k]), grad_relerr(mine, G["grad64_" + k])
        noise = grad_relerr(G["grad_" + k], G["grad64_" + k])
        rec[k] = {"vs_ref32": e32, "vs_ref64": e64, "ref32_vs_ref64": noise}
        assert min(e32, e64) <= max(GRAD_TOL, 2 * noise), (name, k, e32, e64, noise)
    parity_log[f"golden_grads/{name}/{'fused' if fused else 'unfused'}"] = rec


def test_gradient_accumulates_over_views(gs):
    """train.py:463-530: one sigma, several views, one backward."""
    from oracle import gs_oracle as O
    sc_cpu = O.make_scene(800, seed=11, log_scale=-2.5)
    cams = [O.make_camera(80, 64, view=k, n_views=4) for k in range(2)]
    w = [torch.rand(64, 80, 3, generator=torch.Generator().manual_seed(k)) for k in range(2)]
    ref = {k: v.clone().requires_grad_(True) for k, v in sc_cpu.items()}
    sig = O.build_sigma_from_params(ref["scale_raw"], ref["q_raw"])
    loss = 0
    for cam, wk in zip(cams, w):
        col = O.evaluate_sh(ref["f_dc"], ref["f_rest"], ref["pos"], cam["c2w"])
        loss = loss + (O.render(ref["pos"], col, ref["opacity_raw"], sig, cam["c2w"], 64, 80, cam["fx"], cam["fy"],
                                cam["cx"], cam["cy"]) * wk).sum() / 2
    loss.backward()
    mine = {k: v.cuda().requires_grad_(True) for k, v in sc_cpu.items()}
    sig = gs.build_sigma_from_params(mine["scale_raw"], mine["q_raw"])
    loss = 0
    for cam, wk in zip(cams, w):
        c2w = cam["c2w"].cuda()
        col = gs.evaluate_sh(mine["f_dc"], mine["f_rest"], mine["pos"], c2w)
        loss = loss + (gs.render(mine["pos"], col, mine["opacity_raw"], sig, c2w, 64, 80, cam["fx"], cam["fy"],
                                 cam["cx"], cam["cy"]) * wk.cuda()).sum() / 2
    loss.backward()
    for k in PARAMS:
        assert grad_relerr(mine[k].grad.cpu().numpy(), ref[k].grad.numpy()) <= GRAD_TOL, k


# ---------------------------------------------------------------------------------------------------
# full-size properties (BASELINE.json configs): no oracle at this size, size-independent invariants
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n,W,H,ls", [(1_000_000, 1920, 1080, -5.5), (300_000, 1297, 840, -5.0)])
def test_full_size_invariants(gs, n, W, H, ls):
    from oracle import gs_oracle as O
    sc = {k: v.cuda() for k, v in O.make_scene(n, seed=0, log_scale=ls).items()}
    cam = O.make_camera(W, H)
    cam["c2w"] = cam["c2w"].cuda()
    img, frame = _render_frame(gs, sc, cam)
    ex = {k: v.numpy() for k, v in frame.export().items()}
    vis = ex["tiles_touched"] >= 0
    assert frame.n_visible == int(vis.sum())
    assert frame.n_isect == int(ex["tiles_touched"][vis].sum()) == ex["list_id"].shape[0]
    regroup = np.argsort(ex["list_tile"], kind="stable")                   # stored grouped by supertile
    lt, li = ex["list_tile"][regroup].astype(np.int64), ex["list_id"][regroup].astype(np.int64)
    assert (np.diff(lt) >= 0).all()                                        # grouped by tile
    z = ex["depth"][li]
    same = lt[1:] == lt[:-1]
    assert (np.diff(z)[same] >= 0).all()                                   # depth-sorted inside every tile
    tie = same & (np.diff(z) == 0)
    assert (np.diff(li)[tie] > 0).all()                                    # ties in index order
    r = ex["ranges"]
    nonempty = r[:, 1] > r[:, 0]
    assert int((r[:, 1] - r[:, 0]).sum()) == frame.n_isect
    assert np.array_equal(np.unique(lt), np.flatnonzero(nonempty))
    starts = np.sort(r[nonempty, 0])
    assert np.array_equal(np.sort(r[nonempty, 1])[:-1], starts[1:]) and starts[0] == 0   # ranges tile the buffer
    # every (tile, id) pair lies inside the Gaussian's tile rect, and each pair appears once
    tx, ty = lt % ((W + 15) // 16), lt // ((W + 15) // 16)
    rc = ex["rect"][li]
    assert ((tx >= rc[:, 0]) & (tx <= rc[:, 1]) & (ty >= rc[:, 2]) & (ty <= rc[:, 3])).all()
    assert np.unique(lt * n + li).shape[0] == lt.shape[0]
    im = img.cpu().numpy()
    assert np.isfinite(im).all() and im.min() >= 0.0 and im.max() <= 1.0
    # determinism + the unfused path + tile-row bands give the same picture
    img2, _ = _render_frame(gs, sc, cam)
    assert torch.equal(img, img2)
    img_u, _ = _render_frame(gs, sc, cam, fused=False)
    assert float((img_u - img).abs().max()) <= IMG_TOL
    rows = (H + 15) // 16
    top, _ = _render_frame(gs, sc, cam, tile_rows=(0, rows // 3))
    bot, _ = _render_frame(gs, sc, cam, tile_rows=(rows // 3, rows))
    assert torch.equal(top + bot, img)
    spec1, _ = _render_frame(gs, sc, cam, mode="speculative")
    spec2, _ = _render_frame(gs, sc, cam, mode="speculative")
    assert torch.equal(spec1, img) and torch.equal(spec2, img)


# ---------------------------------------------------------------------------------------------------
# live oracle comparison on shapes of the BASELINE configs, scaled so the CPU oracle takes seconds
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n,W,H,ls,view,seed", [(20_000, 325, 211, -3.6, 2, 21),     # C3-like aspect: 1-px edge column
                                                 (12_000, 480, 270, -3.3, 5, 22),     # 1080p aspect, 14-px edge row
                                                 (5_000, 64, 64, -2.2, 1, 23)])       # few tiles, long lists (many batches)
def test_live_oracle_forward_backward(gs, n, W, H, ls, view, seed, parity_log):
    from oracle import gs_oracle as O
    sc = O.make_scene(n, seed=seed, log_scale=ls, unique_depth=True)
    cam = O.make_camera(W, H, view=view, n_views=8)
    w = torch.rand(H, W, 3, generator=torch.Generator().manual_seed(seed))
    ref = {k: v.clone().requires_grad_(True) for k, v in sc.items()}
    img_ref = O.render_from_params(ref["pos"], ref["scale_raw"], ref["q_raw"], ref["opacity_raw"], ref["f_dc"],
                                   ref["f_rest"], cam["c2w"], H, W, cam["fx"], cam["fy"], cam["cx"], cam["cy"])
    (img_ref * w).sum().backward()
    mine = {k: v.cuda().requires_grad_(True) for k, v in sc.items()}
    c2w = cam["c2w"].cuda()
    sigma = gs.build_sigma_from_params(mine["scale_raw"], mine["q_raw"])
    color = gs.evaluate_sh(mine["f_dc"], mine["f_rest"], mine["pos"], c2w)
    img = gs.render(mine["pos"], color, mine["opacity_raw"], sigma, c2w, H, W, cam["fx"], cam["fy"], cam["cx"], cam["cy"])
    (img * w.cuda()).sum().backward()
    d = (img.detach().cpu() - img_ref.detach()).abs()
    n_bad = int((d > IMG_TOL).sum())
    errs = {k: grad_relerr(mine[k].grad.cpu().numpy(), ref[k].grad.numpy()) for k in PARAMS}
    parity_log[f"live_oracle/{n}_{W}x{H}"] = {"max_abs": float(d.max()), "n_gt_tol": n_bad, "n_values": int(d.numel()),
                                               "grad_rel": errs}
    assert n_bad <= max(3, d.numel() // 50_000), (n_bad, float(d.max()))      # threshold-flip pixels only
    assert float(d.max()) <= 0.02
    for k in PARAMS:
        assert errs[k] <= 3 * GRAD_TOL if n_bad else errs[k] <= GRAD_TOL, (k, errs[k], n_bad)


def test_speculative_capacity_overflow_is_redone_with_exact_buffers(gs):
    """speculative mode sizes the lists from a high-water mark and waits only for the frame's counters;
    a frame that does not fit must be rasterized again and give the sync-mode picture."""
    from b200gs import ops
    from oracle import gs_oracle as O
    sc = {k: v.cuda() for k, v in O.make_scene(30_000, seed=5, log_scale=-3.5).items()}
    cam = O.make_camera(320, 200)
    cam["c2w"] = cam["c2w"].cuda()
    ref, fr = _render_frame(gs, sc, cam, mode="sync")
    fr.refresh_stats()
    assert fr.n_isect > 4000
    dev = (sc["pos"].device.index, 30_000, 200, 320)      # the mark is kept per (device, N, H, W)
    saved = ops._high_water.get(dev, 0)
    try:
        ops._high_water[dev] = 1000                      # far too small for this frame
        img, fr2 = _render_frame(gs, sc, cam, mode="speculative")
        assert torch.equal(img, ref) and fr2.n_isect == fr.n_isect and not fr2.overflow
        assert ops._high_water[dev] >= fr.n_isect        # grown: the next frame fits without a redo
        img3, _ = _render_frame(gs, sc, cam, mode="speculative")
        assert torch.equal(img3, ref)
    finally:
        ops._high_water[dev] = max(saved, ops._high_water.get(dev, 0))


def test_frame_sink_matches_the_reference_host_conversion(gs):
    """render_trained.py:357 / inference.py:117: (img.cpu().numpy() * 255).astype(np.uint8)."""
    g = torch.Generator().manual_seed(5)
    for shape in [(1080, 1920, 3), (97, 71, 3), (5, 7, 3), (1, 1, 3)]:
        img = torch.rand(*shape, generator=g)
        img.view(-1)[:: 7] = 1.0
        img.view(-1)[3:: 11] = 0.0
        edges = (torch.arange(img.numel()) % 256).float().div(255.0)          # values sitting on the u8 grid
        img.view(-1)[1:: 5] = edges[1:: 5]
        want = (img.numpy() * 255).astype(np.uint8)
        got = gs.to_uint8(img.cuda())
        assert got.dtype == torch.uint8 and got.shape == img.shape
        assert np.array_equal(got.cpu().numpy(), want)


def test_render_pipeline_gives_the_same_frames(gs):
    """RenderPipeline overlaps the binning of frame i+1 with the blend of frame i on two streams; every frame must
    be bit-identical to the one-frame-at-a-time render."""
    from oracle import gs_oracle as O
    sc = {k: v.cuda() for k, v in O.make_scene(60_000, seed=8, log_scale=-4.2).items()}
    cams = [O.make_camera(640, 360, view=v, n_views=6) for v in range(6)]
    c2ws = [c["c2w"].cuda() for c in cams]
    K = cams[0]
    with torch.no_grad():
        sigma = gs.build_sigma_from_params(sc["scale_raw"], sc["q_raw"])
        want = []
        for c2w in c2ws:
            col = gs.evaluate_sh(sc["f_dc"], sc["f_rest"], sc["pos"], c2w)
            want.append(gs.render(sc["pos"], col, sc["opacity_raw"], sigma, c2w, 360, 640, K["fx"], K["fy"], K["cx"], K["cy"]))
        torch.cuda.synchronize()
        pipe = gs.RenderPipeline()
        got = []
        for rep in range(2):                       # several rounds: buffers are recycled across streams
            for c2w in c2ws:
                col = gs.evaluate_sh(sc["f_dc"], sc["f_rest"], sc["pos"], c2w)
                img = pipe.render(sc["pos"], col, sc["opacity_raw"], sigma, c2w, 360, 640, K["fx"], K["fy"], K["cx"], K["cy"])
                with torch.cuda.stream(pipe.blend_stream):
                    got.append(img.clone())
        # submit / result with frames in flight, consumed on a third stream through the completion events
        side = torch.cuda.Stream()
        tickets = []
        for c2w in c2ws:
            col = gs.evaluate_sh(sc["f_dc"], sc["f_rest"], sc["pos"], c2w)
            tickets.append(pipe.submit(sc["pos"], col, sc["opacity_raw"], sigma, c2w, 360, 640, K["fx"], K["fy"], K["cx"], K["cy"]))
            if len(tickets) >= 3:
                t = tickets.pop(0)
                img = pipe.result(t)
                with torch.cuda.stream(side):
                    side.wait_event(pipe.done_event(t))
                    got.append(img.clone())
        for t in tickets:
            img = pipe.result(t)
            with torch.cuda.stream(side):
                side.wait_event(pipe.done_event(t))
                got.append(img.clone())
        pipe.synchronize()
        side.synchronize()
        # a frame that overflows its speculative capacity while others are in flight is redone transparently
        from b200gs import ops
        ops._high_water[(0, 60_000, 360, 640)] = 2000
        col = gs.evaluate_sh(sc["f_dc"], sc["f_rest"], sc["pos"], c2ws[0])
        redo = pipe.render(sc["pos"], col, sc["opacity_raw"], sigma, c2ws[0], 360, 640, K["fx"], K["fy"], K["cx"], K["cy"])
        pipe.synchronize()
        assert torch.equal(redo, want[0])
    for i, g in enumerate(got):
        assert torch.equal(g, want[i % len(want)]), i


@pytest.mark.parametrize("kw", [
    dict(near=0.8, far=3.4, pix_guard=8, chi_square_clip=4.0, alpha_max=0.9, alpha_cutoff=1 / 64., min_conis=1e-4),
    dict(near=0.01, far=100.0, pix_guard=64, chi_square_clip=9.21, alpha_max=0.999, alpha_cutoff=1 / 255.),   # README values
    dict(chi_square_clip=1.5, alpha_cutoff=0.2),            # cutoff gate tighter than the chi-square gate for most splats
    dict(alpha_cutoff=0.0),                                 # every alpha passes the cutoff
])
def test_non_default_render_arguments_against_live_oracle(gs, kw):
    """The keyword arguments of render.py:62-64 other than T: forward image and all six gradients against the oracle
    (the blend's single-compare gate, the frustum test and the pre-cull all depend on them)."""
    from oracle import gs_oracle as O
    sc = O.make_scene(9_000, seed=17, log_scale=-3.4, unique_depth=True)
    cam = O.make_camera(200, 136, view=3, n_views=8)
    w = torch.rand(136, 200, 3, generator=torch.Generator().manual_seed(2))
    ref = {k: v.clone().requires_grad_(True) for k, v in sc.items()}
    img_ref = O.render_from_params(ref["pos"], ref["scale_raw"], ref["q_raw"], ref["opacity_raw"], ref["f_dc"],
                                   ref["f_rest"], cam["c2w"], 136, 200, cam["fx"], cam["fy"], cam["cx"], cam["cy"], **kw)
    (img_ref * w).sum().backward()
    mine = {k: v.cuda().requires_grad_(True) for k, v in sc.items()}
    c2w = cam["c2w"].cuda()
    sigma = gs.build_sigma_from_params(mine["scale_raw"], mine["q_raw"])
    color = gs.evaluate_sh(mine["f_dc"], mine["f_rest"], mine["pos"], c2w)
    img = gs.render(mine["pos"], color, mine["opacity_raw"], sigma, c2w, 136, 200, cam["fx"], cam["fy"], cam["cx"],
                    cam["cy"], **kw)
    (img * w.cuda()).sum().backward()
    d = (img.detach().cpu() - img_ref.detach()).abs()
    n_bad = int((d > IMG_TOL).sum())
    assert n_bad <= 3, (n_bad, float(d.max()))
    assert float(img_ref.max()) > 0.05                       # the case renders something
    for k in PARAMS:
        err = grad_relerr(mine[k].grad.cpu().numpy(), ref[k].grad.numpy())
        assert err <= 3 * GRAD_TOL if n_bad else err <= GRAD_TOL, (k, err, n_bad)


def test_backward_is_linear_in_the_image_gradient_at_full_size(gs):
    """No oracle at 1M Gaussians / 1080p: the backward of a fixed frame is a linear map of dL/dimage, so
    bwd(a g1 + b g2) = a bwd(g1) + b bwd(g2) for all six parameter tensors (float atomics: 2e-4 of the max-norm)."""
    from oracle import gs_oracle as O
    sc = O.make_scene(1_000_000, seed=0, log_scale=-5.5)
    cam = O.make_camera(1920, 1080, view=2, n_views=16)
    leaves = {k: v.cuda().requires_grad_(True) for k, v in sc.items()}
    c2w = cam["c2w"].cuda()
    sigma = gs.build_sigma_from_params(leaves["scale_raw"], leaves["q_raw"])
    color = gs.evaluate_sh(leaves["f_dc"], leaves["f_rest"], leaves["pos"], c2w)
    img = gs.render(leaves["pos"], color, leaves["opacity_raw"], sigma, c2w, 1080, 1920, cam["fx"], cam["fy"], cam["cx"], cam["cy"])
    g = torch.Generator(device="cuda").manual_seed(3)
    g1 = torch.randn(img.shape, device="cuda", generator=g)
    g2 = torch.rand(img.shape, device="cuda", generator=g)
    params = [leaves[k] for k in PARAMS]
    d1 = torch.autograd.grad(img, params, g1, retain_graph=True)
    d2 = torch.autograd.grad(img, params, g2, retain_graph=True)
    d3 = torch.autograd.grad(img, params, 0.75 * g1 - 2.0 * g2)
    for k, a, b, c in zip(PARAMS, d1, d2, d3):
        assert torch.isfinite(c).all(), k
        want = 0.75 * a - 2.0 * b
        scale = float(torch.maximum(a.abs().max(), b.abs().max()))
        assert scale > 0, k
        assert float((c - want).abs().max()) <= 2e-4 * scale, (k, float((c - want).abs().max()), scale)


def test_deferred_tensors_behave_like_the_tensors_they_stand_for(gs, tmp_path):
    """build_sigma_from_params / evaluate_sh return deferred tensors (api._Deferred) that only `render` can consume
    without running the stand-alone kernel; everything else a script may do with them must give the eager result:
    .cpu(), indexing, arithmetic, torch.save / torch.load, requires_grad / grad_fn, backward through them."""
    from oracle import gs_oracle as O
    sc = O.make_scene(2000, seed=3, log_scale=-3.0)
    cam = O.make_camera(96, 64)
    lv = {k: v.cuda().requires_grad_(True) for k, v in sc.items()}
    c2w = cam["c2w"].cuda()
    ref_sigma = O.build_sigma_from_params(sc["scale_raw"], sc["q_raw"])
    ref_color = O.evaluate_sh(sc["f_dc"], sc["f_rest"], sc["pos"], cam["c2w"])
    sigma = gs.build_sigma_from_params(lv["scale_raw"], lv["q_raw"])
    color = gs.evaluate_sh(lv["f_dc"], lv["f_rest"], lv["pos"], c2w)
    assert sigma.shape == (2000, 3, 3) and color.shape == (2000, 3) and sigma.dtype == torch.float32 and sigma.is_cuda
    assert len(sigma) == 2000 and sigma.dim() == 3 and color.numel() == 6000
    assert float((sigma.cpu() - ref_sigma).abs().max()) <= 2e-6 * float(ref_sigma.abs().max())
    assert float((color[5:9].cpu() - ref_color[5:9]).abs().max()) <= 1e-6                       # indexing
    assert float(((color * 2 + 1).cpu() - (ref_color * 2 + 1)).abs().max()) <= 2e-6             # arithmetic
    assert sigma.requires_grad and color.requires_grad
    torch.save({"sigma": sigma, "color": color.detach()}, tmp_path / "d.pt")                     # serialisation
    back = torch.load(tmp_path / "d.pt")
    assert float((back["sigma"].cpu() - ref_sigma).abs().max()) <= 2e-6 * float(ref_sigma.abs().max())
    assert float((back["color"].cpu() - ref_color).abs().max()) <= 1e-6
    # autograd through a materialised deferred tensor reaches the leaves
    s2 = gs.build_sigma_from_params(lv["scale_raw"], lv["q_raw"])
    (s2.sum() + gs.evaluate_sh(lv["f_dc"], lv["f_rest"], lv["pos"], c2w).sum()).backward()
    assert lv["scale_raw"].grad is not None and lv["f_rest"].grad is not None and float(lv["f_dc"].grad.abs().max()) > 0
    with torch.no_grad():
        s3 = gs.build_sigma_from_params(lv["scale_raw"], lv["q_raw"])
        assert not s3.requires_grad
        assert float((s3.cpu() - ref_sigma).abs().max()) <= 2e-6 * float(ref_sigma.abs().max())


def test_non_fp32_inputs_are_rejected(gs):
    """The reference follows pos.dtype (render.py:318); this path computes in fp32 only and says so instead of
    silently rounding an fp64 scene."""
    from oracle import gs_oracle as O
    sc = {k: v.cuda().double() for k, v in O.make_scene(64, seed=1).items()}
    cam = O.make_camera(32, 32)
    with pytest.raises(TypeError, match="float32"):
        gs.build_sigma_from_params(sc["scale_raw"], sc["q_raw"])
    with pytest.raises(TypeError, match="float32"):
        gs.evaluate_sh(sc["f_dc"], sc["f_rest"], sc["pos"], cam["c2w"].cuda().double())
    with pytest.raises(TypeError, match="float32"):
        gs.render(sc["pos"], torch.zeros(64, 3, device="cuda", dtype=torch.float64), sc["opacity_raw"],
                  torch.zeros(64, 3, 3, device="cuda", dtype=torch.float64), cam["c2w"].cuda(), 32, 32, 28., 28., 16., 16.)


def test_mismatched_row_counts_raise_instead_of_reading_out_of_bounds(gs):
    from oracle import gs_oracle as O
    sc = {k: v.cuda() for k, v in O.make_scene(100, seed=1).items()}
    cam = O.make_camera(32, 32)
    c2w = cam["c2w"].cuda()
    sigma = gs.build_sigma_from_params(sc["scale_raw"], sc["q_raw"])
    color = gs.evaluate_sh(sc["f_dc"], sc["f_rest"], sc["pos"], c2w)
    with pytest.raises(ValueError):
        gs.render(sc["pos"][:50], color, sc["opacity_raw"][:50], sigma, c2w, 32, 32, 28., 28., 16., 16.)


def test_grad_bucket_backward_writes_into_the_flat_buffer(gs):
    """b200gs.dist.GradBucket on one GPU: with one view per backward the render backward writes each leaf's gradient
    straight into its slot of the flat buffer (autograd adopts the view as .grad, no copy), the values are those of a
    plain backward, and the slots are handed out again after allreduce() / zero_grad()."""
    from b200gs.dist import GradBucket
    from oracle import gs_oracle as O
    sc = O.make_scene(3000, seed=4, log_scale=-3.0)
    cam = O.make_camera(160, 112, view=1, n_views=6)
    c2w = cam["c2w"].cuda()
    w = torch.rand(112, 160, 3, generator=torch.Generator().manual_seed(0)).cuda()

    def backward(lv):
        sigma = gs.build_sigma_from_params(lv["scale_raw"], lv["q_raw"])
        color = gs.evaluate_sh(lv["f_dc"], lv["f_rest"], lv["pos"], c2w)
        img = gs.render(lv["pos"], color, lv["opacity_raw"], sigma, c2w, 112, 160, cam["fx"], cam["fy"], cam["cx"], cam["cy"])
        (img * w).sum().backward()
    plain = {k: v.cuda().requires_grad_(True) for k, v in sc.items()}
    backward(plain)
    lv = {k: v.cuda().requires_grad_(True) for k, v in sc.items()}
    bucket = GradBucket([lv[k] for k in PARAMS])
    for rep in range(2):
        bucket.zero_grad()
        backward(lv)
        for i, k in enumerate(PARAMS):
            assert lv[k].grad.data_ptr() == bucket.views[i].data_ptr(), (rep, k)
        bucket.allreduce()
        for k in PARAMS:
            scale = float(plain[k].grad.abs().max())
            assert float((lv[k].grad - plain[k].grad).abs().max()) <= 2e-5 * scale, k


def test_inplace_change_between_forward_and_backward_raises(gs):
    """The backward recomputes the projection from the inputs' live memory, so an in-place modification after the
    forward must raise (as autograd does for tensors it saves), and a deferred sigma whose sources changed must not be
    evaluated from the new values."""
    from oracle import gs_oracle as O
    sc = O.make_scene(500, seed=2, log_scale=-3.0)
    cam = O.make_camera(64, 48)
    lv = {k: v.cuda().requires_grad_(True) for k, v in sc.items()}
    c2w = cam["c2w"].cuda()
    sigma = gs.build_sigma_from_params(lv["scale_raw"], lv["q_raw"])
    color = gs.evaluate_sh(lv["f_dc"], lv["f_rest"], lv["pos"], c2w)
    img = gs.render(lv["pos"], color, lv["opacity_raw"], sigma, c2w, 48, 64, cam["fx"], cam["fy"], cam["cx"], cam["cy"])
    with torch.no_grad():
        lv["pos"].add_(0.01)
    with pytest.raises(RuntimeError, match="modified by an inplace operation"):
        img.sum().backward()
    s2 = gs.build_sigma_from_params(lv["scale_raw"], lv["q_raw"])
    with torch.no_grad():
        lv["scale_raw"].mul_(1.1)
    with pytest.raises(RuntimeError, match="modified in place"):
        s2.cpu()
