"""CPU: the C-ABI library builds, loads and exports every symbol include/b200gs.h declares; the host
package fails loudly off-GPU; install() rebinds a reference-shaped package."""
import os
import re
import sys
import types

import pytest
import torch

from common import ROOT


def test_library_exports_every_declared_symbol():
    import importlib.util
    spec = importlib.util.spec_from_file_location(
        "b200gs_build", os.path.join(ROOT, "3d-gaussian-splatting-for-novel-view-synthesis_b200", "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    mod.build()
    from b200gs import _lib
    lib = _lib.load()
    header = open(os.path.join(ROOT, "include", "b200gs.h")).read()
    declared = set(re.findall(r"\b(b200gs_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations found"
    for name in declared:
        assert hasattr(lib, name), f"libb200gs.so does not export {name}"
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    assert lib.b200gs_abi_version() == 1
    sz = _lib.Sizes()
    assert lib.b200gs_workspace_sizes(1000, 128, 128, 5000, sz) == 0 and sz.frame_bytes > 0 and sz.isect_bytes > 0
    assert lib.b200gs_workspace_sizes(-1, 128, 128, 0, sz) != 0


def test_argument_checks_answer_before_any_cuda_call():
    """Error behaviour of the C ABI that does not need a device: bad arguments come back as B200GS_ERR_* codes (no
    exception crosses the boundary, include/b200gs.h), with a message in b200gs_last_error()."""
    import ctypes
    from b200gs import _lib
    lib = _lib.load()
    ERR_ARG, ERR_WS = -1, -2
    # clip_grad_norm_ with several tensors: null table, negative length, a null tensor, too many tensors, small workspace
    one = (ctypes.c_int64 * 1)(4096)
    ptr = (ctypes.c_void_p * 1)(0x1000)
    assert lib.b200gs_clip_grad_norm_multi(None, None, 1, 1.0, None, 0, None, None) == ERR_ARG
    assert lib.b200gs_clip_grad_norm_multi(ptr, (ctypes.c_int64 * 1)(-1), 1, 1.0, None, 0, None, None) == ERR_ARG
    assert lib.b200gs_clip_grad_norm_multi((ctypes.c_void_p * 1)(None), one, 1, 1.0, None, 0, None, None) == ERR_ARG
    many = _lib.CLIP_MAX_TENSORS + 1
    assert lib.b200gs_clip_grad_norm_multi((ctypes.c_void_p * many)(*[0x1000] * many), (ctypes.c_int64 * many)(*[8] * many),
                                           many, 1.0, ptr, 1 << 20, None, None) == ERR_ARG
    assert b"B200GS_CLIP_MAX_TENSORS" in lib.b200gs_last_error()
    assert lib.b200gs_clip_grad_norm_multi(ptr, one, 1, 1.0, ptr, 16, None, None) == ERR_WS
    # the workspace grows with the number of blocks and covers the single-tensor size
    two = (ctypes.c_int64 * 2)(4096, 1)
    w1, w2 = lib.b200gs_clip_workspace_bytes_multi(one, 1), lib.b200gs_clip_workspace_bytes_multi(two, 2)
    assert w1 == lib.b200gs_clip_workspace_bytes(4096) and w2 >= w1
    big = (ctypes.c_int64 * 2)(10_000_000, 10_000_000)
    assert lib.b200gs_clip_workspace_bytes_multi(big, 2) >= 256 + 4 * (2 * (10_000_000 // 4096) + 1)
    # the single-tensor entry point and the optimizer step check their arguments the same way
    assert lib.b200gs_clip_grad_norm(None, 5, 1.0, None, 0, None, None) == ERR_ARG
    assert lib.b200gs_adam_step(None, 2, 0.9, 0.999, 1e-15, None) == ERR_ARG


def test_render_entry_point_rejects_bad_calls_with_error_codes():
    """b200gs_render_project (include/b200gs.h; the C side of gaussian_splatting.render.render, render.py:62-64): the
    argument checks run before the first CUDA call, in the order camera -> Gaussians -> workspace."""
    import ctypes
    from b200gs import _lib
    lib = _lib.load()
    ERR_ARG, ERR_WS, ERR_TILE = -1, -2, -4
    P = 0x1000                                   # any non-null, 16-byte aligned address: nothing is dereferenced

    def cam(**kw):
        c = _lib.Camera(c2w=P, H=64, W=64, fx=50., fy=50., cx=32., cy=32., near_plane=0.01, far_plane=100., pix_guard=32.,
                        min_conis=1e-6, chi_square_clip=6.25, alpha_max=0.99, alpha_cutoff=1 / 128., tile=16,
                        tile_row_begin=0, tile_row_end=0, flags=0)
        for k, v in kw.items():
            setattr(c, k, v)
        return c

    def gauss(**kw):
        g = _lib.Gaussians(n=10, pos=P, opacity_raw=P, scale_raw=P, q_raw=P, sigma=None, f_dc=P, f_rest=P, color=None)
        for k, v in kw.items():
            setattr(g, k, v)
        return g

    def call(g, c, ws=P, nbytes=1 << 40):
        return lib.b200gs_render_project(ctypes.byref(g) if g is not None else None,
                                         ctypes.byref(c) if c is not None else None, ws, nbytes, None, None)
    assert call(gauss(), None) == ERR_ARG
    assert call(gauss(), cam(c2w=None)) == ERR_ARG
    assert call(gauss(), cam(tile=8)) == ERR_TILE and b"T=16" in lib.b200gs_last_error()     # render(..., T=8)
    assert call(gauss(), cam(H=0)) == ERR_ARG
    assert call(None, cam()) == ERR_ARG
    assert call(gauss(n=-1), cam()) == ERR_ARG
    assert call(gauss(pos=None), cam()) == ERR_ARG
    assert call(gauss(scale_raw=None), cam()) == ERR_ARG          # neither (scale_raw, q_raw) nor sigma
    assert call(gauss(f_dc=None), cam()) == ERR_ARG               # neither (f_dc, f_rest) nor color
    assert call(gauss(q_raw=P + 4), cam()) == ERR_ARG and b"16-byte" in lib.b200gs_last_error()
    assert call(gauss(), cam(), ws=None) == ERR_ARG
    assert call(gauss(), cam(), nbytes=64) == ERR_WS
    sz = _lib.Sizes()
    assert lib.b200gs_workspace_sizes(10, 64, 64, 100, sz) == 0
    assert call(gauss(), cam(), nbytes=sz.frame_bytes - 1) == ERR_WS


def test_struct_layouts_match_header():
    import ctypes
    from b200gs import _lib
    assert ctypes.sizeof(_lib.FrameStats) == 64
    assert ctypes.sizeof(_lib.Gaussians) == 8 + 8 * 8
    assert ctypes.sizeof(_lib.Grads) == 8 * 8
    assert ctypes.sizeof(_lib.Camera) == 8 + 8 + 11 * 8 + 16
    assert ctypes.sizeof(_lib.PeerLayout) == 3 * 8 * 8 + 16
    assert ctypes.sizeof(_lib.PeerGroup) == 8 + 16 * 8 + 8
    assert ctypes.sizeof(_lib.PeerTensor) == 32
    assert ctypes.sizeof(_lib.Route) == 12 + 17 * 4 + 16 * 8 + 8 + 8       # static_assert of the same figure in api.cu


def test_no_cpu_fallback():
    import b200gs
    z = torch.zeros(4, 3)
    with pytest.raises(b200gs.B200GSError):
        b200gs.render(z, z, torch.zeros(4), torch.zeros(4, 3, 3), torch.eye(4), 16, 16, 1., 1., 8., 8.)
    with pytest.raises(b200gs.B200GSError):
        b200gs.build_sigma_from_params(z, torch.zeros(4, 4))
    with pytest.raises(b200gs.B200GSError):
        b200gs.evaluate_sh(z, torch.zeros(4, 45), z, torch.eye(4))
    with pytest.raises(b200gs.B200GSError):
        b200gs.compute_loss(torch.zeros(8, 8, 3), torch.zeros(8, 8, 3))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "3d-gaussian-splatting-for-novel-view-synthesis_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text and "gs_oracle" not in text, f


def test_install_rebinds_reference_shaped_package(monkeypatch):
    import b200gs
    pkg = types.ModuleType("fake_gs")
    pkg.__path__ = []
    mods = {}
    for sub, attrs in (("render", ("render",)), ("gaussian", ("build_sigma_from_params",)),
                       ("spherical_harmonics", ("evaluate_sh",)), ("losses", ("compute_loss", "l1_loss", "ssim_loss"))):
        m = types.ModuleType(f"fake_gs.{sub}")
        for attr in attrs:
            setattr(m, attr, lambda *a, **k: "reference")
        mods[sub] = m
        monkeypatch.setitem(sys.modules, f"fake_gs.{sub}", m)
        if sub != "losses":                            # the reference re-exports the render-path names only
            setattr(pkg, attrs[0], getattr(m, attrs[0]))   # re-export, shadows the submodule name like the reference
    monkeypatch.setitem(sys.modules, "fake_gs", pkg)
    b200gs.install("fake_gs")
    try:
        assert mods["render"].render is b200gs.render and pkg.render is b200gs.render
        assert mods["gaussian"].build_sigma_from_params is b200gs.build_sigma_from_params
        assert pkg.evaluate_sh is b200gs.evaluate_sh
        assert mods["losses"].compute_loss is b200gs.compute_loss and mods["losses"].ssim_loss is b200gs.ssim_loss
        assert mods["losses"].l1_loss is b200gs.l1_loss
    finally:
        b200gs.uninstall()
    assert mods["render"].render() == "reference"


def test_deferred_tensor_behaves_like_its_value():
    """evaluate_sh / build_sigma_from_params return deferred tensors; anything but `render` that touches one
    must see the real values, with autograd intact."""
    from b200gs.api import _Deferred
    a = torch.arange(6.).reshape(2, 3).requires_grad_(True)
    calls = []

    def thunk():
        calls.append(1)
        return a * 2
    d = _Deferred(thunk, (2, 3), torch.float32, torch.device("cpu"), True)
    assert d.shape == (2, 3) and d.dtype == torch.float32 and d.dim() == 2 and d.size(1) == 3 and len(d) == 2
    assert not calls, "metadata access must not run the kernel"
    (d + 1).sum().backward()
    assert calls == [1] and torch.equal(a.grad, torch.full((2, 3), 2.0))
    assert torch.equal(d, a * 2) and torch.equal(d[1], (a * 2)[1]) and d.detach().numpy().sum() == 30.0
    assert calls == [1], "materialises once"
    with torch.no_grad():
        e = _Deferred(lambda: a * 3, (2, 3), torch.float32, torch.device("cpu"), False)
    assert not (e * 1.0).requires_grad      # grad mode captured at creation
