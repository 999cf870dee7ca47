"""Rebind the reference's render-path callables (and the loss that consumes the image) to the b200gs implementations.

The reference has no plugin mechanism: its scripts do `from gaussian_splatting.render import render`
etc. at import time (scripts/train.py:41-45, scripts/render_trained.py:22-25, scripts/inference.py:33-36).
`install()` imports the reference package (it must be importable, e.g. the reference checkout on
sys.path) and replaces the attributes on its modules and on the package re-exports
(gaussian_splatting/__init__.py:7-21), so scripts imported afterwards pick up the CUDA path unchanged.
"""
from __future__ import annotations

import importlib
import sys

_PATCHES = (
    ("gaussian_splatting.render", "render"),
    ("gaussian_splatting.gaussian", "build_sigma_from_params"),
    ("gaussian_splatting.spherical_harmonics", "evaluate_sh"),
    # the consumer of the image in every training iteration (scripts/train.py:44,511)
    ("gaussian_splatting.losses", "compute_loss"),
    ("gaussian_splatting.losses", "l1_loss"),
    ("gaussian_splatting.losses", "ssim_loss"),
)
_saved = {}


def install(package: str = "gaussian_splatting", optimizer: bool = False):
    """optimizer=True additionally rebinds torch.optim.Adam and torch.nn.utils.clip_grad_norm_ to the fused
    versions (the reference script builds `optim.Adam(...)` itself, scripts/train.py:394-401,536)."""
    from . import api, losses
    if optimizer:
        import torch
        from . import optim as _optim
        _saved.setdefault(("torch.optim", "Adam"), torch.optim.Adam)
        _saved.setdefault(("torch.nn.utils", "clip_grad_norm_"), torch.nn.utils.clip_grad_norm_)
        torch.optim.Adam = _optim.FusedAdam
        torch.nn.utils.clip_grad_norm_ = _optim.clip_grad_norm_
    impl = {"compute_loss": losses.compute_loss, "l1_loss": losses.l1_loss, "ssim_loss": losses.ssim_loss}
    importlib.import_module(package)
    pkg = sys.modules[package]
    for mod_name, attr in _PATCHES:
        mod_name = mod_name.replace("gaussian_splatting", package, 1)
        # NB: the package __init__ rebinds `gaussian_splatting.render` to the function, so the module has
        # to come from sys.modules, not from attribute access on the package.
        importlib.import_module(mod_name)
        mod = sys.modules[mod_name]
        _saved.setdefault((mod_name, attr), getattr(mod, attr))
        fn = impl.get(attr) or getattr(api, attr)
        setattr(mod, attr, fn)
        if hasattr(pkg, attr):
            _saved.setdefault((package, attr), getattr(pkg, attr))
            setattr(pkg, attr, fn)
    return pkg


def uninstall():
    for (mod_name, attr), fn in list(_saved.items()):
        mod = sys.modules.get(mod_name)
        if mod is not None:
            setattr(mod, attr, fn)
        del _saved[(mod_name, attr)]
