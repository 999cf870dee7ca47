"""b200gs - B200-native (sm_100a) differentiable Gaussian-splat rasterizer.

Drop-in for the render path of ashu1069/3D-Gaussian-Splatting-for-Novel-View-Synthesis:
`build_sigma_from_params`, `evaluate_sh`, `render`, and the loss that consumes the image (`compute_loss`,
`l1_loss`, `ssim_loss`) - same names/arguments as `gaussian_splatting` -,
`install()` to rebind them inside an imported reference package, `python -m b200gs.run <script>` to run
a reference script unchanged on top of it.
"""
from .api import RenderPipeline, build_sigma_from_params, evaluate_sh, render, to_uint8
from .losses import compute_loss, compute_loss_tensors, l1_loss, ssim_loss
from .install import install, uninstall
from .optim import FusedAdam, clip_grad_norm_
from .peer import PeerAdam, peer_allreduce_gradients
from .io import load_cameras, load_gaussians, save_gaussians
from .densify import densify_and_prune, densify_tensors
from ._lib import B200GSError, LIB_PATH, load as load_library

__all__ = ["build_sigma_from_params", "evaluate_sh", "render", "compute_loss", "compute_loss_tensors", "l1_loss",
           "ssim_loss", "to_uint8", "RenderPipeline", "FusedAdam", "clip_grad_norm_", "PeerAdam", "peer_allreduce_gradients", "load_gaussians", "save_gaussians", "load_cameras", "densify_and_prune", "densify_tensors", "install", "uninstall", "B200GSError",
           "LIB_PATH", "load_library"]
