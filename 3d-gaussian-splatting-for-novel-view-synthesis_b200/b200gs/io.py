"""On-disk formats either side of the render path: the trained-Gaussian files the
reference's scripts write and read, and the camera files of a data directory.  Host-side I/O only - nothing
here computes; tensors are staged through pinned memory straight into the contiguous fp32 device arrays the
kernels take (the SoA-across-parameters layout of DESIGN.md section 2).

  checkpoint dict   scripts/train.py:197-208      {'iteration', 'pos', 'opacity_raw', 'f_dc', 'f_rest', 'scale_raw', 'q_raw'}
                                                   as `checkpoint_{iteration:06d}.pt` / `checkpoint_final.pt`
  loose files       scripts/train.py:591-597      pos_{it}.pt opacity_raw_{it}.pt f_dc_{it}.pt f_rest_{it}.pt
                                                   scale_raw_{it}.pt and - sic - q_rot_{it}.pt
  discovery order   scripts/render_trained.py:116-182   checkpoint file -> the six loose files -> the latest
                                                   checkpoint_*.pt -> FileNotFoundError
  cameras           scripts/render_trained.py:185-202, gaussian_splatting/data_loader.py:30-47,196-216
                                                   cam_meta.npy (pickled dict: fx fy [cx cy] height width), poses.npy [K,4,4]
"""
from __future__ import annotations

from pathlib import Path
from typing import Dict, Optional, Union

import numpy as np
import torch

PARAMS = ("pos", "opacity_raw", "f_dc", "f_rest", "scale_raw", "q_raw")
SHAPES = {"pos": (3,), "opacity_raw": (), "f_dc": (3,), "f_rest": (45,), "scale_raw": (3,), "q_raw": (4,)}
_LOOSE_NAME = {"pos": "pos", "opacity_raw": "opacity_raw", "f_dc": "f_dc", "f_rest": "f_rest", "scale_raw": "scale_raw",
               "q_raw": "q_rot"}          # scripts/train.py:597 saves q_raw as q_rot_{it}.pt


def _torch_load(path):
    try:
        return torch.load(path, map_location="cpu", mmap=True, weights_only=True)
    except (RuntimeError, TypeError, ValueError):        # legacy (non-zip) serialization cannot be memory-mapped
        return torch.load(path, map_location="cpu", weights_only=True)


def _to_device(t: torch.Tensor, device) -> torch.Tensor:
    t = t.detach()
    if t.dtype != torch.float32 or not t.is_contiguous():
        t = t.to(torch.float32).contiguous()
    device = torch.device(device)
    if device.type != "cuda":
        return t.clone()
    return t.pin_memory().to(device, non_blocking=True)


def find_gaussian_files(checkpoint_dir: Union[str, Path], iteration: Union[str, int] = "final"):
    """The reference's discovery rule (scripts/render_trained.py:116-162).  Returns ('checkpoint', path) or
    ('loose', {name: path})."""
    d = Path(checkpoint_dir)
    if not d.exists():
        raise FileNotFoundError(f"Checkpoint directory does not exist: {d}")
    if iteration == "final":
        ckpt, suffix = d / "checkpoint_final.pt", "final"
    else:
        ckpt, suffix = d / f"checkpoint_{int(iteration):06d}.pt", str(int(iteration))
    if ckpt.exists():
        return "checkpoint", ckpt
    loose = {k: d / f"{_LOOSE_NAME[k]}_{suffix}.pt" for k in PARAMS}
    if all(p.exists() for p in loose.values()):
        return "loose", loose
    available = sorted(d.glob("checkpoint_*.pt"))
    if available:
        return "checkpoint", available[-1]
    raise FileNotFoundError(f"Could not find checkpoint files for iteration {iteration} in {d} "
                            f"(files: {[f.name for f in d.glob('*.pt')]})")


def load_gaussians(checkpoint_dir: Union[str, Path], iteration: Union[str, int] = "final", device="cuda",
                   requires_grad: bool = False) -> Dict[str, torch.Tensor]:
    """The six parameter tensors of a trained scene, on `device`, contiguous fp32:
    pos [N,3], opacity_raw [N], f_dc [N,3], f_rest [N,45], scale_raw [N,3], q_raw [N,4]  (+ 'iteration' when the
    file records it)."""
    kind, where = find_gaussian_files(checkpoint_dir, iteration)
    out: Dict[str, torch.Tensor] = {}
    if kind == "checkpoint":
        ck = _torch_load(where)
        for k in PARAMS:
            out[k] = ck[k]
        it = ck.get("iteration")
    else:
        for k in PARAMS:
            out[k] = _torch_load(where[k])
        it = None
    n = out["pos"].shape[0]
    for k in PARAMS:
        if tuple(out[k].shape) != (n,) + SHAPES[k]:
            raise ValueError(f"{k}: expected shape {(n,) + SHAPES[k]}, file holds {tuple(out[k].shape)}")
        out[k] = _to_device(out[k], device)
        if requires_grad:
            out[k].requires_grad_(True)
    if torch.device(device).type == "cuda":
        torch.cuda.current_stream(torch.device(device)).synchronize()       # the pinned staging buffers may go
    if it is not None:
        out["iteration"] = int(it)
    return out


def save_gaussians(params: Dict[str, torch.Tensor], output_dir: Union[str, Path], iteration: Union[str, int],
                   loose_files: bool = True) -> Path:
    """Writes what scripts/train.py writes (:197-208, :587-604): the checkpoint dict and, optionally, the six loose
    files (`iteration='final'` -> checkpoint_final.pt, as the final save does)."""
    d = Path(output_dir)
    d.mkdir(parents=True, exist_ok=True)
    cpu = {k: params[k].detach().to("cpu", torch.float32).contiguous() for k in PARAMS}
    if iteration == "final":
        path, suffix, it = d / "checkpoint_final.pt", "final", int(params.get("iteration", 0)) if "iteration" in params else 0
    else:
        path, suffix, it = d / f"checkpoint_{int(iteration):06d}.pt", str(int(iteration)), int(iteration)
    torch.save({"iteration": it, **cpu}, path)
    if loose_files:
        for k in PARAMS:
            torch.save(cpu[k], d / f"{_LOOSE_NAME[k]}_{suffix}.pt")
    return path


def load_cameras(data_dir: Union[str, Path], scale_factor: float = 1.0, device: Optional[str] = None) -> dict:
    """cam_meta.npy (+ poses.npy when present) of a data directory, with the intrinsics scaled the way the scripts
    do it (scripts/render_trained.py:194-202: H = int(height * s), fx *= s, cx defaults to width / 2).  Returns
    {'H','W','fx','fy','cx','cy'[, 'poses': float32 tensor [K,4,4] on `device`]}."""
    d = Path(data_dir)
    meta_path = d / "cam_meta.npy"
    if not meta_path.exists():
        raise FileNotFoundError(f"Camera metadata not found at {meta_path}")
    cam = np.load(meta_path, allow_pickle=True).item()
    Hs, Ws = cam["height"], cam["width"]
    out = {"H": int(Hs * scale_factor), "W": int(Ws * scale_factor), "fx": float(cam["fx"] * scale_factor),
           "fy": float(cam["fy"] * scale_factor), "cx": float(cam.get("cx", Ws / 2) * scale_factor),
           "cy": float(cam.get("cy", Hs / 2) * scale_factor)}
    poses_path = d / "poses.npy"
    if poses_path.exists():
        poses = torch.from_numpy(np.load(poses_path).astype(np.float32))
        if poses.ndim != 3 or poses.shape[1:] != (4, 4):
            raise ValueError(f"poses.npy: expected [K,4,4], got {tuple(poses.shape)}")
        out["poses"] = poses if device is None else _to_device(poses, device)
    return out
