"""Adaptive density control: the reference's `GaussianModel.densify_and_prune`
(scripts/train.py:89-195) as two kernels' worth of stream compaction (csrc/densify.cu).

    b200gs.densify_and_prune(model, grads, opacity_threshold=0.01, max_grad=0.01, scale_threshold=0.01,
                             max_screen_size=20)

has the reference method's signature (with the model as first argument, so `GaussianModel.densify_and_prune =
b200gs.densify_and_prune` is the whole patch a maintainer needs - the class lives in the training script itself, which
is why `install()` cannot rebind it) and its effects: the six `nn.Parameter`s of `model` are replaced by new ones, the
entries of the `grads` dict are pruned in place, split copies are displaced with `torch.randn_like` noise drawn exactly
where the reference draws it (same generator state afterwards).

One behaviour of the reference is reproduced on purpose: when a split and a clone happen in the same call it raises
(`_clone_points` indexes the already grown tensors with the pre-split mask, scripts/train.py:139-146,187-195).
`strict=False` applies the evident intent instead (clones are taken from the pre-split rows).
"""
from __future__ import annotations

import ctypes
from typing import Dict, Optional

import torch

from . import _lib, ops

PARAMS = ("pos", "opacity_raw", "f_dc", "f_rest", "scale_raw", "q_raw")      # order of csrc/densify.cu
_WIDTH = {"pos": (3,), "opacity_raw": (), "f_dc": (3,), "f_rest": (45,), "scale_raw": (3,), "q_raw": (4,)}


def densify_tensors(params: Dict[str, torch.Tensor], pos_grad: Optional[torch.Tensor], opacity_threshold=0.01,
                    max_grad=0.01, scale_threshold=0.01, strict: bool = True, noise: Optional[torch.Tensor] = None):
    """The tensor-level operation.  params: the six tensors [N, ...] (CUDA fp32); pos_grad: [N,3] or None (prune
    only).  Returns (new tensors dict, info) with info = {'keep': bool[N] mask of surviving rows, 'n_keep', 'n_split',
    'n_clone'}.  `noise` ([n_split,3]) replaces the torch.randn_like draw (tests)."""
    lib = _lib.load()
    src = {}
    for k in PARAMS:
        t = params[k].detach()
        if not t.is_cuda:
            raise _lib.B200GSError("b200gs.densify: CUDA tensors only (no CPU fallback)")
        src[k] = ops._f32c(t)
    n = src["pos"].shape[0]
    dev = src["pos"].device
    for k in PARAMS:
        if tuple(src[k].shape) != (n,) + _WIDTH[k]:
            raise ValueError(f"{k}: expected shape {(n,) + _WIDTH[k]}, got {tuple(src[k].shape)}")
    g = None if pos_grad is None else ops._f32c(pos_grad.detach())
    if g is not None and tuple(g.shape) != (n, 3):
        raise ValueError("pos gradient must be [N,3]")
    nbytes = int(lib.b200gs_densify_workspace_bytes(n))
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    counts = torch.zeros(3, dtype=torch.int32).pin_memory()
    with torch.cuda.device(dev):
        stream = ops._stream(dev)
        _lib.check(lib.b200gs_densify_plan(n, ops._ptr(src["opacity_raw"]), ops._ptr(src["scale_raw"]), ops._ptr(g),
                                           float(opacity_threshold), float(max_grad), float(scale_threshold), ops._ptr(ws),
                                           nbytes, ctypes.c_void_p(counts.data_ptr()), stream), "densify_plan")
        torch.cuda.current_stream(dev).synchronize()          # the one host sync: the new row count sizes the outputs
        n_keep, n_split, n_clone = (int(c) for c in counts)
        if strict and n_split > 0 and n_clone > 0:
            # scripts/train.py:187-195: self.pos[mask] with mask of the pre-split length on the grown tensor
            raise IndexError(f"The shape of the mask [{n_keep}] at index 0 does not match the shape of the indexed tensor "
                             f"[{n_keep + n_split}, 3] at index 0")
        if noise is None:
            # drawn where the reference draws it: torch.randn_like(new_pos) inside _split_points (train.py:164), only if
            # anything is split
            noise = torch.randn((n_split, 3), dtype=torch.float32, device=dev) if n_split > 0 else None
        elif n_split > 0:
            noise = ops._f32c(noise.to(dev))
            if tuple(noise.shape) != (n_split, 3):
                raise ValueError(f"noise must be [{n_split},3]")
        n_out = n_keep + n_split + n_clone
        out = {k: torch.empty((n_out,) + _WIDTH[k], dtype=torch.float32, device=dev) for k in PARAMS}
        in6 = (ctypes.c_void_p * 6)(*[src[k].data_ptr() for k in PARAMS])
        out6 = (ctypes.c_void_p * 6)(*[out[k].data_ptr() for k in PARAMS])
        _lib.check(lib.b200gs_densify_apply(n, ops._ptr(ws), nbytes, in6, out6, ops._ptr(noise) if n_split > 0 else None,
                                            stream), "densify_apply")
        keep = ws[256:256 + 4 * n].view(torch.int32).ne(0) if n > 0 else torch.zeros(0, dtype=torch.bool, device=dev)
    return out, {"keep": keep, "n_keep": n_keep, "n_split": n_split, "n_clone": n_clone}


def densify_and_prune(model, grads, opacity_threshold=0.01, max_grad=0.01, scale_threshold=0.01, max_screen_size=20,
                      strict: bool = True):
    """`GaussianModel.densify_and_prune` (scripts/train.py:89-141): replaces model.pos / opacity_raw / f_dc / f_rest /
    scale_raw / q_raw by new `nn.Parameter`s and prunes the entries of `grads` (a dict of gradient tensors or None)."""
    params = {k: getattr(model, k) for k in PARAMS}
    pos_grad = None if grads is None else grads.get("pos")
    out, info = densify_tensors(params, pos_grad, opacity_threshold, max_grad, scale_threshold, strict=strict)
    if grads is not None:                                       # train.py:122-126
        for key in grads:
            if grads[key] is not None:
                grads[key] = grads[key][info["keep"]]
    for k in PARAMS:
        setattr(model, k, torch.nn.Parameter(out[k]))
    return info
