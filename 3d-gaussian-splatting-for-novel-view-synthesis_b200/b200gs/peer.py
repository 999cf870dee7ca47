"""Data-parallel optimizer step over NVLink peer memory (training row).

The data-parallel form of the reference's iteration (scripts/train.py:530-538) is

    total_loss.backward()  ->  SUM of the six gradient tensors over the ranks  ->
    clip_grad_norm_(model.pos, 1.0)  ->  optimizer.step()

`PeerAdam` does everything after `backward()` in ONE kernel per rank over peer-mapped memory: every rank
owns 1/world of every tensor, reads the owned gradient elements from every rank's staging buffer over NVLink
(reduce-scatter), clips, updates its shard of the Adam moments and stores the new parameter values into every
rank's parameter buffer (all-gather) - csrc/peer.cu.  The parameters are re-homed into a peer-visible buffer
(`p.data` becomes a view of it), so the render path reads them where the optimizer of any rank writes them.

`peer_allreduce_gradients` is the plain SUM all-reduce of `.grad` with the same machinery (no NCCL on the
data path), for callers that keep their own optimizer.

Plumbing only through torch: the peer-visible areas come from `torch.distributed._symmetric_memory` (CUDA VMM
allocations whose handles torch exchanges through the process group's store).  world = 1 (or no process
group) runs the same kernels on a local area.
There is no CPU or NCCL fallback on the data path.
"""
from __future__ import annotations

import ctypes
import os
from typing import Iterable, List, Optional, Sequence

import torch
import torch.distributed as dist

from . import _lib, ops


def slice_bounds(numel: int, world: int, rank: int):
    """[begin, end) of the elements of a tensor that `rank` owns (the rule of csrc/peer.cu, host arithmetic)."""
    lay = _layout([numel], world)
    per = int(lay.per[0])
    b = min(numel, rank * per)
    return b, min(numel, b + per)


def _layout(numels: Sequence[int], world: int) -> _lib.PeerLayout:
    lib = _lib.load()
    lay = _lib.PeerLayout()
    arr = (ctypes.c_int64 * max(1, len(numels)))(*[int(n) for n in numels])
    _lib.check(lib.b200gs_peer_layout_compute(arr, len(numels), int(world), ctypes.byref(lay)), "peer_layout_compute")
    return lay


class PeerArea:
    """One peer-visible area per rank, mapped into every process of the group."""
    _multicast_in_use = False

    def __init__(self, numels: Sequence[int], device: torch.device, group=None, transport: Optional[str] = None,
                 multicast: Optional[bool] = None):
        if device.type != "cuda":
            raise _lib.B200GSError("b200gs.peer: CUDA tensors only (no CPU fallback)")
        lib = _lib.load()
        self.device = device
        self.group = group
        multi = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
        self.world = dist.get_world_size(group) if multi else 1
        self.rank = dist.get_rank(group) if multi else 0
        if self.world > _lib.MAX_PEERS:
            raise _lib.B200GSError(f"b200gs.peer: at most {_lib.MAX_PEERS} ranks (one NVLink domain)")
        self.numels = [int(n) for n in numels]
        self.layout = _layout(self.numels, self.world)
        self.nbytes = int(lib.b200gs_peer_area_bytes(ctypes.byref(self.layout)))
        self._keep = []          # whatever keeps the peer mappings alive
        self.transport = "local"
        self.multicast_ptr = 0   # NVLS multicast mapping of all areas (0: none)
        if multicast is None:
            # Through the switch a rank's own copy crosses its links too: per GPU and direction (1 + 1/p) S bytes against
            # 2 (p-1)/p S for plain peer loads / stores.  Measured (tools/peer_bench.py, N = 1M): 2 ranks 0.84 vs 0.58 ms,
            # 4 ranks 0.79 vs 0.80 ms, 8 ranks 0.85 vs 1.06 ms.  OPT-IN at every world size all the same: with two
            # multimem-mapped areas in one process parameter updates were lost at 4 ranks and the
            # cause is not established, so the default is the path whose ordering argument is complete - plain peer
            # loads / stores, ordered by release/acquire flag barriers.
            multicast = os.environ.get("B200GS_PEER_MULTICAST") == "1"
        if multicast and PeerArea._multicast_in_use:
            # one multicast-mapped area per process: with a second one (the optimizer's and an all-reduce area, both
            # through multimem) parameter updates were lost at 4 ranks; not understood, so not allowed
            multicast = False
        if not multi:
            self.buf = torch.zeros(self.nbytes, dtype=torch.uint8, device=device)
            ptrs = [self.buf.data_ptr()]
        else:
            if transport not in (None, "symm_mem"):
                raise ValueError("b200gs.peer: the only transport is 'symm_mem'")
            try:
                ptrs = self._map_symm_mem()
            except Exception as e:              # noqa: BLE001 - re-raised with the context a user needs
                raise _lib.B200GSError(f"b200gs.peer: could not map the peers' memory ({type(e).__name__}: {e}); the "
                                       "ranks must sit on one NVLink/P2P domain") from e
            self.transport = "symm_mem"
        self.c_group = _lib.PeerGroup(world=self.world, rank=self.rank)
        for q, p in enumerate(ptrs):
            self.c_group.area[q] = p
        self.c_group.multicast = (self.multicast_ptr or None) if multicast else None
        if self.c_group.multicast:
            PeerArea._multicast_in_use = True
        self.epoch = ctypes.c_uint32(0)
        n = int(self.layout.flat_total)
        ctrl = _lib.PEER_CTRL_BYTES
        self.flat_params = self.buf[ctrl:ctrl + 4 * n].view(torch.float32)
        self.flat_grads = self.buf[ctrl + 4 * n:ctrl + 8 * n].view(torch.float32)

    # -- transports ------------------------------------------------------------------------------------------
    def _map_symm_mem(self) -> List[int]:
        import torch.distributed._symmetric_memory as symm_mem
        pg = self.group if self.group is not None else dist.group.WORLD
        with torch.cuda.device(self.device):
            buf = symm_mem.empty(self.nbytes, dtype=torch.uint8, device=self.device)
            hdl = symm_mem.rendezvous(buf, pg)
            buf.zero_()
            torch.cuda.synchronize(self.device)
        dist.barrier(group=self.group)
        self.buf = buf
        self._keep.append(hdl)
        try:
            self.multicast_ptr = int(hdl.multicast_ptr) if hdl.has_multicast_support else 0
        except Exception:          # noqa: BLE001 - no multicast query: plain peer loads and stores
            self.multicast_ptr = 0
        return [int(p) for p in hdl.buffer_ptrs]

    # -- helpers ---------------------------------------------------------------------------------------------
    def view(self, flat: torch.Tensor, index: int, shape) -> torch.Tensor:
        off = int(self.layout.offset[index])
        return flat[off:off + self.numels[index]].view(shape)

    def barrier(self):
        lib = _lib.load()
        with torch.cuda.device(self.device):
            _lib.check(lib.b200gs_peer_barrier(ctypes.byref(self.c_group), ctypes.byref(self.epoch),
                                               ops._stream(self.device)), "peer_barrier")


def _check_param(p: torch.Tensor, who: str):
    if not p.is_cuda or p.dtype != torch.float32 or not p.is_contiguous():
        raise _lib.B200GSError(f"{who}: parameters must be contiguous CUDA fp32 tensors (no CPU / mixed-precision fallback)")


class PeerAdam(torch.optim.Optimizer):
    """`torch.optim.Adam` (weight_decay = 0, amsgrad = False) for one-process-per-GPU data-parallel training,
    fused with the gradient reduction over the ranks and with `clip_grad_norm_`.

        opt = b200gs.PeerAdam([{'params': [model.pos], 'lr': ...}, ...], eps=1e-15,
                              clip_params=[model.pos], max_norm=1.0)
        loss.backward(); opt.step()           # no all-reduce, no clip_grad_norm_ call: step() does both

    * every rank must construct it with the same parameter shapes, in the same order, and call `step()` the same
      number of times; rank 0's parameter values are broadcast at construction;
    * `p.data` of every parameter is re-homed into the peer-visible buffer (same values, new storage);
    * the local `.grad` of a parameter is this rank's contribution (None = zeros); with `write_grads=True` it
      holds the reduced (and clipped) gradient after `step()`, as it would after all-reduce + clip;
    * gradients that come out of `b200gs.render`'s backward are written straight into the peer-visible staging
      buffer (`.grad` is then a view of it) - no staging copy; any other gradient tensor is copied there by `step()`;
    * the moments are sharded over the ranks (each rank keeps 1/world of `exp_avg` / `exp_avg_sq`);
    * opt-in (`multicast=True` / `B200GS_PEER_MULTICAST=1`): the areas are also mapped through an NVLS multicast
      address and the kernel uses `multimem.ld_reduce` / `multimem.st` - the switch sums the gradients and replicates
      the parameters, so the bytes a GPU moves stop growing with the number of ranks (0.85 vs 1.06 ms at 8 ranks).
      Off by default at every world size until the lost updates seen with two multicast areas are explained.
    """

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0, amsgrad=False, *, group=None,
                 clip_params: Optional[Iterable[torch.Tensor]] = None, max_norm: float = 0.0, write_grads: bool = False,
                 transport: Optional[str] = None, multicast: Optional[bool] = None):
        if weight_decay != 0 or amsgrad:
            raise NotImplementedError("b200gs.PeerAdam implements the reference's configuration: weight_decay=0, amsgrad=False")
        if not 0.0 <= lr or not 0.0 <= eps or not 0.0 <= betas[0] < 1.0 or not 0.0 <= betas[1] < 1.0:
            raise ValueError("invalid Adam hyper-parameter")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=0, amsgrad=False))
        self._plist = [p for g in self.param_groups for p in g["params"]]
        if not self._plist:
            raise ValueError("PeerAdam: no parameters")
        if len(self._plist) > _lib.PEER_MAX_TENSORS:
            raise NotImplementedError(f"b200gs.PeerAdam handles up to {_lib.PEER_MAX_TENSORS} tensors (the reference has six)")
        if len({(tuple(g["betas"]), float(g["eps"])) for g in self.param_groups}) != 1:
            raise NotImplementedError("b200gs.PeerAdam: betas and eps must be the same for every group")
        for p in self._plist:
            _check_param(p, "b200gs.PeerAdam")
        dev = self._plist[0].device
        self.area = PeerArea([p.numel() for p in self._plist], dev, group=group, transport=transport, multicast=multicast)
        clip_ids = {id(p) for p in (clip_params or [])}
        self._clip = [1 if id(p) in clip_ids else 0 for p in self._plist]
        self.max_norm = float(max_norm) if clip_ids else 0.0
        self.write_grads = bool(write_grads)
        # re-home the parameters
        with torch.no_grad():
            for i, p in enumerate(self._plist):
                v = self.area.view(self.area.flat_params, i, p.shape)
                v.copy_(p.data)
                p.data = v
            if self.area.world > 1:
                dist.broadcast(self.area.flat_params, src=dist.get_global_rank(group, 0) if group is not None else 0,
                               group=group)
                torch.cuda.synchronize(dev)
                dist.barrier(group=group)
        # the render backward writes these leaves' gradients straight into the staging buffer (ops.register_grad_sink)
        for i, p in enumerate(self._plist):
            ops.register_grad_sink(p, self.area.view(self.area.flat_grads, i, p.shape))
        n_shard = max(1, int(self.area.layout.shard_total))
        self.exp_avg = torch.zeros(n_shard, dtype=torch.float32, device=dev)
        self.exp_avg_sq = torch.zeros(n_shard, dtype=torch.float32, device=dev)
        self.total_norm = torch.zeros((), dtype=torch.float32, device=dev)
        self._steps = 0

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        lib = _lib.load()
        self._steps += 1
        lr_of = {id(p): float(g["lr"]) for g in self.param_groups for p in g["params"]}
        table = (_lib.PeerTensor * len(self._plist))()
        keep = []
        for i, p in enumerate(self._plist):
            g = p.grad
            if g is not None:
                if not g.is_cuda or g.dtype != torch.float32 or g.is_sparse:
                    raise _lib.B200GSError("b200gs.PeerAdam: gradients must be dense CUDA fp32 tensors")
                if not g.is_contiguous():
                    g = g.contiguous()
                    if self.write_grads:
                        p.grad = g
                keep.append(g)
            elif self.write_grads:
                g = p.grad = torch.zeros_like(p)
                keep.append(g)
            table[i] = _lib.PeerTensor(g.data_ptr() if g is not None else None, p.numel(), lr_of[id(p)], self._steps,
                                       self._clip[i])
        grp = self.param_groups[0]
        dev = self.area.device
        with torch.cuda.device(dev):
            _lib.check(lib.b200gs_peer_adam_step(ctypes.byref(self.area.c_group), ctypes.byref(self.area.layout), table,
                                                 len(self._plist), ops._ptr(self.exp_avg), ops._ptr(self.exp_avg_sq),
                                                 grp["betas"][0], grp["betas"][1], grp["eps"], self.max_norm,
                                                 1 if self.write_grads else 0, ctypes.byref(self.area.epoch),
                                                 ops._ptr(self.total_norm), ops._stream(dev)), "peer_adam_step")
        ops.release_grad_sinks(self._plist)
        return loss

    def __del__(self):
        try:
            ops.unregister_grad_sinks(self._plist)
        except Exception:          # noqa: BLE001 - interpreter shutdown
            pass


_allreduce_areas = {}


def peer_allreduce_gradients(params: Iterable[torch.Tensor], group=None) -> None:
    """SUM all-reduce of `.grad` of every parameter over peer memory, in place (a parameter without a gradient
    contributes zeros and receives the sum).  Same contract as `b200gs.dist.allreduce_gradients`, no NCCL on
    the data path.  Every rank must call it with the same parameter shapes in the same order."""
    plist = list(params)
    if not plist:
        return
    if len(plist) > _lib.PEER_MAX_TENSORS:
        raise NotImplementedError(f"peer_allreduce_gradients handles up to {_lib.PEER_MAX_TENSORS} tensors per call")
    for p in plist:
        _check_param(p, "b200gs.peer_allreduce_gradients")
    dev = plist[0].device
    key = (dev.index, id(group), tuple(p.numel() for p in plist))
    area = _allreduce_areas.get(key)
    if area is None:
        area = _allreduce_areas[key] = PeerArea([p.numel() for p in plist], dev, group=group, multicast=False)
    lib = _lib.load()
    table = (_lib.PeerTensor * len(plist))()
    for i, p in enumerate(plist):
        if p.grad is None:
            p.grad = torch.zeros_like(p)
        elif not p.grad.is_contiguous():
            p.grad = p.grad.contiguous()
        if not p.grad.is_cuda or p.grad.dtype != torch.float32:
            raise _lib.B200GSError("b200gs.peer_allreduce_gradients: gradients must be dense CUDA fp32 tensors")
        table[i] = _lib.PeerTensor(p.grad.data_ptr(), p.numel(), 0.0, 1, 0)
    with torch.cuda.device(dev):
        _lib.check(lib.b200gs_peer_allreduce(ctypes.byref(area.c_group), ctypes.byref(area.layout), table, len(plist),
                                             ctypes.byref(area.epoch), ops._stream(dev)), "peer_allreduce")
