"""Regression check for the open question of round 1: with TWO NVLS-multicast-mapped
peer areas in one process - the optimizer's and an all-reduce area - parameter updates were lost at 4 ranks.  Round 2
signals the barrier behind multimem stores through the multicast mapping itself (peer_barrier_mc_kernel); this tool
recreates the two-area situation on purpose (the library normally refuses a second multicast area) and checks, step by
step, the multicast all-reduce against NCCL and the multicast PeerAdam step against a local torch.optim.Adam replay.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29551 \
        tools/peer_two_areas_check.py [N] [steps]
"""
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "3d-gaussian-splatting-for-novel-view-synthesis_b200"))
import torch
import torch.distributed as dist
import b200gs
from b200gs import _lib, ops, peer

SHAPES = dict(pos=(3,), opacity_raw=(), f_dc=(3,), f_rest=(45,), scale_raw=(3,), q_raw=(4,))
LRS = dict(pos=1.6e-4, opacity_raw=0.05, f_dc=2.5e-3, f_rest=1.25e-4, scale_raw=5e-3, q_raw=1e-3)


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 500_000
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()
    g0 = torch.Generator().manual_seed(1)
    ref = {k: torch.nn.Parameter(torch.randn((n,) + s, generator=g0).to(dev)) for k, s in SHAPES.items()}
    mine = {k: torch.nn.Parameter(v.detach().clone()) for k, v in ref.items()}
    opt_ref = torch.optim.Adam([{"params": [ref[k]], "lr": LRS[k]} for k in SHAPES], lr=1e-3, eps=1e-15)
    opt = b200gs.PeerAdam([{"params": [mine[k]], "lr": LRS[k]} for k in SHAPES], lr=1e-3, eps=1e-15,
                          clip_params=[mine["pos"]], max_norm=1.0, multicast=True)
    # the second multicast-mapped area (normally refused)
    peer.PeerArea._multicast_in_use = False
    area2 = peer.PeerArea([mine[k].numel() for k in SHAPES], dev, multicast=True)
    both_mc = bool(opt.area.c_group.multicast) and bool(area2.c_group.multicast)
    report = {"allreduce_max_rel_err": 0.0, "adam_max_rel_err": 0.0, "adam_bad_elements": 0, "allreduce_bad_elements": 0}
    for step in range(1, steps + 1):
        total = {k: None for k in SHAPES}
        for r in range(world):
            g = torch.Generator(device=dev).manual_seed(1000 * step + r)
            for k, s in SHAPES.items():
                x = torch.randn((n,) + s, generator=g, device=dev) * 1e-3
                if r == rank:
                    mine[k].grad = x.clone()
                total[k] = x if total[k] is None else total[k] + x
        # (1) all-reduce of copies of the gradients through area 2 (multicast) against the sum every rank can rebuild
        copies = {k: mine[k].grad.clone() for k in SHAPES}
        table = (_lib.PeerTensor * len(SHAPES))()
        for i, k in enumerate(SHAPES):
            table[i] = _lib.PeerTensor(copies[k].data_ptr(), copies[k].numel(), 0.0, 1, 0)
        _lib.check(lib.b200gs_peer_allreduce(ctypes.byref(area2.c_group), ctypes.byref(area2.layout), table, len(SHAPES),
                                             ctypes.byref(area2.epoch), ops._stream(dev)), "peer_allreduce")
        torch.cuda.synchronize()
        for k in SHAPES:
            d = (copies[k] - total[k]).abs()
            report["allreduce_max_rel_err"] = max(report["allreduce_max_rel_err"], float(d.max() / total[k].abs().max()))
            report["allreduce_bad_elements"] += int((d > 2e-6 * total[k].abs().max()).sum())
        # (2) the multicast optimizer step against a local replay
        for k in SHAPES:
            ref[k].grad = total[k]
        torch.nn.utils.clip_grad_norm_(ref["pos"], max_norm=1.0)
        opt_ref.step()
        opt.step()
        torch.cuda.synchronize()
        dist.barrier()
        for k in SHAPES:
            a, b = mine[k].detach(), ref[k].detach()
            d = (a - b).abs()
            report["adam_max_rel_err"] = max(report["adam_max_rel_err"], float(d.max() / b.abs().max()))
            report["adam_bad_elements"] += int((d > 2e-6 * b.abs().max()).sum())
    out = [None] * world
    dist.all_gather_object(out, report)
    if rank == 0:
        print(json.dumps({"n": n, "world": world, "steps": steps, "both_areas_multicast": both_mc,
                          "allreduce_max_rel_err": max(r["allreduce_max_rel_err"] for r in out),
                          "allreduce_bad_elements": sum(r["allreduce_bad_elements"] for r in out),
                          "adam_max_rel_err": max(r["adam_max_rel_err"] for r in out),
                          "adam_bad_elements": sum(r["adam_bad_elements"] for r in out)}))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
