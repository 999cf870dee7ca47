"""Diagnostic: b200gs.PeerAdam against a local torch.optim.Adam replay on every rank, step by step, with per-tensor
mismatch counts and the owners of the mismatching elements.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29531 \
        tools/peer_step_check.py [N] [steps] [multicast 0|1]
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "3d-gaussian-splatting-for-novel-view-synthesis_b200"))
import torch
import torch.distributed as dist
import b200gs
from b200gs import peer

SHAPES = dict(pos=(3,), opacity_raw=(), f_dc=(3,), f_rest=(45,), scale_raw=(3,), q_raw=(4,))
LRS = dict(pos=1.6e-4, opacity_raw=0.05, f_dc=2.5e-3, f_rest=1.25e-4, scale_raw=5e-3, q_raw=1e-3)


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    multicast = (sys.argv[3] == "1") if len(sys.argv) > 3 else None
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    g0 = torch.Generator().manual_seed(1)
    ref = {k: torch.nn.Parameter(torch.randn((n,) + s, generator=g0).to(dev)) for k, s in SHAPES.items()}
    mine = {k: torch.nn.Parameter(v.detach().clone()) for k, v in ref.items()}
    opt_ref = torch.optim.Adam([{"params": [ref[k]], "lr": LRS[k]} for k in SHAPES], lr=1e-3, eps=1e-15)
    opt = b200gs.PeerAdam([{"params": [mine[k]], "lr": LRS[k]} for k in SHAPES], lr=1e-3, eps=1e-15,
                          clip_params=[mine["pos"]], max_norm=1.0, multicast=multicast)
    report, worst = [], {}
    for step in range(1, steps + 1):
        total = {k: None for k in SHAPES}
        for r in range(world):                      # every rank can rebuild every rank's gradient
            seed = 7 if os.environ.get("SAME_GRADS") == "1" else 1000 * step + r     # SAME_GRADS: one gradient for all ranks and steps
            g = torch.Generator(device=dev).manual_seed(seed)                        # same stream of numbers on every GPU
            for k, s in SHAPES.items():
                x = torch.randn((n,) + s, generator=g, device=dev) * 1e-3
                if r == rank:
                    mine[k].grad = x.clone()
                total[k] = x if total[k] is None else total[k] + x      # rank order, like the plain peer path
        for k in SHAPES:
            ref[k].grad = total[k]
        torch.nn.utils.clip_grad_norm_(ref["pos"], max_norm=1.0)
        opt_ref.step()
        opt.step()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        row = {"step": step}
        for k in SHAPES:
            a, b = mine[k].detach().flatten(), ref[k].detach().flatten()
            err = float((a - b).abs().max() / b.abs().max())
            worst[k] = max(worst.get(k, 0.0), err)                      # max-norm relative error over all steps
            bad = ((a - b).abs() > 2e-6 * b.abs().max()).nonzero().flatten()
            if bad.numel():
                per = peer.slice_bounds(a.numel(), world, 0)[1]
                owners = torch.bincount(torch.clamp(bad // max(per, 1), max=world - 1), minlength=world).tolist()
                row[k] = {"bad": int(bad.numel()), "owners": owners, "first": int(bad[0]), "last": int(bad[-1]),
                          "max_abs": float((a - b).abs().max())}
        report.append(row)
    # replicas: every rank must hold bit-identical parameters after the last step
    digest = torch.stack([mine[k].detach().double().sum() for k in SHAPES] +
                         [mine[k].detach().view(torch.int32).long().sum().double() for k in SHAPES])
    out, worsts, digests = [None] * world, [None] * world, [None] * world
    if world > 1:
        dist.all_gather_object(out, report)
        dist.all_gather_object(worsts, worst)
        dist.all_gather_object(digests, digest.cpu().tolist())
    else:
        out, worsts, digests = [report], [worst], [digest.cpu().tolist()]
    if rank == 0:
        bad = [[row for row in rep if len(row) > 1] for rep in out]
        print(json.dumps({"n": n, "world": world, "steps": steps, "multicast": bool(opt.area.c_group.multicast),
                          "steps_with_mismatch_per_rank": [len(b) for b in bad],
                          "max_rel_err_vs_local_torch_adam_replay": {k: max(w[k] for w in worsts) for k in SHAPES},
                          "tolerance": 2e-6, "replicas_bit_identical": all(d == digests[0] for d in digests),
                          "first_mismatches": [b[:2] for b in bad]}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
