"""Optimizer half of a data-parallel training iteration at the headline size (N = 1M Gaussians, 59 floats each):

    a) NCCL SUM all-reduce of the six gradient tensors + b200gs.clip_grad_norm_(pos) + b200gs.FusedAdam.step()
    b) b200gs.PeerAdam.step()  (reduce-scatter + clip + Adam + all-gather in one kernel over NVLink peer memory)
    c) the all-reduce alone: NCCL vs b200gs.peer_allreduce_gradients

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tools/peer_bench.py [N] [multicast 0|1]

CUDA events on the launching stream, max over ranks, one JSON line from rank 0.
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "3d-gaussian-splatting-for-novel-view-synthesis_b200"))
import torch
import torch.distributed as dist
import b200gs
from b200gs.dist import allreduce_gradients

SHAPES = dict(pos=(3,), opacity_raw=(), f_dc=(3,), f_rest=(45,), scale_raw=(3,), q_raw=(4,))
LRS = dict(pos=1.6e-6, opacity_raw=0.05, f_dc=2.5e-3, f_rest=1.25e-4, scale_raw=5e-3, q_raw=1e-3)


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
    multicast = (sys.argv[2] == "1") if len(sys.argv) > 2 else None
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def make():
        g = torch.Generator().manual_seed(1)
        return {k: torch.nn.Parameter(torch.randn((n,) + s, generator=g).to(dev)) for k, s in SHAPES.items()}
    pa, pb = make(), make()
    grads = {k: torch.randn_like(p) * 1e-3 for k, p in pa.items()}
    opt_a = b200gs.FusedAdam([{"params": [pa[k]], "lr": LRS[k]} for k in SHAPES], lr=1e-3, eps=1e-15)
    opt_b = b200gs.PeerAdam([{"params": [pb[k]], "lr": LRS[k]} for k in SHAPES], lr=1e-3, eps=1e-15,
                            clip_params=[pb["pos"]], max_norm=1.0, multicast=multicast)
    opt_bw = None

    def set_grads(ps):
        for k, p in ps.items():
            p.grad = grads[k].clone()

    def step_a():
        allreduce_gradients(pa.values())
        b200gs.clip_grad_norm_(pa["pos"], max_norm=1.0)
        opt_a.step()

    def step_b():
        opt_b.step()

    def ar_nccl():
        allreduce_gradients(pa.values())

    def ar_peer():
        b200gs.peer_allreduce_gradients(pa.values())

    def timed(fn, ps, iters=20, warm=5):
        for _ in range(warm):
            set_grads(ps)
            fn()
        ms = []
        for _ in range(iters):
            set_grads(ps)
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            torch.cuda.synchronize()
            ms.append(a.elapsed_time(b))
        t = torch.tensor(sorted(ms)[len(ms) // 2], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)

    res = {"n": n, "world": world, "transport": opt_b.area.transport, "multicast": bool(opt_b.area.c_group.multicast), "bytes_per_rank": 59 * n * 4}
    res["nccl_allreduce_clip_fusedadam_ms"] = timed(step_a, pa)
    res["peer_adam_ms"] = timed(step_b, pb)
    res["nccl_allreduce_ms"] = timed(ar_nccl, pa)
    res["peer_allreduce_ms"] = timed(ar_peer, pa)
    # parameters after the same number of steps on the same gradients must agree between the two routes
    res["param_rel_diff"] = max(float((pa[k] - pb[k]).abs().max() / pa[k].abs().max()) for k in SHAPES)
    if rank == 0:
        print(json.dumps(res), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
