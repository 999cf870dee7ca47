"""Data-parallel optimizer step over peer memory (b200gs.PeerAdam, b200gs.peer_allreduce_gradients; csrc/peer.cu).

Reference behaviour: scripts/train.py:530-538 per rank + a SUM of the gradients over the ranks, i.e.
all-reduce -> torch.nn.utils.clip_grad_norm_(pos, 1.0) -> torch.optim.Adam(six groups, eps=1e-15).step().
The checker is torch's own all-reduce / clip / Adam (third-party code of the reference, requirements.txt:1).

  not gpu : the slice arithmetic (host code of the library) and the no-fallback contract
  gpu     : world = 1 runs the same kernels on a local area: parity with torch.optim.Adam + clip_grad_norm_
  gpu x2  : two ranks over NVLink peer memory against NCCL all-reduce + torch clip + torch Adam
"""
import os
import socket

import pytest
import torch

SHAPES = dict(pos=(3,), opacity_raw=(), f_dc=(3,), f_rest=(45,), scale_raw=(3,), q_raw=(4,))
LRS = dict(pos=1.6e-4, opacity_raw=0.05, f_dc=2.5e-3, f_rest=1.25e-4, scale_raw=5e-3, q_raw=1e-3)


def test_slices_partition_every_tensor():
    from b200gs import peer
    for numel in (0, 1, 3, 4, 5, 4095, 4096, 4097, 100_003, 45_000_003):
        for world in (1, 2, 3, 4, 8, 16):
            prev = 0
            for r in range(world):
                b, e = peer.slice_bounds(numel, world, r)
                assert b == prev and b <= e <= numel and b % 4 == 0 or b == numel
                prev = e
            assert prev == numel
    lay = peer._layout([3_000_000, 1_000_000, 45_000_001], 3)
    assert list(lay.offset[:3]) == [0, 3_000_000, 4_000_000] and lay.flat_total == 4_000_000 + 45_000_032
    assert all(o % 32 == 0 for o in lay.shard_offset[:3]) and lay.shard_total >= sum(lay.per[:3])


def test_peer_layout_rejects_bad_arguments():
    from b200gs import _lib, peer
    with pytest.raises(_lib.B200GSError):
        peer._layout([1] * 9, 2)
    with pytest.raises(_lib.B200GSError):
        peer._layout([1], 17)
    with pytest.raises(_lib.B200GSError):
        peer._layout([-1], 2)


def test_peer_has_no_cpu_fallback():
    import b200gs
    p = torch.nn.Parameter(torch.zeros(8, 3))
    with pytest.raises(b200gs.B200GSError):
        b200gs.PeerAdam([{"params": [p]}], lr=1e-3)
    p.grad = torch.ones_like(p)
    with pytest.raises(b200gs.B200GSError):
        b200gs.peer_allreduce_gradients([p])


def _make_params(n, dev, seed):
    g = torch.Generator().manual_seed(seed)
    return {k: torch.nn.Parameter(torch.randn((n,) + s, generator=g).to(dev)) for k, s in SHAPES.items()}


def _make_grads(n, step, rank, scale):
    g = torch.Generator().manual_seed(1000 * step + rank)
    # six decades of dynamic range, like real per-Gaussian gradients
    return {k: (torch.randn((n,) + s, generator=g) * torch.exp(torch.randn((n,) + s, generator=g) * 2.0) * scale)
            for k, s in SHAPES.items()}


def _rel(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


@pytest.mark.gpu
@pytest.mark.parametrize("n,scale", [(1, 1.0), (1023, 1e-6), (4099, 1.0), (100_003, 0.05)])
def test_peer_adam_world1_matches_torch_adam_and_clip(n, scale):
    import b200gs
    dev = torch.device("cuda", 0)
    ref = _make_params(n, dev, 3)
    mine = {k: torch.nn.Parameter(v.detach().clone()) for k, v in ref.items()}
    opt_ref = torch.optim.Adam([{"params": [ref[k]], "lr": LRS[k]} for k in SHAPES], lr=1e-3, eps=1e-15)
    opt = b200gs.PeerAdam([{"params": [mine[k]], "lr": LRS[k]} for k in SHAPES], lr=1e-3, eps=1e-15,
                          clip_params=[mine["pos"]], max_norm=1.0, write_grads=True)
    assert opt.area.transport == "local" and opt.area.world == 1
    for step in range(1, 6):
        grads = _make_grads(n, step, 0, scale)
        for k in SHAPES:
            ref[k].grad = grads[k].to(dev)
            mine[k].grad = grads[k].to(dev)
        norm_ref = torch.nn.utils.clip_grad_norm_(ref["pos"], max_norm=1.0)
        opt_ref.step()
        opt.step()
        assert abs(float(opt.total_norm) - float(norm_ref)) <= 2e-6 * float(norm_ref) + 1e-12
        assert _rel(mine["pos"].grad, ref["pos"].grad) <= 2e-6          # write_grads: the clipped gradient
    for k in SHAPES:
        assert _rel(mine[k].detach(), ref[k].detach()) <= 2e-6, k
        assert mine[k].data_ptr() >= opt.area.flat_params.data_ptr()     # re-homed into the area
    # moments: world = 1 -> the shard is the whole tensor
    lay = opt.area.layout
    for i, k in enumerate(SHAPES):
        so, ne = int(lay.shard_offset[i]), mine[k].numel()
        st = opt_ref.state[ref[k]]
        assert _rel(opt.exp_avg[so:so + ne], st["exp_avg"].flatten()) <= 2e-6
        assert _rel(opt.exp_avg_sq[so:so + ne], st["exp_avg_sq"].flatten()) <= 4e-6


@pytest.mark.gpu
def test_peer_allreduce_world1_is_identity_and_missing_grad_is_zero():
    import b200gs
    dev = torch.device("cuda", 0)
    ps = _make_params(5001, dev, 9)
    keep = {}
    for k, p in ps.items():
        if k != "q_raw":
            p.grad = torch.randn_like(p)
            keep[k] = p.grad.clone()
    b200gs.peer_allreduce_gradients(ps.values())
    torch.cuda.synchronize()
    for k, p in ps.items():
        assert torch.equal(p.grad, keep[k]) if k in keep else float(p.grad.abs().max()) == 0.0


# ---- two ranks ------------------------------------------------------------------------------------------------
def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out, n, multicast):
    import torch.distributed as dist
    import b200gs
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        ref = _make_params(n, dev, 3)
        mine = {k: torch.nn.Parameter(v.detach().clone()) for k, v in ref.items()}
        opt_ref = torch.optim.Adam([{"params": [ref[k]], "lr": LRS[k]} for k in SHAPES], lr=1e-3, eps=1e-15)
        opt = b200gs.PeerAdam([{"params": [mine[k]], "lr": LRS[k]} for k in SHAPES], lr=1e-3, eps=1e-15,
                              clip_params=[mine["pos"]], max_norm=1.0, write_grads=True, multicast=multicast)
        out[f"transport{rank}"] = opt.area.transport
        out[f"multicast{rank}"] = bool(opt.area.c_group.multicast)
        worst = 0.0
        for step in range(1, 5):
            grads = _make_grads(n, step, rank, 0.02 if step % 2 else 1e-6)     # clipping active / inactive
            for k in SHAPES:
                ref[k].grad = grads[k].to(dev)
                mine[k].grad = None if (k == "q_raw" and rank == 1 and step == 2) else grads[k].to(dev)
            if step == 2 and rank == 1:
                ref["q_raw"].grad.zero_()
            for k in SHAPES:
                dist.all_reduce(ref[k].grad)
            torch.nn.utils.clip_grad_norm_(ref["pos"], max_norm=1.0)
            opt_ref.step()
            opt.step()
            worst = max(worst, _rel(mine["pos"].grad, ref["pos"].grad), _rel(mine["f_rest"].grad, ref["f_rest"].grad))
        torch.cuda.synchronize()
        out[f"grad_err{rank}"] = worst
        out[f"param_err{rank}"] = {k: _rel(mine[k].detach(), ref[k].detach()) for k in SHAPES}
        out[f"params{rank}"] = {k: mine[k].detach().cpu() for k in SHAPES}
        # plain all-reduce over peer memory against NCCL
        a = _make_params(n, dev, 50 + rank)
        for k, p in a.items():
            p.grad = torch.randn_like(p)
        want = {k: p.grad.clone() for k, p in a.items()}
        for k in SHAPES:
            dist.all_reduce(want[k])
        for _ in range(2):                    # twice: the second call reuses the area
            for k, p in a.items():
                p.grad = torch.randn(p.shape, generator=torch.Generator().manual_seed(7 + rank)).to(dev)
            want = {k: p.grad.clone() for k, p in a.items()}
            for k in SHAPES:
                dist.all_reduce(want[k])
            b200gs.peer_allreduce_gradients(a.values())
        torch.cuda.synchronize()
        out[f"allreduce_err{rank}"] = max(_rel(a[k].grad, want[k]) for k in SHAPES)
        out[f"allreduce{rank}"] = {k: a[k].grad.cpu() for k in SHAPES}
    finally:
        dist.destroy_process_group()


@pytest.mark.gpu
@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("multicast", [False, True])     # plain peer loads / stores; NVLS multimem (if the box has it)
def test_two_rank_peer_adam_matches_nccl_allreduce_plus_torch_adam(multicast):
    import torch.multiprocessing as mp
    world, port, n = 2, _free_port(), 50_001
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(world, port, out, n, multicast), nprocs=world, join=True)
        out = dict(out)
    for r in range(world):
        assert out[f"grad_err{r}"] <= 2e-6, out[f"grad_err{r}"]
        for k, e in out[f"param_err{r}"].items():
            assert e <= 2e-6, (r, k, e)
        assert out[f"allreduce_err{r}"] <= 1e-6
    for k in SHAPES:                          # the replicas stay bit-identical
        assert torch.equal(out["params0"][k], out["params1"][k]), k
        assert torch.equal(out["allreduce0"][k], out["allreduce1"][k]), k


@pytest.mark.gpu
def test_render_backward_writes_gradients_into_the_staging_buffer():
    """Gradient sinks (b200gs/ops.py): with a PeerAdam constructed over the leaves, the backward of b200gs.render writes
    the leaf gradients straight into the peer-visible staging buffer and autograd adopts those tensors as `.grad`
    (no staging copy in step()); values, accumulation over several views and the optimizer step stay what they were."""
    import b200gs
    from oracle import gs_oracle as O
    dev = torch.device("cuda", 0)
    names = ("pos", "opacity_raw", "f_dc", "f_rest", "scale_raw", "q_raw")
    sc = O.make_scene(6000, seed=8, log_scale=-3.4)
    cams = [O.make_camera(200, 136, view=v, n_views=4) for v in range(2)]
    w = [torch.rand(136, 200, 3, generator=torch.Generator().manual_seed(20 + v)).to(dev) for v in range(2)]

    def loss_of(leaves, v):
        cam = cams[v]
        c2w = cam["c2w"].to(dev)
        sg = b200gs.build_sigma_from_params(leaves["scale_raw"], leaves["q_raw"])
        col = b200gs.evaluate_sh(leaves["f_dc"], leaves["f_rest"], leaves["pos"], c2w)
        img = b200gs.render(leaves["pos"], col, leaves["opacity_raw"], sg, c2w, cam["H"], cam["W"], cam["fx"], cam["fy"],
                            cam["cx"], cam["cy"])
        return (img * w[v]).sum()

    plain = {k: torch.nn.Parameter(sc[k].to(dev)) for k in names}
    mine = {k: torch.nn.Parameter(sc[k].to(dev)) for k in names}
    opt = b200gs.PeerAdam([{"params": [mine[k]], "lr": LRS[k]} for k in names], lr=1e-3, eps=1e-15,
                          clip_params=[mine["pos"]], max_norm=1.0)
    sinks = {k: opt.area.view(opt.area.flat_grads, i, mine[k].shape) for i, k in enumerate(names)}
    # one view: every gradient lands in its sink
    loss_of(plain, 0).backward()
    loss_of(mine, 0).backward()
    for k in names:
        assert mine[k].grad.data_ptr() == sinks[k].data_ptr(), k
        assert _rel(mine[k].grad, plain[k].grad) <= 1e-5, k
    # the step on those gradients = clip + Adam on a copy of them
    ref = {k: torch.nn.Parameter(sc[k].to(dev)) for k in names}
    for k in names:
        ref[k].grad = mine[k].grad.clone()
    opt_ref = torch.optim.Adam([{"params": [ref[k]], "lr": LRS[k]} for k in names], lr=1e-3, eps=1e-15)
    torch.nn.utils.clip_grad_norm_(ref["pos"], max_norm=1.0)
    opt_ref.step()
    opt.step()
    for k in names:
        assert _rel(mine[k].detach(), ref[k].detach()) <= 2e-6, k
    # two views, one backward (scripts/train.py:471-530): the sum, whichever tensor autograd ends up keeping
    for leaves in (plain, mine):
        for p in leaves.values():
            p.grad = None
    with torch.no_grad():
        for k in names:
            plain[k].copy_(mine[k])
    (loss_of(plain, 0) + loss_of(plain, 1)).backward()
    (loss_of(mine, 0) + loss_of(mine, 1)).backward()
    for k in names:
        assert _rel(mine[k].grad, plain[k].grad) <= 1e-5, k
    # two backward calls without a step in between: the second one accumulates into the first
    for p in mine.values():
        p.grad = None
    opt.step()                                  # (consumes nothing: no gradients; frees the sinks)
    with torch.no_grad():
        for k in names:
            plain[k].copy_(mine[k])
    for p in plain.values():
        p.grad = None
    for v in (0, 1):
        loss_of(plain, v).backward()
        loss_of(mine, v).backward()
    for k in names:
        assert mine[k].grad.data_ptr() == sinks[k].data_ptr(), k
        assert _rel(mine[k].grad, plain[k].grad) <= 1e-5, k
