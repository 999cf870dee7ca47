"""Multi-GPU parity (needs >= 2 CUDA devices; skipped otherwise): NCCL data-parallel gradients equal the
single-GPU sum over the same views, and a tile-row sharded frame equals the full frame bit for bit."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
PARAMS = ("pos", "scale_raw", "q_raw", "opacity_raw", "f_dc", "f_rest")


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _views(n_views, W, H):
    from oracle import gs_oracle as O
    return [O.make_camera(W, H, view=v, n_views=n_views) for v in range(n_views)]


def _loss_for_views(gs, leaves, cams, views, W, H, dev, n_total):
    sigma = gs.build_sigma_from_params(leaves["scale_raw"], leaves["q_raw"])
    total = None
    for v in views:
        cam = cams[v]
        c2w = cam["c2w"].to(dev)
        col = gs.evaluate_sh(leaves["f_dc"], leaves["f_rest"], leaves["pos"], c2w)
        img = gs.render(leaves["pos"], col, leaves["opacity_raw"], sigma, c2w, H, W, cam["fx"], cam["fy"], cam["cx"], cam["cy"])
        w = torch.rand(H, W, 3, generator=torch.Generator().manual_seed(100 + v)).to(dev)
        term = (img * w).sum() / n_total          # each view's loss is divided by the GLOBAL batch (train.py:514-521)
        total = term if total is None else total + term
    return total


def _worker(rank, world, port, out):
    import torch.distributed as dist
    import b200gs
    from b200gs.dist import allreduce_gradients, render_tile_row_sharded, shard_views
    from oracle import gs_oracle as O
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        W, H, n_views = 320, 200, 4
        sc = O.make_scene(20000, seed=5, log_scale=-3.6)
        cams = _views(n_views, W, H)
        leaves = {k: sc[k].to(dev).requires_grad_(True) for k in PARAMS}
        loss = _loss_for_views(b200gs, leaves, cams, shard_views(n_views, rank, world), W, H, dev, n_views)
        loss.backward()
        allreduce_gradients(leaves.values())
        out[f"grads{rank}"] = {k: leaves[k].grad.cpu() for k in PARAMS}
        # tile-row sharded single frame
        with torch.no_grad():
            c2w = cams[1]["c2w"].to(dev)
            sigma = b200gs.build_sigma_from_params(leaves["scale_raw"], leaves["q_raw"])
            col = b200gs.evaluate_sh(leaves["f_dc"], leaves["f_rest"], leaves["pos"], c2w)
            fn = lambda tile_rows: b200gs.render(leaves["pos"], col, leaves["opacity_raw"], sigma, c2w, H, W, cams[1]["fx"],
                                                 cams[1]["fy"], cams[1]["cx"], cams[1]["cy"], tile_rows=tile_rows)
            out[f"img{rank}"] = render_tile_row_sharded(fn, (H + 15) // 16).cpu()
            # the same frame through TileRowRenderer: bands stored straight into rank 0's buffer over peer memory
            from b200gs.dist import TileRowRenderer
            # ... with the Gaussians replicated (every rank culls all of them to its band), and routed: every rank
            # projects its slice and writes the splat records into the bands' workspaces over peer memory
            tr_rep = TileRowRenderer(H, W, dev, routed=False)
            tr = TileRowRenderer(H, W, dev)
            assert tr.routed and not tr_rep.routed
            for rep in range(3):               # several frames: the buffers are reused, the barriers must order them
                cam_k = cams[1 + rep % 2]
                c2w_k = cam_k["c2w"].to(dev)
                for name, t in (("tr", tr), ("trrep", tr_rep)):
                    col_k = b200gs.evaluate_sh(leaves["f_dc"], leaves["f_rest"], leaves["pos"], c2w_k)
                    img = t.render(leaves["pos"], col_k, leaves["opacity_raw"], sigma, c2w_k, cam_k["fx"], cam_k["fy"],
                                   cam_k["cx"], cam_k["cy"])
                    torch.cuda.synchronize()
                    if rank == 0:
                        out[f"{name}{rep}"] = img.cpu()
            # records on a side stream behind the keys (opt-in), and a frame whose counter check is deferred
            assert not tr.split_records
            tr.split_records = True
            img = tr.render(leaves["pos"], col, leaves["opacity_raw"], sigma, c2w, cams[1]["fx"], cams[1]["fy"],
                            cams[1]["cx"], cams[1]["cy"], defer_check=True)
            assert not tr.finish()
            torch.cuda.synchronize()
            if rank == 0:
                out["tr_onepass"] = img.cpu()
            tr.split_records = False
            # precomputed sigma / colour tensors (no tags): the routed slices are views of them
            from b200gs import api
            img = tr.render(leaves["pos"], api._real(col).clone(), leaves["opacity_raw"], api._real(sigma).clone(), c2w,
                            cams[1]["fx"], cams[1]["fy"], cams[1]["cx"], cams[1]["cy"])
            torch.cuda.synchronize()
            if rank == 0:
                out["tr_unfused"] = img.cpu()
            w8 = tr.row_weights(leaves["pos"], col, leaves["opacity_raw"], sigma, c2w, cams[1]["fx"], cams[1]["fy"],
                                cams[1]["cx"], cams[1]["cy"])
            tr.set_weights(w8)
            img = tr.render(leaves["pos"], col, leaves["opacity_raw"], sigma, c2w, cams[1]["fx"], cams[1]["fy"],
                            cams[1]["cx"], cams[1]["cy"])
            torch.cuda.synchronize()
            if rank == 0:
                out["tr_weighted"] = img.cpu()
                out["tr_bands"] = list(tr.bands)
        # one flat bucket + one NCCL all-reduce gives the same gradients as the six all-reduces
        from b200gs.dist import GradBucket
        leaves2 = {k: sc[k].to(dev).requires_grad_(True) for k in PARAMS}
        bucket = GradBucket([leaves2[k] for k in PARAMS])
        _loss_for_views(b200gs, leaves2, cams, shard_views(n_views, rank, world), W, H, dev, n_views).backward()
        bucket.allreduce()       # (several views per backward: whichever gradient arrives first is adopted, the rest is copied)
        assert leaves2["f_rest"].grad.data_ptr() == bucket.views[PARAMS.index("f_rest")].data_ptr()
        out[f"bgrads{rank}"] = {k: leaves2[k].grad.cpu() for k in PARAMS}
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_dp_gradients_and_tile_row_bands_match_single_gpu():
    import torch.multiprocessing as mp
    import b200gs
    from oracle import gs_oracle as O
    world, port = 2, _free_port()
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
        out = dict(out)
    # single-GPU reference: all views on one device
    dev = torch.device("cuda", 0)
    W, H, n_views = 320, 200, 4
    sc = O.make_scene(20000, seed=5, log_scale=-3.6)
    cams = _views(n_views, W, H)
    leaves = {k: sc[k].to(dev).requires_grad_(True) for k in PARAMS}
    _loss_for_views(b200gs, leaves, cams, list(range(n_views)), W, H, dev, n_views).backward()
    for k in PARAMS:
        ref = leaves[k].grad.cpu()
        for r in range(world):
            got = out[f"grads{r}"][k]
            err = float((got - ref).abs().max() / ref.abs().max())
            assert err <= 1e-5, (k, r, err)          # float atomics reorder sums: not bit-exact
        assert torch.equal(out["grads0"][k], out["grads1"][k])    # the all-reduce leaves identical replicas
    with torch.no_grad():
        c2w = cams[1]["c2w"].to(dev)
        sigma = b200gs.build_sigma_from_params(leaves["scale_raw"], leaves["q_raw"])
        col = b200gs.evaluate_sh(leaves["f_dc"], leaves["f_rest"], leaves["pos"], c2w)
        full = b200gs.render(leaves["pos"], col, leaves["opacity_raw"], sigma, c2w, H, W, cams[1]["fx"], cams[1]["fy"],
                             cams[1]["cx"], cams[1]["cy"]).cpu()
    assert torch.equal(out["img0"], full) and torch.equal(out["img1"], full)
    assert torch.equal(out["tr0"], full) and torch.equal(out["tr2"], full) and torch.equal(out["tr_weighted"], full)
    assert torch.equal(out["trrep0"], full) and torch.equal(out["trrep2"], full) and torch.equal(out["tr_onepass"], full)
    with torch.no_grad():
        from b200gs import api
        unfused = b200gs.render(leaves["pos"], api._real(col).clone(), leaves["opacity_raw"], api._real(sigma).clone(), c2w,
                                H, W, cams[1]["fx"], cams[1]["fy"], cams[1]["cx"], cams[1]["cy"]).cpu()
    assert torch.equal(out["tr_unfused"], unfused)
    assert out["tr_bands"][0][0] == 0 and out["tr_bands"][-1][1] == (H + 15) // 16
    with torch.no_grad():
        c2w = cams[2]["c2w"].to(dev)
        col = b200gs.evaluate_sh(leaves["f_dc"], leaves["f_rest"], leaves["pos"], c2w)
        other = b200gs.render(leaves["pos"], col, leaves["opacity_raw"], sigma, c2w, H, W, cams[2]["fx"], cams[2]["fy"],
                              cams[2]["cx"], cams[2]["cy"]).cpu()
    assert torch.equal(out["tr1"], other) and torch.equal(out["trrep1"], other)
    for k in PARAMS:
        for r in range(world):
            err = float((out[f"bgrads{r}"][k] - leaves[k].grad.cpu()).abs().max() / leaves[k].grad.abs().max())
            assert err <= 1e-5, (k, r, err)
