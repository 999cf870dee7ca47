"""Parity of the two largest BASELINE.json configurations against the oracle, forward (too large for the test suite's
minute budget, so a tool; its output is merged into profiles/PARITY_r02.json):

  C4  orbit scene, 3M Gaussians, 1920x1080           (b200gs.render, one frame)
  C5  6M Gaussians, 3840x2160                         (b200gs.dist.TileRowRenderer, i.e. the band path when run under
                                                       torchrun, the full-frame path on one GPU)

    python tools/parity_configs.py [c4] [c5]          (one GPU)
    torchrun --nnodes=1 --nproc-per-node N ... tools/parity_configs.py c5

Survivors, depths, radii, tile rects, per-tile sorted lists: exact; image: values beyond 1e-4 counted.  The oracle needs
a few minutes of CPU time and ~20 GB of host memory for C5.  TEST INFRASTRUCTURE (imports oracle/).
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "3d-gaussian-splatting-for-novel-view-synthesis_b200"))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
import b200gs  # noqa: E402
from b200gs import ops  # noqa: E402
from oracle import gs_oracle as O  # noqa: E402
from oracle import parity as PAR  # noqa: E402

CONFIGS = {"c4": dict(n=3_000_000, W=1920, H=1080, ls=-5.5, view=5), "c5": dict(n=6_000_000, W=3840, H=2160, ls=-6.0, view=0)}


def main():
    names = [a for a in sys.argv[1:] if a in CONFIGS] or ["c4", "c5"]
    world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    out = {}
    for name in names:
        cfg = CONFIGS[name]
        n, W, H = cfg["n"], cfg["W"], cfg["H"]
        sc = O.make_scene(n, seed=0, log_scale=cfg["ls"])
        cam = O.make_camera(W, H, view=cfg["view"], n_views=16)
        scd = {k: v.to(dev) for k, v in sc.items()}
        c2w = cam["c2w"].to(dev)
        with torch.no_grad():
            # the frame through the product path of this configuration
            if name == "c5":
                from b200gs.dist import TileRowRenderer
                tr = TileRowRenderer(H, W, dev)
                sigma = b200gs.build_sigma_from_params(scd["scale_raw"], scd["q_raw"])
                color = b200gs.evaluate_sh(scd["f_dc"], scd["f_rest"], scd["pos"], c2w)
                img = tr.render(scd["pos"], color, scd["opacity_raw"], sigma, c2w, cam["fx"], cam["fy"], cam["cx"], cam["cy"])
            else:
                sigma = b200gs.build_sigma_from_params(scd["scale_raw"], scd["q_raw"])
                color = b200gs.evaluate_sh(scd["f_dc"], scd["f_rest"], scd["pos"], c2w)
                img = b200gs.render(scd["pos"], color, scd["opacity_raw"], sigma, c2w, H, W, cam["fx"], cam["fy"], cam["cx"], cam["cy"])
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            if rank != 0:
                continue
            img = img.cpu().numpy()
            # integer stages of the full frame on one GPU (introspection export)
            g, keep = ops._gaussians(scd["pos"], scd["opacity_raw"], scd["scale_raw"], scd["q_raw"], None, scd["f_dc"], scd["f_rest"], None)
            fr = ops.Frame(g, keep, ops.RenderConfig(H=H, W=W, fx=cam["fx"], fy=cam["fy"], cx=cam["cx"], cy=cam["cy"]), c2w.contiguous(), dev)
            img_full = fr.render("sync")
            fr.refresh_stats()
            ex = {k: v.numpy() for k, v in fr.export().items()}
            same = bool((torch.from_numpy(img).to(dev) == img_full).all())
            t0 = time.perf_counter()
            torch.set_num_threads(os.cpu_count() or 1)
            img_ref, proj, bins = O.render(sc["pos"], O.evaluate_sh(sc["f_dc"], sc["f_rest"], sc["pos"], cam["c2w"]),
                                           sc["opacity_raw"], O.build_sigma_from_params(sc["scale_raw"], sc["q_raw"]),
                                           cam["c2w"], H, W, cam["fx"], cam["fy"], cam["cx"], cam["cy"], return_stages=True)
            img64 = None
            if "--fp64" in sys.argv:         # the reference algorithm in fp64: the yardstick for the threshold flips
                s64 = {k: v.double() for k, v in sc.items()}
                img64 = O.render_from_params(s64["pos"], s64["scale_raw"], s64["q_raw"], s64["opacity_raw"], s64["f_dc"],
                                             s64["f_rest"], cam["c2w"].double(), H, W, cam["fx"], cam["fy"], cam["cx"],
                                             cam["cy"]).numpy()
                del s64
            rep = PAR.compare_frame(ex, fr.n_isect, fr.n_visible, img, proj, bins, img_ref.numpy(), image_ref64=img64)
            rep.update(oracle_seconds=round(time.perf_counter() - t0, 1), n_gpus=world,
                       product_frame_equals_single_gpu_frame=same,
                       path="TileRowRenderer" if name == "c5" else "render")
            out[f"{name.upper()}_{n // 1_000_000}M_{W}x{H}_fwd" + (f"_{world}gpu" if world > 1 else "")] = rep
            print(json.dumps({name: rep}), flush=True)
        del scd
        torch.cuda.empty_cache()
    if rank == 0:
        path = os.path.join(ROOT, "gpurun_out", f"parity_configs_{world}gpu.json")
        os.makedirs(os.path.dirname(path), exist_ok=True)
        json.dump(out, open(path, "w"), indent=1)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
