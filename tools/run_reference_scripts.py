"""Manual integration check: runs the reference's OWN scripts, unchanged, on top of b200gs.

    B200GS_REFERENCE_ROOT=/path/to/reference python tools/run_reference_scripts.py [--iterations 150]

Builds a small synthetic data directory in the reference's training format (images/*.png, cam_meta.npy, poses.npy,
pointcloud.npy; gaussian_splatting/data_loader.py:153-284) whose target images are renders of a seeded "ground truth"
scene, then executes

    python -m b200gs.run <reference>/scripts/train.py          --data_dir D --output_dir O --iterations N --scale_factor 1.0
    python -m b200gs.run <reference>/scripts/render_trained.py --checkpoint_dir O --data_dir D --orbit_frames 12 --benchmark_only

once with the stock torch.optim.Adam and once with B200GS_PATCH_ADAM=1, and prints the tail of each log.  The reference
checkout is not part of this repository and does not exist on the benchmark box, so this is not a pytest test; a log of
the last run lives in profiles/.
"""
import argparse
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "3d-gaussian-splatting-for-novel-view-synthesis_b200")
sys.path.insert(0, ROOT)
sys.path.insert(0, PKG)


def make_dataset(d, n_views=12, W=320, H=200):
    import numpy as np
    import torch
    from PIL import Image
    import b200gs
    from oracle import gs_oracle as O       # scene / camera generator only
    os.makedirs(os.path.join(d, "images"), exist_ok=True)
    sc = {k: v.cuda() for k, v in O.make_scene(20_000, seed=11, log_scale=-3.6).items()}
    cams = [O.make_camera(W, H, view=v, n_views=n_views) for v in range(n_views)]
    poses = []
    with torch.no_grad():
        sigma = b200gs.build_sigma_from_params(sc["scale_raw"], sc["q_raw"])
        for v, cam in enumerate(cams):
            c2w = cam["c2w"].cuda()
            col = b200gs.evaluate_sh(sc["f_dc"], sc["f_rest"], sc["pos"], c2w)
            img = b200gs.render(sc["pos"], col, sc["opacity_raw"], sigma, c2w, H, W, cam["fx"], cam["fy"], cam["cx"], cam["cy"])
            Image.fromarray(b200gs.to_uint8(img).cpu().numpy()).save(os.path.join(d, "images", f"view_{v:03d}.png"))
            poses.append(cam["c2w"].numpy())
    K = cams[0]
    np.save(os.path.join(d, "cam_meta.npy"), {"fx": K["fx"], "fy": K["fy"], "cx": K["cx"], "cy": K["cy"], "height": H, "width": W},
            allow_pickle=True)
    np.save(os.path.join(d, "poses.npy"), np.stack(poses).astype(np.float32))
    np.save(os.path.join(d, "pointcloud.npy"), sc["pos"].cpu().numpy()[::4].astype(np.float32))


def run(cmd, env, log):
    print("$", " ".join(cmd), flush=True)
    p = subprocess.run(cmd, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    open(log, "w").write(p.stdout)
    tail = [ln for ln in p.stdout.replace("\r", "\n").splitlines() if ln.strip()][-14:]
    print("\n".join(tail))
    print(f"--> exit code {p.returncode}\n", flush=True)
    return p.returncode


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iterations", type=int, default=150)
    ap.add_argument("--logdir", default=os.path.join(ROOT, "gpurun_out"))
    ap.add_argument("--profile", action="store_true", help="run the training script under cProfile and print the top entries")
    args = ap.parse_args()
    ref = os.environ.get("B200GS_REFERENCE_ROOT") or os.path.join(ROOT, "baseline", "_ref")    # the staged, unmodified copy
    if not os.path.exists(os.path.join(ref, "scripts", "train.py")):
        raise SystemExit("set B200GS_REFERENCE_ROOT to a checkout of the reference repository (or stage one: "
                         "python baseline/stage_reference.py)")
    os.makedirs(args.logdir, exist_ok=True)
    work = tempfile.mkdtemp(prefix="b200gs_scripts_")
    data = os.path.join(work, "data")
    make_dataset(data)
    env = dict(os.environ, PYTHONPATH=PKG + os.pathsep + os.environ.get("PYTHONPATH", ""), B200GS_REFERENCE_ROOT=ref,
               PYTHONDONTWRITEBYTECODE="1")
    rc = 0
    for tag, extra in (("stock_adam", {}), ("fused_adam", {"B200GS_PATCH_ADAM": "1"})):
        out = os.path.join(work, "out_" + tag)
        e = dict(env, **extra)
        prof = ["-m", "cProfile", "-o", os.path.join(work, tag + ".prof")] if args.profile else []
        rc |= run([sys.executable, *prof, "-m", "b200gs.run", os.path.join(ref, "scripts", "train.py"), "--data_dir", data,
                   "--output_dir", out, "--iterations", str(args.iterations), "--scale_factor", "1.0"], e,
                  os.path.join(args.logdir, f"ref_train_{tag}.log"))
        if args.profile:
            import pstats
            st = pstats.Stats(os.path.join(work, tag + ".prof"))
            print(f"----- cProfile, {tag}: top 25 by cumulative time -----")
            st.sort_stats("cumulative").print_stats(25)
        rc |= run([sys.executable, "-m", "b200gs.run", os.path.join(ref, "scripts", "render_trained.py"), "--checkpoint_dir", out,
                   "--data_dir", data, "--orbit_frames", "12", "--benchmark_only", "--output_dir", os.path.join(work, "renders_" + tag)],
                  e, os.path.join(args.logdir, f"ref_render_{tag}.log"))
    return rc


if __name__ == "__main__":
    sys.exit(main())
