"""Unchanged-script data parallelism, checked on real GPUs: scripts/train.py of the (staged, unmodified) reference under
`python -m b200gs.run --dp`, once as 1 process x batch 2 and once as 2 processes x batch 1 (torchrun).  Both visit the
same pairs of views per iteration (shared permutation, rank r takes entries r, r + world, ...), so after K iterations
the parameters must agree up to what float atomics and a different summation order do to Adam; the 1-process run is
repeated to measure that floor.

    python tools/dp_launcher_check.py [--iterations 30] [--gpus 2]

Prints one JSON document (also written to gpurun_out/dp_launcher_check.json).  Needs baseline/_ref (the staged
reference) and >= 2 GPUs.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "3d-gaussian-splatting-for-novel-view-synthesis_b200")
sys.path.insert(0, ROOT)
sys.path.insert(0, PKG)
sys.path.insert(0, os.path.join(ROOT, "tools"))
KEYS = ("pos", "opacity_raw", "f_dc", "f_rest", "scale_raw", "q_raw")


def run(cmd, env, log):
    p = subprocess.run(cmd, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    open(log, "w").write(p.stdout)
    if p.returncode:
        print(p.stdout[-3000:])
        raise SystemExit(f"{' '.join(cmd)} -> exit code {p.returncode}")


def compare(a, b):
    import torch
    out = {}
    for k in KEYS:
        x, y = a[k].float(), b[k].float()
        if x.shape != y.shape:
            out[k] = {"shape_a": list(x.shape), "shape_b": list(y.shape)}
            continue
        d = (x - y).abs()
        out[k] = {"max_abs": float(d.max()), "mean_abs": float(d.mean()), "frac_within_1e-5": float((d <= 1e-5).float().mean()),
                  "max_abs_value": float(x.abs().max())}
    return out


def main():
    import torch
    from run_reference_scripts import make_dataset
    ap = argparse.ArgumentParser()
    ap.add_argument("--iterations", type=int, default=30)
    ap.add_argument("--gpus", type=int, default=2)
    args = ap.parse_args()
    ref = os.environ.get("B200GS_REFERENCE_ROOT") or os.path.join(ROOT, "baseline", "_ref")
    script = os.path.join(ref, "scripts", "train.py")
    if not os.path.exists(script):
        raise SystemExit("no staged reference: run `python baseline/stage_reference.py` in the build container")
    logdir = os.path.join(ROOT, "gpurun_out")
    os.makedirs(logdir, exist_ok=True)
    work = tempfile.mkdtemp(prefix="b200gs_dp_")
    data = os.path.join(work, "data")
    make_dataset(data)
    env = dict(os.environ, PYTHONPATH=PKG + os.pathsep + os.environ.get("PYTHONPATH", ""), B200GS_REFERENCE_ROOT=ref,
               PYTHONDONTWRITEBYTECODE="1", CUDA_VISIBLE_DEVICES=",".join(str(i) for i in range(args.gpus)))
    common = ["--data_dir", data, "--iterations", str(args.iterations), "--scale_factor", "1.0"]
    single = [sys.executable, "-m", "b200gs.run", "--dp", script] + common + ["--batch_size", str(args.gpus)]
    multi = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}", "--master-addr",
             "127.0.0.1", "--master-port", "29611", "-m", "b200gs.run", "--dp", script] + common + ["--batch_size", "1"]
    outs = {}
    for tag, cmd in (("single_a", single), ("single_b", single), ("multi", multi)):
        out = os.path.join(work, tag)
        run(cmd + ["--output_dir", out], dict(env, CUDA_VISIBLE_DEVICES="0") if tag.startswith("single") else env,
            os.path.join(logdir, f"dp_check_{tag}.log"))
        outs[tag] = torch.load(os.path.join(out, "checkpoint_final.pt"), map_location="cpu")
    doc = {"iterations": args.iterations, "gpus": args.gpus, "n_gaussians": {k: int(v["pos"].shape[0]) for k, v in outs.items()},
           "what": f"1 process x batch {args.gpus} against {args.gpus} processes x batch 1 through `b200gs.run --dp`, reference "
                   "scripts/train.py byte-for-byte unchanged; `floor` = the 1-process run against its own repetition",
           "multi_vs_single": compare(outs["single_a"], outs["multi"]),
           "floor_single_vs_single": compare(outs["single_a"], outs["single_b"])}
    print(json.dumps(doc, indent=1))
    json.dump(doc, open(os.path.join(logdir, "dp_launcher_check.json"), "w"), indent=1)
    return 0


if __name__ == "__main__":
    sys.exit(main())
