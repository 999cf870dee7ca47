"""Generates tests/golden/io/: a tiny trained-scene directory written BY THE REFERENCE's own code
(scripts/train.py: GaussianModel.save_checkpoint :197-208; the loose-file block :591-597 is inline in train(), so its six
torch.save lines are executed here verbatim from the reference source) plus a cam_meta.npy / poses.npy pair in the
format gaussian_splatting/data_loader.py:30-47 documents, and an .npz with the values for the test.

Test infrastructure only; needs /root/reference (build container).  python oracle/make_golden_io.py
"""
import os
import re
import sys

import numpy as np
import torch

REF = os.environ.get("B200GS_REFERENCE_ROOT", "/root/reference")
sys.dont_write_bytecode = True
sys.path.insert(0, REF)
sys.path.insert(0, os.path.join(REF, "scripts"))
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "io")


def main():
    import importlib.util
    from pathlib import Path
    spec = importlib.util.spec_from_file_location("ref_train", os.path.join(REF, "scripts", "train.py"))
    ref_train = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref_train)
    g = torch.Generator().manual_seed(42)
    n = 7
    init = {"pos": torch.randn(n, 3, generator=g), "opacity_raw": torch.randn(n, generator=g),
            "f_dc": torch.randn(n, 3, generator=g), "f_rest": torch.randn(n, 45, generator=g) * 0.1,
            "scale_raw": torch.randn(n, 3, generator=g) - 3.0, "q_raw": torch.randn(n, 4, generator=g)}
    model = ref_train.GaussianModel(init, device="cpu")
    output_dir = Path(OUT)
    output_dir.mkdir(parents=True, exist_ok=True)
    iteration = 3000
    model.save_checkpoint(output_dir / f"checkpoint_{iteration:06d}.pt", iteration)
    # the loose-file block of train(): executed from the reference's source text
    src = open(os.path.join(REF, "scripts", "train.py")).read()
    lines = [ln.strip() for ln in src.splitlines() if re.match(r"\s*torch\.save\(model\.\w+\.cpu\(\), output_dir / f'", ln)]
    assert len(lines) == 6, lines
    for ln in lines:
        exec(ln, {"torch": torch, "model": model, "output_dir": output_dir, "iteration": iteration})
    model.save_checkpoint(output_dir / "checkpoint_final.pt", 3007)
    cam = {"fx": 230.4, "fy": 231.0, "height": 200, "width": 320}            # no cx / cy: the scripts default them
    np.save(output_dir / "cam_meta.npy", cam, allow_pickle=True)
    poses = np.tile(np.eye(4, dtype=np.float32), (3, 1, 1))
    poses[:, :3, 3] = np.array([[0, 0, -3], [1, 0, -3], [0, 1, -2]], dtype=np.float32)
    np.save(output_dir / "poses.npy", poses)
    np.savez(output_dir / "values.npz", **{k: v.detach().numpy() for k, v in model.get_params().items()})
    print(sorted(os.listdir(OUT)))


if __name__ == "__main__":
    main()
