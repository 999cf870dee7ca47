"""Stages the UNMODIFIED reference checkout under baseline/_ref/ so that `bench.py --impl reference` can time the
reference's own `evaluate_sh` + `render` (its stock CPU code path) on the GPU box, where /root/reference does not exist.

    python baseline/stage_reference.py [--src /root/reference]

The reference has no setup.py / pyproject.toml (a `pip install` has nothing to build), and it is pure Python, so the
"install" is a byte-for-byte copy of its importable package and scripts.  baseline/_ref/ is git-ignored (no reference
source enters the history) but NOT gpurun-ignored, so it travels with the snapshot.  MANIFEST.json records the sha256 of
every staged file next to the sha256 of its source, which is how a reader checks that nothing was edited.
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
WHAT = ("gaussian_splatting", "scripts", "LICENSE", "requirements.txt")


def _sha(path: str) -> str:
    return hashlib.sha256(open(path, "rb").read()).hexdigest()


def stage(src: str = "/root/reference", dest: str = DEST) -> dict:
    if not os.path.isdir(os.path.join(src, "gaussian_splatting")):
        raise FileNotFoundError(f"{src} is not a checkout of the reference (no gaussian_splatting/ package)")
    os.makedirs(dest, exist_ok=True)
    manifest = {"source": src, "files": {}}
    for name in WHAT:
        s = os.path.join(src, name)
        if not os.path.exists(s):
            continue
        d = os.path.join(dest, name)
        if os.path.isdir(s):
            if os.path.isdir(d):
                shutil.rmtree(d)
            shutil.copytree(s, d, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
            for root, _, files in os.walk(d):
                for f in files:
                    rel = os.path.relpath(os.path.join(root, f), dest)
                    manifest["files"][rel] = {"sha256": _sha(os.path.join(root, f)), "source_sha256": _sha(os.path.join(src, rel))}
        else:
            shutil.copyfile(s, d)
            manifest["files"][name] = {"sha256": _sha(d), "source_sha256": _sha(s)}
    bad = [k for k, v in manifest["files"].items() if v["sha256"] != v["source_sha256"]]
    if bad:
        raise RuntimeError(f"staged files differ from their sources: {bad}")
    json.dump(manifest, open(os.path.join(dest, "MANIFEST.json"), "w"), indent=1, sort_keys=True)
    return manifest


def staged_root(dest: str = DEST):
    """Path to put on sys.path to `import gaussian_splatting` (the staged copy), or None when nothing is staged."""
    return dest if os.path.exists(os.path.join(dest, "gaussian_splatting", "__init__.py")) else None


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--src", default="/root/reference")
    a = ap.parse_args()
    m = stage(a.src)
    print(f"staged {len(m['files'])} files from {a.src} into {DEST}")
    sys.exit(0)
