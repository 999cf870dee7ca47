"""Run a reference script unchanged on the b200gs render path:

    python -m b200gs.run /path/to/reference/scripts/render_trained.py --checkpoint_dir ... --data_dir ... [args]

Equivalent to `install()` followed by executing the script as __main__ (B200GS_PATCH_ADAM=1 also swaps
torch.optim.Adam / clip_grad_norm_ for the fused versions).  The reference checkout must
be importable: its root is derived from the script location (<root>/scripts/x.py) or taken from
$B200GS_REFERENCE_ROOT.
"""
from __future__ import annotations

import os
import runpy
import sys


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    if not argv:
        print(__doc__)
        return 2
    script = os.path.abspath(argv[0])
    root = os.environ.get("B200GS_REFERENCE_ROOT") or os.path.dirname(os.path.dirname(script))
    if root not in sys.path:
        sys.path.insert(0, root)
    from .install import install
    install(optimizer=os.environ.get("B200GS_PATCH_ADAM", "0") == "1")
    sys.argv = [script] + argv[1:]
    runpy.run_path(script, run_name="__main__")
    return 0


if __name__ == "__main__":
    sys.exit(main())
