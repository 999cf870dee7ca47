"""Run a reference script unchanged on the b200gs render path:

    python -m b200gs.run /path/to/reference/scripts/render_trained.py --checkpoint_dir ... --data_dir ... [args]

Equivalent to `install()` followed by executing the script as __main__ (B200GS_PATCH_ADAM=1 also swaps
torch.optim.Adam / clip_grad_norm_ for the fused versions).  The reference checkout must
be importable: its root is derived from the script location (<root>/scripts/x.py) or taken from
$B200GS_REFERENCE_ROOT.

Data-parallel training, script still byte-for-byte unchanged (one process per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29500 \\
        -m b200gs.run --dp /path/to/reference/scripts/train.py --data_dir ... --batch_size 1 [args]

`--dp` turns the single-GPU loop of scripts/train.py into data parallelism over the training views without touching it
(the reference itself only prints the GPU count, scripts/train.py:285-291):
  * every rank sees ONE GPU as cuda:0 (the script hard-codes `cuda:0`, train.py:372) and joins an NCCL group;
  * all ranks are seeded identically (B200GS_DP_SEED, default 0), so the model initialisation (data_loader.py:327,346)
    and the split noise of densification (train.py:164) are the same everywhere - the replicas never diverge;
  * `DataLoader(shuffle=True)` (train.py:318-324) is given a sampler that draws ONE permutation of the views per epoch
    from a generator shared by all ranks and hands rank r the entries r, r + world, ... - so an iteration of the job
    processes world x batch_size distinct views, and a run with p processes x batch b visits the same views per
    iteration as 1 process x batch p*b;
  * the script averages a batch's losses over ITS batch size (train.py:514-521); the launcher completes the average over
    the global batch: the first of `clip_grad_norm_` (train.py:536) / `optimizer.step()` (train.py:538) after a backward
    sums the six gradient tensors over the ranks (one flat bucket, one NCCL all-reduce - b200gs.dist.GradBucket) and
    divides by the world size; clipping, Adam and the densification statistics (train.py:544-557) then see the same
    reduced gradient on every rank;
  * only rank 0 writes checkpoints and prints progress.
Without torchrun (`python -m b200gs.run --dp ...`) the same code runs as a world of one.
"""
from __future__ import annotations

import os
import runpy
import sys


class ShardedRandomSampler:
    """Sampler for `DataLoader(shuffle=True)` under `--dp`: ONE permutation of the views per epoch, drawn from a generator
    that every rank seeds identically, of which rank r takes entries r, r + world, ... (padded by wrapping around, like
    DistributedSampler, so that every rank sees the same number of views per epoch).  p ranks x batch b therefore visit
    the same views per iteration as 1 rank x batch p*b."""

    def __init__(self, n: int, rank: int, world: int, seed: int):
        import torch
        self.n, self.rank, self.world = int(n), int(rank), int(world)
        self.gen = torch.Generator().manual_seed(int(seed) + 7919)      # same stream on every rank

    def __iter__(self):
        import torch
        perm = torch.randperm(self.n, generator=self.gen).tolist()
        total = -(-self.n // self.world) * self.world
        perm = (perm * (total // max(self.n, 1) + 1))[:total]
        return iter(perm[self.rank::self.world])

    def __len__(self):
        return -(-self.n // self.world)


class GradientReducer:
    """Reduces the gradients of the live optimizer's parameters ONCE per iteration - at the first of `clip_grad_norm_`
    (train.py:536) / `optimizer.step()` (train.py:538) after a backward - through a `b200gs.dist.GradBucket` (one flat
    buffer, one all-reduce) and divides by the world size, which completes the script's per-batch loss average
    (train.py:514-521) to the global batch."""

    def __init__(self, group=None):
        self.group, self.bucket, self.params, self.reduced = group, None, None, False

    def track(self, params):
        """A (new) optimizer was constructed over `params` (the script re-creates it after every densification)."""
        if self.params is not None and self.params[0].is_cuda:
            from . import ops
            ops.unregister_grad_sinks(self.params)
        self.bucket, self.params, self.reduced = None, list(params), False

    def ensure_reduced(self):
        if self.reduced or not self.params:
            return
        if self.bucket is None:
            from .dist import GradBucket
            self.bucket = GradBucket(self.params, group=self.group)
        self.bucket.allreduce(average=True)
        self.reduced = True

    def next_iteration(self):
        self.reduced = False

    def wrap_optimizer(self, cls):
        reducer = self

        class DataParallel(cls):
            def __init__(self, params, *a, **kw):
                super().__init__(params, *a, **kw)
                reducer.track([p for g in self.param_groups for p in g["params"]])

            def step(self, closure=None):
                reducer.ensure_reduced()
                out = super().step(closure)
                reducer.next_iteration()
                return out

            def zero_grad(self, set_to_none=True):
                reducer.next_iteration()
                return super().zero_grad(set_to_none)
        DataParallel.__name__ = cls.__name__
        return DataParallel

    def wrap_clip(self, fn):
        def clip_grad_norm_(parameters, *a, **kw):
            self.ensure_reduced()
            return fn(parameters, *a, **kw)
        return clip_grad_norm_


def _setup_dp():
    """Everything `--dp` changes, applied BEFORE the script is imported.  Returns (rank, world)."""
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and os.environ.get("B200GS_DP_KEEP_VISIBLE") != "1":
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        ids = [v for v in vis.split(",") if v] if vis else None
        os.environ["CUDA_VISIBLE_DEVICES"] = ids[local] if ids and local < len(ids) else str(local)
    import random
    import numpy as np
    import torch
    import torch.distributed as dist
    import torch.utils.data as tud
    seed = int(os.environ.get("B200GS_DP_SEED", "0"))
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)                       # CPU and every CUDA generator
    if world > 1:
        torch.cuda.set_device(0)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", 0))

    # ---- views: one shared permutation per epoch, rank r takes entries r, r + world, ... ----------------------------
    orig_init = tud.DataLoader.__init__

    def dl_init(self, dataset, *a, **kw):
        if kw.get("shuffle") and kw.get("sampler") is None and kw.get("batch_sampler") is None:
            kw["shuffle"] = False
            kw["sampler"] = ShardedRandomSampler(len(dataset), rank, world, seed)
        orig_init(self, dataset, *a, **kw)
    tud.DataLoader.__init__ = dl_init

    # ---- gradients: reduced once per iteration by whichever of clip_grad_norm_ / optimizer.step comes first ---------
    reducer = GradientReducer()
    torch.optim.Adam = reducer.wrap_optimizer(torch.optim.Adam)
    torch.nn.utils.clip_grad_norm_ = reducer.wrap_clip(torch.nn.utils.clip_grad_norm_)

    # ---- side effects: rank 0 only ---------------------------------------------------------------------------------------
    if rank != 0:
        torch.save = lambda *a, **kw: None
        sys.stdout = open(os.devnull, "w")
        os.environ.setdefault("TQDM_DISABLE", "1")
    return rank, world


def _install_trace():
    """B200GS_RUN_TRACE=1: wall-clock totals of the host-side calls a training iteration makes (where does the host wait?),
    printed when the script ends.  A debugging aid; changes nothing else."""
    import atexit
    import time
    import torch
    acc = {}

    def wrap(owner, name, label=None):
        fn = getattr(owner, name)
        label = label or f"{getattr(owner, '__name__', owner)}.{name}"

        def timed(*a, **kw):
            t0 = time.perf_counter()
            try:
                return fn(*a, **kw)
            finally:
                e = acc.setdefault(label, [0, 0.0])
                e[0] += 1
                e[1] += time.perf_counter() - t0
        setattr(owner, name, timed)
    wrap(torch.cuda, "empty_cache")
    wrap(torch.cuda, "synchronize")
    wrap(torch.Tensor, "backward", "Tensor.backward")
    wrap(torch.Tensor, "item", "Tensor.item")
    wrap(torch.Tensor, "cpu", "Tensor.cpu")
    wrap(torch.Tensor, "to", "Tensor.to")
    wrap(torch.nn.utils, "clip_grad_norm_")
    wrap(torch.optim.Adam, "step", "Adam.step")
    wrap(torch.optim.Adam, "__init__", "Adam.__init__")
    for mod, attr in (("gaussian_splatting.render", "render"), ("gaussian_splatting.spherical_harmonics", "evaluate_sh"),
                      ("gaussian_splatting.gaussian", "build_sigma_from_params"), ("gaussian_splatting.losses", "compute_loss")):
        wrap(sys.modules[mod], attr, attr)
    t_start = time.perf_counter()

    def report():
        total = time.perf_counter() - t_start
        print(f"[b200gs.run trace] wall {total:.3f} s", file=sys.stderr)
        for k, (n, t) in sorted(acc.items(), key=lambda kv: -kv[1][1]):
            print(f"[b200gs.run trace] {k:28s} calls {n:6d}  total {t:8.3f} s  mean {t / max(n, 1) * 1e3:8.3f} ms", file=sys.stderr)
    atexit.register(report)


def _install_timeline():
    """B200GS_RUN_TRACE=2: start/end timestamps of four calls per training iteration (backward, clip_grad_norm_,
    optimizer.step, empty_cache) - the lightest instrumentation that still tells WHEN the host waits.  Printed at exit
    as per-iteration totals in windows of 10 iterations."""
    import atexit
    import time
    import torch
    ev = []          # (label, t0, t1)

    def wrap(owner, name, label):
        fn = getattr(owner, name)

        def timed(*a, **kw):
            t0 = time.perf_counter()
            try:
                return fn(*a, **kw)
            finally:
                ev.append((label, t0, time.perf_counter()))
        setattr(owner, name, timed)
    wrap(torch.Tensor, "backward", "backward")
    wrap(torch.nn.utils, "clip_grad_norm_", "clip")
    wrap(torch.optim.Adam, "step", "step")
    wrap(torch.cuda, "empty_cache", "empty_cache")

    def report():
        steps = [e for e in ev if e[0] == "step"]
        print(f"[b200gs.run timeline] {len(steps)} iterations", file=sys.stderr)
        for w in range(0, len(steps), 10):
            lo = steps[w][1]
            hi = steps[min(w + 10, len(steps)) - 1][2]
            n = min(w + 10, len(steps)) - w
            inside = {}
            for label, t0, t1 in ev:
                if lo <= t0 <= hi:
                    inside[label] = inside.get(label, 0.0) + (t1 - t0)
            print(f"[b200gs.run timeline] it {w:4d}-{w + n - 1:4d}: {(hi - lo) / n * 1e3:8.3f} ms/it  " +
                  "  ".join(f"{k} {v / n * 1e3:7.3f}" for k, v in sorted(inside.items())), file=sys.stderr)
    atexit.register(report)


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    dp = bool(argv) and argv[0] == "--dp"
    if dp:
        argv = argv[1:]
    if not argv:
        print(__doc__)
        return 2
    script = os.path.abspath(argv[0])
    root = os.environ.get("B200GS_REFERENCE_ROOT") or os.path.dirname(os.path.dirname(script))
    if root not in sys.path:
        sys.path.insert(0, root)
    from .install import install
    install(optimizer=os.environ.get("B200GS_PATCH_ADAM", "0") == "1")     # first: --dp wraps whatever Adam is installed
    world = 1
    if dp:
        _, world = _setup_dp()
    if os.environ.get("B200GS_RUN_TRACE") == "1":
        _install_trace()
    elif os.environ.get("B200GS_RUN_TRACE") == "2":
        _install_timeline()
    sys.argv = [script] + argv[1:]
    try:
        runpy.run_path(script, run_name="__main__")
    finally:
        if dp and world > 1:
            import torch.distributed as dist
            if dist.is_initialized():
                dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
