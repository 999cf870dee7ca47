#!/usr/bin/env python
"""Benchmark of the render path: render frames/s (+ train iterations/s) on the BASELINE.json headline
workload - 1M Gaussians, 1920x1080, SH degree 3, synthetic seeded scene.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200gs|reference] [--mode render|train|tile_rows]
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

A "step" is one pass of the hot path over one view:
  render step (default mode) = evaluate_sh + render under no_grad   (the reference's own timed region,
                scripts/render_trained.py:337-349; build_sigma is outside, as at :192)
  train step  (--mode train) = the whole iteration of scripts/train.py:463-538: build_sigma + evaluate_sh + render +
                L1/SSIM loss + backward + SUM of the gradients over the ranks + clip_grad_norm_(pos) + Adam
  tile_rows   (--mode tile_rows) = ONE 6M-Gaussian 3840x2160 frame (BASELINE.json configs[4]) split into bands of tile
                rows over the ranks (strong scaling; b200gs.dist.TileRowRenderer)
Multi-GPU: one process per GPU; render frames are sharded round-robin (no collective), training views are
data-parallel (weak scaling: per-GPU work fixed).  Rank 0 prints ONE JSON line.  The default mode also measures the
other two on a few steps and reports them inside `config` (the keys a driver keeps).

--impl reference times the UNMODIFIED reference (staged under baseline/_ref by baseline/stage_reference.py; its own
`evaluate_sh` + `render`, torch CPU, all host cores) on full headline frames; when nothing is staged it falls back to
the oracle port and says so (`cpu_baseline.kind`).
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "3d-gaussian-splatting-for-novel-view-synthesis_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402

WORKLOAD = dict(name="1M Gaussians, 1920x1080, SH3, log-scale -5.5 (headline)", n=1_000_000, W=1920, H=1080,
                log_scale=-5.5, sh_degree=3, n_views=16, seed=0)
PARAMS = ("pos", "scale_raw", "q_raw", "opacity_raw", "f_dc", "f_rest")
METRIC = "render FPS (+ train it/s), 1M Gaussians 1080p SH3"


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json)", float(p.get("sm_max_mhz", 1965.0))
    return 6650.0, "fallback (B200_PROFILING.md)", 1965.0


# ---------------------------------------------------------------------------------------------------
# clocks sampler (nvidia-smi during the timed region)
# ---------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:  # noqa: BLE001
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


class Deadline:
    """A rank-local deadline around optional work: `start()` arms it; if `cancel()` does not come first, `on_expire()`
    runs on a timer thread and the process ends with exit code 0 (`exit_fn`).  Exactly one side wins: `cancel()`
    returns False once the timer thread has taken over, and the caller must then leave the printing to it."""

    def __init__(self, seconds: float, on_expire, exit_fn=os._exit):
        self._lock = threading.Lock()
        self._taken = False
        self._on_expire, self._exit = on_expire, exit_fn
        self._timer = threading.Timer(max(0.0, seconds), self._fire)
        self._timer.daemon = True
        self._started = False

    def start(self):
        self._started = True
        self._timer.start()

    def _fire(self):
        with self._lock:
            if self._taken:
                return
            self._taken = True
        try:
            self._on_expire()
        finally:
            self._exit(0)

    def cancel(self) -> bool:
        with self._lock:
            if self._taken:
                return False
            self._taken = True
        if self._started:
            self._timer.cancel()
        return True


# ---------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle port on the host cores, bounded sample
# ---------------------------------------------------------------------------------------------------
def cpu_scene():
    from oracle import gs_oracle as O
    wl = WORKLOAD
    sc = O.make_scene(wl["n"], seed=wl["seed"], log_scale=wl["log_scale"], sh_degree=wl["sh_degree"])
    sc["sigma"] = O.build_sigma_from_params(sc["scale_raw"], sc["q_raw"])
    cams = [O.make_camera(wl["W"], wl["H"], view=v, n_views=wl["n_views"]) for v in range(wl["n_views"])]
    return sc, cams


def staged_reference():
    """The unmodified reference package, imported from baseline/_ref (None when nothing is staged)."""
    root = os.path.join(ROOT, "baseline", "_ref")
    if not os.path.exists(os.path.join(root, "gaussian_splatting", "__init__.py")):
        return None
    if root not in sys.path:
        sys.path.insert(0, root)
    import gaussian_splatting  # noqa: PLC0415
    return gaussian_splatting


def run_reference(args):
    """The reference's own CPU code path on the headline frame, timed the way its script times a frame
    (scripts/render_trained.py:337-349: evaluate_sh + render per frame; build_sigma outside, :192).  Every step is one
    FULL frame; the run stops after `--steps` frames or when the time budget (B200GS_REF_BUDGET_S, default 240 s) is
    used up, whichever comes first, and reports the steps it actually executed."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sc, cams = cpu_scene()
    budget = float(os.environ.get("B200GS_REF_BUDGET_S", "240"))
    ref = None if os.environ.get("B200GS_REF_FORCE_PORT") == "1" else staged_reference()
    if ref is not None:
        kind, what = "reference", "unmodified reference (baseline/_ref): gaussian_splatting.evaluate_sh + render, torch CPU"
        ev, rd = ref.evaluate_sh, ref.render
        with torch.no_grad():
            sc["sigma"] = ref.build_sigma_from_params(sc["scale_raw"], sc["q_raw"])
    else:
        from oracle import gs_oracle as O
        kind, what = "port", "oracle port (oracle/gs_oracle.py; no reference staged under baseline/_ref)"
        ev, rd = O.evaluate_sh, O.render

    def frame(cam, s):
        with torch.no_grad():
            col = ev(s["f_dc"], s["f_rest"], s["pos"], cam["c2w"])
            return rd(s["pos"], col, s["opacity_raw"], s["sigma"], cam["c2w"], cam["H"], cam["W"],
                      cam["fx"], cam["fy"], cam["cx"], cam["cy"], pix_guard=32, chi_square_clip=6.25, alpha_cutoff=1 / 128.)
    t_begin = time.perf_counter()
    # warm-up: the same code path on a 1/16-size frame (thread pools, allocator), then - only if the budget allows a
    # full frame beyond the timed ones - nothing more: a CPU frame has no caches worth warming at 5-20 s per frame
    from oracle import gs_oracle as O
    small = O.make_scene(WORKLOAD["n"] // 16, seed=1, log_scale=WORKLOAD["log_scale"])
    small["sigma"] = (ref.build_sigma_from_params if ref is not None else O.build_sigma_from_params)(small["scale_raw"], small["q_raw"])
    warm = 0
    for _ in range(max(1, min(args.warmup, 2))):
        frame(O.make_camera(WORKLOAD["W"] // 4, WORKLOAD["H"] // 4), small)
        warm += 1
    times = []
    for i in range(max(1, args.steps)):
        t0 = time.perf_counter()
        frame(cams[i % len(cams)], sc)
        times.append(time.perf_counter() - t0)
        if time.perf_counter() - t_begin + times[-1] > budget:
            break
    sec = sum(times) / len(times)
    fps = 1.0 / sec
    line = {"impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
            "steps": len(times), "warmup": warm, "ms_per_step": sec * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD["name"], "device": "host CPU", "steps_requested": args.steps,
                       "warmup_note": f"{warm} warm-up frame(s) of a 1/16-size scene at 480x270 (a full CPU frame "
                                      "takes seconds; the timed steps are full headline frames)",
                       "timed_region": "per frame: evaluate_sh + render (scripts/render_trained.py:337-349)",
                       "frame_seconds": [round(t, 3) for t in times]},
            "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "torch_threads": torch.get_num_threads(),
                             "kind": kind, "sample": f"{len(times)} full headline frame(s), {what}"},
            "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)
    return 0


# ---------------------------------------------------------------------------------------------------
# b200gs arm
# ---------------------------------------------------------------------------------------------------
def algorithmic_bytes(N, V, I, P, tiles, S=0):
    """Compulsory HBM bytes per kernel group for one view (each stage reads its inputs once and writes its
    outputs once; a sort counts as one read + one write of its pairs whatever the pass count) - DESIGN.md.
    S = number of (supertile, Gaussian) pairs."""
    return {
        "emit_super": 12 * N + 8 * V + 8 * S,
        "super_sort": 8 * S + 8 * S,
        "split_tiles": 2 * (4 * S + 8 * S) + 4 * I + 8 * tiles,
        "preprocess_fwd": 236 * N + 68 * V + 12 * (N - V),
        "evaluate_sh": 204 * N + 12 * N,
        "depth_sort": 4 * N + 8 * N,
        "scan": 4 * N + 4 * N + 4 * N,
        "blend_fwd": 4 * I + 36 * I + 12 * P + 8 * P,
        "blend_bwd": 4 * I + 36 * I + 36 * I + 12 * P + 8 * P,
        "preprocess_bwd": 236 * N + 48 * N + 236 * N,
        "build_sigma": 28 * N + 36 * N,
        "adam_step": 28 * 59 * N,                 # (param, grad, exp_avg, exp_avg_sq) in, (param, exp_avg, exp_avg_sq) out
        "clip_grad_norm": 12 * N,                 # norm pass over pos.grad (the scale pass only runs when clipping)
        "l1_ssim_fwd": 24 * P + 36 * P,           # pred + target in, three partial-derivative maps out
        "l1_ssim_bwd": 36 * P + 24 * P + 12 * P,  # maps + pred + target in, dL/dpred out
    }


TILE_ROWS_WORKLOAD = dict(name="6M Gaussians, 3840x2160, SH3, log-scale -6.0 (BASELINE.json configs[4])", n=6_000_000,
                          W=3840, H=2160, log_scale=-6.0, seed=0)
LR0 = {"pos": 1.6e-4 * 0.01, "opacity_raw": 0.05, "f_dc": 2.5e-3, "f_rest": 2.5e-3 / 20.0, "scale_raw": 5e-3, "q_raw": 1e-3}


def ncu_counts():
    """Per-launch counters of the dominant kernels on view 0 of the headline workload, from the committed
    `ncu --set full` capture (profiles/ncu_traffic.json): instruction counts are a property of the seeded scene and the
    built kernels, so they are read from the capture, while every TIME in the line is measured live."""
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    return json.load(open(path)) if os.path.exists(path) else {}


class Env:
    """Process-wide state shared by the three modes."""

    def __init__(self, args):
        import torch.distributed as dist
        import b200gs
        self.dist, self.gs = dist, b200gs
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device; the b200gs arm has no CPU fallback (use --impl reference)")
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)
        self.lib = b200gs.load_library()
        os.environ.setdefault("B200GS_CAPACITY_MODE", "speculative")
        self.K, self.Wm = args.steps, args.warmup

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(self, x: float) -> float:
        if self.world == 1:
            return x
        t = torch.tensor([x], device=self.dev, dtype=torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def timed(self, fn, steps, warm):
        """`steps` calls of fn between two CUDA events on the current stream, barrier + synchronize on both sides,
        max over ranks -> (milliseconds, kernels launched by libb200gs)."""
        for i in range(warm):
            fn(i)
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = self.lib.b200gs_kernel_launch_count()
        e0.record()
        for i in range(steps):
            fn(warm + i)
        e1.record()
        self.barrier()
        return self.max_over_ranks(e0.elapsed_time(e1)), self.lib.b200gs_kernel_launch_count() - l0

    def finish(self):
        if self.world > 1:
            try:
                self.dist.destroy_process_group()
            except Exception:                     # noqa: BLE001 - the line is out
                pass


def headline_scene(env):
    from oracle import gs_oracle as O   # scene generator + cpu_baseline leg only (never on the product path)
    wl = WORKLOAD
    sc_cpu = O.make_scene(wl["n"], seed=wl["seed"], log_scale=wl["log_scale"], sh_degree=wl["sh_degree"])
    sc = {k: v.to(env.dev) for k, v in sc_cpu.items()}
    cams = [O.make_camera(wl["W"], wl["H"], view=v, n_views=wl["n_views"]) for v in range(wl["n_views"])]
    return sc_cpu, sc, cams


# ---- the whole training iteration (scripts/train.py:463-538), data-parallel over the views --------------------------
def measure_train_iteration(env, sc, cams, steps, warm, want_nccl=True):
    """Returns {"peer": ms or None, "nccl": ms or None, ...}: the full iteration with the optimizer half over NVLink
    peer memory (b200gs.PeerAdam; world 1: the same kernels on a local area) and with NCCL (one flat bucket ->
    one all_reduce) + clip + FusedAdam.  Runs last in a process: it moves the parameters."""
    gs, dev, world, rank = env.gs, env.dev, env.world, env.rank
    from b200gs.dist import GradBucket
    H, W = cams[0]["H"], cams[0]["W"]
    intr = cams[0]
    c2w_dev = [c["c2w"].to(dev) for c in cams]
    view_of = lambda step: (rank + world * step) % len(cams)
    target_dev = torch.rand(H, W, 3, generator=torch.Generator().manual_seed(4321)).to(dev)
    out = {"peer": None, "nccl": None, "peer_error": None, "peer_transport": None}

    def fwd_bwd(leaves, i):
        c2w = c2w_dev[view_of(i)]
        sg = gs.build_sigma_from_params(leaves["scale_raw"], leaves["q_raw"])
        col = gs.evaluate_sh(leaves["f_dc"], leaves["f_rest"], leaves["pos"], c2w)
        img = gs.render(leaves["pos"], col, leaves["opacity_raw"], sg, c2w, H, W, intr["fx"], intr["fy"], intr["cx"], intr["cy"])
        loss, _ = gs.compute_loss_tensors(img, target_dev)
        (loss / world).backward()

    if want_nccl:
        leaves = {k: sc[k].clone().requires_grad_(True) for k in PARAMS}
        opt = gs.FusedAdam([{"params": [leaves[k]], "lr": LR0[k], "name": k} for k in PARAMS], lr=1e-3, eps=1e-15)
        bucket = GradBucket([leaves[k] for k in PARAMS]) if world > 1 else None

        def it_nccl(i):
            opt.zero_grad(set_to_none=True)
            fwd_bwd(leaves, i)
            if bucket is not None:
                bucket.allreduce()
            gs.clip_grad_norm_(leaves["pos"], max_norm=1.0)
            opt.step()
        out["nccl"], out["launches"] = env.timed(it_nccl, steps, max(warm, 4))
        if bucket is not None:
            from b200gs import ops
            ops.unregister_grad_sinks(bucket.params)
        del opt, leaves, bucket
    try:
        leaves = {k: sc[k].clone().requires_grad_(True) for k in PARAMS}
        opt_p = gs.PeerAdam([{"params": [leaves[k]], "lr": LR0[k], "name": k} for k in PARAMS], lr=1e-3, eps=1e-15,
                            clip_params=[leaves["pos"]], max_norm=1.0)
        out["peer_transport"] = opt_p.area.transport + ("+nvls" if opt_p.area.c_group.multicast else "")

        def it_peer(i):
            opt_p.zero_grad(set_to_none=True)
            fwd_bwd(leaves, i)
            opt_p.step()
        out["peer"], launches = env.timed(it_peer, steps, max(warm, 4))
        out.setdefault("launches", launches)
    except Exception as e:                    # noqa: BLE001 - reported in the JSON line
        out["peer_error"] = f"{type(e).__name__}: {e}"[:300]
    return out


# ---- one large frame split into tile-row bands (BASELINE.json configs[4]) --------------------------------------------
def measure_tile_rows(env, steps, warm):
    from oracle import gs_oracle as O
    from b200gs.dist import TileRowRenderer
    gs, dev = env.gs, env.dev
    wl = TILE_ROWS_WORKLOAD
    sc = {k: v.to(dev) for k, v in O.make_scene(wl["n"], seed=wl["seed"], log_scale=wl["log_scale"]).items()}
    cam = O.make_camera(wl["W"], wl["H"], view=0, n_views=16)
    c2w = cam["c2w"].to(dev)
    with torch.no_grad():
        sigma = gs.build_sigma_from_params(sc["scale_raw"], sc["q_raw"])
        tr = TileRowRenderer(wl["H"], wl["W"], dev)

        defer = os.environ.get("B200GS_TILE_ROWS_DEFER", "1") != "0"

        def frame(_i):
            # frames are queued back to back: the one host wait of a frame (for its intersection counters) moves to the
            # start of the next call, so the host runs ahead of the GPU (TileRowRenderer.render, defer_check)
            col = gs.evaluate_sh(sc["f_dc"], sc["f_rest"], sc["pos"], c2w)
            return tr.render(sc["pos"], col, sc["opacity_raw"], sigma, c2w, cam["fx"], cam["fy"], cam["cx"], cam["cy"],
                             defer_check=defer)
        weights = None
        if env.world > 1:      # balance the bands by the intersections per tile row of this view
            col = gs.evaluate_sh(sc["f_dc"], sc["f_rest"], sc["pos"], c2w)
            weights = tr.row_weights(sc["pos"], col, sc["opacity_raw"], sigma, c2w, cam["fx"], cam["fy"], cam["cx"], cam["cy"])
            tr.set_weights(weights)
        ms, launches = env.timed(frame, steps, max(warm, 3))
        tr.finish()
        redone = tr.redone
        img = frame(0)
        tr.finish()
        torch.cuda.synchronize()
        checksum = float(img.double().sum()) if env.rank == tr.root else 0.0
        env.barrier()      # (the root's checksum must not delay its first barrier of the profiled frames below)
        stats = dict(V=tr.last_frame.n_visible, I_band=tr.last_frame.n_isect)
        # per kernel group, this rank (CUDA events inside the library around every group; a separate, untimed pass)
        lib = env.lib
        lib.b200gs_profile_enable(1)
        for i in range(5):
            frame(i)
        tr.finish()
        torch.cuda.synchronize()
        ms_buf, call_buf = (ctypes.c_float * 32)(), (ctypes.c_int32 * 32)()
        k = lib.b200gs_profile_collect(ms_buf, call_buf, 32)
        lib.b200gs_profile_enable(0)
        regions = {lib.b200gs_profile_region_name(r).decode(): round(1e3 * ms_buf[r] / 5, 1) for r in range(k) if call_buf[r]}
        if env.world > 1:                      # every rank's table: the slowest rank sets the frame time
            every = [None] * env.world
            env.dist.all_gather_object(every, regions)
            regions = {"rank%d" % q: t for q, t in enumerate(every)}
        env.barrier()
        # measured alternatives (one knob changed at a time; every rank reads the same environment)
        variants = {}
        if os.environ.get("B200GS_TILE_ROWS_SWEEP") and env.world > 1:
            was_split = tr.split_records
            for name, knob, val in (("records_on_a_side_stream", "split", True),
                                    ("one_pass_with_per_warp_runs", "B200GS_ROUTE_WRITE", "warp"),
                                    ("blend_strided_pixel_stores", "B200GS_ROW_STORES", "0"),
                                    ("host_waits_for_counters_every_frame", "defer", False)):
                if knob == "split":
                    tr.split_records = val
                elif knob == "defer":
                    defer = val
                else:
                    os.environ[knob] = val
                    tr.split_records = tr.split_records and knob != "B200GS_ROUTE_WRITE"
                vms, _ = env.timed(frame, 20, 3)
                tr.finish()
                variants[name] = round(vms / 20, 4)
                tr.split_records = was_split
                defer = os.environ.get("B200GS_TILE_ROWS_DEFER", "1") != "0"
                if knob.startswith("B200GS_"):
                    del os.environ[knob]
    del sc, sigma
    torch.cuda.empty_cache()
    return {"ms_per_frame": ms / steps, "frames_per_s": steps / (ms * 1e-3), "launches": int(launches), "bands": tr.bands,
            "balanced": weights is not None, "checksum_root": checksum, "routed": bool(tr.routed), "records_on_side_stream": bool(tr.split_records), "deferred_check": defer,
            "frames_redone": int(redone), "regions_us_per_frame": regions, "variants_ms_per_frame": variants, **stats}


def legs_at_deadline(extras: dict):
    """What the default line reports for the two optional legs when their deadline expires: a finished leg keeps its
    numbers, an unfinished one is marked (`extras` = {"tile_rows": result or None, "train": measure_train_iteration's
    dict, all None until it returns})."""
    tr_leg, tn_leg = extras["tile_rows"], dict(extras["train"])
    if tr_leg is None:
        tr_leg = {"error": "deadline: the tile-row leg did not finish (B200GS_BENCH_EXTRAS_DEADLINE_S)"}
    if tn_leg.get("peer") is None and tn_leg.get("nccl") is None:
        tn_leg["peer_error"] = "deadline: the training leg did not finish (B200GS_BENCH_EXTRAS_DEADLINE_S)"
    return tr_leg, tn_leg


def run_b200gs(args):
    env = Env(args)
    if args.mode == "train":
        return run_mode_train(env, args)
    if args.mode == "tile_rows":
        return run_mode_tile_rows(env, args)
    return run_mode_render(env, args)


def base_line(env, metric, value, unit, ms_total, steps, scaling, config):
    return {"metric": metric, "value": value, "unit": unit, "n_gpus": env.world, "steps": steps, "warmup": env.Wm,
            "ms_per_step": ms_total / steps, "higher_is_better": True, "scaling": scaling, "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": config}


def run_mode_train(env, args):
    sc_cpu, sc, cams = headline_scene(env)
    sampler = ClockSampler(env.local)
    if env.rank == 0:
        sampler.start()
    r = measure_train_iteration(env, sc, cams, env.K, env.Wm, want_nccl=True)
    clocks = sampler.stop() if env.rank == 0 else None
    if env.rank == 0:
        best = min(x for x in (r["peer"], r["nccl"]) if x is not None)
        line = base_line(env, "train views/s (whole iteration of scripts/train.py:463-538), 1M Gaussians 1080p SH3",
                         env.world * env.K / (best * 1e-3), "views/s", best, env.K, "weak",
                         {"workload": WORKLOAD["name"], "parallelism": f"views data-parallel over {env.world} rank(s), "
                          "Gaussians replicated; gradients summed over the ranks every iteration",
                          "it_per_s": env.K / (best * 1e-3),
                          "ms_per_iteration_peer_adam": None if r["peer"] is None else r["peer"] / env.K,
                          "ms_per_iteration_nccl_bucket": None if r["nccl"] is None else r["nccl"] / env.K,
                          "peer_transport": r["peer_transport"], "peer_error": r["peer_error"],
                          "l2_policy": "inputs larger than L2: 236 MB of parameters + 236 MB of gradients per step"})
        line.update(gpu_launches=int(r.get("launches", 0)), clocks=clocks)
        emit(line)
    env.finish()
    return 0


def run_mode_tile_rows(env, args):
    sampler = ClockSampler(env.local)
    if env.rank == 0:
        sampler.start()
    r = measure_tile_rows(env, env.K, env.Wm)
    clocks = sampler.stop() if env.rank == 0 else None
    if env.rank == 0:
        line = base_line(env, "4K single-frame render, 6M Gaussians, tile rows sharded over the GPUs", r["frames_per_s"],
                         "frames/s", r["ms_per_frame"] * env.K, env.K, "strong",
                         {"workload": TILE_ROWS_WORKLOAD["name"], "bands": r["bands"], "balanced_by_row_weights": r["balanced"],
                          "V": r["V"], "I_band_rank0": r["I_band"], "checksum_root": r["checksum_root"],
                          "routed": r["routed"], "records_on_side_stream": r["records_on_side_stream"],
                          "deferred_check": r["deferred_check"], "frames_redone": r["frames_redone"],
                          "kernel_groups_us_per_frame": r["regions_us_per_frame"],
                          "variants_ms_per_frame": r["variants_ms_per_frame"],
                          "parallelism": (f"one band of tile rows per rank ({env.world}); every rank projects 1/{env.world} of "
                                          "the Gaussians and routes the splat records to the bands over NVLink peer memory "
                                          "(sort-middle), bands stored straight into rank 0's frame buffer, no collective")
                          if r["routed"] else
                                         (f"one band of tile rows per rank ({env.world}); Gaussians replicated; bands stored "
                                          "straight into rank 0's frame buffer over NVLink peer memory, no collective"),
                          "l2_policy": "inputs larger than L2: 1.4 GB of parameters per frame"})
        line.update(gpu_launches=r["launches"], clocks=clocks)
        emit(line)
    env.finish()
    return 0


def run_mode_render(env, args):
    import b200gs
    from b200gs import _lib, ops
    from oracle import gs_oracle as O   # scene generator + cpu_baseline leg only (never on the product path)
    world, rank, dev, lib = env.world, env.rank, env.dev, env.lib
    barrier, max_over_ranks, timed = env.barrier, env.max_over_ranks, env.timed
    wl = WORKLOAD
    H, W, K, Wm = wl["H"], wl["W"], env.K, env.Wm
    sc_cpu, sc, cams = headline_scene(env)
    c2w_dev = [c["c2w"].to(dev) for c in cams]
    c2w_pin = [c["c2w"].clone().pin_memory() for c in cams]
    intr = cams[0]
    view_of = lambda step: (rank + world * step) % len(cams)

    def render_step(c2w):
        colors = b200gs.evaluate_sh(sc["f_dc"], sc["f_rest"], sc["pos"], c2w)
        return b200gs.render(sc["pos"], colors, sc["opacity_raw"], sigma, c2w, H, W, intr["fx"], intr["fy"],
                             intr["cx"], intr["cy"])

    sampler = ClockSampler(env.local)
    if rank == 0:
        sampler.start()

    # ---- render: device-resident inputs ("value") --------------------------------------------------------
    with torch.no_grad():
        sigma = b200gs.build_sigma_from_params(sc["scale_raw"], sc["q_raw"])
        torch.cuda.synchronize()
        ms_render_1s, _ = timed(lambda i: render_step(c2w_dev[view_of(i)]), K, Wm)
        # the reference's own timed region (scripts/render_trained.py:337-349): synchronize, start the clock, evaluate_sh +
        # render, synchronize, stop the clock - what an UNCHANGED script measures per frame
        for i in range(Wm):
            render_step(c2w_dev[view_of(i)])
        barrier()
        t_sync = 0.0
        for i in range(K):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            render_step(c2w_dev[view_of(Wm + i)])
            torch.cuda.synchronize()
            t_sync += time.perf_counter() - t0
        s_sync_per_frame = max_over_ranks(t_sync)
        # ---- the headline: frames software-pipelined by b200gs.RenderPipeline - project + binning of frame i+1 on a
        #      high-priority stream while frame i is blended on a second stream (an orbit render; frames independent)
        pipe = b200gs.RenderPipeline(dev)

        def pipe_submit(c2w):
            colors = b200gs.evaluate_sh(sc["f_dc"], sc["f_rest"], sc["pos"], c2w)
            return pipe.submit(sc["pos"], colors, sc["opacity_raw"], sigma, c2w, H, W, intr["fx"], intr["fy"],
                               intr["cx"], intr["cy"])
        pending = []

        def pipe_step(c2w, deliver=None):
            """Queues a frame and collects the one queued before it (one frame of lag keeps the host off the
            critical path); `deliver(image, done_event)` consumes a finished frame."""
            pending.append(pipe_submit(c2w))
            if len(pending) > 1:
                t = pending.pop(0)
                img = pipe.result(t)
                if deliver is not None:
                    deliver(img, pipe.done_event(t))

        def pipe_flush(deliver=None):
            while pending:
                t = pending.pop(0)
                img = pipe.result(t)
                if deliver is not None:
                    deliver(img, pipe.done_event(t))
            pipe.synchronize()

        def timed_pipe(steps, warm):
            for i in range(warm):
                pipe_step(c2w_dev[view_of(i)])
            pipe_flush()
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            l0 = lib.b200gs_kernel_launch_count()
            main = torch.cuda.current_stream(dev)
            e0.record(main)
            pipe.wait_event(e0)
            for i in range(steps):
                pipe_step(c2w_dev[view_of(warm + i)])
            pipe_flush()
            pipe.join(main)
            e1.record(main)
            barrier()
            return max_over_ranks(e0.elapsed_time(e1)), lib.b200gs_kernel_launch_count() - l0
        # the pipeline's first ~25 frames allocate its workspace ring and image buffers: untimed warm-up covers them
        ms_render, launches_render = timed_pipe(K, max(Wm, 30))
        # ---- render e2e: pose from pinned host memory in, finished frame to pinned host memory out, every step.  The
        #      D2H copy of a finished frame runs on a copy stream, waiting for exactly that frame (done_event); a ring
        #      of pinned images, each reused once its copy has completed.  Two deliveries are measured:
        #      uint8 frames (b200gs.to_uint8 on the device = the conversion the reference scripts apply to every frame
        #      they keep, render_trained.py:357 / inference.py:117; bit-identical, 3 B/pixel on PCIe) -> `e2e`, and
        #      the raw fp32 image (12 B/pixel) -> `e2e_f32_frames`.
        copy_stream = torch.cuda.Stream(dev)
        n_ring = 3
        copy_done = [torch.cuda.Event() for _ in range(n_ring)]
        for e in copy_done:
            e.record(copy_stream)
        delivered = [0]

        def make_deliver(pins, convert):
            def deliver(img, done):
                k = delivered[0] % n_ring
                delivered[0] += 1
                copy_done[k].synchronize()                   # the frame that used this pinned image has been delivered
                with torch.cuda.stream(copy_stream):
                    copy_stream.wait_event(done)
                    img.record_stream(copy_stream)
                    pins[k].copy_(convert(img), non_blocking=True)
                    copy_done[k].record(copy_stream)
            return deliver

        def e2e_step(i, deliver):
            with torch.cuda.stream(pipe.next_front_stream):
                c2w = c2w_pin[view_of(i)].to(dev, non_blocking=True)
            pipe_step(c2w, deliver)

        def timed_e2e(deliver):
            for i in range(Wm):
                e2e_step(i, deliver)
            pipe_flush(deliver)
            copy_stream.synchronize()
            barrier()
            t0 = time.perf_counter()
            for i in range(K):
                e2e_step(Wm + i, deliver)
            pipe_flush(deliver)
            copy_stream.synchronize()
            barrier()
            return max_over_ranks(time.perf_counter() - t0)
        f32_pin = [torch.empty((H, W, 3), dtype=torch.float32).pin_memory() for _ in range(n_ring)]
        s_e2e_f32 = timed_e2e(make_deliver(f32_pin, lambda im: im))
        del f32_pin
        u8_pin = [torch.empty((H, W, 3), dtype=torch.uint8).pin_memory() for _ in range(n_ring)]
        s_e2e_u8 = timed_e2e(make_deliver(u8_pin, b200gs.to_uint8))

    # ---- train: fwd + bwd with a weighted-sum loss (per-kernel table) ------------------------------------------------
    leaves = {k: sc[k].clone().requires_grad_(True) for k in PARAMS}
    target_dev = torch.rand(H, W, 3, generator=torch.Generator().manual_seed(4321)).to(dev)

    def train_step_loss(i):
        for p in leaves.values():
            p.grad = None
        c2w = c2w_dev[view_of(i)]
        sg = b200gs.build_sigma_from_params(leaves["scale_raw"], leaves["q_raw"])
        col = b200gs.evaluate_sh(leaves["f_dc"], leaves["f_rest"], leaves["pos"], c2w)
        img = b200gs.render(leaves["pos"], col, leaves["opacity_raw"], sg, c2w, H, W, intr["fx"], intr["fy"],
                            intr["cx"], intr["cy"])
        loss, _ = b200gs.compute_loss_tensors(img, target_dev)
        loss.backward()
        return loss
    ms_train_loss, launches_train = timed(lambda i: train_step_loss(i), K, Wm)
    clocks = sampler.stop() if rank == 0 else None

    # ---- per-kernel CUDA-event times (separate pass with an event pair around every kernel group) ------------
    lib.b200gs_profile_enable(1)
    with torch.no_grad():
        for i in range(K):
            render_step(c2w_dev[view_of(i)])
    torch.cuda.synchronize()
    nreg = 32
    ms_buf, call_buf = (ctypes.c_float * nreg)(), (ctypes.c_int32 * nreg)()

    def collect():
        n_regions = lib.b200gs_profile_collect(ms_buf, call_buf, nreg)
        return {lib.b200gs_profile_region_name(r).decode(): (ms_buf[r] / max(1, call_buf[r]), call_buf[r])
                for r in range(n_regions) if call_buf[r]}
    fwd_regions = collect()
    for i in range(K):
        train_step_loss(i)
    torch.cuda.synchronize()
    train_regions = collect()
    opt = b200gs.FusedAdam([{"params": [leaves[k]], "lr": LR0[k], "name": k} for k in PARAMS], lr=1e-3, eps=1e-15)
    for i in range(8):
        if i == 4:                     # first steps allocate the optimizer state: not representative
            torch.cuda.synchronize()
            lib.b200gs_profile_collect(ms_buf, call_buf, nreg)
        opt.zero_grad(set_to_none=True)
        train_step_loss(i)
        b200gs.clip_grad_norm_(leaves["pos"], max_norm=1.0)
        opt.step()
    torch.cuda.synchronize()
    opt_regions = {k: v for k, v in collect().items() if k in ("adam_step", "clip_grad_norm")}
    lib.b200gs_profile_enable(0)
    del opt, leaves
    # ---- frame statistics of view 0 (V, I) for the byte model + the GPU side of the parity check --------------------
    with torch.no_grad():
        g, keep = ops._gaussians(sc["pos"], sc["opacity_raw"], sc["scale_raw"], sc["q_raw"], None, sc["f_dc"],
                                 sc["f_rest"], None)
        cfg = ops.RenderConfig(H=H, W=W, fx=intr["fx"], fy=intr["fy"], cx=intr["cx"], cy=intr["cy"])
        fr = ops.Frame(g, keep, cfg, c2w_dev[0], dev)
        img_view0 = fr.render("sync")
        fr.refresh_stats()
    V, I, N, P, S = fr.n_visible, fr.n_isect, wl["n"], H * W, fr.n_super
    tiles = ((W + 15) // 16) * ((H + 15) // 16)

    # ---- e2e through the C ABI with HOST buffers (whole Gaussian set uploaded every call) -------------------------
    host_fps = None
    if rank == 0:
        hp = {k: sc_cpu[k].contiguous().pin_memory() for k in sc_cpu}
        gh = _lib.Gaussians(n=N, pos=hp["pos"].data_ptr(), opacity_raw=hp["opacity_raw"].data_ptr(),
                            scale_raw=hp["scale_raw"].data_ptr(), q_raw=hp["q_raw"].data_ptr(), sigma=None,
                            f_dc=hp["f_dc"].data_ptr(), f_rest=hp["f_rest"].data_ptr(), color=None)
        camc = cfg.to_c(c2w_pin[0])
        img_host = torch.empty((H, W, 3), dtype=torch.float32).pin_memory()
        st = _lib.FrameStats()
        for i in range(2):
            _lib.check(lib.b200gs_render_host(ctypes.byref(gh), ctypes.byref(camc), c2w_pin[i].data_ptr(),
                                              img_host.data_ptr(), ctypes.byref(st)), "render_host")
        t0 = time.perf_counter()
        reps = max(3, min(K, 10))
        for i in range(reps):
            _lib.check(lib.b200gs_render_host(ctypes.byref(gh), ctypes.byref(camc), c2w_pin[i % len(c2w_pin)].data_ptr(),
                                              img_host.data_ptr(), ctypes.byref(st)), "render_host")
        host_fps = reps / (time.perf_counter() - t0)
        del hp

    # ---- CPU baseline beside it + PARITY of the headline frame (rank 0, N=1 only): the oracle computes the whole
    #      frame of view 0 on the host cores (timed: cpu_baseline) and its stages are compared with the GPU frame -----
    cpu_base = parity = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import parity as PAR
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        ex = {k: v.numpy() for k, v in fr.export().items()}
        with torch.no_grad():
            sig_cpu = O.build_sigma_from_params(sc_cpu["scale_raw"], sc_cpu["q_raw"])
            t0 = time.perf_counter()
            col_cpu = O.evaluate_sh(sc_cpu["f_dc"], sc_cpu["f_rest"], sc_cpu["pos"], cams[0]["c2w"])
            img_cpu, proj, bins = O.render(sc_cpu["pos"], col_cpu, sc_cpu["opacity_raw"], sig_cpu, cams[0]["c2w"], H, W,
                                           intr["fx"], intr["fy"], intr["cx"], intr["cy"], return_stages=True)
            t_cpu = time.perf_counter() - t0
        cpu_base = {"value": 1.0 / t_cpu, "unit": "frames/s", "cores": cores, "torch_threads": torch.get_num_threads(),
                    "kind": "port", "sample": f"1 full headline frame (view 0, all {int(bins.uniq_tiles.shape[0])} non-empty "
                                              f"tiles): evaluate_sh + render of oracle/gs_oracle.py, {t_cpu:.2f} s; the "
                                              "oracle is bit-identical to the unmodified reference on this frame "
                                              "(profiles/PARITY_r02.json); `--impl reference` times the reference itself"}
        parity = PAR.compare_frame(ex, fr.n_isect, fr.n_visible, img_view0.cpu().numpy(), proj, bins, img_cpu.numpy())
        parity["yardstick"] = ("the unmodified reference's own fp32 run differs from its fp64 run in 520 of these 6 220 800 "
                               "values by more than 1e-4 (max 0.0191): threshold flips of q <= 6.25 / alpha >= 1/128 / T > 5e-5 "
                               "(oracle/, computed in the build container; profiles/PARITY_r02.json)")
        parity["what"] = ("GPU frame of view 0 (fused route, through the C ABI) against the oracle's full frame: survivors, "
                          "depth, radii, tile rects, per-tile sorted lists (exact) and the image (<= tol abs)")
        del ex, proj, bins, img_cpu
    del fr, img_view0

    # ---- the other two north-star modes on a few steps (reported inside `config`; `--mode` runs them at full length).
    #      Everything the headline line needs is measured by now: the two legs run under a deadline, so that a leg that
    #      hangs (one rank failing inside a peer-memory exchange leaves the others at a barrier) costs its own numbers, not
    #      the line - on expiry rank 0 prints the line with the leg marked, and every rank leaves with exit code 0 ----------
    extra_steps = max(4, min(K, 10))
    extras = {"tile_rows": None, "train": {"peer": None, "nccl": None, "peer_error": None, "peer_transport": None}}

    def build_line(tile_rows, train):
        hbm, peak_src, sm_max = peaks()
        counts = ncu_counts()
        bytes_model = algorithmic_bytes(N, V, I, P, tiles, S)
        table = {}
        for name, (ms, calls) in {**opt_regions, **train_regions, **fwd_regions}.items():
            b = bytes_model.get(name)
            row = {"ms": round(ms, 5), "calls": calls, "alg_bytes": b,
                   "gbs": None if b is None else round(b / (ms * 1e-3) / 1e9, 1),
                   "frac_hbm": None if b is None else round(b / (ms * 1e-3) / 1e9 / hbm, 4)}
            c = counts.get(name) or {}
            if isinstance(c, dict) and c.get("inst_executed"):
                # warp instructions issued per second against the issue peak: SMs x 4 schedulers x SM clock
                row["frac_issue"] = round(c["inst_executed"] / (148 * 4 * sm_max * 1e6 * ms * 1e-3), 4)
            table[name] = row
        step_kernels_ms = sum(v[0] for v in fwd_regions.values())
        dom = max(fwd_regions, key=lambda k: fwd_regions[k][0])
        dom_ms = fwd_regions[dom][0]
        dom_counts = counts.get(dom) if isinstance(counts.get(dom), dict) else {}
        issue_peak = 148 * 4 * sm_max * 1e6            # warp instructions per second the four schedulers of every SM can issue
        hbm_obj = {"achieved": bytes_model[dom] / (dom_ms * 1e-3) / 1e9, "peak": hbm, "unit": "GB/s",
                   "frac": bytes_model[dom] / (dom_ms * 1e-3) / 1e9 / hbm, "alg_bytes_per_launch": bytes_model[dom],
                   "peak_source": peak_src}
        if dom_counts.get("inst_executed"):
            ach = dom_counts["inst_executed"] / (dom_ms * 1e-3)
            roofline = {"kernel": dom, "bound": "fp32_issue", "achieved": ach / 1e9, "peak": issue_peak / 1e9,
                        "unit": "G warp-instructions/s", "frac": ach / issue_peak,
                        "traffic": dom_counts.get("dram_bytes"), "ms_per_launch": dom_ms,
                        "share_of_step_kernel_time": dom_ms / step_kernels_ms,
                        "inst_executed_per_launch": dom_counts["inst_executed"],
                        "inst_source": "ncu --set full capture of the same workload/view (profiles/ncu_traffic.json); time measured live",
                        "peak_source": f"148 SMs x 4 issue slots x {sm_max:.0f} MHz (max SM clock, MEASURED_PEAKS.json)",
                        "hbm": hbm_obj,
                        "note": "the blend is FP32-issue / MUFU bound (SURVEY.md 8d): the HBM object is kept for reference only"}
        else:
            roofline = {"kernel": dom, "bound": "hbm", **hbm_obj, "traffic": dom_counts.get("dram_bytes") if dom_counts else None,
                        "ms_per_launch": dom_ms, "share_of_step_kernel_time": dom_ms / step_kernels_ms}
        fps = world * K / (ms_render * 1e-3)
        best_train = min([x for x in (train["peer"], train["nccl"]) if x is not None], default=None)
        config = {"workload": wl["name"], "N": N, "H": H, "W": W, "views": wl["n_views"], "V_view0": V, "I_view0": I,
                  "super_pairs_view0": S,
                  "parallelism": f"frames/views sharded round-robin over {world} rank(s); per rank frames are software-"
                                 "pipelined by b200gs.RenderPipeline (project + binning of frame i+1 on a high-priority "
                                 "stream while frame i is blended on a second stream); Gaussians replicated",
                  "capacity_mode": os.environ.get("B200GS_CAPACITY_MODE"),
                  "l2_policy": "inputs larger than L2: every step streams 236 MB of parameters (L2 = 126 MB) and a different view",
                  # the other ways of counting frames, in keys a driver keeps
                  "single_stream_fps": world * K / (ms_render_1s * 1e-3),
                  "sync_per_frame_fps": world * K / s_sync_per_frame,
                  "sync_per_frame_note": "the reference's own timed region (render_trained.py:337-349): synchronize before and "
                                         "after every evaluate_sh + render, host clock",
                  "e2e_f32_frames_fps": world * K / s_e2e_f32,
                  "train_full_iteration": None if best_train is None else {
                      "views_per_s": world * extra_steps / (best_train * 1e-3), "ms_per_iteration": best_train / extra_steps,
                      "ms_peer_adam": None if train["peer"] is None else train["peer"] / extra_steps,
                      "ms_nccl_bucket": None if train["nccl"] is None else train["nccl"] / extra_steps,
                      "transport": train["peer_transport"], "peer_error": train["peer_error"], "steps": extra_steps,
                      "what": "scripts/train.py:463-538 on the fused path, views data-parallel (bench.py --mode train)"},
                  "tile_rows_4k": None if tile_rows is None else (tile_rows if "error" in tile_rows else {
                      "frames_per_s": tile_rows["frames_per_s"], "ms_per_frame": tile_rows["ms_per_frame"],
                      "bands": tile_rows["bands"], "steps": extra_steps,
                      "what": "6M Gaussians, 3840x2160, one band of tile rows per rank (bench.py --mode tile_rows)"}),
                  "parity": parity}
        line = base_line(env, METRIC, fps, "frames/s", ms_render, K, "weak", config)
        line.update({
            "single_stream": {"value": world * K / (ms_render_1s * 1e-3), "unit": "frames/s", "ms_per_step": ms_render_1s / K,
                              "note": "one frame at a time on one stream (frame latency)"},
            "train": {"with_l1_ssim_loss": {"value": K / (ms_train_loss * 1e-3), "unit": "it/s", "ms_per_step": ms_train_loss / K,
                                            "step": "build_sigma + evaluate_sh + render + b200gs.compute_loss (fused L1 + SSIM, "
                                                    "losses.py:158) + backward, one GPU's share"}},
            "e2e": {"value": world * K / s_e2e_u8, "unit": "frames/s", "h2d_bytes_per_step": 64, "d2h_bytes_per_step": H * W * 3,
                    "api": "b200gs.evaluate_sh + b200gs.RenderPipeline (pose from pinned host memory in; finished frame "
                           "delivered to pinned host memory as uint8 through b200gs.to_uint8 - the conversion the reference "
                           "scripts apply to every frame they keep, render_trained.py:357, bit-identical - 3 B/pixel over PCIe)"},
            "e2e_f32_frames": {"value": world * K / s_e2e_f32, "unit": "frames/s", "h2d_bytes_per_step": 64,
                               "d2h_bytes_per_step": H * W * 12,
                               "api": "the same loop delivering the raw fp32 image (12 B/pixel): PCIe-bound on one GPU and "
                                      "host-fabric-bound beyond two (profiles/r02_d2h_ceiling.json)"},
            "e2e_host_buffers": {"value": host_fps, "unit": "frames/s", "h2d_bytes_per_step": 236 * N + 64,
                                 "d2h_bytes_per_step": H * W * 12,
                                 "api": "b200gs_render_host (C ABI, every parameter array uploaded from host memory each call)"},
            "gpu_launches": int(launches_render),
            "gpu_launches_train": int(launches_train),
            "roofline": roofline,
            "kernels": table,
            "clocks": clocks,
            "cpu_baseline": cpu_base,
            "parity": parity,
        })
        return line

    def on_deadline():
        if rank != 0:
            return
        tr_leg, tn_leg = legs_at_deadline(extras)
        line = build_line(tr_leg, tn_leg)
        line["config"]["extras_deadline"] = {"tile_rows_finished": extras["tile_rows"] is not None,
                                             "train": tn_leg["peer_error"]}
        emit(line)

    deadline = Deadline(float(os.environ.get("B200GS_BENCH_EXTRAS_DEADLINE_S", "300")), on_deadline)
    if not args.no_extras:
        deadline.start()
        try:
            extras["tile_rows"] = measure_tile_rows(env, extra_steps, 3)
        except Exception as e:                    # noqa: BLE001 - reported, must not take the line with it
            extras["tile_rows"] = {"error": f"{type(e).__name__}: {e}"[:300]}
        extras["train"] = measure_train_iteration(env, sc, cams, extra_steps, 4, want_nccl=True)
    if not deadline.cancel():
        time.sleep(3600)                          # the deadline thread is printing the line and ends the process
    if rank == 0:
        emit(build_line(extras["tile_rows"], extras["train"]))
    env.finish()
    return 0


_REAL_STDOUT = None


def emit(line: dict):
    """Print the ONE JSON line on the real stdout (libraries such as NCCL print banners on fd 1, so fd 1 is
    pointed at stderr for the duration of the run)."""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is not None:
        os.write(_REAL_STDOUT, data)
    else:
        sys.stdout.write(data.decode())
        sys.stdout.flush()


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200gs", choices=["b200gs", "reference"])
    ap.add_argument("--mode", default="render", choices=["render", "train", "tile_rows"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="default mode: skip the short train / tile-row measurements")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200gs" else args.warmup
    if args.gpus > 1 and "WORLD_SIZE" not in os.environ:
        # launched directly: re-exec under torchrun, one process per GPU
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", os.environ.get("MASTER_PORT", "29531"),
               os.path.abspath(__file__)] + sys.argv[1:]
        os.dup2(_REAL_STDOUT, 1)
        return subprocess.call(cmd)
    if args.impl == "reference":
        return run_reference(args)
    return run_b200gs(args)


if __name__ == "__main__":
    sys.exit(main())
