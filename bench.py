#!/usr/bin/env python
"""Benchmark of the render path: render frames/s (+ train iterations/s) on the BASELINE.json headline
workload - 1M Gaussians, 1920x1080, SH degree 3, synthetic seeded scene.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200gs|reference]
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

A "step" is one pass of the hot path over one view:
  render step = evaluate_sh + render under no_grad   (the reference's own timed region,
                scripts/render_trained.py:337-349; build_sigma is outside, as at :192)
  train step  = build_sigma + evaluate_sh + render + weighted-sum loss + backward (+ NCCL all-reduce of the
                six gradient tensors when N > 1)      (scripts/train.py:463-530)
Multi-GPU: one process per GPU; render frames are sharded round-robin (no collective), training views are
data-parallel (weak scaling: per-GPU work fixed).  Rank 0 prints ONE JSON line.

--impl reference times the reference's algorithm on the host CPU cores (the oracle port, since the
reference checkout does not exist on the GPU box) on a bounded sample of the same workload.
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


# ---------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle port on the host cores, bounded sample
# ---------------------------------------------------------------------------------------------------
def cpu_frame_sample(sc, cam, budget_s: float):
    """One bounded sample of the headline frame on the CPU: projection + binning of ALL Gaussians, then the
    reference's per-tile blend loop on every k-th non-empty tile, k chosen to fit the budget.  Returns the
    estimated seconds per full frame (project + bin + blend_time * k) and a description."""
    from oracle import gs_oracle as O
    with torch.no_grad():
        t0 = time.perf_counter()
        color = O.evaluate_sh(sc["f_dc"], sc["f_rest"], sc["pos"], cam["c2w"])
        proj = O.project(sc["pos"], color, sc["opacity_raw"], sc["sigma"], cam["c2w"], cam["H"], cam["W"], cam["fx"],
                         cam["fy"], cam["cx"], cam["cy"])
        bins = O.bin_tiles(proj)
        t1 = time.perf_counter()
        n_tiles = int(bins.uniq_tiles.shape[0])
        # calibrate the per-tile cost on a handful of tiles, then pick the stride
        probe = max(1, n_tiles // 24)
        O.blend(proj, bins, tile_stride=probe)
        t2 = time.perf_counter()
        per_tile = (t2 - t1) / max(1, len(range(0, n_tiles, probe)))
        want = max(8, int(budget_s / max(per_tile, 1e-6)))
        stride = max(1, n_tiles // want)
        t3 = time.perf_counter()
        O.blend(proj, bins, tile_stride=stride)
        t4 = time.perf_counter()
        blended = len(range(0, n_tiles, stride))
    est = (t1 - t0) + (t4 - t3) * (n_tiles / blended)
    desc = (f"evaluate_sh+project+bin of all {sc['pos'].shape[0]} Gaussians ({t1 - t0:.2f} s) + reference per-tile blend "
            f"loop on {blended} of {n_tiles} non-empty tiles ({t4 - t3:.2f} s), extrapolated linearly in tiles")
    return est, desc, dict(V=proj.stage_counts["visible"], I=proj.stage_counts["intersections"])


def cpu_scene():
    from oracle import gs_oracle as O
    wl = WORKLOAD
    sc = O.make_scene(wl["n"], seed=wl["seed"], log_scale=wl["log_scale"], sh_degree=wl["sh_degree"])
    sc["sigma"] = O.build_sigma_from_params(sc["scale_raw"], sc["q_raw"])
    cams = [O.make_camera(wl["W"], wl["H"], view=v, n_views=wl["n_views"]) for v in range(wl["n_views"])]
    return sc, cams


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sc, cams = cpu_scene()
    total_steps = args.steps + args.warmup
    budget = min(20.0, max(1.0, 150.0 / max(1, total_steps) - 4.0))
    ests, desc, counts = [], "", {}
    t_begin = time.perf_counter()
    for i in range(total_steps):
        est, desc, counts = cpu_frame_sample(sc, cams[i % len(cams)], budget)
        if i >= args.warmup:
            ests.append(est)
        if time.perf_counter() - t_begin > 420 and len(ests) >= 1:      # hard stop: stay within minutes
            break
    sec = sum(ests) / len(ests)
    fps = 1.0 / sec
    line = {"impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
            "steps": len(ests), "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD["name"], "device": "host CPU", **counts},
            "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "torch_threads": torch.get_num_threads(),
                             "kind": "port", "sample": desc},
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


def run_b200gs(args):
    import torch.distributed as dist
    import b200gs
    from b200gs import _lib, ops
    from b200gs.dist import allreduce_gradients
    from oracle import gs_oracle as O   # scene generator + cpu_baseline leg only (never on the product path)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the b200gs arm has no CPU fallback (use --impl reference)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = b200gs.load_library()
    os.environ.setdefault("B200GS_CAPACITY_MODE", "speculative")
    wl = WORKLOAD
    H, W, K, Wm = wl["H"], wl["W"], args.steps, args.warmup

    sc_cpu = O.make_scene(wl["n"], seed=wl["seed"], log_scale=wl["log_scale"], sh_degree=wl["sh_degree"])
    sc = {k: v.to(dev) for k, v in sc_cpu.items()}
    cams = [O.make_camera(W, H, view=v, n_views=wl["n_views"]) for v in range(wl["n_views"])]
    c2w_dev = [c["c2w"].to(dev) for c in cams]
    c2w_pin = [c["c2w"].clone().pin_memory() for c in cams]
    intr = cams[0]
    view_of = lambda step: (rank + world * step) % len(cams)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def render_step(c2w):
        colors = b200gs.evaluate_sh(sc["f_dc"], sc["f_rest"], sc["pos"], c2w)
        return b200gs.render(sc["pos"], colors, sc["opacity_raw"], sigma, c2w, H, W, intr["fx"], intr["fy"],
                             intr["cx"], intr["cy"])

    def timed(fn, steps, warm):
        for i in range(warm):
            fn(i)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = lib.b200gs_kernel_launch_count()
        e0.record()
        for i in range(steps):
            fn(warm + i)
        e1.record()
        barrier()
        return max_over_ranks(e0.elapsed_time(e1)), lib.b200gs_kernel_launch_count() - l0

    # Frames are independent, so a render job keeps `n_streams` frames in flight, one CUDA stream each: the
    # latency-bound binning kernels of one frame overlap the issue-bound blend of the other (throughput mode;
    # the single-stream number - one frame at a time, the latency view - is reported next to it).
    n_streams = max(1, int(os.environ.get("B200GS_BENCH_STREAMS", "2")))
    streams = [torch.cuda.Stream(dev) for _ in range(n_streams)]

    def timed_streams(fn, steps, warm):
        main = torch.cuda.current_stream(dev)
        for i in range(warm):
            with torch.cuda.stream(streams[i % n_streams]):
                fn(i)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = lib.b200gs_kernel_launch_count()
        e0.record(main)
        for st in streams:
            st.wait_event(e0)
        for i in range(steps):
            with torch.cuda.stream(streams[i % n_streams]):
                fn(warm + i)
        for st in streams:
            main.wait_stream(st)
        e1.record(main)
        barrier()
        return max_over_ranks(e0.elapsed_time(e1)), lib.b200gs_kernel_launch_count() - l0

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()

    # ---- render: device-resident inputs ("value") --------------------------------------------------------
    with torch.no_grad():
        sigma = b200gs.build_sigma_from_params(sc["scale_raw"], sc["q_raw"])
        torch.cuda.synchronize()
        ms_render_1s, _ = timed(lambda i: render_step(c2w_dev[view_of(i)]), K, Wm)
        ms_render_2s, _ = timed_streams(lambda i: render_step(c2w_dev[view_of(i)]), K, Wm)
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
        # ---- render e2e: pose from pinned host memory in, image to pinned host memory out, every step; the
        #      D2H copy of a frame overlaps the next frame on the other stream (one pinned image per stream) ----
        img_pin = [torch.empty((H, W, 3), dtype=torch.float32).pin_memory() for _ in range(n_streams)]

        # the D2H copy of a finished frame runs on a copy stream, waiting for exactly that frame (done_event); a ring
        # of pinned images, each reused once its copy has completed
        copy_stream = torch.cuda.Stream(dev)
        n_ring = 3
        img_pin = [torch.empty((H, W, 3), dtype=torch.float32).pin_memory() for _ in range(n_ring)]
        copy_done = [torch.cuda.Event() for _ in range(n_ring)]
        for e in copy_done:
            e.record(copy_stream)
        delivered = [0]

        def deliver_f32(img, done):
            k = delivered[0] % n_ring
            delivered[0] += 1
            copy_done[k].synchronize()                   # the frame that used this pinned image has been delivered
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(done)
                img.record_stream(copy_stream)
                img_pin[k].copy_(img, non_blocking=True)
                copy_done[k].record(copy_stream)

        def e2e_step(i, deliver):
            with torch.cuda.stream(pipe.next_front_stream):
                c2w = c2w_pin[view_of(i)].to(dev, non_blocking=True)
            pipe_step(c2w, deliver)
        for i in range(Wm):
            e2e_step(i, deliver_f32)
        pipe_flush(deliver_f32)
        copy_stream.synchronize()
        barrier()
        t0 = time.perf_counter()
        for i in range(K):
            e2e_step(Wm + i, deliver_f32)
        pipe_flush(deliver_f32)
        copy_stream.synchronize()
        barrier()
        s_e2e = max_over_ranks(time.perf_counter() - t0)
        # the same loop delivering uint8 frames (b200gs.to_uint8: the conversion the reference scripts do on the
        # host, render_trained.py:357, done on the device): 3 B/pixel over PCIe instead of 12
        u8_pin = [torch.empty((H, W, 3), dtype=torch.uint8).pin_memory() for _ in range(n_ring)]

        def deliver_u8(img, done):
            k = delivered[0] % n_ring
            delivered[0] += 1
            copy_done[k].synchronize()
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(done)
                img.record_stream(copy_stream)
                u8_pin[k].copy_(b200gs.to_uint8(img), non_blocking=True)
                copy_done[k].record(copy_stream)
        for i in range(Wm):
            e2e_step(i, deliver_u8)
        pipe_flush(deliver_u8)
        copy_stream.synchronize()
        barrier()
        t0 = time.perf_counter()
        for i in range(K):
            e2e_step(Wm + i, deliver_u8)
        pipe_flush(deliver_u8)
        copy_stream.synchronize()
        barrier()
        s_e2e_u8 = max_over_ranks(time.perf_counter() - t0)

    # ---- train: fwd + bwd (+ all-reduce) --------------------------------------------------------------------
    leaves = {k: sc[k].clone().requires_grad_(True) for k in PARAMS}
    wimg = torch.rand(H, W, 3, generator=torch.Generator().manual_seed(1234))
    wimg_dev = (wimg / world).to(dev)
    wimg_pin = (wimg / world).pin_memory()

    def train_step(i, w=None):
        for p in leaves.values():
            p.grad = None
        c2w = c2w_dev[view_of(i)]
        sg = b200gs.build_sigma_from_params(leaves["scale_raw"], leaves["q_raw"])
        col = b200gs.evaluate_sh(leaves["f_dc"], leaves["f_rest"], leaves["pos"], c2w)
        img = b200gs.render(leaves["pos"], col, leaves["opacity_raw"], sg, c2w, H, W, intr["fx"], intr["fy"],
                            intr["cx"], intr["cy"])
        loss = (img * (wimg_dev if w is None else w)).sum()
        loss.backward()
        allreduce_gradients(leaves.values())
        return loss
    ms_train, launches_train = timed(lambda i: train_step(i), K, Wm)

    # the same step with the reference's own training loss (losses.py:158: 0.8 L1 + 0.2 (1 - SSIM) against a target
    # image) through the fused loss kernels, loss values left on the device (scripts/train.py:511 reads them back)
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
        (loss / world).backward()
        allreduce_gradients(leaves.values())
        return loss
    ms_train_loss, _ = timed(lambda i: train_step_loss(i), K, Wm)

    # e2e: every step's target image comes from pinned host memory (H2D inside the timed region) and the loss
    # is read back.  As a DataLoader with pin_memory would, the copy of step i+1's target runs on a copy stream
    # while step i computes (two device buffers).
    copy_stream = torch.cuda.Stream(dev)
    w_dev = [torch.empty_like(wimg_dev) for _ in range(2)]
    w_ready = [torch.cuda.Event() for _ in range(2)]
    w_free = [torch.cuda.Event() for _ in range(2)]

    def prefetch_target(i):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(w_free[i % 2])             # the step that used this buffer is done with it
            w_dev[i % 2].copy_(wimg_pin, non_blocking=True)
            w_ready[i % 2].record(copy_stream)

    def train_e2e_step(i):
        prefetch_target(i + 1)
        torch.cuda.current_stream(dev).wait_event(w_ready[i % 2])
        loss = train_step(i, w_dev[i % 2])
        w_free[i % 2].record(torch.cuda.current_stream(dev))
        return float(loss.item())                        # D2H read of the step's result
    for e in w_free:
        e.record(torch.cuda.current_stream(dev))
    prefetch_target(0)
    for i in range(min(Wm, 3)):
        train_e2e_step(i)
    barrier()
    t0 = time.perf_counter()
    for i in range(K):
        train_e2e_step(min(Wm, 3) + i)
    barrier()
    s_train_e2e = max_over_ranks(time.perf_counter() - t0)
    clocks = sampler.stop() if rank == 0 else None

    # ---- per-kernel CUDA-event times (separate pass with an event pair around every kernel group) ------------
    lib.b200gs_profile_enable(1)
    with torch.no_grad():
        for i in range(K):
            render_step(c2w_dev[view_of(i)])
    torch.cuda.synchronize()
    nreg = 24
    ms_buf, call_buf = (ctypes.c_float * nreg)(), (ctypes.c_int32 * nreg)()
    n_regions = lib.b200gs_profile_collect(ms_buf, call_buf, nreg)
    fwd_regions = {lib.b200gs_profile_region_name(r).decode(): (ms_buf[r] / max(1, call_buf[r]), call_buf[r])
                   for r in range(n_regions) if call_buf[r]}
    for i in range(K):
        train_step_loss(i)
    torch.cuda.synchronize()
    n_regions = lib.b200gs_profile_collect(ms_buf, call_buf, nreg)
    train_regions = {lib.b200gs_profile_region_name(r).decode(): (ms_buf[r] / max(1, call_buf[r]), call_buf[r])
                     for r in range(n_regions) if call_buf[r]}
    # ---- the whole training iteration of scripts/train.py:463-538 on the fused path: build_sigma + evaluate_sh + render
    #      + L1/SSIM loss + backward (+ all-reduce) + clip_grad_norm_(pos, 1.0) + Adam over the six groups (the
    #      reference's learning rates, eps 1e-15).  Runs last: it moves the parameters.
    lr0 = {"pos": 1.6e-4 * 0.01, "opacity_raw": 0.05, "f_dc": 2.5e-3, "f_rest": 2.5e-3 / 20.0, "scale_raw": 5e-3, "q_raw": 1e-3}
    opt = b200gs.FusedAdam([{"params": [leaves[k]], "lr": lr0[k], "name": k} for k in PARAMS], lr=1e-3, eps=1e-15)

    def train_full(i):
        opt.zero_grad(set_to_none=True)
        train_step_loss_nograd_reset(i)
        b200gs.clip_grad_norm_(leaves["pos"], max_norm=1.0)
        opt.step()

    def train_step_loss_nograd_reset(i, reduce=True):
        c2w = c2w_dev[view_of(i)]
        sg = b200gs.build_sigma_from_params(leaves["scale_raw"], leaves["q_raw"])
        col = b200gs.evaluate_sh(leaves["f_dc"], leaves["f_rest"], leaves["pos"], c2w)
        img = b200gs.render(leaves["pos"], col, leaves["opacity_raw"], sg, c2w, H, W, intr["fx"], intr["fy"],
                            intr["cx"], intr["cy"])
        loss, _ = b200gs.compute_loss_tensors(img, target_dev)
        (loss / world).backward()
        if reduce:
            allreduce_gradients(leaves.values())
    for i in range(4):                      # first steps allocate the optimizer state: not representative
        train_full(i)
    torch.cuda.synchronize()
    lib.b200gs_profile_collect(ms_buf, call_buf, nreg)
    for i in range(6):
        train_full(4 + i)
    torch.cuda.synchronize()
    n_regions = lib.b200gs_profile_collect(ms_buf, call_buf, nreg)
    opt_regions = {lib.b200gs_profile_region_name(r).decode(): (ms_buf[r] / max(1, call_buf[r]), call_buf[r])
                   for r in range(n_regions) if call_buf[r] and lib.b200gs_profile_region_name(r).decode() in
                   ("adam_step", "clip_grad_norm")}
    lib.b200gs_profile_enable(0)
    ms_train_full, _ = timed(lambda i: train_full(i), K, Wm)
    # ---- frame statistics of view 0 (V, I) for the byte model ---------------------------------------------------
    with torch.no_grad():
        g, keep = ops._gaussians(sc["pos"], sc["opacity_raw"], sc["scale_raw"], sc["q_raw"], None, sc["f_dc"],
                                 sc["f_rest"], None)
        cfg = ops.RenderConfig(H=H, W=W, fx=intr["fx"], fy=intr["fy"], cx=intr["cx"], cy=intr["cy"])
        fr = ops.Frame(g, keep, cfg, c2w_dev[0], dev)
        fr.render("sync")
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

    # ---- CPU baseline beside it (rank 0, N=1 only) ---------------------------------------------------------------
    cpu_base = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        sc_cpu["sigma"] = O.build_sigma_from_params(sc_cpu["scale_raw"], sc_cpu["q_raw"])
        est, desc, _ = cpu_frame_sample(sc_cpu, cams[0], budget_s=12.0)
        cpu_base = {"value": 1.0 / est, "unit": "frames/s", "cores": cores, "torch_threads": torch.get_num_threads(),
                    "kind": "port", "sample": desc}

    # ---- N > 1: the same iteration with the optimizer half done over NVLink peer memory - gradient reduce-scatter +
    # clip + Adam on the owned shard + parameter all-gather in one kernel (b200gs.PeerAdam, csrc/peer.cu) instead of NCCL
    # all-reduce + clip + Adam.  Runs after everything else (it re-homes the parameters into the peer-visible buffer)
    # and may fail without taking the other numbers with it.
    ms_train_peer = peer_transport = peer_error = None
    if world > 1:
        try:
            opt_p = b200gs.PeerAdam([{"params": [leaves[k]], "lr": lr0[k], "name": k} for k in PARAMS], lr=1e-3, eps=1e-15,
                                    clip_params=[leaves["pos"]], max_norm=1.0)
            peer_transport = opt_p.area.transport + ("+nvls" if opt_p.area.c_group.multicast else "")

            def train_full_peer(i):
                opt_p.zero_grad(set_to_none=True)
                train_step_loss_nograd_reset(i, reduce=False)
                opt_p.step()
            ms_train_peer, _ = timed(lambda i: train_full_peer(i), K, Wm)
        except Exception as e:                    # noqa: BLE001 - reported in the JSON line
            ms_train_peer = None
            peer_error = f"{type(e).__name__}: {e}"[:300]

    if rank != 0:
        if world > 1:
            try:
                dist.destroy_process_group()
            except Exception:                     # noqa: BLE001 - nothing left to report from this rank
                pass
        return 0

    hbm, peak_src, sm_max = peaks()
    bytes_model = algorithmic_bytes(N, V, I, P, tiles, S)
    table = {}
    for name, (ms, calls) in {**opt_regions, **train_regions, **fwd_regions}.items():
        b = bytes_model.get(name)
        table[name] = {"ms": round(ms, 5), "calls": calls, "alg_bytes": b,
                       "gbs": None if b is None else round(b / (ms * 1e-3) / 1e9, 1),
                       "frac_hbm": None if b is None else round(b / (ms * 1e-3) / 1e9 / hbm, 4)}
    step_kernels_ms = sum(v[0] for v in fwd_regions.values())
    dom = max(fwd_regions, key=lambda k: fwd_regions[k][0])
    dom_ms = fwd_regions[dom][0]
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(tpath):
        traffic = json.load(open(tpath)).get(dom)
    fps = world * K / (ms_render * 1e-3)
    fp32_peak = 148 * 128 * sm_max * 1e6          # FP32 lane-instructions per second at the max SM clock
    pairs = float(I) * 256.0                       # (pixel, splat) evaluations if no tile exits early
    line = {
        "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": world, "steps": K, "warmup": Wm,
        "ms_per_step": ms_render / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl["name"], "N": N, "H": H, "W": W, "views": wl["n_views"], "V_view0": V, "I_view0": I, "super_pairs_view0": S,
                   "parallelism": f"frames/views sharded round-robin over {world} rank(s); per rank frames are software-"
                                  "pipelined by b200gs.RenderPipeline (project + binning of frame i+1 on a high-priority "
                                  "stream while frame i is blended on a second stream); Gaussians replicated",
                   "capacity_mode": os.environ.get("B200GS_CAPACITY_MODE"),
                   "l2_policy": "inputs larger than L2: every step streams 236 MB of parameters (L2 = 126 MB) and a different view"},
        "single_stream": {"value": world * K / (ms_render_1s * 1e-3), "unit": "frames/s", "ms_per_step": ms_render_1s / K,
                          "note": "one frame at a time on one stream (frame latency)"},
        "two_streams": {"value": world * K / (ms_render_2s * 1e-3), "unit": "frames/s",
                        "note": "whole frames alternating on two equal-priority streams"},
        "train": {"value": K / (ms_train * 1e-3), "unit": "it/s", "views_per_s": world * K / (ms_train * 1e-3),
                  "ms_per_step": ms_train / K,
                  "step": "build_sigma + evaluate_sh + render + weighted-sum loss + backward" +
                          (" + NCCL sum all-reduce of 6 gradient tensors (236 MB)" if world > 1 else ""),
                  "with_l1_ssim_loss": {"value": K / (ms_train_loss * 1e-3), "unit": "it/s", "ms_per_step": ms_train_loss / K,
                                        "step": "the same with b200gs.compute_loss (fused L1 + SSIM, losses.py:158) "
                                                "instead of the weighted sum"},
                  "full_iteration": {"value": K / (ms_train_full * 1e-3), "unit": "it/s", "ms_per_step": ms_train_full / K,
                                     "views_per_s": world * K / (ms_train_full * 1e-3),
                                     "step": "train.py:463-538 on the fused path: render + L1/SSIM loss + backward + " +
                                             ("NCCL sum all-reduce of 6 gradient tensors + " if world > 1 else "") +
                                             "clip_grad_norm_(pos) + fused Adam over the six parameter groups"},
                  **({"full_iteration_peer": {
                      "value": K / (ms_train_peer * 1e-3), "unit": "it/s", "ms_per_step": ms_train_peer / K,
                      "views_per_s": world * K / (ms_train_peer * 1e-3), "transport": peer_transport,
                      "step": "the same iteration with b200gs.PeerAdam: gradient reduce-scatter + clip + Adam on the owned "
                              "shard + parameter all-gather in ONE kernel over NVLink peer memory (no NCCL on the data path)"}}
                     if ms_train_peer else ({"full_iteration_peer": {"error": peer_error}} if peer_error else {})),
                  "e2e": {"value": K / s_train_e2e, "unit": "it/s", "h2d_bytes_per_step": H * W * 12,
                          "d2h_bytes_per_step": 4}},
        "e2e": {"value": world * K / s_e2e, "unit": "frames/s", "h2d_bytes_per_step": 64, "d2h_bytes_per_step": H * W * 12,
                "api": "b200gs.evaluate_sh + b200gs.RenderPipeline.render (pose from pinned host memory, image to "
                       "pinned host memory, two pinned images in rotation)"},
        "e2e_u8_frames": {"value": world * K / s_e2e_u8, "unit": "frames/s", "h2d_bytes_per_step": 64,
                          "d2h_bytes_per_step": H * W * 3,
                          "api": "the same, frames delivered as uint8 through b200gs.to_uint8 (device-side frame sink)"},
        "e2e_host_buffers": {"value": host_fps, "unit": "frames/s", "h2d_bytes_per_step": 236 * N + 64,
                             "d2h_bytes_per_step": H * W * 12,
                             "api": "b200gs_render_host (C ABI, every parameter array uploaded from host memory each call)"},
        "gpu_launches": int(launches_render),
        "gpu_launches_train": int(launches_train),
        "roofline": {"kernel": dom, "bound": "hbm", "achieved": bytes_model[dom] / (dom_ms * 1e-3) / 1e9, "peak": hbm,
                     "unit": "GB/s", "frac": bytes_model[dom] / (dom_ms * 1e-3) / 1e9 / hbm, "traffic": traffic,
                     "peak_source": peak_src, "alg_bytes_per_launch": bytes_model[dom], "ms_per_launch": dom_ms,
                     "share_of_step_kernel_time": dom_ms / step_kernels_ms,
                     "note": "blend is FP32-issue/MUFU bound, not HBM bound (SURVEY.md 8d): see roofline_fp32"},
        "roofline_fp32": {"kernel": "blend_fwd", "pair_evals_upper": pairs,
                          "gpairs_per_s": pairs / (fwd_regions.get("blend_fwd", (dom_ms, 0))[0] * 1e-3) / 1e9,
                          "fp32_lane_instr_peak_per_s": fp32_peak,
                          "lane_instr_per_pair_at_peak": fp32_peak / (pairs / (fwd_regions.get("blend_fwd", (dom_ms, 0))[0] * 1e-3))},
        "kernels": table,
        "clocks": clocks,
        "cpu_baseline": cpu_base,
    }
    emit(line)
    if world > 1:
        try:
            dist.destroy_process_group()
        except Exception:                         # noqa: BLE001 - the line is out
            pass
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
    ap.add_argument("--no-cpu-baseline", action="store_true")
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
