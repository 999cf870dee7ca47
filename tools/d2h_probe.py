"""Device-to-host ceiling of the box, per rank count: every rank copies frame-sized buffers (fp32 1080p = 24.9 MB,
uint8 = 6.2 MB) from its GPU into its own pinned host ring as fast as the copy engine goes - no rendering at all.
This is the upper bound of any end-to-end frame rate that delivers frames to host memory; bench.py's `e2e` numbers are
to be read against it.

    python tools/d2h_probe.py                      (1 rank)
    torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 tools/d2h_probe.py

Rank 0 prints one JSON line: aggregate GB/s and frames/s for both frame sizes at this rank count.
"""
import json
import os
import time

import torch
import torch.distributed as dist


def main():
    world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    out = {"n_gpus": world}
    for name, nbytes in (("f32_1080p", 1080 * 1920 * 12), ("u8_1080p", 1080 * 1920 * 3)):
        src = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        ring = [torch.empty(nbytes, dtype=torch.uint8).pin_memory() for _ in range(3)]
        stream = torch.cuda.Stream(dev)
        reps = 200

        def loop(n):
            with torch.cuda.stream(stream):
                for i in range(n):
                    ring[i % 3].copy_(src, non_blocking=True)
            stream.synchronize()
        loop(20)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        loop(reps)
        if world > 1:
            dist.barrier()
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        out[name] = {"aggregate_GBps": world * reps * nbytes / dt / 1e9, "aggregate_frames_per_s": world * reps / dt,
                     "per_gpu_GBps": reps * nbytes / dt / 1e9}
    if rank == 0:
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
