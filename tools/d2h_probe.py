"""Device-to-host copy rate of one 1080p frame (fp32 and uint8) into pinned memory."""
import torch, time
H, W = 1080, 1920
img = torch.rand(H, W, 3, device="cuda")
for parts in (1, 2, 4):
    pins = [torch.empty((H // parts) * W * 3, dtype=torch.float32).pin_memory() for _ in range(parts)]
    streams = [torch.cuda.Stream() for _ in range(parts)]
    flat = img.view(-1)
    chunk = flat.numel() // parts
    torch.cuda.synchronize()
    for rep in range(2):
        t0 = time.perf_counter()
        for it in range(50):
            for p in range(parts):
                with torch.cuda.stream(streams[p]):
                    pins[p].copy_(flat[p * chunk:(p + 1) * chunk], non_blocking=True)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / 50
    print(parts, "streams:", round(dt * 1e6), "us per frame,", round(flat.numel() * 4 / dt / 1e9, 1), "GB/s")
