"""Does the adam_step time depend on WHERE its four arrays sit relative to each other?  One tensor of 45 M floats
(f_rest at N = 1M: 180 MB per array) carved out of one big buffer at controlled relative offsets."""
import ctypes, sys
sys.path.insert(0, "3d-gaussian-splatting-for-novel-view-synthesis_b200")
import torch
from b200gs import _lib
lib = _lib.load()
n = 45_000_000
nbytes = n * 4
big = torch.empty(3 << 30, dtype=torch.uint8, device="cuda")
big.zero_()
base = big.data_ptr()
base = (base + (2 << 20) - 1) // (2 << 20) * (2 << 20)          # 2 MB aligned
print("base % 1GB =", hex(base % (1 << 30)))
stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)

def run(offsets, reps=12):
    t = (_lib.AdamTensor * 1)()
    p, g, m, v = [base + o for o in offsets]
    t[0] = _lib.AdamTensor(p, g, m, v, n, 1e-3, 5, 0)
    for _ in range(3):
        lib.b200gs_adam_step(t, 1, 0.9, 0.999, 1e-15, stream)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        lib.b200gs_adam_step(t, 1, 0.9, 0.999, 1e-15, stream)
    b.record(); torch.cuda.synchronize()
    us = a.elapsed_time(b) / reps * 1e3
    return round(us, 1), round(28 * n / us / 1e3)     # GB/s

MB = 1 << 20
span = (nbytes + 2 * MB - 1) // (2 * MB) * (2 * MB)              # 180 MB rounded up to 2 MB
for name, stagger in (("2MB-aligned, no stagger", 0), ("+256 B", 256), ("+4 KB", 4096), ("+32 KB", 32768), ("+64 KB", 65536),
                      ("+256 KB", 262144), ("+512 KB", 524288), ("+1 MB", MB), ("unaligned (+ 7 777 280 B)", 7777280)):
    offs = [k * (span + stagger) for k in range(4)]
    print(f"{name:28s}", run(offs))
for gap_mb in (182, 184, 192, 256, 512):
    offs = [k * gap_mb * MB for k in range(4)]
    print(f"gap {gap_mb} MB".ljust(28), run(offs))
