"""Micro-benchmark of the scan / radix-sort primitives through the C ABI (CUDA events, L2 flushed)."""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "3d-gaussian-splatting-for-novel-view-synthesis_b200"))
import numpy as np  # noqa: E402
import torch  # noqa: E402
from b200gs import _lib  # noqa: E402

lib = _lib.load()
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def bench_sort(n, bits, reps=20):
    rng = np.random.default_rng(0)
    keys = torch.from_numpy(rng.integers(0, 2 ** bits, size=n, dtype=np.int64).astype(np.uint32).view(np.int32)).cuda()
    vals = torch.arange(n, dtype=torch.int32, device="cuda")
    ko, vo = torch.empty_like(keys), torch.empty_like(vals)
    nb = lib.b200gs_sort_scratch_bytes(n)
    scratch = torch.empty(nb, dtype=torch.uint8, device="cuda")
    ts = []
    for r in range(reps + 3):
        k, v = keys.clone(), vals.clone()
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.check(lib.b200gs_radix_sort_pairs(k.data_ptr(), v.data_ptr(), ko.data_ptr(), vo.data_ptr(), n, 0, bits,
                                               scratch.data_ptr(), nb, st))
        e1.record()
        torch.cuda.synchronize()
        if r >= 3:
            ts.append(e0.elapsed_time(e1) * 1e3)
    ref = torch.sort(keys.view(torch.int32).to(torch.int64) & 0xFFFFFFFF, stable=True)
    if not os.environ.get("B200GS_SORT_DBG"):
        assert torch.equal(vo.to(torch.int64), ref.indices), "sort wrong"
    ts.sort()
    passes = (bits + 7) // 8
    print(f"sort n={n} bits={bits} passes={passes}: median {ts[len(ts)//2]:.1f} us  min {ts[0]:.1f} us  "
          f"({n * 16 * passes / ts[len(ts)//2] / 1e3:.0f} GB/s realistic traffic)")


def bench_scan(n, reps=20):
    x = torch.randint(0, 30, (n,), dtype=torch.int32, device="cuda")
    out = torch.empty_like(x)
    tot = torch.zeros(1, dtype=torch.int32, device="cuda")
    nb = lib.b200gs_scan_scratch_bytes(n)
    scratch = torch.empty(nb, dtype=torch.uint8, device="cuda")
    ts = []
    for r in range(reps + 3):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.check(lib.b200gs_exclusive_scan_u32(x.data_ptr(), out.data_ptr(), n, tot.data_ptr(), scratch.data_ptr(), nb, st))
        e1.record()
        torch.cuda.synchronize()
        if r >= 3:
            ts.append(e0.elapsed_time(e1) * 1e3)
    assert int(tot) == int(x.sum())
    ts.sort()
    print(f"scan n={n}: median {ts[len(ts)//2]:.1f} us")


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "one":
    bench_sort(int(sys.argv[2]), int(sys.argv[3]), reps=1)
    sys.exit(0)
if __name__ == "__main__":
    for n, bits in [(1_000_000, 32), (4_400_000, 13), (13_000_000, 13), (32_500_000, 15), (3_000_000, 32)]:
        bench_sort(n, bits)
    for n in (1_000_000, 6_000_000):
        bench_scan(n)
