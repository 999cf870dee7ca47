#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
B200GS_FWD_HALF=1 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_config_parity.py -q -x > gpurun_out/k_pytest_half.log 2>&1; tail -3 gpurun_out/k_pytest_half.log
B200GS_FWD_HALF=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/k_bench_half.json 2> gpurun_out/k_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/k_bench_full.json 2>> gpurun_out/k_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
for f in ("k_bench_half.json","k_bench_full.json"):
    d=json.load(open("gpurun_out/"+f)); c=d["config"]
    print(f, "value %.0f single %.0f sync %.0f e2e %.0f blend_fwd %.4f"%(d["value"], c["single_stream_fps"], c["sync_per_frame_fps"], d["e2e"]["value"], d["kernels"]["blend_fwd"]["ms"]))
PY
echo done
