#!/bin/bash
# split kernels with more parts per supertile for bands: parity + the 8-rank emulation on one GPU
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_routed.py tests/test_gpu_config_parity.py tests/test_gpu_parity.py -x -q > gpurun_out/x_pytest.log 2>&1; tail -2 gpurun_out/x_pytest.log
for P in 4 8 0; do
  B200GS_SPLIT_PARTS=$P timeout 300 python tools/routed_probe.py 8 2>/dev/null > gpurun_out/x_probe_parts$P.json
  python - <<PY
import json
d = json.load(open("gpurun_out/x_probe_parts$P.json"))
print("parts=$P", {k: d[k] for k in ("dst_us", "dst_regions_band0", "dst_regions_band4", "equal_full_frame")})
PY
done
