#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
for v in 2 1; do
B200GS_PRE_CTAS=$v timeout 600 python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/l_bench_overlap_pre$v.json 2> gpurun_out/l_bench.err; echo "bench rc=$?"
done
python - <<'PY'
import json
for f in ("l_bench_overlap_pre2.json","l_bench_overlap_pre1.json"):
    d=json.load(open("gpurun_out/"+f)); c=d["config"]
    print(f, "value %.0f single %.0f sync %.0f e2e %.0f pre %.4f"%(d["value"], c["single_stream_fps"], c["sync_per_frame_fps"], d["e2e"]["value"], d["kernels"]["preprocess_fwd"]["ms"]))
PY
echo done
