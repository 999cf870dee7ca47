#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -k "pipeline" > gpurun_out/p_pytest.log 2>&1; tail -3 gpurun_out/p_pytest.log
for b in 1 2 3 4 6; do
B200GS_PIPE_BATCH=$b timeout 600 python bench.py --steps 24 --warmup 6 --no-extras --no-cpu-baseline > gpurun_out/p_bench_batch$b.json 2> gpurun_out/p_bench.err; echo "bench batch=$b rc=$?"
done
python - <<'PY'
import json
for b in (1,2,3,4,6):
    try:
        d=json.load(open("gpurun_out/p_bench_batch%d.json"%b)); c=d["config"]
        print("batch",b,"value %.0f e2e %.0f f32 %.0f"%(d["value"], d["e2e"]["value"], c["e2e_f32_frames_fps"]))
    except Exception as e: print("batch",b,"ERR",e)
PY
