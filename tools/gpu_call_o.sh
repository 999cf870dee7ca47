#!/bin/bash
# the driver's round-end sequence on one fresh box: smoke, GPU tests, reference arm, bench
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/o_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/o_smoke.log
timeout 1800 python -m pytest tests/ -x -q -m gpu > gpurun_out/o_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/o_pytest.log
( time timeout 1700 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/o_bench_reference.json 2> gpurun_out/o_bench_reference.err ) 2> gpurun_out/o_ref_time.txt; echo "ref rc=$?"; cat gpurun_out/o_ref_time.txt | tail -3
( time timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/o_bench.json 2> gpurun_out/o_bench.err ) 2> gpurun_out/o_bench_time.txt; echo "bench rc=$?"; cat gpurun_out/o_bench_time.txt | tail -3
python - <<'PY'
import json
r=json.load(open("gpurun_out/o_bench_reference.json")); d=json.load(open("gpurun_out/o_bench.json"))
print("reference", r["value"], r["steps"], r["ms_per_step"], r["cpu_baseline"]["kind"], r["cpu_baseline"]["cores"])
print("b200gs value", d["value"], "e2e", d["e2e"]["value"], "ratio", d["value"]/r["value"], "e2e ratio", d["e2e"]["value"]/r["value"])
PY
echo done
