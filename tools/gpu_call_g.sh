#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/g_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/g_pytest.log; tail -4 gpurun_out/g_pytest.log
timeout 300 python tools/band_times.py > gpurun_out/g_band_times.log 2>&1; echo "band rc=$?"; cat gpurun_out/g_band_times.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/g_bench.json 2> gpurun_out/g_bench.err; echo "bench rc=$?"
echo done
