#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q > gpurun_out/u_pytest.log 2>&1; tail -3 gpurun_out/u_pytest.log
timeout 300 python tools/kernel_times.py 20 0 > gpurun_out/u_kernel_times.log 2>&1; cat gpurun_out/u_kernel_times.log | tail -3
for cfg in "" "B200GS_SORT_BPS=3" "B200GS_SORT_SHAPE=wide" ; do
echo "== pipe $cfg"; env $cfg timeout 300 python tools/pipe_fps.py 2>&1 | tail -1
done
