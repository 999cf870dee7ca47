#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_routed.py -x -q > gpurun_out/v_pytest.log 2>&1; tail -3 gpurun_out/v_pytest.log
timeout 300 python tools/kernel_times.py 20 0 > gpurun_out/v_kernel_times.log 2>&1; cat gpurun_out/v_kernel_times.log | tail -3
timeout 300 python tools/pipe_fps.py 2>&1 | tail -1
timeout 300 python tools/routed_probe.py 8 > gpurun_out/v_routed_probe8d.json 2> gpurun_out/v_routed_probe8.err; tail -3 gpurun_out/v_routed_probe8.err; cat gpurun_out/v_routed_probe8d.json
