#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_config_parity.py tests/test_gpu_parity.py -q -k "band or full_size" > gpurun_out/h_pytest.log 2>&1; tail -3 gpurun_out/h_pytest.log
timeout 300 python tools/band_times.py > gpurun_out/h_band_times.log 2>&1; echo "band rc=$?"; cat gpurun_out/h_band_times.log
B200GS_BAND_SELECT_MAX_PCT=100 timeout 300 python tools/band_times.py > gpurun_out/h_band_times_select100.log 2>&1; echo "band rc=$?"; tail -1 gpurun_out/h_band_times_select100.log
echo done
