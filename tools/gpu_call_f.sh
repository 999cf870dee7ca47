#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_config_parity.py tests/test_gpu_parity.py -q -k "band or full_size or pipeline or golden" > gpurun_out/f_pytest.log 2>&1; tail -4 gpurun_out/f_pytest.log
timeout 300 python tools/band_times.py > gpurun_out/f_band_times.log 2>&1; echo "band rc=$?"; cat gpurun_out/f_band_times.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"blend|preprocess|onesweep|split_super|scan_emit|l1_ssim|adam_step|grad_s" -c 48 -o gpurun_out/f_full python tools/profile_target.py 2 1 > gpurun_out/f_ncu_full.log 2>&1; echo "ncu full rc=$?"
ncu -i gpurun_out/f_full.ncu-rep --page raw --csv > gpurun_out/f_full_raw.csv 2>/dev/null; echo "export rc=$?"
echo done
