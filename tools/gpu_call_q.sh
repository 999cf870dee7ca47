#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
free -g | head -2; nproc
timeout 1500 python tools/parity_configs.py c4 c5 --fp64 > gpurun_out/q_parity_configs.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/q_parity_configs.log | cut -c1-900
