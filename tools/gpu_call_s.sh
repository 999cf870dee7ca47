#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"blend_fwd|blend_bwd|onesweep|split_super|scan_emit" -s 12 -c 16 -o gpurun_out/s_full python tools/profile_target.py 3 1 > gpurun_out/s_ncu_full.log 2>&1; echo "ncu full rc=$?"
ncu -i gpurun_out/s_full.ncu-rep --page raw --csv > gpurun_out/s_full_raw.csv 2>/dev/null; echo "export rc=$?"
ncu -i gpurun_out/s_full.ncu-rep --page source --csv > gpurun_out/s_full_src.csv 2>/dev/null; echo "export src rc=$?"
ls -la gpurun_out/s_full*
echo done
