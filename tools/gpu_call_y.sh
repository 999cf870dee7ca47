#!/bin/bash
# final ncu launch list of round 2 (3 frames + 2 whole training iterations), after the same command ran clean without ncu
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 200 python tools/profile_target.py 3 2 > gpurun_out/y_plain.log 2>&1; echo "plain rc=$?"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/y_launches.csv python tools/profile_target.py 3 2 > gpurun_out/y_ncu.log 2>&1; echo "ncu rc=$?"
python tools/summarize_ncu.py launches gpurun_out/y_launches.csv gpurun_out/y_launches.md; head -12 gpurun_out/y_launches.md
