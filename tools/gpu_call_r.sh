#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
T4="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
timeout 300 $T4 --master-port 29751 tools/peer_two_areas_check.py 500000 6 2> gpurun_out/r_two_areas.err | grep "^{" > gpurun_out/r_two_areas_4gpu.json; echo "rc=$?"; cat gpurun_out/r_two_areas_4gpu.json; tail -3 gpurun_out/r_two_areas.err | cut -c1-300
timeout 300 $T4 --master-port 29752 tools/peer_two_areas_check.py 1000000 8 2>> gpurun_out/r_two_areas.err | grep "^{" > gpurun_out/r_two_areas_4gpu_1m.json; echo "rc=$?"; cat gpurun_out/r_two_areas_4gpu_1m.json
