#!/bin/bash
# 2 GPUs: green multi-GPU test log + tile-row / train modes with the final band kernels
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_peer.py -q -rA > gpurun_out/i_pytest_2gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/i_pytest_2gpu.log
tail -16 gpurun_out/i_pytest_2gpu.log
timeout 600 $TR --master-port 29643 bench.py --gpus 2 --steps 20 --warmup 5 --mode tile_rows > gpurun_out/i_bench_tile_rows_2gpu.json 2> gpurun_out/i_tile.err; echo "tile rows rc=$?"
cat gpurun_out/i_bench_tile_rows_2gpu.json | cut -c1-400
echo done
