#!/bin/bash
# 8 GPUs: the default bench (render + short train / tile-row legs), the scaling points at 4 GPUs, NVLS train, D2H ceilings,
# PeerAdam against a local torch replay with and without multicast.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
T8="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
T4="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
nvidia-smi topo -m > gpurun_out/j_topo.txt 2>&1
timeout 900 $T8 --master-port 29701 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/j_bench_8gpu.json 2> gpurun_out/j_bench_8gpu.err; echo "bench8 rc=$?"
timeout 400 $T8 --master-port 29702 bench.py --gpus 8 --steps 30 --warmup 5 --mode tile_rows > gpurun_out/j_tile_rows_8gpu.json 2> gpurun_out/j_tile8.err; echo "tile8 rc=$?"
CUDA_VISIBLE_DEVICES=0,1,2,3 timeout 400 $T4 --master-port 29703 bench.py --gpus 4 --steps 30 --warmup 5 --mode tile_rows > gpurun_out/j_tile_rows_4gpu.json 2> gpurun_out/j_tile4.err; echo "tile4 rc=$?"
CUDA_VISIBLE_DEVICES=0,1,2,3 timeout 400 $T4 --master-port 29704 bench.py --gpus 4 --steps 20 --warmup 5 --mode train > gpurun_out/j_train_4gpu.json 2> gpurun_out/j_train4.err; echo "train4 rc=$?"
timeout 400 $T8 --master-port 29705 bench.py --gpus 8 --steps 20 --warmup 5 --mode train > gpurun_out/j_train_8gpu.json 2> gpurun_out/j_train8.err; echo "train8 rc=$?"
B200GS_PEER_MULTICAST=1 timeout 400 $T8 --master-port 29706 bench.py --gpus 8 --steps 20 --warmup 5 --mode train > gpurun_out/j_train_8gpu_nvls.json 2> gpurun_out/j_train8n.err; echo "train8 nvls rc=$?"
CUDA_VISIBLE_DEVICES=0,1,2,3 timeout 120 $T4 --master-port 29707 tools/d2h_probe.py 2>/dev/null | grep "^{" > gpurun_out/j_d2h_4.json; echo "d2h4 rc=$?"
timeout 120 $T8 --master-port 29708 tools/d2h_probe.py 2>/dev/null | grep "^{" > gpurun_out/j_d2h_8.json; echo "d2h8 rc=$?"
timeout 200 $T8 --master-port 29709 tools/peer_step_check.py 1000000 6 0 2>/dev/null | grep "^{" > gpurun_out/j_peer_check_8_plain.json; echo "check plain rc=$?"
timeout 200 $T8 --master-port 29710 tools/peer_step_check.py 1000000 6 1 2>/dev/null | grep "^{" > gpurun_out/j_peer_check_8_nvls.json; echo "check nvls rc=$?"
echo done
