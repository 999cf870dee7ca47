#!/bin/bash
# GPU call D (2 GPUs): multi-GPU parity tests, unchanged-script DP launcher check, bench modes at N=2, D2H probe.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_peer.py -q > gpurun_out/d_pytest_2gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/d_pytest_2gpu.log
tail -5 gpurun_out/d_pytest_2gpu.log
timeout 600 python tools/dp_launcher_check.py --iterations 30 --gpus 2 > gpurun_out/d_dp_check.log 2>&1; echo "dp check rc=$?"
timeout 300 $TR --master-port 29541 tools/peer_step_check.py 1000000 8 0 > gpurun_out/d_peer_check_2.json 2> gpurun_out/d_peer_check_2.err; echo "peer check rc=$?"
timeout 900 $TR --master-port 29542 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/d_bench_2gpu.json 2> gpurun_out/d_bench_2gpu.err; echo "bench rc=$?"
timeout 600 $TR --master-port 29543 bench.py --gpus 2 --steps 20 --warmup 5 --mode tile_rows > gpurun_out/d_bench_tile_rows_2gpu.json 2> gpurun_out/d_bench_tile_rows_2gpu.err; echo "tile rows rc=$?"
timeout 600 $TR --master-port 29544 bench.py --gpus 2 --steps 20 --warmup 5 --mode train > gpurun_out/d_bench_train_2gpu.json 2> gpurun_out/d_bench_train_2gpu.err; echo "train rc=$?"
timeout 200 $TR --master-port 29545 tools/d2h_probe.py > gpurun_out/d_d2h_2.json 2>/dev/null; echo "d2h rc=$?"
timeout 400 python tools/run_reference_scripts.py --iterations 150 > gpurun_out/d_ref_scripts_clean.log 2>&1; echo "ref scripts rc=$?"
echo done
