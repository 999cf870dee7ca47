#!/bin/bash
# final single-GPU records of round 2: whole -m gpu suite, tile-row frame on one GPU, default bench
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/w_pytest_gpu.log 2>&1; tail -3 gpurun_out/w_pytest_gpu.log
timeout 600 python bench.py --mode tile_rows --steps 30 --warmup 5 > gpurun_out/w_tr1.json 2> gpurun_out/w_tr1.err; cat gpurun_out/w_tr1.json | cut -c1-400
timeout 900 python bench.py --steps 30 --warmup 5 > gpurun_out/w_bench.json 2> gpurun_out/w_bench.err; tail -2 gpurun_out/w_bench.err; cat gpurun_out/w_bench.json | cut -c1-600
