#!/bin/bash
# GPU call A (1 GPU): parity tests, bench (all modes at N=1), reference-script run with a profile of the fused-Adam path.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/a_smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/a_pytest.log
tail -5 gpurun_out/a_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/a_bench.json 2> gpurun_out/a_bench.err; echo "bench rc=$?"
tail -c 600 gpurun_out/a_bench.err
timeout 300 python bench.py --mode tile_rows --steps 20 --warmup 5 > gpurun_out/a_bench_tile_rows.json 2> gpurun_out/a_bench_tile_rows.err; echo "tile_rows rc=$?"
timeout 300 python bench.py --mode train --steps 20 --warmup 5 > gpurun_out/a_bench_train.json 2> gpurun_out/a_bench_train.err; echo "train rc=$?"
timeout 400 python tools/run_reference_scripts.py --profile --iterations 150 > gpurun_out/a_ref_scripts.log 2>&1; echo "ref scripts rc=$?"
timeout 120 python tools/d2h_probe.py > gpurun_out/a_d2h_1.json 2>&1
echo done
