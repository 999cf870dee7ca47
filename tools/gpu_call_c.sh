#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/c_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c_pytest.log
tail -6 gpurun_out/c_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/c_bench_ppl2_split128.json 2> gpurun_out/c_bench.err; echo "bench rc=$?"
B200GS_BWD_PPL=1 B200GS_SPLIT_THREADS=256 timeout 600 python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/c_bench_ppl1_split256.json 2>> gpurun_out/c_bench.err; echo "bench rc=$?"
B200GS_BWD_PPL=4 timeout 600 python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/c_bench_ppl4.json 2>> gpurun_out/c_bench.err; echo "bench rc=$?"
B200GS_RUN_TRACE=1 timeout 400 python tools/run_reference_scripts.py --iterations 150 > gpurun_out/c_ref_scripts.log 2>&1; echo "ref scripts rc=$?"
echo done
