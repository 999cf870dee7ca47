#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -k "grad_bucket or deferred or non_fp32 or mismatched or golden" > gpurun_out/e_pytest.log 2>&1; tail -3 gpurun_out/e_pytest.log
timeout 300 python tools/band_times.py > gpurun_out/e_band_times.log 2>&1; echo "band rc=$?"
B200GS_RUN_TRACE=2 timeout 400 python tools/run_reference_scripts.py --iterations 150 > gpurun_out/e_ref_scripts_timeline.log 2>&1; echo "ref scripts rc=$?"
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/e_bench.json 2> gpurun_out/e_bench.err; echo "bench rc=$?"
# ncu: launch list of the profile target, then one full capture of every kernel of 2 frames + 1 training iteration
timeout 300 python tools/profile_target.py 2 1 > gpurun_out/e_target.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/e_launches.csv python tools/profile_target.py 3 2 > gpurun_out/e_ncu_launches.log 2>&1; echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gs::|stats_to_host" -c 60 -o gpurun_out/e_full python tools/profile_target.py 2 1 > gpurun_out/e_ncu_full.log 2>&1; echo "ncu full rc=$?"
ncu -i gpurun_out/e_full.ncu-rep --page raw --csv > gpurun_out/e_full_raw.csv 2>/dev/null; echo "export rc=$?"
ls -la gpurun_out/e_full* | head
echo done
