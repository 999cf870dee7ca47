#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/b_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/b_pytest.log
tail -8 gpurun_out/b_pytest.log
timeout 300 python tools/adam_sweep.py > gpurun_out/b_adam_sweep.jsonl 2>&1; echo "sweep rc=$?"
echo done
