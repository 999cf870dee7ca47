#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 300 python tools/sanitize_target.py > gpurun_out/m_sanitize_plain.log 2>&1; echo "plain rc=$?"; tail -3 gpurun_out/m_sanitize_plain.log
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 7 python tools/sanitize_target.py > gpurun_out/m_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -6 gpurun_out/m_memcheck.log
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 7 python tools/sanitize_target.py > gpurun_out/m_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -4 gpurun_out/m_racecheck.log
echo done
