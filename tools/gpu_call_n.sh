#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
for flags in "-DGS_LOSS_PACKED=0 -DGS_LOSS_FASTLOAD=0" "-DGS_LOSS_PACKED=1 -DGS_LOSS_FASTLOAD=0" "-DGS_LOSS_PACKED=0 -DGS_LOSS_FASTLOAD=1" "-DGS_LOSS_PACKED=1 -DGS_LOSS_FASTLOAD=1"; do
  B200GS_LOSS_FLAGS="$flags" python 3d-gaussian-splatting-for-novel-view-synthesis_b200/build.py --force > /dev/null 2>&1
  echo "== $flags"
  timeout 300 python tools/loss_bench.py 2>&1 | tail -1
done
python 3d-gaussian-splatting-for-novel-view-synthesis_b200/build.py --force > /dev/null 2>&1
