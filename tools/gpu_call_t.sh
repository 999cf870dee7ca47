#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
L=tools/sortlab
{
for cfg in wide:512:1 narrow:256:2; do
IFS=: read sh th bps <<< "$cfg"
echo "===== shape $sh bps $bps"
export B200GS_SORT_SHAPE=$sh B200GS_SORT_BPS=$bps
timeout 60 $L/sort_lab_trace 1000000 9 $th 1000000 0 $bps | head -12
timeout 60 $L/sort_lab_trace 4000 9 $th 4000 0 $bps | head -12
timeout 120 $L/sort_lab_mb3 20
done
} > gpurun_out/t_sorttrace10.log 2>&1
cat gpurun_out/t_sorttrace10.log
