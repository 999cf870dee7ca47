#!/bin/bash
# builds tools/sortlab/sort_lab_mb{2,3} (ONESWEEP_MIN_BLOCKS variants) for sm_100a
cd "$(dirname "$0")"
CS=../../3d-gaussian-splatting-for-novel-view-synthesis_b200/csrc
for mb in 2 3; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo --expt-relaxed-constexpr -I ../../include \
       -DONESWEEP_MIN_BLOCKS=$mb sort_lab.cu $CS/scan_sort.cu -o sort_lab_mb$mb || exit 1
done
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo --expt-relaxed-constexpr -I ../../include \
     sort_trace.cu -o sort_lab_trace || exit 1
