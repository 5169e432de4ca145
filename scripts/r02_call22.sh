#!/bin/bash
# 1-GPU call: small-p kernels vs the segmented general-p kernels on the C2 / C5 / C3 shapes and two more row lengths
cd "$(dirname "$0")/.."; O=gpurun_out/r02_v; mkdir -p $O
for s in C3 C5 M300 M1000p2 C2; do
  timeout 200 python scripts/shape_ab.py $s 2>> $O/err.txt | tail -1
  B200NB_FORCE_GENERIC=1 timeout 200 python scripts/shape_ab.py $s 2>> $O/err.txt | tail -1
done > $O/shape_ab.txt; cat $O/shape_ab.txt; tail -3 $O/err.txt
