#!/bin/bash
# round 2, visit q: halving-round kernels at 4 resident blocks (128 registers) and with L1 instead of L2 prefetches
TAG=r02q
OUT=gpurun_out/$TAG
mkdir -p $OUT
for V in "" ba4 bal1 ba4l1 ""; do
  C12381_LIB_VARIANT=$V timeout 600 python tools/gpu/msm_time.py G1:20,G1:22,G2:18 >> $OUT/ba_variants.txt 2>&1; echo "[$V] exit $?" | tee -a $OUT/status.txt
done
cat $OUT/ba_variants.txt
