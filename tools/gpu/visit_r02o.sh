#!/bin/bash
# round 2, visit o: thread-per-instance pairing kernel at higher occupancy, now that the blocks run in lockstep
TAG=r02o
OUT=gpurun_out/$TAG
mkdir -p $OUT
for V in "" pmb3 pt192 pt256; do
  C12381_LIB_VARIANT=$V timeout 900 python tools/gpu/pairing_ab.py 4 16384,37888,56832,65536 > $OUT/pairing_ab_${V:-default}.txt 2>&1; echo "ab[$V] exit $?" | tee -a $OUT/status.txt
  echo "== ${V:-default}"; grep "thread-per" $OUT/pairing_ab_${V:-default}.txt
done
