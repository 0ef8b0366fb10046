#!/bin/bash
# round 2, visit r: thread-per-instance pairing kernel with called (not inlined) Fp2 products at 8 / 12 / 16 warps per SM
TAG=r02r
OUT=gpurun_out/$TAG
mkdir -p $OUT
C12381_LIB_VARIANT=pc timeout 900 python tools/gpu/pairing_ab.py 4 37888,65536 > $OUT/pairing_ab_pc.txt 2>&1; echo "ab[pc] exit $?" | tee -a $OUT/status.txt
C12381_LIB_VARIANT=pmb3c timeout 900 python tools/gpu/pairing_ab.py 4 56832,65536 > $OUT/pairing_ab_pmb3c.txt 2>&1; echo "ab[pmb3c] exit $?" | tee -a $OUT/status.txt
C12381_LIB_VARIANT=pmb4c timeout 900 python tools/gpu/pairing_ab.py 4 75776,65536 > $OUT/pairing_ab_pmb4c.txt 2>&1; echo "ab[pmb4c] exit $?" | tee -a $OUT/status.txt
for V in pc pmb3c pmb4c; do echo "== $V"; grep "thread-per" $OUT/pairing_ab_$V.txt; done
timeout 300 python tools/gpu/msm_time.py G1:20,G2:18 > $OUT/msm_time.txt 2>&1; cat $OUT/msm_time.txt
