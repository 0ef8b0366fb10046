#!/bin/bash
# round 2, visit s: thread-per-instance pairing kernel as ONE lockstep block of 384 / 512 threads per SM
TAG=r02s
OUT=gpurun_out/$TAG
mkdir -p $OUT
C12381_LIB_VARIANT=pt384 timeout 900 python tools/gpu/pairing_ab.py 4 56832,65536 > $OUT/pairing_ab_pt384.txt 2>&1; echo "ab[pt384] exit $?" | tee -a $OUT/status.txt
C12381_LIB_VARIANT=pt384c timeout 900 python tools/gpu/pairing_ab.py 4 56832,65536 > $OUT/pairing_ab_pt384c.txt 2>&1; echo "ab[pt384c] exit $?" | tee -a $OUT/status.txt
C12381_LIB_VARIANT=pt512 timeout 900 python tools/gpu/pairing_ab.py 4 75776,65536 > $OUT/pairing_ab_pt512.txt 2>&1; echo "ab[pt512] exit $?" | tee -a $OUT/status.txt
for V in pt384 pt384c pt512; do echo "== $V"; grep "thread-per" $OUT/pairing_ab_$V.txt; done
