#!/bin/bash
# round 2, visit 3k: timing experiment - the pairing tower with unreduced additions (wrong values; upper bound of what bounds-tracked lazy additions could give)
TAG=r03k
OUT=gpurun_out/$TAG
mkdir -p $OUT
PAIRING_TPI_ONLY=1 timeout 600 python tools/gpu/pairing_ab.py 4 37888,65536 > $OUT/pairing_default.txt 2>&1; echo "default exit $?" | tee -a $OUT/status.txt
cat $OUT/pairing_default.txt
PAIRING_TPI_ONLY=1 C12381_LIB_VARIANT=lazyexp timeout 600 python tools/gpu/pairing_ab.py 4 37888,65536 > $OUT/pairing_lazyexp.txt 2>&1; echo "lazyexp exit $?" | tee -a $OUT/status.txt
cat $OUT/pairing_lazyexp.txt
