#!/bin/bash
TAG=r02v
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 600 python tools/gpu/e2e_sweep.py 20 > $OUT/e2e_sweep.txt 2>&1; echo "sweep exit $?" | tee -a $OUT/status.txt
cat $OUT/e2e_sweep.txt
