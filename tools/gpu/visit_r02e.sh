#!/bin/bash
# round 2, visit e: first run of the flat batch-affine halving rounds - parity subset, then the A/B
TAG=r02e
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu --timeout=800 -x -k "msm or skew" > $OUT/pytest_msm.log 2>&1; echo "pytest msm exit $?" | tee -a $OUT/status.txt
tail -15 $OUT/pytest_msm.log
timeout 900 python tools/gpu/msm_ab.py > $OUT/msm_ab.txt 2>&1; echo "ab exit $?" | tee -a $OUT/status.txt
cat $OUT/msm_ab.txt
