#!/bin/bash
# round 2, visit g: batch-affine rounds after the region fix - parity subset, per-kernel durations (ncu launch list), A/B
TAG=r02g
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu --timeout=800 -x -k "msm or skew" > $OUT/pytest_msm.log 2>&1; echo "pytest msm exit $?" | tee -a $OUT/status.txt
tail -5 $OUT/pytest_msm.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_ba.csv python tools/gpu/msm_once.py G1 20 7 1 2 > $OUT/ncu_ba.log 2>&1; echo "ncu exit $?" | tee -a $OUT/status.txt
timeout 900 python tools/gpu/msm_ab.py G1:20 > $OUT/msm_ab.txt 2>&1; echo "ab exit $?" | tee -a $OUT/status.txt
cat $OUT/msm_ab.txt
