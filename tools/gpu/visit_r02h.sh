#!/bin/bash
# round 2, visit h: knob sweep of the halving rounds, default build and the block-lockstep variant
TAG=r02h
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu --timeout=800 -x -k "msm_golden or skew" > $OUT/pytest_msm.log 2>&1; echo "pytest msm exit $?" | tee -a $OUT/status.txt
tail -3 $OUT/pytest_msm.log
timeout 900 python tools/gpu/ba_tune.py G1:20 > $OUT/ba_tune.txt 2>&1; echo "tune exit $?" | tee -a $OUT/status.txt
C12381_LIB_VARIANT=balock timeout 900 python tools/gpu/ba_tune.py G1:20 > $OUT/ba_tune_lockstep.txt 2>&1; echo "tune[lockstep] exit $?" | tee -a $OUT/status.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_ba.csv python tools/gpu/msm_once.py G1 20 -1 1 2 > $OUT/ncu_ba.log 2>&1; echo "ncu exit $?" | tee -a $OUT/status.txt
sort -t: -k3 -n $OUT/ba_tune.txt | head -50
