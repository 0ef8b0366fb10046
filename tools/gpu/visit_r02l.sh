#!/bin/bash
# round 2, visit l: plane-sum bucket reduction - whole parity suite, tail by segment length, launch list
TAG=r02l
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu --timeout=800 -x -k "msm or skew" > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" | tee -a $OUT/status.txt
tail -4 $OUT/pytest_gpu.log
timeout 600 python tools/gpu/tail_tune.py > $OUT/tail_tune.txt 2>&1; echo "tail exit $?" | tee -a $OUT/status.txt
cat $OUT/tail_tune.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv python tools/gpu/msm_once.py G1 20 -1 2 2 > $OUT/ncu.log 2>&1; echo "ncu exit $?" | tee -a $OUT/status.txt
