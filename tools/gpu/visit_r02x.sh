#!/bin/bash
TAG=r02x
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 900 python tools/gpu/groups_check.py > $OUT/groups_check.txt 2>&1; echo "groups check exit $?" | tee -a $OUT/status.txt
cat $OUT/groups_check.txt
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu --timeout=800 -x -k "msm or skew" > $OUT/pytest_msm.log 2>&1; echo "pytest msm exit $?" | tee -a $OUT/status.txt
tail -3 $OUT/pytest_msm.log
