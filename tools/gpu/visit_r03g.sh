#!/bin/bash
# round 2, visit 3g: split tail A/B
TAG=r03g
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 900 python tools/gpu/split_tail_ab.py G1:20,G1:18,G1:22,G2:18,G2:20 > $OUT/split_tail_ab.txt 2>&1; echo "split tail ab exit $?" | tee -a $OUT/status.txt
cat $OUT/split_tail_ab.txt
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu --timeout=800 -x -k "msm or skew or mul" > $OUT/pytest_msm.log 2>&1; echo "pytest msm exit $?" | tee -a $OUT/status.txt
tail -3 $OUT/pytest_msm.log
timeout 900 python tools/gpu/groups_check.py > $OUT/groups_check.txt 2>&1; echo "groups check exit $?" | tee -a $OUT/status.txt
tail -3 $OUT/groups_check.txt
