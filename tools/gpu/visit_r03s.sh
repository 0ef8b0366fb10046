#!/bin/bash
# round 2, visit 3s: heavy buckets folded by a block (k_fold_heavy): skewed scalars
TAG=r03s
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu --timeout=800 -x -k "msm or skew or mul" > $OUT/pytest_msm.log 2>&1; echo "pytest msm exit $?" | tee -a $OUT/status.txt
tail -3 $OUT/pytest_msm.log
timeout 900 python tools/gpu/front_end_ab.py G1:20,G1:16,G2:18,G2:20 > $OUT/front_end_ab.txt 2>&1; echo "front end ab exit $?" | tee -a $OUT/status.txt
grep -v "front_end=sort\|front_end=count:" $OUT/front_end_ab.txt
