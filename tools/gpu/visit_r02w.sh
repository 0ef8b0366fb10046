#!/bin/bash
# round 2, visit w: round 0 per upload group, rounds >= 1 on the merged lists
TAG=r02w
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu --timeout=800 -x -k "msm or skew" > $OUT/pytest_msm.log 2>&1; echo "pytest msm exit $?" | tee -a $OUT/status.txt
tail -3 $OUT/pytest_msm.log
timeout 600 python tools/gpu/e2e_probe.py G1 20 > $OUT/e2e_probe_g1.txt 2>&1; echo "e2e g1 exit $?" | tee -a $OUT/status.txt
cat $OUT/e2e_probe_g1.txt
timeout 600 python tools/gpu/e2e_probe.py G2 18 > $OUT/e2e_probe_g2.txt 2>&1; echo "e2e g2 exit $?" | tee -a $OUT/status.txt
cat $OUT/e2e_probe_g2.txt
timeout 600 python tools/gpu/msm_time.py G1:20,G1:18,G1:22,G2:18,G2:20 > $OUT/msm_time.txt 2>&1; cat $OUT/msm_time.txt
