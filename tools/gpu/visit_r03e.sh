#!/bin/bash
# round 2, visit 3e: scatter with one gather, merged plan beside the scatter
TAG=r03e
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 900 python tools/gpu/front_end_ab.py G1:20,G1:16,G1:10,G2:18,G2:20 > $OUT/front_end_ab.txt 2>&1; echo "front end ab exit $?" | tee -a $OUT/status.txt
grep -v "equal\|small" $OUT/front_end_ab.txt
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu --timeout=800 -x -k "msm or skew or mul" > $OUT/pytest_msm.log 2>&1; echo "pytest msm exit $?" | tee -a $OUT/status.txt
tail -3 $OUT/pytest_msm.log
timeout 900 python tools/gpu/groups_check.py > $OUT/groups_check.txt 2>&1; echo "groups check exit $?" | tee -a $OUT/status.txt
tail -3 $OUT/groups_check.txt
timeout 600 python tools/gpu/e2e_probe.py G1 20 > $OUT/e2e_probe_g1.txt 2>&1; echo "e2e g1 exit $?" | tee -a $OUT/status.txt
cat $OUT/e2e_probe_g1.txt
timeout 600 python tools/gpu/e2e_probe.py G2 18 > $OUT/e2e_probe_g2.txt 2>&1; echo "e2e g2 exit $?" | tee -a $OUT/status.txt
cat $OUT/e2e_probe_g2.txt
