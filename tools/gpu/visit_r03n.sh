#!/bin/bash
# round 2, visit 3n: up to eight upload groups
TAG=r03n
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 900 python tools/gpu/groups_check.py > $OUT/groups_check.txt 2>&1; echo "groups check exit $?" | tee -a $OUT/status.txt
tail -6 $OUT/groups_check.txt
timeout 600 python tools/gpu/e2e_probe.py G1 20 > $OUT/e2e_probe_g1.txt 2>&1; echo "e2e g1 exit $?" | tee -a $OUT/status.txt
cat $OUT/e2e_probe_g1.txt
timeout 600 python tools/gpu/e2e_probe.py G2 18 > $OUT/e2e_probe_g2.txt 2>&1; echo "e2e g2 exit $?" | tee -a $OUT/status.txt
cat $OUT/e2e_probe_g2.txt
timeout 600 python tools/gpu/e2e_probe.py G1 22 > $OUT/e2e_probe_g1_22.txt 2>&1; echo "e2e g1 2^22 exit $?" | tee -a $OUT/status.txt
cat $OUT/e2e_probe_g1_22.txt
