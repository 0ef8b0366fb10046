#!/bin/bash
# round 2, visit 3c: window choice at the small end of the sweep; latency of the batch-of-1 drop-ins
TAG=r03c
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 600 python tools/gpu/latency_probe.py > $OUT/latency_probe.txt 2>&1; echo "latency exit $?" | tee -a $OUT/status.txt
cat $OUT/latency_probe.txt
timeout 900 python tools/gpu/window_sweep.py G1 > $OUT/window_sweep_g1.txt 2>&1; echo "window sweep g1 exit $?" | tee -a $OUT/status.txt
grep -- "auto\|->" $OUT/window_sweep_g1.txt
timeout 900 python tools/gpu/window_sweep.py G2 > $OUT/window_sweep_g2.txt 2>&1; echo "window sweep g2 exit $?" | tee -a $OUT/status.txt
grep -- "auto\|->" $OUT/window_sweep_g2.txt
