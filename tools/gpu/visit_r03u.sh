#!/bin/bash
# round 2, visit 3u: ncu captures of the probes the roofline denominators come from (VERDICT r01 task 4)
TAG=r03u
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 300 python tools/gpu/probe_once.py > $OUT/probe_plain.txt 2>&1; echo "plain exit $?" | tee -a $OUT/status.txt
cat $OUT/probe_plain.txt
for K in k_probe_imad k_probe_madc k_probe_wide k_probe_fp; do
  timeout 600 ncu --set full --clock-control none -k regex:$K -s 1 -c 1 -o $OUT/prof_$K -f python tools/gpu/probe_once.py > $OUT/ncu_$K.log 2>&1; echo "ncu $K exit $?" | tee -a $OUT/status.txt
  ncu -i $OUT/prof_$K.ncu-rep --page raw --csv > $OUT/prof_$K.raw.csv 2>/dev/null && rm -f $OUT/prof_$K.ncu-rep
done
