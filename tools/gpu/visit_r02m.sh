#!/bin/bash
# round 2, visit m: ncu --set full on the plane sums, the finish kernel and the halving rounds' unwinding pass
TAG=r02m
OUT=gpurun_out/$TAG
mkdir -p $OUT
CMD="python tools/gpu/msm_once.py G1 20 -1 2 2"
for K in k_reduce_planes k_finish k_reduce_level0 k_ba_bwd k_ba_fwd; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$K -s 0 -c 1 -o $OUT/prof_$K -f $CMD > $OUT/ncu_$K.log 2>&1
  echo "ncu $K exit $?" | tee -a $OUT/status.txt
  ncu -i $OUT/prof_$K.ncu-rep --page raw --csv > $OUT/prof_$K.raw.csv 2>/dev/null
  ncu -i $OUT/prof_$K.ncu-rep --page source --csv > $OUT/prof_$K.source.csv 2>/dev/null
  rm -f $OUT/prof_$K.ncu-rep
done
ls -la $OUT
