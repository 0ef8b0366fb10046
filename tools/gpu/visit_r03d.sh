#!/bin/bash
# round 2, visit 3d: window choice around the 2^16-entry boundary, batch-of-1 latencies after the one-term-MSM route, all GPU tests,
# launch list + ncu captures of the counting front end
TAG=r03d
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 600 python tools/gpu/latency_probe.py > $OUT/latency_probe.txt 2>&1; echo "latency exit $?" | tee -a $OUT/status.txt
cat $OUT/latency_probe.txt
timeout 900 python tools/gpu/window_sweep.py G1 13,14,15,16,17 12,13,16 > $OUT/window_sweep_g1.txt 2>&1; echo "window sweep g1 exit $?" | tee -a $OUT/status.txt
cat $OUT/window_sweep_g1.txt
timeout 900 python tools/gpu/window_sweep.py G2 12,13,14,15,16,17 13,16 > $OUT/window_sweep_g2.txt 2>&1; echo "window sweep g2 exit $?" | tee -a $OUT/status.txt
cat $OUT/window_sweep_g2.txt
timeout 2400 python -m pytest tests -q -m gpu -x --timeout=1500 > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" | tee -a $OUT/status.txt
tail -3 $OUT/pytest_gpu.log
CMD="python tools/gpu/msm_once.py G1 20"
timeout 300 $CMD > $OUT/plain_for_ncu.log 2>&1; echo "plain exit $?" | tee -a $OUT/status.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv $CMD > $OUT/ncu_launches.log 2>&1; echo "ncu launches exit $?" | tee -a $OUT/status.txt
for K in k_recode_count k_bucket_scatter; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$K -s 1 -c 1 -o $OUT/prof_$K -f $CMD > $OUT/ncu_full_$K.log 2>&1; echo "ncu $K exit $?" | tee -a $OUT/status.txt
  ncu -i $OUT/prof_$K.ncu-rep --page raw --csv > $OUT/prof_$K.raw.csv 2>/dev/null && rm -f $OUT/prof_$K.ncu-rep
done
