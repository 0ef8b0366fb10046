#!/bin/bash
# round 2, visit t: ncu --set full of the thread-per-instance pairing kernel at 16 lockstep warps per SM (one 512-thread block)
TAG=r02t
OUT=gpurun_out/$TAG
mkdir -p $OUT
C12381_LIB_VARIANT=pt512 timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:k_pairing$' -s 1 -c 1 -o $OUT/prof_pairing_pt512 -f python tools/gpu/pairing_once.py 75776 > $OUT/ncu_pt512.log 2>&1; echo "ncu pt512 exit $?" | tee -a $OUT/status.txt
ncu -i $OUT/prof_pairing_pt512.ncu-rep --page raw --csv > $OUT/prof_pairing_pt512.raw.csv 2>/dev/null
ncu -i $OUT/prof_pairing_pt512.ncu-rep --page source --csv > $OUT/prof_pairing_pt512.source.csv 2>/dev/null
rm -f $OUT/prof_pairing_pt512.ncu-rep
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:k_pairing$' -s 1 -c 1 -o $OUT/prof_pairing_default -f python tools/gpu/pairing_once.py 37888 > $OUT/ncu_default.log 2>&1; echo "ncu default exit $?" | tee -a $OUT/status.txt
ncu -i $OUT/prof_pairing_default.ncu-rep --page raw --csv > $OUT/prof_pairing_default.raw.csv 2>/dev/null
ncu -i $OUT/prof_pairing_default.ncu-rep --page source --csv > $OUT/prof_pairing_default.source.csv 2>/dev/null
rm -f $OUT/prof_pairing_default.ncu-rep
ls -la $OUT
