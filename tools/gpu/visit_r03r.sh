#!/bin/bash
# round 2, visit 3r: warp-wide Fp2 products in the Horner chain (k_finish, k_fixed_base_windows); G2 auto rounds from 2^21 entries
TAG=r03r
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_bbs_plus.py -q -m gpu --timeout=800 -x > $OUT/pytest.log 2>&1; echo "pytest exit $?" | tee -a $OUT/status.txt
tail -3 $OUT/pytest.log
timeout 600 python tools/gpu/msm_time.py G2:10,G2:14,G2:17,G2:18,G2:20,G1:20 > $OUT/msm_time.txt 2>&1; cat $OUT/msm_time.txt
timeout 600 python tools/gpu/window_sweep.py G2 1,10,18 13,16 > $OUT/window_sweep_g2.txt 2>&1; grep auto $OUT/window_sweep_g2.txt
timeout 600 python tools/gpu/latency_probe.py > $OUT/latency_probe.txt 2>&1; cat $OUT/latency_probe.txt
timeout 600 python tools/gpu/bbs_probe.py > $OUT/bbs_probe.txt 2>&1; tail -12 $OUT/bbs_probe.txt
