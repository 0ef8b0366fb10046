#!/bin/bash
# One GPU-box visit: smoke, parity tests, bench (+ optional variant builds), then — only after the plain bench
# exited 0 — the ncu passes.   Usage (repo root, on the box):  [SKIP_NCU=1] [VARIANTS="call"] bash tools/gpu_round.sh <tag>
TAG=${1:-r01}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi > $OUT/nvidia-smi.txt 2>&1
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke exit $?" | tee -a $OUT/status.txt
timeout 2400 python -m pytest tests -q -m gpu -x --timeout=1500 > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" | tee -a $OUT/status.txt
tail -5 $OUT/pytest_gpu.log
timeout 900 python bench.py > $OUT/bench.json 2> $OUT/bench.err; BE=$?; echo "bench exit $BE" | tee -a $OUT/status.txt
tail -5 $OUT/bench.err
for V in $VARIANTS; do
  C12381_LIB_VARIANT=$V timeout 900 python bench.py > $OUT/bench_$V.json 2>> $OUT/bench.err; echo "bench[$V] exit $?" | tee -a $OUT/status.txt
done
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_reference.json 2>> $OUT/bench.err; echo "bench ref exit $?" | tee -a $OUT/status.txt
python - <<PY
import json,glob
for f in sorted(glob.glob("$OUT/bench*.json")):
    try:
        d=json.load(open(f)); r=d.get("roofline",{}); s=d.get("secondary",{})
        print(f, "value=%.3g ms=%.3f e2e=%.3g acc_ms=%s frac=%s pairings/s=%s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], r.get("kernel_ms"), r.get("frac"), s.get("value")))
    except Exception as e: print(f, "unreadable", e)
PY
if [ "${SKIP_NCU:-0}" = "0" ] && [ $BE -eq 0 ]; then
  CMD="python bench.py --steps 2 --warmup 1 --cpu-sample-log-n 10 --pairing-instances 16384 --g2-log-n 14 --bbs-log-b 10 --sweep-max-log-n 12"
  timeout 600 $CMD > $OUT/plain_for_ncu.log 2>&1 &&
  timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv $CMD > $OUT/ncu_launches.log 2>&1
  echo "ncu launches exit $?" | tee -a $OUT/status.txt
  timeout 1500 ncu --set full --clock-control none --import-source on -k regex:k_accumulate -s 2 -c 1 -o $OUT/prof_accumulate -f $CMD > $OUT/ncu_full_acc.log 2>&1
  echo "ncu full accumulate exit $?" | tee -a $OUT/status.txt
  # the dominant launch of the headline step: the first unwinding pass of the halving rounds (n = 2^20, default settings)
  timeout 600 python tools/gpu/msm_once.py G1 20 > $OUT/plain_msm_once.log 2>&1 &&
  timeout 1500 ncu --set full --clock-control none --import-source on -k regex:k_ba_bwd -s 7 -c 1 -o $OUT/prof_k_ba_bwd -f python tools/gpu/msm_once.py G1 20 > $OUT/ncu_full_ba_bwd.log 2>&1
  echo "ncu full k_ba_bwd exit $?" | tee -a $OUT/status.txt
  ncu -i $OUT/prof_k_ba_bwd.ncu-rep --page raw --csv > $OUT/prof_k_ba_bwd.raw.csv 2>/dev/null && rm -f $OUT/prof_k_ba_bwd.ncu-rep
  timeout 1500 ncu --set full --clock-control none --import-source on -k regex:k_pairing_coop -s 1 -c 1 -o $OUT/prof_pairing_coop -f $CMD > $OUT/ncu_full_pair_coop.log 2>&1
  echo "ncu full pairing (cooperative, 2^14 instances) exit $?" | tee -a $OUT/status.txt
  ncu -i $OUT/prof_pairing_coop.ncu-rep --page raw --csv > $OUT/prof_pairing_coop.raw.csv 2>/dev/null && rm -f $OUT/prof_pairing_coop.ncu-rep
  CMD2="python bench.py --steps 2 --warmup 1 --cpu-sample-log-n 10 --pairing-instances 65536 --g2-log-n 14 --bbs-log-b 0 --sweep-max-log-n 0"
  timeout 1500 ncu --set full --clock-control none --import-source on -k 'regex:k_pairing$' -s 1 -c 1 -o $OUT/prof_pairing -f $CMD2 > $OUT/ncu_full_pair.log 2>&1
  echo "ncu full pairing (thread-per-instance, 2^16 instances) exit $?" | tee -a $OUT/status.txt
  ncu -i $OUT/prof_pairing.ncu-rep --page raw --csv > $OUT/prof_pairing.raw.csv 2>/dev/null && rm -f $OUT/prof_pairing.ncu-rep   # gpurun_out/ is capped at 64 MiB
  ncu -i $OUT/prof_accumulate.ncu-rep --page raw --csv > $OUT/prof_accumulate.raw.csv 2>/dev/null
fi
cat $OUT/status.txt
