#!/bin/bash
# One GPU-box visit: smoke, parity tests, bench, then (only after the plain bench exited 0) the ncu passes.
# Usage (from the repo root on the box):  bash tools/gpu_round.sh [tag]
TAG=${1:-r01}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi > $OUT/nvidia-smi.txt 2>&1
timeout 600 python __graft_entry__.py --smoke > $OUT/smoke.log 2>&1; echo "smoke exit $?" | tee -a $OUT/status.txt
timeout 2400 python -m pytest tests -q -m gpu -x --timeout=1500 > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" | tee -a $OUT/status.txt
tail -5 $OUT/pytest_gpu.log
timeout 900 python bench.py > $OUT/bench.json 2> $OUT/bench.err; BE=$?; echo "bench exit $BE" | tee -a $OUT/status.txt
tail -c 1500 $OUT/bench.json; tail -5 $OUT/bench.err
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_reference.json 2>> $OUT/bench.err; echo "bench ref exit $?" | tee -a $OUT/status.txt
if [ "${SKIP_NCU:-0}" = "0" ] && [ $BE -eq 0 ]; then
  timeout 600 python bench.py --steps 2 --warmup 1 --no-secondary --cpu-sample-log-n 10 > $OUT/plain_for_ncu.log 2>&1 &&
  timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv \
      python bench.py --steps 2 --warmup 1 --no-secondary --cpu-sample-log-n 10 > $OUT/ncu_launches.log 2>&1
  echo "ncu launches exit $?" | tee -a $OUT/status.txt
  timeout 1500 ncu --set full --clock-control none --import-source on -k regex:k_accumulate -s 3 -c 1 -o $OUT/prof_accumulate -f \
      python bench.py --steps 2 --warmup 1 --no-secondary --cpu-sample-log-n 10 > $OUT/ncu_full.log 2>&1
  echo "ncu full exit $?" | tee -a $OUT/status.txt
fi
cat $OUT/status.txt
