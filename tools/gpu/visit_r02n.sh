#!/bin/bash
# round 2, visit n: whole GPU suite, the bench line with the new accounting, launch list of the bench command
TAG=r02n
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke exit $?" | tee -a $OUT/status.txt
timeout 1500 python -m pytest tests -q -m gpu --timeout=1400 > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" | tee -a $OUT/status.txt
tail -4 $OUT/pytest_gpu.log
timeout 900 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit $?" | tee -a $OUT/status.txt
tail -3 $OUT/bench.err
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_reference.json 2>> $OUT/bench.err; echo "bench ref exit $?" | tee -a $OUT/status.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/launches.csv python bench.py --no-secondary --steps 2 --warmup 1 --cpu-sample-log-n 10 > $OUT/ncu_bench.log 2>&1; echo "ncu exit $?" | tee -a $OUT/status.txt
python tools/gpu/sweep_probe.py > $OUT/sweep_probe.txt 2>&1
cat $OUT/status.txt
