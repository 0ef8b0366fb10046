#!/bin/bash
# round 2, visit a: probes of the integer-multiply pipe, the whole GPU suite, the new bench line, 128-bit gather A/B
TAG=r02a
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi > $OUT/nvidia-smi.txt 2>&1
timeout 300 tools/gpu/_bin/imad_probes 2000 > $OUT/imad_probes.jsonl 2>&1; echo "probes exit $?" | tee -a $OUT/status.txt
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke exit $?" | tee -a $OUT/status.txt
timeout 2400 python -m pytest tests -q -m gpu --timeout=1500 > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" | tee -a $OUT/status.txt
tail -15 $OUT/pytest_gpu.log
timeout 900 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit $?" | tee -a $OUT/status.txt
tail -5 $OUT/bench.err
C12381_LIB_VARIANT=noalign timeout 600 python bench.py --no-secondary --cpu-sample-log-n 12 > $OUT/bench_noalign.json 2>> $OUT/bench.err; echo "bench[noalign] exit $?" | tee -a $OUT/status.txt
timeout 600 python bench.py --no-secondary --cpu-sample-log-n 12 > $OUT/bench_align.json 2>> $OUT/bench.err; echo "bench[align] exit $?" | tee -a $OUT/status.txt
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_reference.json 2>> $OUT/bench.err; echo "bench ref exit $?" | tee -a $OUT/status.txt
python tools/gpu/sweep_probe.py > $OUT/sweep_probe.txt 2>&1
cat $OUT/status.txt
