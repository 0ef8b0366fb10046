#!/bin/bash
# round 2, visit p (2 GPUs): the multi-rank paths after the batch-affine / upload-group changes
TAG=r02p
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi -L > $OUT/gpus.txt
timeout 900 python -m pytest tests -q -m gpu -x -k "distributed" > $OUT/pytest_distributed.log 2>&1; echo "pytest distributed exit $?" | tee -a $OUT/status.txt
tail -3 $OUT/pytest_distributed.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29602 bench.py --gpus 2 --steps 10 --warmup 3 > $OUT/scale_n2.json 2> $OUT/scale_n2.err; echo "bench n2 exit $?" | tee -a $OUT/status.txt
tail -3 $OUT/scale_n2.err
python - <<PY
import json
d=json.loads(open("$OUT/scale_n2.json").read().strip().splitlines()[-1])
r=d["roofline"]
print("N=2 value=%.4g ms=%.3f e2e=%.4g (%.3f ms)" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"]), "multi_rank_ok", r.get("multi_rank_result_ok"), "strong", r.get("strong_ms"), r.get("strong_speedup"), r.get("strong_result_ok"), "pairings/s", r.get("pairings_per_s"), "g2", r.get("g2_msm_points_per_s"), r.get("g2_msm_multi_rank_result_ok"), "bbs", r.get("bbs_verifications_per_s"))
PY
