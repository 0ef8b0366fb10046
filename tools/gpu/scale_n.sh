#!/bin/bash
# bench.py under torchrun on N GPUs of this box (N from the environment) + the NCCL sharding test
mkdir -p gpurun_out/$TAG
N=${N:-2}
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600+N)) bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/$TAG/scale_n$N.json 2> gpurun_out/$TAG/scale_n$N.err
echo "N=$N exit $?"
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/$TAG/scale_n$N.json").read().strip().splitlines()[-1])
    r=d["roofline"]
    print("N=$N", "value=%.4g ms=%.3f e2e=%.4g e2e_ms=%.3f" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"]), "pairings/s=%.4g" % d["secondary"]["value"], "g2=%.4g" % d["secondary_g2_msm"]["value"], "bbs=%.4g" % d["secondary_bbs_plus_verify"]["value"])
    print(" multi_rank_result_ok", r.get("multi_rank_result_ok"), "strong_ms", r.get("strong_ms"), "strong_speedup", r.get("strong_speedup"), "strong_ok", r.get("strong_result_ok"), "g2 strong", r.get("g2_msm_strong_ms"), "numa", d["e2e"].get("numa_node_bound"))
except Exception as e: print("N=$N unreadable", e)
PY
timeout 900 python -m pytest tests -q -m gpu -x -k "distributed" 2>&1 | tail -2
