mkdir -p gpurun_out/$TAG
NG=$(nvidia-smi -L | wc -l)
echo "gpus: $NG"
for N in 1 2 4 8; do
  [ $N -le $NG ] || continue
  if [ $N -eq 1 ]; then
    timeout 900 python bench.py --gpus 1 --steps 10 --warmup 3 > gpurun_out/$TAG/scale_n1.json 2> gpurun_out/$TAG/scale_n1.err
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600+N)) bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/$TAG/scale_n$N.json 2> gpurun_out/$TAG/scale_n$N.err
  fi
  echo "N=$N exit $?"
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/$TAG/scale_n$N.json").read().strip().splitlines()[-1])
    print("N=$N", "value=%.4g ms=%.3f e2e=%.4g" % (d["value"], d["ms_per_step"], d["e2e"]["value"]), "pairings/s=%.4g" % d["secondary"]["value"], "g2=%.4g" % d["secondary_g2_msm"]["value"], "bbs=%.4g" % d["secondary_bbs_plus_verify"]["value"])
except Exception as e: print("N=$N unreadable", e)
PY
done
timeout 900 python -m pytest tests -q -m gpu -x -k "distributed" 2>&1 | tail -2
