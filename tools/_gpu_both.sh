mkdir -p gpurun_out/$TAG
timeout 600 python tools/_sweep_probe.py 2>&1 | tee gpurun_out/$TAG/sweep_probe.txt | tail -20 | cut -c1-230
TAG=$TAG SKIP_NCU=1 bash tools/_gpu_quick.sh
python - <<PY
import json
d=json.load(open("gpurun_out/$TAG/bench.json"))
print(" sweep", [(s["log_n"], round(s["ms"],3), "%.3g" % s["points_per_s"], s["window_bits"]) for s in d.get("secondary_g1_sweep",[])])
PY
