mkdir -p gpurun_out/$TAG
timeout 600 python tools/_pairing_bench.py 4 16384,65536 2>&1 | tee gpurun_out/$TAG/pairing_bench.txt | tail -6
TAG=$TAG SKIP_NCU=1 bash tools/_gpu_quick.sh
python - <<PY
import json
d=json.load(open("gpurun_out/$TAG/bench.json"))
print(" sweep", [(s["log_n"], round(s["ms"],3), "%.3g" % s["points_per_s"], s["window_bits"]) for s in d.get("secondary_g1_sweep",[])])
print(" cpu", d.get("cpu_baseline"))
PY
