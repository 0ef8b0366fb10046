SKIP_NCU=1 bash tools/gpu_round.sh r01f
python - <<'PY'
import json
d=json.load(open("gpurun_out/r01f/bench.json"))
print("phases", d["roofline"].get("phases_ms"))
print("hbm", d["roofline"].get("hbm_phase"))
print("secondary", {k:v for k,v in d["secondary"].items() if k!="cpu_baseline"})
PY
