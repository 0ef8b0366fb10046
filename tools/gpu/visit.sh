TAG=${TAG:-r01g}
VARIANTS="${VARIANTS}" bash tools/gpu_round.sh $TAG
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/$TAG/bench*.json")):
    d=json.load(open(f))
    if "roofline" not in d: continue
    print(f)
    print(" phases", {k: round(v,3) for k,v in d["roofline"].get("phases_ms",{}).items()})
    print(" e2e ms", d["e2e"].get("ms_per_step"), " hbm", d["roofline"].get("hbm_phase",{}).get("achieved_gbs"))
    s=d.get("secondary",{})
    print(" pairings/s", s.get("value"), "ms", s.get("ms"), "frac", s.get("frac_of_int32_mad_peak"))
    print(" g2", d.get("secondary_g2_msm"))
PY
python - <<PY
import json
d=json.load(open("gpurun_out/$TAG/bench.json"))
print(" bbs", d.get("secondary_bbs_plus_verify"))
PY
