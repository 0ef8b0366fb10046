"""GPU box: executed-instruction mix of one kernel from an ncu report's SASS source page.
python tools/gpu/opcode_mix.py <report.ncu-rep> <out.json>
Aggregates `ncu -i rep --page source --csv` (per-SASS-line 'Instructions Executed', warp-level) by opcode family."""
import collections, csv, io, json, subprocess, sys
rep, out = sys.argv[1], sys.argv[2]
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr = next(i for i, r in enumerate(rows) if "Source" in r and any("Instructions Executed" in c for c in r))
H = rows[hdr]
si = H.index("Source")
ei = next(i for i, c in enumerate(H) if c.strip() == "Instructions Executed")
wi = next((i for i, c in enumerate(H) if c.strip().startswith("Warp Stall Sampling (All")), None)
mix, stall = collections.Counter(), collections.Counter()
for r in rows[hdr + 1:]:
    if len(r) <= max(si, ei):
        continue
    op = r[si].strip().split()
    if not op:
        continue
    name = op[1] if op[0].startswith("@") and len(op) > 1 else op[0]
    parts = name.split(".")
    fam = ".".join(parts[:2]) if parts[0] in ("IMAD", "IADD3", "LDL", "STL", "LDG", "STG", "SHFL", "LOP3", "ISETP") and len(parts) > 1 and parts[1] in ("WIDE", "X", "HI", "MOV", "IADD", "SHL", "U32", "LUT") else parts[0]
    try:
        n = int(float(r[ei].replace(",", "") or 0))
    except ValueError:
        continue
    mix[fam] += n
    if wi is not None:
        try:
            stall[fam] += int(float(r[wi].replace(",", "") or 0))
        except ValueError:
            pass
tot = sum(mix.values())
res = {"total_warp_instructions": tot, "mix": {k: [v, round(100.0 * v / tot, 2)] for k, v in mix.most_common(25)},
       "stall_samples": {k: v for k, v in stall.most_common(15)}}
json.dump(res, open(out, "w"), indent=1)
print(json.dumps(res, indent=1))
