"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list: the kernels of the LAST MSM call, in order.
    python tools/launch_list.py gpurun_out/<tag>/launches.csv [first-kernel-substring]"""
import csv, re, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
rows = list(csv.DictReader(lines))
key = sys.argv[2] if len(sys.argv) > 2 else "k_recode"
idx = [i for i, x in enumerate(rows) if key in x["Kernel Name"]]
tot = 0.0
for x in rows[idx[-1] if idx else 0:]:
    v = float(x["Metric Value"].replace(",", ""))
    u = x["Metric Unit"]
    v = v / 1000 if u == "ns" else v * 1000 if u == "ms" else v
    tot += v
    name = re.sub(r"^void ", "", x["Kernel Name"])
    print(f"{re.sub(r'[<(].*', '', name)[:26]:26s} {v:9.1f} us  grid {x['Grid Size']:>16s} block {x['Block Size']}")
print(f"total {tot:.1f} us")
