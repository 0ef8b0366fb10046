"""GPU box: bucket lists by counting (knob 5 = 0) against the segmented radix sort (knob 5 = 1): same results, per-phase times.
python tools/gpu/front_end_ab.py G1:20,G1:16,G2:18"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from crypto12381_b200 import _lib, device as dv
_lib.init(0)
lib = _lib.lib()
dev = torch.device("cuda", 0)
def rs(n, seed, skew=None):
    rng = np.random.default_rng(seed)
    a = rng.integers(0, 256, size=(n, 32), dtype=np.uint8); a[:, 0] = rng.integers(0, 0x73, size=n, dtype=np.uint8)
    if skew == "equal": a[:] = a[0]
    if skew == "small": a[:, :30] = 0
    return torch.from_numpy(a).reshape(-1).to(dev)
bad = 0
for a in sys.argv[1].split(","):
    name, logn = a.split(":")[0], int(a.split(":")[1])
    fb, msm = (dv.g1_fixed_base_mul_batch, dv.g1_msm) if name == "G1" else (dv.g2_fixed_base_mul_batch, dv.g2_msm)
    n = (1 << logn) - (3 if logn > 4 else 0)
    p = fb(rs(n, 1))
    for skew in (None, "equal", "small"):
        s = rs(n, 2, skew)
        res = {}
        for fe in (1, 2, 0):          # 1: radix sort; 2: counting, points parsed in line; 0: counting, points parsed on a side stream (default)
            lib.c12381_set_knob(5, 1 if fe == 1 else 0)
            lib.c12381_set_knob(6, 0 if fe == 2 else 1)
            for _ in range(3): out = msm(p, s)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10): out = msm(p, s)
            e1.record(); torch.cuda.synchronize()
            ph = dv.last_msm_stats()["phases_ms"]
            res[fe] = bytes(out.cpu().numpy())
            print(f"{name} n={n} scalars={skew or 'random'} front_end={('count+aside', 'sort', 'count')[fe]}: {e0.elapsed_time(e1)/10:.3f} ms  " +
                  " ".join(f"{k}={v:.3f}" for k, v in ph.items()), flush=True)
        if res[0] != res[1] or res[0] != res[2]:
            bad += 1
            print("  RESULTS DIFFER", flush=True)
lib.c12381_set_knob(5, 0); lib.c12381_set_knob(6, 1)
print("mismatches:", bad)
sys.exit(1 if bad else 0)
