"""GPU box: how many warps a lane's halving round is sized for (knob 8, percent of 148 x 12 resident warps) x largest J (knob 1).
python tools/gpu/fill_sweep.py G1:20,G2:18"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from crypto12381_b200 import _lib, device as dv
_lib.init(0)
lib = _lib.lib()
dev = torch.device("cuda", 0)
def rs(n, seed):
    rng = np.random.default_rng(seed)
    a = rng.integers(0, 256, size=(n, 32), dtype=np.uint8); a[:, 0] = rng.integers(0, 0x73, size=n, dtype=np.uint8)
    return torch.from_numpy(a).reshape(-1).to(dev)
for a in sys.argv[1].split(","):
    name, logn = a.split(":")[0], int(a.split(":")[1])
    fb, msm = (dv.g1_fixed_base_mul_batch, dv.g1_msm) if name == "G1" else (dv.g2_fixed_base_mul_batch, dv.g2_msm)
    n = 1 << logn
    p, s = fb(rs(n, 1)), rs(n, 2)
    want = None
    for jmax in (32,):
        for pct in (100, 125, 80, 100):
            lib.c12381_set_knob(1, jmax); lib.c12381_set_knob(8, pct)
            for _ in range(3): out = msm(p, s)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10): out = msm(p, s)
            e1.record(); torch.cuda.synchronize()
            ph = dv.last_msm_stats()["phases_ms"]
            r = bytes(out.cpu().numpy())
            want = want or r
            print(f"{name} n=2^{logn} jmax={jmax} fill={pct}%: {e0.elapsed_time(e1)/10:.3f} ms  accumulate {ph['accumulate']:.3f} {'OK' if r == want else 'DIFFERS'}", flush=True)
lib.c12381_set_knob(1, 32); lib.c12381_set_knob(8, 100)
