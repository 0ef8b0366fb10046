"""Bucket-reduction tail by segment length (knob 3 = threads the segment running sums should fill): python tools/gpu/tail_tune.py [G1:20,...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from crypto12381_b200 import _lib, device as dv
_lib.init(0)
L = _lib.lib()
dev = torch.device("cuda", 0)
def rs(n, seed):
    rng = np.random.default_rng(seed)
    a = rng.integers(0, 256, size=(n, 32), dtype=np.uint8); a[:, 0] = rng.integers(0, 0x73, size=n, dtype=np.uint8)
    return torch.from_numpy(a).reshape(-1).to(dev)
cases = (("G1", 20), ("G2", 18), ("G1", 14), ("G1", 10))
if len(sys.argv) > 1:
    cases = tuple((a.split(":")[0], int(a.split(":")[1])) for a in sys.argv[1].split(","))
for name, logn in cases:
    fb, msm = (dv.g1_fixed_base_mul_batch, dv.g1_msm) if name == "G1" else (dv.g2_fixed_base_mul_batch, dv.g2_msm)
    n = 1 << logn
    p, s = fb(rs(n, 1)), rs(n, 2)
    res = set()
    for wave in (0, 16384, 32768, 65536, 131072, 262144):
        L.c12381_set_knob(3, wave)
        out = msm(p, s); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): out = msm(p, s)
        e1.record(); torch.cuda.synchronize()
        ph = dv.last_msm_stats()["phases_ms"]
        res.add(bytes(out.cpu().numpy()))
        print(f"{name} n=2^{logn} seg-wave={wave}: {e0.elapsed_time(e1)/5:.3f} ms  segment sums {ph['reduce1']:.3f} planes {ph['reduce2']:.3f} finish {ph['finish']:.3f} tail {ph['reduce1']+ph['reduce2']+ph['finish']:.3f}", flush=True)
    assert len(res) == 1, "results differ"
L.c12381_set_knob(3, 0)
