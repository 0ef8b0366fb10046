"""Knob sweep of the batch-affine halving rounds on one GPU (same results required): python tools/gpu/ba_tune.py [G1:20,...]"""
import itertools, os, sys
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
cases = (("G1", 20),)
if len(sys.argv) > 1:
    cases = tuple((a.split(":")[0], int(a.split(":")[1])) for a in sys.argv[1].split(","))
def run(msm, p, s, reps=5):
    out = msm(p, s); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): out = msm(p, s)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, bytes(out.cpu().numpy())
for name, logn in cases:
    fb, msm = (dv.g1_fixed_base_mul_batch, dv.g1_msm) if name == "G1" else (dv.g2_fixed_base_mul_batch, dv.g2_msm)
    n = 1 << logn
    p, s = fb(rs(n, 1)), rs(n, 2)
    L.c12381_set_msm_batch_affine(0)
    base_ms, want = run(msm, p, s)
    print(f"{name} n=2^{logn} XYZZ only: {base_ms:.3f} ms  acc {dv.last_msm_stats()['phases_ms']['accumulate']:.3f}", flush=True)
    L.c12381_set_msm_batch_affine(-1)
    for pipes, waves, jmax, tail in itertools.product((2, 3, 1), (1, 2), (32, 48), (3, 2, 4)):
        L.c12381_set_msm_pipelines(pipes); L.c12381_set_knob(0, waves); L.c12381_set_knob(1, jmax); L.c12381_set_knob(2, tail)
        ms, got = run(msm, p, s)
        ph = dv.last_msm_stats()["phases_ms"]
        print(f"{name} n=2^{logn} pipes={pipes} waves={waves} jmax={jmax} tail={tail}: {ms:.3f} ms  acc {ph['accumulate']:.3f} order {ph['bounds_order']:.3f} {'OK' if got == want else 'RESULT DIFFERS'}", flush=True)
L.c12381_set_msm_pipelines(2); L.c12381_set_knob(0, 1); L.c12381_set_knob(1, 32); L.c12381_set_knob(2, 3)
