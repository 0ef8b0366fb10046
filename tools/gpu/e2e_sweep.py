"""GPU box: host-pointer G1 MSM entry, upload groups x XYZZ-tail knob: python tools/gpu/e2e_sweep.py [log2 n]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from crypto12381_b200 import _lib, device as dv
_lib.init(0)
lib = _lib.lib()
dev = torch.device("cuda", 0)
def rs(n, seed):
    rng = np.random.default_rng(seed)
    a = rng.integers(0, 256, size=(n, 32), dtype=np.uint8); a[:, 0] = rng.integers(0, 0x73, size=n, dtype=np.uint8)
    return torch.from_numpy(a).reshape(-1)
n = 1 << (int(sys.argv[1]) if len(sys.argv) > 1 else 20)
h_s = rs(n, 2).pin_memory()
h_p = dv.g1_fixed_base_mul_batch(rs(n, 1).to(dev)).cpu().pin_memory()
h_out = torch.empty(49, dtype=torch.uint8).pin_memory()
res = set()
for groups in (1, 2, 3, 4):
    for tail in (3, 2, 4):
        lib.c12381_set_knob(4, groups); lib.c12381_set_knob(2, tail)
        ts = []
        for it in range(8):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            _lib.check(lib.c12381_g1_msm(h_p.data_ptr(), h_s.data_ptr(), n, h_out.data_ptr()))
            ts.append((time.perf_counter() - t0) * 1e3)
        res.add(bytes(h_out.numpy()))
        st = dv.last_msm_stats()
        print(f"n=2^{int(np.log2(n))} groups={groups} tail={tail}: best {min(ts[1:]):.3f} median {sorted(ts[1:])[3]:.3f} ms  rounds {st['ba_rounds']} accumulate-phase {st['phases_ms']['accumulate']:.3f}", flush=True)
assert len(res) == 1
lib.c12381_set_knob(4, 4); lib.c12381_set_knob(2, 3)
