"""GPU box: per-phase times of the G1 MSM across the small end of the sweep."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from crypto12381_b200 import _lib, device as dv
_lib.init(0)
dev = torch.device("cuda", 0)
def rs(n, seed):
    rng = np.random.default_rng(seed)
    a = rng.integers(0, 256, size=(n, 32), dtype=np.uint8); a[:, 0] = rng.integers(0, 0x73, size=n, dtype=np.uint8)
    return torch.from_numpy(a).reshape(-1).to(dev)
n = 1 << 18
P, S = dv.g1_fixed_base_mul_batch(rs(n, 1)), rs(n, 2)
for ln in (4, 6, 8, 10, 12, 13, 14, 15, 16, 17, 18):
    m = 1 << ln
    for c in ([0] if ln not in (14,) else [0, 8, 10, 12, 13, 14, 16]):
        _lib.lib().c12381_set_msm_window(c)
        dv.g1_msm(P[:96*m], S[:32*m]); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): dv.g1_msm(P[:96*m], S[:32*m])
        e1.record(); torch.cuda.synchronize()
        st = dv.last_msm_stats()
        print(f"n=2^{ln} c={st['window_bits']}: {e0.elapsed_time(e1)/5:.3f} ms", {k: round(v, 3) for k, v in st["phases_ms"].items()}, flush=True)
_lib.lib().c12381_set_msm_window(0)
