"""GPU box: the G1 (and G2) MSM at the small end of the sweep under every forced window width - what the window choice should
pick.  python tools/gpu/window_sweep.py [G1|G2] [log sizes] [widths]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from crypto12381_b200 import _lib, device as dv
_lib.init(0)
dev = torch.device("cuda", 0)
g1 = (sys.argv[1] if len(sys.argv) > 1 else "G1") == "G1"
def rs(n, seed):
    rng = np.random.default_rng(seed)
    a = rng.integers(0, 256, size=(n, 32), dtype=np.uint8); a[:, 0] = rng.integers(0, 0x73, size=n, dtype=np.uint8)
    return torch.from_numpy(a).reshape(-1).to(dev)
n = 1 << 18
fb, msm, psz = (dv.g1_fixed_base_mul_batch, dv.g1_msm, 96) if g1 else (dv.g2_fixed_base_mul_batch, dv.g2_msm, 192)
P, S = fb(rs(n, 1)), rs(n, 2)
LNS = [int(x) for x in sys.argv[2].split(',')] if len(sys.argv) > 2 else [1, 4, 6, 8, 10, 12, 14, 16, 18]
CS = [int(x) for x in sys.argv[3].split(',')] if len(sys.argv) > 3 else list(range(4, 17))
for ln in LNS:
    m = 1 << ln
    best = None
    for c in [0] + CS:
        _lib.lib().c12381_set_msm_window(c)
        r = msm(P[:psz*m], S[:32*m]); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): msm(P[:psz*m], S[:32*m])
        e1.record(); torch.cuda.synchronize()
        st = dv.last_msm_stats()
        t = e0.elapsed_time(e1) / 5
        ph = st["phases_ms"]
        tag = "auto" if c == 0 else f"c={c}"
        if c and (best is None or t < best[0]): best = (t, c)
        print(f"{'G1' if g1 else 'G2'} n=2^{ln} {tag} (c={st['window_bits']}): {t:.3f} ms  front {ph['recode']+ph['sort']+ph['bounds_order']+ph['parse']:.3f} acc {ph['accumulate']:.3f} "
              f"r1 {ph['reduce1']:.3f} r2 {ph['reduce2']:.3f} fin {ph['finish']:.3f}", flush=True)
    print(f"  -> best c={best[1]} {best[0]:.3f} ms", flush=True)
_lib.lib().c12381_set_msm_window(0)
