"""GPU box: forced halving-round counts around the automatic threshold.  python tools/gpu/rounds_sweep.py G1 16,17,18,19"""
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
g1 = sys.argv[1] == "G1"
fb, msm, psz = (dv.g1_fixed_base_mul_batch, dv.g1_msm, 96) if g1 else (dv.g2_fixed_base_mul_batch, dv.g2_msm, 192)
lns = [int(x) for x in sys.argv[2].split(",")]
n = 1 << max(lns)
P, S = fb(rs(n, 1)), rs(n, 2)
for ln in lns:
    m = 1 << ln
    want = None
    for rounds in (-1, 0, 1, 2, 3, 4, 5):
        lib.c12381_set_msm_batch_affine(rounds)
        for _ in range(2): out = msm(P[:psz*m], S[:32*m])
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(8): out = msm(P[:psz*m], S[:32*m])
        e1.record(); torch.cuda.synchronize()
        st = dv.last_msm_stats()
        r = bytes(out.cpu().numpy()); want = want or r
        print(f"{sys.argv[1]} n=2^{ln} rounds={'auto' if rounds < 0 else rounds} (c={st['window_bits']}): {e0.elapsed_time(e1)/8:.3f} ms  accumulate {st['phases_ms']['accumulate']:.3f} {'OK' if r == want else 'DIFFERS'}", flush=True)
lib.c12381_set_msm_batch_affine(-1)
