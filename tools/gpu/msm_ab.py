"""MSM-only A/B on the GPU box: batch-affine rounds 0/1/2 at n = 2^20 (G1) and 2^18 (G2)."""
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
for name, logn, fb, msm in (("G1", 20, dv.g1_fixed_base_mul_batch, dv.g1_msm), ("G2", 18, dv.g2_fixed_base_mul_batch, dv.g2_msm), ("G1", 16, dv.g1_fixed_base_mul_batch, dv.g1_msm), ("G1", 22, dv.g1_fixed_base_mul_batch, dv.g1_msm)):
    n = 1 << logn
    p, s = fb(rs(n, 1)), rs(n, 2)
    res = {}
    for r in (0, 1, 2):
        _lib.lib().c12381_set_msm_batch_affine(r)
        out = msm(p, s); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): out = msm(p, s)
        e1.record(); torch.cuda.synchronize()
        st = dv.last_msm_stats()
        res[r] = bytes(out.cpu().numpy())
        print(f"{name} n=2^{logn} rounds={r}: {e0.elapsed_time(e1)/5:.3f} ms  accumulate-phase {st['phases_ms']['accumulate']:.3f} ms  phases {dict((k, round(v,3)) for k,v in st['phases_ms'].items())}", flush=True)
    assert res[0] == res[1] == res[2], "results differ"
