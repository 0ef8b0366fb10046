"""MSM-only A/B on the GPU box: batch-affine halving rounds (0 = XYZZ only, -1 = automatic, forced counts) x pipelines."""
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
cases = (("G1", 20), ("G2", 18), ("G1", 18), ("G1", 16), ("G1", 22), ("G2", 20))
if len(sys.argv) > 1:
    cases = tuple((a.split(":")[0], int(a.split(":")[1])) for a in sys.argv[1].split(","))
combos = ((0, 1), (-1, 1), (-1, 2), (-1, 4), (4, 2), (5, 2), (6, 2), (8, 2))
for name, logn in cases:
    fb, msm = (dv.g1_fixed_base_mul_batch, dv.g1_msm) if name == "G1" else (dv.g2_fixed_base_mul_batch, dv.g2_msm)
    n = 1 << logn
    p, s = fb(rs(n, 1)), rs(n, 2)
    res = {}
    for r, pipes in combos:
        _lib.lib().c12381_set_msm_batch_affine(r)
        _lib.lib().c12381_set_msm_pipelines(pipes)
        out = msm(p, s); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): out = msm(p, s)
        e1.record(); torch.cuda.synchronize()
        st = dv.last_msm_stats()
        res[(r, pipes)] = bytes(out.cpu().numpy())
        print(f"{name} n=2^{logn} rounds={r} pipes={pipes}: {e0.elapsed_time(e1)/5:.3f} ms  accumulate-phase {st['phases_ms']['accumulate']:.3f} ms  phases {dict((k, round(v,3)) for k,v in st['phases_ms'].items())}", flush=True)
    assert len(set(res.values())) == 1, "results differ"
_lib.lib().c12381_set_msm_batch_affine(-1)
_lib.lib().c12381_set_msm_pipelines(2)
