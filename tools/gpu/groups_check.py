"""GPU box: the host-pointer MSM entries give the same result for every upload-group count, repeatedly (races show up as flakes)."""
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
    return torch.from_numpy(a).reshape(-1)
bad = 0
for name, logn in (("G1", 20), ("G1", 19), ("G2", 18), ("G1", 21)):
    g1 = name == "G1"
    n = (1 << logn) - 3
    h_s = rs(n, 2).pin_memory()
    d_p = (dv.g1_fixed_base_mul_batch if g1 else dv.g2_fixed_base_mul_batch)(rs(n, 1).to(dev))
    h_p = d_p.cpu().pin_memory()
    h_out = torch.empty(49 if g1 else 97, dtype=torch.uint8).pin_memory()
    host = lib.c12381_g1_msm if g1 else lib.c12381_g2_msm
    want = bytes((dv.g1_msm if g1 else dv.g2_msm)(d_p, h_s.to(dev)).cpu().numpy())
    lib.c12381_set_msm_batch_affine(0)
    ref = bytes((dv.g1_msm if g1 else dv.g2_msm)(d_p, h_s.to(dev)).cpu().numpy())
    lib.c12381_set_msm_batch_affine(-1)
    assert ref == want, "device entry: halving rounds differ from XYZZ only"
    for groups in (1, 2, 3, 4, 6, 8):
        lib.c12381_set_knob(4, groups)
        for rep in range(6):
            _lib.check(host(h_p.data_ptr(), h_s.data_ptr(), n, h_out.data_ptr()))
            ok = bytes(h_out.numpy()) == want
            bad += not ok
            if not ok: print(f"{name} n={n} groups={groups} rep={rep}: DIFFERS", flush=True)
    print(f"{name} n={n}: checked", flush=True)
lib.c12381_set_knob(4, 4)
print("mismatches:", bad)
sys.exit(1 if bad else 0)
