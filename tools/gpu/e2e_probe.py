"""GPU box: phases of the host-pointer G1 MSM entry (two-phase accumulation behind the upload)."""
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
n = 1 << 20
h_s = rs(n, 2).pin_memory()
d_p = dv.g1_fixed_base_mul_batch(rs(n, 1).to(dev))
h_p = d_p.cpu().pin_memory()
h_out = torch.empty(49, dtype=torch.uint8).pin_memory()
for it in range(4):
    t0 = time.perf_counter()
    _lib.check(lib.c12381_g1_msm(h_p.data_ptr(), h_s.data_ptr(), n, h_out.data_ptr()))
    dt = (time.perf_counter() - t0) * 1e3
    st = dv.last_msm_stats()
    print(f"host entry {dt:.3f} ms; device total {st['total_ms']:.3f}; phases", {k: round(v, 3) for k, v in st["phases_ms"].items()}, flush=True)
d_s = h_s.to(dev)
for it in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    r = dv.g1_msm(d_p, d_s); torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) * 1e3
    st = dv.last_msm_stats()
    print(f"device entry {dt:.3f} ms; phases", {k: round(v, 3) for k, v in st["phases_ms"].items()}, flush=True)
assert bytes(r.cpu().numpy()) == bytes(h_out.numpy())
