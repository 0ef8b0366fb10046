"""GPU box: the host-pointer MSM entry (uploads inside the timed call) by number of upload groups, against the device entry.
    python tools/gpu/e2e_probe.py [G1|G2] [log2 n]"""
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
name = sys.argv[1] if len(sys.argv) > 1 else "G1"
n = 1 << (int(sys.argv[2]) if len(sys.argv) > 2 else 20)
g1 = name == "G1"
h_s = rs(n, 2).pin_memory()
d_p = (dv.g1_fixed_base_mul_batch if g1 else dv.g2_fixed_base_mul_batch)(rs(n, 1).to(dev))
h_p = d_p.cpu().pin_memory()
h_out = torch.empty(49 if g1 else 97, dtype=torch.uint8).pin_memory()
host = lib.c12381_g1_msm if g1 else lib.c12381_g2_msm
results = set()
for groups in (1, 2, 4, 5, 6, 8, 4):
    lib.c12381_set_knob(4, groups)
    best = 1e9
    for it in range(6):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        _lib.check(host(h_p.data_ptr(), h_s.data_ptr(), n, h_out.data_ptr()))
        dt = (time.perf_counter() - t0) * 1e3
        if it: best = min(best, dt)
    results.add(bytes(h_out.numpy()))
    st = dv.last_msm_stats()
    print(f"{name} n={n} host entry, upload groups {groups}: best {best:.3f} ms; device span {st['total_ms']:.3f}; phases", {k: round(v, 3) for k, v in st["phases_ms"].items()}, flush=True)
d_s = h_s.to(dev)
for it in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    r = (dv.g1_msm if g1 else dv.g2_msm)(d_p, d_s); torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) * 1e3
    st = dv.last_msm_stats()
    print(f"device entry {dt:.3f} ms; phases", {k: round(v, 3) for k, v in st["phases_ms"].items()}, flush=True)
results.add(bytes(r.cpu().numpy()))
assert len(results) == 1, "results differ"
print("all results equal")
