"""Time the device-resident MSM of the loaded library build: python tools/gpu/msm_time.py G1:20[,G2:18,...] [label]"""
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
label = sys.argv[2] if len(sys.argv) > 2 else os.environ.get("C12381_LIB_VARIANT", "default") or "default"
for a in sys.argv[1].split(","):
    name, logn = a.split(":")[0], int(a.split(":")[1])
    fb, msm = (dv.g1_fixed_base_mul_batch, dv.g1_msm) if name == "G1" else (dv.g2_fixed_base_mul_batch, dv.g2_msm)
    n = 1 << logn
    p, s = fb(rs(n, 1)), rs(n, 2)
    for _ in range(3): out = msm(p, s)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): out = msm(p, s)
    e1.record(); torch.cuda.synchronize()
    ph = dv.last_msm_stats()["phases_ms"]
    print(f"[{label}] {name} n=2^{logn}: {e0.elapsed_time(e1)/10:.3f} ms  accumulate {ph['accumulate']:.3f}  result {bytes(out.cpu().numpy()).hex()[:16]}", flush=True)
