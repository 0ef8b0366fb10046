"""Quick pairing-only benchmark (GPU box): python tools/_pairing_bench.py [instances] [k]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from crypto12381_b200 import _lib, device as dv
B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
k = int(sys.argv[2]) if len(sys.argv) > 2 else 4
_lib.init(0)
dev = torch.device("cuda", 0)
def rs(n, seed):
    rng = np.random.default_rng(seed)
    a = rng.integers(0, 256, size=(n, 32), dtype=np.uint8); a[:, 0] = rng.integers(0, 0x73, size=n, dtype=np.uint8)
    return torch.from_numpy(a).reshape(-1).to(dev)
g1, g2 = dv.g1_fixed_base_mul_batch(rs(B * k, 1)), dv.g2_fixed_base_mul_batch(rs(B * k, 2))
gt = torch.empty(B * 576, dtype=torch.uint8, device=dev)
for mode, fn in (("product", dv.pairing_product_batch),):
    fn(g1, g2, k, gt); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(g1, g2, k, gt); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"variant={os.environ.get('C12381_LIB_VARIANT','')} scalar={os.environ.get('C12381_PAIRING','')} B={B} k={k} {mode}: {ms:.2f} ms  {B*k/ms/1e3:.3f} M pairings/s")
import hashlib
print("digest", hashlib.sha256(bytes(gt[:576*64].cpu().numpy())).hexdigest()[:16])
