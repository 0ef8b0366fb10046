"""One warmed-up thread-per-instance pairing-product launch for a profiler: python tools/gpu/pairing_once.py B [k]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from crypto12381_b200 import _lib, device as dv
B = int(sys.argv[1]); k = int(sys.argv[2]) if len(sys.argv) > 2 else 4
_lib.init(0)
dev = torch.device("cuda", 0)
def rs(n, seed):
    rng = np.random.default_rng(seed)
    a = rng.integers(0, 256, size=(n, 32), dtype=np.uint8); a[:, 0] = rng.integers(0, 0x73, size=n, dtype=np.uint8)
    return torch.from_numpy(a).reshape(-1).to(dev)
g1, g2 = dv.g1_fixed_base_mul_batch(rs(B * k, 1)), dv.g2_fixed_base_mul_batch(rs(B * k, 2))
_lib.lib().c12381_set_pairing_kernel(int(os.environ.get("PAIRING_KERNEL", "1")))
gt = torch.empty(B * 576, dtype=torch.uint8, device=dev)
for _ in range(2):
    dv.pairing_product_batch(g1, g2, k, gt)
torch.cuda.synchronize()
print(bytes(gt[:16].cpu().numpy()).hex())
