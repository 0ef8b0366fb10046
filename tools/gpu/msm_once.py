"""One warmed-up MSM call for a profiler: python tools/gpu/msm_once.py G1 20 [rounds] [pipes] [repeats]"""
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
name, logn = sys.argv[1], int(sys.argv[2])
rounds = int(sys.argv[3]) if len(sys.argv) > 3 else -1
pipes = int(sys.argv[4]) if len(sys.argv) > 4 else 2
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 2
fb, msm = (dv.g1_fixed_base_mul_batch, dv.g1_msm) if name == "G1" else (dv.g2_fixed_base_mul_batch, dv.g2_msm)
n = 1 << logn
p, s = fb(rs(n, 1)), rs(n, 2)
_lib.lib().c12381_set_msm_batch_affine(rounds)
_lib.lib().c12381_set_msm_pipelines(pipes)
for _ in range(reps):
    out = msm(p, s)
torch.cuda.synchronize()
print(bytes(out.cpu().numpy()).hex(), dv.last_msm_stats())
