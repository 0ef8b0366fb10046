"""Pairing-only benchmark (GPU box): both kernel families over a range of batch sizes.
python tools/_pairing_bench.py [k]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from crypto12381_b200 import _lib, device as dv
k = int(sys.argv[1]) if len(sys.argv) > 1 else 4
sizes = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [1, 64, 1024, 4096, 9472, 16384, 65536]
_lib.init(0)
dev = torch.device("cuda", 0)
def rs(n, seed):
    rng = np.random.default_rng(seed)
    a = rng.integers(0, 256, size=(n, 32), dtype=np.uint8); a[:, 0] = rng.integers(0, 0x73, size=n, dtype=np.uint8)
    return torch.from_numpy(a).reshape(-1).to(dev)
Bmax = max(sizes)
g1a, g2a = dv.g1_fixed_base_mul_batch(rs(Bmax * k, 1)), dv.g2_fixed_base_mul_batch(rs(Bmax * k, 2))
for B in sizes:
    g1, g2 = g1a[:B * k * 96], g2a[:B * k * 192]
    outs = {}
    for mode, name in ((1, "thread-per-instance"), (2, "cooperative"))[:1 if os.environ.get("PAIRING_TPI_ONLY") else 2]:
        _lib.lib().c12381_set_pairing_kernel(mode)
        gt = torch.empty(B * 576, dtype=torch.uint8, device=dev)
        dv.pairing_product_batch(g1, g2, k, gt); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); dv.pairing_product_batch(g1, g2, k, gt); e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        outs[mode] = bytes(gt.cpu().numpy())
        print(f"B={B:6d} k={k} {name:20s}: {ms:9.3f} ms  {B*k/ms/1e3:8.4f} M pairings/s", flush=True)
    assert os.environ.get("C12381_LIB_VARIANT") or len(outs) < 2 or outs[1] == outs[2], "kernel families disagree"
