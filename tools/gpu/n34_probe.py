"""GPU box: throughput of the §8f N3 / N4 entries - hash to G1, and GT exponentiation (plain ladder against the Galbraith-Scott split).
python tools/gpu/n34_probe.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from crypto12381_b200 import _lib, device as dv
_lib.init(0)
dev = torch.device("cuda", 0)
L = _lib.lib()


def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def rs(n, seed):
    rng = np.random.default_rng(seed)
    a = rng.integers(0, 256, size=(n, 32), dtype=np.uint8); a[:, 0] = rng.integers(0, 0x73, size=n, dtype=np.uint8)
    return torch.from_numpy(a).reshape(-1).to(dev)


for logb in (12, 16, 18):
    B, ml = 1 << logb, 64
    msgs = torch.from_numpy(np.random.default_rng(3).integers(0, 256, size=B * ml, dtype=np.uint8)).to(dev)
    out = torch.empty(B * 49, dtype=torch.uint8, device=dev)
    ms = timed(lambda: _lib.check(L.c12381_hash_to_g1_batch_dev(msgs.data_ptr(), ml, B, out.data_ptr(), None)))
    print(f"hash_to_g1 B=2^{logb}: {ms:8.3f} ms  {B / ms / 1e3:8.3f} M points/s", flush=True)
for logb in (10, 14):
    B = 1 << logb
    g1, g2 = dv.g1_fixed_base_mul_batch(rs(B, 1)), dv.g2_fixed_base_mul_batch(rs(B, 2))
    gt = dv.pairing_product_batch(g1, g2, 1)
    e = rs(B, 5)
    o1, o2 = torch.empty_like(gt), torch.empty_like(gt)
    for mode, name in ((1, "thread-per-instance"), (2, "cooperative")):
        L.c12381_set_pairing_kernel(mode)
        ms = timed(lambda: dv.gt_pow_batch(gt, e, o1))
        print(f"gt_pow    B=2^{logb} {name:20s}: {ms:8.3f} ms  {B / ms / 1e3:8.3f} M/s", flush=True)
    L.c12381_set_pairing_kernel(0)
    ms = timed(lambda: dv.gt_pow_gs_batch(gt, e, o2))
    print(f"gt_pow_gs B=2^{logb} {'thread-per-instance':20s}: {ms:8.3f} ms  {B / ms / 1e3:8.3f} M/s", flush=True)
    assert bytes(o1.cpu().numpy()) == bytes(o2.cpu().numpy())
