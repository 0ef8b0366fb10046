"""GPU box: split tail (knob 7) on / off: same results, times.  python tools/gpu/split_tail_ab.py G1:20,G2:18"""
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
for a in sys.argv[1].split(","):
    name, logn = a.split(":")[0], int(a.split(":")[1])
    g1 = name == "G1"
    fb, msm = (dv.g1_fixed_base_mul_batch, dv.g1_msm) if g1 else (dv.g2_fixed_base_mul_batch, dv.g2_msm)
    host = lib.c12381_g1_msm if g1 else lib.c12381_g2_msm
    n = (1 << logn) - 5
    p, s = fb(rs(n, 1).to(dev)), rs(n, 2).to(dev)
    h_p, h_s = p.cpu().pin_memory(), s.cpu().pin_memory()
    h_out = torch.empty(49 if g1 else 97, dtype=torch.uint8).pin_memory()
    res = {}
    for split in (0, 2, 0, 2):
        lib.c12381_set_knob(7, split)
        for _ in range(3): out = msm(p, s)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): out = msm(p, s)
        e1.record(); torch.cuda.synchronize()
        ph = dv.last_msm_stats()["phases_ms"]
        import time
        best = 1e9
        for _ in range(8):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            _lib.check(host(h_p.data_ptr(), h_s.data_ptr(), n, h_out.data_ptr()))
            best = min(best, (time.perf_counter() - t0) * 1e3)
        r = (bytes(out.cpu().numpy()), bytes(h_out.numpy()))
        res.setdefault(split, r)
        if r != res[split] or r[0] != r[1] or r != res[0]:
            bad += 1; print("  RESULTS DIFFER", flush=True)
        print(f"{name} n={n} split_tail={split}: device {e0.elapsed_time(e1)/10:.3f} ms, host entry best {best:.3f} ms  " + " ".join(f"{k}={v:.3f}" for k, v in ph.items()), flush=True)
lib.c12381_set_knob(7, 2)
print("mismatches:", bad)
sys.exit(1 if bad else 0)
