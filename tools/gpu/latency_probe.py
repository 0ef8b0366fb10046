"""GPU box: latency of the batch-of-1 drop-ins through the host-pointer C ABI (what one call of the reference's bridge function
costs when it lands on the GPU).  The compiled reference on ONE host core, from the bench's cpu_baseline legs: multiply(point1&)
0.69 ms, a pairing 1.6 ms, sum_of_products 66 us per term."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from crypto12381_b200 import _lib, bridge
_lib.init(0)
def rs(n, seed):
    rng = np.random.default_rng(seed)
    a = rng.integers(0, 256, size=(n, 32), dtype=np.uint8); a[:, 0] = rng.integers(0, 0x73, size=n, dtype=np.uint8)
    return a.tobytes()
def med(fn, reps=30):
    for _ in range(3): fn()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter(); fn(); ts.append((time.perf_counter() - t0) * 1e3)
    ts.sort()
    return ts[len(ts) // 2], ts[0]
P1 = bridge.generator_power(rs(4, 1)); P2 = bridge.generator_power2(rs(4, 2)); S = rs(4, 3)
rows = []
rows.append(("multiply(point1&, big)            [c12381_g1_mul_batch, n = 1]", lambda: bridge.multiply(P1[:96], S[:32])))
rows.append(("double_multiply(point1 x2, big x2) [c12381_g1_msm, n = 2]", lambda: bridge.double_multiply(P1[:96], P1[96:192], S[:32], S[32:64])))
rows.append(("generator_power g^x                [c12381_g1_fixed_base_mul_batch, n = 1]", lambda: bridge.generator_power(S[:32])))
rows.append(("multiply(point2&, big)            [c12381_g2_mul_batch, n = 1]", lambda: bridge.multiply2(P2[:192], S[:32])))
rows.append(("pair_ate                           [c12381_miller_batch, 1 x 1]", lambda: bridge.pair_ate(P2[:192], P1[:96])))
rows.append(("pair_double_ate                    [c12381_miller_batch, 1 x 2]", lambda: bridge.pair_double_ate(P2[:192], P1[:96], P2[192:384], P1[96:192])))
f = bridge.pair_ate(P2[:192], P1[:96])
rows.append(("pair_final_exponentiation          [c12381_final_exp_batch, 1]", lambda: bridge.pair_final_exponentiation(f)))
for n in (2, 16, 128, 1024):
    Pn = bridge.generator_power(rs(n, 10 + n)); Sn = rs(n, 20 + n)
    rows.append((f"sum_of_products n = {n:<5}             [c12381_g1_msm]", (lambda Pn=Pn, Sn=Sn: bridge.sum_of_products(Pn, Sn))))
print("GPU (host-pointer entry, pageable host buffers, one call at a time): median / best ms")
for name, fn in rows:
    m, b = med(fn)
    print(f"  {name}: {m:.3f} / {b:.3f}", flush=True)
