"""GPU box: per-phase times of bbs_plus.verify_batch_device at 2^16 signatures x 10 blocks (BASELINE configs[4]).
python tools/gpu/bbs_probe.py [log_b]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from crypto12381_b200 import _lib, bbs_plus, device as dv
_lib.init(0)
dev = torch.device("cuda", 0)
Bs, nmsg = 1 << (int(sys.argv[1]) if len(sys.argv) > 1 else 16), 10


def rs(n, seed):
    rng = np.random.default_rng(seed)
    a = rng.integers(0, 256, size=(n, 32), dtype=np.uint8); a[:, 0] = rng.integers(0, 0x73, size=n, dtype=np.uint8)
    return a


def timed(name, fn, reps=3):
    out = fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): out = fn()
    e1.record(); torch.cuda.synchronize()
    print(f"{name:42s} {e0.elapsed_time(e1) / reps:9.3f} ms", flush=True)
    return out


gens = dv.g1_fixed_base_mul_batch(torch.from_numpy(rs(nmsg + 2, 7100)).reshape(-1).to(dev))
g2b = dv.g2_fixed_base_mul_batch(torch.from_numpy(rs(1, 7200)).reshape(-1).to(dev))
wb = dv.g2_decompress_batch(dv.g2_mul_batch(g2b, torch.from_numpy(rs(1, 7300)).reshape(-1).to(dev)))
bases_g2 = torch.cat((wb, g2b))
neg_g2 = torch.frombuffer(bytearray(bbs_plus._neg_g2(bytes(g2b.cpu().numpy()))), dtype=torch.uint8).to(dev)
xs, rr = rs(Bs, 7400), rs(Bs, 7500)
ms = np.random.default_rng(7).integers(0, 256, size=(Bs, nmsg, 32), dtype=np.uint8); ms[:, :, 0] = 1
one = np.zeros((Bs, 1, 32), dtype=np.uint8); one[:, :, 31] = 1
sc_g1 = torch.from_numpy(np.concatenate((one, rr.reshape(Bs, 1, 32), ms), axis=1).reshape(-1)).to(dev)
sc_g2 = torch.from_numpy(np.concatenate((one, xs.reshape(Bs, 1, 32)), axis=1).reshape(-1)).to(dev)
Bp = dv.g1_multi_fixed_base_batch(gens, sc_g1)
sigA = dv.g1_compress_batch(Bp)          # any on-curve points: the verdicts are not the subject here
A = timed("decompress A (g1_decompress_batch)", lambda: dv.g1_decompress_batch(sigA))
W = timed("w + x g2 (g2_multi_fixed_base_batch)", lambda: dv.g2_multi_fixed_base_batch(bases_g2, sc_g2))
Bq = timed("g1 + r h0 + sum m_j h_j (g1_multi_fixed_base)", lambda: dv.g1_multi_fixed_base_batch(gens, sc_g1))
g1s = torch.stack((A.view(Bs, 96), Bq.view(Bs, 96)), dim=1).reshape(-1)
g2s = torch.cat((W.view(Bs, 192), neg_g2.view(1, 192).expand(Bs, 192)), dim=1).reshape(-1)
for mode, name in ((1, "thread-per-instance"), (2, "cooperative"), (0, "automatic")):
    _lib.lib().c12381_set_pairing_kernel(mode)
    timed(f"2-pair pairing check ({name})", lambda: dv.pairing_check_batch(g1s, g2s, 2))
timed("whole verify_batch_device", lambda: bbs_plus.verify_batch_device(gens, bases_g2, neg_g2, sigA, sc_g1, sc_g2))
