"""GPU, world_size 2: the product's own sharded MSM (crypto12381_b200.distributed defaults = the CUDA entries) against the
golden vector, the seeds and the single-GPU result.
  * nccl, one rank per GPU - skipped on a box with fewer than two GPUs;
  * gloo, BOTH ranks on cuda:0 - runs on the one-GPU test box: the same CUDA entries (partial -> all-gather -> merge) and the
    host-pointer form, only the 96-byte all-gather goes through gloo (NCCL refuses two ranks on one device)."""
import os
import socket
import sys

import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _worker(rank, world, port, ret, backend="nccl"):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    sys.path.insert(0, HERE)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    gpu = rank if backend == "nccl" else 0
    os.environ["LOCAL_RANK"] = str(gpu)
    torch.cuda.set_device(gpu)
    if backend == "nccl":
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", gpu))
    else:
        dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from conftest import load_golden
        from crypto12381_b200 import _lib, device as dv
        from crypto12381_b200.distributed import (g1_msm_sharded, g1_msm_sharded_host, g2_msm_sharded, g2_msm_sharded_host, gather_results,
                                                  shard_bounds)
        _lib.init(gpu)
        dev = torch.device("cuda", gpu)
        t = lambda b: torch.frombuffer(bytearray(b), dtype=torch.uint8).to(dev)
        ok = True
        for group, n, fixed, sharded, single in (("g1", 1024, dv.g1_fixed_base_mul_batch, g1_msm_sharded, dv.g1_msm),
                                                 ("g2", 33, dv.g2_fixed_base_mul_batch, g2_msm_sharded, dv.g2_msm)):
            case = [c for c in load_golden("msm.json")["cases"] if c["group"] == group and c["n"] == n][0]
            ks, ss = bytes.fromhex(case["point_scalars"]), bytes.fromhex(case["scalars"])
            lo, hi = shard_bounds(n, world, rank)
            pts = fixed(t(ks[32 * lo:32 * hi]))
            total = sharded(pts, t(ss[32 * lo:32 * hi]))
            ok = ok and bytes(total.cpu().numpy()) == bytes.fromhex(case["result"])
            whole = single(fixed(t(ks)), t(ss))       # every rank also computes the unsharded sum on its own GPU
            ok = ok and bytes(whole.cpu().numpy()) == bytes(total.cpu().numpy())
        # a sum the golden vectors do not hold: 2^14 terms per rank, checked against the seeds (sum s_i k_i mod r, then g^total)
        import numpy as np
        R = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001

        def scalars(n, seed):
            a = np.random.default_rng(seed).integers(0, 256, size=(n, 32), dtype=np.uint8)
            a[:, 0] %= 0x73
            return a

        n = 1 << 14
        allk = [scalars(n, 10 + r) for r in range(world)]
        alls = [scalars(n, 20 + r) for r in range(world)]
        tot = sum(int.from_bytes(k.tobytes(), "big") * int.from_bytes(s_.tobytes(), "big") for kk, ss_ in zip(allk, alls) for k, s_ in zip(kk, ss_)) % R
        tot_t = t(tot.to_bytes(32, "big"))
        for fixed, sharded, sharded_host, compress in ((dv.g1_fixed_base_mul_batch, g1_msm_sharded, g1_msm_sharded_host, dv.g1_compress_batch),
                                                       (dv.g2_fixed_base_mul_batch, g2_msm_sharded, g2_msm_sharded_host, dv.g2_compress_batch)):
            d_s = torch.from_numpy(alls[rank]).reshape(-1).to(dev)
            pts = fixed(torch.from_numpy(allk[rank]).reshape(-1).to(dev))
            want = bytes(compress(fixed(tot_t)).cpu().numpy())
            ok = ok and bytes(sharded(pts, d_s).cpu().numpy()) == want
            ok = ok and sharded_host(pts.cpu().pin_memory(), d_s.cpu().pin_memory()) == want       # host buffers in, host bytes out
        g = gather_results(torch.full((3,), rank, dtype=torch.uint8, device=dev))
        ok = ok and g.cpu().tolist() == [0, 0, 0, 1, 1, 1]
        dv.sync_status()
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


@pytest.mark.gpu
def test_sharded_msm_two_ranks_nccl():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
    assert ret[0] is True and ret[1] is True


@pytest.mark.gpu
def test_sharded_msm_two_ranks_one_gpu_gloo():
    if not torch.cuda.is_available():
        pytest.fail("the -m gpu tests need a CUDA device")
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, port, ret, "gloo"), nprocs=2, join=True)
    assert ret[0] is True and ret[1] is True
