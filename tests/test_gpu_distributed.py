"""GPU, world_size 2, nccl: the product's own sharded MSM (crypto12381_b200.distributed defaults = the CUDA entries)
against the golden vector and against the single-GPU result.  Skipped on a box with fewer than two GPUs."""
import os
import socket
import sys

import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _worker(rank, world, port, ret):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    sys.path.insert(0, HERE)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["LOCAL_RANK"] = str(rank)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from conftest import load_golden
        from crypto12381_b200 import _lib, device as dv
        from crypto12381_b200.distributed import g1_msm_sharded, g2_msm_sharded, gather_results, shard_bounds
        _lib.init(rank)
        dev = torch.device("cuda", rank)
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
