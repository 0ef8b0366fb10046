"""CPU, world_size 2, gloo: the host logic of the sharded paths (slice bounds, the all-gather of per-rank partials,
merge order, the final gather).  The per-rank compute is injected (the host mirror of the kernel bodies stands in
for the CUDA entries, which need a GPU); the product's own defaults are exercised by the -m gpu tests."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def test_shard_bounds():
    from crypto12381_b200.distributed import shard_bounds
    for n in (0, 1, 7, 8, 1000, 1 << 20):
        for world in (1, 2, 3, 8):
            cuts = [shard_bounds(n, world, r) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == n
            assert all(cuts[i][1] == cuts[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in cuts]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_bounds(4, 2, 2)


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, HERE)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import hostmirror_lib as hm
        from conftest import load_golden
        from crypto12381_b200.distributed import gather_results, shard_bounds, sharded_msm

        case = [c for c in load_golden("msm.json")["cases"] if c["group"] == "g1" and c["n"] == 257][0]
        ks, ss = bytes.fromhex(case["point_scalars"]), bytes.fromhex(case["scalars"])
        lo, hi = shard_bounds(257, world, rank)
        pts = hm.g1_fixed_base(ks[32 * lo:32 * hi])
        t = lambda b: torch.frombuffer(bytearray(b), dtype=torch.uint8)

        def partial_fn(p, s):   # rank-local MSM (stand-in for device.g1_msm_partial; 49-byte encoding here)
            return t(hm.g1_msm(bytes(p.numpy()), bytes(s.numpy()), 6))

        def sum_fn(gathered):   # all ranks: add the partials in rank order
            comps = [bytes(gathered[i * 49:(i + 1) * 49].numpy()) for i in range(gathered.numel() // 49)]
            from oracle import bls12381_oracle as o   # test-only checker arithmetic on 2 points
            acc = None
            for c in comps:
                acc = o.g1_add(acc, o.g1_decompress(c))
            return t(o.g1_compress(acc))

        total = sharded_msm(t(pts), t(ss[32 * lo:32 * hi]), 49, partial_fn, sum_fn)
        ok = bytes(total.numpy()) == bytes.fromhex(case["result"])
        g = gather_results(torch.full((3,), rank, dtype=torch.uint8))
        ok = ok and g.tolist() == [0, 0, 0, 1, 1, 1]
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


def test_sharded_msm_two_ranks_gloo():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
    assert ret[0] is True and ret[1] is True
