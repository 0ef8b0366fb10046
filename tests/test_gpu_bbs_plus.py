"""GPU: batched wire-format conversion, products over shared bases, and the BBS+ batch verification built from them
(BASELINE configs[4]; reference: examples/bbs-plus/src/bbs+.cpp:38-73), checked against the oracle restatement and the
compiled reference."""
import random

import pytest
import torch

from conftest import load_golden
from oracle import bls12381_oracle as o   # checker only
from oracle import ref

pytestmark = pytest.mark.gpu
H = bytes.fromhex
R = o.R


def be32(v):
    return int(v).to_bytes(32, "big")


@pytest.fixture(scope="module")
def gpu():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from crypto12381_b200 import _lib, bbs_plus, bridge
    _lib.init(0)
    return bridge, bbs_plus


def test_conversions_round_trip_and_golden(gpu):
    bridge, _ = gpu
    g = load_golden("points.json")
    assert bridge.to_bytes(H(g["g1_affine"])) == H(g["g1_compressed"])
    assert bridge.to_bytes2(H(g["g2_affine"])) == H(g["g2_compressed"])
    assert bridge.from_bytes(H(g["g1_compressed"]) + bytes(49)) == H(g["g1_affine"]) + bytes(96)
    assert bridge.from_bytes2(H(g["g2_compressed"]) + bytes(97)) == H(g["g2_affine"]) + bytes(192)
    rnd = random.Random(5)
    ks = b"".join(be32(rnd.randrange(R)) for _ in range(3000))
    p1, p2 = bridge.generator_power(ks), bridge.generator_power2(ks[:32 * 500])
    assert bridge.from_bytes(bridge.to_bytes(p1)) == p1
    assert bridge.from_bytes2(bridge.to_bytes2(p2)) == p2
    # the reference rejects these (unit-tests/g1_point.cpp:132-138, g2_point.cpp:112-118): status, not a result
    from crypto12381_b200._lib import C12381Error, EINPUT
    for bad, fn in ((b"\xff" * 49, bridge.from_bytes), (b"\x80" + bytes(96), bridge.from_bytes2)):
        with pytest.raises(C12381Error) as e:
            fn(bad)
        assert e.value.code == EINPUT


def test_conversions_vs_oracle(gpu):
    bridge, _ = gpu
    rnd = random.Random(6)
    pts = [o.g1_mul(o.G1_GEN, rnd.randrange(1, R)) for _ in range(4)]
    enc = b"".join(o.g1_compress(p) for p in pts)
    assert bridge.from_bytes(enc) == b"".join(o.g1_to_affine_bytes(p) for p in pts)
    q = [o.g2_mul(o.G2_GEN, rnd.randrange(1, R)) for _ in range(3)]
    enc2 = b"".join(o.g2_compress(p) for p in q)
    assert bridge.from_bytes2(enc2) == b"".join(o.g2_to_affine_bytes(p) for p in q)


def test_products_over_shared_bases(gpu):
    bridge, _ = gpu
    rnd = random.Random(7)
    m, B = 12, 257
    bases = bridge.generator_power(b"".join(be32(rnd.randrange(R)) for _ in range(m - 1))) + bytes(96)   # one identity base
    vals = [[rnd.randrange(R) for _ in range(m)] for _ in range(B)]
    vals[0] = [0] * m
    vals[1] = [R - 1] * m
    vals[2] = [1] + [0] * (m - 1)
    flat = b"".join(be32(v) for row in vals for v in row)
    got = bridge.products_over_bases(bases, flat)
    # each instance is an m-term sum of products: the MSM entry computes the same value
    for b in (0, 1, 2, 3, 100, B - 1):
        want = bridge.sum_of_products(bases, b"".join(be32(v) for v in vals[b]))
        assert bridge.to_bytes(got[96 * b:96 * b + 96]) == want
    if ref.available():
        want = b"".join(ref.g1_msm(bases, b"".join(be32(v) for v in vals[b])) for b in range(0, B, 16))
        assert b"".join(bridge.to_bytes(got[96 * b:96 * b + 96]) for b in range(0, B, 16)) == want
    bases2 = bridge.generator_power2(b"".join(be32(rnd.randrange(R)) for _ in range(2)))
    flat2 = b"".join(be32(1) + be32(rnd.randrange(R)) for _ in range(33))
    got2 = bridge.products_over_bases2(bases2, flat2)
    for b in (0, 32):
        assert bridge.to_bytes2(got2[192 * b:192 * b + 192]) == bridge.sum_of_products2(bases2, flat2[64 * b:64 * b + 64])


@pytest.mark.skipif(not ref.available(), reason="oracle/_ref/libref12381.so not present")
def test_bbs_plus_verify_reference_made_signatures(gpu):
    """Signatures made by the REFERENCE's arithmetic (oracle/ref_shim.cpp: bbs+.cpp:38-55 on the unmodified bridge - parse, the live
    double_multiply product loop, multiply by 1 / (gamma + x)) go to the CUDA `verify_batch`; and the reference's own verify
    (bbs+.cpp:57-73) confirms every verdict, valid and tampered - nothing in this test was signed by the GPU."""
    bridge, bbs = gpu
    rnd = random.Random(88)
    n, B, t = 10, 96, ref.hardware_threads()
    gen_k = b"".join(be32(rnd.randrange(1, R)) for _ in range(n + 2))
    gens_a = ref.g1_fixed_base_mul(gen_k, t)                     # g1, h0, h_0 .. h_9: affine, made by the reference
    g2_a = ref.g2_fixed_base_mul(be32(rnd.randrange(1, R)))
    gamma = rnd.randrange(1, R)
    pk = ref.g2_mul_batch(g2_a, be32(gamma))                     # w = g2^gamma, 97 B
    w_a, ok = ref.g2_decompress(pk)
    assert ok
    gens_c, g2_c = ref.g1_compress(gens_a), ref.g2_compress(g2_a)
    pp = bbs.PublicParameters(gens_c[:49] + g2_c + gens_c[49:98], [gens_c[49 * i:49 * i + 49] for i in range(2, n + 2)])
    assert pp.g1 + pp.h0 + pp.h == gens_a and pp.g2 == g2_a
    msgs = [bytes(rnd.randrange(256) for _ in range(31 * n)) for _ in range(B)]      # n full blocks each
    msgs[1] = b"short"                                                              # ... and ragged ones
    msgs[2] = bytes(rnd.randrange(256) for _ in range(31 * 3 + 7))
    xs, rs = [rnd.randrange(R) for _ in range(B)], [rnd.randrange(R) for _ in range(B)]
    blocks = [bbs.encode_to_zp(m) for m in msgs]
    rows = b"".join(be32(1) + be32(r) + b"".join(be32(m) for m in ms) + bytes(32 * (n - len(ms))) for r, ms in zip(rs, blocks))
    xrows = b"".join(be32(1) + be32(x) for x in xs)
    A = ref.bbs_sign_batch(gens_a[:96], gens_a[96:192], gens_a[192:], be32(gamma), rows, n, t, xs=b"".join(be32(x) for x in xs))
    sigs = [A[49 * i:49 * i + 49] + xs[i].to_bytes(48, "big") + rs[i].to_bytes(48, "big") for i in range(B)]
    ref_verify = lambda enc, rows_, xrows_: list(ref.bbs_verify_batch(gens_a[:96], g2_a, gens_a[96:192], gens_a[192:], w_a, n, enc, rows_, xrows_, t))
    assert ref_verify(A, rows, xrows) == [1] * B                 # the reference accepts its own signatures
    assert bbs.verify_batch(pp, pk, msgs, sigs) == [True] * B    # ... and so does the CUDA path
    # the GPU's sign_batch makes the same bytes
    assert bbs.sign_batch(pp, gamma.to_bytes(48, "big"), msgs, xs, rs) == sigs
    # tampered: message changed, x changed, r changed, A swapped, A negated, A = identity
    bad_msgs, bad = list(msgs), [bytearray(s_) for s_ in sigs]
    bad_msgs[4] = bytes([msgs[4][0] ^ 1]) + msgs[4][1:]
    bad[6][49:97] = ((xs[6] + 1) % R).to_bytes(48, "big")
    bad[8][97:145] = ((rs[8] + 1) % R).to_bytes(48, "big")
    bad[10][:49] = sigs[11][:49]
    bad[12][0] ^= 1                                                # the other square root: -A
    bad[14][:49] = bytes(49)
    bad = [bytes(b_) for b_ in bad]
    got = bbs.verify_batch(pp, pk, bad_msgs, bad)
    assert got == [i not in (4, 6, 8, 10, 12, 14) for i in range(B)]
    bad_blocks = [bbs.encode_to_zp(m) for m in bad_msgs]
    bad_rows = b"".join(be32(1) + be32(int.from_bytes(s_[97:145], "big")) + b"".join(be32(m) for m in ms) + bytes(32 * (n - len(ms)))
                        for s_, ms in zip(bad, bad_blocks))
    bad_xrows = b"".join(be32(1) + be32(int.from_bytes(s_[49:97], "big")) for s_ in bad)
    assert ref_verify(b"".join(s_[:49] for s_ in bad), bad_rows, bad_xrows) == [int(v) for v in got]


def test_bbs_plus_sign_verify_batch(gpu):
    bridge, bbs = gpu
    rnd = random.Random(8)
    n = 16
    gens = bridge.to_bytes(bridge.generator_power(b"".join(be32(rnd.randrange(1, R)) for _ in range(n + 2))))
    g2 = bridge.to_bytes2(bridge.generator_power2(be32(rnd.randrange(1, R))))
    pp = bbs.PublicParameters(gens[:49] + g2 + gens[49:98], [gens[49 * i:49 * i + 49] for i in range(2, n + 2)])
    gamma = rnd.randrange(1, R)
    pk = bridge.multiply2(pp.g2, be32(gamma))            # w = g2^γ, 97 B
    sk = gamma.to_bytes(48, "big")
    B = 64
    msgs = [b"Hello, BBS+!"] + [bytes(rnd.randrange(256) for _ in range(rnd.randrange(1, 31 * n))) for _ in range(B - 1)]
    xs, rs = [rnd.randrange(R) for _ in range(B)], [rnd.randrange(R) for _ in range(B)]
    sigs = bbs.sign_batch(pp, sk, msgs, xs, rs)
    assert all(len(s) == 145 for s in sigs)
    assert bbs.verify_batch(pp, pk, msgs, sigs) == [True] * B
    # tampering: wrong message, wrong x, A replaced, A not on the curve, x out of range
    bad_msgs, bad_sigs = list(msgs), list(sigs)
    bad_msgs[3] = msgs[3] + b"!"
    bad_sigs[5] = sigs[5][:49] + ((xs[5] + 1) % R).to_bytes(48, "big") + sigs[5][97:]
    bad_sigs[7] = sigs[8][:49] + sigs[7][49:]
    bad_sigs[9] = b"\xff" * 49 + sigs[9][49:]
    bad_sigs[11] = sigs[11][:49] + R.to_bytes(48, "big") + sigs[11][97:]
    want = [i not in (3, 5, 7, 9, 11) for i in range(B)]
    assert bbs.verify_batch(pp, pk, bad_msgs, bad_sigs) == want
    # one signature against the oracle's restatement of verify (bbs+.cpp:72) with its own pairing
    A = o.g1_decompress(sigs[0][:49])
    W = o.g2_add(o.g2_from_affine_bytes(bridge.from_bytes2(pk)), o.g2_mul(o.g2_from_affine_bytes(pp.g2), xs[0]))
    ms = bbs.encode_to_zp(msgs[0])
    Bp = o.g1_add(o.g1_add(o.g1_from_affine_bytes(pp.g1), o.g1_mul(o.g1_from_affine_bytes(pp.h0), rs[0])),
                  o.g1_mul(o.g1_from_affine_bytes(pp.h[:96]), ms[0]))
    assert len(ms) == 1 and ms[0] == (1 << 248) | int.from_bytes(b"Hello, BBS+!" + bytes(31 - 12), "big")
    assert o.pairing(A, W) == o.pairing(Bp, o.g2_from_affine_bytes(pp.g2))


def test_subgroup_membership_batch(gpu):
    bridge, _ = gpu
    from test_hostmirror import _curve_points_outside_the_subgroups
    out1, out2 = _curve_points_outside_the_subgroups()
    rnd = random.Random(9)
    ks = b"".join(be32(rnd.randrange(1, R)) for _ in range(200))
    p1 = bridge.generator_power(ks) + out1 + bytes(96)
    p2 = bridge.generator_power2(ks[:32 * 50]) + out2 + bytes(192)
    v1, v2 = bridge.is_member(p1), bridge.is_member2(p2)
    assert v1 == bytes([1] * 200 + [0] * 3 + [0]) and v2 == bytes([1] * 50 + [0] * 2 + [0])
    if ref.available():
        assert ref.g1_member(p1[:-96]) == v1[:-1] and ref.g2_member(p2[:-192]) == v2[:-1]


def test_sha3_512_and_hash_to_zp_batch(gpu):
    """Batched Fiat-Shamir hashing against hashlib (FIPS 202) and, for the digest, the reference's own SHA3 through its bridge."""
    import hashlib
    bridge, _ = gpu
    rnd = random.Random(14)
    for L, B in ((0, 3), (49, 1000), (72, 17), (145, 513), (49 + 97 + 576, 64)):
        msgs = bytes(rnd.randrange(256) for _ in range(L * B))
        want = [hashlib.sha3_512(msgs[L * i:L * (i + 1)]).digest() for i in range(B)]
        if L:
            assert bridge.sha3_512(msgs, L) == b"".join(want)
            assert bridge.hash_to_zp(msgs, L) == b"".join((int.from_bytes(d, "big") % R).to_bytes(32, "big") for d in want)
    assert bridge.sha3_512(b"", 0) == b""


def test_hash_to_g1_batch(gpu, golden_hashing):
    """`hash(...) -> G1` in batch (G1Point::from_hash, g1_point.hpp:219-234) against vectors from the compiled reference, the
    reference itself on fresh messages, and - at a size the CPU cannot follow - subgroup membership of every output."""
    bridge, _ = gpu
    g = golden_hashing
    H = bytes.fromhex
    for m, want in zip(g["messages"], g["points"]):
        assert bridge.hash_to_g1(H(m), len(H(m))) == H(want)
    assert bridge.map_to_g1(H(g["elements"])) == H(g["mapped"])
    with pytest.raises(Exception):
        bridge.map_to_g1(o.P.to_bytes(48, "big"))      # u >= p is refused
    rnd = random.Random(15)
    L, B = 97, 300
    msgs = bytes(rnd.randrange(256) for _ in range(L * B))
    got = bridge.hash_to_g1(msgs, L)
    if ref.available():
        assert got == ref.hash_to_g1(msgs, L, B, ref.hardware_threads())
    else:
        from oracle import bls12381_oracle as o
        assert got[:49 * 8] == b"".join(o.g1_compress(o.hash_to_g1(msgs[L * i:L * (i + 1)])) for i in range(8))
    B = 1 << 14
    msgs = bytes(rnd.randrange(256) for _ in range(32 * B))
    enc = bridge.hash_to_g1(msgs, 32)
    pts = bridge.from_bytes(enc)
    assert bridge.is_member(pts) == bytes([1] * B)
    # the device-pointer entries (CUDA tensors in, CUDA tensors out) give the same bytes
    from crypto12381_b200 import device as dv
    t = torch.frombuffer(bytearray(msgs), dtype=torch.uint8).cuda()
    assert bytes(dv.hash_to_g1_batch(t, 32).cpu().numpy()) == enc
    assert bytes(dv.hash_to_zp_batch(t, 32).cpu().numpy()) == bridge.hash_to_zp(msgs, 32)
    assert bytes(dv.sha3_512_batch(t, 32).cpu().numpy()) == bridge.sha3_512(msgs, 32)
    dv.sync_status()
    assert bridge.hash_to_g1(b"", 0) == b""


def test_bbs_plus_aggregate_verification(gpu):
    """The random-linear-combination batch check agrees with the per-signature verdicts: accepts a valid batch, rejects one
    with a single tampered signature (whatever its position)."""
    bridge, bbs = gpu
    rnd = random.Random(10)
    n = 10
    gens = bridge.to_bytes(bridge.generator_power(b"".join(be32(rnd.randrange(1, R)) for _ in range(n + 2))))
    g2 = bridge.to_bytes2(bridge.generator_power2(be32(rnd.randrange(1, R))))
    pp = bbs.PublicParameters(gens[:49] + g2 + gens[49:98], [gens[49 * i:49 * i + 49] for i in range(2, n + 2)])
    gamma = rnd.randrange(1, R)
    pk = bridge.multiply2(pp.g2, be32(gamma))
    B = 300
    msgs = [bytes(rnd.randrange(256) for _ in range(rnd.randrange(1, 31 * n))) for _ in range(B)]
    xs, rs = [rnd.randrange(R) for _ in range(B)], [rnd.randrange(R) for _ in range(B)]
    sigs = bbs.sign_batch(pp, gamma.to_bytes(48, "big"), msgs, xs, rs)
    assert bbs.verify_batch_aggregate(pp, pk, msgs, sigs, b"seed-1") is True
    for pos in (0, 137, B - 1):
        bad = list(sigs)
        bad[pos] = sigs[pos][:49] + ((xs[pos] + 1) % R).to_bytes(48, "big") + sigs[pos][97:]
        assert bbs.verify_batch_aggregate(pp, pk, msgs, bad, b"seed-2") is False
        assert bbs.verify_batch(pp, pk, msgs, bad) == [i != pos for i in range(B)]
    swapped = list(msgs)
    swapped[5], swapped[6] = msgs[6], msgs[5]
    assert bbs.verify_batch_aggregate(pp, pk, swapped, sigs, b"seed-3") is False
