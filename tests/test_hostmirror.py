"""CPU: the product's kernel BODIES (crypto12381_b200/csrc/*.cuh compiled for the host by tests/hostmirror) against
the golden vectors of the reference.  What this cannot cover — the inline-PTX field primitives (emulated
instruction by instruction in tools/gen_fp_ptx.py), the radix-sort kernels and the launch plumbing — is covered by
the `-m gpu` tests through the C ABI."""
import ctypes
import random
import subprocess
import sys
import os

import pytest

import hostmirror_lib as hm
import parity_suite as ps
from conftest import chunks
from oracle import ref

P = 0x1A0111EA397FE69A4B1BA7B6434BACD764774B84F38512BF6730D2A0F6B0F6241EABFFFEB153FFFFB9FEFFFFFFFFAAAB


class MirrorBackend:
    fixed_base1 = staticmethod(hm.g1_fixed_base)
    fixed_base2 = staticmethod(hm.g2_fixed_base)
    mul1 = staticmethod(hm.g1_mul)
    mul2 = staticmethod(hm.g2_mul)
    final_exp = staticmethod(hm.final_exp)
    gt_mul = staticmethod(hm.gt_mul)
    gt_pow = staticmethod(hm.gt_pow)
    gt_pow_gs = staticmethod(hm.gt_pow_gs)

    @staticmethod
    def msm1(points, scalars, c=0):
        n = len(scalars) // 32
        out = ctypes.create_string_buffer(49)
        assert hm.lib().hm_g1_msm_auto(points, scalars, n, c, out) == 0
        return out.raw

    @staticmethod
    def msm2(points, scalars, c=0):
        n = len(scalars) // 32
        out = ctypes.create_string_buffer(97)
        assert hm.lib().hm_g2_msm_auto(points, scalars, n, c, out) == 0    # the device entry's plan: GLS split in four
        plain = hm.g2_msm(points, scalars, c or max(4, hm.lib().hm_choose_window(n)))   # and the unsplit pipeline
        assert plain == out.raw
        return out.raw

    @staticmethod
    def miller(g1, g2, k):
        return hm.pairing_product(g1, g2, k, 0)

    @staticmethod
    def product(g1, g2, k):
        return hm.pairing_product(g1, g2, k, 1)


def test_ptx_generator_emulation():
    """The inline-PTX Fp primitives, instruction list emulated against big-integer arithmetic."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "tools", "gen_fp_ptx.py"), "--check-only"],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr
    assert "emulation OK" in out.stdout


def test_fp_ops():
    rnd = random.Random(7)
    vals = [0, 1, 2, P - 1, P - 2, (P + 1) // 2] + [rnd.randrange(P) for _ in range(40)]
    b48 = lambda v: v.to_bytes(48, "big")
    for a in vals:
        b = rnd.choice(vals)
        assert hm.fp_op(0, b48(a), b48(b)) == b48(a * b % P)
        assert hm.fp_op(1, b48(a), b48(b)) == b48((a + b) % P)
        assert hm.fp_op(2, b48(a), b48(b)) == b48((a - b) % P)
        assert hm.fp_op(4, b48(a)) == b48(-a % P)
        assert hm.fp_op(3, b48(a)) == b48(pow(a, P - 2, P))
        s = int.from_bytes(hm.fp_op(5, b48(a * a % P)), "big")
        assert s in (a, P - a)


def test_points_golden():
    ps.check_points(MirrorBackend)


def test_msm_golden():
    ps.check_msm(MirrorBackend, max_n=300, windows=(0, 2, 5, 8))


def test_msm_1024_auto_window():
    ps.check_msm(MirrorBackend, windows=(0,))


def test_window_choice():
    l = hm.lib()
    assert [l.hm_choose_window(n) for n in (1, 1 << 10, 1 << 16, 1 << 20, 1 << 24)] == sorted(
        l.hm_choose_window(n) for n in (1, 1 << 10, 1 << 16, 1 << 20, 1 << 24))
    assert l.hm_choose_window(1 << 20) == 16 and 4 <= l.hm_choose_window(1) <= 13
    assert l.hm_choose_window_glv(1 << 20) == 16


def test_glv_split():
    """k = +-k0 +- k1 x^2 (mod r), both halves below 2^127 (so 8 signed 16-bit windows never overflow)."""
    l = hm.lib()
    X2 = 0xD201000000010000 ** 2
    rnd = random.Random(11)
    edge = [0, 1, 2, ps.R - 1, ps.R - 2, X2, X2 - 1, X2 + 1, X2 // 2, X2 // 2 + 1, X2 // 2 - 1, X2 * (X2 // 2), X2 * (X2 // 2 + 1),
            X2 * (X2 // 2) + X2 // 2 + 1, X2 * (X2 - 1), (X2 - 2) * X2 + X2 // 2 + 5, 1 << 127, (1 << 128) - 1, 1 << 254]
    for k in edge + [rnd.randrange(ps.R) for _ in range(3000)]:
        k %= ps.R
        a0, a1 = ctypes.create_string_buffer(16), ctypes.create_string_buffer(16)
        sg = (ctypes.c_uint32 * 2)()
        l.hm_glv_split(k.to_bytes(32, "big"), a0, a1, sg)
        k0, k1 = int.from_bytes(a0.raw, "big"), int.from_bytes(a1.raw, "big")
        assert k0 < 1 << 127 and k1 < 1 << 127, hex(k)
        assert ((-k0 if sg[0] else k0) + (-k1 if sg[1] else k1) * X2 - k) % ps.R == 0, hex(k)


def test_gls_split():
    """k = sum_i +-k_i z^i (mod r) with every |k_i| below 2^63 (four signed 16-bit windows never overflow)."""
    l = hm.lib()
    Z = 0xD201000000010000
    rnd = random.Random(12)
    edge = [0, 1, 2, ps.R - 1, ps.R - 2, Z, Z - 1, Z + 1, Z // 2, Z // 2 + 1, Z * Z, Z ** 3, Z ** 3 * (Z // 2 + 1), Z ** 3 * (Z // 2) + Z * Z * (Z // 2 + 1),
            (Z // 2 + 1) * (1 + Z + Z * Z + Z ** 3), (Z - 1) * (1 + Z + Z * Z) , 1 << 63, (1 << 64) - 1, 1 << 254]
    for k in edge + [rnd.randrange(ps.R) for _ in range(3000)]:
        k %= ps.R
        mags = ctypes.create_string_buffer(32)
        sg = (ctypes.c_uint32 * 4)()
        l.hm_gls_split(k.to_bytes(32, "big"), mags, sg)
        ks = [int.from_bytes(mags.raw[8 * q:8 * q + 8], "big") for q in range(4)]
        assert all(m < 1 << 63 for m in ks), hex(k)
        assert (sum((-m if sg[q] else m) * Z ** q for q, m in enumerate(ks)) - k) % ps.R == 0, hex(k)


def test_pairing_golden():
    ps.check_pairing(MirrorBackend)


@pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built")
def test_pod_conversions_against_reference_structs():
    """MIRACL structs (R = 2^406 residues, 58-bit digits, lazily reduced) <-> wire formats."""
    l = hm.lib()
    sc = ref.random_scalars("pod conversion seed", 6)
    a1, a2 = ref.g1_fixed_base_mul(sc), ref.g2_fixed_base_mul(sc)
    a1 += bytes(96)
    a2 += bytes(192)
    for unnorm in (False, True):
        p1, p2 = ref.make_point1(a1, unnorm), ref.make_point2(a2, unnorm)
        w1, w2 = ctypes.create_string_buffer(96 * 7), ctypes.create_string_buffer(192 * 7)
        assert l.hm_pod_points_to_wire(0, p1, 7, w1) == 0 and w1.raw == a1
        assert l.hm_pod_points_to_wire(1, p2, 7, w2) == 0 and w2.raw == a2
    # and back: the structs we write are accepted by the reference and mean the same points
    q1, q2 = ctypes.create_string_buffer(192 * 7), ctypes.create_string_buffer(384 * 7)
    assert l.hm_wire_to_pod_points(0, a1, 7, q1) == 0 and l.hm_wire_to_pod_points(1, a2, 7, q2) == 0
    assert ref.point1_to_c49(q1, 6) == ref.g1_compress(a1[:96 * 6])
    assert ref.point2_to_c97(q2, 6) == ref.g2_compress(a2[:192 * 6])
    bigs = ref.make_big(sc)
    s32 = ctypes.create_string_buffer(32 * 6)
    assert l.hm_pod_bigs_to_scalars(bigs, 6, s32) == 0 and s32.raw == sc
    # fp12: wire -> struct -> reference serialiser, and struct -> wire
    gt = ref.pairing_product_batch(a1[:96 * 2], a2[:192 * 2], 1, 1)
    pods = ctypes.create_string_buffer(776 * 2)
    l.hm_wire_to_pod_fp12(gt, 2, pods)
    assert ref.fp12_to_bytes(pods, 2) == gt
    back = ctypes.create_string_buffer(576 * 2)
    assert l.hm_pod_fp12_to_wire(pods, 2, back) == 0 and back.raw == gt


def test_decompress_golden(golden_points):
    """Wire<F>::decompress (the body of k_decompress) inverts the reference's compressed encodings; bad encodings are refused."""
    l = hm.lib()
    H = bytes.fromhex
    g1c, g2c = H(golden_points["g1_compressed"]), H(golden_points["g2_compressed"])
    o1 = ctypes.create_string_buffer(96 * (len(g1c) // 49 + 1))
    assert l.hm_g1_decompress(g1c + bytes(49), len(g1c) // 49 + 1, o1) == 0
    assert o1.raw == H(golden_points["g1_affine"]) + bytes(96)
    o2 = ctypes.create_string_buffer(192 * (len(g2c) // 97 + 1))
    assert l.hm_g2_decompress(g2c + bytes(97), len(g2c) // 97 + 1, o2) == 0
    assert o2.raw == H(golden_points["g2_affine"]) + bytes(192)
    # unit-tests/g1_point.cpp:132-138 (0xff...), g2_point.cpp:112-118 (0x80 00...)
    bad = ctypes.create_string_buffer(192)
    assert l.hm_g1_decompress(b"\xff" * 49, 1, bad) == -1
    assert l.hm_g2_decompress(b"\x80" + bytes(96), 1, bad) == -1
    # flipping the sign bit gives the negated point
    flipped = bytes([g1c[0] ^ 1]) + g1c[1:49]
    assert l.hm_g1_decompress(flipped, 1, bad) == 0
    assert bad.raw[:48] == o1.raw[:48] and (int.from_bytes(bad.raw[48:96], "big") + int.from_bytes(o1.raw[48:96], "big")) % ps.P == 0


def test_msm_batch_affine_rounds(golden_msm):
    """The batch-affine halving rounds (the slot mapping and the pair bodies of k_ba_fwd / k_ba_bwd, then the accumulation over
    the reduced lists) give the same sums, including on repeated points, P and -P, identities and zero scalars; small windows
    make the bucket lists long enough for several rounds."""
    l = hm.lib()
    H = bytes.fromhex
    for case in golden_msm["cases"]:
        if case["n"] > 300:
            continue
        g1 = case["group"] == "g1"
        pts = (hm.g1_fixed_base if g1 else hm.g2_fixed_base)(H(case["point_scalars"]))
        for rounds in (1, 2, 5):
            for c in (2, 3, 5):
                out = ctypes.create_string_buffer(49 if g1 else 97)
                fn = l.hm_g1_msm_ba if g1 else l.hm_g2_msm_ba
                assert fn(pts, H(case["scalars"]), case["n"], c, rounds, out) == 0
                assert out.raw == H(case["result"]), (case["group"], case["n"], rounds, c)
    for key, want in (("edge_g1", None), ("cancel_g1", bytes(49))):
        e = golden_msm[key]
        n = len(H(e["scalars"])) // 32
        for rounds in (1, 3, 9):
            for c in (2, 4):
                out = ctypes.create_string_buffer(49)
                assert l.hm_g1_msm_ba(H(e["points"]), H(e["scalars"]), n, c, rounds, out) == 0
                assert out.raw == (want if want is not None else H(e["result"])), (key, rounds, c)


def test_msm_upload_groups(golden_msm):
    """Terms cut into upload groups: every (group, window) pair is a sort segment with buckets of its own (msm_list_plan, the
    recode layout and its padding), a group's lists name that group's terms only, and the groups' buckets merge into the same sum -
    also when the term count is not a multiple of the group count, and with fewer terms than groups."""
    l = hm.lib()
    H = bytes.fromhex
    for case in golden_msm["cases"]:
        if case["n"] > 300:
            continue
        g1 = case["group"] == "g1"
        pts = (hm.g1_fixed_base if g1 else hm.g2_fixed_base)(H(case["point_scalars"]))
        for groups, rounds, c in ((2, 0, 4), (2, 2, 3), (3, 1, 5), (4, 3, 2)):
            out = ctypes.create_string_buffer(49 if g1 else 97)
            fn = l.hm_g1_msm_groups if g1 else l.hm_g2_msm_groups
            assert fn(pts, H(case["scalars"]), case["n"], c, rounds, groups, out) == 0, (case["group"], case["n"], groups, rounds, c)
            assert out.raw == H(case["result"]), (case["group"], case["n"], groups, rounds, c)
    e = golden_msm["edge_g1"]
    out = ctypes.create_string_buffer(49)
    assert l.hm_g1_msm_groups(H(e["points"]), H(e["scalars"]), len(H(e["scalars"])) // 32, 4, 2, 3, out) == 0 and out.raw == H(e["result"])


def test_msm_bucket_lists_by_counting(golden_msm):
    """The device's default front end: ranks from per-bucket counters, bounds from one scan, lists filled by (start + rank) -
    msm_recode_each / msm_count_key / msm_entry_term, the bodies k_recode_count and k_bucket_scatter run.  The arrival order of
    the entries (the atomics' order on the device) is scrambled with different strides and must not change a byte; the (term |
    sign) derived from an entry's position must be the one the sorted front end attaches to it."""
    l = hm.lib()
    H = bytes.fromhex
    for case in golden_msm["cases"]:
        if case["n"] > 300:
            continue
        g1 = case["group"] == "g1"
        pts = (hm.g1_fixed_base if g1 else hm.g2_fixed_base)(H(case["point_scalars"]))
        for groups, rounds, c, arrival in ((1, 0, 4, 1), (1, 0, 5, 7), (1, 2, 3, 1001), (2, 1, 4, 13), (3, 3, 2, 5), (4, 0, 6, 3)):
            out = ctypes.create_string_buffer(49 if g1 else 97)
            fn = l.hm_g1_msm_counting if g1 else l.hm_g2_msm_counting
            assert fn(pts, H(case["scalars"]), case["n"], c, rounds, groups, arrival, out) == 0, (case["group"], case["n"], groups, rounds, c, arrival)
            assert out.raw == H(case["result"]), (case["group"], case["n"], groups, rounds, c, arrival)
    for key, want in (("edge_g1", None), ("cancel_g1", bytes(49))):
        e = golden_msm[key]
        n = len(H(e["scalars"])) // 32
        for groups, rounds, arrival in ((1, 0, 1), (1, 3, 11), (3, 2, 7)):
            out = ctypes.create_string_buffer(49)
            assert l.hm_g1_msm_counting(H(e["points"]), H(e["scalars"]), n, 4, rounds, groups, arrival, out) == 0
            assert out.raw == (want if want is not None else H(e["result"])), (key, groups, rounds, arrival)


def test_msm_chunked_accumulation(golden_msm):
    """Bucket lists cut into chunks (k_accumulate over virtual buckets + k_fold), down to chunks of one and three entries."""
    l = hm.lib()
    H = bytes.fromhex
    for case in golden_msm["cases"]:
        if case["group"] != "g1" or case["n"] > 300:
            continue
        pts = hm.g1_fixed_base(H(case["point_scalars"]))
        for c, chunk in ((2, 1), (3, 3), (5, 4), (8, 32)):
            out = ctypes.create_string_buffer(49)
            assert l.hm_g1_msm_chunked(pts, H(case["scalars"]), case["n"], c, chunk, out) == 0
            assert out.raw == H(case["result"]), (case["n"], c, chunk)
    e = golden_msm["edge_g1"]
    out = ctypes.create_string_buffer(49)
    assert l.hm_g1_msm_chunked(H(e["points"]), H(e["scalars"]), len(H(e["scalars"])) // 32, 3, 2, out) == 0 and out.raw == H(e["result"])
    # window widths whose top window would be nearly empty are never chosen (G1 pieces are 127 bits: 14, 9, 7, 6, 5 are out)
    assert {l.hm_choose_window_glv(1 << k) for k in range(4, 25)} <= {4, 8, 10, 11, 12, 13, 15, 16}


def _curve_points_outside_the_subgroups():
    """On-curve points with small x: with cofactors of ~2^126 (G1) and ~2^381 (G2) they are not in the r-torsion subgroups."""
    from oracle import bls12381_oracle as o
    g1, g2 = [], []
    x = 1
    while len(g1) < 3:
        y = o.fp_sqrt((x * x * x + 4) % o.P)
        if y is not None:
            assert o.g1_add(o.g1_mul((x, y), o.R - 1), (x, y)) is not None     # [r]P != O
            g1.append(o.g1_to_affine_bytes((x, y)))
        x += 1
    a = 1
    while len(g2) < 2:
        xx = (a, 1)
        y = o.f2_sqrt(o.f2_add(o.f2_mul(o.f2_sqr(xx), xx), o.G2_B))
        if y is not None:
            g2.append(o.g2_to_affine_bytes((xx, y)))
        a += 1
    return b"".join(g1), b"".join(g2)


def test_subgroup_membership(golden_points):
    """subgroup_member (the body of k_subgroup_check) against MIRACL's PAIR_G1member / PAIR_G2member."""
    l = hm.lib()
    H = bytes.fromhex
    out1, out2 = g1out, g2out = _curve_points_outside_the_subgroups()
    p1 = H(golden_points["g1_affine"])[:96 * 4] + out1 + bytes(96)
    p2 = H(golden_points["g2_affine"])[:192 * 3] + out2 + bytes(192)
    v1, v2 = ctypes.create_string_buffer(len(p1) // 96), ctypes.create_string_buffer(len(p2) // 192)
    assert l.hm_g1_member(p1, len(p1) // 96, v1) == 0 and l.hm_g2_member(p2, len(p2) // 192, v2) == 0
    assert v1.raw == bytes([1] * 4 + [0] * 3 + [0]) and v2.raw == bytes([1] * 3 + [0] * 2 + [0])
    from oracle import ref
    if ref.available():
        assert ref.g1_member(p1[:-96]) == v1.raw[:-1] and ref.g2_member(p2[:-192]) == v2.raw[:-1]


def test_sha3_512_and_hash_to_zp():
    """The bodies of k_sha3_512 against hashlib and the reference's only hashing vector (unit-tests/miracl_core_interface.cpp:10-33)."""
    import hashlib
    l = hm.lib()
    rnd = random.Random(13)
    empty = ctypes.create_string_buffer(64)
    l.hm_sha3_512(b"", 0, empty)
    assert empty.raw.hex().startswith("a69f73cca23a9ac5c8b567dc185a756e97c982164fe25859e0d1dcc1475c80a6")
    for n in (0, 1, 31, 71, 72, 73, 143, 144, 145, 576, 1000):
        msg = bytes(rnd.randrange(256) for _ in range(n))
        d, z = ctypes.create_string_buffer(64), ctypes.create_string_buffer(32)
        l.hm_sha3_512(msg, n, d)
        l.hm_hash_to_zp(msg, n, z)
        want = hashlib.sha3_512(msg).digest()
        assert d.raw == want, n
        assert z.raw == (int.from_bytes(want, "big") % ps.R).to_bytes(32, "big"), n


def test_hash_to_g1(golden_hashing):
    """The bodies of k_hash_to_g1 (digest mod p, SSWU with one shared exponentiation, 11-isogeny, cofactor) against vectors from the
    compiled reference (G1Point::from_hash, g1_point.hpp:219-234), the exceptional inputs included."""
    H = bytes.fromhex
    l = hm.lib()
    g = golden_hashing
    for m, want in zip(g["messages"], g["points"]):
        msg, out = H(m), ctypes.create_string_buffer(49)
        l.hm_hash_to_g1(msg, len(msg), out)
        assert out.raw == H(want)
    for u, want in zip(chunks(H(g["elements"]), 48), chunks(H(g["mapped"]), 49)):
        out = ctypes.create_string_buffer(49)
        assert l.hm_map_to_g1(u, out) == 0
        assert out.raw == want
    out = ctypes.create_string_buffer(49)
    assert l.hm_map_to_g1(ps.P.to_bytes(48, "big"), out) == 1     # u >= p is refused
