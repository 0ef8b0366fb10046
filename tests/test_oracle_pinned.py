"""Pins the Python restatement (oracle/bls12381_oracle.py) to the reference:
 (a) against the committed golden vectors (tests/golden/*.json, generated from the compiled, unmodified
     reference by tools/gen_golden.py), always;
 (b) against oracle/_ref/libref12381.so directly on fresh seeded inputs, when that library is present.
CPU only."""
import pytest

from conftest import chunks
from oracle import bls12381_oracle as o
from oracle import ref

H = bytes.fromhex


def ints(b):
    return [int.from_bytes(c, "big") for c in chunks(b, 32)]


def test_constants():
    assert o.P % 4 == 3 and o.R == o.X_ABS ** 4 - o.X_ABS ** 2 + 1
    assert o.g1_on_curve(o.G1_GEN) and o.g2_on_curve(o.G2_GEN)
    assert o.g1_mul(o.G1_GEN, o.R) is None
    assert pow(o.CRU, 3, o.P) == 1
    # final exponent identity behind PAIR_fexp's hard part
    x = -o.X_ABS
    assert 3 * (o.P ** 4 - o.P ** 2 + 1) // o.R == (x - 1) ** 2 * (x + o.P) * (x * x + o.P * o.P - 1) + 3


def test_points_golden(golden_points):
    g = golden_points
    assert H(g["g1_generator"]) == o.g1_to_affine_bytes(o.G1_GEN)
    assert H(g["g2_generator"]) == o.g2_to_affine_bytes(o.G2_GEN)
    ks = ints(H(g["point_scalars"]))
    ss = ints(H(g["scalars"]))
    g1 = [o.g1_mul(o.G1_GEN, k) for k in ks]
    g2 = [o.g2_mul(o.G2_GEN, k) for k in ks[:8]]
    assert H(g["g1_affine"]) == b"".join(map(o.g1_to_affine_bytes, g1))
    assert H(g["g2_affine"]) == b"".join(map(o.g2_to_affine_bytes, g2))
    assert H(g["g1_compressed"]) == b"".join(map(o.g1_compress, g1))
    assert H(g["g2_compressed"]) == b"".join(map(o.g2_compress, g2))
    assert [o.g1_decompress(c) for c in chunks(H(g["g1_compressed"]), 49)] == g1
    assert [o.g2_decompress(c) for c in chunks(H(g["g2_compressed"]), 97)] == g2
    assert H(g["g1_mul"]) == b"".join(o.g1_compress(o.g1_mul(p, s)) for p, s in zip(g1, ss))
    assert H(g["g2_mul"]) == b"".join(o.g2_compress(o.g2_mul(p, s)) for p, s in zip(g2, ss))
    es = ints(H(g["edge_scalars"]))
    assert H(g["g1_mul_edge"]) == b"".join(o.g1_compress(o.g1_mul(p, s)) for p, s in zip(g1, es))
    assert H(g["g2_mul_edge"]) == b"".join(o.g2_compress(o.g2_mul(p, s)) for p, s in zip(g2, es))


def test_msm_golden(golden_msm):
    for case in golden_msm["cases"]:
        ks = ints(H(case["point_scalars"]))
        ss = ints(H(case["scalars"]))
        if case["n"] > 300:
            continue  # the 1024-term case is checked by the faster paths in test_msm_1024
        if case["group"] == "g1":
            pts = [o.g1_mul(o.G1_GEN, k) for k in ks]
            assert H(case["result"]) == o.g1_compress(o.g1_msm_muln(pts, ss))
            if case["n"] <= 64:
                assert H(case["result"]) == o.g1_compress(o.g1_msm_live(pts, ss))
        else:
            pts = [o.g2_mul(o.G2_GEN, k) for k in ks]
            assert H(case["result"]) == o.g2_compress(o.g2_msm(pts, ss))
    e = golden_msm["edge_g1"]
    pts = [o.g1_from_affine_bytes(c) for c in chunks(H(e["points"]), 96)]
    ss = ints(H(e["scalars"]))
    assert H(e["result"]) == H(e["result_live"]) == o.g1_compress(o.g1_msm_muln(pts, ss))
    c = golden_msm["cancel_g1"]
    pts = [o.g1_from_affine_bytes(x) for x in chunks(H(c["points"]), 96)]
    assert H(c["result"]) == bytes(49) == o.g1_compress(o.g1_msm_muln(pts, ints(H(c["scalars"]))))


def test_msm_1024(golden_msm):
    """BASELINE.json configs[0]: the 1024-term G1 sum.  Σ s_i (k_i G) = (Σ s_i k_i mod r) G."""
    case = [c for c in golden_msm["cases"] if c["n"] == 1024][0]
    ks = ints(H(case["point_scalars"]))
    ss = ints(H(case["scalars"]))
    total = sum(k * s for k, s in zip(ks, ss)) % o.R
    assert H(case["result"]) == o.g1_compress(o.g1_mul(o.G1_GEN, total))


def test_pairing_golden(golden_pairing):
    g = golden_pairing
    g1 = [o.g1_from_affine_bytes(c) for c in chunks(H(g["g1"]), 96)]
    g2 = [o.g2_from_affine_bytes(c) for c in chunks(H(g["g2"]), 192)]
    assert H(g["generator_miller"]) == o.gt_to_bytes(o.miller_loop([(o.G1_GEN, o.G2_GEN)]))
    assert H(g["generator_gt"]) == o.gt_to_bytes(o.pairing(o.G1_GEN, o.G2_GEN))
    assert g["generator_gt"].startswith("0f41e586") and g["generator_gt"].endswith("89b6")
    single = [o.miller_loop([(p, q)]) for p, q in zip(g1, g2)]
    assert H(g["single_miller"]) == b"".join(map(o.gt_to_bytes, single))
    gts = [o.final_exp(m) for m in single]
    assert H(g["single_gt"]) == b"".join(map(o.gt_to_bytes, gts))
    dbl = [o.miller_loop([(g1[2 * i], g2[2 * i]), (g1[2 * i + 1], g2[2 * i + 1])]) for i in range(4)]
    assert H(g["double_miller"]) == b"".join(map(o.gt_to_bytes, dbl))
    assert H(g["double_gt"]) == b"".join(o.gt_to_bytes(o.final_exp(m)) for m in dbl)
    # pair*pair equals the product of two independent pairings (unit-tests/liner_pair.cpp:66-79)
    assert o.final_exp(dbl[0]) == o.f12_mul(gts[0], gts[1])
    tri = [o.pairing_product([(g1[3 * i + j], g2[3 * i + j]) for j in range(3)]) for i in range(2)]
    assert H(g["triple_gt"]) == b"".join(map(o.gt_to_bytes, tri))
    quad = [o.pairing_product([(g1[4 * i + j], g2[4 * i + j]) for j in range(4)]) for i in range(2)]
    assert H(g["quad_gt"]) == b"".join(map(o.gt_to_bytes, quad))
    one = o.gt_to_bytes(o.F12_ONE)
    assert H(g["inf_g1_gt"]) == one == o.gt_to_bytes(o.pairing(None, g2[0]))
    assert H(g["inf_g2_gt"]) == one == o.gt_to_bytes(o.pairing(g1[0], None))
    m1 = [o.g1_from_affine_bytes(c) for c in chunks(H(g["mixed_g1"]), 96)]
    m2 = [o.g2_from_affine_bytes(c) for c in chunks(H(g["mixed_g2"]), 192)]
    assert H(g["mixed_quad_gt"]) == o.gt_to_bytes(o.pairing_product(list(zip(m1, m2))))
    ss = ints(H(g["gt_pow_scalars"]))
    assert H(g["gt_pow"]) == b"".join(o.gt_to_bytes(o.f12_pow(a, s)) for a, s in zip(gts, ss))
    assert H(g["gt_mul"]) == b"".join(o.gt_to_bytes(o.f12_mul(gts[i], gts[4 + i])) for i in range(4))
    # bilinearity pair(g1^x, g2^y) == pair(g1, g2)^(x*y)  (unit-tests/liner_pair.cpp:42-64; configs[0])
    x, y = ints(H(g["bilinear_xy"]))
    base = o.gt_from_bytes(H(g["generator_gt"]))
    assert H(g["bilinear_lhs_gt"]) == o.gt_to_bytes(o.f12_pow(base, x * y % o.R))
    assert H(g["bilinear_lhs_gt"]) == o.gt_to_bytes(o.pairing(o.g1_mul(o.G1_GEN, x), o.g2_mul(o.G2_GEN, y)))


@pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built (needs /root/reference)")
def test_against_compiled_reference_fresh_seeds():
    assert ref.struct_sizes() == {"big": 56, "fp": 64, "point1": 192, "point2": 384, "fp12": 776}
    sc = ref.random_scalars("fresh-oracle-check", 6)
    kk = ref.random_scalars("fresh-oracle-check-2", 6)
    ks, ss = ints(sc), ints(kk)
    a1 = ref.g1_fixed_base_mul(sc)
    a2 = ref.g2_fixed_base_mul(sc)
    g1 = [o.g1_mul(o.G1_GEN, k) for k in ks]
    g2 = [o.g2_mul(o.G2_GEN, k) for k in ks]
    assert a1 == b"".join(map(o.g1_to_affine_bytes, g1))
    assert a2 == b"".join(map(o.g2_to_affine_bytes, g2))
    for algo in (0, 1, 2):
        assert ref.g1_msm(a1, kk, algo) == o.g1_compress(o.g1_msm_muln(g1, ss))
    assert ref.g2_msm(a2, kk) == o.g2_compress(o.g2_msm(g2, ss))
    assert ref.pairing_product_batch(a1, a2, 3, 0) == b"".join(
        o.gt_to_bytes(o.miller_loop(list(zip(g1[3 * i:3 * i + 3], g2[3 * i:3 * i + 3])))) for i in range(2))
    assert ref.pairing_product_batch(a1, a2, 2, 1) == b"".join(
        o.gt_to_bytes(o.pairing_product(list(zip(g1[2 * i:2 * i + 2], g2[2 * i:2 * i + 2])))) for i in range(3))
    assert ref.random_scalars("same seed", 3) == ref.random_scalars("same seed", 3)  # unit-tests/random.cpp:8-20


def test_hash_to_g1_golden(golden_hashing):
    """G1Point::from_hash restated (sswu_iso_curve / iso11_map / cofactor) against vectors from the compiled reference, including
    the inputs without an SSWU image (u = 0, Z u^2 = -1), where the reference yields the identity."""
    g = golden_hashing
    for m, want in zip(g["messages"], g["points"]):
        assert o.g1_compress(o.hash_to_g1(H(m))) == H(want)
    us = [int.from_bytes(c, "big") for c in chunks(H(g["elements"]), 48)]
    assert b"".join(o.g1_compress(o.map_to_g1(u)) for u in us) == H(g["mapped"])
    assert H(g["mapped"])[:49 * 3] == bytes(49 * 3)
    # structure: the SSWU point is on E', its image on E, the result in the r-torsion subgroup
    q = o.sswu_iso_curve(us[7])
    assert (q[1] * q[1] - q[0] ** 3 - o.ISO_A * q[0] - o.ISO_B) % o.P == 0
    assert o.g1_on_curve(o.iso11_map(q)) and o.g1_mul_raw(o.map_to_g1(us[7]), o.R) is None


@pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built (needs /root/reference)")
def test_hash_to_g1_against_compiled_reference():
    import random
    rnd = random.Random(99)
    msgs = bytes(rnd.randrange(256) for _ in range(8 * 50))
    assert ref.hash_to_g1(msgs, 50, 8, 2) == b"".join(o.g1_compress(o.hash_to_g1(m)) for m in chunks(msgs, 50))
