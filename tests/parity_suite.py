"""Backend-agnostic parity checks against the committed golden vectors (tests/golden/*.json, produced from the
compiled, unmodified reference by tools/gen_golden.py).  Run once over the host mirror of the kernel bodies (CPU)
and once over the CUDA library through its C ABI (GPU)."""
from conftest import chunks, load_golden

H = bytes.fromhex
R = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001
X_ABS = 0xD201000000010000
P = 0x1A0111EA397FE69A4B1BA7B6434BACD764774B84F38512BF6730D2A0F6B0F6241EABFFFEB153FFFFB9FEFFFFFFFFAAAB
ONE_GT = bytes(575) + b"\x01"


def be32(v):
    return int(v).to_bytes(32, "big")


def check_points(be):
    g = load_golden("points.json")
    k16, s16 = H(g["point_scalars"]), H(g["scalars"])
    g1, g2 = H(g["g1_affine"]), H(g["g2_affine"])
    assert be.fixed_base1(k16) == g1
    assert be.fixed_base2(k16[:32 * 8]) == g2
    assert be.fixed_base1(be32(1)) == H(g["g1_generator"])
    assert be.fixed_base2(be32(1)) == H(g["g2_generator"])
    assert be.fixed_base1(be32(0)) == bytes(96) and be.fixed_base2(be32(0)) == bytes(192)
    assert be.mul1(g1, s16) == H(g["g1_mul"])
    assert be.mul2(g2, s16[:32 * 8]) == H(g["g2_mul"])
    es = H(g["edge_scalars"])
    assert be.mul1(g1[:96 * 8], es) == H(g["g1_mul_edge"])
    assert be.mul2(g2, es) == H(g["g2_mul_edge"])
    # identity in, identity out; compression of k*G with k = 1 reproduces the compressed generator
    assert be.mul1(bytes(96), be32(5)) == bytes(49) and be.mul2(bytes(192), be32(5)) == bytes(97)
    assert be.mul1(g1, b"".join(be32(1) for _ in range(16))) == H(g["g1_compressed"])
    assert be.mul2(g2, b"".join(be32(1) for _ in range(8))) == H(g["g2_compressed"])


def msm_case_points(be, case):
    ks = H(case["point_scalars"])
    return be.fixed_base1(ks) if case["group"] == "g1" else be.fixed_base2(ks)


def check_msm(be, max_n=10 ** 9, windows=(0,)):
    g = load_golden("msm.json")
    for case in g["cases"]:
        if case["n"] > max_n:
            continue
        pts = msm_case_points(be, case)
        for c in windows:
            if case["group"] == "g1":
                assert be.msm1(pts, H(case["scalars"]), c) == H(case["result"]), (case["n"], c)
            else:
                assert be.msm2(pts, H(case["scalars"]), c) == H(case["result"]), (case["n"], c)
    e = g["edge_g1"]   # repeated points, P and -P, identity, zero / one scalars (SURVEY §7.2 exceptional cases)
    for c in windows:
        assert be.msm1(H(e["points"]), H(e["scalars"]), c) == H(e["result"]) == H(e["result_live"])
        x = g["cancel_g1"]
        assert be.msm1(H(x["points"]), H(x["scalars"]), c) == bytes(49)
    assert be.msm1(b"", b"", 0) == bytes(49) and be.msm2(b"", b"", 0) == bytes(97)   # empty sum = identity


def check_pairing(be):
    g = load_golden("pairing.json")
    g1, g2 = H(g["g1"]), H(g["g2"])
    pt = load_golden("points.json")
    gen1, gen2 = H(pt["g1_generator"]), H(pt["g2_generator"])
    assert be.miller(gen1, gen2, 1) == H(g["generator_miller"])
    assert be.product(gen1, gen2, 1) == H(g["generator_gt"])
    sm = be.miller(g1, g2, 1)
    assert sm == H(g["single_miller"])
    assert be.final_exp(sm) == H(g["single_gt"]) == be.product(g1, g2, 1)
    assert be.miller(g1, g2, 2) == H(g["double_miller"])
    assert be.product(g1, g2, 2) == H(g["double_gt"])
    assert be.product(g1[:96 * 6], g2[:192 * 6], 3) == H(g["triple_gt"])
    assert be.product(g1, g2, 4) == H(g["quad_gt"])
    assert be.product(bytes(96), g2[:192], 1) == ONE_GT == H(g["inf_g1_gt"])
    assert be.product(g1[:96], bytes(192), 1) == ONE_GT == H(g["inf_g2_gt"])
    assert be.product(H(g["mixed_g1"]), H(g["mixed_g2"]), 4) == H(g["mixed_quad_gt"])
    gts = H(g["single_gt"])
    assert be.gt_pow(gts, H(g["gt_pow_scalars"])) == H(g["gt_pow"])
    # PAIR_GTpow through the Galbraith-Scott split: same values on GT, edge exponents included
    assert be.gt_pow_gs(gts, H(g["gt_pow_scalars"])) == H(g["gt_pow"])
    edge = [0, 1, 2, R - 1, X_ABS, X_ABS ** 2 % R, X_ABS ** 3 % R, (X_ABS // 2 + 1) * (1 + X_ABS + X_ABS ** 2 + X_ABS ** 3) % R]
    es = b"".join(be32(e) for e in edge)
    assert be.gt_pow_gs(gts, es) == be.gt_pow(gts, es)
    assert be.gt_mul(gts[:576 * 4], gts[576 * 4:]) == H(g["gt_mul"])
    # pair*pair == product of two independent pairings (unit-tests/liner_pair.cpp:66-79)
    assert be.gt_mul(gts[:576], gts[576:1152]) == H(g["double_gt"])[:576]
    # bilinearity pair(g1^x, g2^y) == pair(g1, g2)^(x*y)  (unit-tests/liner_pair.cpp:42-64; BASELINE configs[0])
    xy = H(g["bilinear_xy"])
    x, y = int.from_bytes(xy[:32], "big"), int.from_bytes(xy[32:], "big")
    lhs = be.product(be.fixed_base1(xy[:32]), be.fixed_base2(xy[32:]), 1)
    assert lhs == H(g["bilinear_lhs_gt"]) == be.gt_pow(H(g["generator_gt"]), be32(x * y % R))
    assert lhs == be.gt_pow_gs(H(g["generator_gt"]), be32(x * y % R))
