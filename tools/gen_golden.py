#!/usr/bin/env python3
"""Generate tests/golden/*.json from the UNMODIFIED reference (oracle/_ref/libref12381.so).

Run in the build container (where /root/reference exists and `make -C oracle` has been run):
    python tools/gen_golden.py
The vectors are small, committed, and travel to the GPU box, where /root/reference does not exist.
Every value is produced by the reference bridge + MIRACL-core through oracle/ref_shim.cpp; nothing in
here calls the Python restatement or the CUDA path.  Seeds follow the reference's test style
(`create_random_engine("<literal seed>")`, unit-tests/liner_pair.cpp:44)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def hx(b):
    return b.hex()


def main():
    os.makedirs(OUT, exist_ok=True)
    threads = ref.hardware_threads()

    # --- G1 / G2 points and scalar multiplication -------------------------------------------------
    k16 = ref.random_scalars("golden point seed", 16)
    s16 = ref.random_scalars("golden scalar seed", 16)
    g1pts = ref.g1_fixed_base_mul(k16)
    g2pts = ref.g2_fixed_base_mul(k16[: 32 * 8])
    edge_scalars = b"".join(int(v).to_bytes(32, "big") for v in (
        0, 1, 2, 3, 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000000,  # r-1
        1 << 128, (1 << 255) - 19 & ((1 << 254) - 1), 0xD201000000010000))
    g1_8 = g1pts[: 96 * 8]
    g2_8 = g2pts
    points = {
        "seed_points": "golden point seed",
        "seed_scalars": "golden scalar seed",
        "point_scalars": hx(k16),
        "scalars": hx(s16),
        "g1_generator": hx(ref.g1_generator()),
        "g2_generator": hx(ref.g2_generator()),
        "g1_affine": hx(g1pts),
        "g2_affine": hx(g2pts),
        "g1_compressed": hx(ref.g1_compress(g1pts)),
        "g2_compressed": hx(ref.g2_compress(g2pts)),
        "g1_mul": hx(ref.g1_mul_batch(g1pts, s16)),
        "g2_mul": hx(ref.g2_mul_batch(g2pts, s16[: 32 * 8])),
        "edge_scalars": hx(edge_scalars),
        "g1_mul_edge": hx(ref.g1_mul_batch(g1_8, edge_scalars)),
        "g2_mul_edge": hx(ref.g2_mul_batch(g2_8, edge_scalars)),
    }
    with open(os.path.join(OUT, "points.json"), "w") as f:
        json.dump(points, f, indent=1)

    # --- MSM ----------------------------------------------------------------------------------------
    msm = {"cases": []}
    for n in (1, 2, 3, 16, 64, 257):
        kp = ref.random_scalars(f"msm-g1-points-{n}", n)
        ks = ref.random_scalars(f"msm-g1-{n}", n)
        pts = ref.g1_fixed_base_mul(kp, threads)
        a0 = ref.g1_msm(pts, ks, 0)
        a1 = ref.g1_msm(pts, ks, 1)
        assert a0 == a1, "ECP_muln and the live ECP_mul2 loop disagree"
        msm["cases"].append({"group": "g1", "n": n, "seed_points": f"msm-g1-points-{n}", "seed_scalars": f"msm-g1-{n}",
                             "point_scalars": hx(kp), "scalars": hx(ks), "result": hx(a0)})
    for n in (1, 2, 5, 33):
        kp = ref.random_scalars(f"msm-g2-points-{n}", n)
        ks = ref.random_scalars(f"msm-g2-{n}", n)
        pts = ref.g2_fixed_base_mul(kp, threads)
        msm["cases"].append({"group": "g2", "n": n, "seed_points": f"msm-g2-points-{n}", "seed_scalars": f"msm-g2-{n}",
                             "point_scalars": hx(kp), "scalars": hx(ks), "result": hx(ref.g2_msm(pts, ks))})
    # a 1024-term G1 MSM (BASELINE.json configs[0]); points are NOT stored, only their scalars
    n = 1024
    kp = ref.random_scalars("msm-g1-points-1024", n)
    ks = ref.random_scalars("msm-g1-1024", n)
    pts = ref.g1_fixed_base_mul(kp, threads)
    r0 = ref.g1_msm(pts, ks, 0, threads)
    r1 = ref.g1_msm(pts, ks, 1, threads)
    assert r0 == r1
    msm["cases"].append({"group": "g1", "n": n, "seed_points": "msm-g1-points-1024", "seed_scalars": "msm-g1-1024",
                         "point_scalars": hx(kp), "scalars": hx(ks), "result": hx(r0)})
    # edge: zero scalars, repeated points, P and -P, identity points
    ident = bytes(96)
    p0 = g1pts[:96]
    negp0 = p0[:48] + (0x1A0111EA397FE69A4B1BA7B6434BACD764774B84F38512BF6730D2A0F6B0F6241EABFFFEB153FFFFB9FEFFFFFFFFAAAB
                       - int.from_bytes(p0[48:], "big")).to_bytes(48, "big")
    epts = p0 + p0 + negp0 + ident + g1pts[96:192] + p0 + g1pts[96:192] + g1pts[192:288]
    one = (1).to_bytes(32, "big")
    esc = s16[:32] + s16[:32] + s16[:32] + s16[32:64] + bytes(32) + one + one + s16[64:96]
    msm["edge_g1"] = {"points": hx(epts), "scalars": hx(esc), "result": hx(ref.g1_msm(epts, esc, 0)),
                      "result_live": hx(ref.g1_msm(epts, esc, 1))}
    # all cancels to the identity
    epts2 = p0 + negp0
    esc2 = s16[:32] + s16[:32]
    msm["cancel_g1"] = {"points": hx(epts2), "scalars": hx(esc2), "result": hx(ref.g1_msm(epts2, esc2, 0))}
    with open(os.path.join(OUT, "msm.json"), "w") as f:
        json.dump(msm, f, indent=1)

    # --- pairings -------------------------------------------------------------------------------------
    pair = {}
    g1g = ref.g1_generator()
    g2g = ref.g2_generator()
    pair["generator_gt"] = hx(ref.pairing_product_batch(g1g, g2g, 1, 1))
    pair["generator_miller"] = hx(ref.pairing_product_batch(g1g, g2g, 1, 0))
    pair["g1"] = hx(g1_8)
    pair["g2"] = hx(g2_8)
    pair["single_miller"] = hx(ref.pairing_product_batch(g1_8, g2_8, 1, 0))      # 8 x PAIR_ate
    pair["single_gt"] = hx(ref.pairing_product_batch(g1_8, g2_8, 1, 1))
    pair["double_miller"] = hx(ref.pairing_product_batch(g1_8, g2_8, 2, 0))      # 4 x PAIR_double_ate
    pair["double_gt"] = hx(ref.pairing_product_batch(g1_8, g2_8, 2, 1))
    pair["triple_gt"] = hx(ref.pairing_product_batch(g1_8[: 96 * 6], g2_8[: 192 * 6], 3, 1))
    pair["quad_gt"] = hx(ref.pairing_product_batch(g1_8, g2_8, 4, 1))           # config 4 shape
    # degenerate inputs (unit-tests/liner_pair.cpp:28-40)
    pair["inf_g1_gt"] = hx(ref.pairing_product_batch(bytes(96), g2_8[:192], 1, 1))
    pair["inf_g2_gt"] = hx(ref.pairing_product_batch(g1_8[:96], bytes(192), 1, 1))
    mixed_g1 = g1_8[:96] + bytes(96) + g1_8[192:288] + g1_8[288:384]
    mixed_g2 = g2_8[:192] + g2_8[192:384] + bytes(192) + g2_8[576:768]
    pair["mixed_g1"] = hx(mixed_g1)
    pair["mixed_g2"] = hx(mixed_g2)
    pair["mixed_quad_gt"] = hx(ref.pairing_product_batch(mixed_g1, mixed_g2, 4, 1))
    # GT arithmetic
    gts = ref.pairing_product_batch(g1_8, g2_8, 1, 1)
    pair["gt_pow_scalars"] = hx(s16[: 32 * 8])
    pair["gt_pow"] = hx(ref.gt_pow_batch(gts, s16[: 32 * 8]))
    pair["gt_mul"] = hx(ref.gt_mul_batch(gts[: 576 * 4], gts[576 * 4:]))
    # bilinearity instance exactly as unit-tests/liner_pair.cpp:42-64 states it, seed included
    xy = ref.random_scalars("pairing bilinearity seed", 2)
    gx = ref.g1_fixed_base_mul(xy[:32])
    gy = ref.g2_fixed_base_mul(xy[32:])
    pair["bilinear_xy"] = hx(xy)
    pair["bilinear_lhs_gt"] = hx(ref.pairing_product_batch(gx, gy, 1, 1))
    with open(os.path.join(OUT, "pairing.json"), "w") as f:
        json.dump(pair, f, indent=1)
    # --- hashing to G1 (G1Point::from_hash, g1_point.hpp:219-234) ---------------------------------------------------
    import random
    rnd = random.Random(381)
    P = 0x1A0111EA397FE69A4B1BA7B6434BACD764774B84F38512BF6730D2A0F6B0F6241EABFFFEB153FFFFB9FEFFFFFFFFAAAB
    hashing = {"messages": [], "points": []}
    for n in (1, 3, 32, 71, 72, 73, 145, 300):
        msg = bytes(rnd.randrange(256) for _ in range(n))
        hashing["messages"].append(hx(msg))
        hashing["points"].append(hx(ref.hash_to_g1(msg, n, 1)))
    # field elements: 0 and +-sqrt(-1/11) (the inputs without an SSWU image: the reference yields the identity), small and random ones
    s = pow((-pow(11, -1, P)) % P, (P + 1) // 4, P)
    assert 11 * s * s % P == P - 1
    us = [0, s, P - s, 1, 2, P - 1] + [rnd.randrange(P) for _ in range(10)]
    u48 = b"".join(u.to_bytes(48, "big") for u in us)
    hashing["elements"] = hx(u48)
    hashing["mapped"] = hx(ref.map_to_g1(u48, len(us)))
    with open(os.path.join(OUT, "hashing.json"), "w") as f:
        json.dump(hashing, f, indent=1)
    print("golden vectors written to", OUT)


if __name__ == "__main__":
    main()
