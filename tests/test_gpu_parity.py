"""GPU (-m gpu): the CUDA path, called through the C ABI (crypto12381_b200.bridge / .device -> libc12381_cuda.so),
against (a) the golden vectors of the reference, (b) the compiled reference itself (oracle/_ref, which travels to the
GPU box) on fresh seeded inputs, and (c) size-independent algebraic properties at BASELINE.json's full sizes.
All comparisons are bit-exact on the reference's serialised formats."""
import ctypes
import random

import numpy as np
import pytest
import torch

import parity_suite as ps
from conftest import chunks, load_golden
from oracle import ref

pytestmark = pytest.mark.gpu

R = ps.R
be32 = ps.be32


@pytest.fixture(scope="module")
def cuda():
    if not torch.cuda.is_available():
        pytest.fail("the -m gpu tests need a CUDA device (there is no CPU fallback to test)")
    from crypto12381_b200 import _lib, bridge, device
    _lib.init(0)

    class CudaBackend:
        fixed_base1 = staticmethod(bridge.generator_power)
        fixed_base2 = staticmethod(bridge.generator_power2)
        mul1 = staticmethod(bridge.multiply)
        mul2 = staticmethod(bridge.multiply2)
        final_exp = staticmethod(bridge.pair_final_exponentiation)
        gt_mul = staticmethod(bridge.gt_multiply)
        gt_pow = staticmethod(bridge.gt_pow)
        gt_pow_gs = staticmethod(bridge.gt_pow_gs)
        miller = staticmethod(bridge.miller_batch)
        product = staticmethod(bridge.pairing_product_batch)

        @staticmethod
        def msm1(points, scalars, c=0):
            _lib.lib().c12381_set_msm_window(c)
            try:
                return bridge.sum_of_products(points, scalars)
            finally:
                _lib.lib().c12381_set_msm_window(0)

        @staticmethod
        def msm2(points, scalars, c=0):
            _lib.lib().c12381_set_msm_window(c)
            try:
                return bridge.sum_of_products2(points, scalars)
            finally:
                _lib.lib().c12381_set_msm_window(0)

    CudaBackend.lib = _lib
    CudaBackend.bridge = bridge
    CudaBackend.device = device
    return CudaBackend


def rand_scalars(n, seed):
    """n scalars < r, 32 B big-endian each (top byte <= 0x72 keeps them below r = 0x73ed...)."""
    rng = np.random.default_rng(seed)
    a = rng.integers(0, 256, size=(n, 32), dtype=np.uint8)
    a[:, 0] = rng.integers(0, 0x73, size=n, dtype=np.uint8)
    return a.tobytes()


def ints(b):
    return [int.from_bytes(c, "big") for c in chunks(b, 32)]


# ---- field + roofline probes ----------------------------------------------------------------------------------------
def test_probes_run(cuda):
    for kind in range(5):
        r = cuda.device.probe(kind, 200)
        assert r["gops"] > 0 and r["ms"] > 0


# ---- golden vectors ------------------------------------------------------------------------------------------------
def test_points_golden(cuda):
    ps.check_points(cuda)


@pytest.fixture(params=[(0, 1, 1, 0, 2), (1, 1, 1, 0, 0), (3, 2, 2, 0, 2), (9, 4, 3, 0, 2), (2, 1, 4, 0, 0), (0, 1, 1, 1, 2), (3, 2, 2, 1, 0), (4, 2, 2, 0, 1)],
                ids=["xyzz-only", "batch-affine-1", "batch-affine-3x2-groups2", "batch-affine-9-groups3", "batch-affine-2-groups4",
                     "xyzz-only-sorted-lists", "batch-affine-3x2-groups2-sorted-lists", "batch-affine-4x2-split-tail"])
def ba_rounds(request):
    """(forced batch-affine halving rounds, pipelines they are split into, upload groups of the host entry, front end: bucket
    lists by counting = 0 / by the segmented radix sort = 1, tail: 0 one chain / 1 early split / 2 late split); the defaults are -1
    (rounds chosen from the bucket load), 2, 4, 0, 2"""
    from crypto12381_b200 import _lib
    _lib.lib().c12381_set_msm_batch_affine(request.param[0])
    _lib.lib().c12381_set_msm_pipelines(request.param[1])
    _lib.lib().c12381_set_knob(4, request.param[2])
    _lib.lib().c12381_set_knob(5, request.param[3])
    _lib.lib().c12381_set_knob(7, request.param[4])
    yield request.param
    _lib.lib().c12381_set_msm_batch_affine(-1)
    _lib.lib().c12381_set_msm_pipelines(2)
    _lib.lib().c12381_set_knob(4, 4)
    _lib.lib().c12381_set_knob(5, 0)
    _lib.lib().c12381_set_knob(7, 2)


def test_msm_golden_all_windows(cuda, ba_rounds):
    ps.check_msm(cuda, windows=(0, 2, 3, 5, 8, 9, 13, 16))


@pytest.fixture(params=[1, 2], ids=["thread-per-instance", "cooperative"])
def pairing_kernel(request):
    from crypto12381_b200 import _lib
    _lib.lib().c12381_set_pairing_kernel(request.param)
    yield request.param
    _lib.lib().c12381_set_pairing_kernel(0)


def test_pairing_golden(cuda, pairing_kernel):
    ps.check_pairing(cuda)


# ---- the compiled reference on fresh seeds ----------------------------------------------------------------------------
needs_ref = pytest.mark.skipif(not ref.available(), reason="oracle/_ref/libref12381.so not present")


@needs_ref
@pytest.mark.parametrize("n", [1, 2, 3, 31, 100, 1000, 4097, 20000])
def test_g1_msm_vs_reference(cuda, n, ba_rounds):
    t = ref.hardware_threads()
    ks = ref.random_scalars(f"msm-g1-points-{n}", n)
    ss = ref.random_scalars(f"msm-g1-{n}", n)
    pts = cuda.fixed_base1(ks)
    if n <= 1000:
        assert pts == ref.g1_fixed_base_mul(ks, t)
    assert cuda.msm1(pts, ss) == ref.g1_msm(pts, ss, 0, t)
    if n <= 100:   # the live DSL loop (ECP_mul2 pairs) and the naive sum agree with ECP_muln
        assert cuda.msm1(pts, ss) == ref.g1_msm(pts, ss, 1, t) == ref.g1_msm(pts, ss, 2, t)


@needs_ref
@pytest.mark.parametrize("n", [1, 2, 17, 300, 3000])
def test_g2_msm_vs_reference(cuda, n, ba_rounds):
    t = ref.hardware_threads()
    ks = ref.random_scalars(f"msm-g2-points-{n}", n)
    ss = ref.random_scalars(f"msm-g2-{n}", n)
    pts = cuda.fixed_base2(ks)
    if n <= 300:
        assert pts == ref.g2_fixed_base_mul(ks, t)
    assert cuda.msm2(pts, ss) == ref.g2_msm(pts, ss, t)


# ---- the compiled reference at the sizes the benchmark quotes (differentials had stopped at 20 000 / 3 000 / 16) ----------
@needs_ref
def test_g1_msm_vs_reference_2p18(cuda):
    """G1, n = 2^18: ECP_muln over all host threads (about a second on 16) against the whole CUDA pipeline at its c = 16 plan."""
    n, t = 1 << 18, ref.hardware_threads()
    ks, ss = rand_scalars(n, 1801), rand_scalars(n, 1802)
    pts = cuda.fixed_base1(ks)
    assert cuda.msm1(pts, ss) == ref.g1_msm(pts, ss, 0, t)


@needs_ref
def test_g2_msm_vs_reference_2p16(cuda):
    """G2, n = 2^16: the reference's per-term PAIR_G2mul + ECP2_add loop over all host threads."""
    n, t = 1 << 16, ref.hardware_threads()
    ks, ss = rand_scalars(n, 1601), rand_scalars(n, 1602)
    pts = cuda.fixed_base2(ks)
    assert cuda.msm2(pts, ss) == ref.g2_msm(pts, ss, t)


@needs_ref
def test_pairing_products_vs_reference_1024x4(cuda, pairing_kernel):
    """1024 instances x 4 pairs (BASELINE configs[3]'s shape): raw Miller products and GT values, every instance."""
    B, k, t = 1024, 4, ref.hardware_threads()
    g1, g2 = cuda.fixed_base1(rand_scalars(B * k, 1024)), cuda.fixed_base2(rand_scalars(B * k, 1025))
    assert cuda.product(g1, g2, k) == ref.pairing_product_batch(g1, g2, k, 1, t)
    assert cuda.miller(g1, g2, k) == ref.pairing_product_batch(g1, g2, k, 0, t)


@needs_ref
def test_mul_batch_vs_reference(cuda):
    t = ref.hardware_threads()
    n = 257
    ks, ss = ref.random_scalars("mul-batch-points", n), ref.random_scalars("mul-batch", n)
    p1, p2 = cuda.fixed_base1(ks), cuda.fixed_base2(ks)
    assert cuda.mul1(p1, ss) == ref.g1_mul_batch(p1, ss, t)
    assert cuda.mul2(p2, ss) == ref.g2_mul_batch(p2, ss, t)


@needs_ref
@pytest.mark.parametrize("k", [1, 2, 3, 4, 8])
def test_pairing_products_vs_reference(cuda, k, pairing_kernel):
    t = ref.hardware_threads()
    B = 16
    a, b = ref.random_scalars(f"pair-g1-{k}", B * k), ref.random_scalars(f"pair-g2-{k}", B * k)
    g1, g2 = cuda.fixed_base1(a), cuda.fixed_base2(b)
    assert cuda.miller(g1, g2, k) == ref.pairing_product_batch(g1, g2, k, 0, t)
    gt = cuda.product(g1, g2, k)
    assert gt == ref.pairing_product_batch(g1, g2, k, 1, t)
    assert cuda.final_exp(cuda.miller(g1, g2, k)) == gt
    e = ref.random_scalars("gt-exp", B)
    assert cuda.gt_pow(gt, e) == ref.gt_pow_batch(gt, e, t) == cuda.gt_pow_gs(gt, e)
    assert cuda.gt_mul(gt, gt[576:] + gt[:576]) == ref.gt_mul_batch(gt, gt[576:] + gt[:576])


@needs_ref
def test_reference_pods_pass_through(cuda):
    """The drop-in entries on the reference's own structs (what the forwarding TU of INTEGRATION.md calls)."""
    br = cuda.bridge
    n = 9
    ks, ss = ref.random_scalars("pod-points", n), ref.random_scalars("pod-scalars", n)
    a1, a2 = ref.g1_fixed_base_mul(ks), ref.g2_fixed_base_mul(ks)
    for unnorm in (False, True):
        p1, p2, bg = ref.make_point1(a1, unnorm), ref.make_point2(a2, unnorm), ref.make_big(ss)
        r1 = ctypes.create_string_buffer(192)
        br.sum_of_products_pod(r1, n, p1, bg)
        assert ref.point1_to_c49(r1, 1) == ref.g1_msm(a1, ss)
        r2 = ctypes.create_string_buffer(384)
        br.sum_of_products2_pod(r2, n, p2, bg)
        assert ref.point2_to_c97(r2, 1) == ref.g2_msm(a2, ss)
        # multiply / double_multiply mutate their first argument in place, like the reference
        o1 = ctypes.create_string_buffer(p1.raw[:192], 192)
        br.multiply_pod(o1, ctypes.create_string_buffer(bg.raw[:56], 56))
        assert ref.point1_to_c49(o1, 1) == ref.g1_mul_batch(a1[:96], ss[:32])
        o2 = ctypes.create_string_buffer(p2.raw[:384], 384)
        br.multiply2_pod(o2, ctypes.create_string_buffer(bg.raw[:56], 56))
        assert ref.point2_to_c97(o2, 1) == ref.g2_mul_batch(a2[:192], ss[:32])
        d1 = ctypes.create_string_buffer(p1.raw[:192], 192)
        br.double_multiply_pod(d1, ctypes.create_string_buffer(p1.raw[192:384], 192),
                               ctypes.create_string_buffer(bg.raw[:56], 56), ctypes.create_string_buffer(bg.raw[56:112], 56))
        assert ref.point1_to_c49(d1, 1) == ref.g1_msm(a1[:192], ss[:64])
        f = ctypes.create_string_buffer(776)
        br.pair_ate_pod(f, ctypes.create_string_buffer(p2.raw[:384], 384), ctypes.create_string_buffer(p1.raw[:192], 192))
        assert ref.fp12_to_bytes(f, 1) == ref.pairing_product_batch(a1[:96], a2[:192], 1, 0)
        assert ctypes.c_int.from_buffer(f, 768).value == 5   # FP_DENSE
        br.pair_final_exponentiation_pod(f)
        gt = ref.pairing_product_batch(a1[:96], a2[:192], 1, 1)
        assert ref.fp12_to_bytes(f, 1) == gt
        f2 = ctypes.create_string_buffer(776)
        br.pair_double_ate_pod(f2, ctypes.create_string_buffer(p2.raw[:384], 384), ctypes.create_string_buffer(p1.raw[:192], 192),
                               ctypes.create_string_buffer(p2.raw[384:768], 384), ctypes.create_string_buffer(p1.raw[192:384], 192))
        assert ref.fp12_to_bytes(f2, 1) == ref.pairing_product_batch(a1[:192], a2[:384], 2, 0)
        g = ctypes.create_string_buffer(776)
        br.gt_pow_pod(g, f, ctypes.create_string_buffer(bg.raw[:56], 56))
        assert ref.fp12_to_bytes(g, 1) == ref.gt_pow_batch(gt, ss[:32])
        br.gt_multiply_pod(g, f)
        assert ref.fp12_to_bytes(g, 1) == ref.gt_mul_batch(ref.gt_pow_batch(gt, ss[:32]), gt)


# ---- full-size properties (no oracle needed) --------------------------------------------------------------------------
@pytest.mark.parametrize("log_n", [10, 16, 20, 22])
def test_g1_msm_full_size_linearity(cuda, log_n):
    """Σ s_i (k_i G) == (Σ s_i k_i mod r) G at BASELINE's n = 2^20, on device-resident data (the _dev entries)."""
    n = 1 << log_n
    dv = cuda.device
    ks, ss = rand_scalars(n, 100 + log_n), rand_scalars(n, 200 + log_n)
    d_k = torch.frombuffer(bytearray(ks), dtype=torch.uint8).cuda()
    d_s = torch.frombuffer(bytearray(ss), dtype=torch.uint8).cuda()
    pts = dv.g1_fixed_base_mul_batch(d_k)
    got = bytes(dv.g1_msm(pts, d_s).cpu().numpy())
    dv.sync_status()
    total = sum(a * b for a, b in zip(ints(ks), ints(ss))) % R
    gen = bytes.fromhex(load_golden("points.json")["g1_generator"])
    assert got == cuda.mul1(gen, be32(total))
    # run to run: bit-identical (sort-based buckets, fixed reduction trees)
    assert got == bytes(dv.g1_msm(pts, d_s).cpu().numpy())
    # sharded form with world size 1 == the plain form
    from crypto12381_b200.distributed import g1_msm_sharded
    assert bytes(g1_msm_sharded(pts, d_s).cpu().numpy()) == got
    # two half-size partials merged by the point-sum entry == the whole (what 2 ranks would compute)
    h = n // 2
    parts = torch.cat([dv.g1_msm_partial(pts[:96 * h], d_s[:32 * h]), dv.g1_msm_partial(pts[96 * h:], d_s[32 * h:])])
    assert bytes(dv.g1_sum(parts).cpu().numpy()) == got


def test_g1_msm_sweep_maximum(cuda):
    """The top of BASELINE's sweep, n = 2^24: the one-call sum equals sixteen 2^20-term partials merged by the point-sum
    entry (a different plan per call: the property holds only if both are the true sum), and is bit-identical run to run."""
    n, chunk = 1 << 24, 1 << 20
    dv = cuda.device
    g = torch.Generator(device="cuda").manual_seed(2024)
    d_k = torch.randint(0, 256, (n, 32), dtype=torch.uint8, device="cuda", generator=g)
    d_s = torch.randint(0, 256, (n, 32), dtype=torch.uint8, device="cuda", generator=g)
    d_k[:, 0] %= 0x73
    d_s[:, 0] %= 0x73
    d_k, d_s = d_k.reshape(-1), d_s.reshape(-1)
    pts = dv.g1_fixed_base_mul_batch(d_k)
    del d_k
    whole = bytes(dv.g1_msm(pts, d_s).cpu().numpy())
    dv.sync_status()
    parts = torch.cat([dv.g1_msm_partial(pts[96 * i:96 * (i + chunk)], d_s[32 * i:32 * (i + chunk)]) for i in range(0, n, chunk)])
    assert bytes(dv.g1_sum(parts).cpu().numpy()) == whole
    assert bytes(dv.g1_msm(pts, d_s).cpu().numpy()) == whole
    assert whole != bytes(49)


def test_g2_msm_full_size_linearity(cuda):
    n = 1 << 16
    dv = cuda.device
    ks, ss = rand_scalars(n, 31), rand_scalars(n, 32)
    d_k = torch.frombuffer(bytearray(ks), dtype=torch.uint8).cuda()
    d_s = torch.frombuffer(bytearray(ss), dtype=torch.uint8).cuda()
    pts = dv.g2_fixed_base_mul_batch(d_k)
    got = bytes(dv.g2_msm(pts, d_s).cpu().numpy())
    dv.sync_status()
    total = sum(a * b for a, b in zip(ints(ks), ints(ss))) % R
    gen = bytes.fromhex(load_golden("points.json")["g2_generator"])
    assert got == cuda.mul2(gen, be32(total))
    h = n // 2
    parts = torch.cat([dv.g2_msm_partial(pts[:192 * h], d_s[:32 * h]), dv.g2_msm_partial(pts[192 * h:], d_s[32 * h:])])
    assert bytes(dv.g2_sum(parts).cpu().numpy()) == got


def test_skewed_scalars_and_repeated_points(cuda, ba_rounds):
    """Ragged buckets: every scalar equal (one bucket per window takes all terms), tiny scalars, repeated points."""
    n = 3000
    gen = bytes.fromhex(load_golden("points.json")["g1_generator"])
    ks = rand_scalars(n, 5)
    pts = cuda.fixed_base1(ks)
    s = 0x1234567890ABCDEF1234567890ABCDEF1234567890ABCDEF
    assert cuda.msm1(pts, be32(s) * n) == cuda.mul1(gen, be32(sum(ints(ks)) * s % R))
    small = b"".join(be32(i % 3) for i in range(n))
    assert cuda.msm1(pts, small) == cuda.mul1(gen, be32(sum(k * (i % 3) for i, k in enumerate(ints(ks))) % R))
    same = pts[:96] * n   # n copies of one point: every bucket addition past the first is a doubling or P + kP
    ss = rand_scalars(n, 6)
    assert cuda.msm1(same, ss) == cuda.mul1(pts[:96], be32(sum(ints(ss)) % R))


def test_heavy_buckets_at_size(cuda):
    """2^16 equal scalars (G1) / 2^14 (G2): every window is ONE bucket of thousands of chunks - the lists k_fold hands to
    k_fold_heavy - under the default settings (halving rounds off at this size) and with rounds forced on."""
    from crypto12381_b200 import _lib
    s = 0x5A5A1234567890ABCDEF1234567890ABCDEF1234567890ABCDEF
    for g1, n in ((True, 1 << 16), (False, 1 << 14)):
        ks = rand_scalars(n, 77)
        gen = bytes.fromhex(load_golden("points.json")["g1_generator" if g1 else "g2_generator"])
        pts = (cuda.fixed_base1 if g1 else cuda.fixed_base2)(ks)
        want = (cuda.mul1 if g1 else cuda.mul2)(gen, be32(sum(ints(ks)) * s % R))
        for rounds in (-1, 3):
            _lib.lib().c12381_set_msm_batch_affine(rounds)
            try:
                assert (cuda.msm1 if g1 else cuda.msm2)(pts, be32(s) * n) == want, (g1, rounds)
            finally:
                _lib.lib().c12381_set_msm_batch_affine(-1)


def test_pairing_check_full_batch(cuda, pairing_kernel):
    """2^12 instances x 4 pairs with  Π_j e(a_j G1, b_j G2) · e(-(Σ a_j b_j) G1, G2) == 1; flipped instances fail."""
    B, k = 1 << 12, 4
    dv = cuda.device
    a, b = rand_scalars(B * 3, 41), rand_scalars(B * 3, 42)
    ai, bi = ints(a), ints(b)
    g1s, g2s = bytearray(), bytearray()
    last = b"".join(be32((-sum(ai[3 * i + j] * bi[3 * i + j] for j in range(3))) % R) for i in range(B))
    P = cuda.fixed_base1(a)
    Q = cuda.fixed_base2(b)
    Pl = cuda.fixed_base1(last)
    gen2 = bytes.fromhex(load_golden("points.json")["g2_generator"])
    bad = set(random.Random(3).sample(range(B), 37))
    for i in range(B):
        g1s += P[96 * 3 * i:96 * 3 * (i + 1)] + (Pl[96 * ((i + 1) % B):96 * ((i + 1) % B) + 96] if i in bad else Pl[96 * i:96 * i + 96])
        g2s += Q[192 * 3 * i:192 * 3 * (i + 1)] + gen2
    d1 = torch.frombuffer(g1s, dtype=torch.uint8).cuda()
    d2 = torch.frombuffer(g2s, dtype=torch.uint8).cuda()
    verdict = dv.pairing_check_batch(d1, d2, k).cpu().numpy()
    dv.sync_status()
    assert [i for i in range(B) if verdict[i] == 0] == sorted(bad)
    gt = dv.pairing_product_batch(d1, d2, k).cpu().numpy().tobytes()
    assert all((gt[576 * i:576 * (i + 1)] == ps.ONE_GT) == (i not in bad) for i in range(B))


# ---- one arena, several streams ----------------------------------------------------------------------------
def test_interleaved_streams_share_the_arena(cuda):
    """`_dev` MSMs on two torch streams and host entries on the context's own stream, issued back to back without any
    synchronisation: every call carves from the same scratch arena, so each must be ordered behind the previous one."""
    dv, br = cuda.device, cuda.bridge
    n = 1 << 16
    ks, ss = rand_scalars(n, 901), rand_scalars(n, 902)
    d_k = torch.frombuffer(bytearray(ks), dtype=torch.uint8).cuda()
    d_s = torch.frombuffer(bytearray(ss), dtype=torch.uint8).cuda()
    pts = dv.g1_fixed_base_mul_batch(d_k)
    want = bytes(dv.g1_msm(pts, d_s).cpu().numpy())
    want_half = bytes(dv.g1_msm(pts[:96 * (n // 2)], d_s[:32 * (n // 2)]).cpu().numpy())
    comp = br.to_bytes(bytes(pts[:96 * 64].cpu().numpy()))
    torch.cuda.synchronize()
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    for rep in range(4):
        with torch.cuda.stream(s1):
            a = dv.g1_msm(pts, d_s)
        with torch.cuda.stream(s2):
            b = dv.g1_msm(pts[:96 * (n // 2)], d_s[:32 * (n // 2)])
        aff = br.from_bytes(comp)                       # host entry: context stream, its own copies
        with torch.cuda.stream(s1):
            c = dv.g1_msm(pts, d_s)
        small = br.sum_of_products(bytes(pts[:96 * 33].cpu().numpy()), ss[:32 * 33])
        torch.cuda.synchronize()
        assert bytes(a.cpu().numpy()) == want and bytes(c.cpu().numpy()) == want and bytes(b.cpu().numpy()) == want_half
        assert aff == bytes(pts[:96 * 64].cpu().numpy())
        assert small == bytes(dv.g1_msm(pts[:96 * 33], d_s[:32 * 33]).cpu().numpy())
    dv.sync_status()


def test_device_wrappers_validate_sizes(cuda):
    """A mismatched tensor is a ValueError in Python, never an out-of-bounds access on the device."""
    dv = cuda.device
    z = lambda n: torch.zeros(n, dtype=torch.uint8, device="cuda")
    for fn, args in ((dv.g1_mul_batch, (z(96), z(64))), (dv.g2_mul_batch, (z(192 * 2), z(32))), (dv.gt_mul_batch, (z(576 * 2), z(576))),
                     (dv.gt_pow_batch, (z(576), z(64))), (dv.gt_pow_gs_batch, (z(576 * 2), z(32))), (dv.g1_sum, (z(97),)), (dv.g2_sum, (z(191),)),
                     (dv.final_exp_batch, (z(577),)), (dv.sha3_512_batch, (z(10), 3)), (dv.g1_msm, (z(96), z(32), z(48))),
                     (dv.g1_fixed_base_mul_batch, (z(32), torch.zeros(96, dtype=torch.uint8))), (dv.g1_compress_batch, (z(96), z(50))),
                     (dv.pairing_check_batch, (z(96), z(192), 1, z(2))), (dv.g1_multi_fixed_base_batch, (z(96 * 2), z(32 * 3)))):
        with pytest.raises(ValueError):
            fn(*args)
    with pytest.raises(ValueError):
        dv.g1_msm(torch.zeros(96, dtype=torch.uint8), torch.zeros(32, dtype=torch.uint8))     # CPU tensors


@needs_ref
def test_pod_scalar_at_or_above_r_is_reduced(cuda):
    """multiply(point1&, big) with value >= r: the reference (PAIR_G1mul) reduces mod r, so does the drop-in entry - no error."""
    br = cuda.bridge
    k = ref.random_scalars("pod-big-reduce", 1)
    a1 = ref.g1_fixed_base_mul(k)
    for v in (R, R + 5, 2 * R - 1):
        big = ref.make_big(be32(v))
        o1 = ref.make_point1(a1)
        br.multiply_pod(o1, big)
        assert ref.point1_to_c49(o1, 1) == ref.g1_mul_batch(a1, be32(v % R))


# ---- error behaviour -------------------------------------------------------------------------------------------------
def test_malformed_input_is_reported_not_computed(cuda):
    lib = cuda.lib
    g = load_golden("points.json")
    p = bytearray(bytes.fromhex(g["g1_affine"])[:96])
    p[95] ^= 1                                   # off the curve
    with pytest.raises(lib.C12381Error) as e:
        cuda.msm1(bytes(p), be32(3))
    assert e.value.code == lib.EINPUT
    with pytest.raises(lib.C12381Error) as e:    # non-canonical coordinate (x = p)
        cuda.mul1(ps.be32(0)[:0] + (0x1A0111EA397FE69A4B1BA7B6434BACD764774B84F38512BF6730D2A0F6B0F6241EABFFFEB153FFFFB9FEFFFFFFFFAAAB).to_bytes(48, "big") + bytes(48), be32(1))
    assert e.value.code == lib.EINPUT
    with pytest.raises(lib.C12381Error) as e:    # scalar >= r
        cuda.msm1(bytes.fromhex(g["g1_affine"])[:96], be32(R))
    assert e.value.code == lib.EINPUT
    with pytest.raises(ValueError):
        cuda.product(bytes(96 * 9), bytes(192 * 9), 9)
    # the context stays usable
    assert cuda.msm1(bytes.fromhex(g["g1_affine"])[:96], be32(1)) == bytes.fromhex(g["g1_compressed"])[:49]
