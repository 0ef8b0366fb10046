"""BBS+ signatures in batch on the GPU: the reference's `examples/bbs-plus` (src/bbs+.cpp:38-73) restated over the
batched C-ABI entries, for MANY signatures under one set of public parameters and one key (BASELINE configs[4]).

Wire formats are the example's own (`serialize(...)` output):
    pp.g1_g2_h0 = G1 49 B || G2 97 B || G1 49 B        pp.h[i] = G1 49 B        pk = w, G2 97 B        sk = γ, Zp 48 B
    signature   = A (G1 49 B) || x (Zp 48 B) || r (Zp 48 B)                       message blocks: encode_to<Zp>

verify (bbs+.cpp:57-73):   pair(A, w * g2^x) == pair(g1 * h0^r * Π[n](h[i]^m[i]), g2)
per batch of B signatures this becomes
    A_i            <- from_bytes                 c12381_g1_decompress_batch        (one square root per signature)
    W_i  = w + x_i g2                            c12381_g2_multi_fixed_base_batch  (bases w, g2; scalars 1, x_i)
    B_i  = g1 + r_i h0 + Σ_j m_ij h_j            c12381_g1_multi_fixed_base_batch  (bases g1, h0, h_0..h_(n-1))
    e(A_i, W_i) · e(B_i, -g2) == 1               c12381_pairing_check_batch, k = 2 (one shared final exponentiation)
Instances are independent: with several GPUs they are split across ranks with no collective (distributed.gather_results
collects the verdict bytes if they are wanted in one place).

`sign_batch` exists so tests and the benchmark can make valid inputs; the scalar-field arithmetic (the inverse of γ + x)
stays on the host exactly as the reference keeps `Zp` on the CPU (out of scope)."""
from __future__ import annotations

from typing import List, Sequence

from . import bridge

R = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001
P = 0x1A0111EA397FE69A4B1BA7B6434BACD764774B84F38512BF6730D2A0F6B0F6241EABFFFEB153FFFFB9FEFFFFFFFFAAAB
ZP_BYTES, G1_BYTES, G2_BYTES, SIG_BYTES = 48, 49, 97, 49 + 48 + 48
UNIT = 31   # encode_to<Zp>: 31-byte message blocks, bit 248 set (zp_number.hpp:1011-1037)


def be32(v: int) -> bytes:
    return int(v).to_bytes(32, "big")


def encode_to_zp(message: bytes) -> List[int]:
    """encode_to<Zp>(message): one scalar 2^248 + block per 31-byte block; a short last block is followed by zero bytes,
    as the reference's memcpy into the start of the zeroed 31-byte tail leaves it."""
    out = []
    full = len(message) // UNIT
    for i in range(full):
        out.append((1 << 248) | int.from_bytes(message[UNIT * i:UNIT * (i + 1)], "big"))
    rest = len(message) % UNIT
    if rest:
        out.append((1 << 248) | int.from_bytes(message[-rest:] + bytes(UNIT - rest), "big"))
    return out


def _zp(b: bytes) -> int:
    """parse<Zp>: 48 bytes big-endian, must be < r (zp_number.hpp:230-233 throws otherwise)."""
    v = int.from_bytes(b, "big")
    if len(b) != ZP_BYTES or v >= R:
        raise ValueError("Zp element out of range")
    return v


def _neg_g2(q: bytes) -> bytes:
    """-(x, y) on the 192-byte affine form x.b || x.a || y.b || y.a."""
    if q == bytes(192):
        return q
    yb, ya = int.from_bytes(q[96:144], "big"), int.from_bytes(q[144:192], "big")
    return q[:96] + ((P - yb) % P).to_bytes(48, "big") + ((P - ya) % P).to_bytes(48, "big")


class PublicParameters:
    """parse<G1, G2, G1>(pp.g1_g2_h0), parse<G1>(pp.h) done once (the reference's lazy DSL re-decompresses h[i] for every
    signature): affine forms of g1, g2, h0 and h."""

    def __init__(self, g1_g2_h0: bytes, h: Sequence[bytes]):
        if len(g1_g2_h0) != G1_BYTES + G2_BYTES + G1_BYTES:
            raise ValueError("pp.g1_g2_h0 must be 195 bytes")
        g1s = bridge.from_bytes(g1_g2_h0[:G1_BYTES] + g1_g2_h0[G1_BYTES + G2_BYTES:] + b"".join(h))
        self.g1, self.h0, self.h = g1s[:96], g1s[96:192], g1s[192:]
        self.g2 = bridge.from_bytes2(g1_g2_h0[G1_BYTES:G1_BYTES + G2_BYTES])
        self.n_max = len(h)


def _message_scalars(pp: PublicParameters, messages: Sequence[bytes]):
    blocks = [encode_to_zp(m) for m in messages]
    n = max((len(b) for b in blocks), default=0)
    if n > pp.n_max:
        raise RuntimeError("message is too long")     # bbs+.cpp:47-50,66-69
    return blocks, n


def g1_products(pp: PublicParameters, rs: Sequence[int], blocks: Sequence[Sequence[int]], n: int) -> bytes:
    """B_i = g1 * h0^r_i * Π[n](h[j]^m_ij) for every i (affine 96 B each); shorter messages are padded with zero exponents."""
    bases = pp.g1 + pp.h0 + pp.h[:96 * n]
    values = b"".join(be32(1) + be32(r) + b"".join(be32(m) for m in ms) + bytes(32 * (n - len(ms))) for r, ms in zip(rs, blocks))
    return bridge.products_over_bases(bases, values)


def sign_batch(pp: PublicParameters, sk: bytes, messages: Sequence[bytes], xs: Sequence[int], rs: Sequence[int]) -> List[bytes]:
    """sign (bbs+.cpp:38-55) with the caller's randomness (x_i, r_i): A = (g1 * h0^r * Π h[i]^m[i])^(1 / (γ + x))."""
    gamma = _zp(sk)
    blocks, n = _message_scalars(pp, messages)
    Bs = g1_products(pp, rs, blocks, n)
    exps = b"".join(be32(pow((gamma + x) % R, -1, R)) for x in xs)
    As = bridge.multiply(Bs, exps)      # compressed 49 B each
    return [As[49 * i:49 * i + 49] + int(x).to_bytes(48, "big") + int(r).to_bytes(48, "big") for i, (x, r) in enumerate(zip(xs, rs))]


def verify_batch(pp: PublicParameters, pk: bytes, messages: Sequence[bytes], signatures: Sequence[bytes]) -> List[bool]:
    """verify (bbs+.cpp:57-73) for every (message, signature) pair under one pp / pk.  A signature whose encoding does not
    parse (the reference throws from parse<G1, Zp, Zp>) is reported as False."""
    if len(messages) != len(signatures):
        raise ValueError("messages and signatures differ in length")
    B = len(signatures)
    if B == 0:
        return []
    w = bridge.from_bytes2(pk)
    blocks, n = _message_scalars(pp, messages)
    ok = [len(s) == SIG_BYTES for s in signatures]
    xs, rs, enc = [], [], bytearray()
    for i, s in enumerate(signatures):
        try:
            x, r = (_zp(s[49:97]), _zp(s[97:145])) if ok[i] else (0, 0)
        except ValueError:
            ok[i], x, r = False, 0, 0
        xs.append(x)
        rs.append(r)
        enc += s[:49] if ok[i] else bytes(49)
    try:
        As = bridge.from_bytes(bytes(enc))
    except bridge._lib.C12381Error:
        # at least one A is not a curve point: find them one by one (rare path), keep the rest
        As = bytearray()
        for i in range(B):
            try:
                As += bridge.from_bytes(bytes(enc[49 * i:49 * i + 49]))
            except bridge._lib.C12381Error:
                ok[i] = False
                As += bytes(96)
        As = bytes(As)
    Ws = bridge.products_over_bases2(w + pp.g2, b"".join(be32(1) + be32(x) for x in xs))
    Bs = g1_products(pp, rs, blocks, n)
    ng2 = _neg_g2(pp.g2)
    g1s = b"".join(As[96 * i:96 * i + 96] + Bs[96 * i:96 * i + 96] for i in range(B))
    g2s = b"".join(Ws[192 * i:192 * i + 192] + ng2 for i in range(B))
    verdict = bridge.pairing_check_batch(g1s, g2s, 2)
    return [bool(v) and o for v, o in zip(verdict, ok)]


def verify_batch_aggregate(pp: PublicParameters, pk: bytes, messages: Sequence[bytes], signatures: Sequence[bytes], seed: bytes) -> bool:
    """ONE verdict for the whole batch by a random linear combination - an extension, not something the reference has.
    With W_i = w + x_i g2 every valid signature satisfies e(A_i, w) e(x_i A_i - B_i, g2) = 1, hence for 128-bit weights ρ_i
        e(Σ ρ_i A_i, w) · e(Σ ρ_i x_i A_i - Σ ρ_i B_i, g2) = 1
    and a batch with any invalid signature passes with probability 2^-128.  Cost: the B per-signature products B_i, two
    G1 sums of B and 2B terms (the MSM pipeline) and a single 2-pair pairing check, instead of B pairing checks.
    `seed` must be unpredictable to whoever produced the signatures (the weights are derived from it with SHA3-512)."""
    import hashlib
    B = len(signatures)
    if len(messages) != B:
        raise ValueError("messages and signatures differ in length")
    if B == 0:
        return True
    if any(len(s) != SIG_BYTES for s in signatures):
        return False
    try:
        xs = [_zp(s[49:97]) for s in signatures]
        rs = [_zp(s[97:145]) for s in signatures]
        As = bridge.from_bytes(b"".join(s[:49] for s in signatures))
        w = bridge.from_bytes2(pk)
    except (ValueError, bridge._lib.C12381Error):
        return False
    if bytes(96) in (As[96 * i:96 * i + 96] for i in range(B)):
        return False                                   # A = identity never verifies (e(O, .) = 1 would hide B_i != O)
    blocks, n = _message_scalars(pp, messages)
    Bs = g1_products(pp, rs, blocks, n)
    rho = [int.from_bytes(hashlib.sha3_512(seed + i.to_bytes(8, "big")).digest()[:16], "big") | 1 for i in range(B)]
    s1 = bridge.from_bytes(bridge.sum_of_products(As, b"".join(be32(r) for r in rho)))
    s2 = bridge.from_bytes(bridge.sum_of_products(As + Bs, b"".join(be32(r * x % R) for r, x in zip(rho, xs)) + b"".join(be32(R - r) for r in rho)))
    return bridge.pairing_check_batch(s1 + s2, w + pp.g2, 2) == b"\x01"


# ---- the same pipeline on CUDA-resident tensors (what bench.py times) -----------------------------------------------
def verify_batch_device(bases_g1, bases_g2, neg_g2, sig_A, scalars_g1, scalars_g2):
    """bases_g1: (2 + n) x 96 B (g1, h0, h_j); bases_g2: 2 x 192 B (w, g2); neg_g2: 192 B; sig_A: B x 49 B compressed;
    scalars_g1: B x (2 + n) x 32 B rows (1, r_i, m_ij); scalars_g2: B x 2 x 32 B rows (1, x_i).  All uint8 CUDA tensors.
    Returns the B verdict bytes (a CUDA tensor); malformed input is reported by device.sync_status()."""
    import torch

    from . import device as dv
    B = sig_A.numel() // 49
    A = dv.g1_decompress_batch(sig_A)
    W = dv.g2_multi_fixed_base_batch(bases_g2, scalars_g2)
    Bp = dv.g1_multi_fixed_base_batch(bases_g1, scalars_g1)
    g1s = torch.stack((A.view(B, 96), Bp.view(B, 96)), dim=1).reshape(-1)
    g2s = torch.cat((W.view(B, 192), neg_g2.view(1, 192).expand(B, 192)), dim=1).reshape(-1)
    return dv.pairing_check_batch(g1s, g2s, 2)


def verify_batch_aggregate_device(bases_g1, bases_g2, sig_A, scalars_g1, rho, rho_x, neg_rho):
    """The aggregate check on CUDA-resident tensors: bases_g2 = (w, g2) 2 x 192 B; rho, rho_x = ρ_i x_i mod r, neg_rho = r - ρ_i
    as B x 32 B each (Zp arithmetic stays on the host, as everywhere).  Returns the single verdict byte (a CUDA tensor)."""
    import torch

    from . import device as dv
    A = dv.g1_decompress_batch(sig_A)
    Bp = dv.g1_multi_fixed_base_batch(bases_g1, scalars_g1)
    s1 = dv.g1_decompress_batch(dv.g1_msm(A, rho))
    s2 = dv.g1_decompress_batch(dv.g1_msm(torch.cat((A, Bp)), torch.cat((rho_x, neg_rho))))
    return dv.pairing_check_batch(torch.cat((s1, s2)), bases_g2, 2)
