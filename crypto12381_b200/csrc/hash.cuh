// SHA3-512 and hashing to Zp, host+device bodies.
// Replaces, for batches, the byte-at-a-time SHA3_init / SHA3_process / SHA3_hash of the bridge
// (src/miracl_core_interface.cpp:12-25 -> 3rd-party/miracl-core/hash.cpp:392-554) as driven by hash_state
// (include/crypto12381/set.hpp:317-392), and Zp's from_hash (zp_number.hpp:538-547): the 64-byte digest read as a
// big-endian 512-bit integer, reduced mod r (BIG_dfromBytesLen + BIG_ctdmod).
#pragma once
#include "msm_core.cuh"

namespace c12 {

C12_HD uint64_t rotl64(uint64_t x, int n) { return (x << n) | (x >> (64 - n)); }

// Keccak-f[1600], 24 rounds (FIPS 202)
C12_HD void keccak_f1600(uint64_t (&s)[25])
{
    const uint64_t RC[24] = {0x0000000000000001ull, 0x0000000000008082ull, 0x800000000000808aull, 0x8000000080008000ull, 0x000000000000808bull,
                             0x0000000080000001ull, 0x8000000080008081ull, 0x8000000000008009ull, 0x000000000000008aull, 0x0000000000000088ull,
                             0x0000000080008009ull, 0x000000008000000aull, 0x000000008000808bull, 0x800000000000008bull, 0x8000000000008089ull,
                             0x8000000000008003ull, 0x8000000000008002ull, 0x8000000000000080ull, 0x000000000000800aull, 0x800000008000000aull,
                             0x8000000080008081ull, 0x8000000000008080ull, 0x0000000080000001ull, 0x8000000080008008ull};
    const int ROT[25] = {0, 1, 62, 28, 27, 36, 44, 6, 55, 20, 3, 10, 43, 25, 39, 41, 45, 15, 21, 8, 18, 2, 61, 56, 14};
    for (int round = 0; round < 24; ++round) {
        uint64_t c[5], d[5], b[25];
        for (int x = 0; x < 5; ++x) c[x] = s[x] ^ s[x + 5] ^ s[x + 10] ^ s[x + 15] ^ s[x + 20];
        for (int x = 0; x < 5; ++x) d[x] = c[(x + 4) % 5] ^ rotl64(c[(x + 1) % 5], 1);
        for (int i = 0; i < 25; ++i) s[i] ^= d[i % 5];
        for (int x = 0; x < 5; ++x)
            for (int y = 0; y < 5; ++y) {
                const int i = x + 5 * y;
                const uint64_t v = ROT[i] ? rotl64(s[i], ROT[i]) : s[i];
                b[y + 5 * ((2 * x + 3 * y) % 5)] = v;       // rho + pi
            }
        for (int y = 0; y < 5; ++y)
            for (int x = 0; x < 5; ++x) s[x + 5 * y] = b[x + 5 * y] ^ (~b[(x + 1) % 5 + 5 * y] & b[(x + 2) % 5 + 5 * y]);
        s[0] ^= RC[round];
    }
}

// SHA3-512(msg): rate 72 bytes, domain byte 0x06, final bit 0x80
C12_HD void sha3_512(const uint8_t* msg, size_t len, uint8_t out[64])
{
    uint64_t s[25];
    for (int i = 0; i < 25; ++i) s[i] = 0;
    const size_t rate = 72;
    size_t off = 0;
    while (len - off >= rate) {
        for (size_t i = 0; i < rate; ++i) s[i >> 3] ^= (uint64_t)msg[off + i] << (8 * (i & 7));
        keccak_f1600(s);
        off += rate;
    }
    const size_t rest = len - off;
    for (size_t i = 0; i < rest; ++i) s[i >> 3] ^= (uint64_t)msg[off + i] << (8 * (i & 7));
    s[rest >> 3] ^= (uint64_t)0x06 << (8 * (rest & 7));
    s[(rate - 1) >> 3] ^= (uint64_t)0x80 << (8 * ((rate - 1) & 7));
    keccak_f1600(s);
    for (int i = 0; i < 64; ++i) out[i] = (uint8_t)(s[i >> 3] >> (8 * (i & 7)));
}

// big-endian 512-bit digest mod r -> 32 bytes big-endian (bit-serial: acc = 2 acc + bit, minus r when it reaches r)
C12_HD void digest_mod_r(const uint8_t d[64], uint8_t out32[32])
{
    const uint32_t r[8] = C12_R_LIMBS;
    uint32_t acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int i = 0; i < 512; ++i) {
        const uint32_t bit = (d[i >> 3] >> (7 - (i & 7))) & 1u;
        uint32_t top = acc[7] >> 31;
        for (int k = 7; k > 0; --k) acc[k] = (acc[k] << 1) | (acc[k - 1] >> 31);
        acc[0] = (acc[0] << 1) | bit;
        // acc < 2 r < 2^256 whenever the previous acc < r, so `top` is always 0; kept for clarity of the bound
        uint32_t t[8];
        uint64_t borrow = 0;
        for (int k = 0; k < 8; ++k) {
            uint64_t x = (uint64_t)acc[k] - r[k] - borrow;
            t[k] = (uint32_t)x;
            borrow = (x >> 32) & 1u;
        }
        const bool ge = top || !borrow;
        for (int k = 0; k < 8; ++k) acc[k] = ge ? t[k] : acc[k];
    }
    for (int k = 0; k < 8; ++k) {
        out32[28 - 4 * k] = (uint8_t)(acc[k] >> 24);
        out32[29 - 4 * k] = (uint8_t)(acc[k] >> 16);
        out32[30 - 4 * k] = (uint8_t)(acc[k] >> 8);
        out32[31 - 4 * k] = (uint8_t)acc[k];
    }
}

// hash(message) -> Zp: what `hash(...) -> Zp` / hash_state::to(Zp) yields for the serialised bytes `msg`
C12_HD void hash_to_zp_body(const uint8_t* msg, size_t len, uint8_t out32[32])
{
    uint8_t d[64];
    sha3_512(msg, len, d);
    digest_mod_r(d, out32);
}

// ---- hash to G1: G1Point::from_hash (include/crypto12381/g1_point.hpp:219-234) ---------------------------------
// digest -> big2 -> fixed_time_mod p -> residue -> map_to_point (ECP_map2point, 3rd-party/miracl-core/ecp_BLS12381.cpp:1276,
// 1493-1627: simplified SWU on the 11-isogenous curve E', Z = 11, then the isogeny, projective result) -> multiply_cofactor
// (ECP_cfp :1252-1273, times CURVE_Cof = 1 - x).  The reference's constant-time selections compute the RFC 9380 map with
// sgn0 = parity (FP_sign); like the reference, ONE exponentiation w^((p-3)/4) yields the Legendre symbol, the inverse and the
// square root (FP_qr / FP_inv / FP_sqrt sharing `hint`, :1537-1556).

// a^((p-3)/4), 4-bit windows
C12_HD_NOINLINE Fp fp_pow_pm3d4(const Fp& a)
{
    const uint32_t e[12] = C12_PM3D4_LIMBS;
    Fp tab[16];
    tab[0] = fp_one();
    tab[1] = a;
#pragma unroll 1
    for (int i = 2; i < 16; ++i) tab[i] = fp_mul(tab[i - 1], a);
    Fp r = fp_one();
#pragma unroll 1
    for (int i = 95; i >= 0; --i) {
        if (i != 95) {
            r = fp_sqr(r);
            r = fp_sqr(r);
            r = fp_sqr(r);
            r = fp_sqr(r);
        }
        uint32_t d = (e[i >> 3] >> ((i & 7) * 4)) & 15u;
        if (d) r = fp_mul(r, tab[d]);
    }
    return r;
}

// the 64-byte big-endian digest mod p, Montgomery form: lo (384 bits) + hi (128 bits) 2^384
C12_HD Fp digest_mod_p(const uint8_t d[64])
{
    Fp lo = fp_from_be48(d + 16), hi = fp_zero();
#pragma unroll
    for (int k = 0; k < 4; ++k)
        hi.v[k] = ((uint32_t)d[12 - 4 * k] << 24) | ((uint32_t)d[13 - 4 * k] << 16) | ((uint32_t)d[14 - 4 * k] << 8) | (uint32_t)d[15 - 4 * k];
    // fp_to_mont accepts any 384-bit value: to_mont(hi) = hi 2^384 mod p as a plain value, once more for its Montgomery form
    return fp_add(fp_to_mont(lo), fp_to_mont(fp_to_mont(hi)));
}

C12_HD Fp iso_horner(const uint32_t (*c)[12], int n, const Fp& x, bool monic)
{
    Fp acc = monic ? fp_one() : fp_zero();
#pragma unroll 1
    for (int i = 0; i < n; ++i) {
        Fp k;
#pragma unroll
        for (int j = 0; j < 12; ++j) k.v[j] = c[i][j];
        acc = (i == 0 && !monic) ? k : fp_add(fp_mul(acc, x), k);
    }
    return acc;
}

// map_to_point + multiply_cofactor for a field element u (Montgomery form); the identity when Z^2 u^4 + Z u^2 = 0, as the
// reference yields there (its FP_inv(0) = 0 zeroes the projective Z; checked against the compiled reference)
C12_HD_NOINLINE Proj<Fp> map_to_g1(const Fp& u)
{
    const Fp A = sswu_a_m(), B = sswu_b_m();
    const int sgn = fp_sign(u);
    const Fp t = fp_mul(fp_sqr(u), sswu_z_m());          // Z u^2
    const Fp w = fp_add(fp_sqr(t), t);                   // Z^2 u^4 + Z u^2
    if (fp_is_zero(w)) return proj_inf<Fp>();
    // x1 = -B (w + 1) / (A w) = N / D;  g(x1) = (N^3 + A N D^2 + B D^3) / D^3 = G / D^3
    const Fp N = fp_neg(fp_mul(B, fp_add(w, fp_one())));
    const Fp D = fp_mul(A, w);
    const Fp D2 = fp_sqr(D);
    const Fp G = fp_add(fp_mul(N, fp_add(fp_sqr(N), fp_mul(A, D2))), fp_mul(B, fp_mul(D2, D)));
    const Fp GD = fp_mul(G, D);
    const Fp h = fp_pow_pm3d4(GD);
    const Fp c = fp_mul(GD, h);                          // GD^((p+1)/4): c^2 = +-GD
    const bool qr = fp_eq(fp_sqr(c), GD);
    const Fp invD = fp_mul(fp_mul(fp_sqr(fp_sqr(h)), GD), G);   // h^4 GD = 1 / GD;  times G
    const Fp x1 = fp_mul(N, invD);
    const Fp cd2 = fp_mul(c, fp_sqr(invD));              // qr: sqrt(g(x1)) = c / D^2
    // non-residue: x2 = Z u^2 x1, g(x2) = Z^3 u^6 g(x1) = (-Z^3) u^6 (-GD) / D^4, and c^2 = -GD
    const Fp x = qr ? x1 : fp_mul(t, x1);
    Fp y = qr ? cd2 : fp_mul(fp_mul(cd2, sswu_sqrt_mz3_m()), fp_mul(fp_sqr(u), u));
    if (fp_sign(y) != sgn) y = fp_neg(y);
    // 11-isogeny E' -> E, Horner in the affine x; projective result (:1568-1627)
    const uint32_t xn[C12_ISO_XNUM_N][12] = C12_ISO_XNUM_M;
    const uint32_t xd[C12_ISO_XDEN_N][12] = C12_ISO_XDEN_M;
    const uint32_t yn[C12_ISO_YNUM_N][12] = C12_ISO_YNUM_M;
    const uint32_t yd[C12_ISO_YDEN_N][12] = C12_ISO_YDEN_M;
    const Fp xnum = iso_horner(xn, C12_ISO_XNUM_N, x, false), xden = iso_horner(xd, C12_ISO_XDEN_N, x, true);
    const Fp ynum = fp_mul(iso_horner(yn, C12_ISO_YNUM_N, x, false), y), yden = iso_horner(yd, C12_ISO_YDEN_N, x, true);
    const Proj<Fp> P = Proj<Fp>{fp_mul(xnum, yden), fp_mul(ynum, xden), fp_mul(xden, yden)};
    // times 1 - x = 0xd201000000010001 (double-and-add on the complete formulas)
    const uint64_t k = C12_H_EFF;
    Proj<Fp> r = P;
#pragma unroll 1
    for (int i = 62; i >= 0; --i) {
        r = proj_dbl(r);
        if ((k >> i) & 1ull) r = proj_add(r, P);
    }
    return r;
}

// hash(message) -> G1, compressed 49 bytes
C12_HD void hash_to_g1_body(const uint8_t* msg, size_t len, uint8_t out49[49])
{
    uint8_t d[64];
    sha3_512(msg, len, d);
    Wire<Fp>::compress(out49, proj_to_affine(map_to_g1(digest_mod_p(d))));
}
// map_to_point + multiply_cofactor of a field element given as 48 bytes big-endian (< p)
C12_HD bool map_to_g1_body(const uint8_t* u48, uint8_t out49[49])
{
    const Fp u = fp_from_be48(u48);
    const bool ok = fp_is_canonical(u);
    Wire<Fp>::compress(out49, proj_to_affine(map_to_g1(fp_to_mont(u))));
    return ok;
}

} // namespace c12
