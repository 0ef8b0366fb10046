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

} // namespace c12
