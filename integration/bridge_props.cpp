// Bridge-level restatement of the reference's unit tests for the hot path, run against
// build/libcrypto12381_b200.so: the reference bridge with its hot functions replaced by the B200 library.
// Two kinds of checks on identical seeded inputs (create_random_engine + random_in, the reference's own generator):
//   (1) differential: every replaced function against the reference's own MIRACL definition of the same function
//       (kept linkable as refcpu_<mangled>, see integration/Makefile), compared with the bridge's equal();
//   (2) the algebraic properties the reference tests state (unit-tests/g1_point.cpp:51-98, g2_point.cpp:51-78,
//       set.cpp:257-265, liner_pair.cpp:28-103,129-160), evaluated purely through the replaced functions.
// Exit code 0 = all passed.  Needs a CUDA device.
#include <cstdio>
#include <cstring>
#include <string_view>
#include <vector>

#include <crypto12381/miracl_core_interface.hpp>
#include <crypto12381/random.hpp>

using namespace crypto12381;
using namespace crypto12381::detail::miracl_core;

// ABI-additive entry of the replacement bridge (integration/miracl_core_interface_b200.cpp) - not in the reference header
namespace crypto12381::detail::miracl_core
{
    void sum_of_products(point2& result, int n, point2* points, const big* numbers) noexcept;
    void pair_multi_ate(fp12& result, int n, point2* p2s, point1* p1s) noexcept;
}

// the reference's MIRACL-backed definitions, renamed by objcopy
namespace refcpu
{
    void sum_of_products(point1&, int, point1*, const big*) noexcept asm("refcpu__ZN11crypto123816detail11miracl_core15sum_of_productsERNS1_6point1EiPS2_PA7_Kl");
    void double_multiply(point1&, point1&, big&, big&) noexcept asm("refcpu__ZN11crypto123816detail11miracl_core15double_multiplyERNS1_6point1ES3_RA7_lS5_");
    void multiply(point1&, const big&) noexcept asm("refcpu__ZN11crypto123816detail11miracl_core8multiplyERNS1_6point1ERA7_Kl");
    void multiply(point2&, const big&) noexcept asm("refcpu__ZN11crypto123816detail11miracl_core8multiplyERNS1_6point2ERA7_Kl");
    void multiply(fp12&, fp12&) noexcept asm("refcpu__ZN11crypto123816detail11miracl_core8multiplyERNS1_4fp12ES3_");
    void pow(fp12&, fp12&, const big&) noexcept asm("refcpu__ZN11crypto123816detail11miracl_core3powERNS1_4fp12ES3_RA7_Kl");
    void pair_ate(fp12&, point2&, point1&) noexcept asm("refcpu__ZN11crypto123816detail11miracl_core8pair_ateERNS1_4fp12ERNS1_6point2ERNS1_6point1E");
    void pair_final_exponentiation(fp12&) noexcept asm("refcpu__ZN11crypto123816detail11miracl_core25pair_final_exponentiationERNS1_4fp12E");
    void pair_double_ate(fp12&, point2&, point1&, point2&, point1&) noexcept asm("refcpu__ZN11crypto123816detail11miracl_core15pair_double_ateERNS1_4fp12ERNS1_6point2ERNS1_6point1ES5_S7_");
}

namespace
{
    int failures = 0, checks = 0;
#define CHECK(cond)                                                               \
    do {                                                                          \
        ++checks;                                                                 \
        if (!(cond)) { ++failures; std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond); } \
    } while (0)

    // group order r as the bridge's 58-bit digits (include/crypto12381/zp_number.hpp:146-148)
    const big order = {0x3FFFFFF00000001L, 0x36900BFFF96FFBFL, 0x180809A1D80553BL, 0x14CA675F520CCE7L, 0x73EDA7L, 0x0L, 0x0L};

    void random_scalar(big& out, RandomEngine& random) { random_in(out, order, random); }

    void mul_mod_r(big& out, const big& a, const big& b)
    {
        big2 wide;
        multiply(wide, a, b);   // the big2 overload: Zp plumbing, stays on the CPU
        mod(out, wide, order);
    }
    void add_mod_r(big& out, const big& a, const big& b)
    {
        big2 wide{};
        for (int i = 0; i < 7; ++i) wide[i] = a[i] + b[i];
        big2 t;
        std::memcpy(&t, &wide, sizeof t);
        mod(out, t, order);
    }

    void random_g1(point1& p, RandomEngine& random)   // generator ^ random, as select_in<*G1> does (g1_point.hpp:355-369)
    {
        big k;
        random_scalar(k, random);
        get_default_generator(p);
        refcpu::multiply(p, k);
    }
    void random_g2(point2& p, RandomEngine& random)
    {
        big k;
        random_scalar(k, random);
        get_default_generator(p);
        refcpu::multiply(p, k);
    }
    bool same(point1 a, point1 b) { return equal(a, b) == 1; }
    bool same(point2 a, point2 b) { return equal(a, b) == 1; }
    bool same(fp12 a, fp12 b) { return equal(a, b) == 1; }
}

int main()
{
    using namespace std::string_view_literals;
    auto random = create_random_engine("b200 bridge properties seed"sv);

    // ---- G1: multiply / double_multiply / sum_of_products ------------------------------------------------------
    for (int iter = 0; iter < 4; ++iter)
    {
        point1 P, Q;
        random_g1(P, random);
        random_g1(Q, random);
        big x, y;
        random_scalar(x, random);
        random_scalar(y, random);

        point1 a = P, b = P;
        multiply(a, x);
        refcpu::multiply(b, x);
        CHECK(same(a, b));                                   // differential: PAIR_G1mul

        point1 c = P, c2 = Q, d = P, d2 = Q;
        big x1, y1, x2, y2;
        std::memcpy(x1, x, sizeof x); std::memcpy(y1, y, sizeof y); std::memcpy(x2, x, sizeof x); std::memcpy(y2, y, sizeof y);
        double_multiply(c, c2, x1, y1);
        refcpu::double_multiply(d, d2, x2, y2);
        CHECK(same(c, d));                                   // differential: ECP_mul2

        // (P^x)*(Q^y) equals the separately computed products (unit-tests/g1_point.cpp:80-98)
        point1 px = P, qy = Q;
        multiply(px, x);
        multiply(qy, y);
        add(px, qy);
        CHECK(same(c, px));

        // P^(x+y) == P^x * P^y and P^(x*y) == (P^x)^y (unit-tests/g1_point.cpp:51-78)
        big s, m;
        add_mod_r(s, x, y);
        mul_mod_r(m, x, y);
        point1 ps = P, pm = P, pxy = P, pxpy = P, py = P;
        multiply(ps, s);
        multiply(pm, m);
        multiply(pxy, x);
        multiply(pxy, y);
        multiply(pxpy, x);
        multiply(py, y);
        add(pxpy, py);
        CHECK(same(ps, pxpy));
        CHECK(same(pm, pxy));
    }
    {
        // Π over lazy powers (the MSM entry): differential against ECP_muln, n = 1, 3, 33, 257; and
        // product(P^x, P^y, P^z) == P^(x+y+z) (unit-tests/set.cpp:257-265)
        for (int n : {1, 3, 33, 257})
        {
            std::vector<point1> pts(n);
            std::vector<big> nums(n);
            for (int i = 0; i < n; ++i)
            {
                random_g1(pts[i], random);
                random_scalar(nums[i], random);
            }
            point1 r1, r2;
            sum_of_products(r1, n, pts.data(), reinterpret_cast<const big*>(nums.data()));
            refcpu::sum_of_products(r2, n, pts.data(), reinterpret_cast<const big*>(nums.data()));
            CHECK(same(r1, r2));
        }
        point1 P;
        random_g1(P, random);
        big x, y, z, s;
        random_scalar(x, random);
        random_scalar(y, random);
        random_scalar(z, random);
        add_mod_r(s, x, y);
        add_mod_r(s, s, z);
        point1 pts[3] = {P, P, P};
        big nums[3];
        std::memcpy(nums[0], x, sizeof x); std::memcpy(nums[1], y, sizeof y); std::memcpy(nums[2], z, sizeof z);
        point1 r, e = P;
        sum_of_products(r, 3, pts, nums);
        multiply(e, s);
        CHECK(same(r, e));
        // empty product and identity inputs
        point1 inf, r0;
        get_infinity(inf);
        sum_of_products(r0, 0, pts, nums);
        CHECK(is_infinity(r0));
        point1 i2 = inf;
        multiply(i2, x);
        CHECK(is_infinity(i2));
    }

    // ---- G2 ------------------------------------------------------------------------------------------------------
    for (int iter = 0; iter < 3; ++iter)
    {
        point2 P;
        random_g2(P, random);
        big x, y, s;
        random_scalar(x, random);
        random_scalar(y, random);
        point2 a = P, b = P;
        multiply(a, x);
        refcpu::multiply(b, x);
        CHECK(same(a, b));                                   // differential: PAIR_G2mul
        add_mod_r(s, x, y);
        point2 ps = P, py = P;
        multiply(ps, s);
        multiply(py, y);
        add(a, py);
        CHECK(same(ps, a));                                  // P^(x+y) == P^x * P^y (unit-tests/g2_point.cpp:51-78)
    }

    // ---- the additive G2 seam: sum_of_products(point2&) == the reference's per-term multiply + add loop ------
    {
        const int n = 37;
        std::vector<point2> pts(n);
        std::vector<big> nums(n);
        point2 want;
        get_infinity(want);
        for (int i = 0; i < n; ++i)
        {
            random_g2(pts[i], random);
            random_scalar(nums[i], random);
            point2 t = pts[i];
            refcpu::multiply(t, nums[i]);                    // g2_point.hpp:202-217 as the reference runs it
            add(want, t);                                    // :225-236
        }
        point2 got;
        sum_of_products(got, n, pts.data(), reinterpret_cast<const big*>(nums.data()));
        CHECK(same(got, want));
    }

    // ---- the additive n-pairing seam: one Miller product == the reference's pair_ate values multiplied together -----
    for (int n : {1, 3, 4, 7})
    {
        std::vector<point1> p1s(n);
        std::vector<point2> p2s(n);
        fp12 want;
        for (int i = 0; i < n; ++i)
        {
            random_g1(p1s[i], random);
            random_g2(p2s[i], random);
            point1 p = p1s[i];
            point2 q = p2s[i];
            fp12 m;
            refcpu::pair_ate(m, q, p);
            if (i == 0) want = m; else refcpu::multiply(want, m);
        }
        fp12 got;
        pair_multi_ate(got, n, p2s.data(), p1s.data());
        refcpu::pair_final_exponentiation(want);
        refcpu::pair_final_exponentiation(got);
        CHECK(same(got, want));
    }

    // ---- pairings --------------------------------------------------------------------------------------------------
    {
        point1 P, P2;
        point2 Q, Q2;
        random_g1(P, random);
        random_g1(P2, random);
        random_g2(Q, random);
        random_g2(Q2, random);
        big x, y, xy;
        random_scalar(x, random);
        random_scalar(y, random);
        mul_mod_r(xy, x, y);

        fp12 m1, m2;
        { point1 p = P; point2 q = Q; pair_ate(m1, q, p); }
        { point1 p = P; point2 q = Q; refcpu::pair_ate(m2, q, p); }
        CHECK(same(m1, m2));                                 // differential: PAIR_ate (raw Miller value)
        fp12 e1 = m1, e2 = m2;
        pair_final_exponentiation(e1);
        refcpu::pair_final_exponentiation(e2);
        CHECK(same(e1, e2));                                 // differential: PAIR_fexp
        CHECK(!is_unity(e1));                                // non-degeneracy (unit-tests/liner_pair.cpp:28-40)

        fp12 d1, d2;
        { point1 p = P, p2 = P2; point2 q = Q, q2 = Q2; pair_double_ate(d1, q, p, q2, p2); }
        { point1 p = P, p2 = P2; point2 q = Q, q2 = Q2; refcpu::pair_double_ate(d2, q, p, q2, p2); }
        CHECK(same(d1, d2));                                 // differential: PAIR_double_ate

        // pair*pair equals the product of two independent pairings (unit-tests/liner_pair.cpp:66-79)
        fp12 other;
        { point1 p2 = P2; point2 q2 = Q2; pair_ate(other, q2, p2); }
        pair_final_exponentiation(other);
        fp12 prod = e1;
        multiply(prod, other);
        fp12 dd = d1;
        pair_final_exponentiation(dd);
        CHECK(same(prod, dd));
        { fp12 t = e2; refcpu::multiply(t, other); CHECK(same(prod, t)); }   // differential: FP12_mul

        // bilinearity pair(P^x, Q^y) == pair(P, Q)^(x*y) (unit-tests/liner_pair.cpp:42-64)
        point1 px = P;
        point2 qy = Q;
        multiply(px, x);
        multiply(qy, y);
        fp12 lhs, rhs, rhs2;
        pair_ate(lhs, qy, px);
        pair_final_exponentiation(lhs);
        pow(rhs, e1, xy);
        refcpu::pow(rhs2, e2, xy);
        CHECK(same(rhs, rhs2));                              // differential: FP12_pow
        CHECK(same(lhs, rhs));

        // e(O, Q) = e(P, O) = 1 (unit-tests/liner_pair.cpp:28-40)
        point1 inf1;
        point2 inf2;
        get_infinity(inf1);
        get_infinity(inf2);
        fp12 u;
        { point2 q = Q; pair_ate(u, q, inf1); }
        pair_final_exponentiation(u);
        CHECK(is_unity(u));
        { point1 p = P; pair_ate(u, inf2, p); }
        pair_final_exponentiation(u);
        CHECK(is_unity(u));

        // GT exponent laws (unit-tests/liner_pair.cpp:129-160): (g^x)^y == g^(xy), g^x * g^y == g^(x+y)
        fp12 gx, gxy, gy, gs;
        big s;
        add_mod_r(s, x, y);
        pow(gx, e1, x);
        pow(gxy, gx, y);
        CHECK(same(gxy, rhs));
        pow(gy, e1, y);
        pow(gs, e1, s);
        multiply(gx, gy);
        CHECK(same(gx, gs));
    }

    std::printf("%d checks, %d failures\n", checks, failures);
    return failures ? 1 : 0;
}
