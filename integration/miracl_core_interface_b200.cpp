// Replacement definitions of the HOT bridge functions of crypto12381 (namespace crypto12381::detail::miracl_core,
// declared in <crypto12381/miracl_core_interface.hpp>): identical signatures, arguments passed through unchanged to
// the B200 library (include/c12381_cuda.h, `_miracl` entries).  Everything else in the bridge keeps its reference
// definition: integration/Makefile links the reference's own src/miracl_core_interface.cpp object with these ten
// symbols renamed out of the way (objcopy --redefine-sym), so no reference source is copied or edited and the C++
// surface (Zp, G1, G2, GT, operator^, pair, Π, serialize/parse) compiles against the very same header.
//
// The bridge is `noexcept` with no status for these functions (miracl_core_interface.hpp:102,122,125,151,189-203),
// and the north star forbids a CPU fallback, so a CUDA failure terminates with the library's message.
#include <cstdio>
#include <cstdlib>

#include <crypto12381/miracl_core_interface.hpp>

#include "c12381_cuda.h"

namespace
{
    void ensure_context() noexcept
    {
        if (c12381_device() >= 0) return;
        const char* dev = std::getenv("C12381_DEVICE");
        if (int rc = c12381_init(dev ? std::atoi(dev) : 0); rc != C12381_OK)
        {
            std::fprintf(stderr, "crypto12381-b200: %s (code %d)\n", c12381_last_error(), rc);
            std::abort();
        }
    }

    void must(int rc, const char* what) noexcept
    {
        if (rc == C12381_OK) return;
        std::fprintf(stderr, "crypto12381-b200: %s failed: %s (code %d)\n", what, c12381_last_error(), rc);
        std::abort();
    }
}

namespace crypto12381::detail::miracl_core
{
    // Π[n](h[i]^m[i]) seam (g1_point.hpp:371-404; reference body -> ECP_muln)
    void sum_of_products(point1& result, int n, point1* points, const big* numbers) noexcept
    {
        ensure_context();
        must(c12381_sum_of_products_miracl(&result, n, points, numbers), "sum_of_products");
    }

    // ABI-ADDITIVE: the G2 counterpart of the seam above, not declared by the reference header -
    // what a lazy G2Pow / Π over G2 (g2_point.hpp:202-236, today n x PAIR_G2mul + n x ECP2_add) would call.  One G2 MSM launch.
    void sum_of_products(point2& result, int n, point2* points, const big* numbers) noexcept
    {
        ensure_context();
        must(c12381_sum_of_products2_miracl(&result, n, points, numbers), "sum_of_products(point2)");
    }

    // ABI-ADDITIVE: the Miller value of a whole `pair * pair * ...` chain (liner_pair.hpp:219-230,291-303), n <= 8 pairs
    // with shared squarings - what the DSL folds two at a time through pair_double_ate and multiply(fp12&, fp12&) today.
    void pair_multi_ate(fp12& result, int n, point2* p2s, point1* p1s) noexcept
    {
        ensure_context();
        must(c12381_pair_multi_ate_miracl(&result, n, p2s, p1s), "pair_multi_ate");
    }

    // G1Pow -> G1Point (g1_point.hpp:296-310), select g^x (:355-369); reference body -> PAIR_G1mul
    void multiply(point1& object, const big& value) noexcept
    {
        ensure_context();
        must(c12381_multiply_point1_miracl(&object, &value), "multiply(point1)");
    }

    // G1Pow * G1Pow (g1_point.hpp:317-353): p1 = v1*p1 + v2*p2; reference body -> ECP_mul2
    void double_multiply(point1& p1, point1& p2, big& v1, big& v2) noexcept
    {
        ensure_context();
        must(c12381_double_multiply_miracl(&p1, &p2, &v1, &v2), "double_multiply");
    }

    // G2Point ^ Zp (g2_point.hpp:202-217); reference body -> PAIR_G2mul
    void multiply(point2& object, const big& value) noexcept
    {
        ensure_context();
        must(c12381_multiply_point2_miracl(&object, &value), "multiply(point2)");
    }

    // GTPoint * GTPoint (liner_pair.hpp:130-151); reference body -> FP12_mul
    void multiply(fp12& result, fp12& value) noexcept
    {
        ensure_context();
        must(c12381_fp12_multiply_miracl(&result, &value), "multiply(fp12)");
    }

    // GTPoint ^ Zp (liner_pair.hpp:159-174); reference body -> FP12_pow
    void pow(fp12& result, fp12& base, const big& exponent) noexcept
    {
        ensure_context();
        must(c12381_fp12_pow_miracl(&result, &base, &exponent), "pow(fp12)");
    }

    // GTMiller from a pair (liner_pair.hpp:247-255); reference body -> PAIR_ate
    void pair_ate(fp12& result, point2& p2, point1& p1) noexcept
    {
        ensure_context();
        must(c12381_pair_ate_miracl(&result, &p2, &p1), "pair_ate");
    }

    // GT_point() (liner_pair.hpp:203-209) and operator== (:336-357); reference body -> PAIR_fexp
    void pair_final_exponentiation(fp12& object) noexcept
    {
        ensure_context();
        must(c12381_pair_final_exponentiation_miracl(&object), "pair_final_exponentiation");
    }

    // pair * pair (liner_pair.hpp:291-303); reference body -> PAIR_double_ate
    void pair_double_ate(fp12& result, point2& p2, point1& p1, point2& q2, point1& q1) noexcept
    {
        ensure_context();
        must(c12381_pair_double_ate_miracl(&result, &p2, &p1, &q2, &q1), "pair_double_ate");
    }
}
