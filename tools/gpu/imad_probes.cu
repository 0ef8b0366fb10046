// Integer-multiply pipe probes for sm_100a (standalone; build: tools/gpu/build_probes.sh, run on the GPU box).
// What they answer (VERDICT r01 task 4): how fast does the fmaheavy pipe issue
//   IMAD, IMAD.WIDE (64-bit addend, no carry), IMAD.WIDE with an immediate multiplier, IMAD.WIDE.X carry chains (1 / 2 / 4
//   independent chains per thread), chain heads (carry-out only), and - the question behind them - a whole Montgomery product
//   on 12 x 32-bit saturated limbs (carry chains; the product the library ships) against one on 13 x 30-bit unsaturated limbs
//   (plain IMAD.WIDE into 64-bit columns, carries resolved by shifts on the ALU pipe), at full occupancy and at the
//   two-warps-per-scheduler occupancy of the pairing kernels.
// Every kernel's result is consumed; the unsaturated product is checked on the device against the shipped one.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../crypto12381_b200/csrc/fp2.cuh"
#include "../../crypto12381_b200/csrc/fp30.cuh"

using namespace c12;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__global__ void k_imad(uint32_t* out, int iters, uint32_t seed)
{
    uint32_t a[8];
    for (int i = 0; i < 8; ++i) a[i] = threadIdx.x + seed + i;
    uint32_t m = seed | 1u;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < 8; ++j)
            asm volatile("mad.lo.u32 %0, %0, %8, %1;\n\tmad.lo.u32 %1, %1, %8, %2;\n\tmad.lo.u32 %2, %2, %8, %3;\n\tmad.lo.u32 %3, %3, %8, %4;\n\t"
                         "mad.lo.u32 %4, %4, %8, %5;\n\tmad.lo.u32 %5, %5, %8, %6;\n\tmad.lo.u32 %6, %6, %8, %7;\n\tmad.lo.u32 %7, %7, %8, %0;\n\t"
                         : "+r"(a[0]), "+r"(a[1]), "+r"(a[2]), "+r"(a[3]), "+r"(a[4]), "+r"(a[5]), "+r"(a[6]), "+r"(a[7]) : "r"(m));
    }
    uint32_t x = 0;
    for (int i = 0; i < 8; ++i) x ^= a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = x;
}

// 8 independent 64-bit accumulators, register multiplier / immediate multiplier
template <int IMM> __global__ void k_wide(uint32_t* out, int iters, uint32_t seed)
{
    unsigned long long a[8];
    for (int i = 0; i < 8; ++i) a[i] = threadIdx.x + seed + i;
    uint32_t m = seed | 1u, q = seed * 3u + 5u;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (IMM)
                asm volatile("mad.wide.u32 %0, %8, 0x1eabfffe, %0;\n\tmad.wide.u32 %1, %8, 0xb153ffff, %1;\n\tmad.wide.u32 %2, %8, 0xb9feffff, %2;\n\t"
                             "mad.wide.u32 %3, %8, 0xffffaaab, %3;\n\tmad.wide.u32 %4, %8, 0xf6b0f624, %4;\n\tmad.wide.u32 %5, %8, 0x6730d2a0, %5;\n\t"
                             "mad.wide.u32 %6, %8, 0xf38512bf, %6;\n\tmad.wide.u32 %7, %8, 0x64774b84, %7;\n\t"
                             : "+l"(a[0]), "+l"(a[1]), "+l"(a[2]), "+l"(a[3]), "+l"(a[4]), "+l"(a[5]), "+l"(a[6]), "+l"(a[7]) : "r"(m));
            else
                asm volatile("mad.wide.u32 %0, %8, %9, %0;\n\tmad.wide.u32 %1, %8, %9, %1;\n\tmad.wide.u32 %2, %8, %9, %2;\n\tmad.wide.u32 %3, %8, %9, %3;\n\t"
                             "mad.wide.u32 %4, %8, %9, %4;\n\tmad.wide.u32 %5, %8, %9, %5;\n\tmad.wide.u32 %6, %8, %9, %6;\n\tmad.wide.u32 %7, %8, %9, %7;\n\t"
                             : "+l"(a[0]), "+l"(a[1]), "+l"(a[2]), "+l"(a[3]), "+l"(a[4]), "+l"(a[5]), "+l"(a[6]), "+l"(a[7]) : "r"(m), "r"(q));
            m += (uint32_t)a[7];
        }
    }
    unsigned long long x = 0;
    for (int i = 0; i < 8; ++i) x ^= a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = (uint32_t)x ^ (uint32_t)(x >> 32);
}

// CH independent carry chains of four (mad.lo.cc, madc.hi.cc) pairs each, interleaved instruction by instruction is ptxas' job:
// each chain is its own asm block over its own 8 registers
template <int CH> __global__ void k_chain(uint32_t* out, int iters, uint32_t seed)
{
    uint32_t a[CH][8];
    for (int c = 0; c < CH; ++c)
        for (int i = 0; i < 8; ++i) a[c][i] = threadIdx.x + seed + 8 * c + i;
    uint32_t m = seed | 1u, q = seed * 3u + 5u;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
#pragma unroll
            for (int c = 0; c < CH; ++c)
                asm("mad.lo.cc.u32 %0, %8, %9, %0;\n\tmadc.hi.cc.u32 %1, %8, %9, %1;\n\tmadc.lo.cc.u32 %2, %8, %9, %2;\n\tmadc.hi.cc.u32 %3, %8, %9, %3;\n\t"
                    "madc.lo.cc.u32 %4, %8, %9, %4;\n\tmadc.hi.cc.u32 %5, %8, %9, %5;\n\tmadc.lo.cc.u32 %6, %8, %9, %6;\n\tmadc.hi.u32 %7, %8, %9, %7;\n\t"
                    : "+r"(a[c][0]), "+r"(a[c][1]), "+r"(a[c][2]), "+r"(a[c][3]), "+r"(a[c][4]), "+r"(a[c][5]), "+r"(a[c][6]), "+r"(a[c][7])
                    : "r"(m), "r"(q));
            m += a[0][7];
        }
    }
    uint32_t x = 0;
    for (int c = 0; c < CH; ++c)
        for (int i = 0; i < 8; ++i) x ^= a[c][i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = x;
}

// chain heads only: (mad.lo.cc, madc.hi) pairs - carry out of the low half into the high half, nothing in, nothing out
__global__ void k_heads(uint32_t* out, int iters, uint32_t seed)
{
    uint32_t a[8];
    for (int i = 0; i < 8; ++i) a[i] = threadIdx.x + seed + i;
    uint32_t m = seed | 1u, q = seed * 3u + 5u;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            asm("mad.lo.cc.u32 %0, %8, %9, %0;\n\tmadc.hi.u32 %1, %8, %9, %1;\n\tmad.lo.cc.u32 %2, %8, %9, %2;\n\tmadc.hi.u32 %3, %8, %9, %3;\n\t"
                "mad.lo.cc.u32 %4, %8, %9, %4;\n\tmadc.hi.u32 %5, %8, %9, %5;\n\tmad.lo.cc.u32 %6, %8, %9, %6;\n\tmadc.hi.u32 %7, %8, %9, %7;\n\t"
                : "+r"(a[0]), "+r"(a[1]), "+r"(a[2]), "+r"(a[3]), "+r"(a[4]), "+r"(a[5]), "+r"(a[6]), "+r"(a[7]) : "r"(m), "r"(q));
            m += a[7];
        }
    }
    uint32_t x = 0;
    for (int i = 0; i < 8; ++i) x ^= a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = x;
}

// CH independent chains of Montgomery products per thread: saturated 12 x 32 (the shipped inline-PTX product) ...
template <int CH> __global__ void k_fp32(uint32_t* out, int iters, uint32_t seed)
{
    // operands come from memory (whatever the buffer holds, reduced below p by clearing the top bits): nothing for the compiler to fold
    Fp c[CH], b;
    for (int k = 0; k < 12; ++k) b.v[k] = out[(blockIdx.x * 12 + k) % 1024] ^ seed;
    b.v[11] &= 0x0fffffffu;
    for (int i = 0; i < CH; ++i) {
        for (int k = 0; k < 12; ++k) c[i].v[k] = out[(threadIdx.x * 13 + k + 7 * i) % 1024] + seed;
        c[i].v[11] &= 0x0fffffffu;
    }
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int i = 0; i < CH; ++i) c[i] = fp_mul_inl(c[i], b);
    uint32_t x = 0;
    for (int i = 0; i < CH; ++i)
        for (int k = 0; k < 12; ++k) x ^= c[i].v[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = x;
}
// ... and unsaturated 13 x 30
template <int CH> __global__ void k_fp30(uint32_t* out, int iters, uint32_t seed)
{
    Fp bb;
    for (int k = 0; k < 12; ++k) bb.v[k] = out[(blockIdx.x * 12 + k) % 1024] ^ seed;
    bb.v[11] &= 0x0fffffffu;
    Fp30 c[CH], b = fp30_from_fp(bb);
    for (int i = 0; i < CH; ++i) {
        Fp t;
        for (int k = 0; k < 12; ++k) t.v[k] = out[(threadIdx.x * 13 + k + 7 * i) % 1024] + seed;
        t.v[11] &= 0x0fffffffu;
        c[i] = fp30_from_fp(t);
    }
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int i = 0; i < CH; ++i) c[i] = fp30_mul(c[i], b);
    uint32_t x = 0;
    for (int i = 0; i < CH; ++i)
        for (int k = 0; k < 13; ++k) x ^= c[i].v[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = x;
}

// on-device cross check: x y 2^-390 (unsaturated) against the shipped product, on the canonical values
__global__ void k_check(uint32_t* bad, uint32_t seed)
{
    Fp a = fp_r2(), b = fp_one();
    a.v[0] ^= threadIdx.x * 2654435761u + seed;
    a.v[5] ^= blockIdx.x * 40503u;
    b.v[3] ^= threadIdx.x * 97u + blockIdx.x;
    a = fp_mul(a, fp_r2());
    b = fp_mul(b, a);
    // plain values u = a / R, v = b / R (R = 2^384)
    Fp30 x = fp30_from_fp(a), y = fp30_from_fp(b);
    for (int it = 0; it < 20; ++it) {
        Fp want = fp_mul(fp_mul(a, b), fp30_check_const());     // a b 2^-384 2^-384 2^378... see fp30.cuh: a b 2^-390 as a plain value
        Fp30 got = fp30_mul(x, y);
        Fp g = fp30_to_fp_canonical(got);
        if (!fp_eq(g, want)) atomicAdd(bad, 1u);
        a = fp_add(want, b);
        b = fp_sub(want, a);
        x = fp30_from_fp(a);
        y = fp30_from_fp(b);
    }
}

static double run(const char* name, void (*launch)(uint32_t*, int, int, int), uint32_t* out, int blocks, int threads, int iters, double ops_per_thread_iter, int sms)
{
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
        CK(cudaEventRecord(e0));
        launch(out, blocks, threads, iters);
        CK(cudaGetLastError());
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    double ops = (double)blocks * threads * iters * ops_per_thread_iter;
    double gops = ops / (best * 1e-3) / 1e9;
    int clk = 0;
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    double per_sm_clk = gops * 1e9 / sms / (clk * 1e3);
    printf("{\"probe\": \"%s\", \"blocks_per_sm\": %d, \"threads\": %d, \"ms\": %.4f, \"gops\": %.2f, \"ops_per_sm_per_clk\": %.2f}\n", name, blocks / sms, threads, best,
           gops, per_sm_clk);
    fflush(stdout);
    return gops;
}

#define L(kernel) [](uint32_t* o, int b, int t, int it) { kernel<<<b, t>>>(o, it, 12381u); }

int main(int argc, char** argv)
{
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    uint32_t* out;
    CK(cudaMalloc(&out, (size_t)sms * 8 * 256 * 4));
    uint32_t* bad;
    CK(cudaMalloc(&bad, 4));
    CK(cudaMemset(bad, 0, 4));
    k_check<<<64, 128>>>(bad, 7u);
    CK(cudaDeviceSynchronize());
    uint32_t hbad = 1;
    CK(cudaMemcpy(&hbad, bad, 4, cudaMemcpyDeviceToHost));
    printf("{\"check\": \"fp30_mul against fp_mul on %d products\", \"mismatches\": %u}\n", 64 * 128 * 20, hbad);
    const int it = argc > 1 ? atoi(argv[1]) : 2000;
    for (int occ = 0; occ < 2; ++occ) {
        const int bps = occ == 0 ? 8 : 1;          // 64 warps per SM, then 8 warps per SM (two per scheduler)
        const int blocks = sms * bps;
        run("imad", L(k_imad), out, blocks, 256, it, 64, sms);
        run("imad_wide", L(k_wide<0>), out, blocks, 256, it, 64, sms);
        run("imad_wide_imm", L(k_wide<1>), out, blocks, 256, it, 64, sms);
        run("chain_x1 (fused pairs)", L(k_chain<1>), out, blocks, 256, it, 32, sms);
        run("chain_x2 (fused pairs)", L(k_chain<2>), out, blocks, 256, it, 64, sms);
        run("chain_x4 (fused pairs)", L(k_chain<4>), out, blocks, 256, it, 128, sms);
        run("chain_heads (fused pairs)", L(k_heads), out, blocks, 256, it, 32, sms);
        run("fp_mul 12x32 x1", L(k_fp32<1>), out, blocks, 256, it / 4, 1, sms);
        run("fp_mul 12x32 x2", L(k_fp32<2>), out, blocks, 256, it / 4, 2, sms);
        run("fp_mul 12x32 x3", L(k_fp32<3>), out, blocks, 256, it / 4, 3, sms);
        run("fp_mul 13x30 x1", L(k_fp30<1>), out, blocks, 256, it / 4, 1, sms);
        run("fp_mul 13x30 x2", L(k_fp30<2>), out, blocks, 256, it / 4, 2, sms);
        run("fp_mul 13x30 x3", L(k_fp30<3>), out, blocks, 256, it / 4, 3, sms);
    }
    // 128-thread blocks, 3 per SM: the shape of k_accumulate
    run("fp_mul 12x32 x2 (128 thr x 3)", L(k_fp32<2>), out, sms * 3, 128, it / 4, 2, sms);
    run("fp_mul 13x30 x2 (128 thr x 3)", L(k_fp30<2>), out, sms * 3, 128, it / 4, 2, sms);
    return hbad != 0;
}
