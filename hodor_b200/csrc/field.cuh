// 256-bit prime-field arithmetic for sm_100a: 8 x u32 limbs, Montgomery form with R = 2^256,
// canonical representatives -- the in-memory format of ff_ce's derive(PrimeField) for a 4 x u64
// repr (reference: src/bn256.rs:4-7; Montgomery/R evidence src/experiments/square_root_calculator/
// fp2.rs:10-22).  One field per template instantiation; no runtime dispatch inside kernels.
//
// The multiplier keeps two carry-save accumulators, one for the even-indexed and one for the
// odd-indexed limb products, so that every 32x32->64 product lands on a 64-bit aligned register
// pair (ptxas then fuses each mad.lo.cc/madc.hi.cc pair into one IMAD.WIDE.U32.X).  The division
// by 2^32 after every reduction step is a role swap of the two accumulators, not a data move.
//
// The same algorithm body compiles for the host (carry chains emulated with 64-bit arithmetic)
// so that it can be unit-tested without a GPU; the product never calls the host variant.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define HD __host__ __device__ __forceinline__
#define DEV __device__ __forceinline__
#else
#define HD inline
#define DEV inline
#endif

namespace hodor {

enum FieldId : int { FIELD_BLS12_381_FR = 0, FIELD_BN254_FR = 1, FIELD_STARK252 = 2, NUM_FIELDS = 3 };

struct Fe {
    uint32_t v[8];
};

// ---------------------------------------------------------------------------------------------
// Field parameter packs.  Limbs little-endian, 32 bit.
// ---------------------------------------------------------------------------------------------
struct BlsFr {  // what the reference's src/bn256.rs declares: the BLS12-381 scalar field
    static constexpr int ID = FIELD_BLS12_381_FR;
    static constexpr uint32_t INV = 0xffffffffu;  // -p^-1 mod 2^32
    static constexpr int S = 32, NUM_BITS = 255, GENERATOR = 7;
    HD static constexpr uint32_t P(int i) {
        constexpr uint32_t t[8] = {0x00000001u, 0xffffffffu, 0xfffe5bfeu, 0x53bda402u,
                                   0x09a1d805u, 0x3339d808u, 0x299d7d48u, 0x73eda753u};
        return t[i];
    }
    HD static constexpr uint32_t ONE(int i) {  // R mod p
        constexpr uint32_t t[8] = {0xfffffffeu, 0x00000001u, 0x00034802u, 0x5884b7fau,
                                   0xecbc4ff5u, 0x998c4fefu, 0xacc5056fu, 0x1824b159u};
        return t[i];
    }
    HD static constexpr uint32_t R2(int i) {  // R^2 mod p
        constexpr uint32_t t[8] = {0xf3f29c6du, 0xc999e990u, 0x87925c23u, 0x2b6cedcbu,
                                   0x7254398fu, 0x05d31496u, 0x9f59ff11u, 0x0748d9d9u};
        return t[i];
    }
    HD static constexpr uint32_t NINV256(int i) {  // -p^-1 mod 2^256 (mul_pre operand derivation)
        constexpr uint32_t t[8] = {0xffffffffu, 0xfffffffeu, 0xfffe5bfdu, 0x53ba5bffu, 0x0004ec06u, 0x181b2c17u, 0xd7bf2839u, 0x3d443ab0u};
        return t[i];
    }
    HD static constexpr uint32_t NP(int i) { return i == 0 ? 0u - P(0) : ~P(i); }  // 2^256 - p (p odd)
};

// The field usually called "bn256 Fr" (not what src/bn256.rs holds).  GENERATOR = 7 is the declaration of the Rust
// types a matter-labs caller links for this field (pairing_ce / bellman_ce `bn256::Fr`: PrimeFieldGenerator = "7";
// halo2curves' bn256 Fr uses 7 as well and publishes the resulting ROOT_OF_UNITY, pinned in tests/test_oracle_pins.py),
// so multiplicative_generator(), root_of_unity(), every domain generator and coset shift are bit-compatible with them.
struct Bn254Fr {
    static constexpr int ID = FIELD_BN254_FR;
    static constexpr uint32_t INV = 0xefffffffu;
    static constexpr int S = 28, NUM_BITS = 254, GENERATOR = 7;
    HD static constexpr uint32_t P(int i) {
        constexpr uint32_t t[8] = {0xf0000001u, 0x43e1f593u, 0x79b97091u, 0x2833e848u,
                                   0x8181585du, 0xb85045b6u, 0xe131a029u, 0x30644e72u};
        return t[i];
    }
    HD static constexpr uint32_t ONE(int i) {
        constexpr uint32_t t[8] = {0x4ffffffbu, 0xac96341cu, 0x9f60cd29u, 0x36fc7695u,
                                   0x7879462eu, 0x666ea36fu, 0x9a07df2fu, 0x0e0a77c1u};
        return t[i];
    }
    HD static constexpr uint32_t R2(int i) {
        constexpr uint32_t t[8] = {0xae216da7u, 0x1bb8e645u, 0xe35c59e3u, 0x53fe3ab1u,
                                   0x53bb8085u, 0x8c49833du, 0x7f4e44a5u, 0x0216d0b1u};
        return t[i];
    }
    HD static constexpr uint32_t NINV256(int i) {  // -p^-1 mod 2^256 (mul_pre operand derivation)
        constexpr uint32_t t[8] = {0xefffffffu, 0xc2e1f593u, 0x4c6911b3u, 0x6586864bu, 0x99062391u, 0xe39a9828u, 0x0d8341b2u, 0x73f82f1du};
        return t[i];
    }
    HD static constexpr uint32_t NP(int i) { return i == 0 ? 0u - P(0) : ~P(i); }  // 2^256 - p (p odd)
};

struct Stark252 {  // src/experiments/mod.rs:18-21
    static constexpr int ID = FIELD_STARK252;
    static constexpr uint32_t INV = 0xffffffffu;
    static constexpr int S = 192, NUM_BITS = 252, GENERATOR = 3;
    HD static constexpr uint32_t P(int i) {
        constexpr uint32_t t[8] = {0x00000001u, 0x00000000u, 0x00000000u, 0x00000000u,
                                   0x00000000u, 0x00000000u, 0x00000011u, 0x08000000u};
        return t[i];
    }
    HD static constexpr uint32_t ONE(int i) {
        constexpr uint32_t t[8] = {0xffffffe1u, 0xffffffffu, 0xffffffffu, 0xffffffffu,
                                   0xffffffffu, 0xffffffffu, 0xfffffdf0u, 0x07ffffffu};
        return t[i];
    }
    HD static constexpr uint32_t R2(int i) {
        constexpr uint32_t t[8] = {0x7e000401u, 0xfffffd73u, 0x330fffffu, 0x00000001u,
                                   0xff6f8000u, 0xffffffffu, 0x5e008810u, 0x07ffd4abu};
        return t[i];
    }
    HD static constexpr uint32_t NINV256(int i) {  // -p^-1 mod 2^256 (mul_pre operand derivation)
        constexpr uint32_t t[8] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu, 0x00000010u, 0x08000000u};
        return t[i];
    }
    HD static constexpr uint32_t NP(int i) { return i == 0 ? 0u - P(0) : ~P(i); }  // 2^256 - p (p odd)
};

// ---------------------------------------------------------------------------------------------
// Carry-chain rows.  Device: one asm statement per chain.  Host: 64-bit emulation.
// ---------------------------------------------------------------------------------------------

// acc[0..7] (as four 64-bit columns) += x0*y, x2*y<<64, x4*y<<128, x6*y<<192 ; top += carry out
HD void row_mad(uint32_t (&acc)[8], uint32_t x0, uint32_t x2, uint32_t x4, uint32_t x6, uint32_t y, uint32_t& top) {
#ifdef __CUDA_ARCH__
    asm("mad.lo.cc.u32   %0, %9,  %13, %0;\n\t"
        "madc.hi.cc.u32  %1, %9,  %13, %1;\n\t"
        "madc.lo.cc.u32  %2, %10, %13, %2;\n\t"
        "madc.hi.cc.u32  %3, %10, %13, %3;\n\t"
        "madc.lo.cc.u32  %4, %11, %13, %4;\n\t"
        "madc.hi.cc.u32  %5, %11, %13, %5;\n\t"
        "madc.lo.cc.u32  %6, %12, %13, %6;\n\t"
        "madc.hi.cc.u32  %7, %12, %13, %7;\n\t"
        "addc.u32        %8, %8, 0;"
        : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]), "+r"(acc[6]),
          "+r"(acc[7]), "+r"(top)
        : "r"(x0), "r"(x2), "r"(x4), "r"(x6), "r"(y));
#else
    const uint32_t x[4] = {x0, x2, x4, x6};
    uint64_t c = 0;
    for (int k = 0; k < 4; k++) {
        uint64_t prod = (uint64_t)x[k] * y;
        uint64_t lo = (uint64_t)acc[2 * k] + (uint32_t)prod + c;
        acc[2 * k] = (uint32_t)lo;
        uint64_t hi = (uint64_t)acc[2 * k + 1] + (prod >> 32) + (lo >> 32);
        acc[2 * k + 1] = (uint32_t)hi;
        c = hi >> 32;
    }
    top += (uint32_t)c;
#endif
}

// Same, but the carry out is provably zero (see mont_mul) and dropped.
HD void row_mad_nocarry(uint32_t (&acc)[8], uint32_t x0, uint32_t x2, uint32_t x4, uint32_t x6, uint32_t y) {
#ifdef __CUDA_ARCH__
    asm("mad.lo.cc.u32   %0, %8,  %12, %0;\n\t"
        "madc.hi.cc.u32  %1, %8,  %12, %1;\n\t"
        "madc.lo.cc.u32  %2, %9,  %12, %2;\n\t"
        "madc.hi.cc.u32  %3, %9,  %12, %3;\n\t"
        "madc.lo.cc.u32  %4, %10, %12, %4;\n\t"
        "madc.hi.cc.u32  %5, %10, %12, %5;\n\t"
        "madc.lo.cc.u32  %6, %11, %12, %6;\n\t"
        "madc.hi.u32     %7, %11, %12, %7;"
        : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]), "+r"(acc[6]),
          "+r"(acc[7])
        : "r"(x0), "r"(x2), "r"(x4), "r"(x6), "r"(y));
#else
    uint32_t dummy = 0;
    row_mad(acc, x0, x2, x4, x6, y, dummy);
#endif
}

// Reduction rows for moduli with p0 = 1, p1 = 0xffffffff (BLS12-381 Fr, i.e. src/bn256.rs): the two
// low products are additions.  m*p0 = m: (acc0, acc1) += m, which zeroes acc0 (m = -acc0).
HD void row_mad_p0_one(uint32_t (&acc)[8], uint32_t x2, uint32_t x4, uint32_t x6, uint32_t m, uint32_t& top) {
#ifdef __CUDA_ARCH__
    asm("add.cc.u32      %0, %0, %12;\n\t"
        "addc.cc.u32     %1, %1, 0;\n\t"
        "madc.lo.cc.u32  %2, %9,  %12, %2;\n\t"
        "madc.hi.cc.u32  %3, %9,  %12, %3;\n\t"
        "madc.lo.cc.u32  %4, %10, %12, %4;\n\t"
        "madc.hi.cc.u32  %5, %10, %12, %5;\n\t"
        "madc.lo.cc.u32  %6, %11, %12, %6;\n\t"
        "madc.hi.cc.u32  %7, %11, %12, %7;\n\t"
        "addc.u32        %8, %8, 0;"
        : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]), "+r"(acc[6]),
          "+r"(acc[7]), "+r"(top)
        : "r"(x2), "r"(x4), "r"(x6), "r"(m));
#else
    row_mad(acc, 1u, x2, x4, x6, m, top);
#endif
}
// m*p1 = m*(2^32 - 1): low word lo1 = -m, high word hi1 = m - (m != 0); carry out provably zero
HD void row_mad_p1_ones(uint32_t (&acc)[8], uint32_t lo1, uint32_t hi1, uint32_t x3, uint32_t x5, uint32_t x7,
                        uint32_t m) {
#ifdef __CUDA_ARCH__
    asm("add.cc.u32      %0, %0, %8;\n\t"
        "addc.cc.u32     %1, %1, %9;\n\t"
        "madc.lo.cc.u32  %2, %10, %13, %2;\n\t"
        "madc.hi.cc.u32  %3, %10, %13, %3;\n\t"
        "madc.lo.cc.u32  %4, %11, %13, %4;\n\t"
        "madc.hi.cc.u32  %5, %11, %13, %5;\n\t"
        "madc.lo.cc.u32  %6, %12, %13, %6;\n\t"
        "madc.hi.u32     %7, %12, %13, %7;"
        : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]), "+r"(acc[6]),
          "+r"(acc[7])
        : "r"(lo1), "r"(hi1), "r"(x3), "r"(x5), "r"(x7), "r"(m));
#else
    uint64_t s = (uint64_t)acc[0] + lo1;
    acc[0] = (uint32_t)s;
    s = (uint64_t)acc[1] + hi1 + (s >> 32);
    acc[1] = (uint32_t)s;
    uint64_t c = s >> 32;
    const uint32_t x[3] = {x3, x5, x7};
    for (int k = 0; k < 3; k++) {
        const uint64_t prod = (uint64_t)x[k] * m;
        const uint64_t lo = (uint64_t)acc[2 * k + 2] + (uint32_t)prod + c;
        acc[2 * k + 2] = (uint32_t)lo;
        const uint64_t hi = (uint64_t)acc[2 * k + 3] + (prod >> 32) + (lo >> 32);
        acc[2 * k + 3] = (uint32_t)hi;
        c = hi >> 32;
    }
#endif
}

// acc[0..7] = x0*y, x2*y<<64, ... (no accumulate, no carries)
HD void row_mul(uint32_t (&acc)[8], uint32_t x0, uint32_t x2, uint32_t x4, uint32_t x6, uint32_t y) {
#ifdef __CUDA_ARCH__
    asm("mul.lo.u32 %0, %8,  %12;\n\t"
        "mul.hi.u32 %1, %8,  %12;\n\t"
        "mul.lo.u32 %2, %9,  %12;\n\t"
        "mul.hi.u32 %3, %9,  %12;\n\t"
        "mul.lo.u32 %4, %10, %12;\n\t"
        "mul.hi.u32 %5, %10, %12;\n\t"
        "mul.lo.u32 %6, %11, %12;\n\t"
        "mul.hi.u32 %7, %11, %12;"
        : "=r"(acc[0]), "=r"(acc[1]), "=r"(acc[2]), "=r"(acc[3]), "=r"(acc[4]), "=r"(acc[5]), "=r"(acc[6]),
          "=r"(acc[7])
        : "r"(x0), "r"(x2), "r"(x4), "r"(x6), "r"(y));
#else
    const uint32_t x[4] = {x0, x2, x4, x6};
    for (int k = 0; k < 4; k++) {
        uint64_t prod = (uint64_t)x[k] * y;
        acc[2 * k] = (uint32_t)prod;
        acc[2 * k + 1] = (uint32_t)(prod >> 32);
    }
#endif
}

// The role swap: `lo` is the accumulator that becomes the new low-aligned one, `sh` the one whose
// word 0 is known to be zero and which is being divided by 2^32 twice (once by dropping word 0,
// once by the accumulator offset).  Computes
//     lo[0] += sh[1]                          (carry c)
//     sh'   = (sh >> 64) + c + (x1*y, x3*y<<64, x5*y<<128, x7*y<<192)
HD void row_shift_mad(uint32_t (&lo)[8], uint32_t (&sh)[8], uint32_t x1, uint32_t x3, uint32_t x5, uint32_t x7,
                      uint32_t y) {
#ifdef __CUDA_ARCH__
    asm("add.cc.u32      %0, %0, %2;\n\t"
        "madc.lo.cc.u32  %1, %9,  %13, %3;\n\t"
        "madc.hi.cc.u32  %2, %9,  %13, %4;\n\t"
        "madc.lo.cc.u32  %3, %10, %13, %5;\n\t"
        "madc.hi.cc.u32  %4, %10, %13, %6;\n\t"
        "madc.lo.cc.u32  %5, %11, %13, %7;\n\t"
        "madc.hi.cc.u32  %6, %11, %13, %8;\n\t"
        "madc.lo.cc.u32  %7, %12, %13, 0;\n\t"
        "madc.hi.u32     %8, %12, %13, 0;"
        : "+r"(lo[0]), "+r"(sh[0]), "+r"(sh[1]), "+r"(sh[2]), "+r"(sh[3]), "+r"(sh[4]), "+r"(sh[5]), "+r"(sh[6]),
          "+r"(sh[7])
        : "r"(x1), "r"(x3), "r"(x5), "r"(x7), "r"(y));
#else
    uint64_t t = (uint64_t)lo[0] + sh[1];
    lo[0] = (uint32_t)t;
    uint64_t c = t >> 32;
    const uint32_t x[4] = {x1, x3, x5, x7};
    uint32_t in[8] = {sh[2], sh[3], sh[4], sh[5], sh[6], sh[7], 0, 0};
    for (int k = 0; k < 4; k++) {
        uint64_t prod = (uint64_t)x[k] * y;
        uint64_t l = (uint64_t)in[2 * k] + (uint32_t)prod + c;
        sh[2 * k] = (uint32_t)l;
        uint64_t h = (uint64_t)in[2 * k + 1] + (prod >> 32) + (l >> 32);
        sh[2 * k + 1] = (uint32_t)h;
        c = h >> 32;
    }
#endif
}

// ---------------------------------------------------------------------------------------------
// r = a + b (no reduction), returns carry;   r = a - b, returns borrow (1 if a < b)
// ---------------------------------------------------------------------------------------------
HD uint32_t add256(uint32_t (&r)[8], const uint32_t (&a)[8], const uint32_t (&b)[8]) {
    uint32_t carry;
#ifdef __CUDA_ARCH__
    asm("add.cc.u32  %0, %9,  %17;\n\t"
        "addc.cc.u32 %1, %10, %18;\n\t"
        "addc.cc.u32 %2, %11, %19;\n\t"
        "addc.cc.u32 %3, %12, %20;\n\t"
        "addc.cc.u32 %4, %13, %21;\n\t"
        "addc.cc.u32 %5, %14, %22;\n\t"
        "addc.cc.u32 %6, %15, %23;\n\t"
        "addc.cc.u32 %7, %16, %24;\n\t"
        "addc.u32    %8, 0, 0;"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(carry)
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]), "r"(b[0]),
          "r"(b[1]), "r"(b[2]), "r"(b[3]), "r"(b[4]), "r"(b[5]), "r"(b[6]), "r"(b[7]));
#else
    uint64_t c = 0;
    for (int i = 0; i < 8; i++) {
        c += (uint64_t)a[i] + b[i];
        r[i] = (uint32_t)c;
        c >>= 32;
    }
    carry = (uint32_t)c;
#endif
    return carry;
}

HD uint32_t sub256(uint32_t (&r)[8], const uint32_t (&a)[8], const uint32_t (&b)[8]) {
    uint32_t borrow;
#ifdef __CUDA_ARCH__
    asm("sub.cc.u32  %0, %9,  %17;\n\t"
        "subc.cc.u32 %1, %10, %18;\n\t"
        "subc.cc.u32 %2, %11, %19;\n\t"
        "subc.cc.u32 %3, %12, %20;\n\t"
        "subc.cc.u32 %4, %13, %21;\n\t"
        "subc.cc.u32 %5, %14, %22;\n\t"
        "subc.cc.u32 %6, %15, %23;\n\t"
        "subc.cc.u32 %7, %16, %24;\n\t"
        "subc.u32    %8, 0, 0;"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(borrow)
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]), "r"(b[0]),
          "r"(b[1]), "r"(b[2]), "r"(b[3]), "r"(b[4]), "r"(b[5]), "r"(b[6]), "r"(b[7]));
    borrow &= 1u;  // subc of 0-0-borrow gives 0xffffffff
#else
    uint64_t bw = 0;
    for (int i = 0; i < 8; i++) {
        uint64_t d = (uint64_t)a[i] - b[i] - bw;
        r[i] = (uint32_t)d;
        bw = (d >> 32) & 1;
    }
    borrow = (uint32_t)bw;
#endif
    return borrow;
}

}  // namespace hodor
#include "mont_split.cuh"
#include "shoup_rows.cuh"
namespace hodor {

// ---------------------------------------------------------------------------------------------
// Field<F>: canonical Montgomery arithmetic
// ---------------------------------------------------------------------------------------------
template <class F>
struct Field {
    // Where the modulus lives.  ptxas fuses a mad.lo.cc/madc.hi.cc pair into one IMAD.WIDE.U32.X only
    // when both factors are vector registers; with an immediate (or uniform-register) factor it
    // emits IMAD.X + IMAD.HI.U32.X.  Measured on B200 (tools/microbench.cu, profiles/): IMAD.WIDE
    // issues at half the IMAD rate, so both forms cost the same multiplier time (58.8 vs 56.8
    // Gmul/s in favour of immediates) and the immediate form saves 8 registers per thread.  Define
    // HODOR_MODULUS_IN_REGS=1 to keep the modulus in registers: `opaque_zero` must then be 0 at run
    // time but unknown and per-thread at compile time (threadIdx.x & kernel_param_zero).
#ifndef HODOR_MODULUS_IN_REGS
#define HODOR_MODULUS_IN_REGS 0
#endif
    uint32_t p[8];
    HD explicit Field(uint32_t opaque_zero = 0) {
#pragma unroll
        for (int i = 0; i < 8; i++) p[i] = F::P(i) | (HODOR_MODULUS_IN_REGS ? opaque_zero : 0u);
    }

    HD static Fe one() {
        Fe r;
#pragma unroll
        for (int i = 0; i < 8; i++) r.v[i] = F::ONE(i);
        return r;
    }
    HD static Fe r2() {
        Fe r;
#pragma unroll
        for (int i = 0; i < 8; i++) r.v[i] = F::R2(i);
        return r;
    }
    HD static Fe zero() {
        Fe r;
#pragma unroll
        for (int i = 0; i < 8; i++) r.v[i] = 0;
        return r;
    }

    // t in [0, 2p) -> [0, p)
    HD void reduce_once(uint32_t (&t)[8]) const {
        uint32_t u[8];
        uint32_t borrow = sub256(u, t, p);
#pragma unroll
        for (int i = 0; i < 8; i++) t[i] = borrow ? t[i] : u[i];
    }

    HD Fe add(const Fe& a, const Fe& b) const {
        Fe r;
        add256(r.v, a.v, b.v);  // p < 2^255: a + b < 2^256, no carry
        reduce_once(r.v);
        return r;
    }
    HD Fe sub(const Fe& a, const Fe& b) const {
        Fe r;
        uint32_t borrow = sub256(r.v, a.v, b.v);
        uint32_t u[8];
        add256(u, r.v, p);
#pragma unroll
        for (int i = 0; i < 8; i++) r.v[i] = borrow ? u[i] : r.v[i];
        return r;
    }
    HD Fe neg(const Fe& a) const { return sub(zero(), a); }

    // x / 2 in the field.  Linear, so it is the same operation on Montgomery representations.
    HD Fe halve(const Fe& a) const {
        uint32_t q[8], t[8];
        const uint32_t odd = a.v[0] & 1u;
#pragma unroll
        for (int i = 0; i < 8; i++) q[i] = odd ? p[i] : 0u;
        add256(t, a.v, q);  // a + p < 2p < 2^256
        Fe r;
#pragma unroll
        for (int i = 0; i < 7; i++) r.v[i] = (t[i] >> 1) | (t[i + 1] << 31);
        r.v[7] = t[7] >> 1;
        return r;
    }

    // Montgomery product a*b*R^-1 mod p, canonical.  Inputs must be < p.
    //
    // Invariant kept across the 8 reduction steps: running total T = E + O * 2^32 with E, O >= 0
    // held in `even` / `odd` (8 words each).  One step for multiplier word y:
    //   T <- (T + a*y + m*p) / 2^32,  m = (T + a*y) * INV mod 2^32.
    // Even-indexed limbs of a and p multiply into the low-aligned accumulator, odd-indexed limbs
    // into the one offset by a word.  After adding m*p the low-aligned accumulator has word 0 == 0;
    // dropping that word makes it the offset accumulator of the next step (row_shift_mad), so the
    // two arrays trade roles each step.  Carries out of the low-aligned accumulator (weight 2^256)
    // are exactly one unit of word 7 of the offset accumulator.  The offset accumulator cannot
    // overflow: T + a*y + m*p < (2 + 2^33) p < 2^288 for every p < 0.49 * 2^256.
    //
    // Code size: the inlined body is ~300 SASS instructions, so a pass kernel with ~60 multiplies is
    // 200-280 KB of straight-line code and ncu shows ~20 % `no_instruction` stall samples.  An
    // out-of-line copy (HODOR_MUL_OUT_OF_LINE=1: static, by value, operands in registers) shrinks
    // the kernels to ~50 KB but measured 10 % SLOWER on the 2^24 LDE (45.0 vs 40.9 ms, call and
    // register-shuffle overhead; profiles/r01_experiments.md), so inlining stays the default.
#ifndef HODOR_MUL_OUT_OF_LINE
#define HODOR_MUL_OUT_OF_LINE 0
#endif
#if defined(__CUDA_ARCH__) && HODOR_MUL_OUT_OF_LINE && !HODOR_MODULUS_IN_REGS
    static __device__ __noinline__ Fe mul_out_of_line(Fe a, Fe b) { return Field<F>().mul_inline(a, b); }
    DEV Fe mul(const Fe& a, const Fe& b) const { return mul_out_of_line(a, b); }
#else
    HD Fe mul(const Fe& a, const Fe& b) const { return mul_inline(a, b); }
#endif

    // Which multiplier: 0 = even/odd carry-save (default), 1 = split lo/hi chains (mont_split.cuh;
    // measured slower: ptxas has to emit IMAD + IADD3.X per term because IMAD has no carry out).
#ifndef HODOR_MUL_SPLIT
#define HODOR_MUL_SPLIT 0
#endif
    HD Fe mul_inline(const Fe& a, const Fe& b) const {
#if HODOR_MUL_SPLIT
        return mul_split(a, b);
#else
        return mul_evenodd(a, b);
#endif
    }

    // Product by rows with separate low / high carry chains, then word-serial reduction; see
    // mont_split.cuh for the cost model and the overflow argument.
    // Short-cut reduction rows for p0 = 1, p1 = 0xffffffff (reduce_round below).  Measured on B200 in the
    // same session: the isolated multiplier gains 10 % (58.5 -> 64.2 Gmul/s) but the 2^24 LDE LOSES
    // 6 % (40.8 -> 43.2 ms): the extra dependent ALU instructions and registers (spills at the
    // 128-register cap) cost more than the 16 saved products.  Off by default.
#ifndef HODOR_USE_P01
#define HODOR_USE_P01 0
#endif
    static constexpr bool P01 = HODOR_USE_P01 && F::P(0) == 1u && F::P(1) == 0xffffffffu && F::INV == 0xffffffffu;
    HD Fe mul_split(const Fe& a, const Fe& b) const {
        uint32_t T[16], cy[9];
#pragma unroll
        for (int i = 9; i < 16; i++) T[i] = 0;
#pragma unroll
        for (int i = 0; i < 9; i++) cy[i] = 0;
        sp_row0_lo(T, a.v, b.v[0]);
        sp_row0_hi(T, a.v, b.v[0]);
#pragma unroll
        for (int i = 1; i < 8; i++) {
            sp_row_lo(T + i, a.v, b.v[i]);
            sp_row_hi(T + i, a.v, b.v[i]);
        }
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if constexpr (P01) {
                const uint32_t m = 0u - T[i];
                const uint32_t hi1 = m - (m != 0u ? 1u : 0u);
                sp_red_lo_p01(T + i, cy[i], m, p);
                sp_red_hi_p01(T + i, cy[i + 1], hi1, m, p);
            } else {
                const uint32_t m = T[i] * F::INV;
                sp_red_lo(T + i, cy[i], m, p);
                sp_red_hi(T + i, cy[i + 1], m, p);
            }
        }
        Fe r;
        uint32_t hi[8], c8[8];
#pragma unroll
        for (int i = 0; i < 8; i++) {
            hi[i] = T[8 + i];
            c8[i] = cy[i];  // cy[8] == 0: nothing can leave word 15
        }
        add256(r.v, hi, c8);
        reduce_once(r.v);
        return r;
    }

    // One reduction round: low += m*p (even limbs of p), off += m*p (odd limbs), m = -low[0]/p mod 2^32.
    // For p0 = 1, p1 = 0xffffffff the two low products are additions and m is a negation.
    HD void reduce_round(uint32_t (&low)[8], uint32_t (&off)[8]) const {
        if constexpr (P01) {
            const uint32_t t0 = low[0];
            const uint32_t m = 0u - t0;
            const uint32_t hi1 = m - (m != 0u ? 1u : 0u);
            row_mad_p1_ones(off, t0, hi1, p[3], p[5], p[7], m);
            row_mad_p0_one(low, p[2], p[4], p[6], m, off[7]);
        } else {
            const uint32_t m = low[0] * F::INV;
            row_mad_nocarry(off, p[1], p[3], p[5], p[7], m);
            row_mad(low, p[0], p[2], p[4], p[6], m, off[7]);
        }
    }

    HD Fe mul_evenodd(const Fe& a, const Fe& b) const {
        uint32_t even[8], odd[8];
        // step 0
        row_mul(even, a.v[0], a.v[2], a.v[4], a.v[6], b.v[0]);
        row_mul(odd, a.v[1], a.v[3], a.v[5], a.v[7], b.v[0]);
        reduce_round(even, odd);
#pragma unroll
        for (int i = 1; i < 8; i += 2) {
            {  // odd step: `odd` is low-aligned, `even` is shifted
                row_shift_mad(odd, even, a.v[1], a.v[3], a.v[5], a.v[7], b.v[i]);
                row_mad(odd, a.v[0], a.v[2], a.v[4], a.v[6], b.v[i], even[7]);
                reduce_round(odd, even);
            }
            if (i + 1 < 8) {  // even step: roles back
                row_shift_mad(even, odd, a.v[1], a.v[3], a.v[5], a.v[7], b.v[i + 1]);
                row_mad(even, a.v[0], a.v[2], a.v[4], a.v[6], b.v[i + 1], odd[7]);
                reduce_round(even, odd);
            }
        }
        // after step 7: low-aligned = odd (word 0 == 0), offset = even.  T = (odd >> 32) + even.
        Fe r;
        uint32_t hi[8];
#pragma unroll
        for (int i = 0; i < 7; i++) hi[i] = odd[i + 1];
        hi[7] = 0;
        add256(r.v, even, hi);  // T < 2p < 2^256
        reduce_once(r.v);
        return r;
    }

    HD Fe sqr(const Fe& a) const { return mul(a, a); }

    // -----------------------------------------------------------------------------------------
    // Multiplication by a FIXED operand with a precomputed quotient ("Shoup" form).  Every multiply
    // of the transform kernels is by a table entry (twiddle, coset power, n^-1), so the table holds
    //     w  = the plain integer value of the factor  (a_mont * w mod p is (a*w) in Montgomery form)
    //     wq = floor(w * 2^256 / p)
    // and   q = floor(a * wq / 2^256),   r = a*w - q*p  in [0, 2p)   for every a < 2^256
    // (w*2^256 = wq*p + eps, a*wq = q*2^256 + delta  =>  r = (delta*p + a*eps) / 2^256 < 2p).
    // Cost: the HIGH half of one product and the LOW halves of two (the second by the compile-time
    // words of -p), 36+7 / 28+8 / <= 28+8 32-bit products instead of Montgomery's 64 + 64 + 8.
    // Only words >= 7 of a*wq are formed (shoup_rows.cuh); what is dropped, D, is < 14 * 2^224, so
    // the truncation can change q only when the guard word (word 7) is within 14 of wrapping: then
    // (about 3 in 10^9 multiplies) the exact carry is recomputed out of line.  Result canonical,
    // bit-identical to mul(a, to_mont(w)).
    // -----------------------------------------------------------------------------------------
    static constexpr uint32_t PRE_GUARD = 0xfffffff2u;  // 2^32 - 14

    // floor(D / 2^224) for the dropped part D of a*b (products with i+j <= 5 and the low words of i+j = 6)
    HD static uint32_t pre_dropped_carry(const uint32_t (&a)[8], const uint32_t (&b)[8]) {
        uint64_t carry = 0;  // running value of the columns below, divided by 2^(32*col)
        for (int col = 0; col <= 6; col++) {
            uint64_t lo = carry & 0xffffffffu, hi = carry >> 32;
            for (int i = 0; i <= col; i++) {
                const uint64_t prod = (uint64_t)a[i] * b[col - i];
                lo += prod & 0xffffffffu;
                if (col < 6) hi += prod >> 32;  // column 6's high words were kept (they are in word 7)
            }
            carry = hi + (lo >> 32);
        }
        return (uint32_t)carry;  // < 14
    }
#ifdef __CUDA_ARCH__
    static __device__ __noinline__ uint32_t pre_dropped_carry_slow(Fe a, Fe b) { return pre_dropped_carry(a.v, b.v); }
#endif

    template <uint32_t GUARD = PRE_GUARD>
    HD Fe mul_pre(const Fe& a, const Fe& w, const Fe& wq) const {
        Fe r;
#ifdef __CUDA_ARCH__
        uint32_t q[8], guard;
        shoup_hi_trunc(q, guard, a.v, wq.v);
        if (guard >= GUARD) {
            const uint32_t c = (uint32_t)(((uint64_t)guard + pre_dropped_carry_slow(a, wq)) >> 32);
            uint32_t cv[8] = {c, 0, 0, 0, 0, 0, 0, 0};
            add256(q, q, cv);
        }
        shoup_lo2<F>(r.v, a.v, w.v, q);
#else
        // host: the same integers by plain 64-bit arithmetic (exact q, no truncation)
        uint32_t t[16] = {0};
        for (int i = 0; i < 8; i++) {
            uint64_t c = 0;
            for (int j = 0; j < 8; j++) {
                c += (uint64_t)a.v[i] * wq.v[j] + t[i + j];
                t[i + j] = (uint32_t)c;
                c >>= 32;
            }
            t[i + 8] = (uint32_t)c;
        }
        uint32_t acc[8] = {0};
        for (int pass = 0; pass < 2; pass++) {  // acc = a*w + q*NP mod 2^256
            for (int i = 0; i < 8; i++) {
                uint64_t c = 0;
                for (int j = 0; i + j < 8; j++) {
                    const uint32_t x = pass == 0 ? a.v[i] : t[8 + i];
                    const uint32_t y = pass == 0 ? w.v[j] : F::NP(j);
                    c += (uint64_t)x * y + acc[i + j];
                    acc[i + j] = (uint32_t)c;
                    c >>= 32;
                }
            }
        }
        for (int i = 0; i < 8; i++) r.v[i] = acc[i];
#endif
        reduce_once(r.v);
        return r;
    }

    // (w, wq) from the Montgomery form of the factor: w = w_mont / R, and since
    // w * 2^256 = wq * p + w_mont exactly, wq = w_mont * (-p^-1) mod 2^256.
    HD void make_pre(const Fe& w_mont, Fe& w, Fe& wq) const {
        w = from_mont(w_mont);
        uint32_t acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
        for (int i = 0; i < 8; i++) {
            uint64_t c = 0;
#pragma unroll
            for (int j = 0; j < 8; j++) {
                if (i + j < 8) {
                    c += (uint64_t)w_mont.v[i] * F::NINV256(j) + acc[i + j];
                    acc[i + j] = (uint32_t)c;
                    c >>= 32;
                }
            }
        }
#pragma unroll
        for (int i = 0; i < 8; i++) wq.v[i] = acc[i];
    }

    // plain integer (< p) -> Montgomery (ff_ce from_repr) and back (into_repr)
    HD Fe to_mont(const Fe& a) const { return mul(a, r2()); }
    HD Fe from_mont(const Fe& a) const {
        Fe o = zero();
        o.v[0] = 1;
        return mul(a, o);
    }

    HD Fe pow(const Fe& base, uint64_t e) const {
        Fe acc = one(), b = base;
        while (e) {
            if (e & 1) acc = mul(acc, b);
            b = mul(b, b);
            e >>= 1;
        }
        return acc;
    }

    HD static bool eq(const Fe& a, const Fe& b) {
        uint32_t d = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) d |= a.v[i] ^ b.v[i];
        return d == 0;
    }
    HD bool is_canonical(const Fe& a) const {
        uint32_t u[8];
        return sub256(u, a.v, p) != 0;  // a < p  <=>  a - p borrows
    }
};

}  // namespace hodor
