// Second-generation 256-bit Montgomery multiplier for sm_100a: separate low-half and high-half
// carry chains, product first then word-serial reduction, with a short-cut for moduli whose two
// low 32-bit words are (1, 0xffffffff) -- which is the case for the reference's src/bn256.rs field
// (BLS12-381 Fr: p = ... ffffffff 00000001).
//
// Why: tools/pipebench.cu measures on B200 63.5 IMAD/clk/SM but only ~25 IMAD.WIDE/clk/SM.  ptxas
// fuses an adjacent mad.lo.cc/madc.hi.cc pair on the same product into one IMAD.WIDE.U32.X, which
// therefore costs MORE multiplier-pipe time (5 cycles per warp) than the two plain IMADs it
// replaces (2 + 2).  Here every 32x32 product is issued as one IMAD.X into a chain of low halves
// and one IMAD.HI.X into a chain of high halves, one word higher; the two chains never pair up, so
// nothing is fused.  Cost per multiplication (IMAD issue slots, the binding resource):
//     even/odd carry-save version (field.cuh, first generation):  64 WIDE + ~121 IMAD  ~ 562 cycles
//     this version, generic modulus:   128 (a*b) + 8 (m) + 120 (m*p)               = 512 cycles
//     this version, BLS12-381 Fr:      128 (a*b) + 96 (m*p, p0/p1 by additions)    = 448 cycles
//
// Algorithm.  T[0..15] = a*b by rows (row i adds a*b_i at word i: a chain of lo(a_j*b_i) into
// T[i..i+7], carry into T[i+8]; a chain of hi(a_j*b_i) into T[i+1..i+8]).  Then 8 reduction rounds:
// m = T[i] * (-p^-1) mod 2^32; T += m*p << 32i, which zeroes T[i].  lo(m*p0) is not computed: its
// only effect is the carry out of word i, which is 1 exactly when T[i] != 0.  Carries that leave a
// round's chains land in words >= 8, which no later m depends on, so they are collected in a small
// vector cy[] and added once at the end instead of being rippled every round.  Result
// (T[8..15] + cy) < 2p, one conditional subtraction.  All sums are of non-negative terms bounded by
// a*b + (2^256 - 1) p < 2^512, hence no carry can leave word 15 (nor word i+8 during product row i).
#pragma once
#include <stdint.h>

namespace hodor {

// T[0..7] = lo(a_j * y)
HD void sp_row0_lo(uint32_t* T, const uint32_t (&a)[8], uint32_t y) {
#ifdef __CUDA_ARCH__
    asm("mul.lo.u32 %0, %8,  %16;\n\t"
        "mul.lo.u32 %1, %9,  %16;\n\t"
        "mul.lo.u32 %2, %10, %16;\n\t"
        "mul.lo.u32 %3, %11, %16;\n\t"
        "mul.lo.u32 %4, %12, %16;\n\t"
        "mul.lo.u32 %5, %13, %16;\n\t"
        "mul.lo.u32 %6, %14, %16;\n\t"
        "mul.lo.u32 %7, %15, %16;"
        : "=r"(T[0]), "=r"(T[1]), "=r"(T[2]), "=r"(T[3]), "=r"(T[4]), "=r"(T[5]), "=r"(T[6]), "=r"(T[7])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]), "r"(y));
#else
    for (int j = 0; j < 8; j++) T[j] = (uint32_t)((uint64_t)a[j] * y);
#endif
}

// T[1..7] += hi(a_j * y), j = 0..6 ; T[8] = hi(a_7 * y) + carry
HD void sp_row0_hi(uint32_t* T, const uint32_t (&a)[8], uint32_t y) {
#ifdef __CUDA_ARCH__
    asm("mad.hi.cc.u32  %0, %8,  %16, %0;\n\t"
        "madc.hi.cc.u32 %1, %9,  %16, %1;\n\t"
        "madc.hi.cc.u32 %2, %10, %16, %2;\n\t"
        "madc.hi.cc.u32 %3, %11, %16, %3;\n\t"
        "madc.hi.cc.u32 %4, %12, %16, %4;\n\t"
        "madc.hi.cc.u32 %5, %13, %16, %5;\n\t"
        "madc.hi.cc.u32 %6, %14, %16, %6;\n\t"
        "madc.hi.u32    %7, %15, %16, 0;"
        : "+r"(T[1]), "+r"(T[2]), "+r"(T[3]), "+r"(T[4]), "+r"(T[5]), "+r"(T[6]), "+r"(T[7]), "=r"(T[8])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]), "r"(y));
#else
    uint64_t c = 0;
    for (int j = 0; j < 8; j++) {
        const uint64_t s = (j < 7 ? (uint64_t)T[1 + j] : 0) + (((uint64_t)a[j] * y) >> 32) + c;
        T[1 + j] = (uint32_t)s;
        c = s >> 32;
    }
#endif
}

// T[0..7] += lo(a_j * y) ; T[8] += carry          (T points at word i of the product)
HD void sp_row_lo(uint32_t* T, const uint32_t (&a)[8], uint32_t y) {
#ifdef __CUDA_ARCH__
    asm("mad.lo.cc.u32  %0, %9,  %17, %0;\n\t"
        "madc.lo.cc.u32 %1, %10, %17, %1;\n\t"
        "madc.lo.cc.u32 %2, %11, %17, %2;\n\t"
        "madc.lo.cc.u32 %3, %12, %17, %3;\n\t"
        "madc.lo.cc.u32 %4, %13, %17, %4;\n\t"
        "madc.lo.cc.u32 %5, %14, %17, %5;\n\t"
        "madc.lo.cc.u32 %6, %15, %17, %6;\n\t"
        "madc.lo.cc.u32 %7, %16, %17, %7;\n\t"
        "addc.u32       %8, %8, 0;"
        : "+r"(T[0]), "+r"(T[1]), "+r"(T[2]), "+r"(T[3]), "+r"(T[4]), "+r"(T[5]), "+r"(T[6]), "+r"(T[7]), "+r"(T[8])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]), "r"(y));
#else
    uint64_t c = 0;
    for (int j = 0; j < 8; j++) {
        const uint64_t s = (uint64_t)T[j] + (uint32_t)((uint64_t)a[j] * y) + c;
        T[j] = (uint32_t)s;
        c = s >> 32;
    }
    T[8] += (uint32_t)c;
#endif
}

// T[1..8] += hi(a_j * y)   (no carry can leave word 8, see header)
HD void sp_row_hi(uint32_t* T, const uint32_t (&a)[8], uint32_t y) {
#ifdef __CUDA_ARCH__
    asm("mad.hi.cc.u32  %0, %8,  %16, %0;\n\t"
        "madc.hi.cc.u32 %1, %9,  %16, %1;\n\t"
        "madc.hi.cc.u32 %2, %10, %16, %2;\n\t"
        "madc.hi.cc.u32 %3, %11, %16, %3;\n\t"
        "madc.hi.cc.u32 %4, %12, %16, %4;\n\t"
        "madc.hi.cc.u32 %5, %13, %16, %5;\n\t"
        "madc.hi.cc.u32 %6, %14, %16, %6;\n\t"
        "madc.hi.u32    %7, %15, %16, %7;"
        : "+r"(T[1]), "+r"(T[2]), "+r"(T[3]), "+r"(T[4]), "+r"(T[5]), "+r"(T[6]), "+r"(T[7]), "+r"(T[8])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]), "r"(y));
#else
    uint64_t c = 0;
    for (int j = 0; j < 8; j++) {
        const uint64_t s = (uint64_t)T[1 + j] + (((uint64_t)a[j] * y) >> 32) + c;
        T[1 + j] = (uint32_t)s;
        c = s >> 32;
    }
#endif
}

// ---- reduction round, generic modulus.  T points at word i; m = T[0] * INV.
// low halves: carry out of word 0 is (T[0] != 0); T[1..7] += lo(m * p_j), j = 1..7 ; cy_lo += carry
HD void sp_red_lo(uint32_t* T, uint32_t& cy_lo, uint32_t m, const uint32_t (&p)[8]) {
#ifdef __CUDA_ARCH__
    uint32_t scratch;
    asm("add.cc.u32     %8, %9, 0xffffffff;\n\t"
        "madc.lo.cc.u32 %0, %10, %17, %0;\n\t"
        "madc.lo.cc.u32 %1, %11, %17, %1;\n\t"
        "madc.lo.cc.u32 %2, %12, %17, %2;\n\t"
        "madc.lo.cc.u32 %3, %13, %17, %3;\n\t"
        "madc.lo.cc.u32 %4, %14, %17, %4;\n\t"
        "madc.lo.cc.u32 %5, %15, %17, %5;\n\t"
        "madc.lo.cc.u32 %6, %16, %17, %6;\n\t"
        "addc.u32       %7, %7, 0;"
        : "+r"(T[1]), "+r"(T[2]), "+r"(T[3]), "+r"(T[4]), "+r"(T[5]), "+r"(T[6]), "+r"(T[7]), "+r"(cy_lo), "=&r"(scratch)
        : "r"(T[0]), "r"(p[1]), "r"(p[2]), "r"(p[3]), "r"(p[4]), "r"(p[5]), "r"(p[6]), "r"(p[7]), "r"(m));
#else
    uint64_t c = T[0] != 0 ? 1 : 0;
    for (int j = 1; j < 8; j++) {
        const uint64_t s = (uint64_t)T[j] + (uint32_t)((uint64_t)p[j] * m) + c;
        T[j] = (uint32_t)s;
        c = s >> 32;
    }
    cy_lo += (uint32_t)c;
#endif
}
// high halves: T[1..8] += hi(m * p_j), j = 0..7 ; cy_hi += carry
HD void sp_red_hi(uint32_t* T, uint32_t& cy_hi, uint32_t m, const uint32_t (&p)[8]) {
#ifdef __CUDA_ARCH__
    asm("mad.hi.cc.u32  %0, %9,  %17, %0;\n\t"
        "madc.hi.cc.u32 %1, %10, %17, %1;\n\t"
        "madc.hi.cc.u32 %2, %11, %17, %2;\n\t"
        "madc.hi.cc.u32 %3, %12, %17, %3;\n\t"
        "madc.hi.cc.u32 %4, %13, %17, %4;\n\t"
        "madc.hi.cc.u32 %5, %14, %17, %5;\n\t"
        "madc.hi.cc.u32 %6, %15, %17, %6;\n\t"
        "madc.hi.cc.u32 %7, %16, %17, %7;\n\t"
        "addc.u32       %8, %8, 0;"
        : "+r"(T[1]), "+r"(T[2]), "+r"(T[3]), "+r"(T[4]), "+r"(T[5]), "+r"(T[6]), "+r"(T[7]), "+r"(T[8]), "+r"(cy_hi)
        : "r"(p[0]), "r"(p[1]), "r"(p[2]), "r"(p[3]), "r"(p[4]), "r"(p[5]), "r"(p[6]), "r"(p[7]), "r"(m));
#else
    uint64_t c = 0;
    for (int j = 0; j < 8; j++) {
        const uint64_t s = (uint64_t)T[1 + j] + (((uint64_t)p[j] * m) >> 32) + c;
        T[1 + j] = (uint32_t)s;
        c = s >> 32;
    }
    cy_hi += (uint32_t)c;
#endif
}

// ---- reduction round for p0 = 1, p1 = 0xffffffff (then INV = 0xffffffff and m = -T[0]):
//   lo(m*p0) = m (zeroes word 0, carry = T[0] != 0), hi(m*p0) = 0,
//   lo(m*p1) = -m = T[0],                            hi(m*p1) = m - (m != 0)
// low halves: T[1] += T[0] + carry ; T[2..7] += lo(m * p_j), j = 2..7 ; cy_lo += carry
HD void sp_red_lo_p01(uint32_t* T, uint32_t& cy_lo, uint32_t m, const uint32_t (&p)[8]) {
#ifdef __CUDA_ARCH__
    uint32_t scratch;
    asm("add.cc.u32     %8, %9, 0xffffffff;\n\t"
        "addc.cc.u32    %0, %0, %9;\n\t"
        "madc.lo.cc.u32 %1, %10, %16, %1;\n\t"
        "madc.lo.cc.u32 %2, %11, %16, %2;\n\t"
        "madc.lo.cc.u32 %3, %12, %16, %3;\n\t"
        "madc.lo.cc.u32 %4, %13, %16, %4;\n\t"
        "madc.lo.cc.u32 %5, %14, %16, %5;\n\t"
        "madc.lo.cc.u32 %6, %15, %16, %6;\n\t"
        "addc.u32       %7, %7, 0;"
        : "+r"(T[1]), "+r"(T[2]), "+r"(T[3]), "+r"(T[4]), "+r"(T[5]), "+r"(T[6]), "+r"(T[7]), "+r"(cy_lo), "=&r"(scratch)
        : "r"(T[0]), "r"(p[2]), "r"(p[3]), "r"(p[4]), "r"(p[5]), "r"(p[6]), "r"(p[7]), "r"(m));
#else
    uint64_t c = T[0] != 0 ? 1 : 0;
    uint64_t s = (uint64_t)T[1] + T[0] + c;
    T[1] = (uint32_t)s;
    c = s >> 32;
    for (int j = 2; j < 8; j++) {
        s = (uint64_t)T[j] + (uint32_t)((uint64_t)p[j] * m) + c;
        T[j] = (uint32_t)s;
        c = s >> 32;
    }
    cy_lo += (uint32_t)c;
#endif
}
// high halves: T[2] += hi1 ; T[3..8] += hi(m * p_j), j = 2..7 ; cy_hi += carry
HD void sp_red_hi_p01(uint32_t* T, uint32_t& cy_hi, uint32_t hi1, uint32_t m, const uint32_t (&p)[8]) {
#ifdef __CUDA_ARCH__
    asm("add.cc.u32     %0, %0, %8;\n\t"
        "madc.hi.cc.u32 %1, %9,  %15, %1;\n\t"
        "madc.hi.cc.u32 %2, %10, %15, %2;\n\t"
        "madc.hi.cc.u32 %3, %11, %15, %3;\n\t"
        "madc.hi.cc.u32 %4, %12, %15, %4;\n\t"
        "madc.hi.cc.u32 %5, %13, %15, %5;\n\t"
        "madc.hi.cc.u32 %6, %14, %15, %6;\n\t"
        "addc.u32       %7, %7, 0;"
        : "+r"(T[2]), "+r"(T[3]), "+r"(T[4]), "+r"(T[5]), "+r"(T[6]), "+r"(T[7]), "+r"(T[8]), "+r"(cy_hi)
        : "r"(hi1), "r"(p[2]), "r"(p[3]), "r"(p[4]), "r"(p[5]), "r"(p[6]), "r"(p[7]), "r"(m));
#else
    uint64_t s = (uint64_t)T[2] + hi1;
    T[2] = (uint32_t)s;
    uint64_t c = s >> 32;
    for (int j = 2; j < 8; j++) {
        s = (uint64_t)T[1 + j] + (((uint64_t)p[j] * m) >> 32) + c;
        T[1 + j] = (uint32_t)s;
        c = s >> 32;
    }
    cy_hi += (uint32_t)c;
#endif
}

}  // namespace hodor
