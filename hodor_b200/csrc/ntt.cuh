// NTT / coset-LDE kernels for sm_100a.
//
// Replaces (same result bits, different algorithm) the reference's
//   serial_fft / parallel_fft / best_fft            src/fft/fft.rs:5-125
//   distribute_powers                               src/fft/mod.rs:110-123
//   (coset_)lde_using_multiple_cosets               src/polynomials/mod.rs:418-482, 544-609
//
// Algorithm: recursive Cooley-Tukey over index digits.  The n = 2^log_n point transform is split
// into P passes over HBM; pass p transforms one digit of B_p bits (B_p in 6..9) of the index with a
// block-local 2^B_p-point decimation-in-frequency NTT, multiplies by the inter-pass twiddle
// omega_N^(k*r) and writes back in place; the last pass writes to the digit-reversed (= natural
// order) location.  Inside a block the 2^B-point NTT is done as two or three radix-8/4 groups held
// in registers, exchanged through shared memory.  A block always works on 8 adjacent "columns"
// (or 8 cosets / 8 adjacent output indices in the last pass) so that every global access is a
// 256-byte contiguous run and every shared-memory quarter-warp access is a conflict-free 128 B.
// An L-coset LDE is the same transform with one more, trivial, leading digit (the coset index):
// its "twiddle" is the coset scaling shift_i^j, fused into the loads of pass 1, and the last pass
// interleaves the cosets (out[i + L*k]) exactly as the reference's final gather does.
#pragma once
#include <cuda_runtime.h>
#include "field.cuh"

namespace hodor {

struct TwoLevel {  // base^e = hi[e >> lo_bits] * lo[e & ((1 << lo_bits) - 1)]
    const uint4* lo;
    const uint4* hi;
    uint32_t lo_bits;
};

enum : uint32_t {
    PASS_OUT_CONST = 1u,  // last pass: multiply outputs by out_const (ifft: n^-1)
    PASS_OUT_POW = 2u,    // last pass: multiply output k by out_pow^k (icoset_fft: g^-k, n^-1 folded into lo)
};

// A fixed multiplier in the form Field<F>::mul_pre wants: w = plain value, q = floor(w * 2^256 / p).
// Tables of these are 64 bytes per entry (w then q).
struct FePre {
    Fe w, q;
};

struct NttPass {
    const uint4* in;
    uint4* out;
    uint32_t log_n;    // transform length per coset
    uint32_t s;        // index bits below this pass's digit (0 for the last pass)
    uint32_t log_l;    // log2(number of cosets)
    uint32_t b1;       // width of the first digit (last pass: output index composition)
    uint32_t mid0;     // width of the 2nd digit when it is a middle digit of the last pass, else 0
    uint32_t mid1;     // width of the 3rd digit when it is a middle digit of the last pass, else 0
    uint32_t flags;
    uint32_t tw_shift;  // inter-pass twiddle exponent = (k * r) << tw_shift
    uint32_t zero;      // always 0; opaque to ptxas (keeps the modulus in vector registers)
    uint32_t coset_stride_lo, coset_stride_hi;  // elements between per-coset tables
    TwoLevel tw;        // powers of omega
    // Tables read directly as multipliers are in FePre form (64 B per entry); the two-level tables,
    // whose entries are multiplied with each other first, stay in Montgomery form.
    const uint4* tw_b;  // omega_B^x, x in [0, 2^B)  (FePre)
    const uint4* tw_direct;  // inter-pass twiddles omega_N^x, x in [0, N = 2^(s+B)), when N <= 2^16; else null  (FePre)
    // Expanded tables (the GPU counterpart of the reference's PrecomputedOmegas, src/precomputations/
    // mod.rs:14-66), streamed with the same coalesced addressing as the data; null => multiply two
    // table entries instead.  tw_full[(k << s) + r] = omega^(k * r) for pass 1 (n entries);
    // coset_full[i * n + j] = shift_i^j (L * n entries).  Both FePre.
    const uint4* tw_full;
    const uint4* coset_full;
    TwoLevel coset;     // pass 1 of a scaled transform: shift_i^j tables, coset i at +i*coset_stride
    TwoLevel out_pow;   // PASS_OUT_POW
    const uint4* out_pow_full;  // PASS_OUT_POW: out_pow^k as a flat fixed-operand table (FePre, n entries), or null
    FePre out_const;    // PASS_OUT_CONST
    // Sharded four-step NTT, last pass of step A: instead of `out`, output k goes straight into the receive buffer
    // of the rank that owns it for step B -- peer[k >> peer_chunk_log] + (peer_rank << peer_chunk_log) + (k & mask)
    // -- through NVLink peer stores: the all-to-all happens inside this kernel, tile by tile, under its arithmetic.
    uint4* peer[16];
    uint32_t peer_on, peer_chunk_log, peer_rank;
    FePre wr[7];        // omega_16^k, k = 1..7  (omega_8 = wr[1], omega_4 = wr[3])
};

// One element = 32 bytes = one sector.  sm_100 has 256-bit global loads / stores (LDG.E.256 / STG.E.256, PTX
// ld/st.global.v8.b32): one full-sector access per element instead of two half-sector ones -- half the LSU
// instructions, and a store that crosses NVLink (the peer stores of the sharded NTT) travels as whole sectors.
// Needs 32-byte aligned element arrays (every cudaMalloc'ed / torch buffer is; rows are 32 bytes).
#ifndef HODOR_LDST256
#define HODOR_LDST256 1
#endif
DEV void ld256(const void* ptr, uint32_t (&v)[8]) {
    asm volatile("ld.global.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "l"(ptr));
}
DEV void st256(void* ptr, const uint32_t (&v)[8]) {
    asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(ptr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]),
                 "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
                 : "memory");
}
DEV Fe ld_fe(const uint4* base, size_t idx) {
    Fe r;
#if HODOR_LDST256
    ld256(base + 2 * idx, r.v);
#else
    const uint4 a = base[2 * idx], b = base[2 * idx + 1];
    r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
    r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
#endif
    return r;
}
DEV void st_fe(uint4* base, size_t idx, const Fe& r) {
#if HODOR_LDST256
    st256(base + 2 * idx, r.v);
#else
    base[2 * idx] = make_uint4(r.v[0], r.v[1], r.v[2], r.v[3]);
    base[2 * idx + 1] = make_uint4(r.v[4], r.v[5], r.v[6], r.v[7]);
#endif
}
DEV FePre ld_pre(const uint4* base, size_t idx) {
    FePre r;
#if HODOR_LDST256
    ld256(base + 4 * idx, r.w.v);
    ld256(base + 4 * idx + 2, r.q.v);
#else
    const uint4 a = base[4 * idx], b = base[4 * idx + 1], c = base[4 * idx + 2], d = base[4 * idx + 3];
    r.w.v[0] = a.x; r.w.v[1] = a.y; r.w.v[2] = a.z; r.w.v[3] = a.w;
    r.w.v[4] = b.x; r.w.v[5] = b.y; r.w.v[6] = b.z; r.w.v[7] = b.w;
    r.q.v[0] = c.x; r.q.v[1] = c.y; r.q.v[2] = c.z; r.q.v[3] = c.w;
    r.q.v[4] = d.x; r.q.v[5] = d.y; r.q.v[6] = d.z; r.q.v[7] = d.w;
#endif
    return r;
}
DEV void st_pre(uint4* base, size_t idx, const FePre& r) {
#if HODOR_LDST256
    st256(base + 4 * idx, r.w.v);
    st256(base + 4 * idx + 2, r.q.v);
#else
    base[4 * idx] = make_uint4(r.w.v[0], r.w.v[1], r.w.v[2], r.w.v[3]);
    base[4 * idx + 1] = make_uint4(r.w.v[4], r.w.v[5], r.w.v[6], r.w.v[7]);
    base[4 * idx + 2] = make_uint4(r.q.v[0], r.q.v[1], r.q.v[2], r.q.v[3]);
    base[4 * idx + 3] = make_uint4(r.q.v[4], r.q.v[5], r.q.v[6], r.q.v[7]);
#endif
}
template <class F>
DEV Fe mul_by(const Field<F>& fld, const Fe& a, const FePre& m) {
    return fld.mul_pre(a, m.w, m.q);
}
// element idx of a shared-memory array of contiguous 32-byte elements (ld_fe / st_fe are global-only)
DEV Fe lds_elem(const uint4* sm, size_t idx) {
    const uint4 a = sm[2 * idx], b = sm[2 * idx + 1];
    Fe r;
    r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
    r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
    return r;
}
DEV void sts_elem(uint4* sm, size_t idx, const Fe& r) {
    sm[2 * idx] = make_uint4(r.v[0], r.v[1], r.v[2], r.v[3]);
    sm[2 * idx + 1] = make_uint4(r.v[4], r.v[5], r.v[6], r.v[7]);
}
// shared-memory tile: two planes of uint4 so that a quarter warp touching 8 adjacent columns
// reads 128 contiguous bytes
DEV Fe lds_fe(const uint4* sm, uint32_t plane, uint32_t slot) {
    const uint4 a = sm[slot], b = sm[plane + slot];
    Fe r;
    r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
    r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
    return r;
}
DEV void sts_fe(uint4* sm, uint32_t plane, uint32_t slot, const Fe& r) {
    sm[slot] = make_uint4(r.v[0], r.v[1], r.v[2], r.v[3]);
    sm[plane + slot] = make_uint4(r.v[4], r.v[5], r.v[6], r.v[7]);
}
// a kernel-parameter (constant bank) element as a multiplier operand; see Field<F>::Field on why
// it is not forced into vector registers by default
DEV Fe ld_param(const Fe& c, uint32_t opaque_zero) {
    Fe r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = c.v[i] | (HODOR_MODULUS_IN_REGS ? opaque_zero : 0u);
    return r;
}
DEV FePre ld_param(const FePre& c, uint32_t opaque_zero) {
    FePre r;
    r.w = ld_param(c.w, opaque_zero);
    r.q = ld_param(c.q, opaque_zero);
    return r;
}

#ifndef HODOR_PASS_PREFETCH
#define HODOR_PASS_PREFETCH 0  // 1: L2 prefetch of every HBM operand at kernel start (A/B build, profiles/r02_experiments.md)
#endif
DEV void prefetch_l2(const void* ptr) { asm volatile("prefetch.global.L2 [%0];" ::"l"(ptr)); }

template <class F>
DEV Fe two_level_pow(const Field<F>& fld, const TwoLevel& t, size_t lo_off, size_t hi_off, uint64_t e) {
    const Fe l = ld_fe(t.lo, lo_off + (e & ((1ull << t.lo_bits) - 1)));
    const Fe h = ld_fe(t.hi, hi_off + (e >> t.lo_bits));
    return fld.mul(l, h);
}

constexpr HD int bitrev_c(int x, int bits) {
    int r = 0;
    for (int i = 0; i < bits; i++) r |= ((x >> i) & 1) << (bits - 1 - i);
    return r;
}

// In-register radix-2^LOGR decimation-in-frequency NTT.  On return x[j] = X[bitrev(j)].
template <class F, int LOGR>
DEV void dif_inreg(const Field<F>& fld, Fe (&x)[1 << LOGR], const NttPass& p, uint32_t opaque_zero) {
    constexpr int R = 1 << LOGR;
#pragma unroll
    for (int t = 0; t < LOGR; t++) {
        const int half = R >> (t + 1);
#pragma unroll
        for (int blk = 0; blk < R; blk += 2 * half) {
#pragma unroll
            for (int i = 0; i < half; i++) {
                const Fe u = x[blk + i], v = x[blk + i + half];
                x[blk + i] = fld.add(u, v);
                const Fe d = fld.sub(u, v);
                const int e = (i << t) * (16 / R);  // omega_R^(i<<t) as a power of omega_16
                if (e == 0) {
                    x[blk + i + half] = d;
                } else {
                    x[blk + i + half] = mul_by(fld, d, ld_param(p.wr[e - 1], opaque_zero));
                }
            }
        }
    }
}

// How a 2^B-point block transform is split into register-resident radix-2^W butterflies: up to four
// groups, widths R1..R4 summing to B.  With radix-8 groups a thread keeps 8 elements live through the
// butterfly (124-128 registers: 2 blocks of 256 threads per SM); with radix-4 groups only 4 (68-72
// registers: 3 blocks per SM) for the same number of multiplies per element (B = 8: 3.25 either
// way), and the extra resident warps are what keeps the multiplier pipe fed while other warps sit
// in loads, barriers and add/sub work: measured 24.97 vs 26.66 ms on the 2^24 x 8 coset LDE.
// B = 9 (128 KiB tile: one block per SM whatever the register count) and B = 6 stay radix-8.
#ifndef HODOR_RADIX4_GROUPS
#define HODOR_RADIX4_GROUPS 1
#endif
#ifndef HODOR_B6_RADIX4
#define HODOR_B6_RADIX4 0  // B = 6 as 2+2+2 (2.25 multiplies per element, 12 blocks/SM) instead of 3+3 (2.125, 8 blocks/SM)
#endif
template <int B>
struct Groups {
    static constexpr bool RADIX4 = HODOR_RADIX4_GROUPS && (B == 7 || B == 8 || (HODOR_B6_RADIX4 && B == 6));
    static constexpr int R1 = RADIX4 ? (B == 7 ? 1 : 2) : 3;
    static constexpr int REM = B - 3;
    static constexpr int R2 = RADIX4 ? 2 : (REM <= 3 ? REM : (REM + 1) / 2);
    static constexpr int R3 = RADIX4 ? 2 : REM - R2;
    static constexpr int R4 = B - R1 - R2 - R3;
    static constexpr int NGROUPS = R4 > 0 ? 4 : (R3 > 0 ? 3 : 2);
};

// Resident blocks per SM the pass kernel is compiled for (8 * 2^B / EPT threads, EPT elements per
// thread, 2^B x 256 B of shared memory).  B = 8, the 2^24 workhorse: 3 x 256 threads at <= 85
// registers and 3 x 64 KiB, so that other blocks' multiplier work covers one block's loads and
// barrier waits.  HODOR_B8_EPT4: 2 x 512 threads with 4 elements each (<= 64 registers).
#ifndef HODOR_B8_EPT4
#define HODOR_B8_EPT4 0
#endif
template <int B>
struct PassOccupancy {
    static constexpr int EPT = (HODOR_B8_EPT4 && B == 8 && Groups<8>::RADIX4) ? 4 : 8;
    static constexpr int THREADS = (8 << B) / EPT;
    static constexpr int MIN_BLOCKS =
        B >= 9 ? 1 : (B == 8 ? (EPT == 4 ? 2 : (Groups<8>::RADIX4 ? 3 : 2)) : (B == 7 ? (Groups<7>::RADIX4 ? 6 : 4) : (Groups<6>::RADIX4 ? 12 : 8)));
};

// One group of the block-local NTT: every thread owns 8 elements = 8 >> LOGR butterflies of radix
// 2^LOGR whose members are 2^SL positions apart.  FROM_GLOBAL / TO_GLOBAL: the group reads its inputs
// from / writes its outputs to HBM instead of the shared tile.
//
// Code size matters here: with every multiply inlined and every loop unrolled a pass kernel is
// 210-260 KB of straight-line code, far beyond the instruction cache, and ncu attributes 7-13 % of
// the issue stalls to `no_instruction`.  So only the butterflies (whose twiddles are compile-time
// selected constants) are unrolled; every per-element table multiply -- coset scaling on load,
// inter-group twiddle, inter-pass twiddle on store -- runs in a rolled loop over the thread's own
// slots of the shared tile (owner-only accesses: no barrier needed), one multiplier body each.
#ifndef HODOR_ROLL_UNROLL
#define HODOR_ROLL_UNROLL 1  // table-multiply loops: 1 = fully rolled; 2 = two multiplies in flight per thread
#endif
#if HODOR_ROLL_UNROLL == 2
#define HODOR_ROLLED _Pragma("unroll 2")
#else
#define HODOR_ROLLED _Pragma("unroll 1")
#endif
template <class F, int B, int LOGR, int SL, int TWSH, bool FROM_GLOBAL, bool MUL_ON_LOAD, bool TO_GLOBAL, class LoadG,
          class StoreG>
DEV void ntt_group(const Field<F>& fld, const NttPass& p, uint4* sm, uint32_t tid, uint32_t oz, LoadG&& load_global,
                   StoreG&& store_global) {
    constexpr int R = 1 << LOGR;
    constexpr int EPT = PassOccupancy<B>::EPT;  // elements per thread
    static_assert(EPT >= R, "a thread holds at least one whole butterfly");
    constexpr int NB = EPT / R;
    constexpr uint32_t T = (8u << B) / EPT;  // threads per block
    constexpr uint32_t PLANE = 8u << B;
    constexpr uint32_t STEP = 8u << SL;  // slots between members of one butterfly
#pragma unroll 1
    for (int j = 0; j < NB; j++) {
        const uint32_t q = tid + j * T;
        const uint32_t c = q & 7u, rest = q >> 3;
        const uint32_t lo = rest & ((1u << SL) - 1u), hi = rest >> SL;
        const uint32_t base = (hi << (SL + LOGR)) | lo;
        const uint32_t slot0 = base * 8 + c;
        Fe x[R];
        if constexpr (FROM_GLOBAL && MUL_ON_LOAD) {
HODOR_ROLLED
            for (uint32_t d = 0; d < (uint32_t)R; d++) sts_fe(sm, PLANE, slot0 + d * STEP, load_global(base + (d << SL), c));
        }
#pragma unroll
        for (int d = 0; d < R; d++) {
            if constexpr (FROM_GLOBAL && !MUL_ON_LOAD) x[d] = load_global(base + ((uint32_t)d << SL), c);
            else x[d] = lds_fe(sm, PLANE, slot0 + d * STEP);
        }
        dif_inreg<F, LOGR>(fld, x, p, oz);
#pragma unroll
        for (int k = 0; k < R; k++) sts_fe(sm, PLANE, slot0 + k * STEP, x[bitrev_c(k, LOGR)]);
        if constexpr (SL > 0 || TO_GLOBAL) {
HODOR_ROLLED
            for (uint32_t k = TO_GLOBAL ? 0u : 1u; k < (uint32_t)R; k++) {
                Fe v = lds_fe(sm, PLANE, slot0 + k * STEP);
                if constexpr (SL > 0) {
                    if (k > 0) v = mul_by(fld, v, ld_pre(p.tw_b, (size_t)((k * lo) << TWSH)));
                }
                if constexpr (TO_GLOBAL) store_global(base + (k << SL), c, v);
                else sts_fe(sm, PLANE, slot0 + k * STEP, v);
            }
        }
    }
}

// position (digits kappa1 | kappa2 | kappa3, MSB first) -> local output index
template <int B>
DEV uint32_t local_out_index(uint32_t pos) {
    using G = Groups<B>;
    constexpr int R1 = G::R1, R2 = G::R2, R3 = G::R3, R4 = G::R4;
    const uint32_t k1 = pos >> (R2 + R3 + R4);
    const uint32_t k2 = (pos >> (R3 + R4)) & ((1u << R2) - 1u);
    const uint32_t k3 = (pos >> R4) & ((1u << R3) - 1u);
    const uint32_t k4 = pos & ((1u << R4) - 1u);
    return k1 | (k2 << R1) | (k3 << (R1 + R2)) | (k4 << (R1 + R2 + R3));
}

template <class F, int B, bool SCALE_IN, bool LAST>
__global__ void __launch_bounds__(PassOccupancy<B>::THREADS, PassOccupancy<B>::MIN_BLOCKS) ntt_pass_kernel(const __grid_constant__ NttPass p) {
    using G = Groups<B>;
    extern __shared__ uint4 sm[];
    const uint32_t tid = threadIdx.x;
    const uint32_t oz = tid & p.zero;
    const Field<F> fld(oz);
    const uint32_t ln = p.log_n;
    const size_t n = (size_t)1 << ln;

    // ---- tile geometry -------------------------------------------------------------------
    size_t in_base, out_base;     // element offsets that do not depend on (pos, c)
    uint32_t coset_hi = 0;        // LAST: high part of the coset index; else the coset index
    uint32_t k1_hi = 0, mid = 0;  // LAST only
    uint32_t col0 = 0;            // !LAST: first column of the tile
    const uint32_t li = p.log_l < 3 ? p.log_l : 3;  // coset bits inside c (LAST)
    if constexpr (!LAST) {
        // The coset is the fastest-varying part of the block index: the L cosets of one tile are
        // resident together, so pass 1 of an LDE fetches the shared coefficient tile from HBM once
        // and the other L-1 readers hit L2 (ncu before this change: 4.3 GB read for 0.54 GB of input).
        const uint32_t s = p.s;
        const uint32_t col_groups = 1u << (s - 3);
        const uint32_t tile = blockIdx.x >> p.log_l;
        const uint32_t u = tile >> (s - 3), cg = tile & (col_groups - 1u);
        coset_hi = blockIdx.x & ((1u << p.log_l) - 1u);
        col0 = cg * 8;
        in_base = ((size_t)u << (s + B)) + col0;
        out_base = (size_t)coset_hi * n + in_base;
        if constexpr (!SCALE_IN) in_base = out_base;  // per-coset data already in the work buffer
    } else {
        const uint32_t m = ln - B - p.b1;
        const uint32_t kb = 3 - li;
        uint32_t t = blockIdx.x;
        mid = t & ((1u << m) - 1u);
        t >>= m;
        k1_hi = t & ((1u << (p.b1 - kb)) - 1u);
        coset_hi = t >> (p.b1 - kb);
        in_base = 0;
        out_base = 0;
    }

    auto load_global = [&](uint32_t pos, uint32_t c) -> Fe {
        if constexpr (!LAST) {
            const size_t idx = in_base + ((size_t)pos << p.s) + c;
            Fe v = ld_fe(p.in, idx);
            if constexpr (SCALE_IN) {
                // j = idx (u == 0 in pass 1): a[j] * shift_i^j
                if (p.coset_full != nullptr)
                    v = mul_by(fld, v, ld_pre(p.coset_full, (size_t)coset_hi * n + idx));
                else
                    v = fld.mul(v, two_level_pow(fld, p.coset, (size_t)coset_hi * p.coset_stride_lo,
                                                 (size_t)coset_hi * p.coset_stride_hi, idx));
            }
            return v;
        } else {
            const uint32_t m = ln - B - p.b1;
            const uint32_t i = (coset_hi << li) | (c & ((1u << li) - 1u));
            const uint32_t k1 = (k1_hi << (3 - li)) | (c >> li);
            const size_t row = ((size_t)k1 << m) | mid;
            return ld_fe(p.in, (size_t)i * n + (row << B) + pos);
        }
    };
    auto store_global = [&](uint32_t pos, uint32_t c, Fe v) {
        const uint32_t kloc = local_out_index<B>(pos);
        if constexpr (!LAST) {
            // omega_N^(kloc * r): one lookup when the sub-transform is short enough for a flat table
            // (2 MiB at N = 2^16, L2 resident), else hi * lo from the two-level tables (one more multiply)
            const uint64_t prod = (uint64_t)kloc * (col0 + c);
            if (p.tw_full != nullptr) v = mul_by(fld, v, ld_pre(p.tw_full, ((size_t)kloc << p.s) + col0 + c));
            else if (p.tw_direct != nullptr) v = mul_by(fld, v, ld_pre(p.tw_direct, (size_t)prod));
            else v = fld.mul(v, two_level_pow(fld, p.tw, 0, 0, prod << p.tw_shift));
            st_fe(p.out, out_base + ((size_t)kloc << p.s) + c, v);
        } else {
            const uint32_t i = (coset_hi << li) | (c & ((1u << li) - 1u));
            const uint32_t k1 = (k1_hi << (3 - li)) | (c >> li);
            // middle digits: mid = (k2 | k3) MSB first -> k2 + 2^mid0 * k3
            uint32_t midrev = mid;
            if (p.mid1) midrev = (mid >> p.mid1) | ((mid & ((1u << p.mid1) - 1u)) << p.mid0);
            const size_t k = (size_t)k1 | ((size_t)midrev << p.b1) | ((size_t)kloc << (ln - B));
            if (p.flags & PASS_OUT_CONST) v = mul_by(fld, v, ld_param(p.out_const, oz));
            if (p.flags & PASS_OUT_POW) {
                if (p.out_pow_full != nullptr) v = mul_by(fld, v, ld_pre(p.out_pow_full, k));
                else v = fld.mul(v, two_level_pow(fld, p.out_pow, 0, 0, k));
            }
            if (p.peer_on) {
                const size_t mask = ((size_t)1 << p.peer_chunk_log) - 1;
                st_fe(p.peer[k >> p.peer_chunk_log], ((size_t)p.peer_rank << p.peer_chunk_log) + (k & mask), v);
            } else {
                st_fe(p.out, (size_t)i + (k << p.log_l), v);
            }
        }
    };

    constexpr int R1 = G::R1, R2 = G::R2, R3 = G::R3, R4 = G::R4;
#if HODOR_PASS_PREFETCH
    // Every address this thread will touch in HBM is known now: ask L2 for its 8 inputs (and their coset-scaling
    // entries) and for the 8 inter-pass twiddle entries it multiplies by at the very end, so that the demand loads
    // of the rolled loops find them on chip (L2 hit ~250 cycles instead of ~700+ from DRAM).
    {
        constexpr int EPT = PassOccupancy<B>::EPT;
        constexpr uint32_t T = (8u << B) / EPT;
        if constexpr (!LAST) {
            constexpr int SL1 = B - R1;
#pragma unroll 1
            for (int j = 0; j < (EPT >> R1); j++) {
                const uint32_t q = tid + j * T, c = q & 7u, rest = q >> 3;
                const uint32_t lo = rest & ((1u << SL1) - 1u), hi = rest >> SL1;
                const uint32_t base = (hi << (SL1 + R1)) | lo;
#pragma unroll
                for (int d = 0; d < (1 << R1); d++) {
                    const size_t idx = in_base + ((size_t)(base + ((uint32_t)d << SL1)) << p.s) + c;
                    prefetch_l2(p.in + 2 * idx);
                    if constexpr (SCALE_IN) {
                        if (p.coset_full != nullptr) prefetch_l2(p.coset_full + 4 * ((size_t)coset_hi * n + idx));
                    }
                }
            }
            constexpr int RL = G::NGROUPS == 2 ? R2 : (G::NGROUPS == 3 ? R3 : R4);
            if (p.tw_full != nullptr || p.tw_direct != nullptr) {
#pragma unroll 1
                for (int j = 0; j < (EPT >> RL); j++) {
                    const uint32_t q = tid + j * T, c = q & 7u, base = (q >> 3) << RL;
#pragma unroll
                    for (int k = 0; k < (1 << RL); k++) {
                        const uint32_t kloc = local_out_index<B>(base + k);
                        if (p.tw_full != nullptr) prefetch_l2(p.tw_full + 4 * (((size_t)kloc << p.s) + col0 + c));
                        else prefetch_l2(p.tw_direct + 4 * ((size_t)kloc * (col0 + c)));
                    }
                }
            }
        }
    }
#endif
    ntt_group<F, B, R1, B - R1, 0, true, SCALE_IN, false>(fld, p, sm, tid, oz, load_global, store_global);
    __syncthreads();
    if constexpr (G::NGROUPS == 2) {
        ntt_group<F, B, R2, 0, R1, false, false, true>(fld, p, sm, tid, oz, load_global, store_global);
    } else if constexpr (G::NGROUPS == 3) {
        ntt_group<F, B, R2, R3, R1, false, false, false>(fld, p, sm, tid, oz, load_global, store_global);
        __syncthreads();
        ntt_group<F, B, R3, 0, R1 + R2, false, false, true>(fld, p, sm, tid, oz, load_global, store_global);
    } else {
        ntt_group<F, B, R2, R3 + R4, R1, false, false, false>(fld, p, sm, tid, oz, load_global, store_global);
        __syncthreads();
        ntt_group<F, B, R3, R4, R1 + R2, false, false, false>(fld, p, sm, tid, oz, load_global, store_global);
        __syncthreads();
        ntt_group<F, B, R4, 0, R1 + R2 + R3, false, false, true>(fld, p, sm, tid, oz, load_global, store_global);
    }
    if constexpr (LAST) {
        if (p.peer_on == 1) __threadfence_system();  // optional (HODOR_PEER_FENCE=1), see Ops::ntt
    }
}

// ------------------------------------------------------------------------------------------------
// Small transforms (log_n <= 11): one block per coset, radix-2 DIF in shared memory.
// ------------------------------------------------------------------------------------------------
struct SmallNtt {
    const uint4* in;
    uint4* out;
    uint32_t log_n, log_l, flags, zero;
    uint32_t coset_stride_lo, coset_stride_hi;
    const uint4* tw;  // omega^e, e in [0, n/2]
    TwoLevel coset;   // null lo => no input scaling
    TwoLevel out_pow;
    Fe out_const;
};

template <class F>
__global__ void __launch_bounds__(1024) ntt_small_kernel(const __grid_constant__ SmallNtt p) {
    extern __shared__ uint4 sm[];
    const uint32_t tid = threadIdx.x, nt = blockDim.x;
    const uint32_t oz = tid & p.zero;
    const Field<F> fld(oz);
    const uint32_t ln = p.log_n, n = 1u << ln, coset = blockIdx.x;
    for (uint32_t j = tid; j < n; j += nt) {
        Fe v = ld_fe(p.in, j);
        if (p.coset.lo != nullptr)
            v = fld.mul(v, two_level_pow(fld, p.coset, (size_t)coset * p.coset_stride_lo,
                                         (size_t)coset * p.coset_stride_hi, j));
        sts_elem(sm, j, v);
    }
    __syncthreads();
    for (uint32_t st = 0; st < ln; st++) {
        const uint32_t half = n >> (st + 1);
        for (uint32_t q = tid; q < n / 2; q += nt) {
            const uint32_t i = q & (half - 1u), blk = (q / half) * 2 * half;
            const Fe u = lds_elem(sm, blk + i), v = lds_elem(sm, blk + i + half);
            sts_elem(sm, blk + i, fld.add(u, v));
            Fe d = fld.sub(u, v);
            if (i != 0) d = fld.mul(d, ld_fe(p.tw, (size_t)i << st));
            sts_elem(sm, blk + i + half, d);
        }
        __syncthreads();
    }
    for (uint32_t j = tid; j < n; j += nt) {
        const uint32_t k = ln ? (__brev(j) >> (32 - ln)) : 0u;
        Fe v = lds_elem(sm, j);
        if (p.flags & PASS_OUT_CONST) v = fld.mul(v, ld_param(p.out_const, oz));
        if (p.flags & PASS_OUT_POW) v = fld.mul(v, two_level_pow(fld, p.out_pow, 0, 0, k));
        st_fe(p.out, (size_t)coset + ((size_t)k << p.log_l), v);
    }
}

// ------------------------------------------------------------------------------------------------
// Tables and elementwise helpers
// ------------------------------------------------------------------------------------------------
// out[b * count + i] = bases[b]^i * (scale ? scale[0] : 1)
// PRE: entries in FePre form (64 B) for tables that are read directly as multipliers
template <class F, bool PRE = false>
__global__ void pow_table_kernel(uint4* out, const Fe* bases, const Fe* scale, uint32_t count, uint32_t zero) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const Field<F> fld(threadIdx.x & zero);
    Fe r = fld.pow(bases[blockIdx.y], i);
    if (scale != nullptr) r = fld.mul(r, scale[0]);
    if constexpr (PRE) {
        FePre m;
        fld.make_pre(r, m.w, m.q);
        st_pre(out, (size_t)blockIdx.y * count + i, m);
    } else {
        st_fe(out, (size_t)blockIdx.y * count + i, r);
    }
}

// Expanded tables.  out[b * n + j] = bases_b^j  (coset scaling for every coset b), or with
// `boundary_s` >= 0: out[idx] = omega^((idx >> s) * (idx & (2^s - 1))), the pass-1 inter-pass twiddle
// stored at the address of the element it multiplies.  Entries are FePre (64 B).
template <class F>
__global__ void expand_table_kernel(uint4* out, TwoLevel t, uint32_t stride_lo, uint32_t stride_hi, size_t n,
                                    int boundary_s, uint32_t zero) {
    const Field<F> fld(threadIdx.x & zero);
    const size_t b = blockIdx.y;
    for (size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (size_t)gridDim.x * blockDim.x) {
        uint64_t e = j;
        if (boundary_s >= 0) e = (uint64_t)(j >> boundary_s) * (j & (((size_t)1 << boundary_s) - 1));
        FePre m;
        fld.make_pre(two_level_pow(fld, t, b * stride_lo, b * stride_hi, e), m.w, m.q);
        st_pre(out, b * n + j, m);
    }
}

// a[j] <- a[j] * c * g^j     (distribute_powers, src/fft/mod.rs:110-123; c folded into pw.lo)
template <class F>
__global__ void scale_pow_kernel(uint4* a, size_t n, TwoLevel pw, uint32_t zero) {
    const Field<F> fld(threadIdx.x & zero);
    for (size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (size_t)gridDim.x * blockDim.x) {
        const Fe v = ld_fe(a, j);
        st_fe(a, j, fld.mul(v, two_level_pow(fld, pw, 0, 0, j)));
    }
}

enum : int {
    EW_MUL = 0, EW_ADD = 1, EW_SUB = 2, EW_SCALE = 3,                              // a (op) b; SCALE: b is one element
    EW_ADD_SCALED = 4, EW_ADD_CONST = 5, EW_NEGATE = 6, EW_SQUARE = 7, EW_POW = 8,  // scalar / exponent forms
    EW_NUM_OPS = 9
};
// elementwise Polynomial ops (src/polynomials/mod.rs:59-83, 640-683, 744-771, 817-887).  `scalar` is a
// fixed operand for the whole launch (scale, add_assign_scaled), so it goes through mul_pre.
//   MUL a*b | ADD a+b | SUB a-b | SCALE a*b[0] | ADD_SCALED a + b*scalar | ADD_CONST a + scalar |
//   NEGATE -a | SQUARE a^2 | POW a^exp
template <class F>
__global__ void elementwise_kernel(int op, const uint4* a, const uint4* b, uint4* out, size_t n,
                                   const __grid_constant__ Fe scalar, uint64_t exp, uint32_t zero) {
    const uint32_t oz = threadIdx.x & zero;
    const Field<F> fld(oz);
    FePre sp;
    if (op == EW_SCALE) {
        fld.make_pre(ld_fe(b, 0), sp.w, sp.q);
    } else if (op == EW_ADD_SCALED) {
        fld.make_pre(ld_param(scalar, oz), sp.w, sp.q);
    }
    for (size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (size_t)gridDim.x * blockDim.x) {
        const Fe x = ld_fe(a, j);
        Fe r;
        switch (op) {
            case EW_MUL: r = fld.mul(x, ld_fe(b, j)); break;
            case EW_ADD: r = fld.add(x, ld_fe(b, j)); break;
            case EW_SUB: r = fld.sub(x, ld_fe(b, j)); break;
            case EW_SCALE: r = mul_by(fld, x, sp); break;
            case EW_ADD_SCALED: r = fld.add(x, mul_by(fld, ld_fe(b, j), sp)); break;
            case EW_ADD_CONST: r = fld.add(x, ld_param(scalar, oz)); break;
            case EW_NEGATE: r = fld.neg(x); break;
            case EW_SQUARE: r = fld.mul(x, x); break;
            default: r = fld.pow(x, exp); break;  // EW_POW
        }
        st_fe(out, j, r);
    }
}

// Points of a coset and two maps over them (the setup work of `Prover::new`):
//   x_j = first * ratio^j                                   src/precomputations/mod.rs:14-66 (omegas / coset / omegas_inv)
//   CM_LINEAR   x_j - c                                     src/ali/per_register/mod.rs:214-224 (X - omega^row on the coset)
//   CM_DIVISOR  inv_van[j & van_mask] * prod_r (x_j - roots[r])   src/ali/per_register/mod.rs:60-162: x_j^T - 1 takes only
//               E/T distinct values on the evaluation coset, so their inverses come in as a small table
enum : int { CM_POINTS = 0, CM_LINEAR = 1, CM_DIVISOR = 2 };
template <class F>
__global__ void coset_map_kernel(int mode, uint4* out, size_t n, TwoLevel pw, const __grid_constant__ Fe first,
                                 const __grid_constant__ Fe c, const uint4* roots, uint32_t num_roots, const uint4* inv_van,
                                 uint32_t van_mask, uint32_t zero) {
    const uint32_t oz = threadIdx.x & zero;
    const Field<F> fld(oz);
    FePre fp;
    fld.make_pre(ld_param(first, oz), fp.w, fp.q);
    for (size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (size_t)gridDim.x * blockDim.x) {
        Fe x = mul_by(fld, two_level_pow(fld, pw, 0, 0, j), fp);
        if (mode == CM_LINEAR) {
            x = fld.sub(x, ld_param(c, oz));
        } else if (mode == CM_DIVISOR) {
            Fe d = ld_fe(inv_van, j & van_mask);
            for (uint32_t r = 0; r < num_roots; r++) d = fld.mul(d, fld.sub(x, ld_fe(roots, r)));
            x = d;
        }
        st_fe(out, j, x);
    }
}


}  // namespace hodor
