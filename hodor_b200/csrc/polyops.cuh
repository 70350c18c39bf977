// Polynomial helpers next to the transforms: batch inversion and point evaluation.
//
// Replaces (same result bits)
//   Polynomial<F, Values>::batch_inversion          src/polynomials/mod.rs:889-954
//   Polynomial<F, Coefficients>::evaluate_at        src/polynomials/mod.rs:685-711
//
// batch_inversion is Montgomery's trick as a tree instead of the reference's per-worker chunks: at
// every level thread t of M multiplies up the K elements x[t + M*k] (coalesced across threads),
// keeps the running products, and hands its total to the next level; one Fermat inversion at the
// top; the way back down turns "inverse of my total" into the K inverses with two multiplications
// per element.  3 + 3/(K-1) multiplications per element in all.  A zero anywhere makes the top
// product zero: the status word is set and nothing is written (the reference returns
// Err(SynthesisError::Error) before touching the vector, :919).
#pragma once
#include <cuda_runtime.h>
#include "ntt.cuh"

namespace hodor {

constexpr uint32_t BINV_K = 16;

template <class F>
__global__ void __launch_bounds__(256) binv_up_kernel(const uint4* x, uint4* pre, uint4* tot, size_t m, size_t M,
                                                       uint32_t zero) {
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= M) return;
    const Field<F> fld(threadIdx.x & zero);
    Fe acc = ld_fe(x, t);
    st_fe(pre, t, acc);
    for (uint32_t k = 1; k < BINV_K; k++) {
        const size_t idx = t + M * k;
        if (idx >= m) break;
        acc = fld.mul(acc, ld_fe(x, idx));
        st_fe(pre, idx, acc);
    }
    st_fe(tot, t, acc);
}

// top of the tree: x[0] <- x[0]^-1 (a^(p-2)), or *status = 1 when it is zero
template <class F>
__global__ void binv_top_kernel(uint4* x, int* status, uint32_t zero) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const Field<F> fld(threadIdx.x & zero);
    const Fe a = ld_fe(x, 0);
    if (Field<F>::eq(a, Field<F>::zero())) {
        *status = 1;
        return;
    }
    *status = 0;
    uint32_t e[8];
#pragma unroll
    for (int i = 0; i < 8; i++) e[i] = F::P(i);
    uint32_t borrow = 2u;  // e = p - 2
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const uint32_t old = e[i];
        e[i] = old - borrow;
        borrow = old < borrow ? 1u : 0u;
    }
    Fe acc = Field<F>::one();
    for (int i = 255; i >= 0; i--) {
        acc = fld.mul(acc, acc);
        if ((e[i >> 5] >> (i & 31)) & 1u) acc = fld.mul(acc, a);
    }
    st_fe(x, 0, acc);
}

// tot[t] holds the inverse of thread t's total; out[t + M*k] <- x[t + M*k]^-1  (out may alias x)
template <class F>
__global__ void __launch_bounds__(256) binv_down_kernel(const uint4* x, const uint4* pre, const uint4* tot, uint4* out,
                                                         size_t m, size_t M, const int* status, uint32_t zero) {
    if (*status != 0) return;
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= M) return;
    const Field<F> fld(threadIdx.x & zero);
    Fe s = ld_fe(tot, t);
    uint32_t kmax = 0;
    for (uint32_t k = 1; k < BINV_K; k++)
        if (t + M * k < m) kmax = k;
    for (uint32_t k = kmax; k >= 1; k--) {
        const size_t idx = t + M * k;
        const Fe xv = ld_fe(x, idx);
        const Fe g = ld_fe(pre, idx - M);
        st_fe(out, idx, fld.mul(g, s));
        s = fld.mul(s, xv);
    }
    st_fe(out, t, s);
}

// evaluate_at: partial[b] = sum over the block's threads of g^t * sum_k a[t + M*k] * (g^M)^k
template <class F>
__global__ void __launch_bounds__(256) eval_partial_kernel(const uint4* a, size_t n, size_t M, uint32_t K,
                                                           const __grid_constant__ Fe g, const __grid_constant__ FePre gM,
                                                           uint4* partial, uint32_t zero) {
    __shared__ uint4 sm[2 * 256];
    const uint32_t oz = threadIdx.x & zero;
    const Field<F> fld(oz);
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    Fe acc = Field<F>::zero();
    if (t < M) {
        uint32_t kmax = 0;
        for (uint32_t k = 1; k < K; k++)
            if (t + M * k < n) kmax = k;
        acc = ld_fe(a, t + M * kmax);
        const FePre y = ld_param(gM, oz);
        for (uint32_t k = kmax; k >= 1; k--) acc = fld.add(mul_by(fld, acc, y), ld_fe(a, t + M * (k - 1)));
        acc = fld.mul(acc, fld.pow(ld_param(g, oz), (uint64_t)t));
    }
    sts_elem(sm, threadIdx.x, acc);
    __syncthreads();
    for (uint32_t h = 128; h >= 1; h >>= 1) {
        if (threadIdx.x < h) sts_elem(sm, threadIdx.x, fld.add(lds_elem(sm, threadIdx.x), lds_elem(sm, threadIdx.x + h)));
        __syncthreads();
    }
    if (threadIdx.x == 0) st_fe(partial, blockIdx.x, lds_elem(sm, 0));
}

template <class F>
__global__ void __launch_bounds__(256) eval_final_kernel(const uint4* partial, size_t count, uint4* out, uint32_t zero) {
    __shared__ uint4 sm[2 * 256];
    const Field<F> fld(threadIdx.x & zero);
    Fe acc = Field<F>::zero();
    for (size_t i = threadIdx.x; i < count; i += 256) acc = fld.add(acc, ld_fe(partial, i));
    sts_elem(sm, threadIdx.x, acc);
    __syncthreads();
    for (uint32_t h = 128; h >= 1; h >>= 1) {
        if (threadIdx.x < h) sts_elem(sm, threadIdx.x, fld.add(lds_elem(sm, threadIdx.x), lds_elem(sm, threadIdx.x + h)));
        __syncthreads();
    }
    if (threadIdx.x == 0) st_fe(out, 0, lds_elem(sm, 0));
}

// Self-test of Field::mul_pre: a chain of multiplications through the Montgomery multiplier, through
// mul_pre, and through mul_pre with the guard threshold forced down to 2^31 (so the out-of-line
// carry fix-up runs on about half of all multiplies); *mismatch counts threads that disagree.
template <class F>
__global__ void __launch_bounds__(256) selftest_mul_pre_kernel(unsigned long long* mismatch, uint32_t zero) {
    const uint32_t oz = threadIdx.x & zero;
    const Field<F> fld(oz);
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    Fe g = Field<F>::zero();
    g.v[0] = F::GENERATOR;
    g = fld.to_mont(g);
    Fe x = fld.pow(g, 0x9e3779b9ull * (tid + 1));
    Fe w = fld.pow(g, 0x85ebca6bull * (tid + 3));
    if (tid == 0) w = fld.neg(Field<F>::one());  // p - 1 in Montgomery form
    if (tid == 1) w = Field<F>::one();
    if (tid == 2) x = Field<F>::zero();
    FePre m;
    fld.make_pre(w, m.w, m.q);
    Fe r1 = x, r2 = x, r3 = x;
    bool bad = false;
    for (int i = 0; i < 48; i++) {
        r1 = fld.mul(r1, w);
        r2 = fld.mul_pre(r2, m.w, m.q);
        r3 = fld.template mul_pre<0x80000000u>(r3, m.w, m.q);
        bad |= !Field<F>::eq(r1, r2) || !Field<F>::eq(r1, r3) || !fld.is_canonical(r2);
        w = fld.add(w, r1);  // a fresh multiplier every step
        fld.make_pre(w, m.w, m.q);
    }
    if (bad) atomicAdd(mismatch, 1ull);
}

}  // namespace hodor
