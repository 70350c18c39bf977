// FRI fold for sm_100a: one layer of NaiveFriIop::proof_from_lde_by_values,
// src/fri/fri_on_values.rs:61-101.
//
//   next[idx] = ((v[idx] + v[idx+M/2]) + c * (v[idx] - v[idx+M/2]) * omega_N^(-idx * 2^layer)) * 2^-1
//
// The reference reads omega_N^-e from a precomputed N/2-entry table (:24-40) and multiplies by a
// precomputed 2^-1; here the power comes from a two-level table (2 * sqrt(N) entries, L2 resident)
// and the halving is a shift with a conditional add of p (exact, because halving is linear and so
// commutes with the Montgomery factor).  The challenge c is read from device memory, where the
// Merkle tail kernel of the previous layer left it, so the whole commit chain runs without a host
// round trip.
#pragma once
#include "field.cuh"
#include "merkle.cuh"
#include "ntt.cuh"

namespace hodor {

// `in` holds, at local position t, element  idx(t) = ((t >> blk_log) * idx_stride << blk_log) + idx_offset + (t & (2^blk_log - 1))
// of the layer: blocks of 2^blk_log adjacent elements dealt round-robin (blk_log 0: v[idx_offset + idx_stride * t]).
// (offset 0, stride 1: the whole
// layer; offset r, stride G: rank r's cyclic slice of a layer sharded over G GPUs, whose fold pairs
// (t, t + half) are then both local).
// FLAT: the twiddle omega_N^-e is one 64-byte fixed-operand entry of a flat N/2-entry table (the
// reference's own `omegas_inv`, :24-40) instead of hi * lo from the two-level table; both remaining
// multiplies then go through Field::mul_pre (the challenge is converted once per thread): 1.3
// Montgomery-multiply equivalents per output instead of 3.
template <class F, bool FLAT>
__global__ void __launch_bounds__(256) fri_fold_kernel(const uint4* in, uint4* out, size_t half, TwoLevel winv,
                                                       const uint4* winv_flat, uint32_t layer, const uint4* challenge,
                                                       uint64_t idx_offset, uint64_t idx_stride, uint32_t blk_log, uint32_t zero) {
    const Field<F> fld(threadIdx.x & zero);
    const Fe c = ld_fe(challenge, 0);
    FePre cp;
    if constexpr (FLAT) fld.make_pre(c, cp.w, cp.q);
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < half; t += (size_t)gridDim.x * blockDim.x) {
        const size_t idx = t;
        const Fe f0 = ld_fe(in, idx), f1 = ld_fe(in, idx + half);
        const Fe even = fld.add(f0, f1);
        Fe odd = fld.sub(f0, f1);
        const uint64_t e = ((((uint64_t)t >> blk_log) * idx_stride << blk_log) + idx_offset + ((uint64_t)t & (((uint64_t)1 << blk_log) - 1)))
                           << layer;
        if constexpr (FLAT) {
            odd = mul_by(fld, odd, ld_pre(winv_flat, (size_t)e));
            odd = mul_by(fld, odd, cp);
        } else {
            odd = fld.mul(odd, two_level_pow(fld, winv, 0, 0, e));
            odd = fld.mul(odd, c);
        }
        st_fe(out, idx, fld.halve(fld.add(odd, even)));
    }
}

// One layer fused with the bottom three levels of its Merkle tree (src/fri/fri_on_values.rs:61-119: fold, then
// I::create(next_values)).  A thread owns 8 consecutive outputs: it folds them one by one, stores each value,
// hashes it while it is still in registers and combines the hashes as a thread-serial 2^3 subtree, writing the
// node levels half/2, half/4, half/8 of the new tree.  The layer's values are therefore read by no hashing
// kernel (32 * half bytes of HBM reads and one launch per layer saved) and the fold's multiplier work (2
// fixed-operand multiplies per leaf) issues between the ALU-bound compressions of other warps.
#ifndef HODOR_FOLD_COMMIT_MINBLOCKS
#define HODOR_FOLD_COMMIT_MINBLOCKS 1  // 3: cap the kernel at 85 registers (A/B build, profiles/r02_experiments.md)
#endif
template <class F, bool FLAT>
__global__ void __launch_bounds__(256, HODOR_FOLD_COMMIT_MINBLOCKS) fri_fold_commit_kernel(const uint4* in, uint4* out, uint4* nodes, size_t half, TwoLevel winv,
                                                              const uint4* winv_flat, uint32_t layer, const uint4* challenge,
                                                              const __grid_constant__ B2sState key, uint32_t zero) {
    const Field<F> fld(threadIdx.x & zero);
    const Fe c = ld_fe(challenge, 0);
    FePre cp;
    if constexpr (FLAT) fld.make_pre(c, cp.w, cp.q);
    const size_t groups = half >> 3;
    for (size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x; g < groups; g += (size_t)gridDim.x * blockDim.x) {
        auto leaf = [&](size_t idx) -> Digest {
            const Fe f0 = ld_fe(in, idx), f1 = ld_fe(in, idx + half);
            const Fe even = fld.add(f0, f1);
            Fe odd = fld.sub(f0, f1);
            const uint64_t e = (uint64_t)idx << layer;
            if constexpr (FLAT) {
                odd = mul_by(fld, odd, ld_pre(winv_flat, (size_t)e));
                odd = mul_by(fld, odd, cp);
            } else {
                odd = fld.mul(odd, two_level_pow(fld, winv, 0, 0, e));
                odd = fld.mul(odd, c);
            }
            const Fe v = fld.halve(fld.add(odd, even));
            st_fe(out, idx, v);
            Digest d;
#pragma unroll
            for (int i = 0; i < 8; i++) d.w[i] = v.v[i];
            return tree_hash_leaf(key, d);
        };
        merkle_subtree_fn<3>(key, nodes, half, g << 3, leaf);
    }
}

}  // namespace hodor
