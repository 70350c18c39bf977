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
// I::create(next_values)), second design.  The first one (a thread folded its own 8 consecutive outputs inside the
// hashing recursion: 128 registers, 256-byte strided loads) was slower than the separate kernels
// (profiles/r02_experiments.md).  Here a block folds 2048 consecutive outputs with the plain kernel's coalesced
// accesses -- thread t takes outputs t, t + 256, ... --, stores them to HBM and to a 64 KiB shared tile, and after one
// barrier thread t hashes outputs 8t .. 8t + 7 from the tile as a thread-serial 2^3 subtree (merkle.cuh), writing the
// node levels half/2, half/4, half/8: the structure of the fused last pass of the transform (ntt_commit.cuh), at its
// residency (3 blocks of 256 threads per SM, <= 85 registers).  The layer's values are read by no hashing kernel, and
// the fold's multiplier work issues beside the ALU-bound compressions of the other resident blocks.
// half must be a multiple of 2048.
constexpr uint32_t FOLD_COMMIT_TILE = 2048;
template <class F, bool FLAT>
__global__ void __launch_bounds__(256, 3) fri_fold_commit_kernel(const uint4* in, uint4* out, uint4* nodes, size_t half, TwoLevel winv,
                                                              const uint4* winv_flat, uint32_t layer, const uint4* challenge,
                                                              const __grid_constant__ B2sState key, uint32_t zero) {
    extern __shared__ uint4 tile[];  // element e of the block's 2048 outputs at tile[2e], tile[2e + 1]
    const uint32_t tid = threadIdx.x;
    const Field<F> fld(tid & zero);
    const size_t base = (size_t)blockIdx.x * FOLD_COMMIT_TILE;
    {
        const Fe c = ld_fe(challenge, 0);
        FePre cp;
        if constexpr (FLAT) fld.make_pre(c, cp.w, cp.q);
#pragma unroll 1
        for (uint32_t j = 0; j < FOLD_COMMIT_TILE / 256; j++) {
            const uint32_t e = j * 256 + tid;
            const size_t idx = base + e;
            const Fe f0 = ld_fe(in, idx), f1 = ld_fe(in, idx + half);
            const Fe even = fld.add(f0, f1);
            Fe odd = fld.sub(f0, f1);
            const uint64_t ex = (uint64_t)idx << layer;
            if constexpr (FLAT) {
                odd = mul_by(fld, odd, ld_pre(winv_flat, (size_t)ex));
                odd = mul_by(fld, odd, cp);
            } else {
                odd = fld.mul(odd, two_level_pow(fld, winv, 0, 0, ex));
                odd = fld.mul(odd, c);
            }
            const Fe v = fld.halve(fld.add(odd, even));
            st_fe(out, idx, v);
            sts_elem(tile, e, v);
        }
    }
    __syncthreads();
    const size_t first = base + 8 * (size_t)tid;
    auto leaf = [&](size_t idx) -> Digest {
        const uint32_t e = (uint32_t)(idx - base);
        const uint4 a = tile[2 * e], b = tile[2 * e + 1];
        Digest d;
        d.w[0] = a.x; d.w[1] = a.y; d.w[2] = a.z; d.w[3] = a.w;
        d.w[4] = b.x; d.w[5] = b.y; d.w[6] = b.z; d.w[7] = b.w;
        return tree_hash_leaf(key, d);
    };
    merkle_subtree_fn<3>(key, nodes, half, first, leaf);
}

}  // namespace hodor
