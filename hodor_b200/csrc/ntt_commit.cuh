// Last pass of an NTT / coset LDE fused with the bottom three levels of the Merkle tree over its output:
// `let lde = w.lde(&worker, lde_factor)?; let oracle = I::create(lde.as_ref());` (src/prover/mod.rs:73-80,
// tree: src/iop/blake2s_trivial_iop.rs:131-219) without a second kernel re-reading the 2^27 values.
//
// Why it can pay when the separate kernels are both near their pipe ceilings: the transform is bound by the
// multiplier pipe (72-78 % busy, ALU 41 %), the hashing by the ALU pipe (88 %).  As two kernels they share SMs
// badly (profiles/r02_experiments.md: every split loses more occupancy than the other pipe gives back).  Here
// one block does both, one after the other, at the SAME residency as the plain last pass (3 blocks x 256
// threads at <= 85 registers, one 64 KiB tile): the three resident blocks of an SM drift out of phase, so at
// any moment some warps issue multiplier work and others hashing work.
//
// The last pass takes the 8 cosets (or 8 adjacent outputs) of one row as its 8 columns, so the 8 columns of one
// tile position ARE 8 adjacent leaves of the output -- one aligned 2^3 subtree.  The last butterfly group stores
// every output to HBM as before; the shared tile still holds them (position-major), so after one barrier thread t
// takes position t, reads its 8 leaves back from the tile one at a time and runs the thread-serial subtree of
// merkle.cuh on them: 8 leaf + 7 node compressions, node levels nL/2, nL/4, nL/8 written.  The rest of the tree
// (merkle_upper_levels + tail) is enqueued by the host as for the unfused build.
//
// Measured (profiles/r02_experiments.md): 16.66 ms against 6.70 + 10.80 ms for the two separate kernels at 2^24 x 8, 0.5 ms
// (1.3-1.8 %) per lift-and-commit of that shape in every configuration; for plans that end in a 7- or 6-bit digit
// (128- / 64-thread blocks) the fused kernel is no faster than its two parts and loses against the separate tree, which
// runs beside the next polynomial's transform.  HODOR_FUSE_LAST_COMMIT (read at hodor_cuda_init): 1, the default,
// takes this kernel when the last digit is 8; 2 also for 7; 3 also for 6; 0 never.
#pragma once
#include "merkle.cuh"
#include "ntt.cuh"

namespace hodor {

template <class F, int B>
__global__ void __launch_bounds__(PassOccupancy<B>::THREADS, PassOccupancy<B>::MIN_BLOCKS)
    ntt_last_commit_kernel(const __grid_constant__ NttPass p, uint4* nodes, const __grid_constant__ B2sState key) {
    using G = Groups<B>;
    static_assert(PassOccupancy<B>::THREADS == (1 << B), "one thread per tile position in the hashing phase");
    extern __shared__ uint4 sm[];
    constexpr uint32_t PLANE = 8u << B;
    const uint32_t tid = threadIdx.x;
    const uint32_t oz = tid & p.zero;
    const Field<F> fld(oz);
    const uint32_t ln = p.log_n;
    const size_t n = (size_t)1 << ln;

    // tile geometry of a last pass (ntt.cuh, ntt_pass_kernel<.., LAST = true>)
    const uint32_t li = p.log_l < 3 ? p.log_l : 3;  // coset bits inside the column index
    const uint32_t m = ln - B - p.b1;
    uint32_t t = blockIdx.x;
    const uint32_t mid = t & ((1u << m) - 1u);
    t >>= m;
    const uint32_t k1_hi = t & ((1u << (p.b1 - (3 - li))) - 1u);
    const uint32_t coset_hi = t >> (p.b1 - (3 - li));
    uint32_t midrev = mid;  // middle digits: mid = (k2 | k3) MSB first -> k2 + 2^mid0 * k3
    if (p.mid1) midrev = (mid >> p.mid1) | ((mid & ((1u << p.mid1) - 1u)) << p.mid0);

    auto load_global = [&](uint32_t pos, uint32_t c) -> Fe {
        const uint32_t i = (coset_hi << li) | (c & ((1u << li) - 1u));
        const uint32_t k1 = (k1_hi << (3 - li)) | (c >> li);
        const size_t row = ((size_t)k1 << m) | mid;
        return ld_fe(p.in, (size_t)i * n + (row << B) + pos);
    };
    // index of output (pos, c) in the interleaved natural-order vector: for fixed pos the columns c = 0..7 are
    // 8 adjacent, 8-aligned indices (the coset bits, then the low bits of k1, are the low bits of the index)
    auto out_index = [&](uint32_t pos, uint32_t c) -> size_t {
        const uint32_t kloc = local_out_index<B>(pos);
        const uint32_t i = (coset_hi << li) | (c & ((1u << li) - 1u));
        const uint32_t k1 = (k1_hi << (3 - li)) | (c >> li);
        const size_t k = (size_t)k1 | ((size_t)midrev << p.b1) | ((size_t)kloc << (ln - B));
        return (size_t)i + (k << p.log_l);
    };
    // No output scaling here (the host only takes this kernel for plain / coset LDE outputs), so the value the last
    // group stores for (pos, c) is exactly what it has just read from slot pos * 8 + c of the tile: after the
    // barrier below the tile holds the block's outputs, position-major.
    auto store_global = [&](uint32_t pos, uint32_t c, Fe v) { st_fe(p.out, out_index(pos, c), v); };

    constexpr int R1 = G::R1, R2 = G::R2, R3 = G::R3, R4 = G::R4;
    ntt_group<F, B, R1, B - R1, 0, true, false, false>(fld, p, sm, tid, oz, load_global, store_global);
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
    __syncthreads();

    // hashing phase: thread t <-> tile position t, its 8 columns = leaves [first, first + 8)
    const uint32_t pos = tid;
    const size_t first = out_index(pos, 0);
    auto leaf = [&](size_t idx) -> Digest {
        const uint32_t slot = pos * 8 + (uint32_t)(idx - first);
        const uint4 a = sm[slot], b = sm[PLANE + slot];
        Digest d;
        d.w[0] = a.x; d.w[1] = a.y; d.w[2] = a.z; d.w[3] = a.w;
        d.w[4] = b.x; d.w[5] = b.y; d.w[6] = b.z; d.w[7] = b.w;
        return tree_hash_leaf(key, d);
    };
    merkle_subtree_fn<3>(key, nodes, n << p.log_l, first, leaf);
}

}  // namespace hodor
