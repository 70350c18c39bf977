// Keyed, personalised Blake2s-256 Merkle tree (the IOP oracle) for sm_100a.
//
// Replaces Blake2sIopTree::create, src/iop/blake2s_trivial_iop.rs:131-219, with the hashing of
// :8-16 (parameters), :36-42 (leaf = raw little-endian Montgomery limbs, 32 B), :81-104
// (hash_leaf / hash_node) and the root -> challenge map of :48-60.
//
// On the CPU every hash is two compressions (key block, then data block).  The key-block
// compression does not depend on the data, so its output chaining value is computed once on the
// host and every device hash is a single compression that starts from it.
//
// Node layout is the reference's: one array of n 32-byte digests in heap order, nodes[0] unused
// (zero), nodes[1] the root, the level with w nodes at [w, 2w), the bottom node level at [n/2, n)
// built from pairs of leaf hashes (which are not stored).
#pragma once
#include <cuda_runtime.h>
#include "field.cuh"
#include "ntt.cuh"

namespace hodor {

struct B2sState {
    uint32_t h[8];
};

HD constexpr uint32_t b2s_iv(int i) {
    constexpr uint32_t iv[8] = {0x6A09E667u, 0xBB67AE85u, 0x3C6EF372u, 0xA54FF53Au,
                                0x510E527Fu, 0x9B05688Cu, 0x1F83D9ABu, 0x5BE0CD19u};
    return iv[i];
}
HD constexpr int b2s_sigma(int r, int i) {
    constexpr unsigned char s[10][16] = {
        {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15}, {14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3},
        {11, 8, 12, 0, 5, 2, 15, 13, 10, 14, 3, 6, 7, 1, 9, 4}, {7, 9, 3, 1, 13, 12, 11, 14, 2, 6, 5, 10, 4, 0, 15, 8},
        {9, 0, 5, 7, 2, 4, 10, 15, 14, 1, 11, 12, 6, 8, 3, 13}, {2, 12, 6, 10, 0, 11, 8, 3, 4, 13, 7, 5, 15, 14, 1, 9},
        {12, 5, 1, 15, 14, 13, 4, 10, 0, 7, 6, 3, 9, 2, 8, 11}, {13, 11, 7, 14, 12, 1, 3, 9, 5, 0, 15, 4, 8, 6, 2, 10},
        {6, 15, 14, 9, 11, 3, 0, 8, 12, 2, 13, 7, 1, 4, 10, 5}, {10, 2, 8, 4, 7, 6, 1, 5, 15, 11, 9, 14, 3, 12, 13, 0}};
    return s[r][i];
}

// Rotations.  HODOR_B2S_ROT selects how the device does them (all bit-identical):
//   0  funnel shift (SHF.R.W) for all four amounts
//   1  byte permute (PRMT) for 16 and 8, funnel shift for 12 and 7
//   2  PRMT for 16 and 8; 12 and 7 as x * 2^(32-r) on the multiplier pipe: the two halves of the 64-bit
//      product are x >> r and x << (32-r) and never overlap, so their sum is the rotation
//   3  the multiplier form for all four
#ifndef HODOR_B2S_ROT
#define HODOR_B2S_ROT 0
#endif
#if defined(__CUDACC__) && HODOR_B2S_ROT >= 2
// multipliers 2^(32-r) for r = 16, 12, 8, 7 in constant memory, so that ptxas cannot turn the product back into a shift
__constant__ uint32_t k_b2s_rot_mul[4] = {1u << 16, 1u << 20, 1u << 24, 1u << 25};
#endif
template <int R>
HD uint32_t rotr32c(uint32_t x) {
#ifdef __CUDA_ARCH__
#if HODOR_B2S_ROT == 1 || HODOR_B2S_ROT == 2
    if constexpr (R == 16) return __byte_perm(x, x, 0x1032);
    if constexpr (R == 8) return __byte_perm(x, x, 0x0321);
#endif
#if HODOR_B2S_ROT >= 2
    {
        const unsigned long long w = (unsigned long long)x * k_b2s_rot_mul[R == 16 ? 0 : (R == 12 ? 1 : (R == 8 ? 2 : 3))];
        return (uint32_t)w + (uint32_t)(w >> 32);
    }
#else
    return __funnelshift_r(x, x, R);
#endif
#else
    return (x >> R) | (x << (32 - R));
#endif
}

// Three-input additions a + b + m.  The ALU pipe (LOP3 / SHF / IADD3, 64 lanes per clock per SM) is
// what bounds the compression; ptxas already moves the two-input additions to the multiplier pipe
// as IMAD.IADD.  HODOR_B2S_ADD3 = 1 does the same for the three-input ones, as two multiply-adds by an
// opaque 1 from constant memory.
// Measured on B200 (tools/microbench.cu): node hash 22.8 -> 27.6, leaf hash 25.0 -> 27.2 G compressions/s.
#ifndef HODOR_B2S_ADD3
#define HODOR_B2S_ADD3 1
#endif
#if defined(__CUDACC__) && HODOR_B2S_ADD3
__constant__ uint32_t k_b2s_one = 1u;
#endif
// m_is_zero: known at compile time after unrolling (the zero half of a 32-byte leaf message)
HD uint32_t b2s_add3(uint32_t a, uint32_t b, uint32_t m, bool m_is_zero) {
    if (m_is_zero) return a + b;
#if defined(__CUDA_ARCH__) && HODOR_B2S_ADD3
    return (a * k_b2s_one + b) * k_b2s_one + m;
#else
    return a + b + m;
#endif
}

#define HODOR_B2S_G(a, b, c, d, x, y, zx, zy) \
    v[a] = b2s_add3(v[a], v[b], (x), (zx)); \
    v[d] = rotr32c<16>(v[d] ^ v[a]);  \
    v[c] = v[c] + v[d];               \
    v[b] = rotr32c<12>(v[b] ^ v[c]);  \
    v[a] = b2s_add3(v[a], v[b], (y), (zy)); \
    v[d] = rotr32c<8>(v[d] ^ v[a]);   \
    v[c] = v[c] + v[d];               \
    v[b] = rotr32c<7>(v[b] ^ v[c]);

// One Blake2s compression (RFC 7693 3.2).  m[] indices are compile-time after unrolling, so the
// message stays in registers; with HALF only m[0..7] are non-zero (32-byte leaf message).
template <bool HALF>
HD void b2s_compress(uint32_t (&h)[8], const uint32_t (&m)[16], uint32_t t0, bool last) {
    uint32_t v[16];
#pragma unroll
    for (int i = 0; i < 8; i++) {
        v[i] = h[i];
        v[i + 8] = b2s_iv(i);
    }
    v[12] ^= t0;
    if (last) v[14] = ~v[14];
#pragma unroll
    for (int r = 0; r < 10; r++) {
#define MSGZ(i) (HALF && b2s_sigma(r, i) >= 8)
#define MSG(i) (MSGZ(i) ? 0u : m[b2s_sigma(r, i)])
        HODOR_B2S_G(0, 4, 8, 12, MSG(0), MSG(1), MSGZ(0), MSGZ(1))
        HODOR_B2S_G(1, 5, 9, 13, MSG(2), MSG(3), MSGZ(2), MSGZ(3))
        HODOR_B2S_G(2, 6, 10, 14, MSG(4), MSG(5), MSGZ(4), MSGZ(5))
        HODOR_B2S_G(3, 7, 11, 15, MSG(6), MSG(7), MSGZ(6), MSGZ(7))
        HODOR_B2S_G(0, 5, 10, 15, MSG(8), MSG(9), MSGZ(8), MSGZ(9))
        HODOR_B2S_G(1, 6, 11, 12, MSG(10), MSG(11), MSGZ(10), MSGZ(11))
        HODOR_B2S_G(2, 7, 8, 13, MSG(12), MSG(13), MSGZ(12), MSGZ(13))
        HODOR_B2S_G(3, 4, 9, 14, MSG(14), MSG(15), MSGZ(14), MSGZ(15))
#undef MSG
#undef MSGZ
    }
#pragma unroll
    for (int i = 0; i < 8; i++) h[i] ^= v[i] ^ v[i + 8];
}

// Chaining value after the key block of BASE_BLAKE2S_PARAMS (src/iop/blake2s_trivial_iop.rs:8-16):
// digest 32, key "Squeamish Ossifrage" (19 B), fanout 1, depth 1, personal "Shaftoe".
inline B2sState b2s_keyed_state() {
    B2sState s;
    const char key[] = "Squeamish Ossifrage";
    const char personal[8] = {'S', 'h', 'a', 'f', 't', 'o', 'e', 0};
    for (int i = 0; i < 8; i++) s.h[i] = b2s_iv(i);
    s.h[0] ^= 0x01010000u ^ (19u << 8) ^ 32u;
    uint32_t pw[2] = {0, 0};
    for (int i = 0; i < 8; i++) pw[i / 4] |= (uint32_t)(unsigned char)personal[i] << (8 * (i % 4));
    s.h[6] ^= pw[0];
    s.h[7] ^= pw[1];
    uint32_t m[16] = {0};
    for (int i = 0; i < 19; i++) m[i / 4] |= (uint32_t)(unsigned char)key[i] << (8 * (i % 4));
    b2s_compress<false>(s.h, m, 64u, false);
    return s;
}

struct Digest {
    uint32_t w[8];
};

// hash_leaf: message = 32 raw bytes, total input 64 (key block) + 32
HD Digest hash_leaf32(const B2sState& key, const uint32_t (&leaf)[8]) {
    uint32_t h[8], m[16];
#pragma unroll
    for (int i = 0; i < 8; i++) {
        h[i] = key.h[i];
        m[i] = leaf[i];
        m[i + 8] = 0;
    }
    b2s_compress<true>(h, m, 96u, true);
    Digest d;
#pragma unroll
    for (int i = 0; i < 8; i++) d.w[i] = h[i];
    return d;
}
// hash_node: message = left || right, total input 64 + 64
HD Digest hash_node64(const B2sState& key, const Digest& l, const Digest& r) {
    uint32_t h[8], m[16];
#pragma unroll
    for (int i = 0; i < 8; i++) {
        h[i] = key.h[i];
        m[i] = l.w[i];
        m[i + 8] = r.w[i];
    }
    b2s_compress<false>(h, m, 128u, true);
    Digest d;
#pragma unroll
    for (int i = 0; i < 8; i++) d.w[i] = h[i];
    return d;
}

// Out-of-line copies for the tree kernels.  A thread-serial subtree of 2^3 leaves is 15 compressions;
// inlined that is ~240 KB of straight-line code per kernel and ncu shows 6 `no_instruction` stall
// cycles per issued instruction (profiles/r01_final_ncu_merkle).  Two shared ~16 KB bodies stay in the
// instruction cache; the call passes 16-24 registers, against ~1000 instructions of work.
#ifndef HODOR_B2S_OUT_OF_LINE
#define HODOR_B2S_OUT_OF_LINE 1
#endif
#if defined(__CUDACC__)
static __device__ __noinline__ Digest hash_leaf32_ool(B2sState key, Digest leaf) { return hash_leaf32(key, leaf.w); }
static __device__ __noinline__ Digest hash_node64_ool(B2sState key, Digest l, Digest r) { return hash_node64(key, l, r); }
DEV Digest tree_hash_leaf(const B2sState& key, const Digest& leaf) {
#if HODOR_B2S_OUT_OF_LINE
    return hash_leaf32_ool(key, leaf);
#else
    return hash_leaf32(key, leaf.w);
#endif
}
DEV Digest tree_hash_node(const B2sState& key, const Digest& l, const Digest& r) {
#if HODOR_B2S_OUT_OF_LINE
    return hash_node64_ool(key, l, r);
#else
    return hash_node64(key, l, r);
#endif
}
#endif

DEV Digest ld_digest(const uint4* base, size_t idx) {
    const uint4 a = base[2 * idx], b = base[2 * idx + 1];
    Digest d;
    d.w[0] = a.x; d.w[1] = a.y; d.w[2] = a.z; d.w[3] = a.w;
    d.w[4] = b.x; d.w[5] = b.y; d.w[6] = b.z; d.w[7] = b.w;
    return d;
}
DEV void st_digest(uint4* base, size_t idx, const Digest& d) {
    base[2 * idx] = make_uint4(d.w[0], d.w[1], d.w[2], d.w[3]);
    base[2 * idx + 1] = make_uint4(d.w[4], d.w[5], d.w[6], d.w[7]);
}
// the same for arrays known to live in HBM: one 256-bit access per digest / leaf (ntt.cuh, HODOR_LDST256)
DEV Digest ldg_digest(const uint4* base, size_t idx) {
#if HODOR_LDST256
    Digest d;
    ld256(base + 2 * idx, d.w);
    return d;
#else
    return ld_digest(base, idx);
#endif
}
DEV void stg_digest(uint4* base, size_t idx, const Digest& d) {
#if HODOR_LDST256
    st256(base + 2 * idx, d.w);
#else
    st_digest(base, idx, d);
#endif
}

// Where leaf b of a tree lives.  Default: in[b].  Interleaved (a layer that arrived as G chunks of `chunk`
// elements, chunk r holding the cyclic slice v[r + G*t] of the natural-order vector -- the receive side of
// the all-to-all of the sharded FRI chain): in[(b mod G) * chunk + b / G], so the re-blocking costs index
// arithmetic in this kernel's loads instead of a transposing copy through HBM.
struct LeafMap {
    uint32_t log_g;  // 0: identity
    size_t chunk;
};
DEV size_t leaf_index(const LeafMap& m, size_t b) {
    return m.log_g == 0 ? b : (b & (((size_t)1 << m.log_g) - 1)) * m.chunk + (b >> m.log_g);
}

// Thread-serial subtree over 2^K consecutive inputs starting at input index `first`.
// Inputs are leaves (LEAF) or the digests of the level with `w_in` nodes.  The node covering
// inputs [a*2^j, (a+1)*2^j) lives at heap index (w_in >> j) + a.
template <int K, bool LEAF>
DEV Digest merkle_subtree(const B2sState& key, const uint4* in, uint4* nodes, size_t w_in, size_t first, const LeafMap& lm) {
    if constexpr (K == 0) {
        if constexpr (LEAF) return tree_hash_leaf(key, ldg_digest(in, leaf_index(lm, first)));
        else return ldg_digest(in, leaf_index(lm, first));  // a node level may arrive as cyclic chunks as well
    } else {
        const Digest l = merkle_subtree<K - 1, LEAF>(key, in, nodes, w_in, first, lm);
        const Digest r = merkle_subtree<K - 1, LEAF>(key, in, nodes, w_in, first + ((size_t)1 << (K - 1)), lm);
        const Digest d = tree_hash_node(key, l, r);
        stg_digest(nodes, (w_in >> K) + (first >> K), d);
        return d;
    }
}

// The same subtree with the leaf hashes coming from a functor (leaf index -> hash of that leaf): lets a producer
// kernel hash values it has just computed, while they are still in registers (fri.cuh: fold + commit).
template <int K, class LeafHash>
DEV Digest merkle_subtree_fn(const B2sState& key, uint4* nodes, size_t w_in, size_t first, LeafHash&& leaf_hash) {
    if constexpr (K == 0) {
        return leaf_hash(first);
    } else {
        const Digest l = merkle_subtree_fn<K - 1>(key, nodes, w_in, first, leaf_hash);
        const Digest r = merkle_subtree_fn<K - 1>(key, nodes, w_in, first + ((size_t)1 << (K - 1)), leaf_hash);
        const Digest d = tree_hash_node(key, l, r);
        stg_digest(nodes, (w_in >> K) + (first >> K), d);
        return d;
    }
}

// Root of the complete subtree over the 2^K adjacent leaves starting at `first`; nothing below it is stored.
template <int K>
DEV Digest merkle_block_root(const B2sState& key, const uint4* in, size_t first) {
    if constexpr (K == 0) {
        return tree_hash_leaf(key, ldg_digest(in, first));
    } else {
        const Digest l = merkle_block_root<K - 1>(key, in, first);
        const Digest r = merkle_block_root<K - 1>(key, in, first + ((size_t)1 << (K - 1)));
        return tree_hash_node(key, l, r);
    }
}
// out[g] = root of leaves [g * 2^K, (g + 1) * 2^K): the level a rank of the sharded chain can hash without its
// neighbours' leaves (it holds blocks of 2^K adjacent ones) and then exchanges instead of the values
template <int K>
__global__ void __launch_bounds__(256) merkle_block_roots_kernel(const uint4* in, uint4* out, size_t blocks,
                                                                 const __grid_constant__ B2sState key) {
    for (size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x; g < blocks; g += (size_t)gridDim.x * blockDim.x)
        stg_digest(out, g, merkle_block_root<K>(key, in, g << K));
}

// grid-stride over groups of 2^K inputs; writes K node levels
template <int K, bool LEAF>
__global__ void __launch_bounds__(256) merkle_levels_kernel(const uint4* in, uint4* nodes, size_t w_in,
                                                            const __grid_constant__ B2sState key, const LeafMap lm) {
    const size_t groups = w_in >> K;
    for (size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x; g < groups; g += (size_t)gridDim.x * blockDim.x)
        merkle_subtree<K, LEAF>(key, in, nodes, w_in, g << K, lm);
}

// Finishes a tree inside one block: from `w_in` inputs (leaves if LEAF, else the stored level of
// w_in nodes) up to the root, then root -> challenge (interpret_hash, :48-60) in Montgomery form.
template <class F, bool LEAF>
__global__ void __launch_bounds__(1024) merkle_tail_kernel(const uint4* in, uint4* nodes, uint32_t w_in,
                                                           const __grid_constant__ B2sState key, uint4* root_out,
                                                           uint4* challenge_out, uint32_t zero) {
    const uint32_t tid = threadIdx.x, nt = blockDim.x;
    if (tid == 0) {  // nodes[0] is never written by the reference: stays [0u8; 32]
        Digest z;
#pragma unroll
        for (int i = 0; i < 8; i++) z.w[i] = 0;
        st_digest(nodes, 0, z);
    }
    // The top of the tree (heap indices < 1024) is mirrored in shared memory: the chain of levels is
    // latency bound (one compression, ~0.6 us, per level), and a global store -> barrier -> global load
    // round trip per level would double that.  `nodes` still receives every digest.
    __shared__ uint4 heap[2 * 1024];
    bool children_in_smem = false;  // the level below the one being computed is complete in `heap`
    uint32_t width = w_in / 2;
    if constexpr (LEAF) {
        children_in_smem = 2 * width <= 1024;
        for (uint32_t i = tid; i < width; i += nt) {
            const Digest a = ld_digest(in, 2 * i), b = ld_digest(in, 2 * i + 1);
            const Digest d = hash_node64(key, hash_leaf32(key, a.w), hash_leaf32(key, b.w));
            st_digest(nodes, width + i, d);
            if (children_in_smem) st_digest(heap, width + i, d);
        }
        __syncthreads();
        width /= 2;
    }
    for (; width >= 1; width /= 2) {
        const bool to_smem = 2 * width <= 1024;
        for (uint32_t i = tid; i < width; i += nt) {
            const uint32_t l = 2 * (width + i);
            const Digest a = children_in_smem ? ld_digest(heap, l) : ld_digest(nodes, l);
            const Digest b = children_in_smem ? ld_digest(heap, l + 1) : ld_digest(nodes, (size_t)l + 1);
            const Digest d = hash_node64(key, a, b);
            st_digest(nodes, width + i, d);
            if (to_smem) st_digest(heap, width + i, d);
        }
        __syncthreads();
        children_in_smem = to_smem;
    }
    if (tid == 0) {
        const Digest root = children_in_smem ? ld_digest(heap, 1) : ld_digest(nodes, 1);  // w_in == 1: nothing hashed here
        if (root_out != nullptr) st_digest(root_out, 0, root);
        if (challenge_out != nullptr) {
            // read_be: digest bytes are one big-endian 256-bit integer; word j (LE load) holds
            // bytes 4j..4j+3, so limb (7 - j) = byteswap(word j)
            Fe v;
#pragma unroll
            for (int j = 0; j < 8; j++) v.v[7 - j] = __byte_perm(root.w[j], 0, 0x0123);
            constexpr uint32_t shave = (256 - (F::NUM_BITS - 1)) % 64;  // 256 - CAPACITY
            if constexpr (shave >= 32) {
                v.v[7] = 0;
                v.v[6] &= 0xffffffffu >> (shave - 32);
            } else {
                v.v[7] &= 0xffffffffu >> shave;
            }
            const Field<F> fld(tid & zero);
            st_fe(challenge_out, 0, fld.to_mont(v));  // from_repr
        }
    }
}

}  // namespace hodor
