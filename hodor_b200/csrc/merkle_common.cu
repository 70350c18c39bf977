// Field-independent part of the Merkle build: the wide levels (thread-serial subtrees) and the
// authentication-path gather.  Reference: Blake2sIopTree::create / get_path,
// src/iop/blake2s_trivial_iop.rs:131-219, 251-279.
#include <cstdlib>

#include "context.h"

namespace hodor {

static unsigned grid_for(size_t work_items, unsigned block) {
    size_t g = (work_items + block - 1) / block;
    const size_t cap = 148 * 16;
    if (g > cap) g = cap;
    return (unsigned)(g ? g : 1);
}

template <int K, bool LEAF>
static int launch_levels(Ctx& c, const uint4* in, uint4* nodes, size_t w_in, cudaStream_t st, LeafMap lm = LeafMap{0, 0}) {
    // small levels: narrow blocks, so that a few thousand subtrees still spread over many SMs
    const size_t groups = w_in >> K;
    unsigned block = groups >= 148 * 256 ? 256 : (groups >= 148 * 64 ? 128 : 32);
    unsigned grid = grid_for(groups, block);
    if (c.merkle_backfill && groups >= 148 * 256) {
        // Running beside a transform (Ctx::merkle_backfill): small, short-lived blocks -- 128 threads x 80 registers
        // fit into what three resident pass-kernel blocks leave free on an SM, need no shared memory, and return
        // their slot after one subtree per thread, so the block scheduler can keep the transform's blocks (higher
        // stream priority) resident and fill the gaps with hashing.
        // HODOR_BACKFILL_PERSIST = R > 0: a persistent grid of R blocks per SM instead (grid-stride over the subtrees);
        // with the hashing stream at the HIGHER priority (HODOR_COMMIT_PRIORITY=high) those blocks take their place as
        // transform blocks retire and keep it, so the hashing gets a fixed share of every SM.
        static const int persist = getenv("HODOR_BACKFILL_PERSIST") ? atoi(getenv("HODOR_BACKFILL_PERSIST")) : 0;
        static const int bblock = getenv("HODOR_BACKFILL_BLOCK") ? atoi(getenv("HODOR_BACKFILL_BLOCK")) : 128;
        block = (bblock == 64 || bblock == 128 || bblock == 256) ? (unsigned)bblock : 128u;
        grid = persist > 0 ? 148u * (unsigned)persist : (unsigned)((groups + block - 1) / block);
        auto kern = merkle_levels_kernel<K, LEAF>;
        if (!c.configured_kernels.count((const void*)kern)) {  // same L1 / shared split as the pass kernels it shares SMs with
            HODOR_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
            c.configured_kernels.insert((const void*)kern);
        }
    }
    {
        ProfScope ps(c, st, LEAF ? "merkle_levels_leaf" : "merkle_levels_node");
        merkle_levels_kernel<K, LEAF><<<grid, block, 0, st>>>(in, nodes, w_in, c.key, lm);
    }
    HODOR_CUDA_TRY(cudaGetLastError());
    return HODOR_OK;
}

// Width at which the single-block tail kernel takes over.  The tail is one SM: below this width its
// cost is the latency of the remaining chain of levels, above it the compression throughput of a
// single SM (0.19 G/s: 4095 compressions = 22 us) -- so the multi-block kernels go further down than
// they need to for parallelism's sake.  HODOR_MERKLE_TAIL_MAX overrides (power of two, 2..4096).
static size_t tail_max() {
    static size_t v = 0;
    if (v == 0) {
        v = 1024;  // measured on the 2^24 FRI chain: 4096 -> 4.45 ms, 1024 -> 4.30, 512 -> 4.30, 256 -> 4.33, 128 -> 4.38
        if (const char* e = getenv("HODOR_MERKLE_TAIL_MAX")) {
            const size_t x = (size_t)strtoull(e, nullptr, 10);
            if (x >= 2 && x <= 4096 && (x & (x - 1)) == 0) v = x;
        }
    }
    return v;
}

// Hashes the wide part of the tree.  On return *remaining_width is the width (<= tail_max()) of the
// lowest level that has been written; 0 means nothing was done (n <= tail_max(), or too few leaves
// for a subtree kernel: the tail kernel takes the leaves directly).
int merkle_levels(Ctx& c, const uint4* leaves, size_t n, uint4* nodes, size_t* remaining_width, cudaStream_t st,
                  uint32_t leaf_log_g, size_t leaf_chunk) {
    *remaining_width = 0;
    const size_t tmax = tail_max();
    if (n <= tmax || n < 16) {
        if (leaf_log_g) return fail(HODOR_ERR_INVALID_ARG, "internal: interleaved leaves need a tree wider than the tail kernel");
        return HODOR_OK;
    }
    int rc = launch_levels<3, true>(c, leaves, nodes, n, st, LeafMap{leaf_log_g, leaf_chunk});  // levels n/2, n/4, n/8
    if (rc) return rc;
    return merkle_upper_levels(c, nodes, n >> 3, remaining_width, st);
}

// Continues a tree whose level of w nodes (heap [w, 2w)) is already written, down to the width the single-block
// tail kernel takes over at.  Also the second half of the fused fold + leaf kernel's tree (fri.cuh).
int merkle_upper_levels(Ctx& c, uint4* nodes, size_t w, size_t* remaining_width, cudaStream_t st) {
    const size_t tmax = tail_max();
    int rc = HODOR_OK;
    while (w > tmax) {
        int k = 0;
        while (k < 3 && (w >> (k + 1)) >= tmax) k++;  // land exactly on tmax or above
        if (k == 0) k = 1;
        const uint4* in = nodes + 2 * w;
        switch (k) {
            case 1: rc = launch_levels<1, false>(c, in, nodes, w, st); break;
            case 2: rc = launch_levels<2, false>(c, in, nodes, w, st); break;
            default: rc = launch_levels<3, false>(c, in, nodes, w, st); break;
        }
        if (rc) return rc;
        w >>= k;
    }
    *remaining_width = w;
    return HODOR_OK;
}

size_t merkle_tail_width() { return tail_max(); }

int merkle_block_roots(Ctx& c, const uint4* leaves, size_t blocks, int k, uint4* out, cudaStream_t st) {
    if (blocks == 0) return HODOR_OK;
    const unsigned block = blocks >= 148 * 256 ? 256 : (blocks >= 148 * 64 ? 128 : 32);
    const unsigned grid = grid_for(blocks, block);
    {
        ProfScope ps(c, st, "merkle_block_roots");
        switch (k) {
            case 1: merkle_block_roots_kernel<1><<<grid, block, 0, st>>>(leaves, out, blocks, c.key); break;
            case 2: merkle_block_roots_kernel<2><<<grid, block, 0, st>>>(leaves, out, blocks, c.key); break;
            case 3: merkle_block_roots_kernel<3><<<grid, block, 0, st>>>(leaves, out, blocks, c.key); break;
            default: return fail(HODOR_ERR_INVALID_ARG, "internal: leaf block size must be 2, 4 or 8");
        }
    }
    HODOR_CUDA_TRY(cudaGetLastError());
    return HODOR_OK;
}

int merkle_from_level(Ctx& c, const uint4* level, size_t w, uint4* nodes, size_t* remaining_width, cudaStream_t st,
                      uint32_t log_g, size_t chunk) {
    const size_t tmax = tail_max();
    if (w <= tmax || w < 4) return fail(HODOR_ERR_INVALID_ARG, "internal: level too narrow for the level kernels");
    int k = 0;
    while (k < 3 && (w >> (k + 1)) >= tmax) k++;
    if (k == 0) k = 1;
    const LeafMap lm{log_g, chunk};
    int rc;
    switch (k) {
        case 1: rc = launch_levels<1, false>(c, level, nodes, w, st, lm); break;
        case 2: rc = launch_levels<2, false>(c, level, nodes, w, st, lm); break;
        default: rc = launch_levels<3, false>(c, level, nodes, w, st, lm); break;
    }
    if (rc) return rc;
    return merkle_upper_levels(c, nodes, w >> k, remaining_width, st);
}

// out[0] = hash_leaf(values[index ^ 1]); out[1 + j] = sibling on the way up
__global__ void merkle_path_kernel(const uint4* nodes, const uint4* values, size_t size, size_t index, uint4* out,
                                   const __grid_constant__ B2sState key) {
    const uint32_t j = threadIdx.x;
    uint32_t levels = 0;
    while (((size_t)1 << (levels + 1)) < size) levels++;  // log2(size) - 1 stored sibling levels
    if (j == 0) {
        const Digest leaf = ld_digest(values, index ^ 1);
        st_digest(out, 0, hash_leaf32(key, leaf.w));
    }
    if (j < levels) {
        const size_t heap = ((size + index) >> (1 + j)) ^ 1;
        st_digest(out, 1 + j, ld_digest(nodes, heap));
    }
}

int merkle_path_gather(Ctx& c, const uint4* nodes, const uint4* values, size_t size, size_t index, uint4* out,
                       cudaStream_t st) {
    {
        ProfScope ps(c, st, "merkle_path");
        merkle_path_kernel<<<1, 64, 0, st>>>(nodes, values, size, index, out, c.key);
    }
    HODOR_CUDA_TRY(cudaGetLastError());
    return HODOR_OK;
}

// One block per query: slot q of `out` (66 digests wide) receives the path (leaf-pair hash first, then the
// siblings bottom-up; src/iop/blake2s_trivial_iop.rs:251-279) in digests [0, 64) and the queried value in
// digest 64.
__global__ void merkle_paths_kernel(const uint4* nodes, const uint4* values, size_t size, const uint64_t* indices,
                                    uint4* out, const __grid_constant__ B2sState key) {
    const uint32_t j = threadIdx.x;
    const size_t index = (size_t)indices[blockIdx.x];
    uint4* slot = out + 2 * 66 * (size_t)blockIdx.x;
    uint32_t levels = 0;
    while (((size_t)1 << (levels + 1)) < size) levels++;
    if (j == 0) {
        const Digest leaf = ld_digest(values, index ^ 1);
        st_digest(slot, 0, hash_leaf32(key, leaf.w));
        st_digest(slot, 64, ld_digest(values, index));
    }
    if (j < levels) {
        const size_t heap = ((size + index) >> (1 + j)) ^ 1;
        st_digest(slot, 1 + j, ld_digest(nodes, heap));
    }
}

int merkle_paths_gather(Ctx& c, const uint4* nodes, const uint4* values, size_t size, const uint64_t* d_indices,
                        uint32_t count, uint4* out, cudaStream_t st) {
    if (count == 0) return HODOR_OK;
    {
        ProfScope ps(c, st, "merkle_path");
        merkle_paths_kernel<<<count, 64, 0, st>>>(nodes, values, size, d_indices, out, c.key);
    }
    HODOR_CUDA_TRY(cudaGetLastError());
    return HODOR_OK;
}

}  // namespace hodor
