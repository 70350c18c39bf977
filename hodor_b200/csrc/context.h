// Process-wide context of the hodor_b200 library: one GPU, one default stream, a grow-only
// workspace, a cache of twiddle / coset-power tables.  Host-only header.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <cstdlib>
#include <map>
#include <initializer_list>
#include <mutex>
#include <set>
#include <string>
#include <vector>

#include "../../include/hodor_b200.h"
#include "field.cuh"
#include "merkle.cuh"
#include "ntt.cuh"

namespace hodor {

void set_error(const std::string& msg);
int fail(int code, const std::string& msg);
int cuda_fail(cudaError_t e, const char* what);

#define HODOR_CUDA_TRY(expr)                                  \
    do {                                                      \
        cudaError_t _e = (expr);                              \
        if (_e != cudaSuccess) return cuda_fail(_e, #expr);   \
    } while (0)

// base^e for one or several bases, e in [0, 2^bits): lo[b][e & mask] * hi[b][e >> lo_bits]
struct PowTables {
    uint4* block = nullptr;  // one allocation: bases | lo tables | hi tables
    uint4* lo = nullptr;
    uint4* hi = nullptr;
    uint32_t lo_bits = 0, hi_bits = 0, count = 0;
    size_t bytes = 0;
    TwoLevel two_level() const { return TwoLevel{lo, hi, lo_bits}; }
    uint32_t stride_lo() const { return 1u << lo_bits; }
    uint32_t stride_hi() const { return 1u << hi_bits; }
};

constexpr uint32_t FLAT_MAX_LOG = 21;  // largest sub-transform that gets a flat inter-pass twiddle table (128 MiB)

struct NttTables {
    PowTables pw;            // powers of omega, exponents [0, 2^log_n)
    // tw_b and tw_direct entries are FePre (64 B): read directly as multipliers
    uint4* tw_b[10] = {};    // tw_b[B][x] = omega^(x << (log_n - B)), B in 6..9 (as used by the plan)
    uint4* tw_b_block = nullptr;
    uint4* tw_direct[FLAT_MAX_LOG + 1] = {};  // tw_direct[k][x] = omega^(x << (log_n - k)), x < 2^k: flat inter-pass twiddles
    uint4* tw_direct_block = nullptr;
    FePre wr[7];             // omega_16^k, k = 1..7, fixed-operand form
    size_t bytes = 0;
};

struct Ctx {
    int device = -1;
    cudaStream_t stream = nullptr;
    cudaStream_t copy_in = nullptr, copy_out = nullptr;  // H2D / D2H streams of the pipelined batch entry points
    // Lift-and-commit of several polynomials: the tree of polynomial i is hashed on this LOW-priority stream while the
    // transform of polynomial i+1 runs on `stream` (high priority).  The transform is bound by the multiplier pipe
    // (76 % busy, ALU 41 %), the hashing by the ALU pipe (88 %, multiplier 60 %): sharing SMs lets each use the issue
    // slots and the pipe the other leaves idle.  HODOR_CONCURRENT_COMMIT=0 puts both on one stream.
    cudaStream_t commit_stream = nullptr;
    bool concurrent_commit = true;
    bool merkle_backfill = false;  // set around the tree builds that run beside a transform
    std::mutex mu;
    void* ws = nullptr;
    size_t ws_bytes = 0;
    uint4* small = nullptr;  // 4 KiB of device scratch for scalars
    // Pinned host scratch for small results (roots) of pipelined entry points: a device-to-host copy into the
    // caller's pageable memory blocks the calling thread until the stream has drained, which would serialise
    // the pipeline; results land here asynchronously and are handed over after the final synchronise.
    uint8_t* pinned_small = nullptr;
    static constexpr size_t PINNED_SMALL_BYTES = 64 << 10;
    B2sState key;
    std::map<std::string, NttTables> ntt_tables;
    std::map<std::string, PowTables> pow_tables;
    size_t table_bytes = 0;
    std::atomic<uint64_t> launches{0};

    // host-pointer entry points stage through these grow-only device buffers
    void* io[2] = {nullptr, nullptr};
    size_t io_bytes[2] = {0, 0};

    // optional per-launch timing (bench.py's live roofline): events around every kernel
    struct ProfRec {
        const char* name;
        cudaEvent_t start, stop;
    };
    bool profiling = false;
    std::vector<ProfRec> prof;
    std::vector<cudaEvent_t> event_pool;

    // expanded (one entry per element) twiddle / coset-power tables; bounded by full_budget bytes
    std::map<std::string, std::pair<uint4*, size_t>> full_tables;
    size_t full_bytes = 0;
    size_t full_budget = (size_t)24 << 30;  // HODOR_TABLE_BUDGET_MB overrides; 0 disables

    // freed prototype / tree / staging blocks kept for reuse (a committed 2^27 oracle is 8 GiB); HODOR_POOL_CACHE_MB
    // 96 GiB: the eight committed 2^27 oracles of one lift-and-commit batch (8 GiB each) fit, so a prover cycling through
    // them makes no driver allocation calls (with 64 GiB one block was released and re-allocated per batch: 0.7-2 ms per
    // polynomial, erratic).  The cache is dropped and the allocation retried when the driver runs out of memory.
    size_t pool_cache_cap = (size_t)96 << 30;
    std::vector<std::pair<void*, size_t>> pool_free_list;
    std::map<void*, size_t> pool_live;
    void* pool_alloc(size_t bytes);
    void pool_free(void* p);

    int ensure_workspace(size_t bytes);
    int ensure_io(int which, size_t bytes);

    // host -> device copies of caller buffers: pageable sources are staged through pinned buffers by several
    // host threads (HODOR_STAGE_THREADS, default min(8, cores / 2); 1 = leave it to the driver)
    int h2d(void* dptr, const void* hptr, size_t bytes, cudaStream_t st);
    void* stage[2] = {nullptr, nullptr};
    cudaEvent_t stage_free[2] = {nullptr, nullptr};
    bool stage_used[2] = {false, false};
    int stage_threads = 0;

    // The workspace is shared by every multi-kernel entry point (multi-pass NTT / LDE, batch_inversion,
    // evaluate_at) and `_dev` calls may arrive on different caller streams: the host-side mutex orders the
    // enqueues, not the execution.  ws_acquire makes `st` wait for the last user of the workspace when that
    // was another stream; ws_release records the new last use.  Same-stream callers pay nothing.
    cudaEvent_t ws_event = nullptr;
    cudaStream_t ws_stream = nullptr;
    bool ws_used = false;
    int ws_acquire(size_t bytes, cudaStream_t st);
    int ws_release(cudaStream_t st);

    // kernels whose dynamic shared memory attribute has been raised on this context's device
    std::set<const void*> configured_kernels;
    std::map<uint64_t, Fe> inv_cache;  // (field, log_n) -> omega_N^-1 of the FRI domain

    // lift-and-commit: the last pass of the transform also hashes the bottom three tree levels of its output
    // (ntt_commit.cuh).  hodor_cuda_lde_commit_batch sets `nodes` around the transform; Ops::ntt sets `done` when it
    // took the fused kernel (multi-pass plan, last digit <= 8, no output scaling), else the tree is built as usual.
    // HODOR_FUSE_LAST_COMMIT: 0 off; 1 (default) plans whose last digit is 8 -- measured 36.63 against 37.12 ms per
    // lift-and-commit of a 2^24 x 8 polynomial; 2 also 7 and 3 also 6: measured slower than the separate tree, which
    // runs beside the next polynomial's transform (2^22 x 8: 8.97 against 8.73 ms, 2^20 x 8: 2.20 against 2.10 ms)
    int fuse_last_commit = 1;
    struct FuseCommit {
        uint4* nodes = nullptr;
        bool done = false;
    } fuse_commit;

    bool fuse_fold_commit = false;  // FRI chain: fold + bottom of the next tree in one kernel (HODOR_FUSE_FOLD_COMMIT=1; measured slower, profiles/r02_experiments.md)

    // step A of the sharded NTT: the last pass stores into the peers' receive buffers (ntt.cuh NttPass::peer)
    struct PeerStore {
        bool on = false;
        uint32_t chunk_log = 0, rank = 0;
        uint4* base[16] = {};
    } peer_store;

    struct Comm* comm = nullptr;  // multi-GPU state (sharded.cu); null until hodor_cuda_comm_init
    cudaEvent_t take_event();
};

// RAII: brackets one kernel launch with events when profiling is on
struct ProfScope {
    Ctx& c;
    cudaStream_t st;
    long idx = -1;
    ProfScope(Ctx& ctx_, cudaStream_t st_, const char* name) : c(ctx_), st(st_) {
        c.launches++;
        if (!c.profiling) return;
        Ctx::ProfRec r{name, c.take_event(), c.take_event()};
        cudaEventRecord(r.start, st);
        c.prof.push_back(r);
        idx = (long)c.prof.size() - 1;
    }
    ~ProfScope() {
        if (idx >= 0) cudaEventRecord(c.prof[idx].stop, st);
    }
};

Ctx* ctx();  // nullptr (and error set) when not initialised

struct NttPlan {
    int passes = 0;  // 0 => single-block kernel
    int b[4] = {0, 0, 0, 0};
};
// Digits are balanced in 6..max_digit.  max_digit = 9 gives the fewest passes over HBM, max_digit = 8
// avoids the 512-point tiles (128 KiB of shared memory: one resident block per SM, radix-8 groups at
// 124 registers), at the price of a fourth pass for 2^25..2^27.  HODOR_NTT_MAX_DIGIT overrides.
inline int ntt_max_digit() {
    static int v = 0;
    if (v == 0) {
        v = 8;
        if (const char* e = getenv("HODOR_NTT_MAX_DIGIT")) {
            const int x = atoi(e);
            if (x == 8 || x == 9) v = x;
        }
    }
    return v;
}
inline NttPlan make_plan(uint32_t log_n) {
    NttPlan p;
    if (log_n <= 11) return p;
    const int md = ntt_max_digit();
    p.passes = (int)((log_n + md - 1) / md);
    if ((int)log_n / p.passes < 6 && p.passes > 2) p.passes--;  // digits below 6 are not built: fewer, wider passes
    const int base = (int)log_n / p.passes, rem = (int)log_n % p.passes;
    for (int i = 0; i < p.passes; i++) p.b[i] = base + (i < rem ? 1 : 0);
    return p;
}

// Per-field entry points (one translation unit per field instantiates them).
struct FieldOps {
    // host scalar arithmetic on 8 x u32 Montgomery elements
    void (*h_mul)(const Fe&, const Fe&, Fe&);
    void (*h_add)(const Fe&, const Fe&, Fe&);
    void (*h_sub)(const Fe&, const Fe&, Fe&);
    void (*h_pow)(const Fe&, uint64_t, Fe&);
    int (*h_inverse)(const Fe&, Fe&);
    void (*h_from_repr)(const Fe&, Fe&);
    void (*h_into_repr)(const Fe&, Fe&);
    void (*h_constants)(Fe& modulus, Fe& one, Fe& generator, Fe& root, uint32_t& s, uint32_t& num_bits);
    int (*h_domain_generator)(uint32_t log_n, Fe& out);
    int (*h_root_to_challenge)(const uint8_t* root, Fe& out);

    // device work, enqueued on `st`
    // general transform: `log_l` cosets; shift0/step null => no input scaling (then log_l == 0);
    // out_mode 0 none, 1 multiply by n^-1 (ifft), 2 multiply by n^-1 * out_g^k (icoset)
    int (*ntt)(Ctx&, const uint4* in, uint4* out, uint32_t log_n, uint32_t log_l, const Fe& omega, const Fe* shift0,
               const Fe* step, int out_mode, const Fe* out_g, cudaStream_t st);
    int (*scale_pow)(Ctx&, uint4* a, size_t n, const Fe& g, cudaStream_t st);
    int (*elementwise)(Ctx&, int op, const uint4* a, const uint4* b, uint4* out, size_t n, const Fe* scalar, uint64_t exp,
                       cudaStream_t st);
    // out[j] = map(first * ratio^j): mode 0 the point itself, 1 minus *cst, 2 inv_van[j mod van_len] * prod (x_j - roots[r])
    // (all elements in Montgomery form; roots / inv_van are host arrays)
    int (*coset_map)(Ctx&, int mode, uint4* out, size_t n, const Fe& ratio, const Fe& first, const Fe* cst, const Fe* h_roots,
                     uint32_t num_roots, const Fe* h_inv_van, uint32_t van_len, cudaStream_t st);
    int (*batch_inversion)(Ctx&, uint4* a, size_t n, int* d_status, cudaStream_t st);
    int (*evaluate_at)(Ctx&, const uint4* a, size_t n, const Fe& g, uint4* d_out, cudaStream_t st);
    int (*selftest_mul_pre)(Ctx&, unsigned long long* d_mismatch, cudaStream_t st);
    int (*merkle_tail)(Ctx&, const uint4* in, uint4* nodes, uint32_t w_in, bool leaf, uint4* root, uint4* chal,
                       cudaStream_t st);
    int (*fri_fold)(Ctx&, const uint4* in, size_t n, uint32_t log_n0, uint32_t layer, const uint4* chal, uint4* out,
                    uint64_t idx_offset, uint64_t idx_stride, uint32_t blk_log, cudaStream_t st);
    int (*shard_rows)(Ctx&, const uint4* in, uint4* out, uint32_t log_n, uint32_t log_g, uint32_t rank, const Fe& omega,
                      cudaStream_t st);
    // one FRI layer fused with the bottom of its tree: folds `in` (n values) into `out` (n/2) and writes the node
    // levels n/4, n/8, n/16 of `nodes` (heap of the n/2-leaf tree) from the values it holds in registers
    int (*fri_fold_commit)(Ctx&, const uint4* in, size_t n, uint32_t log_n0, uint32_t layer, const uint4* chal, uint4* out,
                           uint4* nodes, cudaStream_t st);
};
extern const FieldOps kOpsBlsFr, kOpsBn254Fr, kOpsStark252;
const FieldOps* field_ops(int field_id);

// field independent Merkle levels (merkle_common.cu)
// leaf_log_g / leaf_chunk: leaves stored as 2^leaf_log_g cyclic-slice chunks (merkle.cuh LeafMap); 0 = natural order
int merkle_levels(Ctx&, const uint4* leaves, size_t n, uint4* nodes, size_t* remaining_width, cudaStream_t st,
                  uint32_t leaf_log_g = 0, size_t leaf_chunk = 0);
int merkle_upper_levels(Ctx&, uint4* nodes, size_t w, size_t* remaining_width, cudaStream_t st);
// out[g] = root of the subtree over leaves [g << k, (g + 1) << k), k in 1..3
int merkle_block_roots(Ctx&, const uint4* leaves, size_t blocks, int k, uint4* out, cudaStream_t st);
// the tree above a level of w digests given in `level` (natural order, or 2^log_g cyclic chunks of `chunk` digests):
// writes heap [.., w) down to the tail width; *remaining_width as merkle_levels
int merkle_from_level(Ctx&, const uint4* level, size_t w, uint4* nodes, size_t* remaining_width, cudaStream_t st,
                      uint32_t log_g, size_t chunk);
size_t merkle_tail_width();
int merkle_path_gather(Ctx&, const uint4* nodes, const uint4* values, size_t size, size_t index, uint4* out,
                       cudaStream_t st);
int merkle_paths_gather(Ctx&, const uint4* nodes, const uint4* values, size_t size, const uint64_t* d_indices,
                        uint32_t count, uint4* out, cudaStream_t st);


int do_merkle(Ctx& c, const FieldOps* ops, const uint4* leaves, size_t n, uint4* nodes, uint4* root, uint4* chal,
              cudaStream_t st, uint32_t leaf_log_g = 0, size_t leaf_chunk = 0);

}  // namespace hodor

// ---- FRI commit chain: the device-resident FRIProofPrototype (src/fri/mod.rs:107-117) ----------------
struct hodor_fri_proto {
    int field_id = 0;
    uint64_t n = 0;
    uint32_t lde_factor = 0, out_coeffs = 0;
    int steps = 0;
    uint4* block = nullptr;      // one allocation for everything below
    const uint4* lde = nullptr;  // layer-0 values (borrowed device pointer, or inside `owned_lde`)
    uint4* owned_lde = nullptr;
    std::vector<uint4*> nodes;   // nodes[0] = l0, nodes[i] = intermediate i-1
    std::vector<uint4*> values;  // values[0] = lde, values[i] = intermediate i-1
    uint4* roots = nullptr;      // steps + 1 digests
    uint4* chal = nullptr;       // steps + 1 elements (the last one is never used by a fold)
    uint4* final_coeffs = nullptr;  // ifft of the last layer (n >> steps elements)
    uint4* path = nullptr;       // scratch for queries: 64 digests + 1 element
    uint4* proof_scratch = nullptr;  // produce_proof: 2 query slots (66 digests each) per layer
    uint64_t* proof_idx = nullptr;   // produce_proof: 2 indices per layer
};


namespace hodor {
void fri_destroy(hodor_fri_proto* p);
void comm_destroy(Ctx* c);  // sharded.cu
// the whole chain on stream c->stream, synchronised on return; caller holds Ctx::mu
hodor_fri_proto* fri_commit_impl(Ctx* c, const uint64_t* lde, uint64_t n, uint32_t lde_factor, uint32_t out_coeffs,
                                 int lde_on_device, int field_id);
}  // namespace hodor

#define LOCKED_CTX()                          \
    Ctx* c = ctx();                           \
    if (!c) return HODOR_ERR_CUDA;            \
    std::lock_guard<std::mutex> _lk(c->mu)
// element / digest arrays are read and written with 256-bit accesses (ntt.cuh ld256 / st256)
inline bool aligned32(std::initializer_list<const void*> ps) {
    for (const void* p : ps)
        if (reinterpret_cast<uintptr_t>(p) & 31u) return false;
    return true;
}
#define CHECK_DEV_PTRS(...) \
    if (!aligned32({__VA_ARGS__})) return fail(HODOR_ERR_INVALID_ARG, "device element arrays must be 32-byte aligned")
#define GET_OPS(field_id)                         \
    const FieldOps* ops = field_ops(field_id);    \
    if (!ops) return HODOR_ERR_INVALID_ARG


