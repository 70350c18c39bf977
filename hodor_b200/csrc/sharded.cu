// Multi-GPU entry points of the C ABI (include/hodor_b200.h, "several GPUs"): one process per GPU, NCCL
// for the exchange steps, everything else the single-GPU kernels.  NCCL is resolved at run time
// (dlopen of libnccl.so.2), so the library loads -- and every single-GPU entry point works -- on a
// machine without NCCL, and inside a process that already carries one (torch bundles its own
// libnccl.so.2: the loader hands back that same copy).
//
//   hodor_cuda_ntt_sharded       four-step (Bailey) NTT; the reference's parallel_fft is this decomposition
//                                over CPU threads (src/fft/fft.rs:68-125)
//   hodor_cuda_lde_fri_sharded   coset LDE (cosets are independent: src/polynomials/mod.rs:572-587) + the FRI
//                                commit chain (src/fri/fri_on_values.rs:11-159) on cyclic slices
#include <dlfcn.h>
#include <nccl.h>  // types and prototypes only; nothing links against libnccl

#include <memory>

#include "context.h"

namespace hodor {

struct NcclApi {
    void* lib = nullptr;
    decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
    decltype(&ncclCommInitRank) CommInitRank = nullptr;
    decltype(&ncclCommDestroy) CommDestroy = nullptr;
    decltype(&ncclGroupStart) GroupStart = nullptr;
    decltype(&ncclGroupEnd) GroupEnd = nullptr;
    decltype(&ncclSend) Send = nullptr;
    decltype(&ncclRecv) Recv = nullptr;
    decltype(&ncclAllGather) AllGather = nullptr;
    decltype(&ncclGetErrorString) GetErrorString = nullptr;
};

static NcclApi g_nccl;
static std::mutex g_nccl_mu;

static int load_nccl() {
    std::lock_guard<std::mutex> lk(g_nccl_mu);
    if (g_nccl.lib) return HODOR_OK;
    const char* names[] = {getenv("HODOR_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    void* h = nullptr;
    for (const char* n : names) {
        if (n && *n && (h = dlopen(n, RTLD_NOW | RTLD_GLOBAL))) break;
    }
    if (!h) return fail(HODOR_ERR_CUDA, std::string("cannot load libnccl.so.2 (set HODOR_NCCL_LIB): ") + (dlerror() ? dlerror() : "?"));
#define HODOR_NCCL_SYM(field, sym)                                                                   \
    g_nccl.field = (decltype(g_nccl.field))dlsym(h, #sym);                                           \
    if (!g_nccl.field) return fail(HODOR_ERR_CUDA, "libnccl lacks " #sym)
    HODOR_NCCL_SYM(GetUniqueId, ncclGetUniqueId);
    HODOR_NCCL_SYM(CommInitRank, ncclCommInitRank);
    HODOR_NCCL_SYM(CommDestroy, ncclCommDestroy);
    HODOR_NCCL_SYM(GroupStart, ncclGroupStart);
    HODOR_NCCL_SYM(GroupEnd, ncclGroupEnd);
    HODOR_NCCL_SYM(Send, ncclSend);
    HODOR_NCCL_SYM(Recv, ncclRecv);
    HODOR_NCCL_SYM(AllGather, ncclAllGather);
    HODOR_NCCL_SYM(GetErrorString, ncclGetErrorString);
#undef HODOR_NCCL_SYM
    g_nccl.lib = h;
    return HODOR_OK;
}

static int nccl_fail(ncclResult_t r, const char* what) {
    return fail(HODOR_ERR_CUDA, std::string(what) + ": " + (g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "NCCL error"));
}
#define HODOR_NCCL_TRY(expr)                                  \
    do {                                                      \
        ncclResult_t _r = (expr);                             \
        if (_r != ncclSuccess) return nccl_fail(_r, #expr);   \
    } while (0)

struct Comm {
    ncclComm_t comm = nullptr;  // null when world == 1
    int rank = 0, world = 1;
    uint32_t log_g = 0;
    void* buf[2] = {nullptr, nullptr};  // grow-only exchange buffers of the sharded NTT
    size_t buf_bytes[2] = {0, 0};
    uint64_t bytes_sent = 0;            // payload this rank pushed through NCCL since init (bench evidence)
    uint64_t bytes_peer_stored = 0;     // payload stored straight into peers' buffers by the fused last pass
    // NVLink peer access to every rank's receive buffer buf[1] (CUDA IPC; one process per GPU)
    void* peer_recv[16] = {};           // peer_recv[rank] == buf[1]
    size_t peer_bytes = 0;              // size the mappings were made for (0: none)
    int peer_state = 0;                 // 0 untried, 1 mapped, -1 unavailable (fall back to NCCL send/recv)
    uint4* flag = nullptr;              // 64 B of device scratch for the barrier collectives
    cudaEvent_t pipe_ev[3] = {nullptr, nullptr, nullptr};  // compute stream <-> exchange stream of the pipelined layers
    int ensure(int which, size_t bytes) {
        if (bytes <= buf_bytes[which]) return HODOR_OK;
        if (buf[which]) {
            cudaDeviceSynchronize();
            cudaFree(buf[which]);
            buf[which] = nullptr;
            buf_bytes[which] = 0;
        }
        HODOR_CUDA_TRY(cudaMalloc(&buf[which], bytes));
        buf_bytes[which] = bytes;
        return HODOR_OK;
    }
};

static void close_peers(Comm& cm) {
    for (int r = 0; r < cm.world; r++)
        if (r != cm.rank && cm.peer_recv[r]) cudaIpcCloseMemHandle(cm.peer_recv[r]);
    for (auto& p : cm.peer_recv) p = nullptr;
    cm.peer_bytes = 0;
}

void comm_destroy(Ctx* c) {
    if (!c || !c->comm) return;
    cudaDeviceSynchronize();
    close_peers(*c->comm);
    if (c->comm->flag) cudaFree(c->comm->flag);
    for (auto ev : c->comm->pipe_ev)
        if (ev) cudaEventDestroy(ev);
    if (c->comm->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(c->comm->comm);
    for (int i = 0; i < 2; i++)
        if (c->comm->buf[i]) cudaFree(c->comm->buf[i]);
    delete c->comm;
    c->comm = nullptr;
}

// chunk h of `send` (chunk_bytes each) goes to rank h; chunk g of `recv` comes from rank g
static int all_to_all(Comm& cm, const void* send, void* recv, size_t chunk_bytes, cudaStream_t st) {
    if (cm.world == 1) {
        if (send != recv) HODOR_CUDA_TRY(cudaMemcpyAsync(recv, send, chunk_bytes, cudaMemcpyDeviceToDevice, st));
        return HODOR_OK;
    }
    HODOR_NCCL_TRY(g_nccl.GroupStart());
    for (int p = 0; p < cm.world; p++) {
        HODOR_NCCL_TRY(g_nccl.Send((const char*)send + (size_t)p * chunk_bytes, chunk_bytes, ncclUint8, p, cm.comm, st));
        HODOR_NCCL_TRY(g_nccl.Recv((char*)recv + (size_t)p * chunk_bytes, chunk_bytes, ncclUint8, p, cm.comm, st));
    }
    HODOR_NCCL_TRY(g_nccl.GroupEnd());
    cm.bytes_sent += (uint64_t)(cm.world - 1) * chunk_bytes;
    return HODOR_OK;
}
static int all_gather(Comm& cm, const void* send, void* recv, size_t bytes, cudaStream_t st) {
    if (cm.world == 1) {
        if (send != recv) HODOR_CUDA_TRY(cudaMemcpyAsync(recv, send, bytes, cudaMemcpyDeviceToDevice, st));
        return HODOR_OK;
    }
    HODOR_NCCL_TRY(g_nccl.AllGather(send, recv, bytes, ncclUint8, cm.comm, st));
    cm.bytes_sent += (uint64_t)(cm.world - 1) * bytes;
    return HODOR_OK;
}

// A collective on 32 bytes: every rank has reached this point of its stream (and, stream ordered, finished
// everything before it) when it completes.
static int stream_barrier(Comm& cm, cudaStream_t st) {
    if (cm.world == 1) return HODOR_OK;
    if (!cm.flag) HODOR_CUDA_TRY(cudaMalloc((void**)&cm.flag, 32 * 17));
    HODOR_NCCL_TRY(g_nccl.AllGather(cm.flag, cm.flag + 2, 32, ncclUint8, cm.comm, st));
    return HODOR_OK;
}

// Maps every rank's receive buffer into this process (cudaIpc handles exchanged through NCCL).  Collective:
// every rank calls it with the same `bytes` whenever its buffer was (re)allocated.  On any failure the
// communicator falls back to NCCL send/recv for good (peer_state = -1) -- on every rank, by agreement.
static int map_peers(Comm& cm, size_t bytes, cudaStream_t st) {
    if (cm.peer_state == 1 && cm.peer_bytes == bytes) return HODOR_OK;
    if (cm.peer_state == -1) return HODOR_OK;
    close_peers(cm);
    struct Msg {
        cudaIpcMemHandle_t h;
        int ok;
        char pad[128 - sizeof(cudaIpcMemHandle_t) - sizeof(int)];
    };
    static_assert(sizeof(Msg) == 128, "one 128-byte record per rank");
    Msg mine{};
    mine.ok = (getenv("HODOR_NO_PEER_STORES") == nullptr) && cudaIpcGetMemHandle(&mine.h, cm.buf[1]) == cudaSuccess;
    cudaGetLastError();
    Msg* d = nullptr;
    HODOR_CUDA_TRY(cudaMalloc((void**)&d, sizeof(Msg) * (size_t)(cm.world + 1)));
    std::vector<Msg> all(cm.world);
    int rc = HODOR_OK;
    do {
        if (cudaMemcpyAsync(d, &mine, sizeof(Msg), cudaMemcpyHostToDevice, st) != cudaSuccess) { rc = HODOR_ERR_CUDA; break; }
        if (g_nccl.AllGather(d, d + 1, sizeof(Msg), ncclUint8, cm.comm, st) != ncclSuccess) { rc = HODOR_ERR_CUDA; break; }
        if (cudaMemcpyAsync(all.data(), d + 1, sizeof(Msg) * cm.world, cudaMemcpyDeviceToHost, st) != cudaSuccess) { rc = HODOR_ERR_CUDA; break; }
        if (cudaStreamSynchronize(st) != cudaSuccess) { rc = HODOR_ERR_CUDA; break; }
    } while (0);
    cudaFree(d);
    if (rc) return fail(rc, "peer handle exchange failed");
    bool ok = true;
    for (int r = 0; r < cm.world; r++) ok = ok && all[r].ok;
    int opened = 1;
    if (ok) {
        for (int r = 0; r < cm.world && opened; r++) {
            if (r == cm.rank) {
                cm.peer_recv[r] = cm.buf[1];
            } else if (cudaIpcOpenMemHandle(&cm.peer_recv[r], all[r].h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
                cudaGetLastError();
                cm.peer_recv[r] = nullptr;
                opened = 0;
            }
        }
    } else {
        opened = 0;
    }
    // agree: peer stores only if EVERY rank mapped every buffer
    int* d_ok = nullptr;
    HODOR_CUDA_TRY(cudaMalloc((void**)&d_ok, sizeof(int) * 32 * (size_t)(cm.world + 1)));
    std::vector<int> oks(32 * (size_t)cm.world);
    int mine_ok[32] = {opened};
    cudaMemcpyAsync(d_ok, mine_ok, sizeof(mine_ok), cudaMemcpyHostToDevice, st);
    ncclResult_t nr = g_nccl.AllGather(d_ok, d_ok + 32, sizeof(mine_ok), ncclUint8, cm.comm, st);
    cudaMemcpyAsync(oks.data(), d_ok + 32, sizeof(mine_ok) * cm.world, cudaMemcpyDeviceToHost, st);
    cudaError_t ce = cudaStreamSynchronize(st);
    cudaFree(d_ok);
    if (nr != ncclSuccess || ce != cudaSuccess) return fail(HODOR_ERR_CUDA, "peer agreement failed");
    bool all_ok = true;
    for (int r = 0; r < cm.world; r++) all_ok = all_ok && oks[32 * (size_t)r] == 1;
    if (all_ok) {
        cm.peer_state = 1;
        cm.peer_bytes = bytes;
    } else {
        close_peers(cm);
        cm.peer_state = -1;
    }
    return HODOR_OK;
}

// parts[r][t] = the t-th local element of rank r, blocks of 2^blk_log adjacent elements dealt round-robin:
// natural index i = ((t >> blk_log) * G + r) << blk_log | (t & (B - 1)).  Gathers natural order (the tail of the
// sharded chain: a few tens of thousands of elements).
__global__ void interleave_kernel(const uint4* parts, uint4* out, size_t per_rank, uint32_t log_g, uint32_t blk_log) {
    const size_t total = per_rank << log_g;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const size_t c = i & (((size_t)1 << blk_log) - 1), blk = i >> blk_log;
        const size_t r = blk & (((size_t)1 << log_g) - 1), t = ((blk >> log_g) << blk_log) | c;
        out[2 * i] = parts[2 * (r * per_rank + t)];
        out[2 * i + 1] = parts[2 * (r * per_rank + t) + 1];
    }
}

static inline Fe fe_from(const uint64_t* x) {
    Fe r;
    memcpy(r.v, x, 32);
    return r;
}
static uint32_t log2u64(uint64_t n) {
    uint32_t r = 0;
    while (n >>= 1) r++;
    return r;
}

}  // namespace hodor

using namespace hodor;

extern "C" {

int hodor_cuda_comm_unique_id(uint8_t id[128]) {
    int rc = load_nccl();
    if (rc) return rc;
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes in every NCCL 2.x");
    ncclUniqueId u;
    HODOR_NCCL_TRY(g_nccl.GetUniqueId(&u));
    memcpy(id, &u, 128);
    return HODOR_OK;
}

int hodor_cuda_comm_init(int rank, int world, const uint8_t id[128]) {
    LOCKED_CTX();
    if (world < 1 || world > 16 || (world & (world - 1)) || rank < 0 || rank >= world)
        return fail(HODOR_ERR_INVALID_ARG, "comm_init: world must be a power of two <= 16 and 0 <= rank < world");
    if (c->comm) {
        if (c->comm->rank == rank && c->comm->world == world) return HODOR_OK;
        return fail(HODOR_ERR_INVALID_ARG, "comm_init: already initialised with another geometry (hodor_cuda_comm_destroy first)");
    }
    std::unique_ptr<Comm> cm(new Comm());
    cm->rank = rank;
    cm->world = world;
    cm->log_g = log2u64((uint64_t)world);
    if (world > 1) {
        if (id == nullptr) return fail(HODOR_ERR_INVALID_ARG, "comm_init: NULL unique id");
        int rc = load_nccl();
        if (rc) return rc;
        ncclUniqueId u;
        memcpy(&u, id, 128);
        HODOR_NCCL_TRY(g_nccl.CommInitRank(&cm->comm, world, u, rank));
    }
    c->comm = cm.release();
    return HODOR_OK;
}

void hodor_cuda_comm_destroy(void) {
    Ctx* c = ctx();
    if (!c) return;
    std::lock_guard<std::mutex> lk(c->mu);
    comm_destroy(c);
}

int hodor_cuda_comm_info(int* rank, int* world, uint64_t* bytes_sent, uint64_t* bytes_peer_stored) {
    LOCKED_CTX();
    if (!c->comm) return fail(HODOR_ERR_INVALID_ARG, "hodor_cuda_comm_init() has not been called");
    if (rank) *rank = c->comm->rank;
    if (world) *world = c->comm->world;
    if (bytes_sent) *bytes_sent = c->comm->bytes_sent;
    if (bytes_peer_stored) *bytes_peer_stored = c->comm->bytes_peer_stored;
    return HODOR_OK;
}

// Four-step NTT of length n = 2^log_n over the G ranks, n = m * G:
//   in : rank g holds a_g[j1] = a[j1 * G + g]                     (cyclic slice, m elements)
//   A  : B_g[k1] = omega^(g k1) * sum_j1 a_g[j1] (omega^G)^(j1 k1)  (local length-m NTT, twiddle fused in its last pass)
//   X  : all-to-all of m/G-element chunks: rank h receives B_g[h m/G + k] from every g
//   B  : G-point DFT over g:  out_h[k2 * m/G + k] = A[k2 * m + h * m/G + k]
// so rank h ends with the h-th m/G-chunk of every length-m block of the natural-order result.
int hodor_cuda_ntt_sharded(const void* d_local, void* d_out, uint32_t log_n, const uint64_t omega[4], int field_id, void* stream) {
    LOCKED_CTX();
    GET_OPS(field_id);
    if (!c->comm) return fail(HODOR_ERR_INVALID_ARG, "hodor_cuda_comm_init() has not been called");
    CHECK_DEV_PTRS(d_local, d_out);
    Comm& cm = *c->comm;
    if (log_n > 34 || log_n < 2 * cm.log_g || (cm.log_g && log_n < 4))
        return fail(HODOR_ERR_INVALID_ARG, "ntt_sharded: need 2 * log2(world) <= log_n <= 34");
    cudaStream_t st = (cudaStream_t)stream;
    const size_t m = ((size_t)1 << log_n) >> cm.log_g;
    const Fe w = fe_from(omega);
    Fe wm, wg;
    ops->h_pow(w, (uint64_t)1 << cm.log_g, wm);
    ops->h_pow(w, (uint64_t)cm.rank, wg);
    if (cm.world == 1)  // nothing to exchange, and the G-point DFT is the identity
        return ops->ntt(*c, (const uint4*)d_local, (uint4*)d_out, log_n, 0, wm, nullptr, nullptr, 0, nullptr, st);
    const size_t had = cm.buf_bytes[1];
    // The receive buffer grows (the same decision on every rank: all see the same m) while the peers still map the
    // old one: it is retired, not freed, until every rank has dropped its mapping -- map_peers closes them before its
    // first collective, so once that collective has completed here nobody maps the old buffer any more.
    void* retired = nullptr;
    if (m * 32 > had && cm.peer_state == 1 && cm.peer_bytes != 0) {
        HODOR_CUDA_TRY(cudaDeviceSynchronize());  // this rank's own step B of the previous call still reads it
        retired = cm.buf[1];
        cm.buf[1] = nullptr;
        cm.buf_bytes[1] = 0;
    }
    struct Retire {
        void* p;
        ~Retire() {
            if (p) cudaFree(p);
        }
    } retire{retired};
    int rc = cm.ensure(0, m * 32);
    if (!rc) rc = cm.ensure(1, m * 32);
    if (rc) return rc;
    // multi-pass local transforms only: the single-block kernel (m <= 2^11) has no peer-store path
    const bool want_peers = log_n - cm.log_g >= 12;
    if (retired != nullptr || (want_peers && (cm.peer_state == 0 || (cm.peer_state == 1 && cm.buf_bytes[1] != had)))) {
        rc = map_peers(cm, cm.buf_bytes[1], st);
        if (rc) return rc;
    }
    if (want_peers && cm.peer_state == 1) {
        // X fused into A: the last pass of the local transform stores every output into the receive buffer of the
        // rank that needs it (NVLink peer stores, 256-byte runs), so the exchange overlaps the butterflies tile by
        // tile.  Barrier 1: nobody is still reading its receive buffer from the previous call's step B (it ran
        // before this point on every rank's stream).  Barrier 2: every peer's stores have landed.
        rc = stream_barrier(cm, st);
        if (rc) return rc;
        c->peer_store.on = true;
        c->peer_store.chunk_log = log_n - 2 * cm.log_g;
        c->peer_store.rank = (uint32_t)cm.rank;
        for (int r = 0; r < 16; r++) c->peer_store.base[r] = r < cm.world ? (uint4*)cm.peer_recv[r] : nullptr;
        rc = ops->ntt(*c, (const uint4*)d_local, (uint4*)cm.buf[0], log_n - cm.log_g, 0, wm, nullptr, nullptr, 3, &wg, st);
        c->peer_store.on = false;
        if (rc) return rc;
        cm.bytes_peer_stored += (uint64_t)(cm.world - 1) * (m >> cm.log_g) * 32;
        rc = stream_barrier(cm, st);
        if (rc) return rc;
    } else {
        rc = ops->ntt(*c, (const uint4*)d_local, (uint4*)cm.buf[0], log_n - cm.log_g, 0, wm, nullptr, nullptr, 3, &wg, st);
        if (rc) return rc;
        rc = all_to_all(cm, cm.buf[0], cm.buf[1], (m >> cm.log_g) * 32, st);
        if (rc) return rc;
    }
    return ops->shard_rows(*c, (const uint4*)cm.buf[1], (uint4*)d_out, log_n, cm.log_g, (uint32_t)cm.rank, w, st);
}

// ONE coset LDE 2^log_n -> 2^(log_n + log_factor) plus its FRI commit chain over all ranks (the north star).
// d_coeffs (2^log_n elements) is replicated on every rank.  Rank r computes the B = L/G cosets r*B .. r*B + B - 1
// (capped at 8; cosets are independent) with no communication: blocks of B adjacent elements of the natural-order
// LDE dealt round-robin, v[(k*G + r)*B + c].  FRI fold pairs (idx, idx + M/2) stay on one rank (M/2 is a multiple
// of B*G), so every fold is local and the next layer has the same distribution.  A committed layer: the bottom
// log2 B levels of its tree are local (a block is a complete subtree); ONE all-to-all then turns the level of M/B
// digests from round-robin into natural-order blocks -- B times fewer bytes than exchanging the values, and the
// re-indexing happens inside the next kernel's loads --; the local subtree's root is node G + r of the
// reference's heap; an all-gather of the G sub-roots (32 B each) and the top log2 G levels + root -> challenge run
// redundantly on every GPU; the challenge stays in HBM for the next fold.  Below 2^16 values the strictly serial
// rest of the chain is finished by every rank with the single-GPU chain.
// Outputs (host, identical on every rank): roots (steps + 1) * 32 B, challenges steps * 4 u64, final
// coefficients out_coeffs * 4 u64.  Returns the number of folding steps.
int hodor_cuda_lde_fri_sharded(const void* d_coeffs, uint32_t log_n, uint32_t log_factor, int coset, uint32_t out_coeffs,
                               uint8_t* roots, uint64_t* challenges, uint64_t* final_coeffs, int field_id) {
    LOCKED_CTX();
    GET_OPS(field_id);
    if (!c->comm) return fail(HODOR_ERR_INVALID_ARG, "hodor_cuda_comm_init() has not been called");
    CHECK_DEV_PTRS(d_coeffs);
    Comm& cm = *c->comm;
    const uint32_t G = (uint32_t)cm.world, log_g = cm.log_g;
    if (log_g > log_factor) return fail(HODOR_ERR_INVALID_ARG, "lde_fri_sharded: more ranks than cosets");
    if (log_n + log_factor > 34 || out_coeffs == 0 || (out_coeffs & (out_coeffs - 1)) || ((uint64_t)1 << log_n) < out_coeffs)
        return fail(HODOR_ERR_INVALID_ARG, "lde_fri_sharded: bad sizes");
    const uint32_t log_N = log_n + log_factor;
    const size_t N = (size_t)1 << log_N;
    const int steps = (int)(log_n - log2u64(out_coeffs));
    if (steps < 1) return fail(HODOR_ERR_INVALID_ARG, "lde_fri_sharded: zero folding steps (the reference panics here)");
    Fe omega, coset_omega, mod, one, gen, root, shift0, step;
    int rc = ops->h_domain_generator(log_N, coset_omega);
    if (rc) return fail(rc, "LDE domain larger than the field's 2-adicity");
    ops->h_domain_generator(log_n, omega);
    uint32_t s_, nb_;
    ops->h_constants(mod, one, gen, root, s_, nb_);
    // block size B = 2^blk_log adjacent leaves per rank per round: L / G, at most 8 (the level kernels hash up to 2^3)
    uint32_t blk_log = log_factor - log_g;
    if (blk_log > 3 || getenv("HODOR_SHARD_CYCLIC") != nullptr) blk_log = 0;  // B = 1: plain cyclic slices (cosets r, r + G, ..)
    const uint32_t B = 1u << blk_log;
    if (blk_log) {
        ops->h_pow(coset_omega, (uint64_t)cm.rank * B, shift0);  // cosets rank * B + t, t < B
        step = coset_omega;
    } else {
        ops->h_pow(coset_omega, (uint64_t)cm.rank, shift0);  // cosets rank + G * t
        ops->h_pow(coset_omega, (uint64_t)G, step);
    }
    if (coset) ops->h_mul(shift0, gen, shift0);

    cudaStream_t st = c->stream;
    const size_t m0 = N >> log_g;  // local slice of layer 0
    // one block: values A (m0) | values B (m0/2) | exchange buffer (m0) | local nodes (m0) | top 2G | roots | challenges
    const size_t slots = m0 + m0 / 2 + m0 + m0 + 2 * (size_t)G + 2 * (size_t)(steps + 2);
    uint4* block = (uint4*)c->pool_alloc(slots * 32);
    if (!block) return HODOR_ERR_OOM;
    struct Guard {
        Ctx* c;
        cudaStream_t st;
        void* p;
        ~Guard() {
            cudaStreamSynchronize(st);
            c->pool_free(p);
        }
    } guard{c, st, block};
    uint4* val[2] = {block, block + 2 * m0};
    uint4* xbuf = val[1] + 2 * (m0 / 2);
    uint4* nodes = xbuf + 2 * m0;
    uint4* top = nodes + 2 * m0;
    uint4* d_roots = top + 2 * 2 * (size_t)G;
    uint4* d_chal = d_roots + 2 * (size_t)(steps + 2);

    // with cosets rank*B + t (t < B) the local output d_out[t + B*k] is v[(k*G + rank)*B + t]: the block-cyclic slice
    rc = ops->ntt(*c, (const uint4*)d_coeffs, val[0], log_n, blk_log ? blk_log : log_factor - log_g, omega, &shift0, &step, 0, nullptr, st);
    if (rc) return rc;

    // Layers smaller than this are not worth two collectives each: the rest of the (strictly serial) chain is done
    // redundantly by every rank.  HODOR_SHARD_GATHER_LOG2 overrides (sweep in profiles/r02_experiments.md).
    const int gather_log2 = [] {
        const char* e = getenv("HODOR_SHARD_GATHER_LOG2");
        const int v = e ? atoi(e) : 0;
        return v >= 12 && v <= 30 ? v : 16;
    }();
    size_t gather_below = (size_t)1 << gather_log2;
    if (gather_below < (size_t)4096 * B * G) gather_below = (size_t)4096 * B * G;  // the level kernels need > 1024 digests per rank
    size_t size = N;
    int layer = 0, cur = 0;
    while (!(size < gather_below || layer >= steps - 1)) {
        const size_t ml = size >> log_g;  // local values of this layer
        if (G == 1) {
            rc = do_merkle(*c, ops, val[cur], ml, nodes, nullptr, nullptr, st);
        } else if (blk_log == 0) {
            // exchange the values (cyclic slice -> natural-order block), hash from the received chunks
            rc = all_to_all(cm, val[cur], xbuf, (ml >> log_g) * 32, st);
            if (!rc) rc = do_merkle(*c, ops, xbuf, ml, nodes, nullptr, nullptr, st, log_g, ml >> log_g);
        } else {
            // bottom log2 B levels locally: digest k of this rank is node (M/B) + k*G + rank of the reference's heap
            const size_t w = ml >> blk_log;  // digests per rank == width of this rank's natural-order block of that level
            const size_t wc = w >> log_g;    // digests per destination
            uint4* lvl = nodes + 2 * w;      // position [w, 2w) of `nodes`: free until the subtree above is built
            if (wc >= ((size_t)1 << 15) && getenv("HODOR_SHARD_PIPELINE") != nullptr) {
                // OFF by default (measured on 8 GPUs: 24.5 ms against 16.4 ms for the single grouped all-to-all below --
                // G pairwise steps cost more than they hide; profiles/r02_experiments.md).  Pipelined by destination:
                // the digests for rank (r + s) % G are hashed while those of step s - 1 travel (pairwise send / recv
                // on a second stream: at step s every rank sends to r + s and receives from r - s)
                cudaStream_t cs = c->copy_in;
                if (!cm.pipe_ev[0])
                    for (auto& ev : cm.pipe_ev) HODOR_CUDA_TRY(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
                HODOR_CUDA_TRY(cudaEventRecord(cm.pipe_ev[0], st));  // xbuf's previous readers are done
                HODOR_CUDA_TRY(cudaStreamWaitEvent(cs, cm.pipe_ev[0], 0));
                for (uint32_t s = 0; s < G && !rc; s++) {
                    const uint32_t to = ((uint32_t)cm.rank + s) % G, from = ((uint32_t)cm.rank + G - s) % G;
                    rc = merkle_block_roots(*c, val[cur] + 2 * ((size_t)to * wc << blk_log), wc, (int)blk_log, lvl + 2 * (size_t)to * wc, st);
                    if (rc) break;
                    HODOR_CUDA_TRY(cudaEventRecord(cm.pipe_ev[1], st));
                    HODOR_CUDA_TRY(cudaStreamWaitEvent(cs, cm.pipe_ev[1], 0));
                    if (s == 0) {
                        HODOR_CUDA_TRY(cudaMemcpyAsync(xbuf + 2 * (size_t)cm.rank * wc, lvl + 2 * (size_t)cm.rank * wc, wc * 32,
                                                       cudaMemcpyDeviceToDevice, cs));
                    } else {
                        HODOR_NCCL_TRY(g_nccl.GroupStart());
                        HODOR_NCCL_TRY(g_nccl.Send(lvl + 2 * (size_t)to * wc, wc * 32, ncclUint8, (int)to, cm.comm, cs));
                        HODOR_NCCL_TRY(g_nccl.Recv(xbuf + 2 * (size_t)from * wc, wc * 32, ncclUint8, (int)from, cm.comm, cs));
                        HODOR_NCCL_TRY(g_nccl.GroupEnd());
                        cm.bytes_sent += wc * 32;
                    }
                }
                if (!rc) {
                    HODOR_CUDA_TRY(cudaEventRecord(cm.pipe_ev[2], cs));
                    HODOR_CUDA_TRY(cudaStreamWaitEvent(st, cm.pipe_ev[2], 0));
                }
            } else {
                rc = merkle_block_roots(*c, val[cur], w, (int)blk_log, lvl, st);
                if (!rc) rc = all_to_all(cm, lvl, xbuf, wc * 32, st);
            }
            size_t rem = 0;
            if (!rc) rc = merkle_from_level(*c, xbuf, w, nodes, &rem, st, log_g, wc);
            if (!rc) rc = ops->merkle_tail(*c, nodes, nodes, (uint32_t)rem, false, nullptr, nullptr, st);
        }
        if (rc) return rc;
        HODOR_CUDA_TRY(cudaMemsetAsync(top, 0, 2 * (size_t)G * 32, st));
        rc = all_gather(cm, nodes + 2, top + 2 * (size_t)G, 32, st);
        if (rc) return rc;
        rc = ops->merkle_tail(*c, top, top, G, false, d_roots + 2 * layer, d_chal + 2 * layer, st);
        if (rc) return rc;
        rc = ops->fri_fold(*c, val[cur], ml, log_N, (uint32_t)layer, d_chal + 2 * layer, val[cur ^ 1], (uint64_t)cm.rank * B, G, blk_log, st);
        if (rc) return rc;
        cur ^= 1;
        size >>= 1;
        layer++;
    }
    // the tail: gather the layer on every rank, natural order, single-GPU chain from here
    const size_t ml = size >> log_g;
    const uint4* full = val[cur];
    if (G > 1) {
        rc = all_gather(cm, val[cur], xbuf, ml * 32, st);
        if (rc) return rc;
        interleave_kernel<<<(unsigned)((size + 255) / 256 > 1184 ? 1184 : (size + 255) / 256), 256, 0, st>>>(xbuf, nodes, ml, log_g, blk_log);
        HODOR_CUDA_TRY(cudaGetLastError());
        c->launches++;
        full = nodes;
    }
    HODOR_CUDA_TRY(cudaStreamSynchronize(st));
    std::unique_ptr<hodor_fri_proto, void (*)(hodor_fri_proto*)> tail(
        fri_commit_impl(c, (const uint64_t*)full, size, 1u << log_factor, out_coeffs, 1, field_id), fri_destroy);
    if (!tail) return hodor_cuda_last_error_code() ? hodor_cuda_last_error_code() : HODOR_ERR_CUDA;
    if (tail->steps != steps - layer) return fail(HODOR_ERR_CUDA, "internal: tail chain has the wrong number of steps");
    if (roots) {
        if (layer) HODOR_CUDA_TRY(cudaMemcpyAsync(roots, d_roots, (size_t)layer * 32, cudaMemcpyDeviceToHost, st));
        HODOR_CUDA_TRY(cudaMemcpyAsync(roots + (size_t)layer * 32, tail->roots, (size_t)(tail->steps + 1) * 32, cudaMemcpyDeviceToHost, st));
    }
    if (challenges) {
        if (layer) HODOR_CUDA_TRY(cudaMemcpyAsync(challenges, d_chal, (size_t)layer * 32, cudaMemcpyDeviceToHost, st));
        HODOR_CUDA_TRY(cudaMemcpyAsync(challenges + 4 * (size_t)layer, tail->chal, (size_t)tail->steps * 32, cudaMemcpyDeviceToHost, st));
    }
    if (final_coeffs) HODOR_CUDA_TRY(cudaMemcpyAsync(final_coeffs, tail->final_coeffs, (size_t)out_coeffs * 32, cudaMemcpyDeviceToHost, st));
    HODOR_CUDA_TRY(cudaStreamSynchronize(st));
    return steps;
}

}  // extern "C"
