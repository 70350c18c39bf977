// extern "C" surface of libhodor_b200.so (include/hodor_b200.h).  No torch, no CPU compute path:
// every transform / hash entry point needs an initialised CUDA context and fails without one.
#include <cstdio>
#include <cstring>
#include <memory>
#include <thread>

#include "context.h"

namespace hodor {

static thread_local std::string g_last_error;
static thread_local int g_last_code = 0;
static Ctx* g_ctx = nullptr;
static std::mutex g_ctx_mu;

void set_error(const std::string& msg) { g_last_error = msg; }
int fail(int code, const std::string& msg) {
    set_error(msg);
    g_last_code = code;
    return code;
}
int cuda_fail(cudaError_t e, const char* what) {
    set_error(std::string(what) + ": " + cudaGetErrorString(e));
    cudaGetLastError();  // clear the sticky non-fatal error state
    g_last_code = e == cudaErrorMemoryAllocation ? HODOR_ERR_OOM : HODOR_ERR_CUDA;
    return g_last_code;
}

// The current device is a per-host-thread property: a thread other than the one that called
// hodor_cuda_init would otherwise allocate and launch on device 0.
static thread_local int tl_device = -1;
Ctx* ctx() {
    if (g_ctx == nullptr) {
        fail(HODOR_ERR_CUDA, "hodor_cuda_init() has not been called (or failed): no GPU context, and there is no CPU path");
        return nullptr;
    }
    if (tl_device != g_ctx->device) {
        if (cudaSetDevice(g_ctx->device) != cudaSuccess) {
            cuda_fail(cudaGetLastError(), "cudaSetDevice");
            return nullptr;
        }
        tl_device = g_ctx->device;
    }
    return g_ctx;
}

int Ctx::ensure_workspace(size_t bytes) {
    if (bytes <= ws_bytes) return HODOR_OK;
    if (ws) {
        cudaDeviceSynchronize();
        cudaFree(ws);
        ws = nullptr;
        ws_bytes = 0;
    }
    HODOR_CUDA_TRY(cudaMalloc(&ws, bytes));
    ws_bytes = bytes;
    return HODOR_OK;
}

int Ctx::ws_acquire(size_t bytes, cudaStream_t st) {
    int rc = ensure_workspace(bytes);  // growing synchronises the whole device first
    if (rc) return rc;
    if (ws_used && ws_stream != st) HODOR_CUDA_TRY(cudaStreamWaitEvent(st, ws_event, 0));
    return HODOR_OK;
}
int Ctx::ws_release(cudaStream_t st) {
    if (!ws_event) HODOR_CUDA_TRY(cudaEventCreateWithFlags(&ws_event, cudaEventDisableTiming));
    HODOR_CUDA_TRY(cudaEventRecord(ws_event, st));
    ws_stream = st;
    ws_used = true;
    return HODOR_OK;
}

// Host -> device copy of a caller buffer.  Pinned (cudaHostAlloc / cudaHostRegister) memory goes down as one
// asynchronous copy.  Pageable memory -- what a Rust Vec<F> is -- would be staged by the driver through a single
// thread at ~11 GB/s (measured: 47 ms for the 512 MiB of a 2^24 polynomial); here it is staged through two pinned
// 32 MiB buffers filled by several host threads, so the copy runs at PCIe speed and overlaps the kernels already
// enqueued.  Host-synchronous for pageable sources (the source may be reused on return either way only after the
// caller's own stream synchronisation for pinned ones, as with cudaMemcpyAsync).
int Ctx::h2d(void* dptr, const void* hptr, size_t bytes, cudaStream_t st) {
    if (bytes == 0) return HODOR_OK;
    cudaPointerAttributes attr;
    const bool pinned = cudaPointerGetAttributes(&attr, hptr) == cudaSuccess && attr.type != cudaMemoryTypeUnregistered;
    cudaGetLastError();
    if (pinned || bytes < ((size_t)8 << 20) || stage_threads <= 1) {
        HODOR_CUDA_TRY(cudaMemcpyAsync(dptr, hptr, bytes, cudaMemcpyHostToDevice, st));
        return HODOR_OK;
    }
    constexpr size_t CHUNK = (size_t)32 << 20;
    for (int b = 0; b < 2; b++) {
        if (!stage[b]) {
            HODOR_CUDA_TRY(cudaHostAlloc(&stage[b], CHUNK, cudaHostAllocDefault));
            HODOR_CUDA_TRY(cudaEventCreateWithFlags(&stage_free[b], cudaEventDisableTiming));
        }
    }
    size_t off = 0;
    for (int k = 0; off < bytes; k++, off += CHUNK) {
        const int b = k & 1;
        const size_t len = bytes - off < CHUNK ? bytes - off : CHUNK;
        if (stage_used[b]) HODOR_CUDA_TRY(cudaEventSynchronize(stage_free[b]));  // the last DMA out of this buffer (this call's or an earlier one's) has finished
        const char* src = (const char*)hptr + off;
        char* dst = (char*)stage[b];
        const int T = stage_threads;
        const size_t part = ((len + T - 1) / T + 4095) & ~(size_t)4095;
        std::vector<std::thread> th;
        for (int t = 1; t < T; t++) {
            const size_t lo = (size_t)t * part;
            if (lo >= len) break;
            const size_t n = len - lo < part ? len - lo : part;
            th.emplace_back([=] { memcpy(dst + lo, src + lo, n); });
        }
        memcpy(dst, src, len < part ? len : part);
        for (auto& x : th) x.join();
        HODOR_CUDA_TRY(cudaMemcpyAsync((char*)dptr + off, dst, len, cudaMemcpyHostToDevice, st));
        HODOR_CUDA_TRY(cudaEventRecord(stage_free[b], st));
        stage_used[b] = true;
    }
    return HODOR_OK;
}

int Ctx::ensure_io(int which, size_t bytes) {
    if (bytes <= io_bytes[which]) return HODOR_OK;
    if (io[which]) {
        cudaDeviceSynchronize();
        cudaFree(io[which]);
        io[which] = nullptr;
        io_bytes[which] = 0;
    }
    HODOR_CUDA_TRY(cudaMalloc(&io[which], bytes));
    io_bytes[which] = bytes;
    return HODOR_OK;
}

// Size-bucketed cache of device blocks for objects with caller-controlled lifetime (FRI
// prototypes): a prover commits chain after chain of the same size, so steady state makes no
// driver allocation calls.  Blocks are only recycled after the owner synchronised its stream.
void* Ctx::pool_alloc(size_t bytes) {
    for (size_t i = 0; i < pool_free_list.size(); i++) {
        if (pool_free_list[i].second >= bytes && pool_free_list[i].second <= bytes + bytes / 4) {
            void* p = pool_free_list[i].first;
            pool_live[p] = pool_free_list[i].second;
            pool_free_list.erase(pool_free_list.begin() + i);
            return p;
        }
    }
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e != cudaSuccess) {  // drop the cache and retry once
        cudaGetLastError();
        for (auto& b : pool_free_list) cudaFree(b.first);
        pool_free_list.clear();
        e = cudaMalloc(&p, bytes);
    }
    if (e != cudaSuccess) {
        cuda_fail(e, "cudaMalloc(pool)");
        return nullptr;
    }
    pool_live[p] = bytes;
    return p;
}
void Ctx::pool_free(void* p) {
    auto it = pool_live.find(p);
    if (it == pool_live.end()) {
        cudaFree(p);
        return;
    }
    const size_t bytes = it->second;
    pool_live.erase(it);
    if (bytes > pool_cache_cap) {
        cudaFree(p);
        return;
    }
    // keep the newest block; make room by releasing the oldest cached ones (a caller cycling through one shape
    // reuses its block, whatever other shapes were cached before)
    size_t cached = 0;
    for (auto& b : pool_free_list) cached += b.second;
    while (!pool_free_list.empty() && (pool_free_list.size() >= 16 || cached + bytes > pool_cache_cap)) {
        cached -= pool_free_list.front().second;
        cudaFree(pool_free_list.front().first);
        pool_free_list.erase(pool_free_list.begin());
    }
    pool_free_list.emplace_back(p, bytes);
}

cudaEvent_t Ctx::take_event() {
    if (!event_pool.empty()) {
        cudaEvent_t e = event_pool.back();
        event_pool.pop_back();
        return e;
    }
    cudaEvent_t e = nullptr;
    cudaEventCreate(&e);
    return e;
}

const FieldOps* field_ops(int field_id) {
    switch (field_id) {
        case HODOR_FIELD_BLS12_381_FR: return &kOpsBlsFr;
        case HODOR_FIELD_BN254_FR: return &kOpsBn254Fr;
        case HODOR_FIELD_STARK252: return &kOpsStark252;
    }
    set_error("unknown field_id");
    return nullptr;
}

static inline Fe fe_from_u64(const uint64_t* x) {
    Fe r;
    memcpy(r.v, x, 32);
    return r;
}
static inline void fe_to_u64(const Fe& a, uint64_t* out) { memcpy(out, a.v, 32); }

// RAII device buffer for the host-pointer entry points
struct DevBuf {
    void* p = nullptr;
    ~DevBuf() {
        if (p) cudaFree(p);
    }
    int alloc(size_t bytes) { HODOR_CUDA_TRY(cudaMalloc(&p, bytes)); return HODOR_OK; }
    uint4* u4() const { return (uint4*)p; }
};

static bool is_pow2(uint64_t n) { return n && !(n & (n - 1)); }
static uint32_t log2u(uint64_t n) {
    uint32_t r = 0;
    while (n >>= 1) r++;
    return r;
}

// ---------------------------------------------------------------------------------------------
// internal helpers shared by the host and device variants
// ---------------------------------------------------------------------------------------------
static int do_fft(Ctx& c, const FieldOps* ops, const uint4* in, uint4* out, uint32_t log_n, int coset, cudaStream_t st) {
    Fe omega;
    int rc = ops->h_domain_generator(log_n, omega);
    if (rc) return fail(rc, "domain larger than the field's 2-adicity");
    if (!coset) return ops->ntt(c, in, out, log_n, 0, omega, nullptr, nullptr, 0, nullptr, st);
    Fe mod, one, gen, root;
    uint32_t s, nb;
    ops->h_constants(mod, one, gen, root, s, nb);
    return ops->ntt(c, in, out, log_n, 0, omega, &gen, nullptr, 0, nullptr, st);
}
static int do_ifft(Ctx& c, const FieldOps* ops, const uint4* in, uint4* out, uint32_t log_n, int coset, cudaStream_t st) {
    Fe omega, omega_inv;
    int rc = ops->h_domain_generator(log_n, omega);
    if (rc) return fail(rc, "domain larger than the field's 2-adicity");
    ops->h_inverse(omega, omega_inv);
    if (!coset) return ops->ntt(c, in, out, log_n, 0, omega_inv, nullptr, nullptr, 1, nullptr, st);
    Fe mod, one, gen, root, ginv;
    uint32_t s, nb;
    ops->h_constants(mod, one, gen, root, s, nb);
    ops->h_inverse(gen, ginv);
    return ops->ntt(c, in, out, log_n, 0, omega_inv, nullptr, nullptr, 2, &ginv, st);
}
static int do_lde(Ctx& c, const FieldOps* ops, const uint4* in, uint4* out, uint32_t log_n, uint32_t log_f, int coset,
                  cudaStream_t st) {
    if (log_f == 0) return do_fft(c, ops, in, out, log_n, coset, st);  // factor == 1 (src/polynomials/mod.rs:419,545)
    Fe omega, coset_omega;
    int rc = ops->h_domain_generator(log_n + log_f, coset_omega);
    if (rc) return fail(rc, "LDE domain larger than the field's 2-adicity");
    ops->h_domain_generator(log_n, omega);
    Fe mod, one, gen, root;
    uint32_t s, nb;
    ops->h_constants(mod, one, gen, root, s, nb);
    const Fe shift0 = coset ? gen : one;
    return ops->ntt(c, in, out, log_n, log_f, omega, &shift0, &coset_omega, 0, nullptr, st);
}
int do_merkle(Ctx& c, const FieldOps* ops, const uint4* leaves, size_t n, uint4* nodes, uint4* root, uint4* chal,
              cudaStream_t st, uint32_t leaf_log_g, size_t leaf_chunk) {
    if (!is_pow2(n) || n < 2) return fail(HODOR_ERR_INVALID_ARG, "merkle: leaf count must be a power of two >= 2");
    size_t w = 0;
    int rc = merkle_levels(c, leaves, n, nodes, &w, st, leaf_log_g, leaf_chunk);
    if (rc) return rc;
    if (w == 0) return ops->merkle_tail(c, leaves, nodes, (uint32_t)n, true, root, chal, st);
    return ops->merkle_tail(c, nodes, nodes, (uint32_t)w, false, root, chal, st);
}

// Releases everything a context owns (also the partially built one of a failed hodor_cuda_init).  Blocks still
// held by live handles (pool_live) belong to their owners and are not touched.
static void ctx_destroy(Ctx* c) {
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    comm_destroy(c);
    for (auto& kv : c->pow_tables) cudaFree(kv.second.block);
    for (auto& kv : c->ntt_tables) {
        cudaFree(kv.second.pw.block);
        cudaFree(kv.second.tw_b_block);
        cudaFree(kv.second.tw_direct_block);
    }
    if (c->ws) cudaFree(c->ws);
    for (int i = 0; i < 2; i++)
        if (c->io[i]) cudaFree(c->io[i]);
    for (auto& b : c->pool_free_list) cudaFree(b.first);
    for (auto& kv : c->full_tables) cudaFree(kv.second.first);
    for (auto& r : c->prof) {
        cudaEventDestroy(r.start);
        cudaEventDestroy(r.stop);
    }
    for (auto e : c->event_pool) cudaEventDestroy(e);
    if (c->ws_event) cudaEventDestroy(c->ws_event);
    if (c->small) cudaFree(c->small);
    if (c->pinned_small) cudaFreeHost(c->pinned_small);
    for (int b = 0; b < 2; b++) {
        if (c->stage[b]) cudaFreeHost(c->stage[b]);
        if (c->stage_free[b]) cudaEventDestroy(c->stage_free[b]);
    }
    for (cudaStream_t st : {c->stream, c->commit_stream, c->copy_in, c->copy_out})
        if (st) cudaStreamDestroy(st);
    cudaGetLastError();
    delete c;
}

}  // namespace hodor

using namespace hodor;

extern "C" {

// ---- context -----------------------------------------------------------------------------------
int hodor_cuda_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

int hodor_cuda_init(int device) {
    std::lock_guard<std::mutex> lk(g_ctx_mu);
    if (g_ctx) {
        if (g_ctx->device == device) return HODOR_OK;
        return fail(HODOR_ERR_INVALID_ARG, "already initialised on another device (one process drives one GPU)");
    }
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        cudaGetLastError();
        return fail(HODOR_ERR_CUDA, "no CUDA device available; hodor_b200 has no CPU fallback");
    }
    if (device < 0 || device >= n) return fail(HODOR_ERR_INVALID_ARG, "device index out of range");
    HODOR_CUDA_TRY(cudaSetDevice(device));
    tl_device = device;
    cudaDeviceProp prop;
    HODOR_CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        char buf[160];
        snprintf(buf, sizeof buf, "device %d is sm_%d%d; this library carries sm_100a code only", device, prop.major, prop.minor);
        return fail(HODOR_ERR_CUDA, buf);
    }
    std::unique_ptr<Ctx, void (*)(Ctx*)> c(new Ctx(), ctx_destroy);  // an early return below releases what was created
    c->device = device;
    int prio_least = 0, prio_greatest = 0;
    HODOR_CUDA_TRY(cudaDeviceGetStreamPriorityRange(&prio_least, &prio_greatest));
    {
        // default: transforms at the higher priority, the hashing stream fills what they leave free
        int p_main = prio_greatest, p_commit = prio_least;
        if (const char* e = getenv("HODOR_COMMIT_PRIORITY")) {
            if (!strcmp(e, "high")) { p_main = prio_least; p_commit = prio_greatest; }
            else if (!strcmp(e, "equal")) { p_main = p_commit = prio_greatest; }
        }
        HODOR_CUDA_TRY(cudaStreamCreateWithPriority(&c->stream, cudaStreamNonBlocking, p_main));
        HODOR_CUDA_TRY(cudaStreamCreateWithPriority(&c->commit_stream, cudaStreamNonBlocking, p_commit));
    }
    if (const char* e = getenv("HODOR_CONCURRENT_COMMIT")) c->concurrent_commit = atoi(e) != 0;
    HODOR_CUDA_TRY(cudaStreamCreateWithFlags(&c->copy_in, cudaStreamNonBlocking));
    HODOR_CUDA_TRY(cudaStreamCreateWithFlags(&c->copy_out, cudaStreamNonBlocking));
    HODOR_CUDA_TRY(cudaMalloc((void**)&c->small, 4096));
    HODOR_CUDA_TRY(cudaHostAlloc((void**)&c->pinned_small, Ctx::PINNED_SMALL_BYTES, cudaHostAllocDefault));
    c->key = b2s_keyed_state();
    if (const char* mb = getenv("HODOR_TABLE_BUDGET_MB")) c->full_budget = (size_t)strtoull(mb, nullptr, 10) << 20;
    if (const char* e = getenv("HODOR_STAGE_THREADS")) c->stage_threads = atoi(e);
    if (c->stage_threads == 0) {
        const unsigned hw = std::thread::hardware_concurrency();
        c->stage_threads = hw >= 16 ? 8 : (hw >= 4 ? (int)hw / 2 : 1);
    }
    if (const char* e = getenv("HODOR_FUSE_FOLD_COMMIT")) c->fuse_fold_commit = atoi(e) != 0;
    if (const char* e = getenv("HODOR_FUSE_LAST_COMMIT")) c->fuse_last_commit = atoi(e);
    if (const char* mb = getenv("HODOR_POOL_CACHE_MB")) c->pool_cache_cap = (size_t)strtoull(mb, nullptr, 10) << 20;
    g_ctx = c.release();
    return HODOR_OK;
}

void hodor_cuda_shutdown(void) {
    std::lock_guard<std::mutex> lk(g_ctx_mu);
    if (!g_ctx) return;
    ctx_destroy(g_ctx);
    g_ctx = nullptr;
}

const char* hodor_cuda_last_error(void) { return g_last_error.c_str(); }
int hodor_cuda_last_error_code(void) { return g_last_code; }

// Releases the cached (not live) pool blocks back to the driver.
int hodor_cuda_trim(void) {
    LOCKED_CTX();
    HODOR_CUDA_TRY(cudaDeviceSynchronize());
    for (auto& b : c->pool_free_list) cudaFree(b.first);
    c->pool_free_list.clear();
    return HODOR_OK;
}

size_t hodor_cuda_workspace_bytes(void) {
    Ctx* c = g_ctx;
    return c ? c->ws_bytes + c->table_bytes + c->full_bytes + c->io_bytes[0] + c->io_bytes[1] : 0;
}

uint64_t hodor_cuda_launch_count(void) {
    Ctx* c = g_ctx;
    return c ? c->launches.load() : 0;
}

// ---- host scalars --------------------------------------------------------------------------------
int hodor_field_constants(int field_id, uint64_t modulus[4], uint64_t one[4], uint64_t generator[4],
                          uint64_t root_of_unity[4], uint32_t* s, uint32_t* num_bits, uint32_t* capacity) {
    GET_OPS(field_id);
    Fe m, o, g, r;
    uint32_t ss, nb;
    ops->h_constants(m, o, g, r, ss, nb);
    if (modulus) fe_to_u64(m, modulus);
    if (one) fe_to_u64(o, one);
    if (generator) fe_to_u64(g, generator);
    if (root_of_unity) fe_to_u64(r, root_of_unity);
    if (s) *s = ss;
    if (num_bits) *num_bits = nb;
    if (capacity) *capacity = nb - 1;
    return HODOR_OK;
}
int hodor_domain_generator(int field_id, uint32_t log_n, uint64_t out[4]) {
    GET_OPS(field_id);
    Fe g;
    int rc = ops->h_domain_generator(log_n, g);
    if (rc) return fail(rc, "domain larger than the field's 2-adicity");
    fe_to_u64(g, out);
    return HODOR_OK;
}
int hodor_field_mul(int field_id, const uint64_t a[4], const uint64_t b[4], uint64_t out[4]) {
    GET_OPS(field_id);
    Fe r;
    ops->h_mul(fe_from_u64(a), fe_from_u64(b), r);
    fe_to_u64(r, out);
    return HODOR_OK;
}
int hodor_field_add(int field_id, const uint64_t a[4], const uint64_t b[4], uint64_t out[4]) {
    GET_OPS(field_id);
    Fe r;
    ops->h_add(fe_from_u64(a), fe_from_u64(b), r);
    fe_to_u64(r, out);
    return HODOR_OK;
}
int hodor_field_sub(int field_id, const uint64_t a[4], const uint64_t b[4], uint64_t out[4]) {
    GET_OPS(field_id);
    Fe r;
    ops->h_sub(fe_from_u64(a), fe_from_u64(b), r);
    fe_to_u64(r, out);
    return HODOR_OK;
}
int hodor_field_pow(int field_id, const uint64_t a[4], uint64_t e, uint64_t out[4]) {
    GET_OPS(field_id);
    Fe r;
    ops->h_pow(fe_from_u64(a), e, r);
    fe_to_u64(r, out);
    return HODOR_OK;
}
int hodor_field_inverse(int field_id, const uint64_t a[4], uint64_t out[4]) {
    GET_OPS(field_id);
    Fe r;
    int rc = ops->h_inverse(fe_from_u64(a), r);
    if (rc) return fail(rc, "inverse of zero");
    fe_to_u64(r, out);
    return HODOR_OK;
}
int hodor_field_from_repr(int field_id, const uint64_t plain[4], uint64_t out[4]) {
    GET_OPS(field_id);
    Fe r;
    ops->h_from_repr(fe_from_u64(plain), r);
    fe_to_u64(r, out);
    return HODOR_OK;
}
int hodor_field_into_repr(int field_id, const uint64_t mont[4], uint64_t out[4]) {
    GET_OPS(field_id);
    Fe r;
    ops->h_into_repr(fe_from_u64(mont), r);
    fe_to_u64(r, out);
    return HODOR_OK;
}
int hodor_root_to_challenge(const uint8_t root[32], uint64_t out[4], int field_id) {
    GET_OPS(field_id);
    Fe r;
    int rc = ops->h_root_to_challenge(root, r);
    if (rc) return fail(rc, "digest does not reduce into the field");
    fe_to_u64(r, out);
    return HODOR_OK;
}
// O(log n) verifier-side hashing (IopTree::verify / get_path's leaf-pair hash): single hashes on the
// host with the same compression code the kernels use, compiled for the host.
int hodor_hash_leaf(const uint64_t leaf[4], uint8_t out[32]) {
    static const B2sState key = b2s_keyed_state();
    uint32_t w[8];
    memcpy(w, leaf, 32);
    const Digest d = hash_leaf32(key, w);
    memcpy(out, d.w, 32);
    return HODOR_OK;
}
int hodor_hash_node(const uint8_t left[32], const uint8_t right[32], uint8_t out[32]) {
    static const B2sState key = b2s_keyed_state();
    Digest l, r;
    memcpy(l.w, left, 32);
    memcpy(r.w, right, 32);
    const Digest d = hash_node64(key, l, r);
    memcpy(out, d.w, 32);
    return HODOR_OK;
}

// ---- memory --------------------------------------------------------------------------------------
void* hodor_cuda_malloc(size_t bytes) {
    if (!ctx()) return nullptr;
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e != cudaSuccess) {
        cuda_fail(e, "cudaMalloc");
        return nullptr;
    }
    return p;
}
void hodor_cuda_free(void* dptr) {
    if (dptr) cudaFree(dptr);
}
void* hodor_cuda_host_alloc(size_t bytes) {
    if (!ctx()) return nullptr;
    void* p = nullptr;
    cudaError_t e = cudaHostAlloc(&p, bytes, cudaHostAllocDefault);
    if (e != cudaSuccess) {
        cuda_fail(e, "cudaHostAlloc");
        return nullptr;
    }
    return p;
}
void hodor_cuda_host_free(void* hptr) {
    if (hptr) cudaFreeHost(hptr);
}
// The caller's stream handle is used verbatim: NULL is CUDA's legacy default stream (what
// torch.cuda.current_stream().cuda_stream returns for the default stream), not the library's own.
static cudaStream_t pick_stream(Ctx* c, void* stream) {
    (void)c;
    return (cudaStream_t)stream;
}
int hodor_cuda_memcpy_h2d(void* dptr, const void* hptr, size_t bytes, void* stream) {
    Ctx* c = ctx();
    if (!c) return HODOR_ERR_CUDA;
    HODOR_CUDA_TRY(cudaMemcpyAsync(dptr, hptr, bytes, cudaMemcpyHostToDevice, pick_stream(c, stream)));
    return HODOR_OK;
}
int hodor_cuda_memcpy_d2h(void* hptr, const void* dptr, size_t bytes, void* stream) {
    Ctx* c = ctx();
    if (!c) return HODOR_ERR_CUDA;
    HODOR_CUDA_TRY(cudaMemcpyAsync(hptr, dptr, bytes, cudaMemcpyDeviceToHost, pick_stream(c, stream)));
    return HODOR_OK;
}
int hodor_cuda_stream_synchronize(void* stream) {
    Ctx* c = ctx();
    if (!c) return HODOR_ERR_CUDA;
    HODOR_CUDA_TRY(cudaStreamSynchronize(pick_stream(c, stream)));
    return HODOR_OK;
}

// ---- device variants -------------------------------------------------------------------------------
int hodor_cuda_ntt_dev(const void* d_in, void* d_out, uint32_t log_n, const uint64_t omega[4], int field_id, void* stream) {
    LOCKED_CTX();
    GET_OPS(field_id);
    CHECK_DEV_PTRS(d_in, d_out);
    return ops->ntt(*c, (const uint4*)d_in, (uint4*)d_out, log_n, 0, fe_from_u64(omega), nullptr, nullptr, 0, nullptr,
                    pick_stream(c, stream));
}
int hodor_cuda_fft_dev(const void* d_in, void* d_out, uint32_t log_n, int coset, int field_id, void* stream) {
    LOCKED_CTX();
    GET_OPS(field_id);
    CHECK_DEV_PTRS(d_in, d_out);
    return do_fft(*c, ops, (const uint4*)d_in, (uint4*)d_out, log_n, coset, pick_stream(c, stream));
}
int hodor_cuda_ifft_dev(const void* d_in, void* d_out, uint32_t log_n, int coset, int field_id, void* stream) {
    LOCKED_CTX();
    GET_OPS(field_id);
    CHECK_DEV_PTRS(d_in, d_out);
    return do_ifft(*c, ops, (const uint4*)d_in, (uint4*)d_out, log_n, coset, pick_stream(c, stream));
}
int hodor_cuda_lde_dev(const void* d_coeffs, uint32_t log_n, uint32_t log_factor, int coset, void* d_out, int field_id,
                       void* stream) {
    LOCKED_CTX();
    GET_OPS(field_id);
    CHECK_DEV_PTRS(d_coeffs, d_out);
    if (d_coeffs == d_out && log_factor) return fail(HODOR_ERR_INVALID_ARG, "lde: input and output must not alias");
    return do_lde(*c, ops, (const uint4*)d_coeffs, (uint4*)d_out, log_n, log_factor, coset, pick_stream(c, stream));
}
int hodor_cuda_merkle_build_dev(const void* d_leaves, uint64_t n, void* d_nodes, void* d_root, void* d_challenge,
                                int field_id, void* stream) {
    LOCKED_CTX();
    GET_OPS(field_id);
    CHECK_DEV_PTRS(d_leaves, d_nodes, d_root, d_challenge);
    return do_merkle(*c, ops, (const uint4*)d_leaves, n, (uint4*)d_nodes, (uint4*)d_root, (uint4*)d_challenge,
                     pick_stream(c, stream));
}
int hodor_cuda_merkle_build_shard_dev(const void* d_chunks, uint64_t n, uint32_t log_g, void* d_nodes, void* d_root,
                                      void* d_challenge, int field_id, void* stream) {
    LOCKED_CTX();
    GET_OPS(field_id);
    CHECK_DEV_PTRS(d_chunks, d_nodes, d_root, d_challenge);
    if (log_g > 4 || !is_pow2(n) || (n >> log_g) <= 1024)
        return fail(HODOR_ERR_INVALID_ARG, "merkle_build_shard: need log_g <= 4 and more than 1024 leaves per chunk");
    return do_merkle(*c, ops, (const uint4*)d_chunks, n, (uint4*)d_nodes, (uint4*)d_root, (uint4*)d_challenge,
                     pick_stream(c, stream), log_g, n >> log_g);
}
int hodor_cuda_merkle_top_dev(void* d_nodes, uint64_t w, void* d_root, void* d_challenge, int field_id, void* stream) {
    LOCKED_CTX();
    GET_OPS(field_id);
    CHECK_DEV_PTRS(d_nodes, d_root, d_challenge);
    if (!is_pow2(w) || w > 4096) return fail(HODOR_ERR_INVALID_ARG, "merkle_top: width must be a power of two <= 4096");
    return ops->merkle_tail(*c, (const uint4*)d_nodes, (uint4*)d_nodes, (uint32_t)w, false, (uint4*)d_root, (uint4*)d_challenge,
                            pick_stream(c, stream));
}
int hodor_cuda_fri_fold_dev(const void* d_in, uint64_t n, uint64_t initial_domain_size, uint32_t layer,
                            const void* d_challenge, void* d_out, int field_id, void* stream) {
    LOCKED_CTX();
    GET_OPS(field_id);
    CHECK_DEV_PTRS(d_in, d_challenge, d_out);
    if (!is_pow2(n) || n < 2 || !is_pow2(initial_domain_size) || (initial_domain_size >> layer) != n)
        return fail(HODOR_ERR_INVALID_ARG, "fri_fold: n must equal initial_domain_size >> layer, both powers of two");
    return ops->fri_fold(*c, (const uint4*)d_in, n, log2u(initial_domain_size), layer, (const uint4*)d_challenge,
                         (uint4*)d_out, 0, 1, 0, pick_stream(c, stream));
}
int hodor_cuda_fri_fold_shard_dev(const void* d_in, uint64_t n_local, uint64_t initial_domain_size, uint32_t layer,
                                  uint32_t log_g, uint32_t rank, const void* d_challenge, void* d_out, int field_id,
                                  void* stream) {
    LOCKED_CTX();
    GET_OPS(field_id);
    CHECK_DEV_PTRS(d_in, d_challenge, d_out);
    const uint64_t g = (uint64_t)1 << log_g;
    if (!is_pow2(n_local) || n_local < 2 || !is_pow2(initial_domain_size) || rank >= g ||
        ((initial_domain_size >> layer) >> log_g) != n_local)
        return fail(HODOR_ERR_INVALID_ARG, "fri_fold_shard: n_local must equal (initial_domain_size >> layer) / G, >= 2");
    return ops->fri_fold(*c, (const uint4*)d_in, n_local, log2u(initial_domain_size), layer, (const uint4*)d_challenge,
                         (uint4*)d_out, rank, g, 0, pick_stream(c, stream));
}
int hodor_cuda_lde_cosets_dev(const void* d_coeffs, uint32_t log_n, uint32_t log_factor, int coset, uint32_t first_coset,
                              uint32_t coset_stride, uint32_t log_count, void* d_out, int field_id, void* stream) {
    LOCKED_CTX();
    GET_OPS(field_id);
    CHECK_DEV_PTRS(d_coeffs, d_out);
    const uint64_t L = (uint64_t)1 << log_factor, cnt = (uint64_t)1 << log_count;
    if (log_count > log_factor || coset_stride == 0 || first_coset + (cnt - 1) * coset_stride >= L)
        return fail(HODOR_ERR_INVALID_ARG, "lde_cosets: coset subset out of range");
    if (d_coeffs == d_out) return fail(HODOR_ERR_INVALID_ARG, "lde_cosets: input and output must not alias");
    Fe omega, coset_omega, mod, one, gen, root, shift0, step;
    int rc = ops->h_domain_generator(log_n + log_factor, coset_omega);
    if (rc) return fail(rc, "LDE domain larger than the field's 2-adicity");
    ops->h_domain_generator(log_n, omega);
    uint32_t s, nb;
    ops->h_constants(mod, one, gen, root, s, nb);
    ops->h_pow(coset_omega, first_coset, shift0);  // shift_i = [g] * w_{nL}^i, i = first + stride * t
    if (coset) ops->h_mul(shift0, gen, shift0);
    ops->h_pow(coset_omega, coset_stride, step);
    return ops->ntt(*c, (const uint4*)d_coeffs, (uint4*)d_out, log_n, log_count, omega, &shift0, &step, 0, nullptr,
                    pick_stream(c, stream));
}
int hodor_cuda_distribute_powers_dev(void* d_a, uint64_t n, const uint64_t g[4], int field_id, void* stream) {
    LOCKED_CTX();
    GET_OPS(field_id);
    CHECK_DEV_PTRS(d_a);
    if (n == 0) return HODOR_OK;
    return ops->scale_pow(*c, (uint4*)d_a, (size_t)n, fe_from_u64(g), pick_stream(c, stream));
}
// ---- setup work of Prover::new on the device: twiddle vectors and the ALI inverse divisors ------------------------
static int precomputed_omegas_impl(Ctx* c, const FieldOps* ops, void* d_omegas, void* d_coset, void* d_omegas_inv,
                                   uint32_t log_n, cudaStream_t st) {
    if (log_n > 40) return fail(HODOR_ERR_DOMAIN, "precomputed_omegas: domain too large");
    Fe omega, omega_inv, modulus, one, gen, root;
    uint32_t s2 = 0, bits = 0;
    if (ops->h_domain_generator(log_n, omega)) return fail(HODOR_ERR_DOMAIN, "precomputed_omegas: size exceeds the field's 2-adicity");
    ops->h_constants(modulus, one, gen, root, s2, bits);
    if (ops->h_inverse(omega, omega_inv)) return fail(HODOR_ERR_NOT_INVERTIBLE, "precomputed_omegas: generator not invertible");
    const size_t n = (size_t)1 << log_n;
    int rc = HODOR_OK;
    if (d_omegas) rc = ops->coset_map(*c, 0, (uint4*)d_omegas, n, omega, one, nullptr, nullptr, 0, nullptr, 0, st);
    if (!rc && d_coset) rc = ops->coset_map(*c, 0, (uint4*)d_coset, n, omega, gen, nullptr, nullptr, 0, nullptr, 0, st);
    if (!rc && d_omegas_inv) rc = ops->coset_map(*c, 0, (uint4*)d_omegas_inv, n / 2, omega_inv, one, nullptr, nullptr, 0, nullptr, 0, st);
    return rc;
}
static int ali_dense_impl(Ctx* c, const FieldOps* ops, void* d_out, uint32_t log_column, uint32_t log_evaluation,
                          uint64_t start_at, uint64_t span, uint64_t num_rows, uint64_t* divisor_degree, cudaStream_t st) {
    if (d_out == nullptr) return fail(HODOR_ERR_INVALID_ARG, "ali_dense_inverse_divisor: d_out is NULL");
    if (log_column > 40 || log_evaluation > 40 || log_evaluation < log_column)
        return fail(HODOR_ERR_INVALID_ARG, "ali_dense_inverse_divisor: need column domain <= evaluation domain");
    const uint64_t T = 1ull << log_column, E = 1ull << log_evaluation;
    // the reference's usize subtractions (src/ali/per_register/mod.rs:71-75) would underflow (debug panic) otherwise
    if (num_rows > T || span > num_rows || start_at > num_rows - span)
        return fail(HODOR_ERR_INVALID_ARG, "ali_dense_inverse_divisor: start_at + span must fit into num_rows <= column domain");
    const uint64_t num_roots = start_at + (T - (num_rows - span));
    if (num_roots > (1u << 20)) return fail(HODOR_ERR_INVALID_ARG, "ali_dense_inverse_divisor: more than 2^20 excluded rows");
    Fe w_col, w_eval, modulus, one, gen, root;
    uint32_t s2 = 0, bits = 0;
    if (ops->h_domain_generator(log_column, w_col) || ops->h_domain_generator(log_evaluation, w_eval))
        return fail(HODOR_ERR_DOMAIN, "ali_dense_inverse_divisor: size exceeds the field's 2-adicity");
    ops->h_constants(modulus, one, gen, root, s2, bits);
    // roots of the divisor's numerator: omega^k, k in [0, start_at) and [num_rows - span, T)   (:77-93)
    std::vector<Fe> roots;
    roots.reserve((size_t)num_roots);
    Fe r = one;
    for (uint64_t k = 0; k < start_at; k++) {
        roots.push_back(r);
        ops->h_mul(r, w_col, r);
    }
    ops->h_pow(w_col, num_rows - span, r);
    for (uint64_t k = num_rows - span; k < T; k++) {
        roots.push_back(r);
        ops->h_mul(r, w_col, r);
    }
    // x_j^T - 1 on the coset g * <w_eval>: g^T * (w_eval^T)^j, w_eval^T of order E/T -> E/T distinct values   (:112-130)
    const uint64_t L = E / T;
    std::vector<Fe> inv_van((size_t)L);
    Fe gT, wL, cur;
    ops->h_pow(gen, T, gT);
    ops->h_pow(w_eval, T, wL);
    cur = gT;
    for (uint64_t k = 0; k < L; k++) {
        Fe v;
        ops->h_sub(cur, one, v);
        if (ops->h_inverse(v, inv_van[(size_t)k]))
            return fail(HODOR_ERR_NOT_INVERTIBLE, "ali_dense_inverse_divisor: X^T - 1 vanishes on the evaluation coset");
        ops->h_mul(cur, wL, cur);
    }
    if (divisor_degree) *divisor_degree = T - start_at - (T - num_rows) - span;
    return ops->coset_map(*c, 2, (uint4*)d_out, (size_t)E, w_eval, gen, nullptr, roots.data(), (uint32_t)roots.size(), inv_van.data(),
                          (uint32_t)L, st);
}
static int ali_boundary_impl(Ctx* c, const FieldOps* ops, void* d_out, uint32_t log_column, uint32_t log_evaluation,
                             uint64_t row, cudaStream_t st) {
    if (d_out == nullptr) return fail(HODOR_ERR_INVALID_ARG, "ali_boundary_inverse_divisor: d_out is NULL");
    if (log_column > 40 || log_evaluation > 40) return fail(HODOR_ERR_INVALID_ARG, "ali_boundary_inverse_divisor: domain too large");
    Fe w_col, w_eval, modulus, one, gen, root, at;
    uint32_t s2 = 0, bits = 0;
    if (ops->h_domain_generator(log_column, w_col) || ops->h_domain_generator(log_evaluation, w_eval))
        return fail(HODOR_ERR_DOMAIN, "ali_boundary_inverse_divisor: size exceeds the field's 2-adicity");
    ops->h_constants(modulus, one, gen, root, s2, bits);
    ops->h_pow(w_col, row, at);
    const size_t n = (size_t)1 << log_evaluation;
    int rc = ops->coset_map(*c, 1, (uint4*)d_out, n, w_eval, gen, &at, nullptr, 0, nullptr, 0, st);
    if (rc) return rc;
    int* d_status = (int*)(c->small + 64);
    rc = ops->batch_inversion(*c, (uint4*)d_out, n, d_status, st);
    if (rc) return rc;
    int* status = (int*)c->pinned_small;
    HODOR_CUDA_TRY(cudaMemcpyAsync(status, d_status, sizeof(int), cudaMemcpyDeviceToHost, st));
    HODOR_CUDA_TRY(cudaStreamSynchronize(st));
    if (*status != 0) return fail(HODOR_ERR_NOT_INVERTIBLE, "ali_boundary_inverse_divisor: X - omega^row vanishes on the coset");
    return HODOR_OK;
}
int hodor_cuda_precomputed_omegas_dev(void* d_omegas, void* d_coset, void* d_omegas_inv, uint32_t log_n, int field_id,
                                      void* stream) {
    LOCKED_CTX();
    GET_OPS(field_id);
    CHECK_DEV_PTRS(d_omegas, d_coset, d_omegas_inv);
    return precomputed_omegas_impl(c, ops, d_omegas, d_coset, d_omegas_inv, log_n, pick_stream(c, stream));
}
int hodor_cuda_ali_dense_inverse_divisor_dev(void* d_out, uint32_t log_column, uint32_t log_evaluation, uint64_t start_at,
                                             uint64_t span, uint64_t num_rows, uint64_t* divisor_degree, int field_id,
                                             void* stream) {
    LOCKED_CTX();
    GET_OPS(field_id);
    CHECK_DEV_PTRS(d_out);
    return ali_dense_impl(c, ops, d_out, log_column, log_evaluation, start_at, span, num_rows, divisor_degree, pick_stream(c, stream));
}
int hodor_cuda_ali_boundary_inverse_divisor_dev(void* d_out, uint32_t log_column, uint32_t log_evaluation, uint64_t row,
                                                int field_id, void* stream) {
    LOCKED_CTX();
    GET_OPS(field_id);
    CHECK_DEV_PTRS(d_out);
    return ali_boundary_impl(c, ops, d_out, log_column, log_evaluation, row, pick_stream(c, stream));
}
// host-vector forms (what a Rust caller holding Vec<F> binds): computed in the library's I/O buffer, copied out
int hodor_cuda_precomputed_omegas(uint64_t* omegas, uint64_t* coset, uint64_t* omegas_inv, uint32_t log_n, int field_id) {
    LOCKED_CTX();
    GET_OPS(field_id);
    if (log_n > 34) return fail(HODOR_ERR_DOMAIN, "precomputed_omegas: domain too large");
    const size_t n = (size_t)1 << log_n;
    int rc = c->ensure_io(0, n * 32);
    if (rc) return rc;
    uint64_t* outs[3] = {omegas, coset, omegas_inv};
    for (int k = 0; k < 3; k++) {
        if (!outs[k]) continue;
        const size_t cnt = k == 2 ? n / 2 : n;
        if (cnt == 0) continue;
        rc = precomputed_omegas_impl(c, ops, k == 0 ? c->io[0] : nullptr, k == 1 ? c->io[0] : nullptr, k == 2 ? c->io[0] : nullptr,
                                     log_n, c->stream);
        if (rc) return rc;
        HODOR_CUDA_TRY(cudaMemcpyAsync(outs[k], c->io[0], cnt * 32, cudaMemcpyDeviceToHost, c->stream));
        HODOR_CUDA_TRY(cudaStreamSynchronize(c->stream));
    }
    return HODOR_OK;
}
int hodor_cuda_ali_dense_inverse_divisor(uint64_t* out, uint32_t log_column, uint32_t log_evaluation, uint64_t start_at,
                                         uint64_t span, uint64_t num_rows, uint64_t* divisor_degree, int field_id) {
    LOCKED_CTX();
    GET_OPS(field_id);
    if (out == nullptr || log_evaluation > 34) return fail(HODOR_ERR_INVALID_ARG, "ali_dense_inverse_divisor: bad output / domain too large");
    const size_t bytes = ((size_t)1 << log_evaluation) * 32;
    int rc = c->ensure_io(0, bytes);
    if (rc) return rc;
    rc = ali_dense_impl(c, ops, c->io[0], log_column, log_evaluation, start_at, span, num_rows, divisor_degree, c->stream);
    if (rc) return rc;
    HODOR_CUDA_TRY(cudaMemcpyAsync(out, c->io[0], bytes, cudaMemcpyDeviceToHost, c->stream));
    HODOR_CUDA_TRY(cudaStreamSynchronize(c->stream));
    return HODOR_OK;
}
int hodor_cuda_ali_boundary_inverse_divisor(uint64_t* out, uint32_t log_column, uint32_t log_evaluation, uint64_t row,
                                            int field_id) {
    LOCKED_CTX();
    GET_OPS(field_id);
    if (out == nullptr || log_evaluation > 34) return fail(HODOR_ERR_INVALID_ARG, "ali_boundary_inverse_divisor: bad output / domain too large");
    const size_t bytes = ((size_t)1 << log_evaluation) * 32;
    int rc = c->ensure_io(0, bytes);
    if (rc) return rc;
    rc = ali_boundary_impl(c, ops, c->io[0], log_column, log_evaluation, row, c->stream);
    if (rc) return rc;
    HODOR_CUDA_TRY(cudaMemcpyAsync(out, c->io[0], bytes, cudaMemcpyDeviceToHost, c->stream));
    HODOR_CUDA_TRY(cudaStreamSynchronize(c->stream));
    return HODOR_OK;
}
int hodor_cuda_elementwise_dev(int op, const void* d_a, const void* d_b, void* d_out, uint64_t n, int field_id,
                               void* stream) {
    LOCKED_CTX();
    GET_OPS(field_id);
    CHECK_DEV_PTRS(d_a, d_b, d_out);
    if (op < 0 || op > 3) return fail(HODOR_ERR_INVALID_ARG, "elementwise: op must be 0..3 (use hodor_cuda_poly_op_dev for the scalar forms)");
    return ops->elementwise(*c, op, (const uint4*)d_a, (const uint4*)d_b, (uint4*)d_out, n, nullptr, 0, pick_stream(c, stream));
}
int hodor_cuda_poly_op_dev(int op, const void* d_a, const void* d_b, const uint64_t scalar[4], uint64_t exp, void* d_out,
                           uint64_t n, int field_id, void* stream) {
    LOCKED_CTX();
    GET_OPS(field_id);
    CHECK_DEV_PTRS(d_a, d_b, d_out);
    Fe s;
    if (scalar) s = fe_from_u64(scalar);
    return ops->elementwise(*c, op, (const uint4*)d_a, (const uint4*)d_b, (uint4*)d_out, n, scalar ? &s : nullptr, exp,
                            pick_stream(c, stream));
}
int hodor_cuda_batch_inversion_dev(void* d_a, uint64_t n, int* d_status, int field_id, void* stream) {
    LOCKED_CTX();
    GET_OPS(field_id);
    CHECK_DEV_PTRS(d_a);
    if (d_status == nullptr) return fail(HODOR_ERR_INVALID_ARG, "batch_inversion: d_status is NULL");
    return ops->batch_inversion(*c, (uint4*)d_a, (size_t)n, d_status, pick_stream(c, stream));
}
int hodor_cuda_evaluate_at_dev(const void* d_coeffs, uint64_t n, const uint64_t g[4], void* d_out, int field_id,
                               void* stream) {
    LOCKED_CTX();
    GET_OPS(field_id);
    CHECK_DEV_PTRS(d_coeffs, d_out);
    return ops->evaluate_at(*c, (const uint4*)d_coeffs, (size_t)n, fe_from_u64(g), (uint4*)d_out, pick_stream(c, stream));
}
int hodor_cuda_ntt_shard_cols_dev(const void* d_in, void* d_out, uint32_t log_n, uint32_t log_g, uint32_t rank,
                                  const uint64_t omega[4], int field_id, void* stream) {
    LOCKED_CTX();
    GET_OPS(field_id);
    CHECK_DEV_PTRS(d_in, d_out);
    if (log_g > 4 || log_n < 2 * log_g || rank >= (1u << log_g)) return fail(HODOR_ERR_INVALID_ARG, "shard_cols: bad geometry");
    // local transform of length m = n/G with root omega^G, then B_g[k] *= (omega^g)^k
    Fe w = fe_from_u64(omega), wm, wg;
    ops->h_pow(w, (uint64_t)1 << log_g, wm);
    ops->h_pow(w, rank, wg);
    return ops->ntt(*c, (const uint4*)d_in, (uint4*)d_out, log_n - log_g, 0, wm, nullptr, nullptr, log_g ? 3 : 0, &wg,
                    pick_stream(c, stream));
}
int hodor_cuda_ntt_shard_rows_dev(const void* d_in, void* d_out, uint32_t log_n, uint32_t log_g, uint32_t rank,
                                  const uint64_t omega[4], int field_id, void* stream) {
    LOCKED_CTX();
    GET_OPS(field_id);
    CHECK_DEV_PTRS(d_in, d_out);
    return ops->shard_rows(*c, (const uint4*)d_in, (uint4*)d_out, log_n, log_g, rank, fe_from_u64(omega),
                           pick_stream(c, stream));
}

// ---- host variants: H2D, the device variant, D2H ---------------------------------------------------
// Staging goes through the context's grow-only device buffers (io[0] input, io[1] output), so a
// caller looping over same-sized vectors pays cudaMalloc once.
static int host_inplace(Ctx* c, uint64_t* a, size_t n, int (*body)(Ctx&, uint4*, void*), void* arg) {
    int rc = c->ensure_io(0, n * 32);
    if (rc) return rc;
    uint4* d = (uint4*)c->io[0];
    { int hrc = c->h2d(d, a, n * 32, c->stream); if (hrc) return hrc; }
    rc = body(*c, d, arg);
    if (rc) return rc;
    HODOR_CUDA_TRY(cudaMemcpyAsync(a, d, n * 32, cudaMemcpyDeviceToHost, c->stream));
    HODOR_CUDA_TRY(cudaStreamSynchronize(c->stream));
    return HODOR_OK;
}

int hodor_cuda_ntt(uint64_t* a, uint32_t log_n, const uint64_t omega[4], int field_id) {
    LOCKED_CTX();
    GET_OPS(field_id);
    if (log_n > 32) return fail(HODOR_ERR_INVALID_ARG, "log_n > 32");
    struct A { const FieldOps* ops; uint32_t log_n; Fe omega; } arg{ops, log_n, fe_from_u64(omega)};
    return host_inplace(c, a, (size_t)1 << log_n, [](Ctx& cc, uint4* d, void* p) {
        A* a = (A*)p;
        return a->ops->ntt(cc, d, d, a->log_n, 0, a->omega, nullptr, nullptr, 0, nullptr, cc.stream);
    }, &arg);
}
int hodor_cuda_fft(uint64_t* a, uint32_t log_n, int coset, int field_id) {
    LOCKED_CTX();
    GET_OPS(field_id);
    if (log_n > 32) return fail(HODOR_ERR_INVALID_ARG, "log_n > 32");
    struct A { const FieldOps* ops; uint32_t log_n; int coset; } arg{ops, log_n, coset};
    return host_inplace(c, a, (size_t)1 << log_n, [](Ctx& cc, uint4* d, void* p) {
        A* a = (A*)p;
        return do_fft(cc, a->ops, d, d, a->log_n, a->coset, cc.stream);
    }, &arg);
}
int hodor_cuda_ifft(uint64_t* a, uint32_t log_n, int coset, int field_id) {
    LOCKED_CTX();
    GET_OPS(field_id);
    if (log_n > 32) return fail(HODOR_ERR_INVALID_ARG, "log_n > 32");
    struct A { const FieldOps* ops; uint32_t log_n; int coset; } arg{ops, log_n, coset};
    return host_inplace(c, a, (size_t)1 << log_n, [](Ctx& cc, uint4* d, void* p) {
        A* a = (A*)p;
        return do_ifft(cc, a->ops, d, d, a->log_n, a->coset, cc.stream);
    }, &arg);
}
int hodor_cuda_distribute_powers(uint64_t* a, uint64_t n, const uint64_t g[4], int field_id) {
    LOCKED_CTX();
    GET_OPS(field_id);
    if (n == 0) return HODOR_OK;
    struct A { const FieldOps* ops; size_t n; Fe g; } arg{ops, (size_t)n, fe_from_u64(g)};
    return host_inplace(c, a, (size_t)n, [](Ctx& cc, uint4* d, void* p) {
        A* a = (A*)p;
        return a->ops->scale_pow(cc, d, a->n, a->g, cc.stream);
    }, &arg);
}
int hodor_cuda_lde(const uint64_t* coeffs, uint32_t log_n, uint32_t log_factor, int coset, uint64_t* out, int field_id) {
    LOCKED_CTX();
    GET_OPS(field_id);
    if (log_n > 32 || log_n + log_factor > 34) return fail(HODOR_ERR_INVALID_ARG, "lde too large");
    const size_t n = (size_t)1 << log_n, total = n << log_factor;
    int rc = c->ensure_io(0, n * 32);
    if (rc) return rc;
    rc = c->ensure_io(1, total * 32);
    if (rc) return rc;
    { int hrc = c->h2d(c->io[0], coeffs, n * 32, c->stream); if (hrc) return hrc; }
    rc = do_lde(*c, ops, (const uint4*)c->io[0], (uint4*)c->io[1], log_n, log_factor, coset, c->stream);
    if (rc) return rc;
    HODOR_CUDA_TRY(cudaMemcpyAsync(out, c->io[1], total * 32, cudaMemcpyDeviceToHost, c->stream));
    HODOR_CUDA_TRY(cudaStreamSynchronize(c->stream));
    return HODOR_OK;
}
// `count` LDEs of the same shape, software-pipelined over three streams with double-buffered device
// staging: H2D of polynomial i+1 and D2H of polynomial i-1 run while polynomial i is transformed, so
// a prover lifting all its registers (src/prover/mod.rs:73-76 loops `w.lde(..)`) pays
// max(PCIe, compute) per polynomial instead of their sum.  Host buffers should be pinned
// (hodor_cuda_host_alloc) -- pageable memory makes the copies synchronous and serialises the pipe.
int hodor_cuda_lde_batch(const uint64_t* const* coeffs, uint64_t* const* outs, uint32_t count, uint32_t log_n,
                         uint32_t log_factor, int coset, int field_id) {
    LOCKED_CTX();
    GET_OPS(field_id);
    if (log_n > 32 || log_n + log_factor > 34) return fail(HODOR_ERR_INVALID_ARG, "LDE too large");
    if (count == 0) return HODOR_OK;
    if (coeffs == nullptr || outs == nullptr) return fail(HODOR_ERR_INVALID_ARG, "lde_batch: NULL pointer table");
    for (uint32_t i = 0; i < count; i++)
        if (coeffs[i] == nullptr || outs[i] == nullptr) return fail(HODOR_ERR_INVALID_ARG, "lde_batch: NULL buffer");
    const size_t n = (size_t)1 << log_n, total = n << log_factor;
    const uint32_t nbuf = count < 2 ? 1 : 2;
    void* in_buf[2] = {nullptr, nullptr};
    void* out_buf[2] = {nullptr, nullptr};
    cudaEvent_t in_done[2], comp_done[2], out_done[2];
    int rc = HODOR_OK;
    for (uint32_t b = 0; b < nbuf && rc == HODOR_OK; b++) {
        in_buf[b] = c->pool_alloc(n * 32);
        out_buf[b] = c->pool_alloc(total * 32);
        if (!in_buf[b] || !out_buf[b]) rc = HODOR_ERR_OOM;
    }
    for (uint32_t b = 0; b < nbuf; b++) {
        cudaEventCreateWithFlags(&in_done[b], cudaEventDisableTiming);
        cudaEventCreateWithFlags(&comp_done[b], cudaEventDisableTiming);
        cudaEventCreateWithFlags(&out_done[b], cudaEventDisableTiming);
    }
    auto step = [&](uint32_t i) -> int {
        const uint32_t b = i % nbuf;
        if (i >= nbuf) HODOR_CUDA_TRY(cudaStreamWaitEvent(c->copy_in, comp_done[b], 0));  // in_buf[b] consumed
        { int hrc = c->h2d(in_buf[b], coeffs[i], n * 32, c->copy_in); if (hrc) return hrc; }
        HODOR_CUDA_TRY(cudaEventRecord(in_done[b], c->copy_in));
        HODOR_CUDA_TRY(cudaStreamWaitEvent(c->stream, in_done[b], 0));
        if (i >= nbuf) HODOR_CUDA_TRY(cudaStreamWaitEvent(c->stream, out_done[b], 0));  // out_buf[b] drained
        int r = do_lde(*c, ops, (const uint4*)in_buf[b], (uint4*)out_buf[b], log_n, log_factor, coset, c->stream);
        if (r) return r;
        HODOR_CUDA_TRY(cudaEventRecord(comp_done[b], c->stream));
        HODOR_CUDA_TRY(cudaStreamWaitEvent(c->copy_out, comp_done[b], 0));
        HODOR_CUDA_TRY(cudaMemcpyAsync(outs[i], out_buf[b], total * 32, cudaMemcpyDeviceToHost, c->copy_out));
        HODOR_CUDA_TRY(cudaEventRecord(out_done[b], c->copy_out));
        return HODOR_OK;
    };
    for (uint32_t i = 0; i < count && rc == HODOR_OK; i++) rc = step(i);
    // drain everything before the staging buffers go back to the pool, error or not
    cudaStreamSynchronize(c->copy_in);
    cudaStreamSynchronize(c->stream);
    cudaError_t e = cudaStreamSynchronize(c->copy_out);
    for (uint32_t b = 0; b < nbuf; b++) {
        cudaEventDestroy(in_done[b]);
        cudaEventDestroy(comp_done[b]);
        cudaEventDestroy(out_done[b]);
        if (in_buf[b]) c->pool_free(in_buf[b]);
        if (out_buf[b]) c->pool_free(out_buf[b]);
    }
    if (rc) return rc;
    if (e != cudaSuccess) return cuda_fail(e, "lde_batch");
    return HODOR_OK;
}
static int host_poly_op(Ctx* c, const FieldOps* ops, int op, const uint64_t* a, const uint64_t* b, const uint64_t* scalar,
                        uint64_t exp, uint64_t* out, uint64_t n) {
    if (n == 0) return HODOR_OK;
    if (op < 0 || op >= EW_NUM_OPS) return fail(HODOR_ERR_INVALID_ARG, "unknown elementwise op");
    const bool needs_b = op <= EW_ADD_SCALED;
    if (needs_b && b == nullptr) return fail(HODOR_ERR_INVALID_ARG, "elementwise: operand b is NULL");
    const size_t nb = !needs_b ? 0 : (op == EW_SCALE ? 1 : (size_t)n);
    int rc = c->ensure_io(0, n * 32);
    if (rc) return rc;
    rc = c->ensure_io(1, (nb ? nb : 1) * 32);
    if (rc) return rc;
    HODOR_CUDA_TRY(cudaMemcpyAsync(c->io[0], a, n * 32, cudaMemcpyHostToDevice, c->stream));
    if (nb) HODOR_CUDA_TRY(cudaMemcpyAsync(c->io[1], b, nb * 32, cudaMemcpyHostToDevice, c->stream));
    Fe s;
    if (scalar) s = fe_from_u64(scalar);
    rc = ops->elementwise(*c, op, (const uint4*)c->io[0], (const uint4*)c->io[1], (uint4*)c->io[0], n, scalar ? &s : nullptr,
                          exp, c->stream);
    if (rc) return rc;
    HODOR_CUDA_TRY(cudaMemcpyAsync(out, c->io[0], n * 32, cudaMemcpyDeviceToHost, c->stream));
    HODOR_CUDA_TRY(cudaStreamSynchronize(c->stream));
    return HODOR_OK;
}
int hodor_cuda_elementwise(int op, const uint64_t* a, const uint64_t* b, uint64_t* out, uint64_t n, int field_id) {
    LOCKED_CTX();
    GET_OPS(field_id);
    if (op < 0 || op > 3) return fail(HODOR_ERR_INVALID_ARG, "elementwise: op must be 0..3 (use hodor_cuda_poly_op for the scalar forms)");
    return host_poly_op(c, ops, op, a, b, nullptr, 0, out, n);
}
int hodor_cuda_poly_op(int op, const uint64_t* a, const uint64_t* b, const uint64_t scalar[4], uint64_t exp, uint64_t* out,
                       uint64_t n, int field_id) {
    LOCKED_CTX();
    GET_OPS(field_id);
    return host_poly_op(c, ops, op, a, b, scalar, exp, out, n);
}
int hodor_cuda_selftest_mul_pre(int field_id) {
    LOCKED_CTX();
    GET_OPS(field_id);
    unsigned long long* d = (unsigned long long*)(c->small + 70);
    int rc = ops->selftest_mul_pre(*c, d, c->stream);
    if (rc) return rc;
    unsigned long long bad = 0;
    HODOR_CUDA_TRY(cudaMemcpyAsync(&bad, d, sizeof(bad), cudaMemcpyDeviceToHost, c->stream));
    HODOR_CUDA_TRY(cudaStreamSynchronize(c->stream));
    return bad > 0x7fffffffull ? 0x7fffffff : (int)bad;
}
int hodor_cuda_batch_inversion(uint64_t* a, uint64_t n, int field_id) {
    LOCKED_CTX();
    GET_OPS(field_id);
    if (n == 0) return HODOR_OK;
    int rc = c->ensure_io(0, n * 32);
    if (rc) return rc;
    int* d_status = (int*)(c->small + 64);  // scalar scratch, beyond the slots the FRI chain uses
    HODOR_CUDA_TRY(cudaMemcpyAsync(c->io[0], a, n * 32, cudaMemcpyHostToDevice, c->stream));
    rc = ops->batch_inversion(*c, (uint4*)c->io[0], (size_t)n, d_status, c->stream);
    if (rc) return rc;
    int status = 0;
    HODOR_CUDA_TRY(cudaMemcpyAsync(&status, d_status, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    HODOR_CUDA_TRY(cudaStreamSynchronize(c->stream));
    if (status != 0) return fail(HODOR_ERR_NOT_INVERTIBLE, "batch_inversion: the vector contains zero");
    HODOR_CUDA_TRY(cudaMemcpyAsync(a, c->io[0], n * 32, cudaMemcpyDeviceToHost, c->stream));
    HODOR_CUDA_TRY(cudaStreamSynchronize(c->stream));
    return HODOR_OK;
}
int hodor_cuda_evaluate_at(const uint64_t* coeffs, uint64_t n, const uint64_t g[4], uint64_t out[4], int field_id) {
    LOCKED_CTX();
    GET_OPS(field_id);
    int rc = c->ensure_io(0, (n ? n : 1) * 32);
    if (rc) return rc;
    uint4* d_out = c->small + 66;
    if (n) HODOR_CUDA_TRY(cudaMemcpyAsync(c->io[0], coeffs, n * 32, cudaMemcpyHostToDevice, c->stream));
    rc = ops->evaluate_at(*c, (const uint4*)c->io[0], (size_t)n, fe_from_u64(g), d_out, c->stream);
    if (rc) return rc;
    HODOR_CUDA_TRY(cudaMemcpyAsync(out, d_out, 32, cudaMemcpyDeviceToHost, c->stream));
    HODOR_CUDA_TRY(cudaStreamSynchronize(c->stream));
    return HODOR_OK;
}
int hodor_cuda_merkle_build(const uint64_t* leaves, uint64_t n, uint8_t* nodes, int field_id) {
    LOCKED_CTX();
    GET_OPS(field_id);
    if (!is_pow2(n) || n < 2) return fail(HODOR_ERR_INVALID_ARG, "merkle: leaf count must be a power of two >= 2");
    int rc = c->ensure_io(0, n * 32);
    if (rc) return rc;
    rc = c->ensure_io(1, n * 32);
    if (rc) return rc;
    { int hrc = c->h2d(c->io[0], leaves, n * 32, c->stream); if (hrc) return hrc; }
    rc = do_merkle(*c, ops, (const uint4*)c->io[0], n, (uint4*)c->io[1], nullptr, nullptr, c->stream);
    if (rc) return rc;
    HODOR_CUDA_TRY(cudaMemcpyAsync(nodes, c->io[1], n * 32, cudaMemcpyDeviceToHost, c->stream));
    HODOR_CUDA_TRY(cudaStreamSynchronize(c->stream));
    return HODOR_OK;
}

// ---- committed oracles: values + Merkle tree resident in HBM behind a handle ------------------------
// What Prover::prove does with every register: `w.lde(..)` then `I::create(&lde)` (src/prover/mod.rs:73-80,
// :91-95), keeping both for the query phase (:142-151, IOP::query src/iop/blake2s_trivial_iop.rs:324-338).
// Only the root (32 B) crosses PCIe at commit time and log2(n) digests + one value per query.
struct hodor_tree {
    int field_id = 0;
    uint64_t n = 0;
    uint4* block = nullptr;         // one allocation: [values] | nodes | root | challenge | path scratch
    const uint4* values = nullptr;  // inside `block`, or borrowed from the caller
    uint4* nodes = nullptr;
    uint4* root = nullptr;
    uint4* chal = nullptr;
    uint4* path = nullptr;          // 64 digests + 1 element per query, TREE_QUERY_SLOTS queries
};
static constexpr uint32_t TREE_QUERY_SLOTS = 64;

static void tree_destroy(hodor_tree* t) {
    if (!t) return;
    Ctx* c = g_ctx;
    if (t->block) c ? c->pool_free(t->block) : (void)cudaFree(t->block);
    delete t;
}

// allocates the handle; own_values: room for n values in front of the nodes
static hodor_tree* tree_alloc(Ctx* c, int field_id, uint64_t n, bool own_values) {
    std::unique_ptr<hodor_tree, void (*)(hodor_tree*)> t(new hodor_tree(), tree_destroy);
    t->field_id = field_id;
    t->n = n;
    const size_t slots = (own_values ? n : 0) + n + 2 + (size_t)TREE_QUERY_SLOTS * 66;
    t->block = (uint4*)c->pool_alloc(slots * 32);
    if (!t->block) return nullptr;
    uint4* cur = t->block;
    if (own_values) {
        t->values = cur;
        cur += 2 * n;
    }
    t->nodes = cur;
    cur += 2 * n;
    t->root = cur;
    t->chal = cur + 2;
    t->path = cur + 4;
    return t.release();
}

hodor_tree* hodor_cuda_tree_commit(const uint64_t* values, uint64_t n, int values_on_device, uint8_t root[32], int field_id) {
    Ctx* c = ctx();
    if (!c) return nullptr;
    std::lock_guard<std::mutex> lk(c->mu);
    const FieldOps* ops = field_ops(field_id);
    if (!ops) return nullptr;
    if (!is_pow2(n) || n < 2 || values == nullptr) {  // reference: assert!(num_leafs == num_leafs.next_power_of_two())
        fail(HODOR_ERR_INVALID_ARG, "tree_commit: leaf count must be a power of two >= 2");
        return nullptr;
    }
    if (values_on_device && !aligned32({values})) {
        fail(HODOR_ERR_INVALID_ARG, "tree_commit: device element arrays must be 32-byte aligned");
        return nullptr;
    }
    std::unique_ptr<hodor_tree, void (*)(hodor_tree*)> t(tree_alloc(c, field_id, n, !values_on_device), tree_destroy);
    if (!t) return nullptr;
    cudaStream_t st = c->stream;
    if (values_on_device) {
        t->values = (const uint4*)values;
    } else if (c->h2d((void*)t->values, values, n * 32, st)) {
        return nullptr;
    }
    int rc = do_merkle(*c, ops, t->values, n, t->nodes, t->root, t->chal, st);
    if (!rc && root && cudaMemcpyAsync(root, t->root, 32, cudaMemcpyDeviceToHost, st) != cudaSuccess)
        rc = cuda_fail(cudaGetLastError(), "cudaMemcpyAsync(root)");
    if (!rc && cudaStreamSynchronize(st) != cudaSuccess) rc = cuda_fail(cudaGetLastError(), "tree_commit sync");
    if (rc) return nullptr;
    return t.release();
}

// `count` polynomials of one shape: H2D of polynomial i+1 (second stream, double-buffered staging)
// overlaps the LDE + tree build of polynomial i.  trees[i] / roots + 32*i receive the results.
int hodor_cuda_lde_commit_batch(const uint64_t* const* coeffs, uint32_t count, uint32_t log_n, uint32_t log_factor, int coset,
                                int coeffs_on_device, hodor_tree** trees, uint8_t* roots, int field_id) {
    LOCKED_CTX();
    GET_OPS(field_id);
    if (log_n > 32 || log_n + log_factor > 34 || log_n + log_factor < 1)
        return fail(HODOR_ERR_INVALID_ARG, "lde_commit: need 2 <= n * factor <= 2^34");
    if (count == 0) return HODOR_OK;
    if (coeffs == nullptr || trees == nullptr) return fail(HODOR_ERR_INVALID_ARG, "lde_commit: NULL pointer table");
    for (uint32_t i = 0; i < count; i++) {
        if (coeffs[i] == nullptr) return fail(HODOR_ERR_INVALID_ARG, "lde_commit: NULL buffer");
        if (coeffs_on_device) CHECK_DEV_PTRS(coeffs[i]);
        trees[i] = nullptr;
    }
    {
        Fe probe;
        int rc = ops->h_domain_generator(log_n + log_factor, probe);
        if (rc) return fail(rc, "LDE domain larger than the field's 2-adicity");
    }
    const size_t n = (size_t)1 << log_n, total = n << log_factor;
    const uint32_t nbuf = coeffs_on_device ? 0 : (count < 2 ? 1 : 2);
    void* in_buf[2] = {nullptr, nullptr};
    cudaEvent_t in_done[2] = {nullptr, nullptr}, comp_done[2] = {nullptr, nullptr};
    int rc = HODOR_OK;
    for (uint32_t b = 0; b < nbuf && rc == HODOR_OK; b++) {
        in_buf[b] = c->pool_alloc(n * 32);
        if (!in_buf[b]) rc = HODOR_ERR_OOM;
        cudaEventCreateWithFlags(&in_done[b], cudaEventDisableTiming);
        cudaEventCreateWithFlags(&comp_done[b], cudaEventDisableTiming);
    }
    const bool roots_pinned = roots != nullptr && (size_t)count * 32 <= Ctx::PINNED_SMALL_BYTES;
    const bool overlap = c->concurrent_commit && count > 1 && total >= ((size_t)1 << 20);
    cudaEvent_t lde_done[2] = {nullptr, nullptr};
    if (overlap)
        for (auto& ev : lde_done) cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
    auto step = [&](uint32_t i) -> int {
        hodor_tree* t = tree_alloc(c, field_id, total, true);
        if (!t) return HODOR_ERR_OOM;
        trees[i] = t;
        const uint4* src = (const uint4*)coeffs[i];
        if (!coeffs_on_device) {
            const uint32_t b = i % nbuf;
            if (i >= nbuf) HODOR_CUDA_TRY(cudaStreamWaitEvent(c->copy_in, comp_done[b], 0));  // in_buf[b] consumed
            { int hrc = c->h2d(in_buf[b], coeffs[i], n * 32, c->copy_in); if (hrc) return hrc; }
            HODOR_CUDA_TRY(cudaEventRecord(in_done[b], c->copy_in));
            HODOR_CUDA_TRY(cudaStreamWaitEvent(c->stream, in_done[b], 0));
            src = (const uint4*)in_buf[b];
        }
        // fused build: the last pass of the transform hashes the bottom three levels of the tree (ntt_commit.cuh)
        c->fuse_commit.nodes = c->fuse_last_commit > 0 && total >= ((size_t)1 << 13) ? t->nodes : nullptr;
        c->fuse_commit.done = false;
        int r = do_lde(*c, ops, src, (uint4*)t->values, log_n, log_factor, coset, c->stream);
        const bool fused = c->fuse_commit.done;
        c->fuse_commit.nodes = nullptr;
        c->fuse_commit.done = false;
        if (r) return r;
        if (!coeffs_on_device) HODOR_CUDA_TRY(cudaEventRecord(comp_done[i % nbuf], c->stream));
        if (fused) {  // levels total/2 .. total/8 are written: the rest of the tree, root and challenge
            size_t w = 0;
            r = merkle_upper_levels(*c, t->nodes, total >> 3, &w, c->stream);
            if (!r) r = ops->merkle_tail(*c, t->nodes, t->nodes, (uint32_t)w, false, t->root, t->chal, c->stream);
            if (r) return r;
            if (roots) {
                uint8_t* dst = roots_pinned ? c->pinned_small + 32 * (size_t)i : roots + 32 * (size_t)i;
                HODOR_CUDA_TRY(cudaMemcpyAsync(dst, t->root, 32, cudaMemcpyDeviceToHost, c->stream));
            }
            return HODOR_OK;
        }
        // the tree: beside the next polynomial's transform when there is one (see Ctx::commit_stream)
        cudaStream_t ts = c->stream;
        if (overlap) {
            HODOR_CUDA_TRY(cudaEventRecord(lde_done[i % 2], c->stream));
            HODOR_CUDA_TRY(cudaStreamWaitEvent(c->commit_stream, lde_done[i % 2], 0));
            ts = c->commit_stream;
            c->merkle_backfill = true;
        }
        r = do_merkle(*c, ops, t->values, total, t->nodes, t->root, t->chal, ts);
        c->merkle_backfill = false;
        if (r) return r;
        if (roots) {  // via pinned scratch while it lasts: see Ctx::pinned_small
            uint8_t* dst = roots_pinned ? c->pinned_small + 32 * (size_t)i : roots + 32 * (size_t)i;
            HODOR_CUDA_TRY(cudaMemcpyAsync(dst, t->root, 32, cudaMemcpyDeviceToHost, ts));
        }
        return HODOR_OK;
    };
    for (uint32_t i = 0; i < count && rc == HODOR_OK; i++) rc = step(i);
    cudaStreamSynchronize(c->copy_in);
    cudaError_t e = cudaStreamSynchronize(c->stream);
    if (overlap) {
        const cudaError_t e2 = cudaStreamSynchronize(c->commit_stream);
        if (e == cudaSuccess) e = e2;
        for (auto& ev : lde_done)
            if (ev) cudaEventDestroy(ev);
    }
    for (uint32_t b = 0; b < nbuf; b++) {
        if (in_done[b]) cudaEventDestroy(in_done[b]);
        if (comp_done[b]) cudaEventDestroy(comp_done[b]);
        if (in_buf[b]) c->pool_free(in_buf[b]);
    }
    if (rc == HODOR_OK && e != cudaSuccess) rc = cuda_fail(e, "lde_commit");
    if (rc == HODOR_OK && roots_pinned) memcpy(roots, c->pinned_small, (size_t)count * 32);
    if (rc) {
        for (uint32_t i = 0; i < count; i++) {
            tree_destroy(trees[i]);
            trees[i] = nullptr;
        }
    }
    return rc;
}

hodor_tree* hodor_cuda_lde_commit(const uint64_t* coeffs, uint32_t log_n, uint32_t log_factor, int coset, int coeffs_on_device,
                                  uint8_t root[32], int field_id) {
    hodor_tree* t = nullptr;
    const uint64_t* tab[1] = {coeffs};
    if (hodor_cuda_lde_commit_batch(tab, 1, log_n, log_factor, coset, coeffs_on_device, &t, root, field_id)) return nullptr;
    return t;
}

void hodor_cuda_tree_free(hodor_tree* t) {
    Ctx* c = g_ctx;
    if (c) {
        std::lock_guard<std::mutex> lk(c->mu);
        cudaStreamSynchronize(c->stream);
        tree_destroy(t);
    } else {
        tree_destroy(t);
    }
}
uint64_t hodor_cuda_tree_size(const hodor_tree* t) { return t ? t->n : 0; }
const void* hodor_cuda_tree_values(const hodor_tree* t) { return t ? (const void*)t->values : nullptr; }
const void* hodor_cuda_tree_nodes(const hodor_tree* t) { return t ? (const void*)t->nodes : nullptr; }
int hodor_cuda_tree_root(const hodor_tree* t, uint8_t root[32], uint64_t challenge[4]) {
    LOCKED_CTX();
    if (!t) return fail(HODOR_ERR_INVALID_ARG, "null tree handle");
    if (root) HODOR_CUDA_TRY(cudaMemcpyAsync(root, t->root, 32, cudaMemcpyDeviceToHost, c->stream));
    if (challenge) HODOR_CUDA_TRY(cudaMemcpyAsync(challenge, t->chal, 32, cudaMemcpyDeviceToHost, c->stream));
    HODOR_CUDA_TRY(cudaStreamSynchronize(c->stream));
    return HODOR_OK;
}
int hodor_cuda_tree_read(const hodor_tree* t, uint64_t first, uint64_t count, uint64_t* values, uint8_t* nodes) {
    LOCKED_CTX();
    if (!t) return fail(HODOR_ERR_INVALID_ARG, "null tree handle");
    if (first > t->n || count > t->n - first) return fail(HODOR_ERR_INVALID_ARG, "tree_read: range out of bounds");
    if (values) HODOR_CUDA_TRY(cudaMemcpyAsync(values, t->values + 2 * first, count * 32, cudaMemcpyDeviceToHost, c->stream));
    if (nodes) HODOR_CUDA_TRY(cudaMemcpyAsync(nodes, t->nodes + 2 * first, count * 32, cudaMemcpyDeviceToHost, c->stream));
    HODOR_CUDA_TRY(cudaStreamSynchronize(c->stream));
    return HODOR_OK;
}
// `count` queries in one launch and one pair of copies: values[i] (4 u64) and paths + i * 32 * log2(n)
int hodor_cuda_tree_query_batch(const hodor_tree* t, const uint64_t* natural_indices, uint32_t count, uint64_t* values,
                                uint8_t* paths) {
    LOCKED_CTX();
    if (!t) return fail(HODOR_ERR_INVALID_ARG, "null tree handle");
    if (count && natural_indices == nullptr) return fail(HODOR_ERR_INVALID_ARG, "tree_query: NULL index table");
    for (uint32_t i = 0; i < count; i++)
        if (natural_indices[i] >= t->n) return fail(HODOR_ERR_INVALID_ARG, "query index out of range");  // reference: assert!
    const int len = (int)log2u(t->n);
    cudaStream_t st = c->stream;
    for (uint32_t done = 0; done < count; done += TREE_QUERY_SLOTS) {
        const uint32_t m = count - done < TREE_QUERY_SLOTS ? count - done : TREE_QUERY_SLOTS;
        uint64_t* d_idx = (uint64_t*)(c->small + 96);  // 64 indices = 512 B of the 4 KiB scalar scratch
        HODOR_CUDA_TRY(cudaMemcpyAsync(d_idx, natural_indices + done, m * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
        int rc = merkle_paths_gather(*c, t->nodes, t->values, t->n, d_idx, m, t->path, st);
        if (rc) return rc;
        for (uint32_t i = 0; i < m; i++) {
            const uint4* slot = t->path + 2 * 66 * (size_t)i;
            if (paths) HODOR_CUDA_TRY(cudaMemcpyAsync(paths + (size_t)(done + i) * 32 * len, slot, (size_t)len * 32, cudaMemcpyDeviceToHost, st));
            if (values) HODOR_CUDA_TRY(cudaMemcpyAsync(values + 4 * (size_t)(done + i), slot + 2 * 64, 32, cudaMemcpyDeviceToHost, st));
        }
        HODOR_CUDA_TRY(cudaStreamSynchronize(st));
    }
    return len;
}
int hodor_cuda_tree_query(const hodor_tree* t, uint64_t natural_index, uint64_t value[4], uint8_t* path) {
    return hodor_cuda_tree_query_batch(t, &natural_index, 1, value, path);
}

// ---- per-kernel timing (bench.py's live roofline) ------------------------------------------------------
int hodor_cuda_profile_begin(void) {
    LOCKED_CTX();
    for (auto& r : c->prof) {
        c->event_pool.push_back(r.start);
        c->event_pool.push_back(r.stop);
    }
    c->prof.clear();
    c->profiling = true;
    return HODOR_OK;
}
int hodor_cuda_profile_end(char* json_out, size_t cap) {
    LOCKED_CTX();
    c->profiling = false;
    HODOR_CUDA_TRY(cudaDeviceSynchronize());
    std::map<std::string, std::pair<uint64_t, double>> agg;  // name -> (count, total ms)
    for (auto& r : c->prof) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, r.start, r.stop) != cudaSuccess) {
            cudaGetLastError();
            continue;
        }
        auto& a = agg[r.name];
        a.first += 1;
        a.second += ms;
    }
    std::string js = "[";
    bool first = true;
    for (auto& kv : agg) {
        char buf[256];
        snprintf(buf, sizeof buf, "%s{\"name\": \"%s\", \"count\": %llu, \"total_ms\": %.6f}", first ? "" : ", ",
                 kv.first.c_str(), (unsigned long long)kv.second.first, kv.second.second);
        js += buf;
        first = false;
    }
    js += "]";
    for (auto& r : c->prof) {
        c->event_pool.push_back(r.start);
        c->event_pool.push_back(r.stop);
    }
    c->prof.clear();
    if (json_out && cap) {
        if (js.size() + 1 > cap) return fail(HODOR_ERR_INVALID_ARG, "profile buffer too small");
        memcpy(json_out, js.c_str(), js.size() + 1);
    }
    return (int)agg.size();
}

// ---- FRI commit chain --------------------------------------------------------------------------------
}  // extern "C"
namespace hodor {
void fri_destroy(hodor_fri_proto* p) {
    if (!p) return;
    Ctx* c = g_ctx;
    if (p->block) c ? c->pool_free(p->block) : (void)cudaFree(p->block);
    if (p->owned_lde) c ? c->pool_free(p->owned_lde) : (void)cudaFree(p->owned_lde);
    delete p;
}
}  // namespace hodor
extern "C" {

hodor_fri_proto* hodor_cuda_fri_commit(const uint64_t* lde, uint64_t n, uint32_t lde_factor, uint32_t out_coeffs,
                                       int lde_on_device, int field_id) {
    Ctx* c = ctx();
    if (!c) return nullptr;
    std::lock_guard<std::mutex> lk(c->mu);
    return fri_commit_impl(c, lde, n, lde_factor, out_coeffs, lde_on_device, field_id);
}
}  // extern "C"

namespace hodor {
// the chain itself; caller holds Ctx::mu
hodor_fri_proto* fri_commit_impl(Ctx* c, const uint64_t* lde, uint64_t n, uint32_t lde_factor, uint32_t out_coeffs,
                                 int lde_on_device, int field_id) {
    const FieldOps* ops = field_ops(field_id);
    if (!ops) return nullptr;
    // the reference's asserts (src/fri/fri_on_values.rs:42-46) and its roots.pop() on an empty vec
    if (!is_pow2(n) || n < 2 || !is_pow2(lde_factor) || !is_pow2(out_coeffs) || lde_factor > n ||
        (n / lde_factor) / out_coeffs == 0) {
        fail(HODOR_ERR_INVALID_ARG, "fri_commit: n, lde_factor, out_coeffs must be powers of two with n/lde_factor >= out_coeffs");
        return nullptr;
    }
    const int steps = (int)log2u((n / lde_factor) / out_coeffs);
    if (steps < 1) {
        fail(HODOR_ERR_INVALID_ARG, "fri_commit: zero folding steps (the reference panics here: roots.pop() on empty)");
        return nullptr;
    }
    if (lde_on_device && !aligned32({lde})) {
        fail(HODOR_ERR_INVALID_ARG, "fri_commit: device element arrays must be 32-byte aligned");
        return nullptr;
    }
    Fe probe;
    if (ops->h_domain_generator(log2u(n), probe)) {
        fail(HODOR_ERR_DOMAIN, "fri_commit: domain larger than the field's 2-adicity");
        return nullptr;
    }
    std::unique_ptr<hodor_fri_proto, void (*)(hodor_fri_proto*)> p(new hodor_fri_proto(), fri_destroy);
    p->field_id = field_id;
    p->n = n;
    p->lde_factor = lde_factor;
    p->out_coeffs = out_coeffs;
    p->steps = steps;
    cudaStream_t st = c->stream;
    if (lde_on_device) {
        p->lde = (const uint4*)lde;
    } else {
        p->owned_lde = (uint4*)c->pool_alloc(n * 32);
        if (!p->owned_lde) return nullptr;
        if (c->h2d(p->owned_lde, lde, n * 32, st)) return nullptr;
        p->lde = p->owned_lde;
    }
    // layout (in 32-byte slots): l0 nodes n | per layer: nodes + values | roots | challenges | final | path scratch
    size_t slots = n;
    for (int i = 0; i < steps; i++) slots += 2 * (n >> (i + 1));
    const size_t last = n >> steps;
    slots += 2 * (size_t)(steps + 1) + last + 80 + (size_t)(steps + 1) * 2 * 66 + 2 * (size_t)(steps + 2);
    p->block = (uint4*)c->pool_alloc(slots * 32);
    if (!p->block) return nullptr;
    uint4* cur = p->block;
    auto take = [&](size_t count) {
        uint4* r = cur;
        cur += 2 * count;
        return r;
    };
    p->nodes.push_back(take(n));
    p->values.push_back(const_cast<uint4*>(p->lde));
    for (int i = 0; i < steps; i++) {
        p->nodes.push_back(take(n >> (i + 1)));
        p->values.push_back(take(n >> (i + 1)));
    }
    p->roots = take(steps + 1);
    p->chal = take(steps + 1);
    p->final_coeffs = take(last);
    p->path = take(80);
    p->proof_scratch = take((size_t)(steps + 1) * 2 * 66);
    p->proof_idx = (uint64_t*)take(2 * (size_t)(steps + 2) / 4 + 1);

    int rc = do_merkle(*c, ops, p->lde, n, p->nodes[0], p->roots, p->chal, st);
    const uint32_t log_n0 = log2u(n);
    for (int i = 0; i < steps && !rc; i++) {
        const size_t m = n >> i;
        if (c->fuse_fold_commit && m / 2 > merkle_tail_width() && m / 2 >= 2048) {  // the fused kernel works on tiles of 2048 outputs
            // fold + the bottom three levels of the new tree in one kernel, the rest of the tree as usual
            rc = ops->fri_fold_commit(*c, p->values[i], m, log_n0, (uint32_t)i, p->chal + 2 * i, p->values[i + 1], p->nodes[i + 1], st);
            size_t w = 0;
            if (!rc) rc = merkle_upper_levels(*c, p->nodes[i + 1], m / 16, &w, st);
            if (!rc) rc = ops->merkle_tail(*c, p->nodes[i + 1], p->nodes[i + 1], (uint32_t)w, false, p->roots + 2 * (i + 1),
                                           p->chal + 2 * (i + 1), st);
            continue;
        }
        rc = ops->fri_fold(*c, p->values[i], m, log_n0, (uint32_t)i, p->chal + 2 * i, p->values[i + 1], 0, 1, 0, st);
        if (rc) break;
        rc = do_merkle(*c, ops, p->values[i + 1], m / 2, p->nodes[i + 1], p->roots + 2 * (i + 1), p->chal + 2 * (i + 1), st);
    }
    if (!rc) rc = do_ifft(*c, ops, p->values[steps], p->final_coeffs, log2u(last), 0, st);
    if (!rc && cudaStreamSynchronize(st) != cudaSuccess) rc = cuda_fail(cudaGetLastError(), "fri_commit sync");
    if (rc) return nullptr;
    return p.release();
}
}  // namespace hodor

extern "C" {

void hodor_cuda_fri_free(hodor_fri_proto* p) {
    Ctx* c = g_ctx;
    if (c) {
        std::lock_guard<std::mutex> lk(c->mu);
        cudaStreamSynchronize(c->stream);
        fri_destroy(p);
    } else {
        fri_destroy(p);
    }
}
int hodor_cuda_fri_num_steps(const hodor_fri_proto* p) { return p ? p->steps : HODOR_ERR_INVALID_ARG; }
uint64_t hodor_cuda_fri_layer_size(const hodor_fri_proto* p, uint32_t layer) {
    if (!p || layer > (uint32_t)p->steps) return 0;
    return p->n >> layer;
}
int hodor_cuda_fri_summary(const hodor_fri_proto* p, uint8_t* roots, uint64_t* challenges, uint64_t* final_coeffs) {
    LOCKED_CTX();
    if (!p) return fail(HODOR_ERR_INVALID_ARG, "null handle");
    cudaStream_t st = c->stream;
    if (roots) HODOR_CUDA_TRY(cudaMemcpyAsync(roots, p->roots, (size_t)(p->steps + 1) * 32, cudaMemcpyDeviceToHost, st));
    if (challenges) HODOR_CUDA_TRY(cudaMemcpyAsync(challenges, p->chal, (size_t)p->steps * 32, cudaMemcpyDeviceToHost, st));
    if (final_coeffs)
        HODOR_CUDA_TRY(cudaMemcpyAsync(final_coeffs, p->final_coeffs, (size_t)p->out_coeffs * 32, cudaMemcpyDeviceToHost, st));
    HODOR_CUDA_TRY(cudaStreamSynchronize(st));
    return HODOR_OK;
}
int hodor_cuda_fri_layer(const hodor_fri_proto* p, uint32_t layer, uint8_t* nodes, uint64_t* values) {
    LOCKED_CTX();
    if (!p || layer > (uint32_t)p->steps) return fail(HODOR_ERR_INVALID_ARG, "bad handle or layer");
    const size_t m = p->n >> layer;
    cudaStream_t st = c->stream;
    if (nodes) HODOR_CUDA_TRY(cudaMemcpyAsync(nodes, p->nodes[layer], m * 32, cudaMemcpyDeviceToHost, st));
    if (values) HODOR_CUDA_TRY(cudaMemcpyAsync(values, p->values[layer], m * 32, cudaMemcpyDeviceToHost, st));
    HODOR_CUDA_TRY(cudaStreamSynchronize(st));
    return HODOR_OK;
}
int hodor_cuda_fri_query(const hodor_fri_proto* p, uint32_t layer, uint64_t natural_index, uint64_t value[4], uint8_t* path) {
    LOCKED_CTX();
    if (!p || layer > (uint32_t)p->steps) return fail(HODOR_ERR_INVALID_ARG, "bad handle or layer");
    const size_t m = p->n >> layer;
    if (natural_index >= m) return fail(HODOR_ERR_INVALID_ARG, "query index out of range");  // reference: assert!
    cudaStream_t st = c->stream;
    int rc = merkle_path_gather(*c, p->nodes[layer], p->values[layer], m, natural_index, p->path, st);
    if (rc) return rc;
    const int len = (int)log2u(m);
    if (path) HODOR_CUDA_TRY(cudaMemcpyAsync(path, p->path, (size_t)len * 32, cudaMemcpyDeviceToHost, st));
    if (value) HODOR_CUDA_TRY(cudaMemcpyAsync(value, p->values[layer] + 2 * natural_index, 32, cudaMemcpyDeviceToHost, st));
    HODOR_CUDA_TRY(cudaStreamSynchronize(st));
    return len;
}
// FRIProofPrototype::produce_proof (src/fri/query_producer.rs:10-53) in one call: for every committed layer the two
// members of the coset of the running index (sorted, src/iop/trivial_coset_combiner.rs:31-43), their values and
// authentication paths; then Domain::index_and_size_for_next_domain (src/domains/mod.rs:56-70).  One gather launch per
// layer, one device-to-host copy and one synchronisation for the whole proof.
//   indices: 2 * (steps + 1) u64;  values: 2 * (steps + 1) * 4 u64;
//   paths:   for layer l (size n >> l) two paths of log2(n >> l) digests each, concatenated in layer order.
// Returns the total number of path digests written.
int hodor_cuda_fri_produce_proof(const hodor_fri_proto* p, uint64_t natural_first_element_index, uint64_t* indices, uint64_t* values,
                                 uint8_t* paths) {
    LOCKED_CTX();
    if (!p) return fail(HODOR_ERR_INVALID_ARG, "null handle");
    if (natural_first_element_index >= p->n) return fail(HODOR_ERR_INVALID_ARG, "query index out of range");
    cudaStream_t st = c->stream;
    const int layers = p->steps + 1;
    std::vector<uint64_t> idx(2 * (size_t)layers);
    uint64_t size = p->n, cur = natural_first_element_index;
    for (int l = 0; l < layers; l++) {
        const uint64_t pair = (cur + size / 2) % size;
        idx[2 * l] = cur < pair ? cur : pair;
        idx[2 * l + 1] = cur < pair ? pair : cur;
        cur = cur < size / 2 ? cur : cur - size / 2;
        size /= 2;
    }
    HODOR_CUDA_TRY(cudaMemcpyAsync(p->proof_idx, idx.data(), idx.size() * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
    for (int l = 0; l < layers; l++) {
        int rc = merkle_paths_gather(*c, p->nodes[l], p->values[l], p->n >> l, p->proof_idx + 2 * l, 2,
                                     p->proof_scratch + 2 * 66 * 2 * (size_t)l, st);
        if (rc) return rc;
    }
    std::vector<uint8_t> host((size_t)layers * 2 * 66 * 32);
    HODOR_CUDA_TRY(cudaMemcpyAsync(host.data(), p->proof_scratch, host.size(), cudaMemcpyDeviceToHost, st));
    HODOR_CUDA_TRY(cudaStreamSynchronize(st));
    size_t out = 0;
    for (int l = 0; l < layers; l++) {
        const int len = (int)log2u(p->n >> l);
        for (int q = 0; q < 2; q++) {
            const uint8_t* slot = host.data() + ((size_t)l * 2 + q) * 66 * 32;
            if (indices) indices[2 * l + q] = idx[2 * l + q];
            if (values) memcpy(values + 4 * (2 * (size_t)l + q), slot + 64 * 32, 32);
            if (paths) memcpy(paths + out * 32, slot, (size_t)len * 32);
            out += (size_t)len;
        }
    }
    return (int)out;
}
int hodor_cuda_fri_commit_host(const uint64_t* lde, uint64_t n, uint32_t lde_factor, uint32_t out_coeffs, uint8_t* l0_nodes,
                               uint8_t** layer_nodes, uint64_t** layer_values, uint64_t* challenges, uint8_t* final_root,
                               uint64_t* final_coeffs, int field_id) {
    hodor_fri_proto* p = hodor_cuda_fri_commit(lde, n, lde_factor, out_coeffs, 0, field_id);
    if (!p) return g_last_code ? g_last_code : HODOR_ERR_CUDA;
    int rc = HODOR_OK;
    std::vector<uint8_t> roots((size_t)(p->steps + 1) * 32);
    rc = hodor_cuda_fri_summary(p, roots.data(), challenges, final_coeffs);
    if (!rc && final_root) memcpy(final_root, roots.data() + (size_t)p->steps * 32, 32);
    if (!rc && l0_nodes) rc = hodor_cuda_fri_layer(p, 0, l0_nodes, nullptr);
    for (int i = 0; i < p->steps && !rc; i++)
        rc = hodor_cuda_fri_layer(p, (uint32_t)i + 1, layer_nodes ? layer_nodes[i] : nullptr, layer_values ? layer_values[i] : nullptr);
    const int steps = p->steps;
    hodor_cuda_fri_free(p);
    return rc ? rc : steps;
}

}  // extern "C"
