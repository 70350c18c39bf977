// Host launchers, templated on the field.  Included by field_bls.cu / field_bn254.cu /
// field_stark.cu, each of which instantiates one FieldOps table.
#pragma once
#include <cstring>
#include <string>

#include "context.h"
#include "fri.cuh"
#include "merkle.cuh"
#include "ntt.cuh"
#include "ntt_commit.cuh"
#include "polyops.cuh"

namespace hodor {

template <class F>
struct Ops {
    using Fld = Field<F>;

    // ------------------------------------------------------------------ host scalars
    static void h_mul(const Fe& a, const Fe& b, Fe& o) { o = Fld().mul(a, b); }
    static void h_add(const Fe& a, const Fe& b, Fe& o) { o = Fld().add(a, b); }
    static void h_sub(const Fe& a, const Fe& b, Fe& o) { o = Fld().sub(a, b); }
    static void h_pow(const Fe& a, uint64_t e, Fe& o) { o = Fld().pow(a, e); }
    static Fe pow2k(Fe a, uint32_t k) {  // a^(2^k)
        Fld f;
        for (uint32_t i = 0; i < k; i++) a = f.mul(a, a);
        return a;
    }
    static int h_inverse(const Fe& a, Fe& o) {  // a^(p-2)
        Fld f;
        if (Fld::eq(a, Fld::zero())) return HODOR_ERR_INVALID_ARG;
        uint32_t e[8];
        for (int i = 0; i < 8; i++) e[i] = F::P(i);
        for (uint32_t borrow = 2, i = 0; borrow && i < 8; i++) {  // e = p - 2
            const uint32_t old = e[i];
            e[i] = old - borrow;
            borrow = old < borrow ? 1u : 0u;
        }
        Fe acc = Fld::one();
        for (int i = 255; i >= 0; i--) {
            acc = f.mul(acc, acc);
            if ((e[i / 32] >> (i % 32)) & 1) acc = f.mul(acc, a);
        }
        o = acc;
        return HODOR_OK;
    }
    static void h_from_repr(const Fe& a, Fe& o) { o = Fld().to_mont(a); }
    static void h_into_repr(const Fe& a, Fe& o) { o = Fld().from_mont(a); }
    static Fe from_u64(uint64_t x) {
        Fe t = Fld::zero();
        t.v[0] = (uint32_t)x;
        t.v[1] = (uint32_t)(x >> 32);
        return Fld().to_mont(t);
    }
    static Fe generator() { return from_u64(F::GENERATOR); }
    static Fe root_of_unity() {  // generator^((p-1) >> S), ff_ce derive
        Fld f;
        uint32_t e[8];
        for (int i = 0; i < 8; i++) e[i] = F::P(i);
        e[0] -= 1;
        for (int s = 0; s < F::S; s++) {
            for (int i = 0; i < 7; i++) e[i] = (e[i] >> 1) | (e[i + 1] << 31);
            e[7] >>= 1;
        }
        const Fe g = generator();
        Fe acc = Fld::one();
        for (int i = 255; i >= 0; i--) {
            acc = f.mul(acc, acc);
            if ((e[i / 32] >> (i % 32)) & 1) acc = f.mul(acc, g);
        }
        return acc;
    }
    static void h_constants(Fe& modulus, Fe& one, Fe& gen, Fe& root, uint32_t& s, uint32_t& num_bits) {
        for (int i = 0; i < 8; i++) modulus.v[i] = F::P(i);
        one = Fld::one();
        gen = generator();
        root = root_of_unity();
        s = F::S;
        num_bits = F::NUM_BITS;
    }
    static int h_domain_generator(uint32_t log_n, Fe& out) {
        if (log_n > (uint32_t)F::S) return HODOR_ERR_DOMAIN;
        static const Fe root = root_of_unity();
        out = pow2k(root, F::S - log_n);
        return HODOR_OK;
    }
    static int h_root_to_challenge(const uint8_t* d, Fe& out) {
        Fe v;
        for (int j = 0; j < 8; j++)
            v.v[7 - j] = ((uint32_t)d[4 * j] << 24) | ((uint32_t)d[4 * j + 1] << 16) | ((uint32_t)d[4 * j + 2] << 8) |
                         (uint32_t)d[4 * j + 3];
        constexpr uint32_t shave = (256 - (F::NUM_BITS - 1)) % 64;
        if (shave >= 32) {
            v.v[7] = 0;
            v.v[6] &= 0xffffffffu >> (shave - 32);
        } else {
            v.v[7] &= 0xffffffffu >> shave;
        }
        Fld f;
        if (!f.is_canonical(v)) return HODOR_ERR_INVALID_ARG;  // reference: expect("in a field") panics
        out = f.to_mont(v);
        return HODOR_OK;
    }

    // ------------------------------------------------------------------ tables
    static std::string key_of(const char* tag, uint32_t a, uint32_t b, const Fe* elems, int n_elems) {
        std::string k(tag);
        k.push_back((char)F::ID);
        k.append((const char*)&a, 4);
        k.append((const char*)&b, 4);
        for (int i = 0; i < n_elems; i++) k.append((const char*)elems[i].v, 32);
        return k;
    }

    // tables of bases[b]^e, e in [0, 2^bits); `scale` (optional) is folded into the lo tables
    static int build_pow_tables(Ctx& c, PowTables& t, const std::vector<Fe>& bases, uint32_t bits, const Fe* scale,
                                cudaStream_t st) {
        const uint32_t nb = (uint32_t)bases.size();
        t.count = nb;
        t.lo_bits = (bits + 1) / 2;
        t.hi_bits = bits - t.lo_bits;
        const size_t lo_n = (size_t)nb << t.lo_bits, hi_n = (size_t)nb << t.hi_bits;
        const size_t hdr = 2 * (size_t)nb + 1;  // Fe slots: lo bases, hi bases, scale
        t.bytes = (hdr + lo_n + hi_n) * sizeof(Fe);
        HODOR_CUDA_TRY(cudaMalloc((void**)&t.block, t.bytes));
        struct Guard {  // an early return below must not leak the block
            PowTables& t;
            bool keep = false;
            ~Guard() {
                if (keep) return;
                cudaFree(t.block);
                t.block = nullptr;
            }
        } guard{t};
        std::vector<Fe> h(hdr);
        for (uint32_t i = 0; i < nb; i++) {
            h[i] = bases[i];
            h[nb + i] = pow2k(bases[i], t.lo_bits);
        }
        h[2 * nb] = scale ? *scale : Fld::one();
        HODOR_CUDA_TRY(cudaMemcpyAsync(t.block, h.data(), hdr * sizeof(Fe), cudaMemcpyHostToDevice, st));
        HODOR_CUDA_TRY(cudaStreamSynchronize(st));  // h is a stack-owned vector
        const Fe* d_hdr = (const Fe*)t.block;
        t.lo = t.block + 2 * hdr;
        t.hi = t.lo + 2 * lo_n;
        {
            const uint32_t cnt = 1u << t.lo_bits;
            dim3 grid((cnt + 255) / 256, nb);
            ProfScope ps(c, st, "pow_table");
            pow_table_kernel<F><<<grid, 256, 0, st>>>(t.lo, d_hdr, scale ? d_hdr + 2 * nb : nullptr, cnt, 0u);
        }
        {
            const uint32_t cnt = 1u << t.hi_bits;
            dim3 grid((cnt + 255) / 256, nb);
            ProfScope ps(c, st, "pow_table");
            pow_table_kernel<F><<<grid, 256, 0, st>>>(t.hi, d_hdr + nb, nullptr, cnt, 0u);
        }
        HODOR_CUDA_TRY(cudaGetLastError());
        HODOR_CUDA_TRY(cudaStreamSynchronize(st));  // tables may be used from other streams later
        guard.keep = true;
        c.table_bytes += t.bytes;
        return HODOR_OK;
    }

    // Tables are small (a few MiB each) but the cache is kept bounded.  Only called at the top of an
    // entry point, before any table pointer has been fetched, and after a device-wide sync.
    static void maybe_evict(Ctx& c) {
        if (c.pow_tables.size() <= 64 && c.ntt_tables.size() <= 32) return;
        cudaDeviceSynchronize();
        for (auto& kv : c.pow_tables) cudaFree(kv.second.block);
        for (auto& kv : c.ntt_tables) {
            cudaFree(kv.second.pw.block);
            cudaFree(kv.second.tw_b_block);
            cudaFree(kv.second.tw_direct_block);
        }
        c.pow_tables.clear();
        c.ntt_tables.clear();
        c.table_bytes = 0;
        for (auto& kv : c.full_tables) cudaFree(kv.second.first);
        c.full_tables.clear();
        c.full_bytes = 0;
    }

    static int get_pow_tables(Ctx& c, const PowTables** out, const std::vector<Fe>& bases, uint32_t bits, const Fe* scale,
                              cudaStream_t st) {
        std::vector<Fe> k(bases);
        if (scale) k.push_back(*scale);
        const std::string key = key_of("pow", bits, scale ? 1 : 0, k.data(), (int)k.size());
        auto it = c.pow_tables.find(key);
        if (it == c.pow_tables.end()) {
            PowTables t;
            int rc = build_pow_tables(c, t, bases, bits, scale, st);
            if (rc) return rc;
            it = c.pow_tables.emplace(key, t).first;
        }
        *out = &it->second;
        return HODOR_OK;
    }

    static int get_ntt_tables(Ctx& c, const NttTables** out, uint32_t log_n, const Fe& omega, cudaStream_t st) {
        const std::string key = key_of("ntt", log_n, 0, &omega, 1);
        auto it = c.ntt_tables.find(key);
        if (it == c.ntt_tables.end()) {
            NttTables t;
            Fe* d_base = nullptr;  // device copy of the per-B bases
            struct Guard {  // an early return below must not leak what was already allocated for `t`
                Ctx& c;
                NttTables& t;
                Fe*& d_base;
                bool keep = false, pw_counted = false;
                ~Guard() {
                    if (d_base) cudaFree(d_base);
                    if (keep) return;
                    if (t.pw.block) {
                        cudaFree(t.pw.block);
                        if (pw_counted) c.table_bytes -= t.pw.bytes;  // build_pow_tables adds them only when it succeeds
                    }
                    if (t.tw_b_block) cudaFree(t.tw_b_block);
                    if (t.tw_direct_block) cudaFree(t.tw_direct_block);
                }
            } guard{c, t, d_base};
            std::vector<Fe> base{omega};
            int rc = build_pow_tables(c, t.pw, base, log_n == 0 ? 1 : log_n, nullptr, st);
            if (rc) return rc;
            guard.pw_counted = true;
            if (log_n >= 4) {
                Fld f;
                const Fe w16 = pow2k(omega, log_n - 4);
                Fe acc = w16;
                for (int k = 0; k < 7; k++) {
                    f.make_pre(acc, t.wr[k].w, t.wr[k].q);
                    acc = f.mul(acc, w16);
                }
            }
            const NttPlan plan = make_plan(log_n);
            if (plan.passes) {
                size_t total = 0;
                bool need[10] = {};
                for (int i = 0; i < plan.passes; i++) need[plan.b[i]] = true;
                for (int b = 6; b <= 9; b++)
                    if (need[b]) total += (size_t)1 << b;
                t.bytes = total * sizeof(FePre);
                HODOR_CUDA_TRY(cudaMalloc((void**)&t.tw_b_block, t.bytes));
                HODOR_CUDA_TRY(cudaMalloc((void**)&d_base, 4 * sizeof(Fe)));
                uint4* cur = t.tw_b_block;
                Fe hb[4];
                int nb = 0;
                for (int b = 6; b <= 9; b++)
                    if (need[b]) hb[nb++] = pow2k(omega, log_n - b);
                HODOR_CUDA_TRY(cudaMemcpyAsync(d_base, hb, nb * sizeof(Fe), cudaMemcpyHostToDevice, st));
                nb = 0;
                for (int b = 6; b <= 9; b++) {
                    if (!need[b]) continue;
                    t.tw_b[b] = cur;
                    const uint32_t cnt = 1u << b;
                    {
                        ProfScope ps(c, st, "pow_table");
                        pow_table_kernel<F, true><<<dim3((cnt + 255) / 256, 1), 256, 0, st>>>(cur, d_base + nb, nullptr, cnt, 0u);
                    }
                    cur += 4 * (size_t)cnt;
                    nb++;
                }
                HODOR_CUDA_TRY(cudaGetLastError());
                HODOR_CUDA_TRY(cudaStreamSynchronize(st));
                // flat inter-pass twiddle tables for every non-last pass whose sub-transform N = 2^(s+B)
                // is at most 2^FLAT_MAX_LOG (for 2^24 = 8+8+8 that is pass 2: 65536 entries, 4 MiB; for the
                // four-pass plans of 2^25..2^28 pass 2 has N = 2^18..2^21, up to 128 MiB).  Pass 1 of a
                // transform >= 2^20 uses the expanded per-element table instead.
                size_t dtotal = 0;
                uint32_t below = log_n;
                for (int i = 0; i + 1 < plan.passes; i++) {
                    const uint32_t k = below;  // s + B of pass i
                    below -= plan.b[i];
                    if (k <= FLAT_MAX_LOG && !(i == 0 && log_n >= 20) && t.tw_direct[k] == nullptr) {
                        t.tw_direct[k] = (uint4*)1;  // mark
                        dtotal += (size_t)1 << k;
                    }
                }
                if (dtotal) {
                    HODOR_CUDA_TRY(cudaMalloc((void**)&t.tw_direct_block, dtotal * sizeof(FePre)));
                    uint4* dcur = t.tw_direct_block;
                    for (uint32_t k = 0; k <= FLAT_MAX_LOG; k++) {
                        if (t.tw_direct[k] == nullptr) continue;
                        const Fe base = pow2k(omega, log_n - k);
                        HODOR_CUDA_TRY(cudaMemcpyAsync(d_base, &base, sizeof(Fe), cudaMemcpyHostToDevice, st));
                        HODOR_CUDA_TRY(cudaStreamSynchronize(st));  // `base` is a stack temporary
                        t.tw_direct[k] = dcur;
                        const uint32_t cnt = 1u << k;
                        {
                            ProfScope ps(c, st, "pow_table");
                            pow_table_kernel<F, true><<<dim3((cnt + 255) / 256, 1), 256, 0, st>>>(dcur, d_base, nullptr, cnt, 0u);
                        }
                        HODOR_CUDA_TRY(cudaStreamSynchronize(st));
                        dcur += 4 * (size_t)cnt;
                    }
                    HODOR_CUDA_TRY(cudaGetLastError());
                    t.bytes += dtotal * sizeof(FePre);
                }
                c.table_bytes += t.bytes;
            }
            guard.keep = true;
            it = c.ntt_tables.emplace(key, t).first;
        }
        *out = &it->second;
        return HODOR_OK;
    }

    // One entry per element instead of hi * lo, in the fixed-operand form of Field::mul_pre (64 B per
    // entry): trades HBM capacity and (abundant) bandwidth for one Montgomery multiplication per
    // element and a cheaper remaining one -- the path is multiplier bound, not HBM bound.  Returns
    // null (callers fall back to the two-level tables) when the budget would be exceeded.
    static const uint4* get_full_table(Ctx& c, const std::string& key, const TwoLevel& src, uint32_t stride_lo,
                                       uint32_t stride_hi, size_t n, uint32_t count, int boundary_s, cudaStream_t st) {
        auto it = c.full_tables.find(key);
        if (it != c.full_tables.end()) return it->second.first;
        const size_t bytes = n * count * sizeof(FePre);
        if (c.full_budget == 0 || c.full_bytes + bytes > c.full_budget) return nullptr;
        uint4* d = nullptr;
        if (cudaMalloc((void**)&d, bytes) != cudaSuccess) {
            cudaGetLastError();
            return nullptr;
        }
        const unsigned gx = (unsigned)((n + 255) / 256 > 148 * 32 ? 148 * 32 : (n + 255) / 256);
        {
            ProfScope ps(c, st, "expand_table");
            expand_table_kernel<F><<<dim3(gx, count), 256, 0, st>>>(d, src, stride_lo, stride_hi, n, boundary_s, 0u);
        }
        if (cudaGetLastError() != cudaSuccess || cudaStreamSynchronize(st) != cudaSuccess) {
            cudaGetLastError();
            cudaFree(d);
            return nullptr;
        }
        c.full_tables.emplace(key, std::make_pair(d, bytes));
        c.full_bytes += bytes;
        return d;
    }

    // ------------------------------------------------------------------ transforms
    template <int B, bool SCALE_IN, bool LAST>
    static int launch_pass(Ctx& c, const NttPass& p, dim3 grid, cudaStream_t st) {
        auto kern = ntt_pass_kernel<F, B, SCALE_IN, LAST>;
        constexpr size_t smem = (size_t)256 << B;
        if (!c.configured_kernels.count((const void*)kern)) {  // per context: survives neither shutdown nor a re-init elsewhere
            HODOR_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            HODOR_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout,
                                                cudaSharedmemCarveoutMaxShared));
            c.configured_kernels.insert((const void*)kern);
        }
        {
            ProfScope ps(c, st, LAST ? "ntt_pass_last" : (SCALE_IN ? "ntt_pass_first_scaled" : "ntt_pass"));
            kern<<<grid, PassOccupancy<B>::THREADS, smem, st>>>(p);
        }
        HODOR_CUDA_TRY(cudaGetLastError());
        return HODOR_OK;
    }
    // last pass + bottom three levels of the tree over its output (ntt_commit.cuh)
    template <int B>
    static int launch_last_commit(Ctx& c, const NttPass& p, uint4* nodes, dim3 grid, cudaStream_t st) {
        auto kern = ntt_last_commit_kernel<F, B>;
        constexpr size_t smem = (size_t)256 << B;
        if (!c.configured_kernels.count((const void*)kern)) {
            HODOR_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            HODOR_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout,
                                                cudaSharedmemCarveoutMaxShared));
            c.configured_kernels.insert((const void*)kern);
        }
        {
            ProfScope ps(c, st, "ntt_pass_last_commit");
            kern<<<grid, PassOccupancy<B>::THREADS, smem, st>>>(p, nodes, c.key);
        }
        HODOR_CUDA_TRY(cudaGetLastError());
        return HODOR_OK;
    }

    template <bool SCALE_IN, bool LAST>
    static int launch_pass_b(Ctx& c, int b, const NttPass& p, dim3 grid, cudaStream_t st) {
        switch (b) {
            case 6: return launch_pass<6, SCALE_IN, LAST>(c, p, grid, st);
            case 7: return launch_pass<7, SCALE_IN, LAST>(c, p, grid, st);
            case 8: return launch_pass<8, SCALE_IN, LAST>(c, p, grid, st);
            case 9: return launch_pass<9, SCALE_IN, LAST>(c, p, grid, st);
        }
        return fail(HODOR_ERR_INVALID_ARG, "internal: bad pass width");
    }

    static int ntt(Ctx& c, const uint4* in, uint4* out, uint32_t log_n, uint32_t log_l, const Fe& omega,
                   const Fe* shift0, const Fe* step, int out_mode, const Fe* out_g, cudaStream_t st) {
        Fld f;
        if (log_n > 32 || log_n + log_l > 34) return fail(HODOR_ERR_INVALID_ARG, "transform too large");
        maybe_evict(c);
        // omega must be a primitive 2^log_n-th root: omega^(n/2) == -1  (n == 1: omega == 1)
        {
            const Fe minus_one = f.neg(Fld::one());
            const Fe h = log_n ? pow2k(omega, log_n - 1) : omega;
            if (!Fld::eq(h, log_n ? minus_one : Fld::one()))
                return fail(HODOR_ERR_NOT_A_ROOT, "omega is not a primitive 2^log_n-th root of unity");
        }
        const size_t n = (size_t)1 << log_n;
        const uint32_t L = 1u << log_l;
        const NttTables* tw = nullptr;
        int rc = get_ntt_tables(c, &tw, log_n, omega, st);
        if (rc) return rc;

        const PowTables* coset = nullptr;
        if (shift0 != nullptr) {
            std::vector<Fe> bases(L);
            Fe cur = *shift0;
            for (uint32_t i = 0; i < L; i++) {
                bases[i] = cur;
                if (step) cur = f.mul(cur, *step);
            }
            rc = get_pow_tables(c, &coset, bases, log_n == 0 ? 1 : log_n, nullptr, st);
            if (rc) return rc;
        } else if (log_l != 0) {
            return fail(HODOR_ERR_INVALID_ARG, "internal: cosets without shifts");
        }

        uint32_t flags = 0;
        Fe out_const = Fld::zero();
        const PowTables* out_pow = nullptr;
        if (out_mode) {
            Fe ninv;
            h_inverse(from_u64((uint64_t)n), ninv);
            if (out_mode == 1) {
                flags |= PASS_OUT_CONST;
                out_const = ninv;
            } else {
                std::vector<Fe> bases{*out_g};
                rc = get_pow_tables(c, &out_pow, bases, log_n == 0 ? 1 : log_n, out_mode == 3 ? nullptr : &ninv, st);
                if (rc) return rc;
                flags |= PASS_OUT_POW;
            }
        }

        const NttPlan plan = make_plan(log_n);
        if (plan.passes == 0) {
            SmallNtt p{};
            p.in = in;
            p.out = out;
            p.log_n = log_n;
            p.log_l = log_l;
            p.flags = flags;
            p.zero = 0;
            p.out_const = out_const;
            if (out_pow) p.out_pow = out_pow->two_level();
            if (coset) {
                p.coset = coset->two_level();
                p.coset_stride_lo = coset->stride_lo();
                p.coset_stride_hi = coset->stride_hi();
            }
            // the single-block kernel wants omega^e directly, e < n/2: a dedicated flat table
            const PowTables* flat = nullptr;
            {
                const std::string key = key_of("flat", log_n, 0, &omega, 1);
                auto it = c.pow_tables.find(key);
                if (it == c.pow_tables.end()) {
                    PowTables t;
                    t.count = 1;
                    t.lo_bits = log_n;
                    t.bytes = (n + 1) * sizeof(Fe);
                    HODOR_CUDA_TRY(cudaMalloc((void**)&t.block, t.bytes));
                    cudaError_t e = cudaMemcpyAsync(t.block, &omega, sizeof(Fe), cudaMemcpyHostToDevice, st);
                    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
                    t.lo = t.block + 2;
                    if (e == cudaSuccess) {
                        ProfScope ps(c, st, "pow_table");
                        pow_table_kernel<F><<<dim3((unsigned)((n + 255) / 256), 1), 256, 0, st>>>(
                            t.lo, (const Fe*)t.block, nullptr, (uint32_t)n, 0u);
                    }
                    if (e == cudaSuccess) e = cudaGetLastError();
                    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
                    if (e != cudaSuccess) {
                        cudaFree(t.block);
                        return cuda_fail(e, "flat twiddle table");
                    }
                    c.table_bytes += t.bytes;
                    it = c.pow_tables.emplace(key, t).first;
                }
                flat = &it->second;
            }
            p.tw = flat->lo;
            auto kern = ntt_small_kernel<F>;
            const size_t smem = n * sizeof(Fe);
            if (!c.configured_kernels.count((const void*)kern)) {
                HODOR_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 2048 * 32));
                c.configured_kernels.insert((const void*)kern);
            }
            unsigned threads = (unsigned)(n / 2 < 32 ? 32 : (n / 2 > 1024 ? 1024 : n / 2));
            {
                ProfScope ps(c, st, "ntt_small");
                kern<<<L, threads, smem, st>>>(p);
            }
            HODOR_CUDA_TRY(cudaGetLastError());
            return HODOR_OK;
        }

        // ---- multi-pass ----
        rc = c.ws_acquire(n * L * sizeof(Fe), st);
        if (rc) return rc;
        uint4* work = (uint4*)c.ws;
        NttPass p{};
        p.log_n = log_n;
        p.log_l = log_l;
        p.b1 = plan.b[0];
        p.zero = 0;
        p.tw = tw->pw.two_level();
        f.make_pre(out_const, p.out_const.w, p.out_const.q);
        if (out_pow) p.out_pow = out_pow->two_level();
        if (coset) {
            p.coset = coset->two_level();
            p.coset_stride_lo = coset->stride_lo();
            p.coset_stride_hi = coset->stride_hi();
        }
        for (int k = 0; k < 7; k++) p.wr[k] = tw->wr[k];
        if (c.peer_store.on) {  // set by hodor_cuda_ntt_sharded around step A (sharded.cu)
            if (log_l != 0) return fail(HODOR_ERR_INVALID_ARG, "internal: peer stores with cosets");
            // 2: no fence inside the kernel.  The kernel boundary orders the stores before the barrier collective that
            // follows on this stream, and that collective's own system-scope release / acquire (cumulative) orders
            // them before every peer's step B.  1 (HODOR_PEER_FENCE=1): every thread additionally fences at system
            // scope before it exits -- measured +1.1 ms on the 6.9 ms last pass of 2^28 over 2 GPUs, same result bits.
            static const bool fence = getenv("HODOR_PEER_FENCE") && atoi(getenv("HODOR_PEER_FENCE")) == 1;
            p.peer_on = fence ? 1 : 2;
            p.peer_chunk_log = c.peer_store.chunk_log;
            p.peer_rank = c.peer_store.rank;
            for (int k = 0; k < 16; k++) p.peer[k] = c.peer_store.base[k];
        }
        // expanded tables for the transforms where they pay (>= 2^20): pass-1 inter-pass twiddles and
        // the per-coset scaling powers, one entry per element
        const uint4* tw_full = nullptr;
        const uint4* coset_full = nullptr;
        if (log_n >= 20) {
            const uint32_t s0 = log_n - plan.b[0];
            // make room once, before any pointer is handed out: drop every expanded table if this
            // call's tables would not fit beside the cached ones
            const std::string bkey = key_of("bfull", log_n, s0, &omega, 1);
            std::string ckey;
            if (coset) {
                std::vector<Fe> kb{*shift0};
                if (step) kb.push_back(*step);
                ckey = key_of("cfull", log_n, log_l, kb.data(), (int)kb.size());
            }
            std::string okey;
            if (out_pow) {
                std::vector<Fe> kb{*out_g};
                okey = key_of("ofull", log_n, (uint32_t)out_mode, kb.data(), 1);
            }
            const size_t want = (c.full_tables.count(bkey) ? 0 : n * sizeof(FePre)) +
                                (coset && !c.full_tables.count(ckey) ? n * L * sizeof(FePre) : 0) +
                                (out_pow && !c.full_tables.count(okey) ? n * sizeof(FePre) : 0);
            if (c.full_bytes + want > c.full_budget && !c.full_tables.empty()) {
                cudaDeviceSynchronize();
                for (auto& kv : c.full_tables) cudaFree(kv.second.first);
                c.full_tables.clear();
                c.full_bytes = 0;
            }
            tw_full = get_full_table(c, bkey, tw->pw.two_level(), 0, 0, n, 1, (int)s0, st);
            if (coset)
                coset_full = get_full_table(c, ckey, coset->two_level(), coset->stride_lo(), coset->stride_hi(), n, L, -1, st);
            // output scaling out_g^k (icoset_fft, step A of the sharded NTT) as one streamed fixed-operand multiply
            // instead of two Montgomery multiplies.  Fetched here, after the make-room step above: no pointer handed
            // out by get_full_table may be fetched before it.
            if (out_pow) p.out_pow_full = get_full_table(c, okey, out_pow->two_level(), 0, 0, n, 1, -1, st);
        }
        p.coset_full = coset_full;
        uint32_t below = log_n;
        for (int i = 0; i < plan.passes; i++) {
            const int b = plan.b[i];
            below -= b;
            const bool last = (i == plan.passes - 1);
            p.s = below;
            p.tw_b = tw->tw_b[b];
            p.tw_full = (i == 0 && !last) ? tw_full : nullptr;
            p.tw_direct = (!last && below + b <= FLAT_MAX_LOG) ? tw->tw_direct[below + b] : nullptr;
            p.tw_shift = log_n - below - b;
            p.flags = last ? flags : 0;
            if (!last) {
                p.in = (i == 0) ? in : work;
                p.out = work;
                dim3 grid((unsigned)(((size_t)L << log_n) >> (b + 3)), 1);  // tile-major, coset fastest
                if (i == 0 && coset) rc = launch_pass_b<true, false>(c, b, p, grid, st);
                else rc = launch_pass_b<false, false>(c, b, p, grid, st);
            } else {
                p.in = work;
                p.out = out;
                p.mid0 = plan.passes >= 3 ? plan.b[1] : 0;
                p.mid1 = plan.passes >= 4 ? plan.b[2] : 0;
                if (plan.passes == 3) p.mid1 = 0;
                dim3 grid((unsigned)(((size_t)L << log_n) >> (b + 3)), 1);
                if (c.fuse_commit.nodes != nullptr && flags == 0 && !p.peer_on && c.fuse_last_commit > 0 && b >= 9 - c.fuse_last_commit && b <= 8) {
                    switch (b) {
                        case 6: rc = launch_last_commit<6>(c, p, c.fuse_commit.nodes, grid, st); break;
                        case 7: rc = launch_last_commit<7>(c, p, c.fuse_commit.nodes, grid, st); break;
                        default: rc = launch_last_commit<8>(c, p, c.fuse_commit.nodes, grid, st); break;
                    }
                    c.fuse_commit.done = (rc == HODOR_OK);
                } else {
                    rc = launch_pass_b<false, true>(c, b, p, grid, st);
                }
            }
            if (rc) return rc;
        }
        return c.ws_release(st);
    }

    static int scale_pow(Ctx& c, uint4* a, size_t n, const Fe& g, cudaStream_t st) {
        maybe_evict(c);
        uint32_t bits = 1;
        while (((size_t)1 << bits) < n) bits++;
        const PowTables* t = nullptr;
        std::vector<Fe> bases{g};
        int rc = get_pow_tables(c, &t, bases, bits, nullptr, st);
        if (rc) return rc;
        const unsigned grid = (unsigned)((n + 255) / 256 > 148 * 16 ? 148 * 16 : (n + 255) / 256);
        {
            ProfScope ps(c, st, "scale_pow");
            scale_pow_kernel<F><<<grid ? grid : 1, 256, 0, st>>>(a, n, t->two_level(), 0u);
        }
        HODOR_CUDA_TRY(cudaGetLastError());
        return HODOR_OK;
    }

    // out[j] = map(first * ratio^j), j < n (coset_map_kernel); h_roots / h_inv_van are host arrays (CM_DIVISOR), uploaded
    // into the event-guarded workspace
    static int coset_map(Ctx& c, int mode, uint4* out, size_t n, const Fe& ratio, const Fe& first, const Fe* cst,
                         const Fe* h_roots, uint32_t num_roots, const Fe* h_inv_van, uint32_t van_len, cudaStream_t st) {
        if (n == 0) return HODOR_OK;
        maybe_evict(c);
        uint32_t bits = 1;
        while (((size_t)1 << bits) < n) bits++;
        const PowTables* t = nullptr;
        std::vector<Fe> bases{ratio};
        int rc = get_pow_tables(c, &t, bases, bits, nullptr, st);
        if (rc) return rc;
        const uint4 *d_roots = nullptr, *d_van = nullptr;
        const bool divisor = mode == CM_DIVISOR;
        if (divisor) {
            if (van_len == 0 || (van_len & (van_len - 1)) || h_inv_van == nullptr || (num_roots && h_roots == nullptr))
                return fail(HODOR_ERR_INVALID_ARG, "coset_map: bad divisor tables");
            rc = c.ws_acquire(((size_t)num_roots + van_len) * sizeof(Fe), st);
            if (rc) return rc;
            Fe* w = (Fe*)c.ws;
            // pageable sources: the runtime stages them before cudaMemcpyAsync returns
            if (num_roots) HODOR_CUDA_TRY(cudaMemcpyAsync(w, h_roots, (size_t)num_roots * sizeof(Fe), cudaMemcpyHostToDevice, st));
            HODOR_CUDA_TRY(cudaMemcpyAsync(w + num_roots, h_inv_van, (size_t)van_len * sizeof(Fe), cudaMemcpyHostToDevice, st));
            d_roots = (const uint4*)w;
            d_van = (const uint4*)(w + num_roots);
        }
        const unsigned grid = (unsigned)((n + 255) / 256 > 148 * 16 ? 148 * 16 : (n + 255) / 256);
        {
            ProfScope ps(c, st, "coset_map");
            coset_map_kernel<F><<<grid ? grid : 1, 256, 0, st>>>(mode, out, n, t->two_level(), first, cst ? *cst : Fld::zero(), d_roots,
                                                               num_roots, d_van, divisor ? van_len - 1 : 0u, 0u);
        }
        HODOR_CUDA_TRY(cudaGetLastError());
        return divisor ? c.ws_release(st) : HODOR_OK;
    }

    static int elementwise(Ctx& c, int op, const uint4* a, const uint4* b, uint4* out, size_t n, const Fe* scalar,
                           uint64_t exp, cudaStream_t st) {
        if (op < 0 || op >= EW_NUM_OPS) return fail(HODOR_ERR_INVALID_ARG, "unknown elementwise op");
        const bool needs_b = op <= EW_ADD_SCALED, needs_scalar = op == EW_ADD_SCALED || op == EW_ADD_CONST;
        if (n && ((needs_b && b == nullptr) || (needs_scalar && scalar == nullptr)))
            return fail(HODOR_ERR_INVALID_ARG, "elementwise: missing operand for this op");
        const unsigned grid = (unsigned)((n + 255) / 256 > 148 * 16 ? 148 * 16 : (n + 255) / 256);
        {
            ProfScope ps(c, st, "elementwise");
            elementwise_kernel<F><<<grid ? grid : 1, 256, 0, st>>>(op, a, b, out, n, scalar ? *scalar : Fld::zero(), exp, 0u);
        }
        HODOR_CUDA_TRY(cudaGetLastError());
        return HODOR_OK;
    }

    // ------------------------------------------------------------------ batch inversion, evaluation
    // a[i] <- a[i]^-1 in place; *d_status (device int) = 1 and `a` untouched when some a[i] == 0
    static int batch_inversion(Ctx& c, uint4* a, size_t n, int* d_status, cudaStream_t st) {
        if (n == 0) {
            HODOR_CUDA_TRY(cudaMemsetAsync(d_status, 0, sizeof(int), st));
            return HODOR_OK;
        }
        // level l: m[l] inputs -> M[l] = ceil(m[l] / K) totals, which are level l+1's inputs
        std::vector<size_t> m{n};
        while (m.back() > 1) m.push_back((m.back() + BINV_K - 1) / BINV_K);
        const int levels = (int)m.size() - 1;  // m[levels] == 1
        size_t scratch = 0;
        for (int l = 0; l < levels; l++) scratch += m[l] + m[l + 1];  // running products + totals
        int rc = c.ws_acquire((scratch + 1) * sizeof(Fe), st);
        if (rc) return rc;
        uint4* base = (uint4*)c.ws;
        std::vector<uint4*> pre(levels), tot(levels + 1);
        size_t off = 0;
        for (int l = 0; l < levels; l++) {
            pre[l] = base + 2 * off;
            off += m[l];
            tot[l] = base + 2 * off;
            off += m[l + 1];
        }
        auto input = [&](int l) -> uint4* { return l == 0 ? a : tot[l - 1]; };
        for (int l = 0; l < levels; l++) {
            const size_t M = m[l + 1];
            ProfScope ps(c, st, "batch_inv_up");
            binv_up_kernel<F><<<(unsigned)((M + 255) / 256), 256, 0, st>>>(input(l), pre[l], tot[l], m[l], M, 0u);
        }
        {
            ProfScope ps(c, st, "batch_inv_top");
            binv_top_kernel<F><<<1, 32, 0, st>>>(input(levels), d_status, 0u);
        }
        for (int l = levels - 1; l >= 0; l--) {
            const size_t M = m[l + 1];
            ProfScope ps(c, st, "batch_inv_down");
            binv_down_kernel<F><<<(unsigned)((M + 255) / 256), 256, 0, st>>>(input(l), pre[l], tot[l], input(l), m[l], M,
                                                                             d_status, 0u);
        }
        HODOR_CUDA_TRY(cudaGetLastError());
        return c.ws_release(st);
    }

    static int selftest_mul_pre(Ctx& c, unsigned long long* d_mismatch, cudaStream_t st) {
        HODOR_CUDA_TRY(cudaMemsetAsync(d_mismatch, 0, sizeof(unsigned long long), st));
        ProfScope ps(c, st, "selftest_mul_pre");
        selftest_mul_pre_kernel<F><<<148 * 2, 256, 0, st>>>(d_mismatch, 0u);
        HODOR_CUDA_TRY(cudaGetLastError());
        return HODOR_OK;
    }

    // d_out[0] <- sum_j a[j] * g^j
    static int evaluate_at(Ctx& c, const uint4* a, size_t n, const Fe& g, uint4* d_out, cudaStream_t st) {
        Fld f;
        if (n == 0) {
            HODOR_CUDA_TRY(cudaMemsetAsync(d_out, 0, sizeof(Fe), st));
            return HODOR_OK;
        }
        const uint32_t K = n >= ((size_t)1 << 22) ? 128u : (n >= ((size_t)1 << 16) ? 32u : 8u);
        const size_t M = (n + K - 1) / K;
        const unsigned blocks = (unsigned)((M + 255) / 256);
        int rc = c.ws_acquire((size_t)blocks * sizeof(Fe), st);
        if (rc) return rc;
        FePre gM;
        f.make_pre(f.pow(g, (uint64_t)M), gM.w, gM.q);
        {
            ProfScope ps(c, st, "evaluate_at");
            eval_partial_kernel<F><<<blocks, 256, 0, st>>>(a, n, M, K, g, gM, (uint4*)c.ws, 0u);
        }
        {
            ProfScope ps(c, st, "evaluate_at_final");
            eval_final_kernel<F><<<1, 256, 0, st>>>((const uint4*)c.ws, blocks, d_out, 0u);
        }
        HODOR_CUDA_TRY(cudaGetLastError());
        return c.ws_release(st);
    }

    // ------------------------------------------------------------------ Merkle tail, FRI fold
    static int merkle_tail(Ctx& c, const uint4* in, uint4* nodes, uint32_t w_in, bool leaf, uint4* root, uint4* chal,
                           cudaStream_t st) {
        unsigned threads = w_in / 2 < 32 ? 32 : (w_in / 2 > 1024 ? 1024 : w_in / 2);
        {
            ProfScope ps(c, st, "merkle_tail");
            if (leaf) merkle_tail_kernel<F, true><<<1, threads, 0, st>>>(in, nodes, w_in, c.key, root, chal, 0u);
            else merkle_tail_kernel<F, false><<<1, threads, 0, st>>>(in, nodes, w_in, c.key, root, chal, 0u);
        }
        HODOR_CUDA_TRY(cudaGetLastError());
        return HODOR_OK;
    }

    // omega_N^-e tables of the FRI domain N = 2^log_n0 (src/fri/fri_on_values.rs:24-40): the two-level table, and for
    // the domains where the chain is long enough to pay for it the flat fixed-operand table of N/2 entries (32 * N
    // bytes; cached and budgeted like the other expanded tables, two-level fallback otherwise)
    static int fri_tables(Ctx& c, uint32_t log_n0, const PowTables** t, const uint4** flat, cudaStream_t st) {
        maybe_evict(c);
        auto& inv_cache = c.inv_cache;  // a host inversion is ~400 host multiplies
        Fe omega, omega_inv;
        int rc = h_domain_generator(log_n0, omega);
        if (rc) return fail(rc, "FRI domain larger than the field's 2-adicity");
        const uint64_t inv_key = ((uint64_t)F::ID << 32) | log_n0;
        auto cached = inv_cache.find(inv_key);
        if (cached == inv_cache.end()) {
            h_inverse(omega, omega_inv);
            inv_cache.emplace(inv_key, omega_inv);
        } else {
            omega_inv = cached->second;
        }
        std::vector<Fe> bases{omega_inv};
        rc = get_pow_tables(c, t, bases, log_n0 > 1 ? log_n0 - 1 : 1, nullptr, st);
        if (rc) return rc;
        *flat = nullptr;
        if (log_n0 >= 16)
            *flat = get_full_table(c, key_of("ffull", log_n0, 0, &omega_inv, 1), (*t)->two_level(), 0, 0, (size_t)1 << (log_n0 - 1),
                                   1, -1, st);
        return HODOR_OK;
    }

    static int fri_fold(Ctx& c, const uint4* in, size_t n, uint32_t log_n0, uint32_t layer, const uint4* chal, uint4* out,
                        uint64_t idx_offset, uint64_t idx_stride, uint32_t blk_log, cudaStream_t st) {
        const PowTables* t = nullptr;
        const uint4* flat = nullptr;
        int rc = fri_tables(c, log_n0, &t, &flat, st);
        if (rc) return rc;
        const size_t half = n / 2;
        const unsigned grid = (unsigned)((half + 255) / 256 > 148 * 16 ? 148 * 16 : (half + 255) / 256);
        {
            ProfScope ps(c, st, "fri_fold");
            if (flat)
                fri_fold_kernel<F, true><<<grid ? grid : 1, 256, 0, st>>>(in, out, half, t->two_level(), flat, layer, chal,
                                                                           idx_offset, idx_stride, blk_log, 0u);
            else
                fri_fold_kernel<F, false><<<grid ? grid : 1, 256, 0, st>>>(in, out, half, t->two_level(), nullptr, layer,
                                                                            chal, idx_offset, idx_stride, blk_log, 0u);
        }
        HODOR_CUDA_TRY(cudaGetLastError());
        return HODOR_OK;
    }

    static int fri_fold_commit(Ctx& c, const uint4* in, size_t n, uint32_t log_n0, uint32_t layer, const uint4* chal, uint4* out,
                               uint4* nodes, cudaStream_t st) {
        const PowTables* t = nullptr;
        const uint4* flat = nullptr;
        int rc = fri_tables(c, log_n0, &t, &flat, st);
        if (rc) return rc;
        const size_t half = n / 2;
        if (half < FOLD_COMMIT_TILE || half % FOLD_COMMIT_TILE) return fail(HODOR_ERR_INVALID_ARG, "internal: fold_commit needs a multiple of 2048 outputs");
        const unsigned g = (unsigned)(half / FOLD_COMMIT_TILE);
        constexpr size_t smem = (size_t)FOLD_COMMIT_TILE * sizeof(Fe);
        const void* kern = flat ? (const void*)fri_fold_commit_kernel<F, true> : (const void*)fri_fold_commit_kernel<F, false>;
        if (!c.configured_kernels.count(kern)) {
            HODOR_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            HODOR_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
            c.configured_kernels.insert(kern);
        }
        {
            ProfScope ps(c, st, "fri_fold_commit");
            if (flat)
                fri_fold_commit_kernel<F, true><<<g, 256, smem, st>>>(in, out, nodes, half, t->two_level(), flat, layer, chal, c.key, 0u);
            else
                fri_fold_commit_kernel<F, false><<<g, 256, smem, st>>>(in, out, nodes, half, t->two_level(), nullptr, layer, chal, c.key,
                                                                       0u);
        }
        HODOR_CUDA_TRY(cudaGetLastError());
        return HODOR_OK;
    }

    static int shard_rows(Ctx& c, const uint4* in, uint4* out, uint32_t log_n, uint32_t log_g, uint32_t rank,
                          const Fe& omega, cudaStream_t st);

    static FieldOps table() {
        FieldOps o;
        o.h_mul = h_mul;
        o.h_add = h_add;
        o.h_sub = h_sub;
        o.h_pow = h_pow;
        o.h_inverse = h_inverse;
        o.h_from_repr = h_from_repr;
        o.h_into_repr = h_into_repr;
        o.h_constants = h_constants;
        o.h_domain_generator = h_domain_generator;
        o.h_root_to_challenge = h_root_to_challenge;
        o.ntt = ntt;
        o.scale_pow = scale_pow;
        o.elementwise = elementwise;
        o.coset_map = coset_map;
        o.batch_inversion = batch_inversion;
        o.evaluate_at = evaluate_at;
        o.selftest_mul_pre = selftest_mul_pre;
        o.merkle_tail = merkle_tail;
        o.fri_fold = fri_fold;
        o.fri_fold_commit = fri_fold_commit;
        o.shard_rows = shard_rows;
        return o;
    }
};

// ---------------------------------------------------------------------------------------------
// Four-step across G GPUs, step B: after the all-to-all rank h holds M[g][k] = B_g[h*m/G + k],
// g < G, k < m/G (m = n/G); it computes, for every k, the G-point DFT over g:
//   out[k2 * (m/G) + k] = sum_g M[g][k] * omega_G^(g*k2)          = A[(h*m/G + k) + m*k2]
// ---------------------------------------------------------------------------------------------
template <class F, int LOGG>
__global__ void __launch_bounds__(256) shard_rows_kernel(const uint4* in, uint4* out, size_t cols,
                                                         const __grid_constant__ NttPass p) {
    const uint32_t oz = threadIdx.x & p.zero;
    const Field<F> fld(oz);
    constexpr int G = 1 << LOGG;
    for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < cols; k += (size_t)gridDim.x * blockDim.x) {
        Fe x[G];
#pragma unroll
        for (int g = 0; g < G; g++) x[g] = ld_fe(in, (size_t)g * cols + k);
        if constexpr (LOGG > 0) dif_inreg<F, LOGG>(fld, x, p, oz);
#pragma unroll
        for (int k2 = 0; k2 < G; k2++) st_fe(out, (size_t)k2 * cols + k, x[bitrev_c(k2, LOGG)]);
    }
}

template <class F>
int Ops<F>::shard_rows(Ctx& c, const uint4* in, uint4* out, uint32_t log_n, uint32_t log_g, uint32_t rank,
                       const Fe& omega, cudaStream_t st) {
    (void)rank;
    if (log_g > 4 || log_n < 2 * log_g) return fail(HODOR_ERR_INVALID_ARG, "shard_rows: need log_g <= 4 and log_n >= 2*log_g");
    Fld f;
    NttPass p{};
    p.zero = 0;
    // dif_inreg expects wr[k-1] = w16^k where w16^(16/R) is the radix-R root; here R = G and the
    // root is omega_G = omega^(n/G): set w16 = omega^(n/16) (exists when log_n >= 4)
    if (log_g > 0) {
        if (log_n < 4) return fail(HODOR_ERR_INVALID_ARG, "shard_rows: log_n < 4");
        const Fe w16 = pow2k(omega, log_n - 4);
        Fe acc = w16;
        for (int k = 0; k < 7; k++) {
            f.make_pre(acc, p.wr[k].w, p.wr[k].q);
            acc = f.mul(acc, w16);
        }
    }
    const size_t cols = ((size_t)1 << log_n) >> (2 * log_g);
    const unsigned grid = (unsigned)((cols + 255) / 256 > 148 * 8 ? 148 * 8 : (cols + 255) / 256);
    ProfScope ps(c, st, "shard_rows");
    switch (log_g) {
        case 0: shard_rows_kernel<F, 0><<<grid ? grid : 1, 256, 0, st>>>(in, out, cols, p); break;
        case 1: shard_rows_kernel<F, 1><<<grid ? grid : 1, 256, 0, st>>>(in, out, cols, p); break;
        case 2: shard_rows_kernel<F, 2><<<grid ? grid : 1, 256, 0, st>>>(in, out, cols, p); break;
        case 3: shard_rows_kernel<F, 3><<<grid ? grid : 1, 256, 0, st>>>(in, out, cols, p); break;
        case 4: shard_rows_kernel<F, 4><<<grid ? grid : 1, 256, 0, st>>>(in, out, cols, p); break;
    }
    HODOR_CUDA_TRY(cudaGetLastError());
    return HODOR_OK;
}

}  // namespace hodor
