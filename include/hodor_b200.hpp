// hodor_b200.hpp -- C++17 host-side mirror of the reference interface for the hot path, on top of
// the C ABI of hodor_b200.h.  Header only; link with -lhodor_b200.
//
// The reference (matter-labs/hodor @ 76fc894) is Rust and its toolchain is not available in this
// environment, so the host side that a Rust caller would write is mirrored here in C++ with the
// same names, argument meaning and error behaviour:
//
//   Domain<F>                         src/domains/mod.rs:14-70
//   Polynomial<F, Coefficients|Values> src/polynomials/mod.rs:26-34, 343-352, 611-638, 773-815
//   Blake2sIopTree<F>, TrivialBlake2sIOP<F>, TrivialBlake2sIopQuery<F>
//                                     src/iop/mod.rs:49-92, src/iop/blake2s_trivial_iop.rs:107-368
//   NaiveFriIop<F>, FRIProofPrototype<F>, FRIProof<F>
//                                     src/fri/mod.rs:26-154, src/fri/fri_on_values.rs:11-159,
//                                     src/fri/query_producer.rs:10-53
//   Comm, DeviceVec<F>                several GPUs, one process per GPU (no counterpart in the reference: its
//                                     parallelism is the threads of one host, src/fft/multicore.rs)
//
// `Result<_, SynthesisError>` becomes a thrown SynthesisError; the reference's assert!/expect panics
// become std::logic_error.  `Worker` is accepted and ignored (the CUDA grid replaces the thread pool).
#pragma once
#include <algorithm>
#include <array>
#include <cstdint>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <tuple>
#include <utility>
#include <vector>

#include "hodor_b200.h"

namespace hodor_b200 {

struct SynthesisError : std::runtime_error {  // src/lib.rs:40-46
    explicit SynthesisError(const std::string& what) : std::runtime_error(what) {}
};
struct CudaError : std::runtime_error {
    explicit CudaError(const std::string& what) : std::runtime_error(what) {}
};

inline int check(int rc) {
    if (rc >= 0) return rc;
    const std::string msg = hodor_cuda_last_error();
    if (rc == HODOR_ERR_DOMAIN || rc == HODOR_ERR_NOT_INVERTIBLE) throw SynthesisError("General error for now: " + msg);
    if (rc == HODOR_ERR_INVALID_ARG) throw std::logic_error(msg);
    throw CudaError(msg);
}
inline void init(int device = 0) { check(hodor_cuda_init(device)); }

struct Worker {};  // src/fft/multicore.rs:17-103

using Digest = std::array<uint8_t, 32>;

// One field element: 4 little-endian u64 limbs, Montgomery form -- `Fr(FrRepr([u64; 4]))`.
template <int FIELD_ID>
struct Fr {
    static constexpr int ID = FIELD_ID;
    uint64_t l[4] = {0, 0, 0, 0};
    bool operator==(const Fr& o) const { return std::memcmp(l, o.l, 32) == 0; }
    bool operator!=(const Fr& o) const { return !(*this == o); }

    static Fr zero() { return Fr{}; }
    static Fr one() {
        Fr r;
        check(hodor_field_constants(ID, nullptr, r.l, nullptr, nullptr, nullptr, nullptr, nullptr));
        return r;
    }
    static Fr multiplicative_generator() {
        Fr r;
        check(hodor_field_constants(ID, nullptr, nullptr, r.l, nullptr, nullptr, nullptr, nullptr));
        return r;
    }
    static Fr root_of_unity() {
        Fr r;
        check(hodor_field_constants(ID, nullptr, nullptr, nullptr, r.l, nullptr, nullptr, nullptr));
        return r;
    }
    static uint32_t S() {
        uint32_t s = 0;
        check(hodor_field_constants(ID, nullptr, nullptr, nullptr, nullptr, &s, nullptr, nullptr));
        return s;
    }
    static Fr from_u64(uint64_t x) {  // PrimeField::from_str / from_repr
        const uint64_t plain[4] = {x, 0, 0, 0};
        Fr r;
        check(hodor_field_from_repr(ID, plain, r.l));
        return r;
    }
    void mul_assign(const Fr& o) { check(hodor_field_mul(ID, l, o.l, l)); }
    void add_assign(const Fr& o) { check(hodor_field_add(ID, l, o.l, l)); }
    void sub_assign(const Fr& o) { check(hodor_field_sub(ID, l, o.l, l)); }
    void square() { mul_assign(*this); }
    void double_() { add_assign(*this); }
    Fr pow(uint64_t e) const {
        Fr r;
        check(hodor_field_pow(ID, l, e, r.l));
        return r;
    }
    // Field::inverse returns Option<Self>: empty on zero
    std::pair<bool, Fr> inverse() const {
        Fr r;
        if (hodor_field_inverse(ID, l, r.l) != HODOR_OK) return {false, Fr{}};
        return {true, r};
    }
};
using Bn256RsFr = Fr<HODOR_FIELD_BLS12_381_FR>;  // the field src/bn256.rs declares
using Bn254Fr = Fr<HODOR_FIELD_BN254_FR>;
using Stark252Fr = Fr<HODOR_FIELD_STARK252>;

template <class F>
struct Domain {
    uint64_t size = 0;
    uint64_t power_of_two = 0;
    F generator;

    static Domain new_for_size(uint64_t size) {  // :21-44
        uint64_t s = 1, k = 0;
        while (s < size) {
            s <<= 1;
            k++;
        }
        Domain d;
        d.size = s;
        d.power_of_two = k;
        check(hodor_domain_generator(F::ID, (uint32_t)k, d.generator.l));  // Err(SynthesisError::Error)
        return d;
    }
    static std::vector<size_t> coset_for_natural_index_and_size(size_t natural_index, size_t domain_size) {
        if (domain_size <= 1 || (domain_size & (domain_size - 1))) throw std::logic_error("assert!(domain_size.is_power_of_two())");
        std::vector<size_t> c{natural_index, (natural_index + domain_size / 2) % domain_size};
        std::sort(c.begin(), c.end());
        return c;
    }
    static std::pair<size_t, size_t> index_and_size_for_next_domain(size_t natural_index, size_t domain_size) {
        if (domain_size <= 1 || (domain_size & (domain_size - 1))) throw std::logic_error("assert!(domain_size.is_power_of_two())");
        const size_t next = domain_size / 2;
        return {natural_index < next ? natural_index : natural_index - next, next};
    }
};

struct Coefficients {};
struct Values {};

template <class F, class Form>
class Polynomial {
  public:
    uint32_t exp = 0;
    F omega, omegainv, geninv, minv;

    static Polynomial from_coeffs(std::vector<F> v) {
        static_assert(std::is_same<Form, Coefficients>::value, "from_coeffs makes Polynomial<F, Coefficients>");
        return Polynomial(std::move(v));
    }
    static Polynomial from_values(std::vector<F> v) {
        static_assert(std::is_same<Form, Values>::value, "from_values makes Polynomial<F, Values>");
        return Polynomial(std::move(v));
    }
    size_t size() const { return coeffs_.size(); }
    const std::vector<F>& as_ref() const { return coeffs_; }
    std::vector<F>& as_mut() { return coeffs_; }
    std::vector<F> into_coeffs() && { return std::move(coeffs_); }
    bool operator==(const Polynomial& o) const { return coeffs_ == o.coeffs_; }

    void distribute_powers(const Worker&, const F& g) {  // src/fft/mod.rs:110-123
        check(hodor_cuda_distribute_powers(raw(), coeffs_.size(), g.l, F::ID));
    }
    // ---- Polynomial<F, Coefficients> -----------------------------------------------------------
    Polynomial<F, Values> fft(const Worker&) && {  // :611-624
        require<Coefficients>();
        check(hodor_cuda_fft(raw(), exp, 0, F::ID));
        return Polynomial<F, Values>::adopt(std::move(coeffs_));
    }
    Polynomial<F, Values> coset_fft(const Worker&) && {  // :626-631
        require<Coefficients>();
        check(hodor_cuda_fft(raw(), exp, 1, F::ID));
        return Polynomial<F, Values>::adopt(std::move(coeffs_));
    }
    Polynomial<F, Values> lde(const Worker&, size_t factor) && { return do_lde(factor, 0); }        // :343-346
    Polynomial<F, Values> coset_lde(const Worker&, size_t factor) && { return do_lde(factor, 1); }  // :348-352
    // ---- Polynomial<F, Values> -----------------------------------------------------------------
    Polynomial<F, Coefficients> ifft(const Worker&) && {  // :773-798
        require<Values>();
        check(hodor_cuda_ifft(raw(), exp, 0, F::ID));
        return Polynomial<F, Coefficients>::adopt(std::move(coeffs_));
    }
    Polynomial<F, Coefficients> icoset_fft(const Worker&) && {  // :800-807
        require<Values>();
        check(hodor_cuda_ifft(raw(), exp, 1, F::ID));
        return Polynomial<F, Coefficients>::adopt(std::move(coeffs_));
    }
    // ---- elementwise (:59-83, :640-683, :744-771, :817-887) --------------------------------------
    void scale(const Worker&, const F& g) { op(HODOR_OP_SCALE, g.l, nullptr, 0, size()); }
    void negate(const Worker&) { op(HODOR_OP_NEGATE, nullptr, nullptr, 0, size()); }
    void add_assign(const Worker&, const Polynomial& o) { op(HODOR_OP_ADD, o.raw_c(), nullptr, 0, checked(o)); }
    void sub_assign(const Worker&, const Polynomial& o) { op(HODOR_OP_SUB, o.raw_c(), nullptr, 0, checked(o)); }
    void add_assign_scaled(const Worker&, const Polynomial& o, const F& s) { op(HODOR_OP_ADD_SCALED, o.raw_c(), s.l, 0, checked(o)); }
    void mul_assign(const Worker&, const Polynomial& o) {
        require<Values>();
        if (o.size() != size()) throw std::logic_error("assert_eq!(self.coeffs.len(), other.coeffs.len())");
        op(HODOR_OP_MUL, o.raw_c(), nullptr, 0, size());
    }
    void add_constant(const Worker&, const F& c) {
        require<Values>();
        op(HODOR_OP_ADD_CONST, nullptr, c.l, 0, size());
    }
    void square(const Worker&) {
        require<Values>();
        op(HODOR_OP_SQUARE, nullptr, nullptr, 0, size());
    }
    void pow(const Worker& w, uint64_t exp) {
        require<Values>();
        if (exp == 2) return square(w);
        op(HODOR_OP_POW, nullptr, nullptr, exp, size());
    }
    void batch_inversion(const Worker&) {  // :889-954; SynthesisError (vector untouched) on a zero value
        require<Values>();
        check(hodor_cuda_batch_inversion(raw(), coeffs_.size(), F::ID));
    }
    F evaluate_at(const Worker&, const F& g) const {  // :685-711
        require<Coefficients>();
        F out;
        check(hodor_cuda_evaluate_at(reinterpret_cast<const uint64_t*>(coeffs_.data()), coeffs_.size(), g.l, out.l, F::ID));
        return out;
    }
    Polynomial clone() const { return Polynomial(coeffs_); }

    static Polynomial adopt(std::vector<F> v) { return Polynomial(std::move(v)); }

  private:
    std::vector<F> coeffs_;

    explicit Polynomial(std::vector<F> v) : coeffs_(std::move(v)) {  // from_values / from_coeffs, :713-742
        const auto dom = Domain<F>::new_for_size(std::max<size_t>(1, coeffs_.size()));
        coeffs_.resize(dom.size, F::zero());
        exp = (uint32_t)dom.power_of_two;
        omega = dom.generator;
        omegainv = omega.inverse().second;
        geninv = F::multiplicative_generator().inverse().second;
        minv = F::from_u64(dom.size).inverse().second;
    }
    uint64_t* raw() { return reinterpret_cast<uint64_t*>(coeffs_.data()); }
    const uint64_t* raw_c() const { return reinterpret_cast<const uint64_t*>(coeffs_.data()); }
    size_t checked(const Polynomial& o) const {
        if (size() < o.size()) throw std::logic_error("assert!(self.coeffs.len() >= other.coeffs.len())");
        return o.size();
    }
    void op(int code, const uint64_t* b, const uint64_t* scalar, uint64_t exp, size_t n) {
        check(hodor_cuda_poly_op(code, raw(), b, scalar, exp, raw(), n, F::ID));
    }
    template <class Need>
    void require() const {
        static_assert(std::is_same<Form, Need>::value, "wrong polynomial form for this transform");
    }
    Polynomial<F, Values> do_lde(size_t factor, int coset) {
        require<Coefficients>();
        if (factor == 0 || (factor & (factor - 1))) throw std::logic_error("assert!(factor.is_power_of_two())");
        uint32_t log_f = 0;
        while (((size_t)1 << log_f) < factor) log_f++;
        (void)Domain<F>::new_for_size(coeffs_.size() * factor);  // same Err as src/polynomials/mod.rs:435
        std::vector<F> out(coeffs_.size() * factor);
        check(hodor_cuda_lde(raw(), exp, log_f, coset, reinterpret_cast<uint64_t*>(out.data()), F::ID));
        return Polynomial<F, Values>::adopt(std::move(out));
    }
};

// `for w in polys { w.lde(&worker, factor) }` (src/prover/mod.rs:73-76) as ONE pipelined call: the PCIe
// copies of neighbouring polynomials overlap the transform of the current one (hodor_cuda_lde_batch).
template <class F>
std::vector<Polynomial<F, Values>> lde_batch(const std::vector<Polynomial<F, Coefficients>>& polys, const Worker&,
                                             size_t factor, bool coset = false) {
    std::vector<Polynomial<F, Values>> result;
    if (polys.empty()) return result;
    if (factor == 0 || (factor & (factor - 1))) throw std::logic_error("assert!(factor.is_power_of_two())");
    const size_t n = polys[0].size();
    uint32_t log_f = 0;
    while (((size_t)1 << log_f) < factor) log_f++;
    (void)Domain<F>::new_for_size(n * factor);
    std::vector<std::vector<F>> outs(polys.size(), std::vector<F>(n * factor));
    std::vector<const uint64_t*> in_ptrs;
    std::vector<uint64_t*> out_ptrs;
    for (size_t i = 0; i < polys.size(); i++) {
        if (polys[i].size() != n) throw std::logic_error("lde_batch: polynomials must have one size");
        in_ptrs.push_back(reinterpret_cast<const uint64_t*>(polys[i].as_ref().data()));
        out_ptrs.push_back(reinterpret_cast<uint64_t*>(outs[i].data()));
    }
    check(hodor_cuda_lde_batch(in_ptrs.data(), out_ptrs.data(), (uint32_t)polys.size(), polys[0].exp, log_f, coset ? 1 : 0,
                               F::ID));
    for (auto& o : outs) result.push_back(Polynomial<F, Values>::adopt(std::move(o)));
    return result;
}

// ---- setup of Prover::new ----------------------------------------------------------------------
// src/precomputations/mod.rs:7-66
template <class F>
struct PrecomputedOmegas {
    std::vector<F> omegas, coset, omegas_inv;
    static PrecomputedOmegas new_for_domain(const Domain<F>& domain, const Worker&) {
        PrecomputedOmegas p;
        p.omegas.resize(domain.size);
        p.coset.resize(domain.size);
        p.omegas_inv.resize(domain.size / 2);
        check(hodor_cuda_precomputed_omegas(reinterpret_cast<uint64_t*>(p.omegas.data()), reinterpret_cast<uint64_t*>(p.coset.data()),
                                            domain.size >= 2 ? reinterpret_cast<uint64_t*>(p.omegas_inv.data()) : nullptr,
                                            (uint32_t)domain.power_of_two, F::ID));
        return p;
    }
};
struct DenseConstraint {  // src/air/mod.rs:30-33
    size_t start_at = 0, span = 1;
};
// src/ali/per_register/mod.rs:60-162 -> (inverse divisor over g * <evaluation domain>, divisor degree)
template <class F>
std::pair<Polynomial<F, Values>, size_t> inverse_divisor_for_dense_constraint_in_coset(const Domain<F>& column_domain,
                                                                                       const Domain<F>& evaluation_domain,
                                                                                       DenseConstraint dense_constraint,
                                                                                       uint64_t num_rows, const Worker&) {
    std::vector<F> out(evaluation_domain.size);
    uint64_t degree = 0;
    check(hodor_cuda_ali_dense_inverse_divisor(reinterpret_cast<uint64_t*>(out.data()), (uint32_t)column_domain.power_of_two,
                                               (uint32_t)evaluation_domain.power_of_two, dense_constraint.start_at,
                                               dense_constraint.span, num_rows, &degree, F::ID));
    return {Polynomial<F, Values>::from_values(std::move(out)), (size_t)degree};
}
// src/ali/per_register/mod.rs:214-227: 1 / (X - omega^row) over the coset of the constraints domain
template <class F>
Polynomial<F, Values> boundary_constraint_inverse_divisor(const Domain<F>& column_domain, const Domain<F>& constraints_domain,
                                                          uint64_t row, const Worker&) {
    std::vector<F> out(constraints_domain.size);
    check(hodor_cuda_ali_boundary_inverse_divisor(reinterpret_cast<uint64_t*>(out.data()), (uint32_t)column_domain.power_of_two,
                                                  (uint32_t)constraints_domain.power_of_two, row, F::ID));
    return Polynomial<F, Values>::from_values(std::move(out));
}

// ---- IOP -------------------------------------------------------------------------------------
template <class F>
struct Blake2sTreeHasher {  // src/iop/blake2s_trivial_iop.rs:63-105
    static Digest hash_leaf(const F& v) {
        Digest d;
        check(hodor_hash_leaf(v.l, d.data()));
        return d;
    }
    static Digest hash_node(const Digest& l, const Digest& r) {
        Digest d;
        check(hodor_hash_node(l.data(), r.data(), d.data()));
        return d;
    }
};

template <class F>
struct TrivialBlake2sIopQuery {  // :343-368
    size_t index = 0;
    F value_;
    std::vector<Digest> path_;
    size_t natural_index() const { return index; }
    size_t tree_index() const { return index; }
    const F& value() const { return value_; }
    const std::vector<Digest>& path() const { return path_; }
};

template <class F>
class Blake2sIopTree {  // :107-280
  public:
    static Blake2sIopTree create(const std::vector<F>& leafs) {
        const size_t n = leafs.size();
        if (n < 2 || (n & (n - 1))) throw std::logic_error("assert!(num_leafs == num_leafs.next_power_of_two())");
        Blake2sIopTree t;
        t.nodes_.resize(n);
        check(hodor_cuda_merkle_build(reinterpret_cast<const uint64_t*>(leafs.data()), n,
                                      reinterpret_cast<uint8_t*>(t.nodes_.data()), F::ID));
        return t;
    }
    uint64_t size() const { return nodes_.size(); }
    Digest get_root() const { return nodes_[1]; }
    static F encode_root_into_challenge(const Digest& root) {
        F c;
        check(hodor_root_to_challenge(root.data(), c.l, F::ID));
        return c;
    }
    F get_challenge_scalar_from_root() const { return encode_root_into_challenge(get_root()); }
    static bool verify(const Digest& root, const F& leaf_value, const std::vector<Digest>& path, size_t tree_index) {
        Digest h = Blake2sTreeHasher<F>::hash_leaf(leaf_value);
        size_t idx = tree_index;
        for (const Digest& el : path) {
            h = (idx & 1) == 0 ? Blake2sTreeHasher<F>::hash_node(h, el) : Blake2sTreeHasher<F>::hash_node(el, h);
            idx >>= 1;
        }
        return h == root;
    }
    std::vector<Digest> get_path(size_t tree_index, const std::vector<F>& leafs_values) const {
        std::vector<Digest> path{Blake2sTreeHasher<F>::hash_leaf(leafs_values[tree_index ^ 1])};
        for (size_t idx = (nodes_.size() + tree_index) >> 1; idx > 1; idx >>= 1) path.push_back(nodes_[idx ^ 1]);
        return path;
    }
    const std::vector<Digest>& nodes() const { return nodes_; }

  private:
    std::vector<Digest> nodes_;
};

template <class F>
class TrivialBlake2sIOP {  // :282-341
  public:
    using Query = TrivialBlake2sIopQuery<F>;
    static TrivialBlake2sIOP create(const std::vector<F>& leafs) { return TrivialBlake2sIOP{Blake2sIopTree<F>::create(leafs)}; }
    Digest get_root() const { return tree.get_root(); }
    F get_challenge_scalar_from_root() const { return tree.get_challenge_scalar_from_root(); }
    static bool verify_query(const Query& q, const Digest& root) {
        return Blake2sIopTree<F>::verify(root, q.value(), q.path(), q.tree_index());
    }
    Query query(size_t natural_index, const std::vector<F>& leafs) const {
        if (natural_index >= tree.size() || natural_index >= leafs.size()) throw std::logic_error("assert!(natural_index < size)");
        return Query{natural_index, leafs[natural_index], tree.get_path(natural_index, leafs)};
    }
    bool operator==(const TrivialBlake2sIOP& o) const { return get_root() == o.get_root(); }
    Blake2sIopTree<F> tree;
};

// An oracle whose leaves and tree stay in HBM behind a `hodor_tree` handle: what Prover::prove keeps per register
// between the commit phase and the query phase (src/prover/mod.rs:73-95, :142-151).  Same surface as
// TrivialBlake2sIOP; `query` extracts value and path on the device.
template <class F>
class CommittedOracle {
  public:
    using Query = TrivialBlake2sIopQuery<F>;
    // IOP::create on host values (copied in once)
    static CommittedOracle create(const std::vector<F>& leafs) {
        const size_t n = leafs.size();
        if (n < 2 || (n & (n - 1))) throw std::logic_error("assert!(num_leafs == num_leafs.next_power_of_two())");
        Digest root{};
        hodor_tree* h = hodor_cuda_tree_commit(reinterpret_cast<const uint64_t*>(leafs.data()), n, 0, root.data(), F::ID);
        if (!h) check(hodor_cuda_last_error_code() < 0 ? hodor_cuda_last_error_code() : HODOR_ERR_CUDA);
        return CommittedOracle(h, root);
    }
    // `let lde = w.lde(&worker, factor)?; I::create(lde.as_ref())` for every polynomial, in one pipelined call
    static std::vector<CommittedOracle> lde_commit_batch(const std::vector<Polynomial<F, Coefficients>>& polys, size_t factor,
                                                         bool coset = false) {
        std::vector<CommittedOracle> out;
        if (polys.empty()) return out;
        if (factor == 0 || (factor & (factor - 1))) throw std::logic_error("assert!(factor.is_power_of_two())");
        uint32_t log_f = 0;
        while (((size_t)1 << log_f) < factor) log_f++;
        std::vector<const uint64_t*> in;
        for (const auto& p : polys) {
            if (p.size() != polys[0].size()) throw std::logic_error("lde_commit_batch: polynomials must have one size");
            in.push_back(reinterpret_cast<const uint64_t*>(p.as_ref().data()));
        }
        std::vector<hodor_tree*> handles(polys.size(), nullptr);
        std::vector<Digest> roots(polys.size());
        check(hodor_cuda_lde_commit_batch(in.data(), (uint32_t)polys.size(), polys[0].exp, log_f, coset ? 1 : 0, 0, handles.data(),
                                          reinterpret_cast<uint8_t*>(roots.data()), F::ID));
        for (size_t i = 0; i < polys.size(); i++) out.push_back(CommittedOracle(handles[i], roots[i]));
        return out;
    }
    uint64_t size() const { return hodor_cuda_tree_size(handle_.get()); }
    Digest get_root() const { return root_; }
    F get_challenge_scalar_from_root() const {
        F c;
        check(hodor_cuda_tree_root(handle_.get(), nullptr, c.l));
        return c;
    }
    static bool verify_query(const Query& q, const Digest& root) {
        return Blake2sIopTree<F>::verify(root, q.value(), q.path(), q.tree_index());
    }
    Query query(size_t natural_index) const {
        size_t len = 0;
        while (((uint64_t)1 << len) < size()) len++;
        Query q;
        q.index = natural_index;
        q.path_.resize(len);
        check(hodor_cuda_tree_query(handle_.get(), natural_index, q.value_.l, reinterpret_cast<uint8_t*>(q.path_.data())));
        return q;
    }
    std::vector<F> values() const {  // the committed vector, copied back
        std::vector<F> v(size());
        check(hodor_cuda_tree_read(handle_.get(), 0, size(), reinterpret_cast<uint64_t*>(v.data()), nullptr));
        return v;
    }
    const void* device_values() const { return hodor_cuda_tree_values(handle_.get()); }

  private:
    CommittedOracle(hodor_tree* h, const Digest& root) : handle_(h, &hodor_cuda_tree_free), root_(root) {}
    std::shared_ptr<hodor_tree> handle_;
    Digest root_;
};

// ---- FRI -------------------------------------------------------------------------------------
template <class F>
struct FRIProof {  // src/fri/mod.rs:140-154
    std::vector<TrivialBlake2sIopQuery<F>> queries;
    std::vector<Digest> roots;
    std::vector<F> final_coefficients;
    size_t initial_degree_plus_one = 0, output_coeffs_at_degree_plus_one = 0, lde_factor = 0;
    const std::vector<F>& get_final_coefficients() const { return final_coefficients; }
};

template <class F>
class FRIProofPrototype {  // src/fri/mod.rs:107-138; trees and layer values stay in HBM behind the handle
  public:
    std::vector<F> challenges;
    Digest final_root{};
    std::vector<F> final_coefficients;
    size_t initial_degree_plus_one = 0, output_coeffs_at_degree_plus_one = 0, lde_factor = 0;

    FRIProofPrototype(hodor_fri_proto* h, size_t n, size_t factor, size_t out) : handle_(h, &hodor_cuda_fri_free), n_(n) {
        lde_factor = factor;
        output_coeffs_at_degree_plus_one = out;
        initial_degree_plus_one = n / factor;
        steps_ = check(hodor_cuda_fri_num_steps(h));
        roots_.resize(steps_ + 1);
        challenges.resize(steps_);
        final_coefficients.resize(out);
        check(hodor_cuda_fri_summary(h, reinterpret_cast<uint8_t*>(roots_.data()), reinterpret_cast<uint64_t*>(challenges.data()),
                                     reinterpret_cast<uint64_t*>(final_coefficients.data())));
        final_root = roots_.back();
    }
    int num_steps() const { return steps_; }
    std::vector<Digest> get_roots() const { return roots_; }
    Digest get_final_root() const { return final_root; }
    std::vector<F> get_final_coefficients() const { return final_coefficients; }
    std::vector<F> intermediate_values(size_t i) const {  // intermediate_values[i] as a host vector
        std::vector<F> v(n_ >> (i + 1));
        check(hodor_cuda_fri_layer(handle_.get(), (uint32_t)i + 1, nullptr, reinterpret_cast<uint64_t*>(v.data())));
        return v;
    }
    std::vector<Digest> commitment_nodes(size_t layer) const {  // layer 0 = l0_commitment
        std::vector<Digest> v(n_ >> layer);
        check(hodor_cuda_fri_layer(handle_.get(), (uint32_t)layer, reinterpret_cast<uint8_t*>(v.data()), nullptr));
        return v;
    }
    TrivialBlake2sIopQuery<F> query(size_t layer, size_t natural_index) const {
        const size_t size = n_ >> layer;
        size_t len = 0;
        while (((size_t)1 << len) < size) len++;
        TrivialBlake2sIopQuery<F> q;
        q.index = natural_index;
        q.path_.resize(len);
        check(hodor_cuda_fri_query(handle_.get(), (uint32_t)layer, natural_index, q.value_.l,
                                   reinterpret_cast<uint8_t*>(q.path_.data())));
        return q;
    }
    FRIProof<F> produce_proof(size_t natural_first_element_index) const {  // src/fri/query_producer.rs:10-53, one call
        FRIProof<F> proof;
        const size_t layers = (size_t)steps_ + 1;
        size_t depth0 = 0;
        while (((size_t)1 << depth0) < n_) depth0++;
        size_t total = 0;
        for (size_t l = 0; l < layers; l++) total += 2 * (depth0 - l);
        std::vector<uint64_t> idx(2 * layers);
        std::vector<F> vals(2 * layers);
        std::vector<Digest> paths(total);
        check(hodor_cuda_fri_produce_proof(handle_.get(), natural_first_element_index, idx.data(), reinterpret_cast<uint64_t*>(vals.data()),
                                           reinterpret_cast<uint8_t*>(paths.data())));
        size_t off = 0;
        for (size_t l = 0; l < layers; l++) {
            for (int q = 0; q < 2; q++) {
                TrivialBlake2sIopQuery<F> query;
                query.index = idx[2 * l + q];
                query.value_ = vals[2 * l + q];
                query.path_.assign(paths.begin() + off, paths.begin() + off + (depth0 - l));
                off += depth0 - l;
                proof.queries.push_back(std::move(query));
            }
            proof.roots.push_back(roots_[l]);
        }
        proof.final_coefficients = final_coefficients;
        proof.initial_degree_plus_one = initial_degree_plus_one;
        proof.output_coeffs_at_degree_plus_one = output_coeffs_at_degree_plus_one;
        proof.lde_factor = lde_factor;
        return proof;
    }

  private:
    std::unique_ptr<hodor_fri_proto, void (*)(hodor_fri_proto*)> handle_;
    size_t n_;
    int steps_ = 0;
    std::vector<Digest> roots_;
};

template <class F>
struct NaiveFriIop {  // src/fri/mod.rs:63-105
    static constexpr size_t DEGREE = 2;
    static FRIProofPrototype<F> proof_from_lde(const Polynomial<F, Values>& lde_values, size_t lde_factor,
                                               size_t output_coeffs_at_degree_plus_one, const Worker&) {
        hodor_fri_proto* h = hodor_cuda_fri_commit(reinterpret_cast<const uint64_t*>(lde_values.as_ref().data()),
                                                   lde_values.size(), (uint32_t)lde_factor,
                                                   (uint32_t)output_coeffs_at_degree_plus_one, 0, F::ID);
        if (!h) check(hodor_cuda_last_error_code() < 0 ? hodor_cuda_last_error_code() : HODOR_ERR_CUDA);  // DOMAIN -> SynthesisError, ...
        return FRIProofPrototype<F>(h, lde_values.size(), lde_factor, output_coeffs_at_degree_plus_one);
    }
    static FRIProof<F> prototype_into_proof(const FRIProofPrototype<F>& prototype, const Polynomial<F, Values>&,
                                            size_t natural_first_element_index) {
        return prototype.produce_proof(natural_first_element_index);
    }
};

// ---- several GPUs: one process per GPU (hodor_b200.h, "several GPUs") ------------------------------------------
// The reference's only parallelism is the threads of one host (`Worker`); its parallel_fft (src/fft/fft.rs:68-125) is
// the four-step decomposition over CPU threads and its multi-coset LDE runs the cosets on separate threads
// (src/polynomials/mod.rs:572-587).  Here a rank is a process that owns one GPU.  Rendezvous: rank 0 makes the id
// (Comm::unique_id) and ships its 128 bytes to the other ranks by whatever means the host program has (MPI, a file);
// every rank then constructs its Comm.  Same surface as rust/src/cuda/sharded.rs and hodor_b200/multigpu.py.

// A vector of field elements in HBM (32 bytes per element, the reference's own in-memory form).
template <class F>
class DeviceVec {
  public:
    explicit DeviceVec(size_t len) : len_(len), ptr_(hodor_cuda_malloc((len ? len : 1) * 32), &hodor_cuda_free) {
        if (!ptr_) check(hodor_cuda_last_error_code() < 0 ? hodor_cuda_last_error_code() : HODOR_ERR_CUDA);
    }
    static DeviceVec from_host(const std::vector<F>& v) {
        DeviceVec d(v.size());
        check(hodor_cuda_memcpy_h2d(d.ptr_.get(), v.data(), v.size() * 32, nullptr));
        check(hodor_cuda_stream_synchronize(nullptr));
        return d;
    }
    std::vector<F> to_host() const {
        std::vector<F> v(len_);
        check(hodor_cuda_memcpy_d2h(v.data(), ptr_.get(), len_ * 32, nullptr));
        check(hodor_cuda_stream_synchronize(nullptr));
        return v;
    }
    size_t size() const { return len_; }
    void* data() { return ptr_.get(); }
    const void* data() const { return ptr_.get(); }

  private:
    size_t len_;
    std::unique_ptr<void, void (*)(void*)> ptr_;
};

// What FriProofPrototype::get_roots, .challenges and get_final_coefficients give for the unsharded chain.
template <class F>
struct ShardedFriCommitment {
    std::vector<Digest> roots;
    std::vector<F> challenges;
    std::vector<F> final_coefficients;
};

class Comm {
  public:
    static std::array<uint8_t, 128> unique_id() {  // rank 0 only
        std::array<uint8_t, 128> id{};
        check(hodor_cuda_comm_unique_id(id.data()));
        return id;
    }
    // collective; world a power of two <= 16; world == 1 needs neither NCCL nor an id
    Comm(int rank, int world, const uint8_t* id) : rank_(rank), world_(world) { check(hodor_cuda_comm_init(rank, world, id)); }
    ~Comm() { hodor_cuda_comm_destroy(); }
    Comm(const Comm&) = delete;
    Comm& operator=(const Comm&) = delete;
    int rank() const { return rank_; }
    int world() const { return world_; }

    // best_fft (src/fft/fft.rs:5-19) on a vector dealt over the ranks.  local: this rank's cyclic slice
    // a[j * world + rank]; result: the rank-th (n / world^2)-element chunk of every length-(n / world) block of the
    // natural-order output (hodor_b200.h, hodor_cuda_ntt_sharded).  Collective.
    template <class F>
    DeviceVec<F> ntt(const DeviceVec<F>& local, uint32_t log_n, const F& omega) const {
        if (local.size() != (((size_t)1 << log_n) / (size_t)world_)) throw std::logic_error("ntt: wrong slice length");
        DeviceVec<F> out(local.size());
        check(hodor_cuda_ntt_sharded(local.data(), out.data(), log_n, omega.l, F::ID, nullptr));
        check(hodor_cuda_stream_synchronize(nullptr));
        return out;
    }
    // coset_lde (src/polynomials/mod.rs:349-352) followed by proof_from_lde_by_values (src/fri/fri_on_values.rs:11-159)
    // with cosets, folds and the bottom of every tree sharded over the ranks; coeffs replicated.  Collective.
    template <class F>
    ShardedFriCommitment<F> lde_fri(const DeviceVec<F>& coeffs, uint32_t log_n, size_t factor, bool coset,
                                    size_t output_coeffs_at_degree_plus_one) const {
        if (!factor || (factor & (factor - 1)) || !output_coeffs_at_degree_plus_one ||
            (output_coeffs_at_degree_plus_one & (output_coeffs_at_degree_plus_one - 1)))
            throw std::logic_error("assert!(factor.is_power_of_two())");
        uint32_t log_f = 0, log_o = 0;
        while (((size_t)1 << log_f) < factor) log_f++;
        while (((size_t)1 << log_o) < output_coeffs_at_degree_plus_one) log_o++;
        if (log_o >= log_n) throw std::logic_error("zero folding steps (the reference panics here)");
        const size_t steps = log_n - log_o;
        ShardedFriCommitment<F> r;
        r.roots.resize(steps + 1);
        r.challenges.resize(steps);
        r.final_coefficients.resize(output_coeffs_at_degree_plus_one);
        const int got = check(hodor_cuda_lde_fri_sharded(coeffs.data(), log_n, log_f, coset ? 1 : 0, (uint32_t)output_coeffs_at_degree_plus_one,
                                                         reinterpret_cast<uint8_t*>(r.roots.data()),
                                                         reinterpret_cast<uint64_t*>(r.challenges.data()),
                                                         reinterpret_cast<uint64_t*>(r.final_coefficients.data()), F::ID));
        if ((size_t)got != steps) throw CudaError("lde_fri_sharded: unexpected number of folding steps");
        return r;
    }

  private:
    int rank_, world_;
};

}  // namespace hodor_b200
