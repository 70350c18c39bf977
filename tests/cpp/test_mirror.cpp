// C++ parity tests through the host mirror (include/hodor_b200.hpp), shaped after the reference's own
// tests.  Links the product (libhodor_b200.so) and, as the checker only, the
// CPU oracle (oracle/_build/libhodor_oracle.so).  Needs a B200; run by tests/test_cpp_mirror.py.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "hodor_b200.hpp"

using namespace hodor_b200;

extern "C" {  // oracle/hodor_oracle.c (TEST ONLY)
int oracle_random_elements(int field, uint64_t* out, size_t count, uint64_t seed);
int oracle_serial_fft(int field, uint64_t* a, const uint64_t* omega, uint32_t log_n);
int oracle_lde(int field, const uint64_t* coeffs, uint32_t log_n, uint32_t factor, int coset, uint64_t* out, uint32_t cpus);
int oracle_merkle_create(int field, const uint64_t* leaves, size_t n, uint8_t* nodes, uint32_t cpus);
int oracle_batch_inversion(int field, uint64_t* a, size_t n, uint32_t cpus);
int oracle_evaluate_at(int field, const uint64_t* coeffs, size_t n, const uint64_t* g, uint32_t cpus, uint64_t* out);
}

// compile the batch entry of the mirror for one field (exercised through the Python mirror's GPU test)
template std::vector<Polynomial<Bn256RsFr, Values>> hodor_b200::lde_batch<Bn256RsFr>(
    const std::vector<Polynomial<Bn256RsFr, Coefficients>>&, const Worker&, size_t, bool);

static int failures = 0;
#define CHECK(cond)                                                          \
    do {                                                                     \
        if (!(cond)) {                                                       \
            std::printf("FAIL %s:%d  %s\n", __FILE__, __LINE__, #cond);      \
            failures++;                                                      \
        }                                                                    \
    } while (0)

template <class F>
static std::vector<F> random_vec(size_t n, uint64_t seed) {
    std::vector<F> v(n);
    oracle_random_elements(F::ID, reinterpret_cast<uint64_t*>(v.data()), n, seed);
    return v;
}

// test_worker_size (src/fft/mod.rs:281): forward == serial_fft, inverse round trip == input
template <class F>
static void test_fft_roundtrip() {
    const Worker worker;
    for (uint32_t log_n : {0u, 1u, 5u, 10u, 13u, 16u}) {
        const auto a = random_vec<F>((size_t)1 << log_n, 100 + log_n);
        auto coeffs = Polynomial<F, Coefficients>::from_coeffs(a);
        const F omega = coeffs.omega;
        auto values = std::move(coeffs).fft(worker);
        std::vector<F> want = a;
        oracle_serial_fft(F::ID, reinterpret_cast<uint64_t*>(want.data()), omega.l, log_n);
        CHECK(values.as_ref() == want);
        auto back = std::move(values).ifft(worker);
        CHECK(back.as_ref() == a);
        auto cback = Polynomial<F, Coefficients>::from_coeffs(a).coset_fft(worker).icoset_fft(worker);
        CHECK(cback.as_ref() == a);
    }
}

// test_batch_inversion (src/polynomials/mod.rs:959-985): batch inverse == per-element inverse;
// evaluate_at (:685-711) against the oracle and against a coset LDE value
template <class F>
static void test_batch_inversion_and_evaluate() {
    const Worker worker;
    const auto a = random_vec<F>(1 << 10, 41);
    auto values = Polynomial<F, Values>::from_values(a);
    values.batch_inversion(worker);
    for (size_t i = 0; i < a.size(); i += 97) CHECK(values.as_ref()[i] == a[i].inverse().second);
    std::vector<F> want = a;
    CHECK(oracle_batch_inversion(F::ID, reinterpret_cast<uint64_t*>(want.data()), want.size(), 4) == 0);
    CHECK(values.as_ref() == want);
    auto with_zero = a;
    with_zero[77] = F::zero();
    auto bad = Polynomial<F, Values>::from_values(with_zero);
    bool threw = false;
    try {
        bad.batch_inversion(worker);
    } catch (const SynthesisError&) {
        threw = true;
    }
    CHECK(threw);
    CHECK(bad.as_ref() == with_zero);
    // elementwise ops against scalar field arithmetic
    {
        const auto b = random_vec<F>(1 << 10, 42);
        auto x = Polynomial<F, Values>::from_values(a);
        const auto y = Polynomial<F, Values>::from_values(b);
        x.add_assign_scaled(worker, y, a[3]);   // a + b * s
        x.add_constant(worker, b[7]);           // + c
        x.negate(worker);
        x.square(worker);
        x.pow(worker, 5);
        x.scale(worker, a[9]);
        x.mul_assign(worker, y);
        x.sub_assign(worker, y);
        for (size_t i = 0; i < a.size(); i += 131) {
            F t = b[i];
            t.mul_assign(a[3]);
            F e = a[i];
            e.add_assign(t);
            e.add_assign(b[7]);
            F neg = F::zero();
            neg.sub_assign(e);
            neg.square();
            F r = neg.pow(5);
            r.mul_assign(a[9]);
            r.mul_assign(b[i]);
            r.sub_assign(b[i]);
            CHECK(x.as_ref()[i] == r);
        }
    }
    const auto coeffs = Polynomial<F, Coefficients>::from_coeffs(a);
    const F z = a[5];
    F expect;
    oracle_evaluate_at(F::ID, reinterpret_cast<const uint64_t*>(a.data()), a.size(), z.l, 3, expect.l);
    CHECK(coeffs.evaluate_at(worker, z) == expect);
    const auto lde = coeffs.clone().coset_lde(worker, 4);
    F x = F::multiplicative_generator();
    x.mul_assign(Domain<F>::new_for_size(a.size() * 4).generator.pow(9));
    CHECK(coeffs.evaluate_at(worker, x) == lde.as_ref()[9]);
}

// test_lde_correctness / test_coset_lde_correctness (src/polynomials/mod.rs:988, 1036)
template <class F>
static void test_lde_correctness() {
    const Worker worker;
    for (auto [log_n, factor] : {std::pair<uint32_t, size_t>{2, 16}, {8, 8}, {12, 8}, {14, 2}}) {
        const size_t n = (size_t)1 << log_n;
        const auto a = random_vec<F>(n, 7 * log_n + factor);
        for (int coset = 0; coset < 2; coset++) {
            auto poly = Polynomial<F, Coefficients>::from_coeffs(a);
            auto lde = coset ? std::move(poly).coset_lde(worker, factor) : std::move(poly).lde(worker, factor);
            std::vector<F> want(n * factor);
            oracle_lde(F::ID, reinterpret_cast<const uint64_t*>(a.data()), log_n, (uint32_t)factor, coset,
                       reinterpret_cast<uint64_t*>(want.data()), (uint32_t)factor);
            CHECK(lde.as_ref() == want);
            // multi-coset LDE == (coset) NTT of the zero-padded vector
            std::vector<F> padded(n * factor);
            std::copy(a.begin(), a.end(), padded.begin());
            auto p2 = Polynomial<F, Coefficients>::from_coeffs(padded);
            auto full = coset ? std::move(p2).coset_fft(worker) : std::move(p2).fft(worker);
            CHECK(full.as_ref() == lde.as_ref());
        }
    }
    bool threw = false;
    try {
        auto p = Polynomial<F, Coefficients>::from_coeffs(random_vec<F>(4, 1));
        (void)std::move(p).lde(worker, 3);
    } catch (const std::logic_error&) {
        threw = true;
    }
    CHECK(threw);  // assert!(factor.is_power_of_two())
}

// make_small_iop (src/iop/blake2s_trivial_iop.rs:390-408) + bit-exact nodes
template <class F>
static void test_small_iop() {
    std::vector<F> inputs;
    F f = F::one();
    for (int i = 0; i < 64; i++) {
        inputs.push_back(f);
        f.double_();
    }
    const auto iop = TrivialBlake2sIOP<F>::create(inputs);
    const Digest root = iop.get_root();
    for (size_t i = 0; i < 64; i++) {
        const auto q = iop.query(i, inputs);
        CHECK(TrivialBlake2sIOP<F>::verify_query(q, root));
        auto bad = q;
        bad.value_.double_();
        CHECK(!TrivialBlake2sIOP<F>::verify_query(bad, root));
    }
    const auto leaves = random_vec<F>(1 << 13, 5);
    const auto tree = Blake2sIopTree<F>::create(leaves);
    std::vector<Digest> want(leaves.size());
    oracle_merkle_create(F::ID, reinterpret_cast<const uint64_t*>(leaves.data()), leaves.size(),
                         reinterpret_cast<uint8_t*>(want.data()), 4);
    CHECK(tree.nodes() == want);
}

// test_one_fri_step (src/fri/mod.rs:252-331) and the query producer
template <class F>
static void test_one_fri_step() {
    const Worker worker;
    std::vector<F> lde_coeffs;
    F f = F::one();
    for (int i = 0; i < 4; i++) {
        lde_coeffs.push_back(f);
        f.double_();
    }
    const size_t lde_factor = 4;
    auto lde_values = Polynomial<F, Coefficients>::from_coeffs(lde_coeffs).lde(worker, lde_factor);
    const auto proto = NaiveFriIop<F>::proof_from_lde(lde_values, lde_factor, 2, worker);
    CHECK(proto.num_steps() == 1);
    const F challenge = proto.challenges[0];
    std::vector<F> new_coeffs;
    for (size_t k = 0; k < 4; k += 2) {
        F tmp = lde_coeffs[k + 1];
        tmp.mul_assign(challenge);
        tmp.add_assign(lde_coeffs[k]);
        new_coeffs.push_back(tmp);
    }
    CHECK(proto.final_coefficients == new_coeffs);
    auto next_lde = Polynomial<F, Coefficients>::from_coeffs(new_coeffs).lde(worker, lde_factor);
    CHECK(proto.intermediate_values(0) == next_lde.as_ref());
    CHECK(proto.get_roots()[1] == TrivialBlake2sIOP<F>::create(next_lde.as_ref()).get_root());
    // hand interpolation at coset index 3 (:283-300)
    const size_t idx = 3, pair = idx + lde_factor * 2;
    const F divisor = lde_values.omegainv.pow(idx);
    const F two_inv = F::from_u64(2).inverse().second;
    F t0 = lde_values.as_ref()[idx];
    t0.add_assign(lde_values.as_ref()[pair]);
    F t1 = lde_values.as_ref()[idx];
    t1.sub_assign(lde_values.as_ref()[pair]);
    t1.mul_assign(divisor);
    t1.mul_assign(challenge);
    t0.add_assign(t1);
    t0.mul_assign(two_inv);
    CHECK(next_lde.as_ref()[idx] == t0);
    for (size_t start : {1u, 3u, 7u, 12u}) {
        const auto proof = NaiveFriIop<F>::prototype_into_proof(proto, lde_values, start);
        CHECK(proof.queries.size() == 2 * proof.roots.size());
        for (size_t k = 0; k < proof.queries.size(); k++)
            CHECK(TrivialBlake2sIOP<F>::verify_query(proof.queries[k], proof.roots[k / 2]));
    }
    bool threw = false;
    try {
        (void)NaiveFriIop<F>::proof_from_lde(lde_values, 16, 1, worker);  // zero folding steps
    } catch (const std::logic_error&) {
        threw = true;
    }
    CHECK(threw);
}

// the register loop of Prover::prove (src/prover/mod.rs:73-80, :142-151): lde + I::create with everything resident
// in HBM, then openings from the handle; against the oracle's LDE and tree
template <class F>
static void test_committed_oracle() {
    const uint32_t log_n = 10;
    const size_t n = (size_t)1 << log_n, factor = 8;
    std::vector<Polynomial<F, Coefficients>> polys;
    for (int r = 0; r < 3; r++) polys.push_back(Polynomial<F, Coefficients>::from_coeffs(random_vec<F>(n, 700 + r)));
    const auto oracles = CommittedOracle<F>::lde_commit_batch(polys, factor, false);
    CHECK(oracles.size() == 3);
    for (int r = 0; r < 3; r++) {
        std::vector<F> want(n * factor);
        oracle_lde(F::ID, reinterpret_cast<const uint64_t*>(polys[r].as_ref().data()), log_n, (uint32_t)factor, 0,
                   reinterpret_cast<uint64_t*>(want.data()), 16);
        std::vector<Digest> nodes(n * factor);
        oracle_merkle_create(F::ID, reinterpret_cast<const uint64_t*>(want.data()), n * factor, reinterpret_cast<uint8_t*>(nodes.data()), 4);
        CHECK(oracles[r].size() == n * factor);
        CHECK(oracles[r].get_root() == nodes[1]);
        CHECK(oracles[r].values() == want);
        CHECK(oracles[r].get_challenge_scalar_from_root() == Blake2sIopTree<F>::encode_root_into_challenge(nodes[1]));
        for (size_t idx : {(size_t)0, (size_t)1, n * factor / 2 + 5, n * factor - 1}) {
            const auto q = oracles[r].query(idx);
            CHECK(q.value() == want[idx]);
            CHECK(CommittedOracle<F>::verify_query(q, oracles[r].get_root()));
        }
        CHECK(CommittedOracle<F>::create(want).get_root() == nodes[1]);
    }
}

template <class F>
static void test_domain() {
    CHECK(Domain<F>::new_for_size(5).size == 8);
    bool threw = false;
    try {
        (void)Domain<F>::new_for_size(((uint64_t)1 << F::S()) + 1);
    } catch (const SynthesisError&) {
        threw = true;
    }
    CHECK(threw || F::S() >= 63);
    F g = Domain<F>::new_for_size(16).generator;
    CHECK(g.pow(16) == F::one() && g.pow(8) != F::one());
}

// `Prover::new`'s setup vectors against the reference's formulas restated with scalar host arithmetic
// (src/precomputations/mod.rs:14-66, src/ali/per_register/mod.rs:60-162, 214-227)
template <class F>
static void test_precomputations() {
    const Worker worker;
    const auto col = Domain<F>::new_for_size(16), ev = Domain<F>::new_for_size(64);
    const auto pre = PrecomputedOmegas<F>::new_for_domain(ev, worker);
    CHECK(pre.omegas.size() == 64 && pre.coset.size() == 64 && pre.omegas_inv.size() == 32);
    F u = F::one();
    for (size_t i = 0; i < 64; i++) {
        F c = u;
        c.mul_assign(F::multiplicative_generator());
        CHECK(pre.omegas[i] == u && pre.coset[i] == c);
        if (i < 32) {
            F prod = pre.omegas_inv[i];
            prod.mul_assign(u);
            CHECK(prod == F::one());
        }
        u.mul_assign(ev.generator);
    }
    const size_t start_at = 2, span = 3;
    const uint64_t num_rows = 15;
    auto [inv, degree] = inverse_divisor_for_dense_constraint_in_coset(col, ev, DenseConstraint{start_at, span}, num_rows, worker);
    CHECK(degree == 16 - start_at - (16 - num_rows) - span);
    const auto boundary = boundary_constraint_inverse_divisor(col, ev, 5, worker);
    F x = F::multiplicative_generator();
    for (size_t j = 0; j < 64; j++) {
        F den = x.pow(16);
        den.sub_assign(F::one());
        F want = den.inverse().second;
        for (uint64_t k = 0; k < 16; k++) {
            if (k >= start_at && k < num_rows - span) continue;
            F t = x;
            t.sub_assign(col.generator.pow(k));
            want.mul_assign(t);
        }
        CHECK(inv.as_ref()[j] == want);
        F b = x;
        b.sub_assign(col.generator.pow(5));
        CHECK(boundary.as_ref()[j] == b.inverse().second);
        x.mul_assign(ev.generator);
    }
    bool threw = false;
    try {
        inverse_divisor_for_dense_constraint_in_coset(ev, col, DenseConstraint{0, 1}, 64, worker);
    } catch (const std::exception&) {
        threw = true;
    }
    CHECK(threw);
}

template <class F>
static void run_all(const char* name) {
    const int before = failures;
    test_domain<F>();
    test_fft_roundtrip<F>();
    test_lde_correctness<F>();
    test_batch_inversion_and_evaluate<F>();
    test_small_iop<F>();
    test_one_fri_step<F>();
    test_committed_oracle<F>();
    test_precomputations<F>();
    std::printf("%s: %s\n", name, failures == before ? "ok" : "FAILED");
}

// the multi-GPU host at world 1 (no NCCL needed): the sharded entry points must equal the single-GPU ones
template <class F>
static void test_comm_world1() {
    const Worker worker;
    Comm comm(0, 1, nullptr);
    const uint32_t log_n = 13;
    const auto a = random_vec<F>((size_t)1 << log_n, 4242);
    auto poly = Polynomial<F, Coefficients>::from_coeffs(a);
    const F omega = poly.omega;
    const auto got = comm.ntt(DeviceVec<F>::from_host(a), log_n, omega).to_host();
    CHECK(got == std::move(poly).fft(worker).as_ref());
    const size_t factor = 8;
    const auto chain = comm.lde_fri(DeviceVec<F>::from_host(a), log_n, factor, true, 1);
    auto lde = Polynomial<F, Coefficients>::from_coeffs(a).coset_lde(worker, factor);
    const auto proto = NaiveFriIop<F>::proof_from_lde(lde, factor, 1, worker);
    CHECK(chain.roots == proto.get_roots());
    CHECK(chain.challenges == proto.challenges);
    CHECK(chain.final_coefficients == proto.final_coefficients);
}

int main() {
    try {
        init(0);
        test_comm_world1<Bn256RsFr>();
        test_comm_world1<Stark252Fr>();
        run_all<Bn256RsFr>("bn256.rs Fr (BLS12-381 Fr)");
        run_all<Bn254Fr>("BN254 Fr");
        run_all<Stark252Fr>("Stark252");
    } catch (const std::exception& e) {
        std::printf("EXCEPTION: %s\n", e.what());
        return 2;
    }
    std::printf("%d failure(s)\n", failures);
    return failures ? 1 : 0;
}
