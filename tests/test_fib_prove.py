"""BASELINE.json configs[3]: the Fibonacci `Prover::prove` call sequence.  CPU: the big-int model
(oracle/fib_model.py) is self-consistent (quotients are polynomials, every opening verifies).  GPU: the device
replay (hodor_b200/fib_replay.py) reproduces the model's proof object bit for bit."""
import numpy as np
import pytest


def test_fib_model_is_a_valid_prover_run(pymodel):
    """The reference's own checks on this shape: g's top coefficient is zero (test_fib_conversion_into_ali,
    src/ali/per_register/mod.rs:533-590) -- i.e. the constraint quotient is a polynomial --, h1 / h2 are low degree
    (the FRI chain's final coefficient reproduces the last layer), every opening verifies against its root."""
    from oracle import fib_model as FM
    F = pymodel.BLS12_381_FR
    for log_rows in (2, 4, 6):
        pr = FM.prove(F, log_rows, 16, 1)
        T = 1 << log_rows
        assert pr.stages["g_poly"][-1] == 0
        A, B = pr.stages["witness"]
        assert A[:4] == [1, 1, 2, 3][: len(A[:4])] and B[:3] == [1, 2, 3]
        omega = F.domain_generator(log_rows)
        for r, w in enumerate(pr.stages["f"]):  # witness polys re-evaluate to the witness
            assert [pymodel.evaluate(F, w, pow(omega, i, F.p)) for i in range(T)] == [A, B][r]
        for (idx, val, path), root in zip(pr.f_queries, pr.f_iop_roots):
            assert idx == pr.x_index_h1 and pymodel.merkle_verify(F, root, val, path, idx)
        assert pymodel.merkle_verify(F, pr.g_iop_root, pr.g_query[1], pr.g_query[2], pr.g_query[0])
        for qs, roots in ((pr.h1_queries, pr.h1_roots), (pr.h2_queries, pr.h2_roots)):
            assert len(qs) == 2 * len(roots)
            for k, (idx, val, path) in enumerate(qs):
                assert pymodel.merkle_verify(F, roots[k // 2], val, path, idx)
        assert pr.x_index_h1 % 2 == 1 and pr.x_index_h1 % 16 != 0


@pytest.mark.gpu
@pytest.mark.parametrize("log_rows", [2, 5, 8])
def test_fib_replay_matches_model(hodor, oracle, pymodel, log_rows):
    from hodor_b200 import fib_replay as R
    from oracle import fib_model as FM
    fid, F = 0, pymodel.BLS12_381_FR
    want = FM.prove(F, log_rows, 16, 1)
    a, b = R.fibonacci_witness(fid, 1 << log_rows)
    prover = R.FibonacciProver(fid, log_rows, 16, 1)
    got = prover.prove(a, b, keep_stages=True)

    def mont(x):
        return oracle.int_to_limbs(F.to_mont(x))

    def same(limbs, x):
        return np.array_equal(np.asarray(limbs, dtype=np.uint64).reshape(4), mont(x))

    for i in range(4):
        assert same(got.stages["dense_div_head"][i], want.stages["dense_div_head"][i]), "ALI divisor"
    for r in range(2):
        assert np.array_equal(got.stages["f"][r], np.stack([mont(x) for x in want.stages["f"][r]])), "witness polynomial"
    assert got.f_iop_roots == want.f_iop_roots
    assert np.array_equal(got.stages["g_poly"], np.stack([mont(x) for x in want.stages["g_poly"]])), "g"
    assert got.g_iop_root == want.g_iop_root
    assert same(got.stages["z"], want.stages["z"])
    for x, y in zip(got.f_at_z_m, want.f_at_z_m):
        assert same(x, y)
    assert same(got.g_at_z, want.g_at_z)
    for i in range(8):
        assert same(got.stages["h1_head"][i], want.stages["h1_head"][i]) and same(got.stages["h2_head"][i], want.stages["h2_head"][i])
    assert got.h1_roots == want.h1_roots and got.h2_roots == want.h2_roots
    assert np.array_equal(got.h1_final, np.stack([mont(x) for x in want.h1_final]))
    assert np.array_equal(got.h2_final, np.stack([mont(x) for x in want.h2_final]))
    assert (got.x_index_h1, got.x_index_h2) == (want.x_index_h1, want.x_index_h2)
    for q, (idx, val, path) in zip(got.f_queries + [got.g_query], want.f_queries + [want.g_query]):
        assert q.natural_index() == idx and same(q.value(), val) and q.path() == path
    for proof, wq in ((got.fri_proof_h1, want.h1_queries), (got.fri_proof_h2, want.h2_queries)):
        assert len(proof.queries) == len(wq)
        for q, (idx, val, path) in zip(proof.queries, wq):
            assert q.natural_index() == idx and same(q.value(), val) and q.path() == path
    # and the proof object is what a verifier accepts on the FRI side (src/fri/verifier.rs): low-degree check at x1
    assert hodor.NaiveFriIop.verify_proof(got.fri_proof_h1, got.x_index_h1, got.fri_proof_h1.queries[0].value()
                                         if got.fri_proof_h1.queries[0].natural_index() == got.x_index_h1
                                         else got.fri_proof_h1.queries[1].value())
