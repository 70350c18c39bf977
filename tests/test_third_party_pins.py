"""The transforms against an INDEPENDENT third-party implementation: sympy's number-theoretic transform.

`sympy.discrete.transforms.ntt(seq, prime)` computes X_k = sum_j a_j w^(jk) mod p with w = g^((p-1)/n) and g =
`sympy.primitive_root(p)`, the smallest primitive root.  For the two fields the reference declares that is exactly
the reference's convention -- `PrimeFieldGenerator = "7"` for src/bn256.rs:5-6 and "3" for src/experiments/mod.rs:19-20,
both the smallest primitive roots, and Domain::new_for_size's generator root_of_unity^(2^(S - log n)) = g^((p-1)/n)
(src/domains/mod.rs:21-44) -- so sympy's output must equal `Polynomial::fft`'s (src/polynomials/mod.rs:611-624) value
for value.  sympy shares nothing with this repository's oracle, big-int model or CUDA code; together with hashlib's
Blake2s (tests/test_oracle_pins.py) it anchors every building block of the path to code written by someone else.
(pairing_ce's BN254 Fr declares generator 7, not the smallest primitive root 5, so it is not comparable this way.)"""
import numpy as np
import pytest

sympy = pytest.importorskip("sympy")
from sympy.discrete.transforms import intt, ntt  # noqa: E402

FIELDS = {0: ("bls12_381_fr", 7), 2: ("stark252", 3)}


def _plain(oracle, fid, mont):
    return oracle.array_to_ints(oracle.from_mont(fid, mont))


def _modulus(oracle, fid):
    return oracle.limbs_to_int(oracle.field_constants(fid)["p"])


@pytest.mark.parametrize("fid", FIELDS, ids=[v[0] for v in FIELDS.values()])
def test_declared_generator_is_sympys_primitive_root(oracle, fid):
    p = _modulus(oracle, fid)
    assert sympy.primitive_root(p) == FIELDS[fid][1]
    assert _plain(oracle, fid, oracle.field_constants(fid)["generator"][None, :]) == [FIELDS[fid][1]]


@pytest.mark.parametrize("fid", FIELDS, ids=[v[0] for v in FIELDS.values()])
@pytest.mark.parametrize("log_n", [1, 2, 3, 5, 8, 10])
def test_oracle_fft_and_ifft_match_sympy(oracle, fid, log_n):
    p, n = _modulus(oracle, fid), 1 << log_n
    a = oracle.random_elements(fid, n, seed=900 + log_n)
    plain = _plain(oracle, fid, a)
    omega = oracle.domain_generator(fid, log_n)
    assert _plain(oracle, fid, oracle.serial_fft(fid, a, omega, log_n)) == ntt(plain, p)
    assert _plain(oracle, fid, oracle.best_fft(fid, a, omega, log_n, cpus=4)) == ntt(plain, p)
    assert _plain(oracle, fid, oracle.ifft(fid, a, log_n)) == intt(plain, p)


@pytest.mark.parametrize("fid", FIELDS, ids=[v[0] for v in FIELDS.values()])
@pytest.mark.parametrize("coset", [False, True], ids=["lde", "coset_lde"])
def test_oracle_lde_matches_sympy_on_the_padded_vector(oracle, fid, coset):
    """(coset_)lde_using_multiple_cosets (src/polynomials/mod.rs:418-482, 544-609) == the size-nL transform of the
    zero-padded coefficients, scaled by g^j first for the coset form."""
    p, log_n, L = _modulus(oracle, fid), 6, 8
    a = oracle.random_elements(fid, 1 << log_n, seed=77)
    plain = _plain(oracle, fid, a)
    g = FIELDS[fid][1]
    scaled = [x * pow(g, j, p) % p for j, x in enumerate(plain)] if coset else plain
    want = ntt(scaled + [0] * ((L - 1) << log_n), p)
    assert _plain(oracle, fid, oracle.lde(fid, a, log_n, L, coset)) == want


@pytest.mark.gpu
@pytest.mark.parametrize("fid", FIELDS, ids=[v[0] for v in FIELDS.values()])
def test_gpu_transforms_match_sympy(hodor, oracle, fid):
    """The CUDA path itself against sympy: single-block (2^10) and multi-pass (2^12, 2^13) transforms, the inverse,
    and a coset LDE 2^10 x 4 (multi-pass over 2^12 outputs)."""
    p, g = _modulus(oracle, fid), FIELDS[fid][1]
    W = hodor.Worker()
    for log_n in (10, 12, 13):
        a = oracle.random_elements(fid, 1 << log_n, seed=940 + log_n)
        plain = _plain(oracle, fid, a)
        got = hodor.Polynomial.from_coeffs(fid, a).fft(W)
        assert _plain(oracle, fid, got.as_ref()) == ntt(plain, p)
        back = hodor.Polynomial.from_values(fid, a).ifft(W)
        assert _plain(oracle, fid, back.as_ref()) == intt(plain, p)
    log_n, L = 10, 4
    a = oracle.random_elements(fid, 1 << log_n, seed=951)
    plain = _plain(oracle, fid, a)
    scaled = [x * pow(g, j, p) % p for j, x in enumerate(plain)]
    want = ntt(scaled + [0] * ((L - 1) << log_n), p)
    got = hodor.Polynomial.from_coeffs(fid, a).coset_lde(W, L)
    assert _plain(oracle, fid, got.as_ref()) == want


def test_oracle_against_third_party_on_random_shapes(oracle):
    """Random small shapes (hypothesis): sizes 1 .. 2^7, blowups 1 .. 16, plain and coset, both reference-declared
    fields, structured and random inputs -- the oracle's LDE against sympy on the padded (and g^j-scaled) vector, and
    its FRI commit chain against the chain restated from its definition (tests/test_golden.py SympyImpl)."""
    hyp = pytest.importorskip("hypothesis")
    from hypothesis import given, settings, strategies as st

    from test_golden import SympyImpl

    third = SympyImpl(oracle)

    @settings(max_examples=40, deadline=None, derandomize=True)
    @given(fid=st.sampled_from([0, 2]), log_n=st.integers(0, 7), log_l=st.integers(0, 4), coset=st.booleans(),
           seed=st.integers(1, 2**32), shape=st.sampled_from(["random", "zero", "one", "delta", "minus_one"]))
    def lde_case(fid, log_n, log_l, coset, seed, shape):
        n, L = 1 << log_n, 1 << log_l
        p = _modulus(oracle, fid)
        a = oracle.random_elements(fid, n, seed)
        if shape != "random":
            plain = {"zero": [0] * n, "one": [1] * n, "delta": [1] + [0] * (n - 1), "minus_one": [p - 1] * n}[shape]
            a = oracle.to_mont(fid, oracle.ints_to_array(plain))
        got = oracle.lde(fid, a, log_n, L, coset) if L > 1 else oracle.fft(fid, a, log_n, coset=coset)
        if n * L == 1:
            assert np.array_equal(got, a)  # a transform of length one is the identity (also for the coset: g^0)
        else:
            assert np.array_equal(got, third.lde(fid, a, log_n, L, coset))

    @settings(max_examples=25, deadline=None, derandomize=True)
    @given(fid=st.sampled_from([0, 2]), log_n=st.integers(2, 7), log_l=st.integers(1, 3), log_o=st.integers(0, 2),
           seed=st.integers(1, 2**32))
    def fri_case(fid, log_n, log_l, log_o, seed):
        n, L, oc = 1 << log_n, 1 << log_l, 1 << log_o
        if n // L // oc < 2:
            return  # zero folding steps: the reference panics (covered by the argument-error tests)
        v = oracle.random_elements(fid, n, seed)
        want = oracle.fri_commit(fid, v, L, oc)
        roots, chal, final_root, final_coeffs, values = third.fri(fid, v, L, oc)
        assert [bytes(r) for r in roots] == want.roots() and final_root == want.final_root
        assert np.array_equal(chal, want.challenges) and np.array_equal(final_coeffs, want.final_coefficients)
        assert all(np.array_equal(x, y) for x, y in zip(values, want.layer_values))

    lde_case()
    fri_case()
