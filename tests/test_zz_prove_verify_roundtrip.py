"""Prove -> query -> verify round trip through the mirror of the reference's FriIop trait
(src/fri/mod.rs:36-61), shaped after test_fib_fri_iop_verifier (src/fri/mod.rs:364-505): the commit chain
and the queries come from the GPU (hodor_cuda_fri_commit / hodor_cuda_fri_query), the verifier is the
host-side NaiveFriIop.verify_proof (src/fri/verifier.rs:130-290).  Runs last (file name) on purpose."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("fid", [0, 1, 2], ids=["bls12_381_fr", "bn254_fr", "stark252"])
def test_fri_prove_query_verify(hodor, oracle, fid):
    W = hodor.Worker()
    lde_factor, out = 8, 1
    coeffs = oracle.random_elements(fid, 1 << 10, seed=500 + fid)
    lde = hodor.Polynomial.from_coeffs(fid, coeffs).coset_lde(W, lde_factor)
    other = hodor.Polynomial.from_coeffs(fid, oracle.random_elements(fid, 1 << 10, seed=600 + fid)).coset_lde(W, lde_factor)
    proto = hodor.NaiveFriIop.proof_from_lde(lde, lde_factor, out, W)
    n = lde.size()
    for index in (63, 1, n // 2 + 33, n - 1):  # odd indices: the reference's domain check rejects even ones
        proof = hodor.NaiveFriIop.prototype_into_proof(proto, lde, index)
        assert hodor.NaiveFriIop.verify_proof(proof, index, lde.as_ref()[index]) is True
        assert hodor.NaiveFriIop.verify_proof(proof, index, other.as_ref()[index]) is False
    proof = hodor.NaiveFriIop.prototype_into_proof(proto, lde, 63)
    proof.final_coefficients = proof.final_coefficients.copy()
    proof.final_coefficients[0, 0] ^= np.uint64(1)
    assert hodor.NaiveFriIop.verify_proof(proof, 63, lde.as_ref()[63]) is False
