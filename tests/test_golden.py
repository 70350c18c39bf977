"""Committed known-answer vectors (tests/golden/vectors.json, made by tests/golden/make_golden.py from
the big-integer model): the C oracle must reproduce them on the CPU, the CUDA path on the GPU."""
import hashlib
import json
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
CASES = json.load(open(os.path.join(HERE, "golden", "vectors.json")))["cases"]


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def hexint(x: str) -> np.ndarray:
    v = int(x, 16)
    return np.array([(v >> (64 * i)) & (2**64 - 1) for i in range(4)], dtype=np.uint64)


def case_id(c):
    return f"{c['kind']}-f{c['field']}-2p{c['log_n']}"


class OracleImpl:
    def __init__(self, O):
        self.O = O

    def ntt(self, fid, a, ln):
        return self.O.serial_fft(fid, a, self.O.domain_generator(fid, ln), ln)

    def lde(self, fid, a, ln, L, coset):
        return self.O.lde(fid, a, ln, L, coset)

    def merkle(self, fid, a):
        nodes = self.O.merkle_create(fid, a)
        return nodes, self.O.interpret_hash(fid, nodes[1].tobytes())

    def fri(self, fid, a, L, oc):
        p = self.O.fri_commit(fid, a, L, oc)
        return p.roots(), p.challenges, p.final_root, p.final_coefficients, p.layer_values

    def batch_inversion(self, fid, a):
        return self.O.batch_inversion(fid, a)

    def evaluate_at(self, fid, a, z):
        return self.O.evaluate_at(fid, a, z)


class CudaImpl:
    def __init__(self, H):
        self.H = H

    def ntt(self, fid, a, ln):
        return self.H.Polynomial.from_coeffs(fid, a).fft().as_ref()

    def lde(self, fid, a, ln, L, coset):
        p = self.H.Polynomial.from_coeffs(fid, a)
        return (p.coset_lde(None, L) if coset else p.lde(None, L)).as_ref()

    def merkle(self, fid, a):
        t = self.H.Blake2sIopTree.create(fid, a)
        return t.nodes, t.get_challenge_scalar_from_root()

    def fri(self, fid, a, L, oc):
        p = self.H.NaiveFriIop.proof_from_lde(self.H.Polynomial.from_values(fid, a), L, oc, None)
        return p.get_roots(), p.challenges, p.get_final_root(), p.final_coefficients, [v.as_ref() for v in p.intermediate_values]

    def batch_inversion(self, fid, a):
        p = self.H.Polynomial.from_values(fid, a)
        p.batch_inversion(None)
        return p.as_ref()

    def evaluate_at(self, fid, a, z):
        return self.H.Polynomial.from_coeffs(fid, a).evaluate_at(None, z)


def check_case(impl, O, c):
    fid, ln = c["field"], c["log_n"]
    a = O.random_elements(fid, 1 << ln, c["seed"])
    if c["kind"] == "ntt":
        r = impl.ntt(fid, a, ln)
        assert sha(r) == c["sha256"]
        assert np.array_equal(r[0], hexint(c["first"])) and np.array_equal(r[-1], hexint(c["last"]))
    elif c["kind"] == "lde":
        assert sha(impl.lde(fid, a, ln, c["factor"], c["coset"])) == c["sha256"]
    elif c["kind"] == "merkle":
        nodes, chal = impl.merkle(fid, a)
        assert nodes[1].tobytes().hex() == c["root"]
        assert sha(nodes) == c["nodes_sha256"]
        assert np.array_equal(chal, hexint(c["challenge"]))
    elif c["kind"] == "batch_inversion":
        r = impl.batch_inversion(fid, a)
        assert sha(r) == c["sha256"] and np.array_equal(r[0], hexint(c["first"]))
    elif c["kind"] == "evaluate_at":
        z = O.random_elements(fid, 1, c["point_seed"])[0]
        assert np.array_equal(impl.evaluate_at(fid, a, z), hexint(c["value"]))
    elif c["kind"] == "fri":
        roots, chal, final_root, final_coeffs, values = impl.fri(fid, a, c["lde_factor"], c["out_coeffs"])
        assert [bytes(r).hex() for r in roots] == c["roots"]
        assert [hex(O.limbs_to_int(x)) for x in chal] == c["challenges"]
        assert bytes(final_root).hex() == c["final_root"]
        assert [hex(O.limbs_to_int(x)) for x in final_coeffs] == c["final_coefficients"]
        assert [sha(v) for v in values] == c["values_sha256"]


@pytest.mark.parametrize("c", CASES, ids=case_id)
def test_oracle_reproduces_golden(oracle, c):
    check_case(OracleImpl(oracle), oracle, c)


@pytest.mark.gpu
@pytest.mark.parametrize("c", CASES, ids=case_id)
def test_cuda_reproduces_golden(hodor, oracle, c):
    check_case(CudaImpl(hodor), oracle, c)
