"""Committed known-answer vectors (tests/golden/vectors.json, made by tests/golden/make_golden.py from
the big-integer model): the C oracle must reproduce them on the CPU, the CUDA path on the GPU -- and sympy's
number-theoretic transform, an implementation from outside this repository, the NTT and LDE ones."""
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


class SympyImpl:
    """An implementation that shares nothing with this repository: sympy's number-theoretic transform over the prime,
    with sympy's own choice of primitive root (the smallest one = the generator the reference declares for the two
    fields it defines, see tests/test_third_party_pins.py).  Montgomery form in and out, as the fixtures hold it."""

    GENERATOR = {0: 7, 2: 3}

    def __init__(self, O):
        self.O = O

    def _io(self, fid):
        p = self.O.limbs_to_int(self.O.field_constants(fid)["p"])
        rinv = pow(1 << 256, -1, p)
        to_plain = lambda a: [self.O.limbs_to_int(x) * rinv % p for x in a]  # noqa: E731
        to_mont = lambda xs: np.stack([self.O.int_to_limbs((int(x) << 256) % p) for x in xs])  # noqa: E731
        return p, to_plain, to_mont

    def ntt(self, fid, a, ln):
        from sympy.discrete.transforms import ntt
        p, to_plain, to_mont = self._io(fid)
        return to_mont(ntt(to_plain(a), p))

    def lde(self, fid, a, ln, L, coset):
        from sympy.discrete.transforms import ntt
        p, to_plain, to_mont = self._io(fid)
        g = self.GENERATOR[fid] if coset else 1
        scaled = [x * pow(g, j, p) % p for j, x in enumerate(to_plain(a))]
        return to_mont(ntt(scaled + [0] * ((L - 1) << ln), p))


    # --- the FRI commit chain from its DEFINITION, not from the reference's pairing formula: interpolate the layer
    # (sympy intt), fold the coefficients f_even + c * f_odd, evaluate on the half-size domain (sympy ntt); trees and
    # challenges with hashlib's Blake2s (RFC 7693 keyed + personalised, src/iop/blake2s_trivial_iop.rs:8-16, :48-60).
    @staticmethod
    def _h(data: bytes) -> bytes:
        return hashlib.blake2s(data, key=b"Squeamish Ossifrage", person=b"Shaftoe", digest_size=32).digest()

    def _tree(self, fid, plain, p):
        leaves = [self._h(((x << 256) % p).to_bytes(32, "little")) for x in plain]  # encode_leaf: raw Montgomery limbs
        level = [self._h(leaves[2 * i] + leaves[2 * i + 1]) for i in range(len(leaves) // 2)]
        while len(level) > 1:
            level = [self._h(level[2 * i] + level[2 * i + 1]) for i in range(len(level) // 2)]
        root = level[0]
        capacity = self.O.field_constants(fid)["num_bits"] - 1
        top_mask = (2**64 - 1) >> ((256 - capacity) % 64)
        v = int.from_bytes(root, "big") & ((top_mask << 192) | (2**192 - 1))  # read_be, shave the top limb
        return root, v

    def merkle(self, fid, a):
        """Blake2sIopTree::create's heap (src/iop/blake2s_trivial_iop.rs:131-219) with hashlib: nodes[0] zero, nodes[1]
        the root, the level of w nodes at [w, 2w)."""
        p, to_plain, to_mont = self._io(fid)
        n = len(a)
        nodes = [bytes(32)] * n
        leaves = [self._h(np.ascontiguousarray(x).tobytes()) for x in a]
        for i in range(n // 2):
            nodes[n // 2 + i] = self._h(leaves[2 * i] + leaves[2 * i + 1])
        for i in range(n // 2 - 1, 0, -1):
            nodes[i] = self._h(nodes[2 * i] + nodes[2 * i + 1])
        _, v = self._tree(fid, to_plain(a), p)
        return np.frombuffer(b"".join(nodes), np.uint8).reshape(n, 32), to_mont([v])[0]

    def batch_inversion(self, fid, a):
        p, to_plain, to_mont = self._io(fid)
        return to_mont([pow(x, -1, p) for x in to_plain(a)])

    def evaluate_at(self, fid, a, z):
        p, to_plain, to_mont = self._io(fid)
        zz = to_plain([z])[0]
        return to_mont([sum(x * pow(zz, j, p) for j, x in enumerate(to_plain(a))) % p])[0]

    def fri(self, fid, a, L, oc):
        from sympy.discrete.transforms import intt, ntt
        p, to_plain, to_mont = self._io(fid)
        values = to_plain(a)
        steps = (len(values) // L // oc).bit_length() - 1
        roots, challenges, layer_values = [], [], []
        root, c = self._tree(fid, values, p)
        roots.append(root)
        coeffs = intt(values, p)
        for _ in range(steps):
            challenges.append(c)
            coeffs = [(coeffs[2 * j] + c * coeffs[2 * j + 1]) % p for j in range(len(coeffs) // 2)]
            values = ntt(coeffs, p)
            layer_values.append(to_mont(values))
            root, c = self._tree(fid, values, p)
            roots.append(root)
        return roots, to_mont(challenges), roots[-1], to_mont(coeffs[:oc]), layer_values


THIRD_PARTY_CASES = [c for c in CASES if c["field"] in SympyImpl.GENERATOR and c["log_n"] >= 1]


@pytest.mark.parametrize("c", THIRD_PARTY_CASES, ids=case_id)
def test_third_party_code_reproduces_golden(oracle, c):
    """Every committed fixture of the two fields the reference declares, from code written by someone else: sympy's
    transforms, hashlib's Blake2s, Python's modular pow (the FRI chain restated from its definition: interpolate,
    fold the coefficients, re-evaluate -- not from the reference's pairing formula)."""
    pytest.importorskip("sympy")
    check_case(SympyImpl(oracle), oracle, c)


@pytest.mark.gpu
@pytest.mark.parametrize("c", CASES, ids=case_id)
def test_cuda_reproduces_golden(hodor, oracle, c):
    check_case(CudaImpl(hodor), oracle, c)
