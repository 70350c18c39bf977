"""Mirror of the reference's FRI prover surface: trait FriIop and the proof structs in
src/fri/mod.rs:26-154, NaiveFriIop::proof_from_lde_by_values in src/fri/fri_on_values.rs:11-159 and
FRIProofPrototype::produce_proof in src/fri/query_producer.rs:10-53.

The whole commit chain (l0 tree, then per layer: fold, tree, root -> challenge; final iNTT) is
enqueued on one CUDA stream by hodor_cuda_fri_commit and stays in HBM behind a handle.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional

import numpy as np

from . import field as fld
from ._ffi import HodorError, SynthesisError, check, ensure_init, last_error, lib, raise_last, u8p
from .domains import Domain
from .field import _p
from .iop import DeviceIOP, TrivialBlake2sIOP, TrivialBlake2sIopQuery, TrivialCombiner
from .polynomials import VALUES, Polynomial, Worker


class FRIProof:
    """src/fri/mod.rs:140-154"""

    def __init__(self, queries, roots, final_coefficients, initial_degree_plus_one, output_coeffs_at_degree_plus_one,
                 lde_factor, field_id: int = 0):
        self.field_id = field_id
        self.queries = queries
        self.roots = roots
        self.final_coefficients = final_coefficients
        self.initial_degree_plus_one = initial_degree_plus_one
        self.output_coeffs_at_degree_plus_one = output_coeffs_at_degree_plus_one
        self.lde_factor = lde_factor

    def get_final_coefficients(self) -> np.ndarray:
        return self.final_coefficients


class FRIProofPrototype:
    """src/fri/mod.rs:107-138, backed by a device handle."""

    def __init__(self, field_id: int, handle: int, n: int, lde_factor: int, out_coeffs: int):
        self.field_id = field_id
        self._handle = handle
        self._n = n
        self.lde_factor = lde_factor
        self.output_coeffs_at_degree_plus_one = out_coeffs
        self.initial_degree_plus_one = n // lde_factor
        steps = check(lib.hodor_cuda_fri_num_steps(handle))
        self.num_steps = steps
        roots = np.zeros((steps + 1, 32), np.uint8)
        self.challenges = np.zeros((steps, 4), np.uint64)
        self.final_coefficients = np.zeros((out_coeffs, 4), np.uint64)
        check(lib.hodor_cuda_fri_summary(handle, roots.ctypes.data_as(u8p), _p(self.challenges), _p(self.final_coefficients)))
        self._roots = [r.tobytes() for r in roots]
        self.final_root = self._roots[-1]
        self.l0_commitment = DeviceIOP(field_id, self, 0, n, self._roots[0])
        self.intermediate_commitments = [DeviceIOP(field_id, self, i + 1, n >> (i + 1), self._roots[i + 1])
                                         for i in range(steps)]
        self._values: List[Optional[Polynomial]] = [None] * steps

    def free(self) -> None:
        """Return the prototype's device memory (trees, layer values) to the library."""
        h, self._handle = getattr(self, "_handle", None), None
        if h:
            lib.hodor_cuda_fri_free(h)

    def __del__(self):
        self.free()

    # FriProofPrototype trait (:26-30, :119-138)
    def get_roots(self) -> List[bytes]:
        return list(self._roots)

    def get_final_root(self) -> bytes:
        return self.final_root

    def get_final_coefficients(self) -> np.ndarray:
        return self.final_coefficients.copy()

    @property
    def intermediate_values(self) -> List[Polynomial]:
        for i in range(self.num_steps):
            if self._values[i] is None:
                self._values[i] = Polynomial(self.field_id, self._fetch_layer(i + 1, want_values=True)[1], VALUES)
        return self._values  # type: ignore[return-value]

    def _h(self) -> int:
        if not getattr(self, "_handle", None):
            raise RuntimeError("the FRIProofPrototype's device memory was freed")
        return self._handle

    def _fetch_layer(self, layer: int, want_nodes: bool = False, want_values: bool = False):
        self._h()
        size = int(lib.hodor_cuda_fri_layer_size(self._handle, layer))
        nodes = np.zeros((size, 32), np.uint8) if want_nodes else None
        values = np.zeros((size, 4), np.uint64) if want_values else None
        check(lib.hodor_cuda_fri_layer(self._handle, layer, nodes.ctypes.data_as(u8p) if want_nodes else None,
                                       _p(values) if want_values else None))
        return nodes, values

    def _query(self, layer: int, natural_index: int) -> TrivialBlake2sIopQuery:
        self._h()
        size = int(lib.hodor_cuda_fri_layer_size(self._handle, layer))
        value = np.zeros(4, np.uint64)
        path = np.zeros((size.bit_length() - 1, 32), np.uint8)
        n = check(lib.hodor_cuda_fri_query(self._handle, layer, C.c_uint64(natural_index), _p(value), path.ctypes.data_as(u8p)))
        assert n == path.shape[0]
        return TrivialBlake2sIopQuery(natural_index, value, [p.tobytes() for p in path])

    def produce_proof(self, iop_values: Optional[Polynomial], natural_first_element_index: int) -> FRIProof:
        """src/fri/query_producer.rs:10-53 (the leaf values are read from HBM, so `iop_values` is unused)."""
        self._h()
        layers = self.num_steps + 1
        idx = np.zeros(2 * layers, np.uint64)
        vals = np.zeros((2 * layers, 4), np.uint64)
        depth0 = self._n.bit_length() - 1
        total = sum(2 * (depth0 - l) for l in range(layers))
        paths = np.zeros((total, 32), np.uint8)
        got = check(lib.hodor_cuda_fri_produce_proof(self._handle, C.c_uint64(natural_first_element_index), _p(idx), _p(vals),
                                                     paths.ctypes.data_as(u8p)))
        assert got == total
        queries, roots, off = [], [], 0
        for layer in range(layers):
            d = depth0 - layer
            for q in range(2):
                queries.append(TrivialBlake2sIopQuery(int(idx[2 * layer + q]), vals[2 * layer + q].copy(),
                                                      [x.tobytes() for x in paths[off:off + d]]))
                off += d
            roots.append(self._roots[layer])
        return FRIProof(queries, roots, self.final_coefficients.copy(), self.initial_degree_plus_one,
                        self.output_coeffs_at_degree_plus_one, self.lde_factor, self.field_id)


class NaiveFriIop:
    """src/fri/mod.rs:63-105"""

    DEGREE = 2

    @staticmethod
    def proof_from_lde(lde_values: Polynomial, lde_factor: int, output_coeffs_at_degree_plus_one: int,
                       worker: Optional[Worker] = None) -> FRIProofPrototype:
        return NaiveFriIop.proof_from_lde_by_values(lde_values, lde_factor, output_coeffs_at_degree_plus_one, worker)

    @staticmethod
    def proof_from_lde_by_values(lde_values: Polynomial, lde_factor: int, output_coeffs_at_degree_plus_one: int,
                                 worker: Optional[Worker] = None) -> FRIProofPrototype:
        if lde_values.form != VALUES:
            raise TypeError("proof_from_lde needs Polynomial<F, Values>")
        ensure_init()
        n = lde_values.size()
        h = lib.hodor_cuda_fri_commit(lde_values.as_ref().ctypes.data, C.c_uint64(n), lde_factor,
                                      output_coeffs_at_degree_plus_one, 0, lde_values.field_id)
        if not h:
            raise_last()  # the code the library recorded: INVALID_ARG, DOMAIN (-> SynthesisError), OOM, CUDA
        return FRIProofPrototype(lde_values.field_id, h, n, lde_factor, output_coeffs_at_degree_plus_one)

    @staticmethod
    def prototype_into_proof(prototype: FRIProofPrototype, iop_values: Optional[Polynomial],
                             natural_first_element_index: int) -> FRIProof:
        return prototype.produce_proof(iop_values, natural_first_element_index)

    @staticmethod
    def verify_proof(proof: FRIProof, natural_element_index: int, expected_value) -> bool:
        """FriIop::verify_proof (src/fri/mod.rs:97-104) -> verify_proof_queries (src/fri/verifier.rs:130-290).
        O(log n) scalar work on the host (hodor_field_* helpers, hashlib), as in the reference; returns
        False for a proof that does not check out and raises SynthesisError where the reference returns
        Err(InvalidValue).  Two properties of the reference are kept as they are: its domain check
        reports every EVEN query index as "not in the LDE domain" ((w^idx)^(N/2) == 1), and it folds the
        last committed layer once more with the challenge the prover dropped before comparing with the
        final coefficients, which is only consistent for output_coeffs_at_degree_plus_one == 1 -- the one
        shape its own tests use (src/fri/mod.rs:436, index 63)."""
        fid, degree = proof.field_id, NaiveFriIop.DEGREE
        two_inv = fld.inverse(fid, fld.add(fid, fld.one(fid), fld.one(fid)))
        domain = Domain.new_for_size(fid, proof.initial_degree_plus_one * proof.lde_factor)
        x = fld.pow_(fid, domain.generator, natural_element_index)
        if not np.array_equal(fld.pow_(fid, x, domain.size), fld.one(fid)):
            raise SynthesisError(-1, "initial challenge value is not in the LDE domain")
        if np.array_equal(fld.pow_(fid, x, domain.size // 2), fld.one(fid)):
            raise SynthesisError(-1, "initial challenge value is not in the LDE domain")
        omega = domain.generator
        omega_inv = fld.inverse(fid, omega)
        expected: Optional[np.ndarray] = None
        domain_size, domain_idx = domain.size, natural_element_index
        if len(proof.queries) % degree != 0:
            raise SynthesisError(-1, "invalid number of queries")
        expected_value = fld.limbs(expected_value)
        for rnd, root in enumerate(proof.roots):
            queries = proof.queries[rnd * degree:(rnd + 1) * degree]
            if len(queries) < degree:
                break  # zip(roots, chunks_exact) stops at the shorter one
            coset = TrivialCombiner.get_coset_for_natural_index(domain_idx, domain_size)
            if len(coset) != degree:
                raise SynthesisError(-1, "invalid coset size")
            if any(q.natural_index() not in coset for q in queries):
                return False
            if rnd == 0:
                for q in queries:
                    if q.natural_index() == natural_element_index and not np.array_equal(q.value(), expected_value):
                        return False
            for c, q in zip(coset, queries):
                if q.tree_index() != TrivialCombiner.natural_index_into_tree_index(c):
                    raise SynthesisError(-1, f"invalid tree index for element at natural index {c}")
                assert q.natural_index() == c, "coset values and produced queries are expected to be sorted!"
            if not all(TrivialBlake2sIOP.verify_query(q, root) for q in queries):
                return False
            challenge = TrivialBlake2sIOP.encode_root_into_challenge(fid, root)
            f_at_omega = queries[0].value()
            if expected is not None:
                if domain_idx not in coset:
                    return False
                hits = [q for q in queries if q.natural_index() == domain_idx]
                if len(hits) != 1 or not np.array_equal(hits[0].value(), expected):
                    return False
            f_at_minus_omega = queries[1].value()
            divisor = fld.pow_(fid, omega_inv, coset[0])
            even = fld.add(fid, f_at_omega, f_at_minus_omega)
            odd = fld.mul(fid, fld.sub(fid, f_at_omega, f_at_minus_omega), divisor)
            expected = fld.mul(fid, fld.add(fid, fld.mul(fid, odd, challenge), even), two_inv)
            domain_idx, domain_size = Domain.index_and_size_for_next_domain(domain_idx, domain_size)
            omega = fld.mul(fid, omega, omega)
            omega_inv = fld.mul(fid, omega_inv, omega_inv)
        point = fld.pow_(fid, omega, domain_idx)
        acc, power = fld.zero(), fld.one(fid)
        for c in np.asarray(proof.final_coefficients, dtype=np.uint64).reshape(-1, 4):
            acc = fld.add(fid, acc, fld.mul(fid, power, c))
            power = fld.mul(fid, power, point)
        assert expected is not None, "is some"
        return bool(np.array_equal(acc, expected))
