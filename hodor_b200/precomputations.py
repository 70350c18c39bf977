"""Setup work of `Prover::new` on the device: the twiddle vectors of
`PrecomputedOmegas` (src/precomputations/mod.rs:7-66) and the ALI inverse divisors
(src/ali/per_register/mod.rs:60-162 dense constraints, :214-227 boundary rows).

Same names and argument meaning as the reference; the vectors are `DevicePolynomial`s (HBM), each built by one
call of the C ABI (`hodor_cuda_precomputed_omegas_dev`, `hodor_cuda_ali_*_inverse_divisor_dev`).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from ._ffi import check, ensure_init, lib
from .device import DevicePolynomial, _ptr, _stream, empty_elems
from .domains import Domain
from .field import _p
from .polynomials import Polynomial


def _log2(size: int) -> int:
    if size <= 0 or size & (size - 1):
        raise ValueError("domain sizes are powers of two")
    return size.bit_length() - 1


class PrecomputedOmegas:
    """`PrecomputedOmegas<F>`: omegas[i] = omega^i (n), coset[i] = g * omega^i (n), omegas_inv[i] = omega^-i (n/2);
    three `Vec<F>` in the reference, three (len, 4) device tensors here."""

    def __init__(self, field_id: int, omegas, coset, omegas_inv):
        self.field_id, self.omegas, self.coset, self.omegas_inv = field_id, omegas, coset, omegas_inv

    @staticmethod
    def new_for_domain(domain: Domain, worker=None) -> "PrecomputedOmegas":
        """src/precomputations/mod.rs:14-66."""
        ensure_init()
        fid, n = domain.field_id, domain.size
        om, co, inv = empty_elems(n), empty_elems(n), empty_elems(n // 2)
        check(lib.hodor_cuda_precomputed_omegas_dev(_ptr(om), _ptr(co), _ptr(inv) if n >= 2 else None, _log2(n), fid, _stream()))
        return PrecomputedOmegas(fid, om, co, inv)

    def coset_values(self) -> DevicePolynomial:
        """`Polynomial::from_values(precomputations.coset.clone())` (src/ali/per_register/mod.rs:302)."""
        return DevicePolynomial(self.field_id, self.coset.clone(), "Values")


@dataclass(frozen=True)
class DenseConstraint:
    """`DenseConstraint` (src/air/mod.rs): the constraint holds on rows start_at .. num_rows - span - 1."""
    start_at: int = 0
    span: int = 1


def inverse_divisor_for_dense_constraint_in_coset(column_domain: Domain, evaluation_domain: Domain, dense_constraint: DenseConstraint,
                                                  num_rows: int, worker=None):
    """src/ali/per_register/mod.rs:60-162 -> (inverse divisor values on g * <evaluation domain>, divisor degree)."""
    ensure_init()
    fid = column_domain.field_id
    out = empty_elems(evaluation_domain.size)
    degree = C.c_uint64(0)
    check(lib.hodor_cuda_ali_dense_inverse_divisor_dev(_ptr(out), _log2(column_domain.size), _log2(evaluation_domain.size),
                                                       C.c_uint64(dense_constraint.start_at), C.c_uint64(dense_constraint.span),
                                                       C.c_uint64(num_rows), C.byref(degree), fid, _stream()))
    return DevicePolynomial(fid, out, "Values"), int(degree.value)


def boundary_constraint_inverse_divisor(column_domain: Domain, constraints_domain: Domain, row: int, worker=None) -> DevicePolynomial:
    """src/ali/per_register/mod.rs:214-227: 1 / (X - omega^row) on the coset of the constraints domain."""
    ensure_init()
    fid = column_domain.field_id
    out = empty_elems(constraints_domain.size)
    check(lib.hodor_cuda_ali_boundary_inverse_divisor_dev(_ptr(out), _log2(column_domain.size), _log2(constraints_domain.size),
                                                          C.c_uint64(row), fid, _stream()))
    return DevicePolynomial(fid, out, "Values")


# ---- host-vector forms: what a Rust caller holding Vec<F> binds (rust/src/cuda/ali.rs) ---------------------------------
def precomputed_omegas_host(domain: Domain):
    """-> (omegas, coset, omegas_inv) as numpy (n, 4) / (n/2, 4) uint64 arrays."""
    ensure_init()
    n = domain.size
    om, co, inv = np.zeros((n, 4), np.uint64), np.zeros((n, 4), np.uint64), np.zeros((n // 2, 4), np.uint64)
    check(lib.hodor_cuda_precomputed_omegas(_p(om), _p(co), _p(inv) if n >= 2 else None, _log2(n), domain.field_id))
    return om, co, inv


def inverse_divisor_for_dense_constraint_in_coset_host(column_domain: Domain, evaluation_domain: Domain,
                                                       dense_constraint: DenseConstraint, num_rows: int, worker=None):
    ensure_init()
    out = np.zeros((evaluation_domain.size, 4), np.uint64)
    degree = C.c_uint64(0)
    check(lib.hodor_cuda_ali_dense_inverse_divisor(_p(out), _log2(column_domain.size), _log2(evaluation_domain.size),
                                                   C.c_uint64(dense_constraint.start_at), C.c_uint64(dense_constraint.span),
                                                   C.c_uint64(num_rows), C.byref(degree), column_domain.field_id))
    return Polynomial.from_values(column_domain.field_id, out), int(degree.value)


def boundary_constraint_inverse_divisor_host(column_domain: Domain, constraints_domain: Domain, row: int, worker=None) -> Polynomial:
    ensure_init()
    out = np.zeros((constraints_domain.size, 4), np.uint64)
    check(lib.hodor_cuda_ali_boundary_inverse_divisor(_p(out), _log2(column_domain.size), _log2(constraints_domain.size),
                                                      C.c_uint64(row), column_domain.field_id))
    return Polynomial.from_values(column_domain.field_id, out)
