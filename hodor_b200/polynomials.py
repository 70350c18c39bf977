"""Mirror of the reference's typed Polynomial (src/polynomials/mod.rs:26-34 and the transform
methods :343-352, :611-638, :773-815).  Data lives in host numpy arrays of shape (n, 4) uint64 --
the same bytes as the reference's Vec<F> -- and every transform is one call into the C ABI.

`Worker` is accepted wherever the reference takes one and ignored: the CUDA grid replaces the
thread pool (src/fft/multicore.rs).
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np

from . import field as fld
from ._ffi import check, ensure_init, lib
from .domains import Domain
from .field import _p

COEFFICIENTS = "Coefficients"
VALUES = "Values"


class Worker:
    """Placeholder for src/fft/multicore.rs:17-103; carries no threads."""

    def __init__(self, cpus: Optional[int] = None):
        self.cpus = cpus or 1


def _as_elems(a) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.uint64)
    if a.ndim == 1:
        a = a.reshape(-1, 4)
    assert a.ndim == 2 and a.shape[1] == 4
    return a


class Polynomial:
    def __init__(self, field_id: int, data: np.ndarray, form: str):
        data = _as_elems(data)
        dom = Domain.new_for_size(field_id, max(1, data.shape[0]))  # from_values / from_coeffs (:713-742)
        if dom.size != data.shape[0]:
            padded = np.zeros((dom.size, 4), np.uint64)  # values.resize(m, F::zero())
            padded[: data.shape[0]] = data
            data = padded
        self.field_id = field_id
        self.coeffs = data
        self.form = form
        self.exp = dom.power_of_two
        self.omega = dom.generator
        self.omegainv = fld.inverse(field_id, self.omega)
        self.geninv = fld.inverse(field_id, fld.multiplicative_generator(field_id))
        self.minv = fld.inverse(field_id, fld.from_repr(field_id, dom.size))

    # ---- constructors -------------------------------------------------------------------------
    @staticmethod
    def from_coeffs(field_id: int, coeffs) -> "Polynomial":
        return Polynomial(field_id, np.array(coeffs, dtype=np.uint64, copy=True), COEFFICIENTS)

    @staticmethod
    def from_values(field_id: int, values) -> "Polynomial":
        return Polynomial(field_id, np.array(values, dtype=np.uint64, copy=True), VALUES)

    # ---- accessors ----------------------------------------------------------------------------
    def size(self) -> int:
        return self.coeffs.shape[0]

    def as_ref(self) -> np.ndarray:
        return self.coeffs

    def into_coeffs(self) -> np.ndarray:
        return self.coeffs

    def clone(self) -> "Polynomial":
        return Polynomial(self.field_id, self.coeffs.copy(), self.form)

    def __eq__(self, other) -> bool:
        return (isinstance(other, Polynomial) and self.field_id == other.field_id and self.form == other.form
                and np.array_equal(self.coeffs, other.coeffs))

    def _need(self, form: str):
        if self.form != form:
            raise TypeError(f"operation needs Polynomial<F, {form}>, have {self.form}")

    def _retag(self, form: str) -> "Polynomial":
        self.form = form
        return self

    # ---- transforms (consume self, like the reference) -----------------------------------------
    def distribute_powers(self, worker: Optional[Worker], g) -> None:
        """:53-56 -> distribute_powers (src/fft/mod.rs:110-123)."""
        ensure_init()
        check(lib.hodor_cuda_distribute_powers(_p(self.coeffs), C.c_uint64(self.size()), _p(fld.limbs(g)), self.field_id))

    def fft(self, worker: Optional[Worker] = None) -> "Polynomial":
        """:611-624"""
        self._need(COEFFICIENTS)
        ensure_init()
        check(lib.hodor_cuda_fft(_p(self.coeffs), self.exp, 0, self.field_id))
        return self._retag(VALUES)

    def coset_fft(self, worker: Optional[Worker] = None) -> "Polynomial":
        """:626-631"""
        self._need(COEFFICIENTS)
        ensure_init()
        check(lib.hodor_cuda_fft(_p(self.coeffs), self.exp, 1, self.field_id))
        return self._retag(VALUES)

    def coset_fft_for_generator(self, worker: Optional[Worker], gen) -> "Polynomial":
        """:633-638"""
        self._need(COEFFICIENTS)
        self.distribute_powers(worker, gen)
        return self.fft(worker)

    def ifft(self, worker: Optional[Worker] = None) -> "Polynomial":
        """:773-798"""
        self._need(VALUES)
        ensure_init()
        check(lib.hodor_cuda_ifft(_p(self.coeffs), self.exp, 0, self.field_id))
        return self._retag(COEFFICIENTS)

    def icoset_fft(self, worker: Optional[Worker] = None) -> "Polynomial":
        """:800-807"""
        self._need(VALUES)
        ensure_init()
        check(lib.hodor_cuda_ifft(_p(self.coeffs), self.exp, 1, self.field_id))
        return self._retag(COEFFICIENTS)

    def icoset_fft_for_generator(self, worker: Optional[Worker], geninv) -> "Polynomial":
        """:809-815"""
        res = self.ifft(worker)
        res.distribute_powers(worker, geninv)
        return res

    def _lde(self, factor: int, coset: bool) -> "Polynomial":
        self._need(COEFFICIENTS)
        if factor < 1 or factor & (factor - 1):
            raise AssertionError("assert!(factor.is_power_of_two())")  # :434 / :560
        ensure_init()
        log_f = factor.bit_length() - 1
        Domain.new_for_size(self.field_id, self.size() * factor)  # Err(SynthesisError) as in :435 / :561
        out = np.zeros((self.size() * factor, 4), np.uint64)
        check(lib.hodor_cuda_lde(_p(self.coeffs), self.exp, log_f, int(coset), _p(out), self.field_id))
        return Polynomial(self.field_id, out, VALUES)

    def lde(self, worker: Optional[Worker], factor: int) -> "Polynomial":
        """:343-346 -> lde_using_multiple_cosets :418-482"""
        return self._lde(factor, coset=False)

    def coset_lde(self, worker: Optional[Worker], factor: int) -> "Polynomial":
        """:348-352 -> coset_lde_using_multiple_cosets :544-609"""
        return self._lde(factor, coset=True)

    # the reference keeps several algorithmic variants that must agree; here they are one kernel
    lde_using_multiple_cosets = lde
    coset_lde_using_multiple_cosets = coset_lde
    filtering_lde = lde
    coset_filtering_lde = coset_lde

    # ---- elementwise (:640-683, :817-887) --------------------------------------------------------
    def _ew(self, op: int, other: np.ndarray) -> None:
        ensure_init()
        other = _as_elems(other)
        n = other.shape[0] if op != 3 else self.size()
        assert self.size() >= n
        check(lib.hodor_cuda_elementwise(op, _p(self.coeffs), _p(other), _p(self.coeffs), C.c_uint64(n), self.field_id))

    def add_assign(self, worker: Optional[Worker], other: "Polynomial") -> None:
        self._ew(1, other.coeffs)

    def sub_assign(self, worker: Optional[Worker], other: "Polynomial") -> None:
        self._ew(2, other.coeffs)

    def mul_assign(self, worker: Optional[Worker], other: "Polynomial") -> None:
        self._need(VALUES)
        assert self.size() == other.size()
        self._ew(0, other.coeffs)

    def scale(self, worker: Optional[Worker], g) -> None:
        self._ew(3, fld.limbs(g).reshape(1, 4))

    def _op(self, op: int, other=None, scalar=None, exp: int = 0) -> None:
        ensure_init()
        n = self.size() if other is None else _as_elems(other).shape[0]
        assert self.size() >= n
        b = None if other is None else _p(_as_elems(other))
        s = None if scalar is None else _p(fld.limbs(scalar))
        check(lib.hodor_cuda_poly_op(op, _p(self.coeffs), b, s, C.c_uint64(exp), _p(self.coeffs), C.c_uint64(n), self.field_id))

    def add_assign_scaled(self, worker: Optional[Worker], other: "Polynomial", scaling) -> None:
        """:654-669 / :843-858"""
        self._op(4, other.coeffs, scaling)

    def add_constant(self, worker: Optional[Worker], constant) -> None:
        """:831-841"""
        self._need(VALUES)
        self._op(5, scalar=constant)

    def negate(self, worker: Optional[Worker] = None) -> None:
        """:72-83"""
        self._op(6)

    def square(self, worker: Optional[Worker] = None) -> None:
        """:760-771"""
        self._need(VALUES)
        self._op(7)

    def pow(self, worker: Optional[Worker], exp: int) -> None:
        """:744-758"""
        self._need(VALUES)
        self._op(7 if exp == 2 else 8, exp=exp)

    # ---- batch inversion (:889-954) and point evaluation (:685-711) ------------------------------
    def batch_inversion(self, worker: Optional[Worker] = None) -> None:
        """Every value replaced by its inverse; SynthesisError (vector untouched) if one is zero."""
        self._need(VALUES)
        ensure_init()
        check(lib.hodor_cuda_batch_inversion(_p(self.coeffs), C.c_uint64(self.size()), self.field_id))

    def evaluate_at(self, worker: Optional[Worker], g) -> np.ndarray:
        """sum_j coeffs[j] * g^j as 4 Montgomery limbs."""
        self._need(COEFFICIENTS)
        ensure_init()
        out = np.zeros(4, np.uint64)
        check(lib.hodor_cuda_evaluate_at(_p(self.coeffs), C.c_uint64(self.size()), _p(fld.limbs(g)), _p(out), self.field_id))
        return out


def lde_batch(polys, worker: Optional[Worker], factor: int, coset: bool = False):
    """`[w.lde(worker, factor) for w in polys]` (src/prover/mod.rs:73-76) as ONE pipelined call:
    copies of neighbouring polynomials overlap the transform of the current one."""
    polys = list(polys)
    if not polys:
        return []
    fid, n = polys[0].field_id, polys[0].size()
    for q in polys:
        q._need(COEFFICIENTS)
        assert q.field_id == fid and q.size() == n, "lde_batch: polynomials must share field and size"
    if factor < 1 or factor & (factor - 1):
        raise AssertionError("assert!(factor.is_power_of_two())")
    ensure_init()
    Domain.new_for_size(fid, n * factor)
    outs = [np.zeros((n * factor, 4), np.uint64) for _ in polys]
    u64p = C.POINTER(C.c_uint64)
    ins_arr = (u64p * len(polys))(*[_p(q.coeffs) for q in polys])
    outs_arr = (u64p * len(polys))(*[_p(o) for o in outs])
    check(lib.hodor_cuda_lde_batch(ins_arr, outs_arr, len(polys), polys[0].exp, factor.bit_length() - 1, int(coset), fid))
    return [Polynomial(fid, o, VALUES) for o in outs]
