"""Device-resident entry points on torch CUDA tensors (dtype int64/uint64, shape (n, 4)).

torch is only the owner of HBM buffers and streams here; every call goes straight to the `_dev`
functions of the C ABI on torch's current stream, so torch.cuda.Event timing sees the kernels.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import field as fld
from ._ffi import check, ensure_init, lib
from .field import _p


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: torch.Tensor) -> int:
    assert t.is_cuda and t.is_contiguous()
    return t.data_ptr()


def empty_elems(n: int, device=None) -> torch.Tensor:
    return torch.empty((n, 4), dtype=torch.int64, device=device or torch.device("cuda", torch.cuda.current_device()))


def to_device(a: np.ndarray, device=None) -> torch.Tensor:
    return torch.from_numpy(np.ascontiguousarray(a).view(np.int64)).to(device or torch.device("cuda", torch.cuda.current_device()))


def to_host(t: torch.Tensor) -> np.ndarray:
    return t.cpu().numpy().view(np.uint64)


def ntt(src: torch.Tensor, dst: torch.Tensor, log_n: int, omega, field_id: int) -> None:
    ensure_init()
    check(lib.hodor_cuda_ntt_dev(_ptr(src), _ptr(dst), log_n, _p(fld.limbs(omega)), field_id, _stream()))


def fft(src, dst, log_n: int, coset: bool, field_id: int) -> None:
    ensure_init()
    check(lib.hodor_cuda_fft_dev(_ptr(src), _ptr(dst), log_n, int(coset), field_id, _stream()))


def ifft(src, dst, log_n: int, coset: bool, field_id: int) -> None:
    ensure_init()
    check(lib.hodor_cuda_ifft_dev(_ptr(src), _ptr(dst), log_n, int(coset), field_id, _stream()))


def lde(coeffs, log_n: int, log_factor: int, coset: bool, out, field_id: int) -> None:
    ensure_init()
    check(lib.hodor_cuda_lde_dev(_ptr(coeffs), log_n, log_factor, int(coset), _ptr(out), field_id, _stream()))


def merkle_build(leaves, n: int, nodes, field_id: int, root=None, challenge=None) -> None:
    ensure_init()
    check(lib.hodor_cuda_merkle_build_dev(_ptr(leaves), C.c_uint64(n), _ptr(nodes), _ptr(root) if root is not None else None,
                                          _ptr(challenge) if challenge is not None else None, field_id, _stream()))


def fri_fold(src, n: int, initial_domain_size: int, layer: int, challenge, dst, field_id: int) -> None:
    ensure_init()
    check(lib.hodor_cuda_fri_fold_dev(_ptr(src), C.c_uint64(n), C.c_uint64(initial_domain_size), layer, _ptr(challenge),
                                      _ptr(dst), field_id, _stream()))


def fri_commit(lde_values: torch.Tensor, lde_factor: int, out_coeffs: int, field_id: int):
    """Whole commit chain on a device-resident LDE.  Runs on the library's own stream; returns a
    FRIProofPrototype that borrows `lde_values` (keep the tensor alive)."""
    from .fri import FRIProofPrototype
    from ._ffi import HodorError, last_error

    ensure_init()
    torch.cuda.current_stream().synchronize()
    n = lde_values.shape[0]
    h = lib.hodor_cuda_fri_commit(_ptr(lde_values), C.c_uint64(n), lde_factor, out_coeffs, 1, field_id)
    if not h:
        raise HodorError(-1, last_error())
    proto = FRIProofPrototype(field_id, h, n, lde_factor, out_coeffs)
    proto._keepalive = lde_values
    return proto


def launch_count() -> int:
    return int(lib.hodor_cuda_launch_count())


class DevicePolynomial:
    """`Polynomial<F, Coefficients | Values>` (src/polynomials/mod.rs:26-34) whose vector stays in HBM:
    the same methods as `hodor_b200.Polynomial`, each one `_dev` call on torch's current stream, so a
    caller chaining transforms and elementwise steps (what the ALI / DEEP phases of src/ali do between
    the LDEs and FRI) pays PCIe only where it asks for host data."""

    def __init__(self, field_id: int, data, form: str):
        from .domains import Domain

        if not isinstance(data, torch.Tensor):
            data = to_device(np.ascontiguousarray(data, dtype=np.uint64).reshape(-1, 4))
        dom = Domain.new_for_size(field_id, max(1, data.shape[0]))  # from_values / from_coeffs (:713-742)
        if dom.size != data.shape[0]:
            padded = torch.zeros((dom.size, 4), dtype=torch.int64, device=data.device)
            padded[: data.shape[0]] = data
            data = padded
        self.field_id, self.coeffs, self.form, self.exp = field_id, data.contiguous(), form, dom.power_of_two

    @staticmethod
    def from_coeffs(field_id: int, coeffs) -> "DevicePolynomial":
        return DevicePolynomial(field_id, coeffs, "Coefficients")

    @staticmethod
    def from_values(field_id: int, values) -> "DevicePolynomial":
        return DevicePolynomial(field_id, values, "Values")

    def size(self) -> int:
        return self.coeffs.shape[0]

    def to_host(self) -> np.ndarray:
        return to_host(self.coeffs)

    def clone(self) -> "DevicePolynomial":
        return DevicePolynomial(self.field_id, self.coeffs.clone(), self.form)

    def _need(self, form: str) -> None:
        if self.form != form:
            raise TypeError(f"operation needs Polynomial<F, {form}>, have {self.form}")

    # ---- transforms (in place, like the reference's consuming methods) ---------------------------
    def fft(self, worker=None) -> "DevicePolynomial":
        self._need("Coefficients")
        fft(self.coeffs, self.coeffs, self.exp, False, self.field_id)
        self.form = "Values"
        return self

    def coset_fft(self, worker=None) -> "DevicePolynomial":
        self._need("Coefficients")
        fft(self.coeffs, self.coeffs, self.exp, True, self.field_id)
        self.form = "Values"
        return self

    def ifft(self, worker=None) -> "DevicePolynomial":
        self._need("Values")
        ifft(self.coeffs, self.coeffs, self.exp, False, self.field_id)
        self.form = "Coefficients"
        return self

    def icoset_fft(self, worker=None) -> "DevicePolynomial":
        self._need("Values")
        ifft(self.coeffs, self.coeffs, self.exp, True, self.field_id)
        self.form = "Coefficients"
        return self

    def _lde(self, factor: int, coset: bool) -> "DevicePolynomial":
        from .domains import Domain

        self._need("Coefficients")
        if factor < 1 or factor & (factor - 1):
            raise AssertionError("assert!(factor.is_power_of_two())")
        Domain.new_for_size(self.field_id, self.size() * factor)
        out = empty_elems(self.size() * factor, self.coeffs.device)
        lde(self.coeffs, self.exp, factor.bit_length() - 1, coset, out, self.field_id)
        return DevicePolynomial(self.field_id, out, "Values")

    def lde(self, worker, factor: int) -> "DevicePolynomial":
        return self._lde(factor, False)

    def coset_lde(self, worker, factor: int) -> "DevicePolynomial":
        return self._lde(factor, True)

    @staticmethod
    def filled(field_id: int, size: int, value, form: str = "Values") -> "DevicePolynomial":
        """A vector of `size` copies of one element (vec![value; size])."""
        one = to_device(fld.limbs(value).reshape(1, 4))
        return DevicePolynomial(field_id, one.expand(size, 4).contiguous(), form)

    def distribute_powers(self, worker, g) -> None:
        """:54-57 -> src/fft/mod.rs:110-123: a[j] <- a[j] * g^j."""
        ensure_init()
        check(lib.hodor_cuda_distribute_powers_dev(_ptr(self.coeffs), C.c_uint64(self.size()), _p(fld.limbs(g)), self.field_id,
                                                   _stream()))

    # ---- elementwise ---------------------------------------------------------------------------
    def _op(self, op: int, other=None, scalar=None, exp: int = 0) -> None:
        ensure_init()
        n = self.size() if (other is None or op == 3) else other.shape[0]  # op 3: `other` is the one-element scalar
        assert self.size() >= n
        s = None if scalar is None else _p(fld.limbs(scalar))
        check(lib.hodor_cuda_poly_op_dev(op, _ptr(self.coeffs), _ptr(other) if other is not None else None, s,
                                         C.c_uint64(exp), _ptr(self.coeffs), C.c_uint64(n), self.field_id, _stream()))

    def add_assign(self, worker, other: "DevicePolynomial") -> None:
        self._op(1, other.coeffs)

    def sub_assign(self, worker, other: "DevicePolynomial") -> None:
        self._op(2, other.coeffs)

    def mul_assign(self, worker, other: "DevicePolynomial") -> None:
        self._need("Values")
        assert self.size() == other.size()
        self._op(0, other.coeffs)

    def scale(self, worker, g) -> None:
        self._op(3, to_device(fld.limbs(g).reshape(1, 4), self.coeffs.device))

    def add_assign_scaled(self, worker, other: "DevicePolynomial", scaling) -> None:
        self._op(4, other.coeffs, scaling)

    def add_constant(self, worker, constant) -> None:
        self._need("Values")
        self._op(5, scalar=constant)

    def negate(self, worker=None) -> None:
        self._op(6)

    def square(self, worker=None) -> None:
        self._need("Values")
        self._op(7)

    def pow(self, worker, exp: int) -> None:
        self._need("Values")
        self._op(7 if exp == 2 else 8, exp=exp)

    def batch_inversion(self, worker=None) -> None:
        """:889-954.  Reads one status word back (the reference's Err on a zero value)."""
        from ._ffi import ERR_NOT_INVERTIBLE, SynthesisError

        self._need("Values")
        ensure_init()
        status = torch.zeros(1, dtype=torch.int32, device=self.coeffs.device)
        check(lib.hodor_cuda_batch_inversion_dev(_ptr(self.coeffs), C.c_uint64(self.size()), status.data_ptr(), self.field_id,
                                                 _stream()))
        if int(status.item()) != 0:
            raise SynthesisError(ERR_NOT_INVERTIBLE, "batch_inversion: the vector contains zero")

    def evaluate_at(self, worker, g) -> np.ndarray:
        """:685-711; returns the 4 Montgomery limbs on the host."""
        self._need("Coefficients")
        ensure_init()
        out = empty_elems(1, self.coeffs.device)
        check(lib.hodor_cuda_evaluate_at_dev(_ptr(self.coeffs), C.c_uint64(self.size()), _p(fld.limbs(g)), _ptr(out),
                                             self.field_id, _stream()))
        return to_host(out)[0]
