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
