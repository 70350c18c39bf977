"""Callers of the multi-GPU entry points of the C ABI (`hodor_cuda_comm_init`, `hodor_cuda_ntt_sharded`,
`hodor_cuda_lde_fri_sharded`).  The orchestration -- local transforms, NCCL send/recv, folds, subtrees -- is
C++ inside libhodor_b200.so; torch.distributed is used for exactly one thing, shipping NCCL's 128-byte
unique id from rank 0 to the other ranks at start-up (a Rust or C++ host would use MPI or a file).

`hodor_b200/sharded.py` and `sharded_fri.py` keep the same algorithms as torch.distributed programs with a
pluggable compute backend: they are what the world_size-2 gloo tests on the CPU exercise.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Tuple

import numpy as np
import torch

from . import field as fld
from ._ffi import check, ensure_init, lib, u8p
from .field import _p

_comm: Tuple[int, int] | None = None


def comm_init(group=None) -> Tuple[int, int]:
    """Bind the library's communicator to this process' rank.  Without torch.distributed: world 1."""
    global _comm
    import torch.distributed as dist

    ensure_init()
    if _comm is not None:
        return _comm
    if dist.is_available() and dist.is_initialized():
        world, rank = dist.get_world_size(group), dist.get_rank(group)
    else:
        world, rank = 1, 0
    uid = torch.zeros(128, dtype=torch.uint8)
    if world > 1:
        if rank == 0:
            buf = np.zeros(128, np.uint8)
            check(lib.hodor_cuda_comm_unique_id(buf.ctypes.data_as(u8p)))
            uid = torch.from_numpy(buf)
        backend = dist.get_backend(group)
        if backend == "nccl":
            d = uid.cuda()
            dist.broadcast(d, src=0, group=group)
            uid = d.cpu()
        else:
            dist.broadcast(uid, src=0, group=group)
    idb = uid.numpy().copy()
    torch.cuda.synchronize()
    check(lib.hodor_cuda_comm_init(rank, world, idb.ctypes.data_as(u8p)))
    _comm = (rank, world)
    return _comm


def comm_destroy() -> None:
    global _comm
    lib.hodor_cuda_comm_destroy()
    _comm = None


def bytes_sent() -> int:
    """Payload this rank pushed through NCCL since comm_init."""
    n = C.c_uint64(0)
    check(lib.hodor_cuda_comm_info(None, None, C.byref(n), None))
    return int(n.value)


def bytes_peer_stored() -> int:
    """Payload this rank's kernels stored straight into other ranks' buffers over NVLink since comm_init."""
    n = C.c_uint64(0)
    check(lib.hodor_cuda_comm_info(None, None, None, C.byref(n)))
    return int(n.value)


def ntt_sharded(local_in: torch.Tensor, log_n: int, omega, field_id: int, out: torch.Tensor | None = None) -> torch.Tensor:
    """Forward NTT of length 2^log_n over the ranks.  `local_in`: this rank's cyclic slice a[j*G + rank].
    Returns this rank's part of the natural-order result (hodor_b200.sharded.gather_output states the layout)."""
    comm_init()
    if out is None:
        out = torch.empty_like(local_in)
    check(lib.hodor_cuda_ntt_sharded(local_in.data_ptr(), out.data_ptr(), log_n, _p(fld.limbs(omega)), field_id,
                                     torch.cuda.current_stream().cuda_stream))
    return out


def lde_fri_sharded(d_coeffs: torch.Tensor, log_n: int, log_factor: int, coset: bool, out_coeffs: int,
                    field_id: int) -> Tuple[List[bytes], np.ndarray, np.ndarray]:
    """One (coset) LDE + FRI commit chain over all ranks; `d_coeffs` replicated.  Returns (roots, challenges,
    final coefficients), identical on every rank and to the single-GPU chain."""
    comm_init()
    steps = log_n - (out_coeffs.bit_length() - 1)
    roots = np.zeros((max(steps, 0) + 1, 32), np.uint8)
    chals = np.zeros((max(steps, 1), 4), np.uint64)
    fin = np.zeros((out_coeffs, 4), np.uint64)
    torch.cuda.current_stream().synchronize()  # the chain runs on the library's stream
    got = check(lib.hodor_cuda_lde_fri_sharded(d_coeffs.data_ptr(), log_n, log_factor, int(coset), out_coeffs,
                                               roots.ctypes.data_as(u8p), _p(chals), _p(fin), field_id))
    assert got == steps
    return [r.tobytes() for r in roots], chals[:steps], fin
