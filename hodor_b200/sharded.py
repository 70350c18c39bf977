"""Four-step (Bailey) NTT sharded over the G = 2^log_g GPUs of one box: one process per GPU,
torch.distributed for the plumbing, NCCL all-to-all over NVLink for the single transpose.

The reference's own parallel_fft (src/fft/fft.rs:68-125) is this decomposition over CPU threads
with the first factor done as a naive O(C) DFT; here the factors are n = m * G:

  rank g holds  a_g[j1] = a[j1 * G + g]                      (cyclic distribution of the input)
  step A (local, hodor_cuda_ntt_shard_cols_dev):
        B_g[k1] = omega^(g * k1) * sum_j1 a_g[j1] * (omega^G)^(j1 * k1)        k1 < m
  transpose (the only communication): all_to_all of m/G-element chunks; rank h then holds
        M_h[g][k] = B_g[h * m/G + k]                                           k < m/G
  step B (local, hodor_cuda_ntt_shard_rows_dev): G-point DFT over g
        out_h[k2 * m/G + k] = sum_g M_h[g][k] * (omega^m)^(g * k2)
                            = A[k2 * m + h * m/G + k]
  so rank h owns, of every length-m block of the natural-order output, its h-th m/G-chunk
  (block-cyclic with block n / G^2).  `scatter_input` / `gather_output` state the contract in code.

Each rank sends (G-1)/G of its n/G elements once; nothing else crosses NVLink.
"""
from __future__ import annotations

from typing import List, Optional, Protocol

import numpy as np
import torch
import torch.distributed as dist


class ShardBackend(Protocol):
    def shard_cols(self, src: torch.Tensor, log_n: int, log_g: int, rank: int, omega, field_id: int) -> torch.Tensor: ...

    def shard_rows(self, src: torch.Tensor, log_n: int, log_g: int, rank: int, omega, field_id: int) -> torch.Tensor: ...


class CudaBackend:
    """The product path: hand-written CUDA through the C ABI, on torch's current stream."""

    def shard_cols(self, src, log_n, log_g, rank, omega, field_id):
        import ctypes as C  # noqa: F401
        from . import device as dev
        from . import field as fld
        from ._ffi import check, ensure_init, lib
        from .field import _p

        ensure_init()
        dst = torch.empty_like(src)
        check(lib.hodor_cuda_ntt_shard_cols_dev(src.data_ptr(), dst.data_ptr(), log_n, log_g, rank, _p(fld.limbs(omega)),
                                                field_id, dev._stream()))
        return dst

    def shard_rows(self, src, log_n, log_g, rank, omega, field_id):
        from . import device as dev
        from . import field as fld
        from ._ffi import check, ensure_init, lib
        from .field import _p

        ensure_init()
        dst = torch.empty_like(src)
        check(lib.hodor_cuda_ntt_shard_rows_dev(src.data_ptr(), dst.data_ptr(), log_n, log_g, rank, _p(fld.limbs(omega)),
                                                field_id, dev._stream()))
        return dst


def scatter_input(a: np.ndarray, world: int, rank: int) -> np.ndarray:
    """The slice of a natural-order vector that rank `rank` must hold on entry."""
    return np.ascontiguousarray(a.reshape(-1, 4)[rank::world])


def gather_output(parts: List[np.ndarray]) -> np.ndarray:
    """Reassemble the natural-order result from every rank's output."""
    world = len(parts)
    m = parts[0].reshape(-1, 4).shape[0]
    chunk = m // world
    out = np.zeros((m * world, 4), np.uint64)
    for h, part in enumerate(parts):
        p = part.reshape(world, chunk, 4)
        for k2 in range(world):
            out[k2 * m + h * chunk : k2 * m + (h + 1) * chunk] = p[k2]
    return out


def ntt_sharded(local_in: torch.Tensor, log_n: int, omega, field_id: int, group: Optional[dist.ProcessGroup] = None,
                backend: Optional[ShardBackend] = None) -> torch.Tensor:
    """Forward NTT of length 2^log_n distributed over the ranks of `group` (see module docstring
    for the input / output distribution).  `local_in`: (2^log_n / world, 4) int64."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    log_g = world.bit_length() - 1
    if world != 1 << log_g:
        raise ValueError("world size must be a power of two")
    if log_n < 2 * log_g:
        raise ValueError("transform too short to shard: need log_n >= 2 * log2(world)")
    if local_in.shape[0] != (1 << log_n) >> log_g:
        raise ValueError("local_in has the wrong length for this rank")
    backend = backend or CudaBackend()
    b = backend.shard_cols(local_in, log_n, log_g, rank, omega, field_id)
    if world == 1:
        recv = b
    else:
        recv = torch.empty_like(b)
        dist.all_to_all_single(recv, b, group=group)  # equal splits: chunk h of B_g -> rank h
    return backend.shard_rows(recv, log_n, log_g, rank, omega, field_id)
