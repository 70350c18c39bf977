"""ctypes binding of libhodor_b200.so (include/hodor_b200.h).

There is no Python or CPU implementation behind this module: if the shared library is missing the
import fails, and if no sm_100 GPU is present every compute call raises HodorError.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("HODOR_B200_LIB") or os.path.join(_HERE, "libhodor_b200.so")  # override: A/B builds

OK = 0
ERR_INVALID_ARG = -1
ERR_DOMAIN = -2
ERR_CUDA = -3
ERR_OOM = -4
ERR_NOT_A_ROOT = -5
ERR_NOT_INVERTIBLE = -6

BLS12_381_FR = 0  # what the reference's src/bn256.rs declares
BN254_FR = 1
STARK252 = 2
FIELD_NAMES = {BLS12_381_FR: "bls12_381_fr", BN254_FR: "bn254_fr", STARK252: "stark252"}


class HodorError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"hodor_b200 error {code}: {message}")
        self.code = code


class SynthesisError(HodorError):
    """Mirror of the reference's SynthesisError::Error (src/lib.rs:40-46) for domain failures."""


if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
        "(or `make -C hodor_b200/csrc`).  hodor_b200 has no fallback implementation."
    )

lib = C.CDLL(LIB_PATH)

u64p = C.POINTER(C.c_uint64)
u8p = C.POINTER(C.c_uint8)
u32p = C.POINTER(C.c_uint32)
vp = C.c_void_p

_SIGS = {
    "hodor_cuda_device_count": (C.c_int, []),
    "hodor_cuda_init": (C.c_int, [C.c_int]),
    "hodor_cuda_shutdown": (None, []),
    "hodor_cuda_last_error": (C.c_char_p, []),
    "hodor_cuda_last_error_code": (C.c_int, []),
    "hodor_cuda_workspace_bytes": (C.c_size_t, []),
    "hodor_cuda_trim": (C.c_int, []),
    "hodor_cuda_launch_count": (C.c_uint64, []),
    "hodor_cuda_selftest_mul_pre": (C.c_int, [C.c_int]),
    "hodor_cuda_profile_begin": (C.c_int, []),
    "hodor_cuda_profile_end": (C.c_int, [C.c_char_p, C.c_size_t]),
    "hodor_field_constants": (C.c_int, [C.c_int, u64p, u64p, u64p, u64p, u32p, u32p, u32p]),
    "hodor_domain_generator": (C.c_int, [C.c_int, C.c_uint32, u64p]),
    "hodor_field_mul": (C.c_int, [C.c_int, u64p, u64p, u64p]),
    "hodor_field_add": (C.c_int, [C.c_int, u64p, u64p, u64p]),
    "hodor_field_sub": (C.c_int, [C.c_int, u64p, u64p, u64p]),
    "hodor_field_pow": (C.c_int, [C.c_int, u64p, C.c_uint64, u64p]),
    "hodor_field_inverse": (C.c_int, [C.c_int, u64p, u64p]),
    "hodor_field_from_repr": (C.c_int, [C.c_int, u64p, u64p]),
    "hodor_field_into_repr": (C.c_int, [C.c_int, u64p, u64p]),
    "hodor_root_to_challenge": (C.c_int, [u8p, u64p, C.c_int]),
    "hodor_hash_leaf": (C.c_int, [u64p, u8p]),
    "hodor_hash_node": (C.c_int, [u8p, u8p, u8p]),
    "hodor_cuda_malloc": (vp, [C.c_size_t]),
    "hodor_cuda_free": (None, [vp]),
    "hodor_cuda_host_alloc": (vp, [C.c_size_t]),
    "hodor_cuda_host_free": (None, [vp]),
    "hodor_cuda_memcpy_h2d": (C.c_int, [vp, vp, C.c_size_t, vp]),
    "hodor_cuda_memcpy_d2h": (C.c_int, [vp, vp, C.c_size_t, vp]),
    "hodor_cuda_stream_synchronize": (C.c_int, [vp]),
    "hodor_cuda_ntt": (C.c_int, [u64p, C.c_uint32, u64p, C.c_int]),
    "hodor_cuda_fft": (C.c_int, [u64p, C.c_uint32, C.c_int, C.c_int]),
    "hodor_cuda_ifft": (C.c_int, [u64p, C.c_uint32, C.c_int, C.c_int]),
    "hodor_cuda_distribute_powers": (C.c_int, [u64p, C.c_uint64, u64p, C.c_int]),
    "hodor_cuda_lde": (C.c_int, [u64p, C.c_uint32, C.c_uint32, C.c_int, u64p, C.c_int]),
    "hodor_cuda_lde_batch": (C.c_int, [C.POINTER(u64p), C.POINTER(u64p), C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, C.c_int]),
    "hodor_cuda_elementwise": (C.c_int, [C.c_int, u64p, u64p, u64p, C.c_uint64, C.c_int]),
    "hodor_cuda_poly_op": (C.c_int, [C.c_int, u64p, u64p, u64p, C.c_uint64, u64p, C.c_uint64, C.c_int]),
    "hodor_cuda_batch_inversion": (C.c_int, [u64p, C.c_uint64, C.c_int]),
    "hodor_cuda_evaluate_at": (C.c_int, [u64p, C.c_uint64, u64p, u64p, C.c_int]),
    "hodor_cuda_merkle_build": (C.c_int, [u64p, C.c_uint64, u8p, C.c_int]),
    "hodor_cuda_lde_commit": (vp, [vp, C.c_uint32, C.c_uint32, C.c_int, C.c_int, u8p, C.c_int]),
    "hodor_cuda_lde_commit_batch": (C.c_int, [C.POINTER(vp), C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, C.c_int,
                                              C.POINTER(vp), u8p, C.c_int]),
    "hodor_cuda_tree_commit": (vp, [vp, C.c_uint64, C.c_int, u8p, C.c_int]),
    "hodor_cuda_tree_free": (None, [vp]),
    "hodor_cuda_tree_size": (C.c_uint64, [vp]),
    "hodor_cuda_tree_values": (vp, [vp]),
    "hodor_cuda_tree_nodes": (vp, [vp]),
    "hodor_cuda_tree_root": (C.c_int, [vp, u8p, u64p]),
    "hodor_cuda_tree_query": (C.c_int, [vp, C.c_uint64, u64p, u8p]),
    "hodor_cuda_tree_query_batch": (C.c_int, [vp, u64p, C.c_uint32, u64p, u8p]),
    "hodor_cuda_tree_read": (C.c_int, [vp, C.c_uint64, C.c_uint64, u64p, u8p]),
    "hodor_cuda_fri_commit": (vp, [vp, C.c_uint64, C.c_uint32, C.c_uint32, C.c_int, C.c_int]),
    "hodor_cuda_fri_free": (None, [vp]),
    "hodor_cuda_fri_num_steps": (C.c_int, [vp]),
    "hodor_cuda_fri_summary": (C.c_int, [vp, u8p, u64p, u64p]),
    "hodor_cuda_fri_layer": (C.c_int, [vp, C.c_uint32, u8p, u64p]),
    "hodor_cuda_fri_layer_size": (C.c_uint64, [vp, C.c_uint32]),
    "hodor_cuda_fri_query": (C.c_int, [vp, C.c_uint32, C.c_uint64, u64p, u8p]),
    "hodor_cuda_fri_produce_proof": (C.c_int, [vp, C.c_uint64, u64p, u64p, u8p]),
    "hodor_cuda_fri_commit_host": (C.c_int, [u64p, C.c_uint64, C.c_uint32, C.c_uint32, u8p, C.POINTER(u8p),
                                             C.POINTER(u64p), u64p, u8p, u64p, C.c_int]),
    "hodor_cuda_ntt_dev": (C.c_int, [vp, vp, C.c_uint32, u64p, C.c_int, vp]),
    "hodor_cuda_fft_dev": (C.c_int, [vp, vp, C.c_uint32, C.c_int, C.c_int, vp]),
    "hodor_cuda_ifft_dev": (C.c_int, [vp, vp, C.c_uint32, C.c_int, C.c_int, vp]),
    "hodor_cuda_lde_dev": (C.c_int, [vp, C.c_uint32, C.c_uint32, C.c_int, vp, C.c_int, vp]),
    "hodor_cuda_merkle_build_dev": (C.c_int, [vp, C.c_uint64, vp, vp, vp, C.c_int, vp]),
    "hodor_cuda_merkle_build_shard_dev": (C.c_int, [vp, C.c_uint64, C.c_uint32, vp, vp, vp, C.c_int, vp]),
    "hodor_cuda_comm_unique_id": (C.c_int, [u8p]),
    "hodor_cuda_comm_init": (C.c_int, [C.c_int, C.c_int, u8p]),
    "hodor_cuda_comm_destroy": (None, []),
    "hodor_cuda_comm_info": (C.c_int, [C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "hodor_cuda_ntt_sharded": (C.c_int, [vp, vp, C.c_uint32, u64p, C.c_int, vp]),
    "hodor_cuda_lde_fri_sharded": (C.c_int, [vp, C.c_uint32, C.c_uint32, C.c_int, C.c_uint32, u8p, u64p, u64p, C.c_int]),
    "hodor_cuda_merkle_top_dev": (C.c_int, [vp, C.c_uint64, vp, vp, C.c_int, vp]),
    "hodor_cuda_fri_fold_dev": (C.c_int, [vp, C.c_uint64, C.c_uint64, C.c_uint32, vp, vp, C.c_int, vp]),
    "hodor_cuda_fri_fold_shard_dev": (C.c_int, [vp, C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, vp, vp,
                                                C.c_int, vp]),
    "hodor_cuda_lde_cosets_dev": (C.c_int, [vp, C.c_uint32, C.c_uint32, C.c_int, C.c_uint32, C.c_uint32, C.c_uint32, vp,
                                            C.c_int, vp]),
    "hodor_cuda_distribute_powers_dev": (C.c_int, [vp, C.c_uint64, u64p, C.c_int, vp]),
    "hodor_cuda_precomputed_omegas_dev": (C.c_int, [vp, vp, vp, C.c_uint32, C.c_int, vp]),
    "hodor_cuda_precomputed_omegas": (C.c_int, [u64p, u64p, u64p, C.c_uint32, C.c_int]),
    "hodor_cuda_ali_dense_inverse_divisor": (C.c_int, [u64p, C.c_uint32, C.c_uint32, C.c_uint64, C.c_uint64, C.c_uint64,
                                                       C.POINTER(C.c_uint64), C.c_int]),
    "hodor_cuda_ali_boundary_inverse_divisor": (C.c_int, [u64p, C.c_uint32, C.c_uint32, C.c_uint64, C.c_int]),
    "hodor_cuda_ali_dense_inverse_divisor_dev": (C.c_int, [vp, C.c_uint32, C.c_uint32, C.c_uint64, C.c_uint64, C.c_uint64,
                                                           C.POINTER(C.c_uint64), C.c_int, vp]),
    "hodor_cuda_ali_boundary_inverse_divisor_dev": (C.c_int, [vp, C.c_uint32, C.c_uint32, C.c_uint64, C.c_int, vp]),
    "hodor_cuda_elementwise_dev": (C.c_int, [C.c_int, vp, vp, vp, C.c_uint64, C.c_int, vp]),
    "hodor_cuda_poly_op_dev": (C.c_int, [C.c_int, vp, vp, u64p, C.c_uint64, vp, C.c_uint64, C.c_int, vp]),
    "hodor_cuda_batch_inversion_dev": (C.c_int, [vp, C.c_uint64, vp, C.c_int, vp]),
    "hodor_cuda_evaluate_at_dev": (C.c_int, [vp, C.c_uint64, u64p, vp, C.c_int, vp]),
    "hodor_cuda_ntt_shard_cols_dev": (C.c_int, [vp, vp, C.c_uint32, C.c_uint32, C.c_uint32, u64p, C.c_int, vp]),
    "hodor_cuda_ntt_shard_rows_dev": (C.c_int, [vp, vp, C.c_uint32, C.c_uint32, C.c_uint32, u64p, C.c_int, vp]),
}

EXPORTED_SYMBOLS = tuple(_SIGS)

for _name, (_res, _args) in _SIGS.items():
    _fn = getattr(lib, _name)  # AttributeError here == header and library disagree
    _fn.restype = _res
    _fn.argtypes = _args


def last_error() -> str:
    msg = lib.hodor_cuda_last_error()
    return msg.decode() if msg else ""


def raise_last() -> None:
    """For entry points that return a handle or NULL: raise with the code the library recorded."""
    code = int(lib.hodor_cuda_last_error_code()) or ERR_CUDA
    check(code if code < 0 else ERR_CUDA)


def check(rc: int) -> int:
    if rc is None or rc >= 0:
        return rc
    msg = last_error()
    if rc in (ERR_DOMAIN, ERR_NOT_INVERTIBLE):  # the reference returns Err(SynthesisError::Error) for both
        raise SynthesisError(rc, msg)
    raise HodorError(rc, msg)


_initialised: Optional[int] = None


def init(device: Optional[int] = None) -> int:
    """Bind this process to one GPU (LOCAL_RANK by default).  Raises HodorError without a B200."""
    global _initialised
    if device is None:
        device = int(os.environ.get("LOCAL_RANK", "0"))
    if _initialised is not None and _initialised == device:
        return device
    check(lib.hodor_cuda_init(device))
    _initialised = device
    return device


def ensure_init() -> None:
    if _initialised is None:
        init()
