"""TEST INFRASTRUCTURE ONLY -- ctypes front end of oracle/hodor_oracle.c (the CPU restatement of
Hodor's hot path).  Imported by tests/, bench.py's cpu_baseline / --impl reference legs and
__graft_entry__.smoke(); never by hodor_b200/.  See the header of hodor_oracle.c for the parity
status ("parity unpinned") and the reference citations.

Field elements travel as numpy uint64 arrays of shape (n, 4): little-endian limbs, Montgomery form.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import List, Optional, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libhodor_oracle.so")

BLS12_381_FR, BN254_FR, STARK252 = 0, 1, 2
FIELD_NAMES = {0: "bls12_381_fr", 1: "bn254_fr", 2: "stark252"}


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "hodor_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _LIB_PATH


_lib: Optional[C.CDLL] = None

_u64p = C.POINTER(C.c_uint64)
_u8p = C.POINTER(C.c_uint8)


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        _lib = C.CDLL(_LIB_PATH)
    return _lib


def _p64(a: np.ndarray):
    assert a.dtype == np.uint64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_u64p)


def _p8(a: np.ndarray):
    assert a.dtype == np.uint8 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_u8p)


def _check(rc: int, what: str):
    if rc < 0:
        raise ValueError(f"oracle {what} failed: rc={rc}")
    return rc


def default_cpus() -> int:
    return len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)


# ------------------------------------------------------------------------------------------
def field_constants(field: int) -> dict:
    p, r, r2, gen, root = (np.zeros(4, np.uint64) for _ in range(5))
    inv, s, nb = C.c_uint64(), C.c_uint32(), C.c_uint32()
    _check(lib().oracle_field_constants(field, _p64(p), _p64(r), _p64(r2), C.byref(inv), _p64(gen), _p64(root),
                                        C.byref(s), C.byref(nb)), "field_constants")
    return dict(p=p, r=r, r2=r2, inv=inv.value, generator=gen, root_of_unity=root, s=s.value, num_bits=nb.value)


def limbs_to_int(a: np.ndarray) -> int:
    return sum(int(x) << (64 * i) for i, x in enumerate(a.reshape(-1)[:4]))


def int_to_limbs(v: int) -> np.ndarray:
    return np.array([(v >> (64 * i)) & (2**64 - 1) for i in range(4)], dtype=np.uint64)


def ints_to_array(vals) -> np.ndarray:
    out = np.zeros((len(vals), 4), np.uint64)
    for i, v in enumerate(vals):
        out[i] = int_to_limbs(v)
    return out


def array_to_ints(a: np.ndarray) -> List[int]:
    return [limbs_to_int(row) for row in a.reshape(-1, 4)]


def _binop(name, field, a, b):
    a = np.ascontiguousarray(a, np.uint64).reshape(-1, 4)
    b = np.ascontiguousarray(b, np.uint64).reshape(-1, 4)
    out = np.zeros_like(a)
    _check(getattr(lib(), name)(field, _p64(a), _p64(b), _p64(out), C.c_size_t(a.shape[0])), name)
    return out


def mul(field, a, b):
    return _binop("oracle_mul", field, a, b)


def add(field, a, b):
    return _binop("oracle_add", field, a, b)


def sub(field, a, b):
    return _binop("oracle_sub", field, a, b)


def inverse(field, a):
    a = np.ascontiguousarray(a, np.uint64).reshape(4)
    out = np.zeros(4, np.uint64)
    _check(lib().oracle_inverse(field, _p64(a), _p64(out)), "inverse")
    return out


def pow_(field, a, e: int):
    a = np.ascontiguousarray(a, np.uint64).reshape(4)
    out = np.zeros(4, np.uint64)
    _check(lib().oracle_pow(field, _p64(a), C.c_uint64(e), _p64(out)), "pow")
    return out


def to_mont(field, plain):
    plain = np.ascontiguousarray(plain, np.uint64).reshape(-1, 4)
    out = np.zeros_like(plain)
    _check(lib().oracle_to_mont(field, _p64(plain), _p64(out), C.c_size_t(plain.shape[0])), "to_mont")
    return out


def from_mont(field, mont):
    mont = np.ascontiguousarray(mont, np.uint64).reshape(-1, 4)
    out = np.zeros_like(mont)
    _check(lib().oracle_from_mont(field, _p64(mont), _p64(out), C.c_size_t(mont.shape[0])), "from_mont")
    return out


def domain_generator(field, log_n: int) -> np.ndarray:
    out = np.zeros(4, np.uint64)
    _check(lib().oracle_domain_generator(field, log_n, _p64(out)), "domain_generator")
    return out


def serial_fft(field, a, omega, log_n):
    a = np.array(a, np.uint64, copy=True).reshape(-1, 4)
    assert a.shape[0] == 1 << log_n
    _check(lib().oracle_serial_fft(field, _p64(a), _p64(np.ascontiguousarray(omega, np.uint64)), log_n), "serial_fft")
    return a


def serial_fft_radix_4(field, a, omega, log_n):
    a = np.array(a, np.uint64, copy=True).reshape(-1, 4)
    _check(lib().oracle_serial_fft_radix_4(field, _p64(a), _p64(np.ascontiguousarray(omega, np.uint64)), log_n),
           "serial_fft_radix_4")
    return a


def serial_dif_fft(field, a, omega, log_n, non_zero_entries=None):
    """serial_DIT_fft (src/fft/dit_fft/mod.rs:4-53); non_zero_entries prunes a zero tail."""
    a = np.array(a, np.uint64, copy=True).reshape(-1, 4)
    nz = a.shape[0] if non_zero_entries is None else non_zero_entries
    _check(lib().oracle_serial_dif_fft(field, _p64(a), _p64(np.ascontiguousarray(omega, np.uint64)), log_n, C.c_uint64(nz)),
           "serial_dif_fft")
    return a


def best_fft(field, a, omega, log_n, cpus=None, hint: int = -1):
    a = np.array(a, np.uint64, copy=True).reshape(-1, 4)
    cpus = cpus or default_cpus()
    _check(lib().oracle_best_fft(field, _p64(a), _p64(np.ascontiguousarray(omega, np.uint64)), log_n, cpus,
                                 C.c_long(hint)), "best_fft")
    return a


def best_fft_radix_4(field, a, omega, log_n, cpus=None):
    a = np.array(a, np.uint64, copy=True).reshape(-1, 4)
    cpus = cpus or default_cpus()
    _check(lib().oracle_best_fft_radix_4(field, _p64(a), _p64(np.ascontiguousarray(omega, np.uint64)), log_n, cpus),
           "best_fft_radix_4")
    return a


def distribute_powers(field, a, g, cpus=None):
    a = np.array(a, np.uint64, copy=True).reshape(-1, 4)
    cpus = cpus or default_cpus()
    _check(lib().oracle_distribute_powers(field, _p64(a), C.c_size_t(a.shape[0]),
                                          _p64(np.ascontiguousarray(g, np.uint64)), cpus), "distribute_powers")
    return a


def fft(field, a, log_n, cpus=None, coset=False):
    a = np.array(a, np.uint64, copy=True).reshape(-1, 4)
    _check(lib().oracle_fft(field, _p64(a), log_n, cpus or default_cpus(), int(coset)), "fft")
    return a


def ifft(field, a, log_n, cpus=None, coset=False):
    a = np.array(a, np.uint64, copy=True).reshape(-1, 4)
    _check(lib().oracle_ifft(field, _p64(a), log_n, cpus or default_cpus(), int(coset)), "ifft")
    return a


def evaluate_at(field, coeffs, g, cpus=None) -> np.ndarray:
    """Polynomial::evaluate_at (src/polynomials/mod.rs:685-711)."""
    coeffs = np.ascontiguousarray(coeffs, np.uint64).reshape(-1, 4)
    out = np.zeros(4, np.uint64)
    _check(lib().oracle_evaluate_at(field, _p64(coeffs), C.c_size_t(coeffs.shape[0]),
                                    _p64(np.ascontiguousarray(g, np.uint64)), cpus or default_cpus(), _p64(out)),
           "evaluate_at")
    return out


def batch_inversion(field, a, cpus=None):
    """Polynomial::batch_inversion (src/polynomials/mod.rs:889-954); None when an element is zero
    (the reference returns Err(SynthesisError::Error) and leaves the vector untouched)."""
    a = np.array(a, np.uint64, copy=True).reshape(-1, 4)
    rc = lib().oracle_batch_inversion(field, _p64(a), C.c_size_t(a.shape[0]), cpus or default_cpus())
    if rc == -2:
        return None
    _check(rc, "batch_inversion")
    return a


def lde(field, coeffs, log_n, factor, coset, cpus=None):
    coeffs = np.ascontiguousarray(coeffs, np.uint64).reshape(-1, 4)
    assert coeffs.shape[0] == 1 << log_n
    out = np.zeros((coeffs.shape[0] * factor, 4), np.uint64)
    # default: one worker chunk per coset (cpus >= factor), the regime in which the reference's
    # chunk-index generator (src/polynomials/mod.rs:448,575) yields the mathematically right cosets
    cpus = cpus or max(default_cpus(), factor)
    _check(lib().oracle_lde(field, _p64(coeffs), log_n, factor, int(coset), _p64(out), cpus), "lde")
    return out


def filtering_lde(field, coeffs, log_n, factor, coset, cpus=None):
    """Polynomial::(coset_)filtering_lde (src/polynomials/mod.rs:355-368, 484-499) -> serial_lde
    (src/fft/lde.rs:15-126), the zero-aware NTT of the zero-padded vector."""
    coeffs = np.ascontiguousarray(coeffs, np.uint64).reshape(-1, 4)
    assert coeffs.shape[0] == 1 << log_n
    out = np.zeros((coeffs.shape[0] * factor, 4), np.uint64)
    _check(lib().oracle_filtering_lde(field, _p64(coeffs), log_n, factor, int(coset), _p64(out), cpus or default_cpus()),
           "filtering_lde")
    return out


def hash_leaf(field, x) -> bytes:
    out = np.zeros(32, np.uint8)
    lib().oracle_hash_leaf(field, _p64(np.ascontiguousarray(x, np.uint64).reshape(4)), _p8(out))
    return out.tobytes()


def hash_node(l: bytes, r: bytes) -> bytes:
    out = np.zeros(32, np.uint8)
    lib().oracle_hash_node(_p8(np.frombuffer(l, np.uint8).copy()), _p8(np.frombuffer(r, np.uint8).copy()), _p8(out))
    return out.tobytes()


def blake2s(data: bytes) -> bytes:
    out = np.zeros(32, np.uint8)
    buf = np.frombuffer(data, np.uint8).copy() if data else np.zeros(1, np.uint8)
    lib().oracle_blake2s(_p8(buf), C.c_size_t(len(data)), _p8(out))
    return out.tobytes()


def merkle_create(field, leaves, cpus=None) -> np.ndarray:
    leaves = np.ascontiguousarray(leaves, np.uint64).reshape(-1, 4)
    n = leaves.shape[0]
    nodes = np.zeros((n, 32), np.uint8)
    _check(lib().oracle_merkle_create(field, _p64(leaves), C.c_size_t(n), _p8(nodes), cpus or default_cpus()),
           "merkle_create")
    return nodes


def interpret_hash(field, digest) -> np.ndarray:
    d = np.ascontiguousarray(np.frombuffer(bytes(digest), np.uint8)).copy()
    out = np.zeros(4, np.uint64)
    _check(lib().oracle_interpret_hash(field, _p8(d), _p64(out)), "interpret_hash")
    return out


def merkle_path(field, nodes: np.ndarray, leaves: np.ndarray, index: int) -> List[bytes]:
    """get_path (src/iop/blake2s_trivial_iop.rs:251-279)."""
    leaves = leaves.reshape(-1, 4)
    path = [hash_leaf(field, leaves[index ^ 1])]
    idx = (nodes.shape[0] + index) >> 1
    while idx > 1:
        path.append(nodes[idx ^ 1].tobytes())
        idx >>= 1
    return path


def merkle_verify(field, root: bytes, leaf, path, index: int) -> bool:
    """verify (src/iop/blake2s_trivial_iop.rs:236-249)."""
    h, idx = hash_leaf(field, leaf), index
    for el in path:
        h = hash_node(h, el) if idx & 1 == 0 else hash_node(el, h)
        idx >>= 1
    return h == bytes(root)


class FriPrototype:
    """Mirror of FRIProofPrototype (src/fri/mod.rs:107-117), arrays instead of trees."""

    def __init__(self, l0_nodes, layer_nodes, layer_values, challenges, final_root, final_coefficients):
        self.l0_nodes = l0_nodes
        self.layer_nodes = layer_nodes
        self.layer_values = layer_values
        self.challenges = challenges
        self.final_root = final_root
        self.final_coefficients = final_coefficients

    def roots(self) -> List[bytes]:
        return [self.l0_nodes[1].tobytes()] + [n[1].tobytes() for n in self.layer_nodes]


def fri_num_steps(n: int, lde_factor: int, out_coeffs: int) -> int:
    return ((n // lde_factor) // out_coeffs).bit_length() - 1


def fri_commit(field, lde_values, lde_factor, out_coeffs, cpus=None) -> FriPrototype:
    v = np.ascontiguousarray(lde_values, np.uint64).reshape(-1, 4)
    n = v.shape[0]
    steps = fri_num_steps(n, lde_factor, out_coeffs)
    if steps < 1:
        raise ValueError("num_steps must be >= 1 (the reference panics otherwise)")
    l0 = np.zeros((n, 32), np.uint8)
    ln = [np.zeros((n >> (i + 1), 32), np.uint8) for i in range(steps)]
    lv = [np.zeros((n >> (i + 1), 4), np.uint64) for i in range(steps)]
    ch = np.zeros((steps, 4), np.uint64)
    fr = np.zeros(32, np.uint8)
    fc = np.zeros((out_coeffs, 4), np.uint64)
    lnp = (_u8p * steps)(*[_p8(a) for a in ln])
    lvp = (_u64p * steps)(*[_p64(a) for a in lv])
    rc = lib().oracle_fri_commit(field, _p64(v), C.c_size_t(n), lde_factor, out_coeffs, _p8(l0), lnp, lvp, _p64(ch),
                                 _p8(fr), _p64(fc), cpus or default_cpus())
    _check(rc, "fri_commit")
    assert rc == steps
    return FriPrototype(l0, ln, lv, ch, fr.tobytes(), fc)


def random_elements(field, count: int, seed: int = 0x3DBE62598D313D76) -> np.ndarray:
    out = np.zeros((count, 4), np.uint64)
    _check(lib().oracle_random_elements(field, _p64(out), C.c_size_t(count), C.c_uint64(seed)), "random_elements")
    return out
