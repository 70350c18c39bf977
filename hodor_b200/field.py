"""Host-side scalar field helpers: what `ff_ce`'s PrimeField gives the Rust caller (constants,
from_repr / into_repr, mul / inverse / pow on single elements).  All arithmetic runs inside
libhodor_b200.so; elements are numpy uint64[4] little-endian Montgomery limbs, the in-memory layout
of `Fr(FrRepr([u64; 4]))` (reference: src/bn256.rs:4-7).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from functools import lru_cache

import numpy as np

from . import _ffi
from ._ffi import BLS12_381_FR, BN254_FR, STARK252, check, lib, u64p  # noqa: F401

BN256_RS_FR = BLS12_381_FR  # the field `src/bn256.rs` really declares


def _p(a: np.ndarray):
    assert a.dtype == np.uint64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(u64p)


def limbs(x) -> np.ndarray:
    return np.ascontiguousarray(x, dtype=np.uint64).reshape(4)


def int_to_limbs(v: int) -> np.ndarray:
    return np.array([(v >> (64 * i)) & (2**64 - 1) for i in range(4)], dtype=np.uint64)


def limbs_to_int(a) -> int:
    return sum(int(x) << (64 * i) for i, x in enumerate(np.asarray(a).reshape(-1)[:4]))


@dataclass(frozen=True)
class FieldConstants:
    field_id: int
    modulus: int
    one: np.ndarray
    generator: np.ndarray
    root_of_unity: np.ndarray
    S: int
    NUM_BITS: int
    CAPACITY: int


@lru_cache(maxsize=None)
def constants(field_id: int) -> FieldConstants:
    m, o, g, r = (np.zeros(4, np.uint64) for _ in range(4))
    s, nb, cap = C.c_uint32(), C.c_uint32(), C.c_uint32()
    check(lib.hodor_field_constants(field_id, _p(m), _p(o), _p(g), _p(r), C.byref(s), C.byref(nb), C.byref(cap)))
    return FieldConstants(field_id, limbs_to_int(m), o, g, r, s.value, nb.value, cap.value)


def _bin(fn, field_id, a, b):
    out = np.zeros(4, np.uint64)
    check(fn(field_id, _p(limbs(a)), _p(limbs(b)), _p(out)))
    return out


def mul(field_id, a, b):
    return _bin(lib.hodor_field_mul, field_id, a, b)


def add(field_id, a, b):
    return _bin(lib.hodor_field_add, field_id, a, b)


def sub(field_id, a, b):
    return _bin(lib.hodor_field_sub, field_id, a, b)


def pow_(field_id, a, e: int):
    out = np.zeros(4, np.uint64)
    check(lib.hodor_field_pow(field_id, _p(limbs(a)), C.c_uint64(e), _p(out)))
    return out


def inverse(field_id, a):
    """Field::inverse; raises HodorError on zero (the reference returns None)."""
    out = np.zeros(4, np.uint64)
    check(lib.hodor_field_inverse(field_id, _p(limbs(a)), _p(out)))
    return out


def from_repr(field_id, value: int) -> np.ndarray:
    """PrimeField::from_repr / from_str: plain integer -> Montgomery limbs."""
    out = np.zeros(4, np.uint64)
    check(lib.hodor_field_from_repr(field_id, _p(int_to_limbs(value % constants(field_id).modulus)), _p(out)))
    return out


def into_repr(field_id, a) -> int:
    out = np.zeros(4, np.uint64)
    check(lib.hodor_field_into_repr(field_id, _p(limbs(a)), _p(out)))
    return limbs_to_int(out)


def one(field_id) -> np.ndarray:
    return constants(field_id).one.copy()


def zero() -> np.ndarray:
    return np.zeros(4, np.uint64)


def multiplicative_generator(field_id) -> np.ndarray:
    return constants(field_id).generator.copy()


def root_of_unity(field_id) -> np.ndarray:
    return constants(field_id).root_of_unity.copy()
