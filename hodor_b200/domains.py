"""Mirror of the reference's Domain (src/domains/mod.rs:14-70)."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import List, Tuple

import numpy as np

from ._ffi import check, lib
from .field import _p


@dataclass(frozen=True, eq=False)
class Domain:
    field_id: int
    size: int
    power_of_two: int
    generator: np.ndarray

    @staticmethod
    def new_for_size(field_id: int, size: int) -> "Domain":
        """Domain::new_for_size (:21-44).  Raises SynthesisError when 2^k exceeds the 2-adicity."""
        size = 1 if size <= 1 else 1 << (size - 1).bit_length()  # next_power_of_two
        power_of_two = size.bit_length() - 1
        gen = np.zeros(4, np.uint64)
        check(lib.hodor_domain_generator(field_id, C.c_uint32(power_of_two), _p(gen)))
        return Domain(field_id, size, power_of_two, gen)

    @staticmethod
    def coset_for_natural_index_and_size(natural_index: int, domain_size: int) -> List[int]:
        assert domain_size > 1 and domain_size & (domain_size - 1) == 0
        pair = (natural_index + domain_size // 2) % domain_size
        return sorted([natural_index, pair])

    @staticmethod
    def index_and_size_for_next_domain(natural_index: int, domain_size: int) -> Tuple[int, int]:
        assert domain_size > 1 and domain_size & (domain_size - 1) == 0
        next_size = domain_size // 2
        return (natural_index if natural_index < next_size else natural_index - next_size), next_size
