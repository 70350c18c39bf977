"""hodor_b200 -- B200 (sm_100a) hot path for matter-labs/hodor's Polynomial / Domain / IOP / FRI
surface: NTT and coset LDE over the `src/bn256.rs` field, FRI fold, Blake2s Merkle build.

Everything computes inside libhodor_b200.so (hand-written CUDA behind the C ABI of
include/hodor_b200.h).  This package is the host-side mirror of the reference interface: same
names, argument meaning and error behaviour.  There is no CPU fallback.
"""
from . import _ffi
from ._ffi import (BLS12_381_FR, BN254_FR, STARK252, FIELD_NAMES, HodorError, SynthesisError, init)  # noqa: F401
from .field import BN256_RS_FR  # noqa: F401
from .domains import Domain  # noqa: F401
from .polynomials import COEFFICIENTS, VALUES, Polynomial, Worker, lde_batch  # noqa: F401
from .iop import (Blake2sIopTree, Blake2sLeafEncoder, Blake2sTreeHasher, CommittedOracle, DeviceIOP, TrivialBlake2sIOP,  # noqa: F401
                  TrivialBlake2sIopQuery, TrivialCombiner)
from .fri import FRIProof, FRIProofPrototype, NaiveFriIop  # noqa: F401

__version__ = "0.1.0"
