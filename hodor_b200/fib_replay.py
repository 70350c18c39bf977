"""BASELINE.json configs[3]: the hot-path call sequence of `Prover::prove` (src/prover/mod.rs:66-174) for the
Fibonacci AIR of the reference's own end-to-end test (src/prover/mod.rs:178-227), replayed on device-resident
polynomials.

This is NOT the reference's AIR / ARP / ALI machinery (SURVEY.md section 2 keeps that on the caller's side of
the boundary): the constraint system is fixed -- registers A, B; A' = B, B' = A + B; boundary constraints
"Initial A", "Initial B", "Final B" -- and what is replayed is the sequence of `Polynomial` / `IOP` / `FriIop`
calls the generic code makes for it, in the reference's order (mask order, constraint order, challenge order),
every one of them a `_dev` entry point of the C ABI on vectors that stay in HBM:

    2 iNTT (witness interpolation)          arp/per_register/mod.rs:43-61
    2 LDE 2^k -> 2^k * L, 2 trees           prover/mod.rs:73-87
    4 distribute_powers, 4 + 3 coset NTTs,  ali/per_register/mod.rs:246-529
      elementwise passes, 1 icoset NTT
    1 LDE, 1 tree (g)                       prover/mod.rs:91-95
    5 evaluate_at, 3 batch inversions,      ali/per_register/deep.rs:14-149
      elementwise passes over 2^k * L
    2 FRI commit chains                     prover/mod.rs:112-113
    queries                                 prover/mod.rs:120-151

The Fiat-Shamir transcript (transcript/mod.rs:29-79) and `bytes_to_challenge_index`
(verifier/mod.rs:246-263) are O(1) host work and run on the host (hashlib), as they stay Rust in the
reference's deployment.  PCIe carries the witness in and roots / openings out.
"""
from __future__ import annotations

import hashlib
from dataclasses import dataclass, field
from typing import Dict, List, Tuple

import numpy as np
import torch

from . import device as dev
from . import field as fld
from . import precomputations as P
from .device import DevicePolynomial as DP
from .domains import Domain
from .iop import BLAKE2S_KEY, BLAKE2S_PERSONAL, Blake2sLeafEncoder, CommittedOracle

MASKS = [(1, 0), (0, 1), (0, 0), (1, 1)]  # all_masks in IndexSet insertion order: (register, steps)
CONSTRAINTS = [[(-1, 1, 0), (+1, 0, 1)], [(-1, 0, 0), (-1, 1, 0), (+1, 1, 1)]]  # (coeff, register, steps)


class Blake2sTranscript:
    """src/transcript/mod.rs:29-79."""

    def __init__(self, field_id: int):
        self.field_id = field_id
        self.h = hashlib.blake2s(key=BLAKE2S_KEY, person=BLAKE2S_PERSONAL, digest_size=32)

    def commit_bytes(self, b: bytes) -> None:
        self.h.update(bytes(b))

    def commit_field_element(self, x) -> None:
        self.h.update(fld.into_repr(self.field_id, x).to_bytes(32, "big"))

    def get_challenge_bytes(self) -> bytes:
        v = self.h.copy().digest()
        self.h.update(v)
        return v

    def get_challenge(self) -> np.ndarray:
        return Blake2sLeafEncoder.interpret_hash(self.field_id, self.get_challenge_bytes())


def bytes_to_challenge_index(b: bytes, lde_size: int, lde_factor: int) -> int:
    x = int.from_bytes(b[-8:], "big") % lde_size
    if x % lde_factor == 0:
        x = (x + 1) % lde_size
    if x % 2 == 0:
        x = (x + 1) % lde_size
    return x


def fibonacci_witness(field_id: int, num_rows: int) -> Tuple[np.ndarray, np.ndarray]:
    """The trace as Montgomery limbs (additions are linear: starting from R mod p keeps every row in Montgomery
    form).  Host work, outside every timed region: the witness is the prover's input."""
    p = fld.constants(field_id).modulus
    one = fld.limbs_to_int(fld.one(field_id))
    a, b = one, one
    ab, bb = bytearray(), bytearray()
    for _ in range(num_rows):
        ab += a.to_bytes(32, "little")
        bb += b.to_bytes(32, "little")
        a, b = b, (a + b) % p
    return (np.frombuffer(bytes(ab), np.uint64).reshape(num_rows, 4).copy(),
            np.frombuffer(bytes(bb), np.uint64).reshape(num_rows, 4).copy())


@dataclass
class FibProof:
    f_iop_roots: List[bytes]
    g_iop_root: bytes
    f_at_z_m: List[np.ndarray]
    g_at_z: np.ndarray
    h1_roots: List[bytes]
    h2_roots: List[bytes]
    h1_final: np.ndarray
    h2_final: np.ndarray
    x_index_h1: int
    x_index_h2: int
    f_queries: list
    g_query: object
    fri_proof_h1: object
    fri_proof_h2: object
    stages: Dict[str, object] = field(default_factory=dict)


class FibonacciProver:
    """`Prover::new` (src/prover/mod.rs:46-64): the ALI divisor precomputation (ali/per_register/mod.rs:60-227),
    built on the device once per instance."""

    def __init__(self, field_id: int, log_rows: int, lde_factor: int = 16, fri_final_degree_plus_one: int = 1):
        self.fid, self.log_rows, self.L, self.fri_final = field_id, log_rows, lde_factor, fri_final_degree_plus_one
        self.T = 1 << log_rows
        self.N = self.T * lde_factor
        fid, T = field_id, self.T
        self.omega = Domain.new_for_size(fid, T).generator
        self.omega_N = Domain.new_for_size(fid, self.N).generator
        # constraints are evaluated on the coset g * <omega> (coset_lde factor = max constraint degree = 1), so the
        # constraints domain is the column domain; both divisor families are one C-ABI call each
        col = Domain.new_for_size(fid, T)
        # Dense{start_at 0, span 1}: (x - omega^(T-1)) / (x^T - 1)  (:60-162)
        self.dense_div, _ = P.inverse_divisor_for_dense_constraint_in_coset(col, col, P.DenseConstraint(0, 1), T)
        # boundary rows: 1 / (x - omega^row)  (:214-227)
        self.bdiv = {row: P.boundary_constraint_inverse_divisor(col, col, row) for row in (0, T - 1)}
        torch.cuda.synchronize()

    def prove(self, a_col: np.ndarray, b_col: np.ndarray, keep_stages: bool = False) -> FibProof:
        fid, T, L, N = self.fid, self.T, self.L, self.N
        W = None
        neg = lambda x: fld.sub(fid, fld.zero(), x)  # noqa: E731
        tr = Blake2sTranscript(fid)
        boundary = [(0, 0, fld.one(fid)), (1, 0, fld.one(fid)), (1, T - 1, b_col[T - 1])]

        # ---- witness polynomials: iNTT per register ----------------------------------------------------------
        f = [DP.from_values(fid, a_col).ifft(W), DP.from_values(fid, b_col).ifft(W)]
        # ---- f LDEs and their oracles -------------------------------------------------------------------------
        f_ldes = [w.lde(W, L) for w in f]
        f_oracles = [CommittedOracle.create_on_device(fid, l.coeffs) for l in f_ldes]
        for o in f_oracles:
            tr.commit_bytes(o.get_root())

        # ---- calculate_g -----------------------------------------------------------------------------------------
        masked = {}
        for r, s in MASKS:
            m = f[r].clone()
            m.distribute_powers(W, fld.pow_(fid, self.omega, s))
            masked[(r, s)] = m
        cache: Dict[Tuple[int, int], DP] = {}
        zero_vec = lambda n: DP(fid, torch.zeros((n, 4), dtype=torch.int64, device="cuda"), "Values")  # noqa: E731
        g_values, batch = zero_vec(T), zero_vec(T)
        for terms in CONSTRAINTS:
            alpha = tr.get_challenge()
            tr.get_challenge()  # beta (unused at adjustment degree 0)
            cv = zero_vec(T)
            for coeff, r, s in terms:
                if (r, s) not in cache:
                    base = masked[(r, s)].coset_lde(W, 1)
                    base.pow(W, 1)
                    cache[(r, s)] = base
                sub = cache[(r, s)].clone()
                if coeff == -1:
                    sub.negate(W)
                cv.add_assign(W, sub)
            cv.add_constant(W, fld.zero())
            cv.scale(W, alpha)
            batch.add_assign(W, cv)
        batch.mul_assign(W, self.dense_div)
        g_values.add_assign(W, batch)
        for r, row, value in boundary:
            alpha = tr.get_challenge()
            tr.get_challenge()
            w = f[r].clone()
            c0 = dev.to_host(w.coeffs[:1])[0]  # coeffs[0] -= value: one element through the host scalar helpers
            w.coeffs[:1] = dev.to_device(fld.sub(fid, c0, value).reshape(1, 4))
            cv = w.coset_lde(W, 1)
            cv.scale(W, alpha)
            cv.mul_assign(W, self.bdiv[row])
            g_values.add_assign(W, cv)
        g_poly = g_values.icoset_fft(W)

        g_lde = g_poly.lde(W, L)
        g_oracle = CommittedOracle.create_on_device(fid, g_lde.coeffs)
        tr.commit_bytes(g_oracle.get_root())

        # ---- calculate_deep -----------------------------------------------------------------------------------------
        z = tr.get_challenge()
        dom = DP.filled(fid, N, fld.one(fid))
        dom.distribute_powers(W, self.omega_N)  # evaluate_at_domain_for_degree_one's u = omega_N^i
        h1 = zero_vec(N)
        f_at_z_m, inv_div = [], {}
        for r, s in MASKS:
            root = fld.mul(fid, fld.pow_(fid, self.omega, s), z)
            val = f[r].evaluate_at(W, root)
            f_at_z_m.append(val)
            if s not in inv_div:
                q = dom.clone()
                q.add_constant(W, neg(root))
                q.batch_inversion(W)
                inv_div[s] = q
            t = f_ldes[r].clone()
            t.add_constant(W, neg(val))
            alpha = tr.get_challenge()
            t.scale(W, alpha)
            t.mul_assign(W, inv_div[s])
            h1.add_assign(W, t)
        q = dom.clone()
        q.add_constant(W, neg(z))
        q.batch_inversion(W)
        g_at_z = g_poly.evaluate_at(W, z)
        h2 = g_lde.clone()
        h2.add_constant(W, neg(g_at_z))
        h2.mul_assign(W, q)

        # ---- FRI ------------------------------------------------------------------------------------------------------
        p1 = dev.fri_commit(h1.coeffs, L, self.fri_final, fid)
        p2 = dev.fri_commit(h2.coeffs, L, self.fri_final, fid)
        for pr in (p1, p2):
            tr.commit_bytes(pr.get_final_root())
            for c in pr.get_final_coefficients():
                tr.commit_field_element(c)
        x1 = bytes_to_challenge_index(tr.get_challenge_bytes(), N, L)
        x2 = bytes_to_challenge_index(tr.get_challenge_bytes(), N, L)
        proof1 = p1.produce_proof(None, x1)
        proof2 = p2.produce_proof(None, x2)
        f_queries = [o.query(x1) for o in f_oracles]
        g_query = g_oracle.query(x2)
        out = FibProof([o.get_root() for o in f_oracles], g_oracle.get_root(), f_at_z_m, g_at_z, p1.get_roots(), p2.get_roots(),
                       p1.get_final_coefficients(), p2.get_final_coefficients(), x1, x2, f_queries, g_query, proof1, proof2)
        if keep_stages:
            out.stages = {"f": [w.to_host() for w in f], "g_poly": g_poly.to_host(), "z": z, "h1_head": dev.to_host(h1.coeffs[:8]),
                          "h2_head": dev.to_host(h2.coeffs[:8]), "dense_div_head": dev.to_host(self.dense_div.coeffs[:4])}
        torch.cuda.synchronize()
        for o in f_oracles + [g_oracle]:
            o.free()
        p1.free()
        p2.free()
        return out
