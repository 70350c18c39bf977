"""CPU restatement (Python big integers + hashlib) of `Prover::prove` for the Fibonacci AIR -- the
reference's own end-to-end test shape (src/prover/mod.rs:178-227: two registers A, B; constraints
A' = B and B' = A + B; boundary constraints "Initial A", "Initial B", "Final B"), BASELINE.json
configs[3].  TEST INFRASTRUCTURE: only tests/ and bench.py's cpu leg may import this.

Not a port of the AIR / ARP / ALI machinery: the constraint system is fixed, so what is restated is the
exact sequence of hot-path calls the generic code makes for it, with the reference's iteration orders
(IndexSet insertion order of the masks, constraint order, challenge order):

  witness polys   arp/per_register/mod.rs:13-68      best_fft(omega^-1) then * n^-1, per register
  f LDEs, trees   prover/mod.rs:73-87                w.lde(L), I::create, roots -> transcript
  g               ali/per_register/mod.rs:246-529    masks (distribute_powers), coset LDE (factor =
                                                     max constraint degree = 1), alpha per constraint,
                                                     divisors, boundary constraints, icoset_fft
  g LDE, tree     prover/mod.rs:91-95
  DEEP            ali/per_register/deep.rs:14-149    z, f(z m) per mask, 1/(x - z m), h1, h2
  FRI x 2         prover/mod.rs:112-113              proof_from_lde(h1), proof_from_lde(h2)
  queries         prover/mod.rs:120-151              transcript -> two challenge indices, FRI proofs,
                                                     f and g openings
  transcript      transcript/mod.rs:29-79            streaming keyed Blake2s; get_challenge = finalize,
                                                     absorb the digest, interpret it as a field element
  index           verifier/mod.rs:246-263            bytes_to_challenge_index

"parity unpinned": like the rest of oracle/, this follows the reference by reading it; the reference
cannot be executed in this image.  It is checked against the GPU replay (hodor_b200/fib_replay.py) at
sizes this model finishes in seconds.
"""
from __future__ import annotations

import hashlib
from dataclasses import dataclass, field
from typing import Dict, List, Tuple

from . import pymodel as M


class Transcript:
    """Blake2sTranscript (src/transcript/mod.rs:29-79)."""

    def __init__(self, F: M.Field):
        self.F = F
        self.h = hashlib.blake2s(key=M.KEY, person=M.PERSONAL, digest_size=32)

    def commit_bytes(self, b: bytes) -> None:
        self.h.update(b)

    def commit_field_element(self, x: int) -> None:
        self.h.update(x.to_bytes(8 * self.F.limbs, "big"))  # into_repr().write_be

    def get_challenge_bytes(self) -> bytes:
        v = self.h.copy().digest()  # State::finalize does not consume the state
        self.h.update(v)
        return v

    def get_challenge(self) -> int:
        return M.interpret_hash(self.F, self.get_challenge_bytes())  # same read_be / shave / from_repr


def bytes_to_challenge_index(b: bytes, lde_size: int, lde_factor: int) -> int:
    x = int.from_bytes(b[-8:], "big") % lde_size
    if x % lde_factor == 0:
        x = (x + 1) % lde_size
    if x % 2 == 0:
        x = (x + 1) % lde_size
    return x


def fibonacci_witness(F: M.Field, num_rows: int) -> Tuple[List[int], List[int]]:
    """TestTraceSystem::calculate_witness(1, 1, num_rows - 1) without its per-step println!."""
    a, b = [1], [1]
    for _ in range(num_rows - 1):
        a.append(b[-1])
        b.append((a[-2] + b[-1]) % F.p)
    return a, b


# (register, steps) of all_masks in IndexSet insertion order: constraint 0 terms (-B(t), +A(t+1)), constraint 1
# terms (-A(t), -B(t), +B(t+1)); the boundary constraints add (A, 0) and (B, 0), already present.
MASKS = [(1, 0), (0, 1), (0, 0), (1, 1)]
# constraints as lists of (coeff, register, steps): 0 = A(t+1) - B(t);  0 = B(t+1) - A(t) - B(t)
CONSTRAINTS = [[(-1, 1, 0), (+1, 0, 1)], [(-1, 0, 0), (-1, 1, 0), (+1, 1, 1)]]


@dataclass
class FibProof:
    f_iop_roots: List[bytes]
    g_iop_root: bytes
    f_at_z_m: List[int]
    g_at_z: int
    h1_roots: List[bytes]
    h2_roots: List[bytes]
    h1_final: List[int]
    h2_final: List[int]
    x_index_h1: int
    x_index_h2: int
    f_queries: List[Tuple[int, int, List[bytes]]]  # (index, value, path)
    g_query: Tuple[int, int, List[bytes]]
    h1_queries: List[Tuple[int, int, List[bytes]]]
    h2_queries: List[Tuple[int, int, List[bytes]]]
    # intermediates, for stage-by-stage comparison
    stages: Dict[str, object] = field(default_factory=dict)


def fri_queries(F, proto: M.FriPrototype, lde_values, start_index: int):
    """FRIProofPrototype::produce_proof (src/fri/query_producer.rs:10-53)."""
    out = []
    trees = [proto.l0_nodes] + proto.layer_nodes
    values = [lde_values] + proto.layer_values
    size, idx = len(lde_values), start_index
    for nodes, vals in zip(trees, values):
        for q in sorted([idx, (idx + size // 2) % size]):
            out.append((q, vals[q], M.merkle_path(F, nodes, vals, q)))
        idx, size = (idx if idx < size // 2 else idx - size // 2), size // 2
    return out


def prove(F: M.Field, log_rows: int, lde_factor: int = 16, fri_final: int = 1, keep_stages: bool = True) -> FibProof:
    p = F.p
    T = 1 << log_rows
    log_L = lde_factor.bit_length() - 1
    N = T * lde_factor
    omega = F.domain_generator(log_rows)          # column domain == constraints domain (max degree 1)
    g = F.generator
    tr = Transcript(F)
    A, B = fibonacci_witness(F, T)
    boundary = [(0, 0, 1), (1, 0, 1), (1, T - 1, B[T - 1])]  # (register, row, value): Initial A, Initial B, Final B

    # ---- Prover::new -> ALIInstance::from_arp: divisors on the coset g * <omega> (:60-227) ----------------
    xs = [g * pow(omega, i, p) % p for i in range(T)]
    last_root = pow(omega, T - 1, p)  # DenseConstraint{start_at 0, span 1}: the one excluded row
    dense_div = [(x - last_root) * pow(pow(x, T, p) - 1, -1, p) % p for x in xs]
    bdiv = {row: [pow((x - pow(omega, row, p)) % p, -1, p) for x in xs] for row in {0, T - 1}}

    # ---- witness polynomials --------------------------------------------------------------------------------
    f = [M.ifft(F, A, log_rows), M.ifft(F, B, log_rows)]
    f_ldes = [M.lde(F, w, log_rows, lde_factor, False) for w in f]
    f_trees = [M.merkle_create(F, l) for l in f_ldes]
    for t in f_trees:
        tr.commit_bytes(t[1])

    # ---- calculate_g ------------------------------------------------------------------------------------------
    masked = {(r, s): M.distribute_powers(F, f[r], pow(omega, s, p)) for (r, s) in MASKS}
    cache: Dict[Tuple[int, int], List[int]] = {}
    g_values = [0] * T
    batch = [0] * T
    for terms in CONSTRAINTS:
        alpha = tr.get_challenge()
        tr.get_challenge()  # beta: drawn, unused when the adjustment degree is 0
        cv = [0] * T
        for coeff, r, s in terms:
            if (r, s) not in cache:
                cache[(r, s)] = M.lde(F, masked[(r, s)], log_rows, 1, True)  # coset_lde(factor 1) == coset_fft
            base = cache[(r, s)]
            cv = [(c + coeff * v) % p for c, v in zip(cv, base)]
        batch = [(b + alpha * c) % p for b, c in zip(batch, cv)]
    g_values = [b * d % p for b, d in zip(batch, dense_div)]
    for r, row, value in boundary:
        alpha = tr.get_challenge()
        tr.get_challenge()
        w = list(f[r])
        w[0] = (w[0] - value) % p
        cv = M.lde(F, w, log_rows, 1, True)
        g_values = [(gv + alpha * c % p * d) % p for gv, c, d in zip(g_values, cv, bdiv[row])]
    g_poly = M.icoset_fft(F, g_values, log_rows)

    g_lde = M.lde(F, g_poly, log_rows, lde_factor, False)
    g_tree = M.merkle_create(F, g_lde)
    tr.commit_bytes(g_tree[1])

    # ---- calculate_deep ----------------------------------------------------------------------------------------
    z = tr.get_challenge()
    omega_N = F.domain_generator(log_rows + log_L)
    dom = [pow(omega_N, i, p) for i in range(N)]
    h1 = [0] * N
    f_at_z_m, inv_div = [], {}
    for r, s in MASKS:
        root = pow(omega, s, p) * z % p
        val = M.evaluate(F, f[r], root)
        f_at_z_m.append(val)
        if s not in inv_div:
            inv_div[s] = [pow((x - root) % p, -1, p) for x in dom]
        alpha = tr.get_challenge()
        d = inv_div[s]
        fl = f_ldes[r]
        h1 = [(h + (fl[i] - val) * alpha % p * d[i]) % p for i, h in enumerate(h1)]
    g_at_z = M.evaluate(F, g_poly, z)
    h2 = [(g_lde[i] - g_at_z) * pow((dom[i] - z) % p, -1, p) % p for i in range(N)]

    # ---- FRI ------------------------------------------------------------------------------------------------------
    p1 = M.fri_commit(F, h1, lde_factor, fri_final)
    p2 = M.fri_commit(F, h2, lde_factor, fri_final)
    h1_roots = [p1.l0_nodes[1]] + [n[1] for n in p1.layer_nodes]
    h2_roots = [p2.l0_nodes[1]] + [n[1] for n in p2.layer_nodes]
    for pr in (p1, p2):
        tr.commit_bytes(pr.final_root)
        for c in pr.final_coefficients:
            tr.commit_field_element(c)
    x1 = bytes_to_challenge_index(tr.get_challenge_bytes(), N, lde_factor)
    x2 = bytes_to_challenge_index(tr.get_challenge_bytes(), N, lde_factor)

    proof = FibProof(
        f_iop_roots=[t[1] for t in f_trees], g_iop_root=g_tree[1], f_at_z_m=f_at_z_m, g_at_z=g_at_z,
        h1_roots=h1_roots, h2_roots=h2_roots, h1_final=list(p1.final_coefficients), h2_final=list(p2.final_coefficients),
        x_index_h1=x1, x_index_h2=x2,
        f_queries=[(x1, l[x1], M.merkle_path(F, t, l, x1)) for t, l in zip(f_trees, f_ldes)],
        g_query=(x2, g_lde[x2], M.merkle_path(F, g_tree, g_lde, x2)),
        h1_queries=fri_queries(F, p1, h1, x1), h2_queries=fri_queries(F, p2, h2, x2))
    if keep_stages:
        proof.stages = {"witness": (A, B), "f": f, "g_poly": g_poly, "z": z, "h1_head": h1[:8], "h2_head": h2[:8],
                        "dense_div_head": dense_div[:4]}
    return proof
