"""TEST INFRASTRUCTURE ONLY -- big-integer model of Hodor's NTT / LDE / FRI / Merkle path.

This is the *second, independent* restatement (the first is oracle/hodor_oracle.c).  It uses
Python integers and hashlib.blake2s, shares no code with the C oracle or with the CUDA product,
and exists so that the C oracle (which in turn checks the CUDA kernels) can itself be
cross-checked.  Nothing under hodor_b200/ may import this module; only tests/, bench.py's
cpu_baseline leg and __graft_entry__.smoke() may.

PARITY STATUS: "parity unpinned".  The reference (matter-labs/hodor @ 76fc894) is Rust, cannot be
compiled in this image (no cargo/rustc) and its tests hold NO golden vectors for this path
(SURVEY.md section 8c).  What *is* pinned against reference-held constants:
  * Montgomery form with R = 2^256 over 4 little-endian u64 limbs:
    src/experiments/square_root_calculator/fp2.rs:10-22 (MINUS_ONE, NON_RESIDUE) -- see
    tests/test_oracle_pins.py;
  * Montgomery mul / add / sub / inverse over `experiments::Fr`: E_PRECOMPUTED / F_PRECOMPUTED
    (fp2.rs:51-81), the printed output of the reference's own `find_c` test (fp2.rs:358-412),
    reproduced bit for bit -- tests/test_reference_kat.py.
Everything else follows the published algorithms of the un-vendored dependencies
(ff_ce "0.7" derive: Montgomery arithmetic, root_of_unity = g^((p-1)/2^S);
 blake2s_simd "0.5": RFC 7693 keyed+personalised Blake2s) and the reference call sites cited on
each function below.  All reference paths are relative to /root/reference.
"""
from __future__ import annotations

import hashlib
from dataclasses import dataclass
from typing import List, Sequence


# --------------------------------------------------------------------------------------------
# Fields  (src/bn256.rs:4-7, src/experiments/mod.rs:18-21, src/lib.rs:35-38)
# --------------------------------------------------------------------------------------------
@dataclass(frozen=True)
class Field:
    name: str
    p: int
    generator: int  # PrimeFieldGenerator
    limbs: int  # u64 limbs of FrRepr (ff_ce: smallest count with 2p < 2^(64*limbs))

    @property
    def num_bits(self) -> int:
        return self.p.bit_length()

    @property
    def capacity(self) -> int:
        return self.num_bits - 1

    @property
    def s(self) -> int:
        t, s = self.p - 1, 0
        while t % 2 == 0:
            t //= 2
            s += 1
        return s

    @property
    def R(self) -> int:
        return pow(2, 64 * self.limbs, self.p)

    @property
    def root_of_unity(self) -> int:
        # ff_ce derive: ROOT_OF_UNITY = GENERATOR^t, t = (p-1)/2^S
        return pow(self.generator, (self.p - 1) >> self.s, self.p)

    # Montgomery conversions (ff_ce: from_repr multiplies by R^2 and reduces -> x*R mod p)
    def to_mont(self, x: int) -> int:
        return (x % self.p) * self.R % self.p

    def from_mont(self, xm: int) -> int:
        return xm * pow(self.R, -1, self.p) % self.p

    def raw_bytes(self, x: int) -> bytes:
        """encode_leaf: raw Montgomery limbs, little-endian, zero padded to 32 B
        (src/iop/blake2s_trivial_iop.rs:36-42)."""
        return self.to_mont(x).to_bytes(8 * self.limbs, "little").ljust(32, b"\0")

    def domain_generator(self, log_n: int) -> int:
        """Domain::new_for_size (src/domains/mod.rs:21-44)."""
        if log_n > self.s:
            raise ValueError("domain too large for the field's 2-adicity")
        g = self.root_of_unity
        for _ in range(log_n, self.s):
            g = g * g % self.p
        return g


BLS12_381_FR = Field(  # this is what src/bn256.rs actually declares
    "bls12_381_fr",
    52435875175126190479447740508185965837690552500527637822603658699938581184513,
    7,
    4,
)
BN254_FR = Field(
    "bn254_fr",
    21888242871839275222246405745257275088548364400416034343698204186575808495617,
    7,  # PrimeFieldGenerator of pairing_ce's bn256::Fr (the reference itself does not declare this field)
    4,
)
STARK252 = Field(
    "stark252",
    3618502788666131213697322783095070105623107215331596699973092056135872020481,
    3,
    4,
)
F257 = Field("f257", 257, 3, 1)

FIELDS = {f.name: f for f in (BLS12_381_FR, BN254_FR, STARK252, F257)}


# --------------------------------------------------------------------------------------------
# NTT  (src/fft/fft.rs:21-66 serial_fft; definitional DFT as an independent check)
# --------------------------------------------------------------------------------------------
def bitreverse(n: int, l: int) -> int:
    r = 0
    for _ in range(l):
        r = (r << 1) | (n & 1)
        n >>= 1
    return r


def dft(F: Field, a: Sequence[int], omega: int) -> List[int]:
    """out[k] = sum_j a[j] * omega^(j*k)   -- O(n^2), the definition."""
    n, p = len(a), F.p
    return [sum(a[j] * pow(omega, j * k, p) for j in range(n)) % p for k in range(n)]


def serial_fft(F: Field, a: Sequence[int], omega: int, log_n: int) -> List[int]:
    """src/fft/fft.rs:21-66: bit-reverse swap, then log_n DIT stages with running twiddles."""
    p = F.p
    a = list(a)
    n = len(a)
    assert n == 1 << log_n
    for k in range(n):
        rk = bitreverse(k, log_n)
        if k < rk:
            a[k], a[rk] = a[rk], a[k]
    m = 1
    for _ in range(log_n):
        w_m = pow(omega, n // (2 * m), p)
        for k in range(0, n, 2 * m):
            w = 1
            for j in range(m):
                t = a[k + j + m] * w % p
                a[k + j + m] = (a[k + j] - t) % p
                a[k + j] = (a[k + j] + t) % p
                w = w * w_m % p
        m *= 2
    return a


def distribute_powers(F: Field, a: Sequence[int], g: int) -> List[int]:
    """src/fft/mod.rs:110-123: a[j] <- a[j] * g^j."""
    p, out, u = F.p, [], 1
    for v in a:
        out.append(v * u % p)
        u = u * g % p
    return out


def ifft(F: Field, a: Sequence[int], log_n: int) -> List[int]:
    """Polynomial::ifft (src/polynomials/mod.rs:773-798): NTT with omega^-1, then * n^-1."""
    p = F.p
    omega_inv = pow(F.domain_generator(log_n), -1, p)
    minv = pow(len(a), -1, p)
    return [v * minv % p for v in serial_fft(F, a, omega_inv, log_n)]


def icoset_fft(F: Field, a: Sequence[int], log_n: int) -> List[int]:
    """Polynomial::icoset_fft (src/polynomials/mod.rs:800-807)."""
    return distribute_powers(F, ifft(F, a, log_n), pow(F.generator, -1, F.p))


def lde(F: Field, coeffs: Sequence[int], log_n: int, factor: int, coset: bool) -> List[int]:
    """(coset_)lde_using_multiple_cosets (src/polynomials/mod.rs:418-482, 544-609) with one
    worker chunk per coset (num_cpus >= factor; see DESIGN.md on the chunk-index quirk)."""
    n, p = len(coeffs), F.p
    assert n == 1 << log_n
    if factor == 1:
        c = distribute_powers(F, coeffs, F.generator) if coset else list(coeffs)
        return serial_fft(F, c, F.domain_generator(log_n), log_n)
    log_f = factor.bit_length() - 1
    assert factor == 1 << log_f
    coset_omega = F.domain_generator(log_n + log_f)
    omega = F.domain_generator(log_n)
    results = []
    for i in range(factor):
        shift = pow(coset_omega, i, p)
        if coset:
            shift = shift * F.generator % p
        results.append(serial_fft(F, distribute_powers(F, coeffs, shift), omega, log_n))
    return [results[idx % factor][idx // factor] for idx in range(n * factor)]


def evaluate(F: Field, coeffs: Sequence[int], x: int) -> int:
    acc = 0
    for c in reversed(coeffs):
        acc = (acc * x + c) % F.p
    return acc


# --------------------------------------------------------------------------------------------
# Blake2s Merkle IOP  (src/iop/blake2s_trivial_iop.rs)
# --------------------------------------------------------------------------------------------
KEY = b"Squeamish Ossifrage"  # :12
PERSONAL = b"Shaftoe"  # :13


def H(data: bytes) -> bytes:
    """BASE_BLAKE2S_PARAMS state .update(data).finalize()  (:8-16, :86-104)."""
    return hashlib.blake2s(data, key=KEY, person=PERSONAL, digest_size=32).digest()


def hash_leaf(F: Field, x: int) -> bytes:
    return H(F.raw_bytes(x))  # :81-92


def hash_node(l: bytes, r: bytes) -> bytes:
    return H(l + r)  # :94-104


def merkle_create(F: Field, leaves: Sequence[int]) -> List[bytes]:
    """Blake2sIopTree::create (:131-219).  Returns `nodes` (heap order, nodes[0] = 32 zero bytes,
    nodes[1] = root, bottom node level at [n/2, n) built from pairs of leaf hashes)."""
    n = len(leaves)
    assert n >= 2 and n & (n - 1) == 0
    nodes = [bytes(32)] * n
    lh = [hash_leaf(F, x) for x in leaves]
    for i in range(n // 2):
        nodes[n // 2 + i] = hash_node(lh[2 * i], lh[2 * i + 1])
    for i in range(n // 2 - 1, 0, -1):
        nodes[i] = hash_node(nodes[2 * i], nodes[2 * i + 1])
    return nodes


def interpret_hash(F: Field, digest: bytes) -> int:
    """interpret_hash (:48-60): read_be into the repr, mask the top limb by
    u64::MAX >> ((256 - CAPACITY) % 64), from_repr.  Returns the plain (non-Montgomery) value."""
    shave = (256 - F.capacity) % 64
    mask = (2**64 - 1) >> shave
    limbs = [0] * F.limbs
    # PrimeFieldRepr::read_be: most significant limb first, each limb big-endian
    for k in range(F.limbs):
        limbs[F.limbs - 1 - k] = int.from_bytes(digest[8 * k : 8 * k + 8], "big")
    limbs[-1] &= mask
    v = sum(l << (64 * i) for i, l in enumerate(limbs))
    if v >= F.p:
        raise ValueError("not in field (reference would panic: 'in a field')")
    return v


def merkle_path(F: Field, nodes: Sequence[bytes], leaves: Sequence[int], index: int) -> List[bytes]:
    """get_path (:251-279)."""
    path = [hash_leaf(F, leaves[index ^ 1])]
    idx = (len(nodes) + index) >> 1  # heap index of the bottom-level node above the leaf pair
    while idx > 1:
        path.append(nodes[idx ^ 1])
        idx >>= 1
    return path


def merkle_verify(F: Field, root: bytes, leaf: int, path: Sequence[bytes], index: int) -> bool:
    """verify (:236-249)."""
    h, idx = hash_leaf(F, leaf), index
    for el in path:
        h = hash_node(h, el) if idx & 1 == 0 else hash_node(el, h)
        idx >>= 1
    return h == root


# --------------------------------------------------------------------------------------------
# FRI commit chain on values  (src/fri/fri_on_values.rs:11-159)
# --------------------------------------------------------------------------------------------
@dataclass
class FriPrototype:
    l0_nodes: List[bytes]
    layer_nodes: List[List[bytes]]
    layer_values: List[List[int]]
    challenges: List[int]
    final_root: bytes
    final_coefficients: List[int]


def fri_commit(F: Field, lde_values: Sequence[int], lde_factor: int, out_coeffs: int) -> FriPrototype:
    p = F.p
    N = len(lde_values)
    log_N = N.bit_length() - 1
    l0 = merkle_create(F, lde_values)
    omega_inv = pow(F.domain_generator(log_N), -1, p)
    two_inv = pow(2, -1, p)
    initial_degree_plus_one = N // lde_factor
    q = initial_degree_plus_one // out_coeffs
    num_steps = q.bit_length() - 1  # log2_floor
    assert num_steps >= 1, "reference panics on roots.pop() when num_steps == 0"
    challenges = [interpret_hash(F, l0[1])]
    c = challenges[0]
    values = list(lde_values)
    layer_nodes, layer_values, roots = [], [], []
    for i in range(num_steps):
        half = len(values) // 2
        stride = 1 << i
        nxt = []
        for idx in range(half):
            f0, f1 = values[idx], values[idx + half]
            even = (f0 + f1) % p
            odd = (f0 - f1) * pow(omega_inv, idx * stride, p) % p
            nxt.append((odd * c + even) * two_inv % p)
        nodes = merkle_create(F, nxt)
        roots.append(nodes[1])
        c = interpret_hash(F, nodes[1])
        challenges.append(c)
        layer_nodes.append(nodes)
        layer_values.append(nxt)
        values = nxt
    challenges.pop()
    final_root = roots.pop()
    log_last = (len(values)).bit_length() - 1
    final = ifft(F, values, log_last)[:out_coeffs]
    return FriPrototype(l0, layer_nodes, layer_values, challenges, final_root, final)


def fri_commit_through_coefficients(F: Field, lde_values: Sequence[int], lde_factor: int, out_coeffs: int):
    """proof_from_lde_through_coefficients (src/fri/mod.rs:156-248) -- the reference's own
    cross-check path: ifft once, fold coefficient pairs a0 + c*a1, re-LDE every layer."""
    N = len(lde_values)
    log_N = N.bit_length() - 1
    l0 = merkle_create(F, lde_values)
    initial_degree_plus_one = N // lde_factor
    num_steps = (initial_degree_plus_one // out_coeffs).bit_length() - 1
    coeffs = ifft(F, lde_values, log_N)[:initial_degree_plus_one]
    challenges = [interpret_hash(F, l0[1])]
    c = challenges[0]
    layer_nodes, layer_values, roots = [], [], []
    for _ in range(num_steps):
        nxt = [(coeffs[2 * k] + c * coeffs[2 * k + 1]) % F.p for k in range(len(coeffs) // 2)]
        vals = lde(F, nxt, len(nxt).bit_length() - 1, lde_factor, coset=False)
        nodes = merkle_create(F, vals)
        roots.append(nodes[1])
        c = interpret_hash(F, nodes[1])
        challenges.append(c)
        layer_nodes.append(nodes)
        layer_values.append(vals)
        coeffs = nxt
    challenges.pop()
    final_root = roots.pop()
    return FriPrototype(l0, layer_nodes, layer_values, challenges, final_root, coeffs)


# --------------------------------------------------------------------------------------------
# Setup work of Prover::new: PrecomputedOmegas (src/precomputations/mod.rs:14-66) and the ALI inverse
# divisors (src/ali/per_register/mod.rs:60-162 dense constraints, :214-227 boundary rows)
# --------------------------------------------------------------------------------------------
def precomputed_omegas(F: Field, log_n: int):
    """-> (omegas[n], coset[n], omegas_inv[n/2])   (:14-66)"""
    n, p = 1 << log_n, F.p
    omega = F.domain_generator(log_n)
    omega_inv = pow(omega, -1, p)
    omegas, u = [], 1
    for _ in range(n):  # :27-37
        omegas.append(u)
        u = u * omega % p
    omegas_inv, u = [], 1
    for _ in range(n // 2):  # :39-49
        omegas_inv.append(u)
        u = u * omega_inv % p
    coset = [v * F.generator % p for v in omegas]  # :51-59
    return omegas, coset, omegas_inv


def inverse_divisor_for_dense_constraint_in_coset(F: Field, log_column: int, log_evaluation: int, start_at: int, span: int,
                                                  num_rows: int):
    """(:60-162) -> (values over g * <evaluation domain>, divisor_degree)."""
    p = F.p
    T, E = 1 << log_column, 1 << log_evaluation
    divisor_degree = T - start_at - (T - num_rows) - span  # :69-75
    w_col, w_eval = F.domain_generator(log_column), F.domain_generator(log_evaluation)
    roots, root = [], 1
    for _ in range(start_at):  # :81-84
        roots.append(root)
        root = root * w_col % p
    last_step = num_rows - span  # :86-91
    root = pow(w_col, last_step, p)
    for _ in range(last_step, T):
        roots.append(root)
        root = root * w_col % p
    out, x = [], F.generator  # :116-129: x = g * w_eval^i, v = x^T - 1
    for _ in range(E):
        out.append((pow(x, T, p) - 1) % p)
        x = x * w_eval % p
    if any(v == 0 for v in out):
        raise ZeroDivisionError("batch_inversion of a vector with a zero")
    out = [pow(v, -1, p) for v in out]  # :133 batch_inversion
    x = F.generator
    for i in range(E):  # :137-158
        d = out[i]
        for r in roots:
            d = d * ((x - r) % p) % p
        out[i] = d
        x = x * w_eval % p
    return out, divisor_degree


def boundary_constraint_inverse_divisor(F: Field, log_column: int, log_evaluation: int, row: int) -> List[int]:
    """(:214-227): q(X) = X - omega^row evaluated on the coset (coset_evaluate_at_domain_for_degree_one), inverted."""
    p = F.p
    w_col, w_eval = F.domain_generator(log_column), F.domain_generator(log_evaluation)
    root = pow(w_col, row, p)
    out, x = [], F.generator
    for _ in range(1 << log_evaluation):
        out.append((x - root) % p)
        x = x * w_eval % p
    return [pow(v, -1, p) for v in out]


# --------------------------------------------------------------------------------------------
# Test-input generator shared by the oracle, the tests and bench.py:
# SplitMix64 stream, 4 limbs per element (limb 0 first), top limb masked to NUM_BITS, rejection
# sampled, and the accepted limbs are used DIRECTLY as the Montgomery representation.
# --------------------------------------------------------------------------------------------
def splitmix64(state: int):
    while True:
        state = (state + 0x9E3779B97F4A7C15) & (2**64 - 1)
        z = state
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & (2**64 - 1)
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & (2**64 - 1)
        yield z ^ (z >> 31)


def random_mont_elements(F: Field, count: int, seed: int = 0x3DBE62598D313D76) -> List[int]:
    """Returns Montgomery representations (raw 256-bit integers < p)."""
    gen = splitmix64(seed)
    mask = (2**64 - 1) >> (64 * F.limbs - F.num_bits)
    out = []
    while len(out) < count:
        limbs = [next(gen) for _ in range(F.limbs)]
        limbs[-1] &= mask
        v = sum(l << (64 * i) for i, l in enumerate(limbs))
        if v < F.p:
            out.append(v)
    return out
