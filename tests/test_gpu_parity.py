"""GPU parity: the CUDA path, called through the C ABI (host-pointer entry points via the Python
mirror of the reference interface, and the device-pointer entry points), against the CPU oracle on
the same seeded inputs -- bit-exact -- plus size-independent properties at BASELINE.json's sizes.
The shapes follow the reference's own tests."""
import ctypes as C

import numpy as np
import pytest

from conftest import FIELD_IDS, FIELDS

pytestmark = pytest.mark.gpu


def structured_vectors(O, fid, n):
    c = O.field_constants(fid)
    p = O.limbs_to_int(c["p"])
    zero = np.zeros((n, 4), np.uint64)
    one = np.tile(c["r"], (n, 1))
    d0 = zero.copy(); d0[0] = c["r"]
    d1 = zero.copy(); d1[min(1, n - 1)] = c["r"]
    pm1 = np.tile(O.int_to_limbs(p - 1), (n, 1))
    return {"zero": zero, "one": one, "delta0": d0, "delta1": d1, "p-1": pm1}


# ----------------------------------------------------------------------------------------------
# field arithmetic on the device
# ----------------------------------------------------------------------------------------------
@pytest.mark.parametrize("fid", FIELDS, ids=FIELD_IDS)
def test_elementwise_field_ops(hodor, oracle, fid):
    from hodor_b200 import _ffi
    from hodor_b200.field import _p

    n = 1 << 14
    a, b = oracle.random_elements(fid, n, 1), oracle.random_elements(fid, n, 2)
    c = oracle.field_constants(fid)
    p = oracle.limbs_to_int(c["p"])
    edge = oracle.ints_to_array([0, 1, p - 1, oracle.limbs_to_int(c["r"]), p - 2, 2, (p - 1) // 2, (p + 1) // 2])
    a[:8], b[:8] = edge, edge[::-1]
    a[8:16], b[8:16] = edge, edge
    a[16:24], b[16:24] = edge, np.tile(oracle.int_to_limbs(p - 1), (8, 1))
    out = np.zeros_like(a)
    for op, fn in ((0, oracle.mul), (1, oracle.add), (2, oracle.sub)):
        _ffi.check(_ffi.lib.hodor_cuda_elementwise(op, _p(a), _p(b), _p(out), C.c_uint64(n), fid))
        assert np.array_equal(out, fn(fid, a, b)), f"op {op}"
    _ffi.check(_ffi.lib.hodor_cuda_elementwise(3, _p(a), _p(b[:1].copy()), _p(out), C.c_uint64(n), fid))
    assert np.array_equal(out, oracle.mul(fid, a, np.tile(b[0], (n, 1))))


@pytest.mark.parametrize("fid", FIELDS, ids=FIELD_IDS)
def test_polynomial_scalar_ops(hodor, oracle, fid):
    """negate, add_constant, add_assign_scaled, square, pow (src/polynomials/mod.rs:72-83, 654-669,
    744-771, 831-858) against the oracle's scalar field arithmetic."""
    n = 1 << 10
    a, b = oracle.random_elements(fid, n, 11), oracle.random_elements(fid, n, 12)
    c = oracle.field_constants(fid)
    p = oracle.limbs_to_int(c["p"])
    a[:4] = oracle.ints_to_array([0, 1, p - 1, oracle.limbs_to_int(c["r"])])
    s = b[5]
    W = hodor.Worker()
    x = hodor.Polynomial.from_values(fid, a); x.negate(W)
    assert np.array_equal(x.as_ref(), oracle.sub(fid, np.zeros_like(a), a))
    x = hodor.Polynomial.from_values(fid, a); x.add_constant(W, s)
    assert np.array_equal(x.as_ref(), oracle.add(fid, a, np.tile(s, (n, 1))))
    x = hodor.Polynomial.from_values(fid, a); x.add_assign_scaled(W, hodor.Polynomial.from_values(fid, b), s)
    assert np.array_equal(x.as_ref(), oracle.add(fid, a, oracle.mul(fid, b, np.tile(s, (n, 1)))))
    x = hodor.Polynomial.from_values(fid, a); x.square(W)
    assert np.array_equal(x.as_ref(), oracle.mul(fid, a, a))
    x = hodor.Polynomial.from_values(fid, a); x.pow(W, 2)
    assert np.array_equal(x.as_ref(), oracle.mul(fid, a, a))
    for e in (0, 1, 3, 65537, (1 << 64) - 1):
        x = hodor.Polynomial.from_values(fid, a[:64]); x.pow(W, e)
        assert np.array_equal(x.as_ref(), np.stack([oracle.pow_(fid, v, e) for v in a[:64]])), e
    x = hodor.Polynomial.from_values(fid, a); x.scale(W, s)
    assert np.array_equal(x.as_ref(), oracle.mul(fid, a, np.tile(s, (n, 1))))


@pytest.mark.parametrize("fid", FIELDS, ids=FIELD_IDS)
def test_fixed_operand_multiplier_selftest(hodor, fid):
    """Field::mul_pre (the multiplier of every table multiply) against the Montgomery multiplier on
    the device, including with the guard threshold forced low so that the out-of-line carry fix-up
    (taken about 3 times in 10^9 multiplies in production) runs on half of all multiplies."""
    from hodor_b200 import _ffi

    assert _ffi.check(_ffi.lib.hodor_cuda_selftest_mul_pre(fid)) == 0


@pytest.mark.parametrize("fid", FIELDS, ids=FIELD_IDS)
def test_batch_inversion_matches_oracle(hodor, oracle, fid):
    """Polynomial::batch_inversion (src/polynomials/mod.rs:889-954; reference test :959-985):
    every tree shape of the kernel (K = 16 per level), ragged sizes included."""
    from hodor_b200 import _ffi
    from hodor_b200.field import _p

    for n in (1, 2, 5, 16, 17, 255, 256, 257, 4097, 1 << 16, (1 << 18) + 3):
        a = oracle.random_elements(fid, n, seed=900 + n % 97)
        want = oracle.batch_inversion(fid, a)
        got = a.copy()
        _ffi.check(_ffi.lib.hodor_cuda_batch_inversion(_p(got), C.c_uint64(n), fid))
        assert np.array_equal(got, want), f"n={n}"
    # a zero anywhere: Err(SynthesisError::Error), vector untouched (:919)
    a = oracle.random_elements(fid, 1000, seed=7)
    a[613] = 0
    got = a.copy()
    assert _ffi.lib.hodor_cuda_batch_inversion(_p(got), C.c_uint64(1000), fid) == _ffi.ERR_NOT_INVERTIBLE
    assert np.array_equal(got, a)
    poly = hodor.Polynomial.from_values(fid, a)
    with pytest.raises(_ffi.SynthesisError):
        poly.batch_inversion(hodor.Worker())
    assert oracle.batch_inversion(fid, a) is None
    a[613] = a[0]
    poly = hodor.Polynomial.from_values(fid, a[:512])
    poly.batch_inversion(hodor.Worker())
    assert np.array_equal(poly.as_ref(), oracle.batch_inversion(fid, a[:512]))


@pytest.mark.parametrize("fid", FIELDS, ids=FIELD_IDS)
def test_evaluate_at_matches_oracle(hodor, oracle, fid):
    """Polynomial::evaluate_at (src/polynomials/mod.rs:685-711)."""
    from hodor_b200 import _ffi
    from hodor_b200.field import _p

    z = oracle.random_elements(fid, 1, seed=31)[0]
    for n in (1, 2, 7, 8, 9, 255, 256, 257, 2049, 1 << 16, (1 << 16) + 1, (1 << 20) + 5):
        a = oracle.random_elements(fid, n, seed=300 + n % 89)
        out = np.zeros(4, np.uint64)
        _ffi.check(_ffi.lib.hodor_cuda_evaluate_at(_p(a), C.c_uint64(n), _p(z), _p(out), fid))
        assert np.array_equal(out, oracle.evaluate_at(fid, a, z)), f"n={n}"
    a = oracle.random_elements(fid, 1 << 12, seed=3)
    poly = hodor.Polynomial.from_coeffs(fid, a)
    assert np.array_equal(poly.evaluate_at(hodor.Worker(), z), oracle.evaluate_at(fid, a, z))
    # a coset LDE value IS the polynomial evaluated at g * w^idx
    lde = hodor.Polynomial.from_coeffs(fid, a).coset_lde(hodor.Worker(), 4)
    w = oracle.domain_generator(fid, 14)
    g = oracle.field_constants(fid)["generator"]
    for idx in (0, 1, 5, (1 << 14) - 1):
        x = oracle.mul(fid, g, oracle.pow_(fid, w, idx))[0]
        assert np.array_equal(poly.evaluate_at(hodor.Worker(), x), lde.as_ref()[idx])


@pytest.mark.parametrize("fid", FIELDS, ids=FIELD_IDS)
def test_device_polynomial_matches_host_polynomial(hodor, oracle, fid):
    """DevicePolynomial (vector resident in HBM, `_dev` entry points) gives the bits of the
    host-pointer Polynomial on a chain shaped like the ALI / DEEP steps: LDE, elementwise forms, batch
    inversion, transforms back, point evaluation."""
    from hodor_b200 import _ffi
    from hodor_b200.device import DevicePolynomial

    W = hodor.Worker()
    a, b = oracle.random_elements(fid, 1 << 11, seed=81), oracle.random_elements(fid, 1 << 11, seed=82)
    s, z = b[3], a[7]

    def chain(P):
        x = P.from_coeffs(fid, a).coset_lde(W, 4)
        y = P.from_coeffs(fid, b).coset_lde(W, 4)
        x.add_assign_scaled(W, y, s)
        x.add_constant(W, s)
        x.square(W)
        x.pow(W, 3)
        x.negate(W)
        x.batch_inversion(W)
        x.mul_assign(W, y)
        x.sub_assign(W, y)
        x.scale(W, z)
        x.add_assign(W, y)
        c = x.icoset_fft(W)
        return c, c.evaluate_at(W, z)

    hc, hv = chain(hodor.Polynomial)
    dc, dv = chain(DevicePolynomial)
    assert np.array_equal(dc.to_host(), hc.as_ref()) and np.array_equal(dv, hv)
    back = dc.clone().fft(W).ifft(W)
    assert np.array_equal(back.to_host(), hc.as_ref())
    bad = a.copy()
    bad[100] = 0
    dz = DevicePolynomial.from_values(fid, bad)
    with pytest.raises(_ffi.SynthesisError):
        dz.batch_inversion(W)
    assert np.array_equal(dz.to_host(), bad)


def test_batch_inversion_2p24_properties(hodor, oracle):
    """BASELINE-size check without a 2^24 CPU pass: a * a^-1 == 1 everywhere, and inverting twice
    returns the input."""
    from hodor_b200 import _ffi
    from hodor_b200.field import _p

    fid, n = 0, 1 << 24
    a = oracle.random_elements(fid, n, seed=24)
    inv = a.copy()
    _ffi.check(_ffi.lib.hodor_cuda_batch_inversion(_p(inv), C.c_uint64(n), fid))
    prod = np.zeros_like(a)
    _ffi.check(_ffi.lib.hodor_cuda_elementwise(0, _p(a), _p(inv), _p(prod), C.c_uint64(n), fid))
    assert np.array_equal(prod, np.tile(oracle.field_constants(fid)["r"], (n, 1)))
    inv_once = inv[[0, 1, n // 2, n - 1]].copy()
    _ffi.check(_ffi.lib.hodor_cuda_batch_inversion(_p(inv), C.c_uint64(n), fid))
    assert np.array_equal(inv, a)
    idx = range(4)
    a = a[[0, 1, n // 2, n - 1]]
    for i in idx:
        assert np.array_equal(inv_once[i], oracle.inverse(fid, a[i]))


# ----------------------------------------------------------------------------------------------
# NTT
# ----------------------------------------------------------------------------------------------
@pytest.mark.parametrize("fid", FIELDS, ids=FIELD_IDS)
@pytest.mark.parametrize("log_n", list(range(0, 17)))
def test_ntt_matches_oracle(hodor, oracle, fid, log_n):
    """best_fft == serial_fft for every size class: single block (<= 2^11) and 2 passes."""
    n = 1 << log_n
    a = oracle.random_elements(fid, n, seed=100 + log_n)
    omega = oracle.domain_generator(fid, log_n)
    want = oracle.serial_fft(fid, a, omega, log_n)
    got = hodor.Polynomial.from_coeffs(fid, a).fft().as_ref()
    assert np.array_equal(got, want)
    # arbitrary primitive root through the raw best_fft entry point: omega^-1
    from hodor_b200 import _ffi
    from hodor_b200.field import _p
    winv = oracle.inverse(fid, omega)
    buf = a.copy()
    _ffi.check(_ffi.lib.hodor_cuda_ntt(_p(buf), log_n, _p(winv), fid))
    assert np.array_equal(buf, oracle.serial_fft(fid, a, winv, log_n))


@pytest.mark.parametrize("fid", FIELDS, ids=FIELD_IDS)
def test_ntt_structured_inputs(hodor, oracle, fid):
    for log_n in (3, 9, 12, 14):
        omega = oracle.domain_generator(fid, log_n)
        for name, v in structured_vectors(oracle, fid, 1 << log_n).items():
            got = hodor.Polynomial.from_coeffs(fid, v).fft().as_ref()
            assert np.array_equal(got, oracle.serial_fft(fid, v, omega, log_n)), (name, log_n)


@pytest.mark.parametrize("log_n", [18, 19, 20, 21])
def test_ntt_large_matches_oracle_multicore(hodor, oracle, log_n):
    """2 and 3 HBM passes (all digit widths 6..9 are covered by 12..21) vs the reference's
    parallel_fft restatement on all host cores."""
    fid = 0
    a = oracle.random_elements(fid, 1 << log_n, seed=log_n)
    omega = oracle.domain_generator(fid, log_n)
    want = oracle.best_fft(fid, a, omega, log_n)
    assert np.array_equal(hodor.Polynomial.from_coeffs(fid, a).fft().as_ref(), want)


def test_ntt_rejects_non_root(hodor, oracle):
    from hodor_b200 import _ffi
    from hodor_b200.field import _p
    a = oracle.random_elements(0, 16, 1)
    bad = oracle.domain_generator(0, 5)  # order 32, not 16
    assert _ffi.lib.hodor_cuda_ntt(_p(a), 4, _p(bad), 0) == _ffi.ERR_NOT_A_ROOT


@pytest.mark.parametrize("fid", FIELDS, ids=FIELD_IDS)
def test_ifft_and_coset_roundtrips(hodor, oracle, fid):
    """test_worker_size's round trip (src/fft/mod.rs:281) and Polynomial::{ifft, icoset_fft}."""
    for log_n in (0, 1, 4, 10, 13, 16):
        a = oracle.random_elements(fid, 1 << log_n, seed=7 + log_n)
        vals = hodor.Polynomial.from_coeffs(fid, a).fft()
        assert np.array_equal(vals.as_ref(), oracle.fft(fid, a, log_n))
        assert np.array_equal(vals.clone().ifft().as_ref(), a)
        cvals = hodor.Polynomial.from_coeffs(fid, a).coset_fft()
        assert np.array_equal(cvals.as_ref(), oracle.fft(fid, a, log_n, coset=True))
        assert np.array_equal(cvals.clone().icoset_fft().as_ref(), a)
        assert np.array_equal(hodor.Polynomial.from_values(fid, a).ifft().as_ref(), oracle.ifft(fid, a, log_n))
        assert np.array_equal(hodor.Polynomial.from_values(fid, a).icoset_fft().as_ref(),
                              oracle.ifft(fid, a, log_n, coset=True))


@pytest.mark.parametrize("fid", FIELDS, ids=FIELD_IDS)
def test_distribute_powers(hodor, oracle, fid):
    g = oracle.field_constants(fid)["generator"]
    for n in (1, 2, 5, 1000, 1 << 15):
        a = oracle.random_elements(fid, n, seed=n)
        pad = 1 << max(0, (n - 1).bit_length())
        p = hodor.Polynomial.from_coeffs(fid, a)
        p.distribute_powers(None, g)
        want = oracle.distribute_powers(fid, np.concatenate([a, np.zeros((pad - n, 4), np.uint64)]), g)
        assert np.array_equal(p.as_ref(), want)


# ----------------------------------------------------------------------------------------------
# LDE
# ----------------------------------------------------------------------------------------------
@pytest.mark.parametrize("fid", FIELDS, ids=FIELD_IDS)
@pytest.mark.parametrize("coset", [False, True])
def test_lde_matches_oracle(hodor, oracle, fid, coset):
    """test_lde_correctness / test_coset_lde_correctness / test_various_ldes shapes
    (src/polynomials/mod.rs:988-1128), incl. the reference's n = 4, L = 16."""
    for log_n, L in [(2, 16), (0, 2), (1, 4), (3, 1), (5, 2), (8, 8), (11, 4), (12, 8), (13, 16), (14, 2), (15, 32)]:
        a = oracle.random_elements(fid, 1 << log_n, seed=31 * log_n + L)
        p = hodor.Polynomial.from_coeffs(fid, a)
        got = (p.coset_lde(None, L) if coset else p.lde(None, L)).as_ref()
        assert np.array_equal(got, oracle.lde(fid, a, log_n, L, coset)), (log_n, L)


def test_lde_batch_pipelined_matches_single_calls(hodor, oracle):
    """hodor_cuda_lde_batch: N polynomials through the double-buffered copy/compute pipeline give the
    bits of N separate calls (and of the oracle); batch sizes around the double-buffer depth."""
    fid, log_n, L = 0, 13, 4
    polys = [oracle.random_elements(fid, 1 << log_n, seed=60 + i) for i in range(5)]
    for count in (0, 1, 2, 3, 5):
        for coset in (False, True):
            got = hodor.lde_batch([hodor.Polynomial.from_coeffs(fid, a) for a in polys[:count]], hodor.Worker(), L, coset)
            assert len(got) == count
            for a, g in zip(polys, got):
                assert np.array_equal(g.as_ref(), oracle.lde(fid, a, log_n, L, coset))
    big = [oracle.random_elements(fid, 1 << 20, seed=70 + i) for i in range(3)]
    got = hodor.lde_batch([hodor.Polynomial.from_coeffs(fid, a) for a in big], hodor.Worker(), 8, True)
    for a, g in zip(big, got):
        assert np.array_equal(g.as_ref(), hodor.Polynomial.from_coeffs(fid, a).coset_lde(hodor.Worker(), 8).as_ref())


def test_lde_2p20_x8_bit_exact_and_horner(hodor, oracle, pymodel):
    """configs[1] shape at a size the CPU oracle finishes in seconds: coset LDE, blowup 8."""
    fid, log_n, L = 0, 18, 8
    a = oracle.random_elements(fid, 1 << log_n, seed=2024)
    got = hodor.Polynomial.from_coeffs(fid, a).coset_lde(None, L).as_ref()
    assert np.array_equal(got, oracle.lde(fid, a, log_n, L, True))
    # multi-coset LDE == coset NTT of the zero-padded vector (reference identity)
    padded = np.zeros(((1 << log_n) * L, 4), np.uint64)
    padded[: 1 << log_n] = a
    assert np.array_equal(got, hodor.Polynomial.from_coeffs(fid, padded).coset_fft().as_ref())


def test_lde_2p24_x8_properties(hodor, oracle, pymodel):
    """BASELINE.json configs[1] at full size (2^24 -> 2^27, 4 GiB): size-independent checks.
    (a) Horner spot checks of 24 outputs in Python big ints; (b) decimating the LDE by 8 gives the
    plain coset NTT; (c) icoset_fft of that recovers the coefficients."""
    F = pymodel.BLS12_381_FR
    fid, log_n, L = 0, 24, 8
    n = 1 << log_n
    a = oracle.random_elements(fid, n, seed=99)
    # keep the Horner check affordable: make the polynomial sparse-ish but full degree
    a[1 << 12 : n - (1 << 12)] = 0
    lde = hodor.Polynomial.from_coeffs(fid, a).coset_lde(None, L).as_ref()
    assert lde.shape[0] == n * L
    w = F.domain_generator(log_n + 3)
    nz = [i for i in list(range(1 << 12)) + list(range(n - (1 << 12), n))]
    coeffs = {i: F.from_mont(oracle.limbs_to_int(a[i])) for i in nz}
    rng = np.random.default_rng(5)
    for idx in [0, 1, n * L - 1] + [int(x) for x in rng.integers(0, n * L, 21)]:
        x = F.generator * pow(w, idx, F.p) % F.p
        lo = sum(c * pow(x, i, F.p) for i, c in coeffs.items() if i < (1 << 12)) % F.p
        hi = sum(c * pow(x, i - (n - (1 << 12)), F.p) for i, c in coeffs.items() if i >= (1 << 12)) % F.p
        want = (lo + hi * pow(x, n - (1 << 12), F.p)) % F.p
        assert F.from_mont(oracle.limbs_to_int(lde[idx])) == want, idx
    coset0 = np.ascontiguousarray(lde[::L])
    back = hodor.Polynomial.from_values(fid, coset0).icoset_fft().as_ref()
    assert np.array_equal(back, a)


def test_four_pass_plan_2p25(hodor, oracle):
    """Transforms of 2^25 and up run FOUR passes over HBM (2^25 = 7+6+6+6: two middle digits in the
    last pass's output index).  Checked against an independent code path, the reference's own identity
    (test_lde_correctness, src/polynomials/mod.rs:988-1034): the NTT of a zero-padded vector equals the
    multi-coset LDE -- which for 2^24 x 2 runs the three-pass plan per coset -- and by round trips of
    the plain and coset transforms (input scaling and output scaling on the four-pass path)."""
    fid, log_n = 0, 25
    n = 1 << log_n
    a = oracle.random_elements(fid, n // 2, seed=25)
    lde = hodor.Polynomial.from_coeffs(fid, a).lde(hodor.Worker(), 2).as_ref()
    padded = np.zeros((n, 4), np.uint64)
    padded[: n // 2] = a
    full = hodor.Polynomial.from_coeffs(fid, padded).fft(hodor.Worker())
    assert np.array_equal(full.as_ref(), lde)
    back = full.ifft(hodor.Worker())
    assert np.array_equal(back.as_ref(), padded)
    del lde, full
    x = oracle.random_elements(fid, n, seed=26)
    y = hodor.Polynomial.from_coeffs(fid, x).coset_fft(hodor.Worker())
    assert not np.array_equal(y.as_ref()[:1024], x[:1024])
    assert np.array_equal(y.icoset_fft(hodor.Worker()).as_ref(), x)


# ----------------------------------------------------------------------------------------------
# Merkle oracle
# ----------------------------------------------------------------------------------------------
@pytest.mark.parametrize("fid", FIELDS, ids=FIELD_IDS)
@pytest.mark.parametrize("log_n", [1, 2, 3, 4, 7, 11, 12, 13, 14, 15, 16, 17])
def test_merkle_matches_oracle(hodor, oracle, fid, log_n):
    leaves = oracle.random_elements(fid, 1 << log_n, seed=500 + log_n)
    tree = hodor.Blake2sIopTree.create(fid, leaves)
    want = oracle.merkle_create(fid, leaves)
    assert np.array_equal(tree.nodes, want)
    assert np.array_equal(tree.get_challenge_scalar_from_root(), oracle.interpret_hash(fid, want[1].tobytes()))


def test_make_small_tree_and_iop(hodor, oracle):
    """make_small_tree / make_small_iop (src/iop/blake2s_trivial_iop.rs:377-408) on the GPU tree."""
    fid = 2
    one = oracle.field_constants(fid)["r"]
    tree = hodor.Blake2sIopTree.create(fid, np.tile(one, (16, 1)))
    assert tree.get_root().hex() == "fdf489862b4402468d94f026c014e1ca0129f421a55ce5e1df838eab8eefbd22"
    vals = oracle.to_mont(fid, oracle.ints_to_array([1 << i for i in range(64)]))
    iop = hodor.TrivialBlake2sIOP.create(fid, vals)
    root = iop.get_root()
    for i in range(64):
        q = iop.query(i, vals)
        assert hodor.TrivialBlake2sIOP.verify_query(q, root), f"invalid query for leaf {i}"
        assert q.path() == oracle.merkle_path(fid, iop.tree.nodes, vals, i)
    with pytest.raises(AssertionError):
        hodor.Blake2sIopTree.create(fid, vals[:48])  # assert!(num_leafs == next_power_of_two)


def test_merkle_2p22_vs_oracle_and_2p24_paths(hodor, oracle):
    fid = 0
    leaves = oracle.random_elements(fid, 1 << 22, seed=4242)
    tree = hodor.Blake2sIopTree.create(fid, leaves)
    assert np.array_equal(tree.nodes, oracle.merkle_create(fid, leaves))
    # full size of configs[2]'s first tree: every sampled authentication path must verify
    leaves = oracle.random_elements(fid, 1 << 24, seed=4243)
    iop = hodor.TrivialBlake2sIOP.create(fid, leaves)
    root = iop.get_root()
    assert not iop.tree.nodes[0].any()
    rng = np.random.default_rng(8)
    for i in [0, 1, (1 << 24) - 1] + [int(x) for x in rng.integers(0, 1 << 24, 13)]:
        assert hodor.TrivialBlake2sIOP.verify_query(iop.query(i, leaves), root)


# ----------------------------------------------------------------------------------------------
# FRI commit chain
# ----------------------------------------------------------------------------------------------
def assert_fri_equal(O, fid, proto, want):
    assert proto.get_roots() == want.roots()
    assert np.array_equal(proto.challenges, want.challenges)
    assert proto.get_final_root() == want.final_root
    assert np.array_equal(proto.final_coefficients, want.final_coefficients)
    assert np.array_equal(proto.l0_commitment.nodes, want.l0_nodes)
    for i, v in enumerate(proto.intermediate_values):
        assert np.array_equal(v.as_ref(), want.layer_values[i]), f"layer {i} values"
        assert np.array_equal(proto.intermediate_commitments[i].nodes, want.layer_nodes[i]), f"layer {i} nodes"


@pytest.mark.parametrize("fid", FIELDS, ids=FIELD_IDS)
def test_fri_commit_matches_oracle(hodor, oracle, fid):
    """proof_from_lde_by_values on random values (the commit chain is defined for any input)."""
    for log_n, L, oc in [(2, 2, 1), (4, 4, 2), (6, 2, 1), (10, 8, 1), (12, 16, 4), (14, 8, 1), (16, 16, 1)]:
        v = oracle.random_elements(fid, 1 << log_n, seed=900 + log_n)
        proto = hodor.NaiveFriIop.proof_from_lde(hodor.Polynomial.from_values(fid, v), L, oc, hodor.Worker())
        assert_fri_equal(oracle, fid, proto, oracle.fri_commit(fid, v, L, oc))


def test_fri_host_signature_matches_oracle(hodor, oracle):
    """The reference-shaped entry point (everything copied to caller buffers)."""
    from hodor_b200 import _ffi
    from hodor_b200._ffi import u8p, u64p
    from hodor_b200.field import _p
    fid, log_n, L, oc = 0, 12, 8, 2
    n = 1 << log_n
    v = oracle.random_elements(fid, n, seed=1234)
    steps = oracle.fri_num_steps(n, L, oc)
    l0 = np.zeros((n, 32), np.uint8)
    ln = [np.zeros((n >> (i + 1), 32), np.uint8) for i in range(steps)]
    lv = [np.zeros((n >> (i + 1), 4), np.uint64) for i in range(steps)]
    ch = np.zeros((steps, 4), np.uint64)
    fr = np.zeros(32, np.uint8)
    fc = np.zeros((oc, 4), np.uint64)
    lnp = (u8p * steps)(*[a.ctypes.data_as(u8p) for a in ln])
    lvp = (u64p * steps)(*[_p(a) for a in lv])
    rc = _ffi.lib.hodor_cuda_fri_commit_host(_p(v), C.c_uint64(n), L, oc, l0.ctypes.data_as(u8p), lnp, lvp, _p(ch),
                                             fr.ctypes.data_as(u8p), _p(fc), fid)
    assert rc == steps, _ffi.last_error()
    want = oracle.fri_commit(fid, v, L, oc)
    assert np.array_equal(l0, want.l0_nodes) and np.array_equal(ch, want.challenges)
    assert fr.tobytes() == want.final_root and np.array_equal(fc, want.final_coefficients)
    for i in range(steps):
        assert np.array_equal(ln[i], want.layer_nodes[i]) and np.array_equal(lv[i], want.layer_values[i])


def test_one_fri_step_shape_and_low_degree(hodor, oracle):
    """test_one_fri_step (src/fri/mod.rs:252-331) over bn256.rs's field: coefficients 1, 2, 4, 8,
    lde_factor 4, output 2; the fold must equal the LDE of (a0 + c a1, a2 + c a3)."""
    from hodor_b200 import field as f
    fid = 0
    coeffs = oracle.to_mont(fid, oracle.ints_to_array([1, 2, 4, 8]))
    lde_values = hodor.Polynomial.from_coeffs(fid, coeffs).lde(hodor.Worker(), 4)
    proto = hodor.NaiveFriIop.proof_from_lde_by_values(lde_values, 4, 2, hodor.Worker())
    c = proto.challenges[0]
    new_coeffs = np.stack([f.add(fid, coeffs[0], f.mul(fid, c, coeffs[1])), f.add(fid, coeffs[2], f.mul(fid, c, coeffs[3]))])
    assert np.array_equal(proto.final_coefficients, new_coeffs)
    next_lde = hodor.Polynomial.from_coeffs(fid, new_coeffs).lde(hodor.Worker(), 4)
    assert np.array_equal(proto.intermediate_values[0].as_ref(), next_lde.as_ref())
    assert proto.intermediate_commitments[0] == hodor.TrivialBlake2sIOP.create(fid, next_lde.as_ref())
    # query production (src/fri/query_producer.rs): every query verifies against its layer's root
    for start in (1, 3, 7, 12):
        proof = hodor.NaiveFriIop.prototype_into_proof(proto, lde_values, start)
        assert len(proof.queries) == 2 * len(proof.roots)
        for k, q in enumerate(proof.queries):
            assert hodor.TrivialBlake2sIOP.verify_query(q, proof.roots[k // 2])


def test_fri_argument_errors(hodor, oracle):
    v = hodor.Polynomial.from_values(0, oracle.random_elements(0, 16, 1))
    with pytest.raises(hodor.HodorError):
        hodor.NaiveFriIop.proof_from_lde(v, 16, 1, None)  # zero folding steps: the reference panics
    with pytest.raises(hodor.HodorError):
        hodor.NaiveFriIop.proof_from_lde(v, 3, 1, None)  # assert!(lde_factor.is_power_of_two())


def test_fri_chain_2p24_properties(hodor, oracle):
    """BASELINE.json configs[2] at full size: N = 2^24, blowup 8, one output coefficient (21 layers).
    Size-independent checks: (a) every challenge is interpret_hash of the previous root; (b) sampled
    fold outputs recomputed on the host from the layer below; (c) every query of a produced proof
    verifies; (d) the first three and last three layers are bit-exact against the CPU oracle run on
    those layers' inputs; (e) final coefficient == iNTT of the last layer."""
    from hodor_b200 import field as f
    fid, log_n, L, oc = 0, 24, 8, 1
    n = 1 << log_n
    v = oracle.random_elements(fid, n, seed=777)
    proto = hodor.NaiveFriIop.proof_from_lde(hodor.Polynomial.from_values(fid, v), L, oc, None)
    steps = proto.num_steps
    assert steps == 21
    roots = proto.get_roots()
    for i in range(steps):
        assert np.array_equal(proto.challenges[i], oracle.interpret_hash(fid, roots[i]))
    winv = oracle.inverse(fid, oracle.domain_generator(fid, log_n))
    two_inv = oracle.inverse(fid, oracle.to_mont(fid, oracle.ints_to_array([2]))[0])
    rng = np.random.default_rng(3)
    values = [v] + [p.as_ref() for p in proto.intermediate_values]
    for i in range(steps):
        half = n >> (i + 1)
        for idx in {0, half - 1, *[int(x) for x in rng.integers(0, half, 3)]}:
            f0, f1 = values[i][idx], values[i][idx + half]
            odd = f.mul(fid, f.sub(fid, f0, f1), oracle.pow_(fid, winv, idx << i))
            want = f.mul(fid, f.add(fid, f.mul(fid, odd, proto.challenges[i]), f.add(fid, f0, f1)), two_inv)
            assert np.array_equal(values[i + 1][idx], want), (i, idx)
    for layer in (1, 2, steps - 2, steps - 1, steps):
        assert np.array_equal(proto.intermediate_commitments[layer - 1].nodes, oracle.merkle_create(fid, values[layer]))
    small = oracle.fri_commit(fid, values[steps - 3], L, oc)  # the last three folds as their own chain
    assert small.roots()[1:] == roots[steps - 2 :]
    assert np.array_equal(proto.final_coefficients, oracle.ifft(fid, values[steps], 3)[:oc])
    proof = proto.produce_proof(None, 123457)
    for k, q in enumerate(proof.queries):
        assert hodor.TrivialBlake2sIOP.verify_query(q, proof.roots[k // 2])


# ----------------------------------------------------------------------------------------------
# device-pointer entry points
# ----------------------------------------------------------------------------------------------
def test_device_entry_points(hodor, oracle):
    import torch
    from hodor_b200 import device as dev
    fid, log_n, L = 0, 14, 8
    n = 1 << log_n
    a = oracle.random_elements(fid, n, seed=66)
    d_a = dev.to_device(a)
    d_out = dev.empty_elems(n * L)
    dev.lde(d_a, log_n, 3, True, d_out, fid)
    want = oracle.lde(fid, a, log_n, L, True)
    assert np.array_equal(dev.to_host(d_out), want)
    d_nodes = torch.empty((n * L, 4), dtype=torch.int64, device="cuda")
    d_root = torch.empty(4, dtype=torch.int64, device="cuda")
    d_chal = torch.empty(4, dtype=torch.int64, device="cuda")
    dev.merkle_build(d_out, n * L, d_nodes, fid, d_root, d_chal)
    nodes = oracle.merkle_create(fid, want)
    assert np.array_equal(d_nodes.cpu().numpy().view(np.uint8).reshape(-1, 32), nodes)
    assert d_root.cpu().numpy().tobytes() == nodes[1].tobytes()
    assert np.array_equal(d_chal.cpu().numpy().view(np.uint64), oracle.interpret_hash(fid, nodes[1].tobytes()))
    d_next = dev.empty_elems(n * L // 2)
    dev.fri_fold(d_out, n * L, n * L, 0, d_chal, d_next, fid)
    assert np.array_equal(dev.to_host(d_next), oracle.fri_commit(fid, want, L, 1).layer_values[0])
    # in place forward / inverse on the device
    d_b = dev.to_device(a)
    dev.fft(d_b, d_b, log_n, False, fid)
    assert np.array_equal(dev.to_host(d_b), oracle.fft(fid, a, log_n))
    dev.ifft(d_b, d_b, log_n, False, fid)
    assert np.array_equal(dev.to_host(d_b), a)
    proto = dev.fri_commit(d_out, L, 1, fid)
    assert proto.get_roots() == oracle.fri_commit(fid, want, L, 1).roots()
    assert dev.launch_count() > 0


@pytest.mark.parametrize("world", [1, 2, 4, 8])
def test_sharded_steps_single_gpu(hodor, oracle, world):
    """The two local steps of the four-step NTT, with the all-to-all emulated on one GPU: all
    `world` ranks' slices are processed in turn.  (The NCCL path is exercised by bench.py.)"""
    import torch
    from hodor_b200 import device as dev
    from hodor_b200.sharded import CudaBackend, gather_output, scatter_input
    fid, log_n = 0, 16
    log_g = world.bit_length() - 1
    a = oracle.random_elements(fid, 1 << log_n, seed=55)
    omega = oracle.domain_generator(fid, log_n)
    be = CudaBackend()
    cols = [be.shard_cols(dev.to_device(scatter_input(a, world, g)), log_n, log_g, g, omega, fid) for g in range(world)]
    m = (1 << log_n) // world
    chunk = m // world
    outs = []
    for h in range(world):
        recv = torch.cat([cols[g][h * chunk : (h + 1) * chunk] for g in range(world)]).contiguous()
        outs.append(dev.to_host(be.shard_rows(recv, log_n, log_g, h, omega, fid)))
    assert np.array_equal(gather_output(outs), oracle.serial_fft(fid, a, omega, log_n))


def test_sharded_steps_large_local_transform(hodor, oracle):
    """Same emulation at 2^26 over 2 ranks: each rank's column step is a 2^25-point transform, i.e. the
    four-pass plan with the per-rank output scaling fused into its last pass; checked against the direct
    single-GPU NTT of the whole vector."""
    import torch
    from hodor_b200 import device as dev
    from hodor_b200.sharded import CudaBackend, gather_output, scatter_input
    fid, log_n, world = 0, 26, 2
    log_g = 1
    a = oracle.random_elements(fid, 1 << log_n, seed=56)
    omega = oracle.domain_generator(fid, log_n)
    be = CudaBackend()
    cols = [be.shard_cols(dev.to_device(scatter_input(a, world, g)), log_n, log_g, g, omega, fid) for g in range(world)]
    m = (1 << log_n) // world
    chunk = m // world
    outs = []
    for h in range(world):
        recv = torch.cat([cols[g][h * chunk : (h + 1) * chunk] for g in range(world)]).contiguous()
        outs.append(dev.to_host(be.shard_rows(recv, log_n, log_g, h, omega, fid)))
    del cols
    want = hodor.Polynomial.from_coeffs(fid, a).fft(hodor.Worker()).as_ref()
    assert np.array_equal(gather_output(outs), want)


@pytest.mark.parametrize("world", [1, 2, 4, 8])
def test_sharded_lde_fri_building_blocks(hodor, oracle, world):
    """hodor_cuda_lde_cosets_dev / hodor_cuda_fri_fold_shard_dev for every rank of a `world`-GPU box,
    run in turn on one GPU: rank r's coset subset is the cyclic slice LDE[r::G] and its fold is the
    cyclic slice of the next layer.  (The collectives are covered by tests/test_sharded_cpu.py and by
    tools/sharded_check.py under torchrun.)"""
    from hodor_b200 import device as dev
    from hodor_b200.sharded_fri import CudaFriBackend
    fid, log_n, log_f = 0, 13, 3
    L = 1 << log_f
    log_g = world.bit_length() - 1
    be = CudaFriBackend()
    coeffs = oracle.random_elements(fid, 1 << log_n, seed=808)
    full = oracle.lde(fid, coeffs, log_n, L, True)
    proto = oracle.fri_commit(fid, full, L, 1)
    d_coeffs = dev.to_device(coeffs)
    for r in range(world):
        local = be.lde_cosets(d_coeffs, log_n, log_f, True, r, world, log_f - log_g, fid)
        assert np.array_equal(dev.to_host(local), full[r::world])
        nxt = be.fold_shard(local, full.shape[0], 0, log_g, r, proto.challenges[0], fid)
        assert np.array_equal(dev.to_host(nxt), proto.layer_values[0][r::world])
        nxt2 = be.fold_shard(nxt, full.shape[0], 1, log_g, r, proto.challenges[1], fid)
        assert np.array_equal(dev.to_host(nxt2), proto.layer_values[1][r::world])


def test_two_level_fallback_for_large_transforms(oracle):
    """Transforms >= 2^20 normally stream expanded (one entry per element) twiddle / coset tables; with
    HODOR_TABLE_BUDGET_MB=0 the same sizes must take the two-level (hi * lo) path and agree bit for bit."""
    import os
    import subprocess
    import sys
    from conftest import ROOT
    code = (
        "import numpy as np, hodor_b200 as H\n"
        "from oracle import oracle as O\n"
        "H.init(0)\n"
        "a = O.random_elements(0, 1 << 20, seed=321)\n"
        "got = H.Polynomial.from_coeffs(0, a).coset_lde(None, 2).as_ref()\n"
        "assert np.array_equal(got, O.lde(0, a, 20, 2, True)), 'lde'\n"
        "assert np.array_equal(H.Polynomial.from_coeffs(0, a).fft().as_ref(), O.fft(0, a, 20)), 'fft'\n"
        "from hodor_b200 import _ffi\n"
        "print('tables', _ffi.lib.hodor_cuda_workspace_bytes())\n"
    )
    for budget in ("0", "4096"):
        env = dict(os.environ, HODOR_TABLE_BUDGET_MB=budget, PYTHONPATH=ROOT)
        r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=600)
        assert r.returncode == 0, r.stdout + r.stderr


def test_lde_cosets_argument_errors(hodor, oracle):
    from hodor_b200 import _ffi, device as dev
    d = dev.to_device(oracle.random_elements(0, 16, 1))
    out = dev.empty_elems(64)
    lib = _ffi.lib
    assert lib.hodor_cuda_lde_cosets_dev(d.data_ptr(), 4, 2, 1, 3, 2, 1, out.data_ptr(), 0, None) == _ffi.ERR_INVALID_ARG  # 3 + 2 >= 4
    assert lib.hodor_cuda_lde_cosets_dev(d.data_ptr(), 4, 2, 1, 0, 1, 3, out.data_ptr(), 0, None) == _ffi.ERR_INVALID_ARG  # count > L
    assert lib.hodor_cuda_lde_cosets_dev(d.data_ptr(), 4, 30, 1, 0, 1, 0, out.data_ptr(), 0, None) == _ffi.ERR_DOMAIN


def test_sharded_fri_world1_equals_single_gpu_chain(hodor, oracle):
    from hodor_b200 import device as dev
    from hodor_b200.sharded_fri import fri_commit_sharded, lde_sharded
    fid, log_n, log_f = 0, 14, 3
    coeffs = oracle.random_elements(fid, 1 << log_n, seed=909)
    local = lde_sharded(dev.to_device(coeffs), log_n, log_f, True, fid)
    full = oracle.lde(fid, coeffs, log_n, 1 << log_f, True)
    assert np.array_equal(dev.to_host(local), full)
    proto = fri_commit_sharded(local, full.shape[0], 1 << log_f, 1, fid, gather_below=1 << 10)
    want = oracle.fri_commit(fid, full, 1 << log_f, 1)
    assert proto.roots == want.roots() and len(proto.commitments) >= 5
    assert np.array_equal(np.stack(proto.challenges), want.challenges)
    assert np.array_equal(proto.final_coefficients, want.final_coefficients)


# ----------------------------------------------------------------------------------------------
# committed oracles (values + tree resident in HBM), concurrency, pageable memory
# ----------------------------------------------------------------------------------------------
@pytest.mark.parametrize("fid", FIELDS, ids=FIELD_IDS)
def test_committed_oracle_matches_oracle(hodor, oracle, fid):
    """hodor_cuda_tree_commit / _query against Blake2sIopTree::create + get_path of the oracle."""
    for log_n in (1, 2, 5, 10, 11, 14):
        n = 1 << log_n
        vals = oracle.random_elements(fid, n, seed=300 + log_n)
        nodes = oracle.merkle_create(fid, vals)
        orc = hodor.CommittedOracle.create(fid, vals)
        assert orc.size() == n and orc.get_root() == nodes[1].tobytes()
        assert np.array_equal(orc.nodes, nodes) and np.array_equal(orc.values(), vals)
        assert np.array_equal(orc.get_challenge_scalar_from_root(), oracle.interpret_hash(fid, nodes[1].tobytes()))
        idx = sorted({0, n - 1, n // 2, *[int(x) for x in np.random.default_rng(log_n).integers(0, n, 70)]})
        qs = orc.query_batch(idx)  # more than one launch's worth of slots when n is large
        for i, q in zip(idx, qs):
            assert q.natural_index() == i and np.array_equal(q.value(), vals[i])
            assert q.path() == oracle.merkle_path(fid, nodes, vals, i)
            assert hodor.TrivialBlake2sIOP.verify_query(q, orc.get_root())
        assert orc.query(idx[-1]) == qs[-1]
        assert orc == hodor.TrivialBlake2sIOP.create(fid, vals)
        orc.free()
        with pytest.raises(RuntimeError):
            orc.query(0)
    with pytest.raises(AssertionError):
        hodor.CommittedOracle.create(fid, oracle.random_elements(fid, 48, 1))


def test_lde_commit_batch_matches_separate_calls(hodor, oracle):
    """The register loop of Prover::prove (src/prover/mod.rs:73-80): lde then I::create, pipelined."""
    import torch
    from hodor_b200 import device as dev
    fid, log_n, L = 0, 13, 8
    polys = [oracle.random_elements(fid, 1 << log_n, seed=80 + i) for i in range(5)]
    for coset in (False, True):
        for count in (0, 1, 2, 5):
            got = hodor.CommittedOracle.lde_commit_batch([hodor.Polynomial.from_coeffs(fid, a) for a in polys[:count]], L, coset)
            assert len(got) == count
            for a, o in zip(polys, got):
                lde = oracle.lde(fid, a, log_n, L, coset)
                assert np.array_equal(o.values(), lde)
                assert np.array_equal(o.nodes, oracle.merkle_create(fid, lde))
                o.free()
    # FRI straight off a committed oracle's device-resident values (no host copy of the LDE)
    a = oracle.random_elements(fid, 1 << 12, seed=91)
    orc = hodor.CommittedOracle.lde_commit(hodor.Polynomial.from_coeffs(fid, a), L, True)
    lde = oracle.lde(fid, a, 12, L, True)
    from hodor_b200 import _ffi
    h = _ffi.lib.hodor_cuda_fri_commit(orc.device_values_ptr(), C.c_uint64(orc.size()), L, 1, 1, fid)
    assert h, _ffi.last_error()
    proto = hodor.FRIProofPrototype(fid, h, orc.size(), L, 1)
    want = oracle.fri_commit(fid, lde, L, 1)
    assert proto.get_roots() == want.roots() and proto.get_roots()[0] == orc.get_root()
    assert np.array_equal(proto.final_coefficients, want.final_coefficients)
    proto.free()
    orc.free()
    # device-resident coefficients in, and IOP::create on a borrowed device vector
    d = dev.to_device(a)
    out = (C.c_void_p * 1)()
    roots = np.zeros(32, np.uint8)
    ins = (C.c_void_p * 1)(d.data_ptr())
    torch.cuda.synchronize()
    _ffi.check(_ffi.lib.hodor_cuda_lde_commit_batch(ins, 1, 12, 3, 1, 1, out, roots.ctypes.data_as(_ffi.u8p), fid))
    assert roots.tobytes() == want.roots()[0]
    _ffi.lib.hodor_cuda_tree_free(out[0])
    o2 = hodor.CommittedOracle.create_on_device(fid, dev.to_device(lde))
    assert o2.get_root() == want.roots()[0]
    o2.free()
    with pytest.raises(hodor.SynthesisError):  # Domain::new_for_size -> Err past the 2-adicity (BN254: S = 28)
        hodor.CommittedOracle.lde_commit(hodor.Polynomial.from_coeffs(1, oracle.random_elements(1, 1 << 12, 3)), 1 << 17)


def test_c_abi_from_four_host_threads(hodor, oracle):
    """The reference calls its transforms from inside spawned threads (src/polynomials/mod.rs:446-459) and
    `create` makes its own Worker (src/iop/blake2s_trivial_iop.rs:147).  Four host threads hammer the host-pointer
    ABI (transform, Merkle build, FRI commit, committed oracle) concurrently; every result is bit-exact and the
    calls do not deadlock."""
    import threading
    fid = 0
    a = [oracle.random_elements(fid, 1 << 14, seed=500 + t) for t in range(4)]
    want_lde = [oracle.lde(fid, x, 14, 4, True) for x in a]
    want_nodes = [oracle.merkle_create(fid, x) for x in a]
    want_fri = [oracle.fri_commit(fid, x, 4, 1) for x in a]
    errors = []

    def work(t):
        try:
            for rep in range(6):
                got = hodor.Polynomial.from_coeffs(fid, a[t]).coset_lde(hodor.Worker(), 4).as_ref()
                assert np.array_equal(got, want_lde[t]), ("lde", t, rep)
                tree = hodor.Blake2sIopTree.create(fid, a[t])
                assert np.array_equal(tree.nodes, want_nodes[t]), ("merkle", t, rep)
                proto = hodor.NaiveFriIop.proof_from_lde(hodor.Polynomial.from_values(fid, a[t]), 4, 1, None)
                assert proto.get_roots() == want_fri[t].roots(), ("fri", t, rep)
                proto.free()
                orc = hodor.CommittedOracle.lde_commit(hodor.Polynomial.from_coeffs(fid, a[t]), 4, True)
                assert orc.get_root() == oracle.merkle_create(fid, want_lde[t])[1].tobytes(), ("commit", t, rep)
                orc.free()
        except BaseException as e:  # noqa: BLE001
            errors.append(e)

    threads = [threading.Thread(target=work, args=(t,)) for t in range(4)]
    for th in threads:
        th.start()
    for th in threads:
        th.join(timeout=300)
        assert not th.is_alive(), "deadlock: a thread is still inside the C ABI"
    assert not errors, errors


def test_dev_calls_on_two_streams_share_the_workspace_safely(hodor, oracle):
    """`_dev` entry points enqueue on the caller's stream and share one context workspace; calls issued
    back to back on two different streams must not overwrite each other's inter-pass data."""
    import torch
    from hodor_b200 import device as dev
    fid, log_n = 0, 16
    xs = [oracle.random_elements(fid, 1 << log_n, seed=600 + i) for i in range(4)]
    omega = oracle.domain_generator(fid, log_n)
    want = [oracle.best_fft(fid, x, omega, log_n) for x in xs]
    d_in = [dev.to_device(x) for x in xs]
    d_out = [dev.empty_elems(1 << log_n) for _ in xs]
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    torch.cuda.synchronize()
    for rep in range(3):
        for i in range(4):
            with torch.cuda.stream(streams[i % 2]):
                dev.fft(d_in[i], d_out[i], log_n, False, fid)
        torch.cuda.synchronize()
        for i in range(4):
            assert np.array_equal(dev.to_host(d_out[i]), want[i]), (rep, i)


def test_pageable_host_memory_is_accepted(hodor, oracle):
    """A Rust `Vec<F>` is pageable: the host-pointer entry points must give the same bits from plain numpy
    buffers (no pinning) as from pinned ones."""
    from hodor_b200 import _ffi
    from hodor_b200.field import _p
    fid, log_n, log_f = 0, 16, 2
    a = oracle.random_elements(fid, 1 << log_n, seed=77)
    out = np.zeros(((1 << log_n) << log_f, 4), np.uint64)
    _ffi.check(_ffi.lib.hodor_cuda_lde(_p(a), log_n, log_f, 1, _p(out), fid))
    assert np.array_equal(out, oracle.lde(fid, a, log_n, 1 << log_f, True))


# ----------------------------------------------------------------------------------------------
# multi-GPU entry points of the C ABI
# ----------------------------------------------------------------------------------------------
def test_merkle_build_shard_reindexes_cyclic_chunks(hodor, oracle):
    """The leaf kernel of the sharded chain reads a layer that arrived as G cyclic-slice chunks; the tree must be
    the tree of the natural-order vector (no transposing copy in between)."""
    import torch
    from hodor_b200 import _ffi
    from hodor_b200 import device as dev
    fid = 0
    for log_g, log_n in ((1, 13), (2, 14), (3, 14), (4, 15)):
        n, G = 1 << log_n, 1 << log_g
        v = oracle.random_elements(fid, n, seed=40 + log_g)
        chunks = np.concatenate([v[r::G] for r in range(G)])  # chunk r = v[r + G t]
        d = dev.to_device(chunks)
        nodes = dev.empty_elems(n)
        root = torch.zeros(32, dtype=torch.uint8, device="cuda")
        chal = dev.empty_elems(1)
        _ffi.check(_ffi.lib.hodor_cuda_merkle_build_shard_dev(d.data_ptr(), C.c_uint64(n), log_g, nodes.data_ptr(), root.data_ptr(),
                                                              chal.data_ptr(), fid, torch.cuda.current_stream().cuda_stream))
        want = oracle.merkle_create(fid, v)
        assert np.array_equal(nodes.cpu().numpy().view(np.uint8).reshape(n, 32), want)
        assert root.cpu().numpy().tobytes() == want[1].tobytes()
    assert _ffi.lib.hodor_cuda_merkle_build_shard_dev(d.data_ptr(), C.c_uint64(2048), 2, nodes.data_ptr(), None, None, fid, None) \
        == _ffi.ERR_INVALID_ARG


def test_sharded_c_abi_world1(hodor, oracle):
    """hodor_cuda_comm_init(0, 1) needs no NCCL; the sharded entry points then equal the single-GPU ones."""
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    import sharded_check as sc
    from hodor_b200 import multigpu as mg
    assert mg.comm_init() == (0, 1)
    for r in (sc.check_ntt(18, reps=1), sc.check_ntt(12, reps=1), sc.check_lde_fri(16, 3, reps=1), sc.check_lde_fri(12, 4, reps=1)):
        assert r["ok"], r
    roots, chals, fin = mg.lde_fri_sharded(__import__("hodor_b200").device.to_device(oracle.random_elements(0, 1 << 12, 5)), 12, 3, True, 2, 0)
    want = oracle.fri_commit(0, oracle.lde(0, oracle.random_elements(0, 1 << 12, 5), 12, 8, True), 8, 2)
    assert roots == want.roots() and np.array_equal(chals, want.challenges) and np.array_equal(fin, want.final_coefficients)
    assert mg.bytes_sent() == 0


def test_sharded_c_abi_two_gpus_under_torchrun():
    """The NCCL composition itself: tools/sharded_check.py under torchrun on 2 GPUs compares the library's
    four-step NTT and its sharded LDE + FRI chain with the single-GPU results.  Skipped on a one-GPU box."""
    import json
    import os
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29517", os.path.join(root, "tools", "sharded_check.py"), "18", "3", "20"]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=root)
    lines = [json.loads(l) for l in p.stdout.splitlines() if l.startswith("{")]
    assert p.returncode == 0 and len(lines) == 4 and all(l["ok"] for l in lines), (p.returncode, p.stdout[-2000:], p.stderr[-2000:])


def test_sharded_ntt_receive_buffer_regrows_under_torchrun():
    """Sizes in an order that makes the receive buffer of the fused exchange grow twice and then be reused for
    smaller transforms: every growth retires the old buffer until all peers have dropped their CUDA IPC mapping
    of it and re-maps the new one (sharded.cu).  2 GPUs; skipped on a one-GPU box."""
    import json
    import os
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29519", os.path.join(root, "tools", "sharded_check.py")]
    env = dict(os.environ, HODOR_CHECK_NTT_SIZES="14,18,22,20,12,22")
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=root, env=env)
    lines = [json.loads(l) for l in p.stdout.splitlines() if l.startswith("{")]
    assert p.returncode == 0 and len(lines) == 6 and all(l["ok"] for l in lines), (p.returncode, p.stdout[-2000:], p.stderr[-2000:])


@pytest.mark.parametrize("level", ["0", "1", "3"])
def test_fused_last_pass_commit_matches_oracle(level):
    """HODOR_FUSE_LAST_COMMIT (read at init, hence a process of its own): lift-and-commit with the bottom three tree
    levels hashed inside the last pass of the transform (csrc/ntt_commit.cuh) for no plan (0), for plans ending in an
    8-bit digit (1, the default) and for every plan (3) gives the oracle's values and nodes over every last-pass width, blowup
    1..16, three fields -- and each case ran the kernel its plan and the level call for."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, HODOR_FUSE_LAST_COMMIT=level)
    p = subprocess.run([sys.executable, os.path.join(root, "tools", "fused_commit_check.py")], capture_output=True, text=True,
                       timeout=900, cwd=root, env=env)
    lines = [json.loads(l) for l in p.stdout.splitlines() if l.startswith("{")]
    assert p.returncode == 0 and lines and lines[-1]["ok"] and lines[-1]["level"] == int(level), (p.returncode, p.stdout[-3000:], p.stderr[-2000:])
    assert all(l["ok"] for l in lines[:-1])
    fused = [l["fused"] for l in lines[:-1]]
    assert (not any(fused)) if level == "0" else (all(fused) if level == "3" else (any(fused) and not all(fused)))


@pytest.mark.parametrize("fused", ["0", "1"])
def test_fri_chain_fold_commit_switch_matches_oracle(fused):
    """HODOR_FUSE_FOLD_COMMIT (read at init, hence a process of its own): the FRI commit chain with the fold fused with
    the bottom three levels of the next tree (csrc/fri.cuh fri_fold_commit_kernel) and with separate kernels gives
    the oracle's roots, challenges, final coefficients, layer values and nodes -- and the kernel the switch names ran."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, HODOR_FUSE_FOLD_COMMIT=fused)
    p = subprocess.run([sys.executable, os.path.join(root, "tools", "fri_fused_check.py")], capture_output=True, text=True,
                       timeout=900, cwd=root, env=env)
    lines = [json.loads(l) for l in p.stdout.splitlines() if l.startswith("{")]
    assert p.returncode == 0 and lines and lines[-1]["ok"] and lines[-1]["want_fused"] == (fused == "1"), (p.returncode, p.stdout[-3000:], p.stderr[-2000:])
    assert all(l["ok"] and l["fused"] == (fused == "1") for l in lines[:-1])


def test_misaligned_device_pointer_is_rejected(hodor, oracle):
    """Element arrays are read with 256-bit loads: a device pointer that is not 32-byte aligned is an argument
    error, not a misaligned-address fault that would kill the context."""
    import torch
    from hodor_b200 import device as dev
    fid, log_n = 0, 12
    n = 1 << log_n
    flat = torch.zeros(4 * n + 8, dtype=torch.int64, device="cuda")
    skew = flat[2 : 2 + 4 * n].view(n, 4)  # 16 bytes past a 32-byte boundary
    assert skew.data_ptr() % 32 == 16
    good = dev.empty_elems(n)
    omega = hodor.Domain.new_for_size(fid, n).generator
    with pytest.raises(hodor.HodorError):
        dev.ntt(skew, good, log_n, omega, fid)
    with pytest.raises(hodor.HodorError):
        dev.ntt(good, skew, log_n, omega, fid)
    with pytest.raises(hodor.HodorError):
        dev.merkle_build(skew, n, good, fid)
    # the context is still healthy
    a = oracle.random_elements(fid, n, seed=5)
    assert np.array_equal(hodor.Polynomial.from_coeffs(fid, a).fft(hodor.Worker()).as_ref(), oracle.serial_fft(fid, a, omega, log_n))
