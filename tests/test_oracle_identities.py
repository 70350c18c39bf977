"""The reference's own test strategy restated on the oracle: two code paths
must agree, or a round trip must return the input; plus the C oracle against the independent
big-integer model.  CPU only."""
import numpy as np
import pytest

from conftest import FIELD_IDS, FIELDS

MODELS = {0: "BLS12_381_FR", 1: "BN254_FR", 2: "STARK252"}


def plain(F, arr, oracle):
    return [F.from_mont(x) for x in oracle.array_to_ints(arr)]


@pytest.mark.parametrize("fid", FIELDS, ids=FIELD_IDS)
def test_c_oracle_matches_bigint_model(oracle, pymodel, fid):
    F = getattr(pymodel, MODELS[fid])
    a = oracle.random_elements(fid, 64)
    assert oracle.array_to_ints(a) == pymodel.random_mont_elements(F, 64)
    ai = plain(F, a, oracle)
    for ln in (0, 1, 2, 3, 6):
        n = 1 << ln
        om, omp = oracle.domain_generator(fid, ln), F.domain_generator(ln)
        r = oracle.serial_fft(fid, a[:n], om, ln)
        assert plain(F, r, oracle) == pymodel.serial_fft(F, ai[:n], omp, ln) == pymodel.dft(F, ai[:n], omp)
    for L in (1, 2, 8):
        for coset in (False, True):
            r = oracle.lde(fid, a[:8], 3, L, coset)
            assert plain(F, r, oracle) == pymodel.lde(F, ai[:8], 3, L, coset)
            # every output is the polynomial evaluated at shift * w_{nL}^idx (Horner, definition)
            w = F.domain_generator(3 + L.bit_length() - 1)
            shift = F.generator if coset else 1
            assert plain(F, r, oracle) == [pymodel.evaluate(F, ai[:8], shift * pow(w, i, F.p) % F.p) for i in range(8 * L)]
    nodes = oracle.merkle_create(fid, a[:32])
    assert [x.tobytes() for x in nodes] == pymodel.merkle_create(F, ai[:32])
    pr, pm = oracle.fri_commit(fid, a[:64], 4, 2), pymodel.fri_commit(F, ai[:64], 4, 2)
    assert plain(F, pr.challenges, oracle) == pm.challenges
    assert pr.final_root == pm.final_root
    assert plain(F, pr.final_coefficients, oracle) == pm.final_coefficients
    for i in range(len(pm.layer_values)):
        assert plain(F, pr.layer_values[i], oracle) == pm.layer_values[i]
        assert [x.tobytes() for x in pr.layer_nodes[i]] == pm.layer_nodes[i]


def test_config1_roundtrip_2p10(oracle):
    """BASELINE.json configs[0]: 2^10 forward + inverse NTT round trip over bn256.rs's field on the
    CPU reference path (serial_fft, src/fft/fft.rs:21-66), bit-exact."""
    fid, ln = 0, 10
    a = oracle.random_elements(fid, 1 << ln)
    om = oracle.domain_generator(fid, ln)
    fwd = oracle.serial_fft(fid, a, om, ln)
    inv = oracle.serial_fft(fid, fwd, oracle.inverse(fid, om), ln)
    ninv = oracle.inverse(fid, oracle.to_mont(fid, oracle.ints_to_array([1 << ln]))[0])
    back = oracle.mul(fid, inv, np.tile(ninv, (1 << ln, 1)))
    assert np.array_equal(back, a)
    assert np.array_equal(oracle.ifft(fid, fwd, ln), a)


@pytest.mark.parametrize("fid", FIELDS, ids=FIELD_IDS)
def test_radix2_radix4_parallel_agree(oracle, fid):
    """test_sequential_radix4_fft / test_parallel_radix4_fft / test_worker_size (src/fft/mod.rs:66,
    128, 281) at sizes the CPU finishes quickly."""
    ln = 12
    a = oracle.random_elements(fid, 1 << ln, seed=5)
    om = oracle.domain_generator(fid, ln)
    r2 = oracle.serial_fft(fid, a, om, ln)
    assert np.array_equal(oracle.serial_fft_radix_4(fid, a, om, ln), r2)
    for cpus in (1, 2, 3, 4, 7, 8, 16):
        assert np.array_equal(oracle.best_fft(fid, a, om, ln, cpus), r2)
        assert np.array_equal(oracle.best_fft_radix_4(fid, a, om, ln, cpus), r2)
    small = oracle.random_elements(fid, 32, seed=9)  # test_worker_size: n = 32
    om5 = oracle.domain_generator(fid, 5)
    ref = oracle.serial_fft(fid, small, om5, 5)
    for cpus in range(1, 17):
        assert np.array_equal(oracle.best_fft(fid, small, om5, 5, cpus), ref)
    assert np.array_equal(oracle.ifft(fid, ref, 5), small)


@pytest.mark.parametrize("fid", FIELDS, ids=FIELD_IDS)
def test_lde_equals_zero_padded_ntt(oracle, fid):
    """test_lde_correctness / test_coset_lde_correctness (src/polynomials/mod.rs:988, 1036): the
    multi-coset LDE equals the (coset) NTT of the zero-padded coefficient vector."""
    ln, L = 6, 16
    a = oracle.random_elements(fid, 1 << ln, seed=11)
    padded = np.zeros(((1 << ln) * L, 4), np.uint64)
    padded[: 1 << ln] = a
    for coset in (False, True):
        assert np.array_equal(oracle.lde(fid, a, ln, L, coset), oracle.fft(fid, padded, ln + 4, coset=coset))
    for coset in (False, True):  # n = 4, L = 16, the reference's small case
        assert np.array_equal(oracle.lde(fid, a[:4], 2, 16, coset),
                              oracle.fft(fid, np.concatenate([a[:4], np.zeros((60, 4), np.uint64)]), 6, coset=coset))


def test_lde_chunk_index_quirk_is_reproduced(oracle):
    """The reference seeds each worker chunk's coset generator with coset_omega^chunk_index
    (src/polynomials/mod.rs:448, 575) rather than ^(chunk_index * chunk).  With num_cpus >= factor
    (chunk == 1) that is the right coset; with fewer CPUs the reference's output is not an LDE.
    The oracle keeps the quirk; the product implements the well-defined chunk == 1 result."""
    fid, ln, L = 0, 4, 8
    a = oracle.random_elements(fid, 1 << ln, seed=3)
    good = oracle.lde(fid, a, ln, L, True, cpus=8)
    assert np.array_equal(good, oracle.lde(fid, a, ln, L, True, cpus=64))
    assert not np.array_equal(good, oracle.lde(fid, a, ln, L, True, cpus=2))


@pytest.mark.parametrize("fid", FIELDS, ids=FIELD_IDS)
def test_merkle_every_path_verifies(oracle, fid):
    """make_small_iop (src/iop/blake2s_trivial_iop.rs:390-408): 64 leaves 1, 2, 4, ..."""
    vals = oracle.to_mont(fid, oracle.ints_to_array([1 << i for i in range(64)]))
    nodes = oracle.merkle_create(fid, vals)
    root = nodes[1].tobytes()
    for i in range(64):
        path = oracle.merkle_path(fid, nodes, vals, i)
        assert len(path) == 6
        assert oracle.merkle_verify(fid, root, vals[i], path, i)
        assert not oracle.merkle_verify(fid, root, vals[i ^ 3], path, i)
    for cpus in (1, 3, 8):
        assert np.array_equal(oracle.merkle_create(fid, vals, cpus=cpus), nodes)


@pytest.mark.parametrize("fid", [0, 2], ids=["bls12_381_fr", "stark252"])
def test_fri_on_values_vs_on_coefficients(oracle, pymodel, fid):
    """test_fri_on_values_vs_on_coefficients (src/fri/mod.rs:510): both provers agree on a genuine
    low-degree input, and the fold of layer i is the LDE of the pairwise-folded coefficients."""
    F = getattr(pymodel, MODELS[fid])
    ln, L, out = 4, 4, 2
    coeffs = oracle.random_elements(fid, 1 << ln, seed=21)
    lde = oracle.lde(fid, coeffs, ln, L, False)
    pv = oracle.fri_commit(fid, lde, L, out)
    pc = pymodel.fri_commit_through_coefficients(F, plain(F, lde, oracle), L, out)
    assert plain(F, pv.final_coefficients, oracle) == pc.final_coefficients
    assert pv.final_root == pc.final_root
    assert plain(F, pv.challenges, oracle) == pc.challenges
    for i in range(len(pc.layer_values)):
        assert plain(F, pv.layer_values[i], oracle) == pc.layer_values[i]
        assert [x.tobytes() for x in pv.layer_nodes[i]] == pc.layer_nodes[i]


@pytest.mark.parametrize("fid", FIELDS, ids=FIELD_IDS)
def test_batch_inversion_and_evaluate_at(oracle, pymodel, fid):
    """test_batch_inversion (src/polynomials/mod.rs:959-985): batch inverse == per-element inverse, for
    every worker count; a zero element gives the reference's Err and leaves the vector alone.
    evaluate_at (:685-711) == Horner on integers, for every worker count."""
    F = getattr(pymodel, MODELS[fid])
    a = oracle.random_elements(fid, 77, seed=5)
    ai = plain(F, a, oracle)
    want = [pow(x, -1, F.p) for x in ai]
    for cpus in (1, 2, 3, 8, 200):
        got = oracle.batch_inversion(fid, a, cpus=cpus)
        assert plain(F, got, oracle) == want
        for n in (0, 1, 5, 77):
            z = a[3]
            assert F.from_mont(oracle.limbs_to_int(oracle.evaluate_at(fid, a[:n], z, cpus=cpus))) == \
                pymodel.evaluate(F, ai[:n], ai[3])
    assert np.array_equal(oracle.batch_inversion(fid, a[:1])[0], oracle.inverse(fid, a[0]))
    bad = a.copy()
    bad[40] = 0
    assert oracle.batch_inversion(fid, bad) is None


@pytest.mark.parametrize("fid", FIELDS, ids=FIELD_IDS)
def test_dif_variant_and_pruning(oracle, fid):
    """test_sequential_radix4_fft / test_fft_prunning (src/fft/mod.rs:66-126, 187-279): the DIF variant
    (src/fft/dit_fft/mod.rs) gives the radix-2 result, and pruning the butterflies of a zero tail does
    not change it."""
    for ln in (1, 2, 6, 10):
        n = 1 << ln
        a = oracle.random_elements(fid, n, seed=40 + ln)
        om = oracle.domain_generator(fid, ln)
        want = oracle.serial_fft(fid, a, om, ln)
        assert np.array_equal(oracle.serial_dif_fft(fid, a, om, ln), want)
        if ln % 2 == 0:
            assert np.array_equal(oracle.serial_fft_radix_4(fid, a, om, ln), want)
        nz = max(1, n // 16)
        padded = a.copy()
        padded[nz:] = 0
        full = oracle.serial_fft(fid, padded, om, ln)
        assert np.array_equal(oracle.serial_dif_fft(fid, padded, om, ln, non_zero_entries=nz), full)
        assert np.array_equal(oracle.serial_dif_fft(fid, padded, om, ln), full)


@pytest.mark.parametrize("fid", FIELDS, ids=FIELD_IDS)
def test_filtering_lde_equals_multi_coset_lde(oracle, fid):
    """test_lde_correctness / test_coset_lde_correctness / test_various_ldes (src/polynomials/mod.rs:988-1135):
    the zero-aware NTT of the zero-padded vector (filtering_lde, src/fft/lde.rs) == the multi-coset LDE ==
    the plain NTT of the zero-padded (and, for the coset form, g^j-scaled) vector."""
    for ln, L in ((2, 16), (0, 4), (3, 1), (5, 2), (6, 8), (8, 16)):
        a = oracle.random_elements(fid, 1 << ln, seed=70 + ln)
        for coset in (False, True):
            multi = oracle.lde(fid, a, ln, L, coset)
            assert np.array_equal(oracle.filtering_lde(fid, a, ln, L, coset), multi)
            padded = np.zeros(((1 << ln) * L, 4), np.uint64)
            padded[: 1 << ln] = oracle.distribute_powers(fid, a, oracle.field_constants(fid)["generator"]) if coset else a
            tl = ln + L.bit_length() - 1
            assert np.array_equal(oracle.serial_fft(fid, padded, oracle.domain_generator(fid, tl), tl), multi)
