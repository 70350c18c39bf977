"""GPU parity at BASELINE.json's full sizes, bit-exact against the CPU oracle on the SAME unsparsified
inputs: configs[1] (coset LDE 2^24 x 8, all 2^27 outputs), the plain NTT at 2^22 / 2^23 / 2^24 (the
metric's size, plan 8+8+8), configs[2] (FRI commit chain on 2^24 values, blowup 8 and 16: every root,
challenge, final coefficient, every layer's values and every tree's nodes).  The oracle needs 10-30 s
per case on the box's host cores; nothing is sampled."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_coset_lde_2p24_x8_bit_exact(hodor, oracle):
    fid, log_n, L = 0, 24, 8
    a = oracle.random_elements(fid, 1 << log_n, seed=2401)
    got = hodor.Polynomial.from_coeffs(fid, a).coset_lde(hodor.Worker(), L).as_ref()
    want = oracle.lde(fid, a, log_n, L, True)
    assert got.shape == want.shape == ((1 << log_n) * L, 4)
    assert np.array_equal(got, want)
    # the committed form of the same lift (hodor_cuda_lde_commit): identical values, tree == oracle's create
    orc = hodor.CommittedOracle.lde_commit(hodor.Polynomial.from_coeffs(fid, a), L, coset=True)
    for first in (0, 1 << 20, (1 << 27) - (1 << 16)):
        assert np.array_equal(orc.values(first, 1 << 16), want[first:first + (1 << 16)])
    del got
    nodes = oracle.merkle_create(fid, want)
    assert orc.get_root() == nodes[1].tobytes()
    assert np.array_equal(orc.nodes, nodes)
    orc.free()


@pytest.mark.parametrize("log_n", [22, 23, 24])
def test_ntt_full_size_bit_exact(hodor, oracle, log_n):
    fid = 0
    a = oracle.random_elements(fid, 1 << log_n, seed=2200 + log_n)
    omega = oracle.domain_generator(fid, log_n)
    got = hodor.Polynomial.from_coeffs(fid, a).fft(hodor.Worker()).as_ref()
    assert np.array_equal(got, oracle.best_fft(fid, a, omega, log_n))


@pytest.mark.parametrize("L", [8, 16])
def test_fri_chain_2p24_bit_exact(hodor, oracle, L):
    fid, log_n, oc = 0, 24, 1
    v = oracle.random_elements(fid, 1 << log_n, seed=7700 + L)
    proto = hodor.NaiveFriIop.proof_from_lde(hodor.Polynomial.from_values(fid, v), L, oc, hodor.Worker())
    want = oracle.fri_commit(fid, v, L, oc)
    assert proto.num_steps == 24 - (L.bit_length() - 1)
    assert proto.get_roots() == want.roots()
    assert np.array_equal(proto.challenges, want.challenges)
    assert proto.get_final_root() == want.final_root
    assert np.array_equal(proto.final_coefficients, want.final_coefficients)
    assert np.array_equal(proto.l0_commitment.nodes, want.l0_nodes)
    for i in range(proto.num_steps):
        nodes, values = proto._fetch_layer(i + 1, want_nodes=True, want_values=True)
        assert np.array_equal(values, want.layer_values[i]), f"layer {i} values"
        assert np.array_equal(nodes, want.layer_nodes[i]), f"layer {i} nodes"
    proto.free()
