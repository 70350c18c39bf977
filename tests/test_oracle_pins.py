"""Pins the oracle (C restatement + big-int model) against everything the reference holds for this
path.  The reference has no golden vectors -- "parity unpinned" -- so what can be
pinned is: the Montgomery constants hard-coded in src/experiments/square_root_calculator/fp2.rs,
RFC 7693 Blake2s (via hashlib, an independent implementation), and the known answers SURVEY.md 8c
derived for the reference's own test shapes (make_small_tree, test_one_fri_step)."""
import hashlib

import numpy as np
import pytest

from conftest import FIELD_IDS, FIELDS

MODELS = {0: "BLS12_381_FR", 1: "BN254_FR", 2: "STARK252"}


def test_reference_montgomery_constants(oracle, pymodel):
    """fp2.rs:10-22: NON_RESIDUE = Fq(FqRepr([..ffa1, ..ffff, ..ffff, 0x07fffffffffff9b0])) is 3 and
    MINUS_ONE = Fq(FqRepr([0x20, 0, 0, 0x220])) is p-1, both in Montgomery form with R = 2^256."""
    F = pymodel.STARK252
    non_residue = [0xFFFFFFFFFFFFFFA1, 0xFFFFFFFFFFFFFFFF, 0xFFFFFFFFFFFFFFFF, 0x07FFFFFFFFFFF9B0]
    minus_one = [0x20, 0, 0, 0x220]
    three = oracle.to_mont(2, oracle.ints_to_array([3]))[0]
    m1 = oracle.to_mont(2, oracle.ints_to_array([F.p - 1]))[0]
    assert [int(x) for x in three] == non_residue
    assert [int(x) for x in m1] == minus_one
    assert F.to_mont(3) == oracle.limbs_to_int(np.array(non_residue, np.uint64))
    assert F.to_mont(F.p - 1) == oracle.limbs_to_int(np.array(minus_one, np.uint64))


def test_bn256_rs_is_bls12_381_scalar_field(oracle, pymodel):
    """src/bn256.rs:5 modulus, :6 generator 7; 2-adicity 32."""
    c = oracle.field_constants(0)
    assert oracle.limbs_to_int(c["p"]) == 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001
    assert c["s"] == 32 and c["num_bits"] == 255
    assert oracle.limbs_to_int(oracle.from_mont(0, c["generator"])[0]) == 7


def test_published_constants_of_independent_implementations(oracle, pymodel):
    """ff_ce's derive is absent from /root/reference (Cargo.toml:16, no Cargo.lock); what it computes is
    published: ROOT_OF_UNITY = GENERATOR^((p-1)/2^S), stored in Montgomery form with R = 2^256.  Two
    independent, widely deployed implementations of the same two fields hard-code the results, and they are
    reproduced here bit for bit (constants quoted from those crates' sources):
      * zkcrypto `bls12_381::Scalar` (src/scalar.rs): GENERATOR = 7, S = 32, INV, R, R2, ROOT_OF_UNITY as
        Montgomery limbs -- the field src/bn256.rs declares with the same generator 7;
      * halo2curves `bn256::Fr`: GENERATOR = 7, S = 28, ROOT_OF_UNITY (plain integer)."""
    def limbs(x):
        return sum(int(l) << (64 * i) for i, l in enumerate(x))

    c = oracle.field_constants(0)
    assert c["inv"] == 0xFFFFFFFEFFFFFFFF
    assert limbs(c["r"]) == limbs([0x00000001FFFFFFFE, 0x5884B7FA00034802, 0x998C4FEFECBC4FF5, 0x1824B159ACC5056F])
    assert limbs(c["r2"]) == limbs([0xC999E990F3F29C6D, 0x2B6CEDCB87925C23, 0x05D314967254398F, 0x0748D9D99F59FF11])
    assert limbs(c["generator"]) == limbs([0x0000000EFFFFFFF1, 0x17E363D300189C0F, 0xFF9C57876F8457B0, 0x351332208FC5A8C4])
    assert limbs(c["root_of_unity"]) == limbs([0xB9B58D8C5F0E466A, 0x5B1B4C801819D7EC, 0x0AF53AE352A31E64, 0x5BF3ADDA19E9B27B])
    assert pymodel.BLS12_381_FR.to_mont(pymodel.BLS12_381_FR.root_of_unity) == limbs(c["root_of_unity"])
    c = oracle.field_constants(1)
    assert oracle.limbs_to_int(oracle.from_mont(1, c["generator"])[0]) == 7
    assert oracle.limbs_to_int(oracle.from_mont(1, c["root_of_unity"])[0]) == \
        0x03DDB9F5166D18B798865EA93DD31F743215CF6DD39329C8D34F1ED960C37C9C
    assert pymodel.BN254_FR.root_of_unity == 0x03DDB9F5166D18B798865EA93DD31F743215CF6DD39329C8D34F1ED960C37C9C


@pytest.mark.parametrize("fid", FIELDS, ids=FIELD_IDS)
def test_field_constants_agree_with_bigint_model(oracle, pymodel, fid):
    F = getattr(pymodel, MODELS[fid])
    c = oracle.field_constants(fid)
    assert oracle.limbs_to_int(c["p"]) == F.p
    assert oracle.limbs_to_int(c["r"]) == F.R
    assert oracle.limbs_to_int(c["r2"]) == F.R * F.R % F.p
    assert c["inv"] == (-pow(F.p, -1, 2**64)) % 2**64
    assert c["s"] == F.s and c["num_bits"] == F.num_bits
    assert oracle.limbs_to_int(c["root_of_unity"]) == F.to_mont(F.root_of_unity)
    # root_of_unity has order exactly 2^S
    assert pow(F.root_of_unity, 1 << (F.s - 1), F.p) == F.p - 1


def test_blake2s_against_rfc7693_implementation(oracle):
    """blake2s_simd Params{hash_length 32, key, personal} == RFC 7693; hashlib is independent of
    both oracle implementations."""
    kat = hashlib.blake2s(b"", key=b"Squeamish Ossifrage", person=b"Shaftoe", digest_size=32).hexdigest()
    assert kat == "a61dd261a9b23522c19ebdecc9b5755882c1b4f3940d3437029d99120ab1b437"
    assert oracle.blake2s(b"").hex() == kat
    rng = np.random.default_rng(7)
    for ln in (1, 31, 32, 33, 63, 64, 65, 127, 128, 129, 200):
        data = rng.integers(0, 256, ln, dtype=np.uint8).tobytes()
        want = hashlib.blake2s(data, key=b"Squeamish Ossifrage", person=b"Shaftoe", digest_size=32).digest()
        assert oracle.blake2s(data) == want


SURVEY_KAT = {
    # SURVEY.md 8c "Known answers": encode_leaf(one), hash_leaf(one), root of 16 x one, challenge
    0: ("feffffff0100000002480300fab78458f54fbcecef4f8c996f05c5ac59b12418",
        "66af316f9b1a181e1006da977f609f37c346a7d4d40c43e1abf47ad3ef5f9505",
        "661512723ab4cfa09bdd1aad0e9f1cc69356055f99a9528b016f35b8c5fe706b",
        0x261512723AB4CFA09BDD1AAD0E9F1CC69356055F99A9528B016F35B8C5FE706B),
    2: ("e1fffffffffffffffffffffffffffffffffffffffffffffff0fdffffffffff07",
        "d24624c02e2d6f62358acd21592e93264ec96f63ce714e872c7bcde465214360",
        "fdf489862b4402468d94f026c014e1ca0129f421a55ce5e1df838eab8eefbd22",
        0x05F489862B4402468D94F026C014E1CA0129F421A55CE5E1DF838EAB8EEFBD22),
}


@pytest.mark.parametrize("fid", [0, 2], ids=["bls12_381_fr", "stark252"])
def test_make_small_tree_shape(oracle, fid):
    """make_small_tree (src/iop/blake2s_trivial_iop.rs:377-387): 16 x Fr::one()."""
    enc, leaf_hash, root, challenge = SURVEY_KAT[fid]
    one = oracle.field_constants(fid)["r"]
    assert one.tobytes().hex() == enc
    assert oracle.hash_leaf(fid, one).hex() == leaf_hash
    nodes = oracle.merkle_create(fid, np.tile(one, (16, 1)))
    assert nodes[1].tobytes().hex() == root
    assert not nodes[0].any()
    c = oracle.interpret_hash(fid, nodes[1].tobytes())
    assert oracle.limbs_to_int(oracle.from_mont(fid, c)[0]) == challenge


def test_ntt4_known_answer(oracle):
    """NTT_4([1,2,3,4]) over bn256.rs's field."""
    p = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001
    a = oracle.to_mont(0, oracle.ints_to_array([1, 2, 3, 4]))
    w4 = oracle.domain_generator(0, 2)
    assert oracle.limbs_to_int(oracle.from_mont(0, w4)[0]) == 0x8D51CCCE760304D0EC030002760300000001000000000000
    out = oracle.array_to_ints(oracle.from_mont(0, oracle.serial_fft(0, a, w4, 2)))
    assert out == [0xA, 0x73EDA753299D7D4718963E6B1D9BCE637BB7A3FE13F85BFEFFFDFFFEFFFFFFFF, p - 2,
                   0x11AA3999CEC0609A1D8060004EC0600000001FFFFFFFFFFFE]


def test_one_fri_step_f257(pymodel):
    """test_one_fri_step (src/fri/mod.rs:252-331) over the reference's toy field F_257:
    coeffs [1,2,4,8], lde_factor 4, output 2."""
    F = pymodel.F257
    lde = pymodel.lde(F, [1, 2, 4, 8], 2, 4, coset=False)
    assert lde == [15, 0, 97, 85, 93, 0, 71, 174, 252, 0, 34, 206, 158, 4, 59, 53]
    by_values = pymodel.fri_commit(F, lde, 4, 2)
    by_coeffs = pymodel.fri_commit_through_coefficients(F, lde, 4, 2)
    assert by_values.l0_nodes[1].hex() == "23f8af21441f78018052ace5f31145c9ed7ebd7d6c49f1ced9d29ee296e58ec9"
    assert by_values.challenges == [1]
    assert by_values.layer_values == [[15, 0, 68, 51, 248, 6, 195, 212]]
    c = by_values.challenges[0]
    assert by_values.final_coefficients == [(1 + c * 2) % 257, (4 + c * 8) % 257]
    # values-FRI == coefficients-FRI on every field of the prototype (:312-317)
    assert by_values.final_coefficients == by_coeffs.final_coefficients
    assert by_values.final_root == by_coeffs.final_root
    assert by_values.layer_values == by_coeffs.layer_values
    assert by_values.challenges == by_coeffs.challenges
    assert by_values.l0_nodes == by_coeffs.l0_nodes and by_values.layer_nodes == by_coeffs.layer_nodes
