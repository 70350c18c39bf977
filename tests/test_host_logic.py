"""CPU-side checks of the product: the C-ABI library loads and exports every symbol the header
declares, its host scalar helpers agree with the oracle, the reference's error behaviour is kept,
the Montgomery multiplier of csrc/field.cuh (host build) is exact, and -- with no GPU in the box --
every compute entry point refuses loudly instead of computing on the CPU."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import FIELD_IDS, FIELDS, ROOT


def header_functions():
    text = open(os.path.join(ROOT, "include", "hodor_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(hodor_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    import hodor_b200._ffi as ffi

    declared = header_functions()
    assert len(declared) >= 40
    for name in declared:
        assert hasattr(ffi.lib, name), f"{name} declared in include/hodor_b200.h but not exported"
    assert sorted(ffi.EXPORTED_SYMBOLS) == declared, "ctypes signature table and header disagree"


def test_library_has_sm100a_code_only():
    out = subprocess.run(["cuobjdump", "-lelf", os.path.join(ROOT, "hodor_b200", "libhodor_b200.so")],
                         capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


@pytest.mark.parametrize("fid", FIELDS, ids=FIELD_IDS)
def test_host_scalars_match_oracle(oracle, fid):
    import hodor_b200 as H
    from hodor_b200 import field as f

    oc, c = oracle.field_constants(fid), f.constants(fid)
    assert c.modulus == oracle.limbs_to_int(oc["p"])
    assert np.array_equal(c.one, oc["r"]) and np.array_equal(c.generator, oc["generator"])
    assert np.array_equal(c.root_of_unity, oc["root_of_unity"])
    assert (c.S, c.NUM_BITS, c.CAPACITY) == (oc["s"], oc["num_bits"], oc["num_bits"] - 1)
    for ln in (0, 1, 2, 10, 24, min(28, c.S)):
        assert np.array_equal(H.Domain.new_for_size(fid, 1 << ln).generator, oracle.domain_generator(fid, ln))
    a = oracle.random_elements(fid, 8, seed=77)
    for i in range(0, 8, 2):
        assert np.array_equal(f.mul(fid, a[i], a[i + 1]), oracle.mul(fid, a[i : i + 1], a[i + 1 : i + 2])[0])
        assert np.array_equal(f.add(fid, a[i], a[i + 1]), oracle.add(fid, a[i : i + 1], a[i + 1 : i + 2])[0])
        assert np.array_equal(f.sub(fid, a[i], a[i + 1]), oracle.sub(fid, a[i : i + 1], a[i + 1 : i + 2])[0])
        assert np.array_equal(f.inverse(fid, a[i]), oracle.inverse(fid, a[i]))
        assert np.array_equal(f.pow_(fid, a[i], 12345678901), oracle.pow_(fid, a[i], 12345678901))
        assert np.array_equal(f.mul(fid, a[i], f.inverse(fid, a[i])), c.one)
    assert f.into_repr(fid, f.from_repr(fid, 123456789)) == 123456789
    assert np.array_equal(f.from_repr(fid, 7), oracle.to_mont(fid, oracle.ints_to_array([7]))[0])
    rng = np.random.default_rng(1)
    for _ in range(8):
        d = rng.integers(0, 256, 32, dtype=np.uint8).tobytes()
        assert np.array_equal(H.Blake2sLeafEncoder.interpret_hash(fid, d), oracle.interpret_hash(fid, d))
    assert np.array_equal(H.Blake2sLeafEncoder.interpret_hash(fid, b"\xff" * 32), oracle.interpret_hash(fid, b"\xff" * 32))


def test_domain_errors_like_the_reference():
    """Domain::new_for_size -> Err(SynthesisError::Error) past the 2-adicity (src/domains/mod.rs:29-32);
    size is rounded up to a power of two (:22)."""
    import hodor_b200 as H

    assert H.Domain.new_for_size(0, 5).size == 8 and H.Domain.new_for_size(0, 5).power_of_two == 3
    assert H.Domain.new_for_size(0, 1 << 32).power_of_two == 32
    with pytest.raises(H.SynthesisError):
        H.Domain.new_for_size(0, (1 << 32) + 1)
    with pytest.raises(H.SynthesisError):
        H.Domain.new_for_size(1, 1 << 29)
    assert H.Domain.coset_for_natural_index_and_size(5, 8) == [1, 5]
    assert H.Domain.index_and_size_for_next_domain(5, 8) == (1, 4)
    assert H.TrivialCombiner.get_coset_for_natural_index(1, 8) == [1, 5]
    with pytest.raises(H.HodorError):
        from hodor_b200 import field as f
        f.inverse(0, f.zero())


def test_hashlib_side_matches_oracle_hashing(oracle):
    import hodor_b200 as H

    a = oracle.random_elements(0, 4, seed=2)
    assert H.Blake2sTreeHasher.hash_leaf(a[0]) == oracle.hash_leaf(0, a[0])
    l, r = oracle.hash_leaf(0, a[1]), oracle.hash_leaf(0, a[2])
    assert H.Blake2sTreeHasher.hash_node([l, r]) == oracle.hash_node(l, r)


@pytest.fixture(scope="module")
def host_field_shim(tmp_path_factory):
    so = tmp_path_factory.mktemp("shim") / "field_host_shim.so"
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-Wno-unknown-pragmas", "-o", str(so),
                           os.path.join(ROOT, "tests", "host", "field_host_shim.cpp")])
    return C.CDLL(str(so))


@pytest.mark.parametrize("fid", FIELDS, ids=FIELD_IDS)
def test_device_multiplier_algorithm_on_host(oracle, pymodel, host_field_shim, fid):
    """csrc/field.cuh compiled for the host (carry chains emulated): the even/odd Montgomery
    multiplier, add, sub, halve, to/from Montgomery against the oracle, incl. edge values."""
    F = getattr(pymodel, {0: "BLS12_381_FR", 1: "BN254_FR", 2: "STARK252"}[fid])
    a, b = oracle.random_elements(fid, 4000, 1), oracle.random_elements(fid, 4000, 2)
    edge = oracle.ints_to_array([0, 1, F.p - 1, F.R, F.p - 2, 2, (F.p - 1) // 2, (F.p + 1) // 2])
    a[:8], b[:8] = edge, edge[::-1]
    a[8:16], b[8:16] = edge, edge
    a[16:24], b[16:24] = edge, oracle.ints_to_array([F.p - 1] * 8)
    out = np.zeros_like(a)

    def run(op):
        rc = host_field_shim.host_field_op(fid, op, a.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p),
                                           out.ctypes.data_as(C.c_void_p), C.c_size_t(len(a)))
        assert rc == 0
        return out.copy()

    assert np.array_equal(run(0), oracle.mul(fid, a, b))
    assert np.array_equal(run(1), oracle.add(fid, a, b))
    assert np.array_equal(run(2), oracle.sub(fid, a, b))
    ai = oracle.array_to_ints(a)
    assert oracle.array_to_ints(run(3)) == [x * pow(2, -1, F.p) % F.p for x in ai]
    assert np.array_equal(run(4), oracle.to_mont(fid, a))
    assert np.array_equal(run(5), oracle.from_mont(fid, a))
    assert oracle.array_to_ints(run(6)) == [(-x) % F.p for x in ai]


@pytest.mark.parametrize("fid", FIELDS, ids=FIELD_IDS)
def test_fixed_operand_multiplier_on_host(oracle, pymodel, host_field_shim, fid):
    """Field::mul_pre / make_pre (csrc/field.cuh, host build): same bits as the Montgomery product, the
    table quotient is floor(w * 2^256 / p), and the out-of-line carry helper equals the dropped part of
    the truncated high product (model of the device code path: tools/gen_shoup.py)."""
    F = getattr(pymodel, {0: "BLS12_381_FR", 1: "BN254_FR", 2: "STARK252"}[fid])
    n = 4000
    a, b = oracle.random_elements(fid, n, 3), oracle.random_elements(fid, n, 4)
    edge = oracle.ints_to_array([0, 1, F.p - 1, F.R, F.p - 2, 2, (F.p - 1) // 2, (F.p + 1) // 2])
    a[:8], b[:8] = edge, edge[::-1]
    a[8:16], b[8:16] = edge, edge
    a[16:24], b[16:24] = edge, oracle.ints_to_array([F.p - 1] * 8)
    out = np.zeros_like(a)

    def run(op, x=a, y=b):
        rc = host_field_shim.host_field_op(fid, op, x.ctypes.data_as(C.c_void_p), y.ctypes.data_as(C.c_void_p),
                                           out.ctypes.data_as(C.c_void_p), C.c_size_t(len(x)))
        assert rc == 0
        return out.copy()

    assert np.array_equal(run(7), oracle.mul(fid, a, b))
    # the first operand may be ANY 256-bit integer (r = a*w - q*p < 2p for every a < 2^256): result is
    # the canonical representative of a * w / R
    rng0 = np.random.default_rng(9)
    wide = rng0.integers(0, 2**64, size=(n, 4), dtype=np.uint64)
    wide[:2] = np.uint64(0xFFFFFFFFFFFFFFFF)
    got = oracle.array_to_ints(run(7, wide, b))
    Rinv = pow(F.R, -1, F.p)
    assert got == [x * y * Rinv % F.p for x, y in zip(oracle.array_to_ints(wide), oracle.array_to_ints(b))]
    plain = [F.from_mont(x) for x in oracle.array_to_ints(b)]
    assert oracle.array_to_ints(run(8)) == [(w << 256) // F.p for w in plain]
    # truncated high product: kept = words >= 7 (+ high words of column 6); dropped part D < 14 * 2^224 and
    # pre_dropped_carry == D >> 224, for arbitrary 256-bit operands (not only field elements)
    rng = np.random.default_rng(5)
    x = rng.integers(0, 2**64, size=(n, 4), dtype=np.uint64)
    y = rng.integers(0, 2**64, size=(n, 4), dtype=np.uint64)
    x[:3] = y[:3] = np.uint64(0xFFFFFFFFFFFFFFFF)
    x[3:6, :3] = np.uint64(0xFFFFFFFFFFFFFFFF)
    got = run(9, x, y)[:, 0] & np.uint64(0xFFFFFFFF)
    M = (1 << 32) - 1
    for k in range(n):
        A, B = oracle.limbs_to_int(x[k]), oracle.limbs_to_int(y[k])
        aw = [(A >> (32 * i)) & M for i in range(8)]
        bw = [(B >> (32 * i)) & M for i in range(8)]
        kept = 0
        for i in range(8):
            for j in range(8):
                c, pr = i + j, aw[i] * bw[j]
                if c >= 7:
                    kept += pr << (32 * c)
                elif c == 6:
                    kept += (pr >> 32) << (32 * 7)
        D = A * B - kept
        assert 0 <= D < 14 << 224
        assert int(got[k]) == D >> 224
        guard, q = (kept >> 224) & M, kept >> 256
        assert q + ((guard + (D >> 224)) >> 32) == (A * B) >> 256
        if guard < 0xFFFFFFF2:
            assert q == (A * B) >> 256  # the fast path needs no fix-up


def test_generated_carry_chains_are_in_sync():
    """csrc/shoup_rows.cuh is generated: the committed file must be what tools/gen_shoup.py emits."""
    import subprocess
    import sys
    want = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "gen_shoup.py")], capture_output=True, text=True,
                          check=True, env={k: v for k, v in os.environ.items() if k not in ("IMM_ROLE", "IMM_FIRST", "SWAP_HI")}).stdout
    assert open(os.path.join(ROOT, "hodor_b200", "csrc", "shoup_rows.cuh")).read() == want


def test_no_gpu_means_loud_failure_not_cpu_compute():
    """Only meaningful where no GPU is visible (the build container): the product must refuse."""
    import hodor_b200 as H
    from hodor_b200 import _ffi

    if _ffi.lib.hodor_cuda_device_count() > 0:
        pytest.skip("a GPU is visible; the refusal path is exercised on the CPU-only box")
    with pytest.raises(H.HodorError):
        H.init(0)
    a = np.zeros((4, 4), np.uint64)
    with pytest.raises(H.HodorError):
        H.Polynomial.from_coeffs(0, a).fft()
    with pytest.raises(H.HodorError):
        H.Blake2sIopTree.create(0, a)
    assert "no" in _ffi.last_error().lower()


def test_plan_covers_all_sizes():
    """The pass plan used by csrc (context.h make_plan), restated: digits in 6..9 summing to log_n, at
    most four passes, for both plan policies (HODOR_NTT_MAX_DIGIT = 8 default, 9)."""
    for md in (8, 9):
        for ln in range(12, 33):
            passes = (ln + md - 1) // md
            if ln // passes < 6 and passes > 2:
                passes -= 1
            base, rem = divmod(ln, passes)
            digits = [base + (1 if i < rem else 0) for i in range(passes)]
            assert sum(digits) == ln and all(6 <= d <= 9 for d in digits) and passes <= 4, (md, ln, digits)


def _plan(ln, md=8):
    passes = (ln + md - 1) // md
    if ln // passes < 6 and passes > 2:
        passes -= 1
    base, rem = divmod(ln, passes)
    return [base + (1 if i < rem else 0) for i in range(passes)]


def _groups(b):
    """csrc/ntt.cuh Groups<B>: widths of the register-resident butterfly groups of a 2^B tile."""
    if b in (7, 8):
        r = [1 if b == 7 else 2, 2, 2]
    else:
        rem = b - 3
        r2 = rem if rem <= 3 else (rem + 1) // 2
        r = [3, r2, rem - r2]
    r.append(b - sum(r))
    return r


def _local_out_index(b, pos):
    """csrc/ntt.cuh local_out_index<B>: tile position (digits k1 | k2 | k3 | k4, MSB first) -> local output index."""
    r1, r2, r3, r4 = _groups(b)
    k1 = pos >> (r2 + r3 + r4)
    k2 = (pos >> (r3 + r4)) & ((1 << r2) - 1)
    k3 = (pos >> r4) & ((1 << r3) - 1)
    k4 = pos & ((1 << r4) - 1)
    return k1 | (k2 << r1) | (k3 << (r1 + r2)) | (k4 << (r1 + r2 + r3))


def test_last_pass_tile_columns_are_aligned_runs_of_eight_outputs():
    """Index map of the last pass (csrc/ntt.cuh store_global, csrc/ntt_commit.cuh out_index), restated.  What the fused
    last pass + commit kernel relies on: for every plan and every blowup the 8 columns of one tile position are 8
    ADJACENT, 8-ALIGNED outputs of the interleaved natural-order vector -- one complete 2^3 subtree of leaves -- and
    over all blocks every output is produced exactly once."""
    for ln, log_l in [(12, 0), (12, 1), (12, 2), (12, 3), (13, 4), (14, 3), (15, 0), (16, 3), (17, 2), (18, 3), (18, 1), (25, 0)]:  # the last one: four passes (7+6+6+6), two middle digits
        digits = _plan(ln)
        b, b1 = digits[-1], digits[0]
        mid0 = digits[1] if len(digits) >= 3 else 0
        mid1 = digits[2] if len(digits) >= 4 else 0
        li = min(log_l, 3)
        m = ln - b - b1
        blocks = (1 << (ln + log_l)) >> (b + 3)
        seen = np.zeros(1 << (ln + log_l), np.uint8)
        for block in range(blocks):
            t = block
            mid = t & ((1 << m) - 1)
            t >>= m
            k1_hi = t & ((1 << (b1 - (3 - li))) - 1)
            coset_hi = t >> (b1 - (3 - li))
            midrev = ((mid >> mid1) | ((mid & ((1 << mid1) - 1)) << mid0)) if mid1 else mid
            # vectorised over the tile: positions x columns
            pos = np.arange(1 << b)
            kloc = np.array([_local_out_index(b, int(x)) for x in pos], np.int64) if block == 0 else kloc  # noqa: F821
            c = np.arange(8)
            i = (coset_hi << li) | (c & ((1 << li) - 1))
            k1 = (k1_hi << (3 - li)) | (c >> li)
            k = k1[None, :] | (midrev << b1) | (kloc[:, None] << (ln - b))
            out = i[None, :] + (k << log_l)
            assert np.all(out[:, 0] % 8 == 0) and np.all(out == out[:, :1] + c[None, :]), (ln, log_l, block)
            seen[out.ravel()] += 1
        assert np.all(seen == 1), (ln, log_l, digits)


@pytest.mark.parametrize("fid", FIELDS, ids=FIELD_IDS)
def test_fri_verify_proof_on_oracle_built_proofs(oracle, fid):
    """NaiveFriIop.verify_proof (src/fri/verifier.rs:130-290, host-side scalar work through the
    library's hodor_field_* helpers) accepts proofs assembled from the oracle's commit chain of a
    low-degree LDE -- queries laid out as produce_proof does (src/fri/query_producer.rs:10-53) -- and
    rejects a wrong claimed value, a tampered leaf, a tampered path and a wrong final coefficient."""
    from hodor_b200.domains import Domain
    from hodor_b200.fri import FRIProof, NaiveFriIop
    from hodor_b200.iop import TrivialBlake2sIopQuery, TrivialCombiner

    log_n, L, out_coeffs = 5, 4, 1  # the reference's own shape (src/fri/mod.rs:436): one final coefficient, odd query index
    coeffs = oracle.random_elements(fid, 1 << log_n, seed=77)
    lde = oracle.lde(fid, coeffs, log_n, L, True)
    proto = oracle.fri_commit(fid, lde, L, out_coeffs)
    layers = [(proto.l0_nodes, lde)] + list(zip(proto.layer_nodes, proto.layer_values))

    def make_proof(index):
        queries, roots = [], []
        size, idx = lde.shape[0], index
        for nodes, values in layers:
            for c in TrivialCombiner.get_coset_for_natural_index(idx, size):
                queries.append(TrivialBlake2sIopQuery(c, values[c].copy(), oracle.merkle_path(fid, nodes, values, c)))
            roots.append(nodes[1].tobytes())
            idx, size = Domain.index_and_size_for_next_domain(idx, size)
        return FRIProof(queries, roots, proto.final_coefficients.copy(), 1 << log_n, out_coeffs, L, fid)

    # the reference's domain check (verifier.rs:147-157) only lets odd indices through:
    # (w^idx)^(N/2) == 1 for every even idx is reported as "not in the LDE domain"
    from hodor_b200._ffi import SynthesisError
    for index in (0, 2, 36):
        with pytest.raises(SynthesisError):
            NaiveFriIop.verify_proof(make_proof(index), index, lde[index])
    for index in (1, 37, lde.shape[0] // 2 + 5, lde.shape[0] - 1):
        proof = make_proof(index)
        assert len(proof.roots) == len(layers) and len(proof.queries) == 2 * len(layers)
        assert NaiveFriIop.verify_proof(proof, index, lde[index]) is True
        assert NaiveFriIop.verify_proof(proof, index, lde[(index + 1) % lde.shape[0]]) is False
    index = 37
    bad = make_proof(index)
    bad.queries[2]._value = bad.queries[2]._value.copy()
    bad.queries[2]._value[0] ^= np.uint64(1)
    assert NaiveFriIop.verify_proof(bad, index, lde[index]) is False
    bad = make_proof(index)
    bad.queries[1]._path = [bytes(32)] + bad.queries[1]._path[1:]
    assert NaiveFriIop.verify_proof(bad, index, lde[index]) is False
    bad = make_proof(index)
    bad.final_coefficients[0, 0] ^= np.uint64(1)
    assert NaiveFriIop.verify_proof(bad, index, lde[index]) is False
