"""PrecomputedOmegas (src/precomputations/mod.rs:14-66) and the ALI inverse divisors
(src/ali/per_register/mod.rs:60-162, :214-227).  CPU: the big-int restatement satisfies the defining identities.
GPU: the device vectors equal the restatement bit for bit, through the C ABI."""
import numpy as np
import pytest

FIELDS = {0: "BLS12_381_FR", 1: "BN254_FR", 2: "STARK252"}
# (log_column, log_evaluation, start_at, span, num_rows)
DENSE_CASES = [(3, 3, 0, 1, 8), (3, 5, 0, 1, 8), (4, 6, 2, 1, 16), (4, 7, 0, 2, 13), (5, 5, 3, 4, 30), (6, 10, 0, 1, 64)]


def test_model_divisor_identities(pymodel):
    F = pymodel.BLS12_381_FR
    p = F.p
    for lc, le, start_at, span, num_rows in DENSE_CASES:
        T, E = 1 << lc, 1 << le
        inv, degree = pymodel.inverse_divisor_for_dense_constraint_in_coset(F, lc, le, start_at, span, num_rows)
        assert degree == num_rows - start_at - span and len(inv) == E
        w_col, w_eval = F.domain_generator(lc), F.domain_generator(le)
        keep = [pow(w_col, k, p) for k in range(start_at, num_rows - span)]  # rows the constraint holds on
        for j in (0, 1, E // 2, E - 1):
            x = F.generator * pow(w_eval, j, p) % p
            divisor = 1  # the divisor is prod over the rows the constraint holds on
            for r in keep:
                divisor = divisor * (x - r) % p
            assert inv[j] * divisor % p == 1
    # the Fibonacci shape: (x - omega^(T-1)) / (x^T - 1), as oracle/fib_model.py states it
    lc = 4
    T = 1 << lc
    inv, _ = pymodel.inverse_divisor_for_dense_constraint_in_coset(F, lc, lc, 0, 1, T)
    omega = F.domain_generator(lc)
    xs = [F.generator * pow(omega, i, p) % p for i in range(T)]
    assert inv == [(x - pow(omega, T - 1, p)) * pow(pow(x, T, p) - 1, -1, p) % p for x in xs]
    b = pymodel.boundary_constraint_inverse_divisor(F, lc, lc + 2, 3)
    w_eval = F.domain_generator(lc + 2)
    assert all(b[j] * (F.generator * pow(w_eval, j, p) - pow(omega, 3, p)) % p == 1 for j in range(4 * T))


def test_model_precomputed_omegas(pymodel):
    F = pymodel.STARK252
    om, co, inv = pymodel.precomputed_omegas(F, 5)
    assert len(om) == len(co) == 32 and len(inv) == 16
    assert om[0] == 1 and om[16] == F.p - 1 and pow(om[1], 32, F.p) == 1
    assert all(a * b % F.p == 1 for a, b in zip(om[:16], inv))
    assert all(c == o * F.generator % F.p for c, o in zip(co, om))


def _limbs(oracle, F, xs):
    return np.stack([oracle.int_to_limbs(F.to_mont(x)) for x in xs]) if xs else np.zeros((0, 4), np.uint64)


@pytest.mark.gpu
@pytest.mark.parametrize("fid", [0, 1, 2])
def test_precomputed_omegas_on_device(hodor, oracle, pymodel, fid):
    from hodor_b200.precomputations import PrecomputedOmegas
    F = getattr(pymodel, FIELDS[fid])
    for log_n in (0, 1, 4, 9, 13):
        pre = PrecomputedOmegas.new_for_domain(hodor.Domain.new_for_size(fid, 1 << log_n))
        om, co, inv = pymodel.precomputed_omegas(F, log_n)
        from hodor_b200 import device as dev
        assert np.array_equal(dev.to_host(pre.omegas), _limbs(oracle, F, om))
        assert np.array_equal(dev.to_host(pre.coset), _limbs(oracle, F, co))
        assert np.array_equal(dev.to_host(pre.omegas_inv).reshape(-1, 4), _limbs(oracle, F, inv))
        assert np.array_equal(pre.coset_values().to_host(), _limbs(oracle, F, co))
        from hodor_b200.precomputations import precomputed_omegas_host
        h_om, h_co, h_inv = precomputed_omegas_host(hodor.Domain.new_for_size(fid, 1 << log_n))
        assert np.array_equal(h_om, _limbs(oracle, F, om)) and np.array_equal(h_co, _limbs(oracle, F, co))
        assert np.array_equal(h_inv, _limbs(oracle, F, inv))


@pytest.mark.gpu
@pytest.mark.parametrize("fid", [0, 1, 2])
def test_ali_inverse_divisors_on_device(hodor, oracle, pymodel, fid):
    from hodor_b200 import precomputations as P
    F = getattr(pymodel, FIELDS[fid])
    for lc, le, start_at, span, num_rows in DENSE_CASES:
        col, ev = hodor.Domain.new_for_size(fid, 1 << lc), hodor.Domain.new_for_size(fid, 1 << le)
        got, degree = P.inverse_divisor_for_dense_constraint_in_coset(col, ev, P.DenseConstraint(start_at, span), num_rows)
        want, want_degree = pymodel.inverse_divisor_for_dense_constraint_in_coset(F, lc, le, start_at, span, num_rows)
        assert degree == want_degree
        assert np.array_equal(got.to_host(), _limbs(oracle, F, want)), (lc, le, start_at, span, num_rows)
        h_got, h_degree = P.inverse_divisor_for_dense_constraint_in_coset_host(col, ev, P.DenseConstraint(start_at, span), num_rows)
        assert h_degree == want_degree and np.array_equal(h_got.as_ref(), _limbs(oracle, F, want))
    for lc, le, row in [(3, 3, 0), (3, 5, 7), (5, 9, 11), (4, 12, 15)]:
        col, ev = hodor.Domain.new_for_size(fid, 1 << lc), hodor.Domain.new_for_size(fid, 1 << le)
        got = P.boundary_constraint_inverse_divisor(col, ev, row)
        want = _limbs(oracle, F, pymodel.boundary_constraint_inverse_divisor(F, lc, le, row))
        assert np.array_equal(got.to_host(), want)
        assert np.array_equal(P.boundary_constraint_inverse_divisor_host(col, ev, row).as_ref(), want)


@pytest.mark.gpu
def test_ali_divisor_argument_errors(hodor):
    from hodor_b200 import precomputations as P
    col, ev = hodor.Domain.new_for_size(0, 16), hodor.Domain.new_for_size(0, 64)
    with pytest.raises(hodor.HodorError):   # evaluation domain smaller than the column domain
        P.inverse_divisor_for_dense_constraint_in_coset(ev, col, P.DenseConstraint(0, 1), 64)
    with pytest.raises(hodor.HodorError):   # more rows than the column domain holds
        P.inverse_divisor_for_dense_constraint_in_coset(col, ev, P.DenseConstraint(0, 1), 17)
    with pytest.raises(hodor.HodorError):   # start_at + span beyond num_rows (the reference's usize subtraction underflows)
        P.inverse_divisor_for_dense_constraint_in_coset(col, ev, P.DenseConstraint(12, 8), 16)
