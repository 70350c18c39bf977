"""Small pass over every kernel family for compute-sanitizer (memcheck / racecheck):
    compute-sanitizer --tool racecheck python tools/sanitizer_target.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import hodor_b200 as H
from oracle import oracle as O

H.init(0)
W = H.Worker()
for fid in (0, 1, 2):
    for log_n in (5, 12, 13, 14, 15, 16, 18):   # single block; 2 passes with B = 6, 7, 8, 9 tiles
        a = O.random_elements(fid, 1 << log_n, seed=log_n)
        v = H.Polynomial.from_coeffs(fid, a).fft(W)
        if log_n <= 14:
            assert np.array_equal(v.as_ref(), O.fft(fid, a, log_n))
        back = v.icoset_fft(W).coset_fft(W).ifft(W)
        assert np.array_equal(back.as_ref(), a)
    a = O.random_elements(fid, 1 << 12, seed=3)
    lde = H.Polynomial.from_coeffs(fid, a).coset_lde(W, 8)
    assert np.array_equal(lde.as_ref(), O.lde(fid, a, 12, 8, True))
    proto = H.NaiveFriIop.proof_from_lde(lde, 8, 1, W)
    assert proto.get_roots() == O.fri_commit(fid, lde.as_ref(), 8, 1).roots()
    inv = H.Polynomial.from_values(fid, a)
    inv.batch_inversion(W)
    assert np.array_equal(inv.as_ref(), O.batch_inversion(fid, a))
    z = a[1]
    assert np.array_equal(H.Polynomial.from_coeffs(fid, a).evaluate_at(W, z), O.evaluate_at(fid, a, z))
    outs = H.lde_batch([H.Polynomial.from_coeffs(fid, a) for _ in range(3)], W, 4, True)
    assert all(np.array_equal(o.as_ref(), O.lde(fid, a, 12, 4, True)) for o in outs)
# round-2 entry points: committed oracles (lift and commit, concurrent hashing stream), one-call query production,
# the fused fold + commit kernel (off by default), setup vectors, the shard building blocks
from hodor_b200 import precomputations as P  # noqa: E402
from oracle import pymodel as M  # noqa: E402

for fid in (0, 2):
    polys = [O.random_elements(fid, 1 << 12, seed=60 + i) for i in range(3)]
    orcs = H.CommittedOracle.lde_commit_batch([H.Polynomial.from_coeffs(fid, a) for a in polys], 8, True)
    for a, orc in zip(polys, orcs):
        lde = O.lde(fid, a, 12, 8, True)
        nodes = O.merkle_create(fid, lde)
        assert orc.get_root() == nodes[1].tobytes() and np.array_equal(orc.values(), lde)
        for q, i in zip(orc.query_batch([0, 77, (1 << 15) - 1]), [0, 77, (1 << 15) - 1]):
            assert q.path() == O.merkle_path(fid, nodes, lde, i)
        orc.free()
    lde_poly = H.Polynomial.from_coeffs(fid, polys[0]).coset_lde(W, 8)
    proto = H.NaiveFriIop.proof_from_lde(lde_poly, 8, 1, W)
    proof = H.NaiveFriIop.prototype_into_proof(proto, lde_poly, 12345)
    assert all(H.TrivialBlake2sIOP.verify_query(q, proof.roots[k // 2]) for k, q in enumerate(proof.queries))
    F = {0: M.BLS12_381_FR, 2: M.STARK252}[fid]
    col, ev = H.Domain.new_for_size(fid, 16), H.Domain.new_for_size(fid, 256)
    got, _ = P.inverse_divisor_for_dense_constraint_in_coset(col, ev, P.DenseConstraint(1, 2), 15)
    want, _ = M.inverse_divisor_for_dense_constraint_in_coset(F, 4, 8, 1, 2, 15)
    assert np.array_equal(got.to_host(), np.stack([O.int_to_limbs(F.to_mont(x)) for x in want]))
    got = P.boundary_constraint_inverse_divisor(col, ev, 3)
    want = M.boundary_constraint_inverse_divisor(F, 4, 8, 3)
    assert np.array_equal(got.to_host(), np.stack([O.int_to_limbs(F.to_mont(x)) for x in want]))
    P.PrecomputedOmegas.new_for_domain(ev)
big = O.random_elements(0, 1 << 20, seed=9)   # 3 passes incl. expanded tables
assert np.array_equal(H.Polynomial.from_coeffs(0, big).coset_lde(W, 2).as_ref()[::2][:64],
                      H.Polynomial.from_coeffs(0, big).coset_fft(W).as_ref()[:64])
print("sanitizer target ok")
