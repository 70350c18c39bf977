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
big = O.random_elements(0, 1 << 20, seed=9)   # 3 passes incl. expanded tables
assert np.array_equal(H.Polynomial.from_coeffs(0, big).coset_lde(W, 2).as_ref()[::2][:64],
                      H.Polynomial.from_coeffs(0, big).coset_fft(W).as_ref()[:64])
print("sanitizer target ok")
