"""Known answers the REFERENCE itself holds for its field arithmetic, recomputed here.

src/experiments/square_root_calculator/fp2.rs hard-codes, as raw Montgomery limbs of `experiments::Fr`
(the Stark-252 field, R = 2^256), the output of its own test `find_c` (fp2.rs:358-412): with
Fq2 = Fq[u] / (u^2 - NON_RESIDUE), c = the first of 1+u, 2+u, ... whose xi(c) = c^((q^2-1)/2) is not one,
d = c^((q-1)/2), E = (d c)^-1 and F = (d c)^2; `E_PRECOMPUTED` (fp2.rs:51-65) and `F_PRECOMPUTED`
(fp2.rs:67-81) are what that test printed with `into_raw_repr`.  Reproducing them takes ~3000 dependent
Montgomery multiplications, additions, subtractions and one inversion over the reference's own field
type, so they pin a1 (mul / add / sub / inverse, Montgomery form, canonical reduction) to bits the
reference holds -- unlike the rest of the golden fixtures, which come from this repository's own models.
The exponents are the reference's constants too (fp2.rs:26-49) and are checked against the modulus.

The same procedure runs on three implementations: the C oracle (CPU), the library's host scalar helpers
(CPU, inside libhodor_b200.so) and the CUDA elementwise / batch-inversion kernels (`-m gpu`)."""
import numpy as np
import pytest

STARK = 2

# fp2.rs:10-22
NON_RESIDUE = [0xFFFFFFFFFFFFFFA1, 0xFFFFFFFFFFFFFFFF, 0xFFFFFFFFFFFFFFFF, 0x07FFFFFFFFFFF9B0]
MINUS_ONE = [0x0000000000000020, 0x0, 0x0, 0x0000000000000220]
# fp2.rs:26-49 (little-endian u64 words)
Q_SQUARED_MINUS_ONE_BY_TWO = [0x0, 0x0, 0x0, 0x0800000000000011, 0x0, 0x8000000000000000, 0x8800000000000090,
                              0x20000000000000]
Q_MINUS_ONE_BY_TWO = [0x0, 0x0, 0x8000000000000000, 0x400000000000008]
Q_MINUS_ONE_BY_FOUR = [0x0, 0x0, 0x4000000000000000, 0x200000000000004]
# fp2.rs:51-81
E_PRECOMPUTED = ([0, 0, 0, 0],
                 [0xB11079DBFDB6981F, 0x701BB8E5E5B53751, 0x6FE88E46D707BCAB, 0x02365BB6D67E6298])
F_PRECOMPUTED = ([0xFFFFFFFFFFFFFF41, 0xFFFFFFFFFFFFFFFF, 0xFFFFFFFFFFFFFFFF, 0x07FFFFFFFFFFF350],
                 [0, 0, 0, 0])


def _words(ws):
    return sum(int(w) << (64 * i) for i, w in enumerate(ws))


class Fq2:
    """fp2.rs's Fq2 over an `ops` backend working on (k, 4) uint64 arrays of Montgomery limbs; every one
    of the k lanes carries the same computation, so a vector backend is exercised on all its lanes."""

    def __init__(self, ops, lanes):
        self.o, self.k = ops, lanes
        self.nr = np.tile(np.array(NON_RESIDUE, np.uint64), (lanes, 1))

    def const(self, limbs):
        return np.tile(np.array(limbs, np.uint64), (self.k, 1))

    def mul(self, a, b):                       # fp2.rs:212-228 (Karatsuba form)
        o = self.o
        aa, bb = o.mul(a[0], b[0]), o.mul(a[1], b[1])
        s = o.mul(o.add(a[0], a[1]), o.add(b[0], b[1]))
        return o.add(aa, o.mul(bb, self.nr)), o.sub(o.sub(s, aa), bb)

    def square(self, a):
        return self.mul(a, a)

    def pow(self, a, exp_words, one):          # ff::Field::pow: MSB-first square and multiply
        e, res, started = _words(exp_words), one, False
        for i in reversed(range(64 * len(exp_words))):
            if started:
                res = self.square(res)
            if (e >> i) & 1:
                started = True
                res = self.mul(res, a)
        return res

    def inverse(self, a):                      # fp2.rs:230-254: conj(a) / norm(a)
        o = self.o
        n = o.sub(o.mul(a[0], a[0]), o.mul(o.mul(a[1], a[1]), self.nr))
        ni = o.inv(n)
        return o.mul(a[0], ni), o.mul(o.sub(self.const([0] * 4), a[1]), ni)


def find_c(ops, one_limbs, lanes=1):
    """fp2.rs:358-412.  Returns (c, E, F) as pairs of (lanes, 4) limb arrays."""
    K = Fq2(ops, lanes)
    one = K.const(one_limbs)
    zero = K.const([0] * 4)
    one2 = (one, zero)
    c = (one, one)
    for _ in range(8):
        xi = K.pow(c, Q_SQUARED_MINUS_ONE_BY_TWO, one2)
        if not (np.array_equal(xi[0], one) and np.array_equal(xi[1], zero)):
            break
        c = (ops.add(c[0], one), c[1])
    else:
        raise AssertionError("no non-residue found")
    d = K.pow(c, Q_MINUS_ONE_BY_TWO, one2)
    dc = K.mul(d, c)
    e = K.inverse(dc)
    f = K.square(dc)
    may_be_one = K.mul(e, dc)
    assert np.array_equal(may_be_one[0], one) and np.array_equal(may_be_one[1], zero)
    return c, e, f


def _check(c, e, f, one_limbs, lanes):
    def rows(limbs):
        return np.tile(np.array(limbs, np.uint64), (lanes, 1))
    assert np.array_equal(e[0], rows(E_PRECOMPUTED[0])) and np.array_equal(e[1], rows(E_PRECOMPUTED[1]))
    assert np.array_equal(f[0], rows(F_PRECOMPUTED[0])) and np.array_equal(f[1], rows(F_PRECOMPUTED[1]))


def test_reference_exponents_belong_to_the_declared_modulus(oracle):
    q = oracle.limbs_to_int(oracle.field_constants(STARK)["p"])
    assert q == 2**251 + 17 * 2**192 + 1          # src/experiments/mod.rs:19
    assert _words(Q_MINUS_ONE_BY_TWO) == (q - 1) // 2
    assert _words(Q_MINUS_ONE_BY_FOUR) == (q - 1) // 4
    assert _words(Q_SQUARED_MINUS_ONE_BY_TWO) == (q * q - 1) // 2
    assert _words(MINUS_ONE) == (q - 1) * 2**256 % q and _words(NON_RESIDUE) == 3 * 2**256 % q


def test_find_c_constants_with_the_c_oracle(oracle):
    class Ops:
        mul = staticmethod(lambda a, b: oracle.mul(STARK, a, b))
        add = staticmethod(lambda a, b: oracle.add(STARK, a, b))
        sub = staticmethod(lambda a, b: oracle.sub(STARK, a, b))
        inv = staticmethod(lambda a: np.stack([oracle.inverse(STARK, x) for x in a]).reshape(-1, 4))

    one = oracle.field_constants(STARK)["r"]
    c, e, f = find_c(Ops, one, lanes=2)
    assert oracle.array_to_ints(oracle.from_mont(STARK, c[0][:1])) == [3]      # c = 3 + u
    _check(c, e, f, one, 2)


def test_find_c_constants_with_the_bigint_model(pymodel):
    F = pymodel.STARK252
    to_arr = lambda xs: np.array([[(x >> (64 * i)) & (2**64 - 1) for i in range(4)] for x in xs], np.uint64)
    to_int = lambda a: [_words(r) for r in a]
    Rinv = pow(F.R, -1, F.p)

    class Ops:      # Montgomery-form integers: a (*) b = a b R^-1
        mul = staticmethod(lambda a, b: to_arr([x * y * Rinv % F.p for x, y in zip(to_int(a), to_int(b))]))
        add = staticmethod(lambda a, b: to_arr([(x + y) % F.p for x, y in zip(to_int(a), to_int(b))]))
        sub = staticmethod(lambda a, b: to_arr([(x - y) % F.p for x, y in zip(to_int(a), to_int(b))]))
        inv = staticmethod(lambda a: to_arr([pow(x * Rinv % F.p, -1, F.p) * F.R % F.p for x in to_int(a)]))

    one = [(F.R % F.p >> (64 * i)) & (2**64 - 1) for i in range(4)]
    c, e, f = find_c(Ops, one, lanes=1)
    _check(c, e, f, one, 1)


def test_find_c_constants_with_the_library_host_helpers():
    """hodor_field_mul / add / sub / inverse: host scalar code inside libhodor_b200.so (no GPU needed)."""
    from hodor_b200 import field as fld

    class Ops:
        mul = staticmethod(lambda a, b: np.stack([fld.mul(STARK, x, y) for x, y in zip(a, b)]))
        add = staticmethod(lambda a, b: np.stack([fld.add(STARK, x, y) for x, y in zip(a, b)]))
        sub = staticmethod(lambda a, b: np.stack([fld.sub(STARK, x, y) for x, y in zip(a, b)]))
        inv = staticmethod(lambda a: np.stack([fld.inverse(STARK, x) for x in a]))

    one = fld.one(STARK)
    c, e, f = find_c(Ops, one, lanes=1)
    _check(c, e, f, one, 1)


@pytest.mark.gpu
def test_find_c_constants_on_the_gpu(hodor):
    """The same chain through the C ABI's device kernels: `hodor_cuda_elementwise` (mul / add / sub) and
    `hodor_cuda_batch_inversion`, on vectors of 37 lanes (ragged: not a multiple of any tile)."""
    import ctypes as C

    from hodor_b200 import _ffi
    from hodor_b200 import field as fld
    from hodor_b200.field import _p

    lanes = 37

    def ew(op, a, b):
        a, b = np.ascontiguousarray(a), np.ascontiguousarray(b)
        out = np.zeros_like(a)
        _ffi.check(_ffi.lib.hodor_cuda_elementwise(op, _p(a), _p(b), _p(out), C.c_uint64(a.shape[0]), STARK))
        return out

    def batch_inv(a):
        a = np.ascontiguousarray(a).copy()
        _ffi.check(_ffi.lib.hodor_cuda_batch_inversion(_p(a), C.c_uint64(a.shape[0]), STARK))
        return a

    class Ops:
        mul = staticmethod(lambda a, b: ew(0, a, b))
        add = staticmethod(lambda a, b: ew(1, a, b))
        sub = staticmethod(lambda a, b: ew(2, a, b))
        inv = staticmethod(batch_inv)

    one = fld.one(STARK)
    c, e, f = find_c(Ops, one, lanes=lanes)
    _check(c, e, f, one, lanes)
