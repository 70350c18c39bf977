// Host build of hodor_b200/csrc/field.cuh (carry chains emulated in C) so the even/odd Montgomery
// multiplier can be checked on a machine without a GPU.  Test-only; never shipped.
#include <cstddef>
#include "../../hodor_b200/csrc/field.cuh"
using namespace hodor;

template <class F>
static void run(int op, const Fe* a, const Fe* b, Fe* out, size_t n) {
    const Field<F> f;
    for (size_t i = 0; i < n; i++) {
        switch (op) {
            case 0: out[i] = f.mul(a[i], b[i]); break;
            case 1: out[i] = f.add(a[i], b[i]); break;
            case 2: out[i] = f.sub(a[i], b[i]); break;
            case 3: out[i] = f.halve(a[i]); break;
            case 4: out[i] = f.to_mont(a[i]); break;
            case 5: out[i] = f.from_mont(a[i]); break;
            case 6: out[i] = f.neg(a[i]); break;
            case 7: {  // fixed-operand multiplier: a * b with b converted to (w, floor(w * 2^256 / p))
                Fe w, q;
                f.make_pre(b[i], w, q);
                out[i] = f.mul_pre(a[i], w, q);
                break;
            }
            case 8: {  // the quotient word table itself: out = q of make_pre(b)
                Fe w;
                f.make_pre(b[i], w, out[i]);
                break;
            }
            case 9: {  // pre_dropped_carry(a, b) in word 0
                out[i] = Field<F>::zero();
                out[i].v[0] = Field<F>::pre_dropped_carry(a[i].v, b[i].v);
                break;
            }
        }
    }
}
extern "C" int host_field_op(int field, int op, const uint32_t* a, const uint32_t* b, uint32_t* out, size_t n) {
    const Fe *A = (const Fe*)a, *B = (const Fe*)b;
    Fe* O = (Fe*)out;
    switch (field) {
        case 0: run<BlsFr>(op, A, B, O, n); return 0;
        case 1: run<Bn254Fr>(op, A, B, O, n); return 0;
        case 2: run<Stark252>(op, A, B, O, n); return 0;
    }
    return -1;
}
