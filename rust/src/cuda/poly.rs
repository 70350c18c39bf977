//! `Polynomial`-level entry points: one library call per reference method instead of one `best_fft`
//! per coset (src/polynomials/mod.rs:418-482, 544-609 clone the coefficients per coset, transform
//! each clone and interleave; the library does all L cosets in one launch sequence and writes the
//! interleaved vector directly).
use crate::domains::Domain;
use crate::fft::multicore::Worker;
use crate::polynomials::*;
use crate::SynthesisError;

use super::ffi::{self, CudaField};

/// `Polynomial::lde` / `coset_lde` (src/polynomials/mod.rs:343-352).
pub fn cuda_lde<F: CudaField>(
    poly: Polynomial<F, Coefficients>,
    _worker: &Worker,
    factor: usize,
    coset: bool,
) -> Result<Polynomial<F, Values>, SynthesisError> {
    assert!(factor.is_power_of_two()); // :434, :560
    ffi::init();
    let n = poly.size();
    let _ = Domain::<F>::new_for_size((n * factor) as u64)?; // same Err as :435 / :561
    let mut out = vec![F::zero(); n * factor];
    ffi::check(unsafe {
        ffi::hodor_cuda_lde(
            ffi::as_u64(poly.as_ref()),
            poly.exp,
            factor.trailing_zeros(),
            coset as i32,
            ffi::as_u64_mut(&mut out),
            F::FIELD_ID,
        )
    })?;
    Polynomial::from_values(out)
}

/// The register loop of `Prover::prove` (src/prover/mod.rs:73-76) in one pipelined call.
pub fn cuda_lde_batch<F: CudaField>(
    polys: &[Polynomial<F, Coefficients>],
    _worker: &Worker,
    factor: usize,
    coset: bool,
) -> Result<Vec<Polynomial<F, Values>>, SynthesisError> {
    assert!(factor.is_power_of_two());
    if polys.is_empty() {
        return Ok(vec![]);
    }
    ffi::init();
    let n = polys[0].size();
    assert!(polys.iter().all(|p| p.size() == n));
    let _ = Domain::<F>::new_for_size((n * factor) as u64)?;
    let mut outs: Vec<Vec<F>> = polys.iter().map(|_| vec![F::zero(); n * factor]).collect();
    let in_ptrs: Vec<*const u64> = polys.iter().map(|p| ffi::as_u64(p.as_ref())).collect();
    let out_ptrs: Vec<*mut u64> = outs.iter_mut().map(|o| ffi::as_u64_mut(o)).collect();
    ffi::check(unsafe {
        ffi::hodor_cuda_lde_batch(
            in_ptrs.as_ptr(),
            out_ptrs.as_ptr(),
            polys.len() as u32,
            polys[0].exp,
            factor.trailing_zeros(),
            coset as i32,
            F::FIELD_ID,
        )
    })?;
    outs.into_iter().map(Polynomial::from_values).collect()
}

/// `Polynomial::<F, Values>::batch_inversion` (src/polynomials/mod.rs:889-954).  A zero value is
/// HODOR_ERR_NOT_INVERTIBLE -> `SynthesisError::Error` with the vector untouched, as at :919.
pub fn cuda_batch_inversion<F: CudaField>(poly: &mut Polynomial<F, Values>, _worker: &Worker) -> Result<(), SynthesisError> {
    ffi::init();
    let n = poly.size() as u64;
    ffi::check(unsafe { ffi::hodor_cuda_batch_inversion(ffi::as_u64_mut(poly.as_mut()), n, F::FIELD_ID) }).map(|_| ())
}

/// `Polynomial::<F, Coefficients>::evaluate_at` (src/polynomials/mod.rs:685-711).
pub fn cuda_evaluate_at<F: CudaField>(poly: &Polynomial<F, Coefficients>, _worker: &Worker, g: F) -> F {
    ffi::init();
    let mut out = F::zero();
    let rc = unsafe {
        ffi::hodor_cuda_evaluate_at(
            ffi::as_u64(poly.as_ref()),
            poly.size() as u64,
            ffi::elem(&g),
            &mut out as *mut F as *mut u64,
            F::FIELD_ID,
        )
    };
    assert!(rc == ffi::OK, "hodor_cuda_evaluate_at failed: {}", ffi::last_error());
    out
}

// ---- elementwise methods (src/polynomials/mod.rs:59-83, 640-683, 744-771, 817-887) ---------------------
// op codes of include/hodor_b200.h (HODOR_OP_*)
pub const OP_MUL: i32 = 0;
pub const OP_ADD: i32 = 1;
pub const OP_SUB: i32 = 2;
pub const OP_SCALE: i32 = 3;
pub const OP_ADD_SCALED: i32 = 4;
pub const OP_ADD_CONST: i32 = 5;
pub const OP_NEGATE: i32 = 6;
pub const OP_SQUARE: i32 = 7;
pub const OP_POW: i32 = 8;

fn poly_op<F: CudaField>(op: i32, a: &mut [F], b: Option<&[F]>, scalar: Option<&F>, exp: u64) {
    ffi::init();
    // `other` may be shorter than `self` in the reference's add / sub (:640-652): only the common prefix is touched
    let n = b.map(|b| if op == OP_SCALE { a.len() } else { b.len() }).unwrap_or(a.len());
    assert!(a.len() >= n);
    let rc = unsafe {
        ffi::hodor_cuda_poly_op(
            op,
            ffi::as_u64(a),
            b.map(|b| ffi::as_u64(b)).unwrap_or(std::ptr::null()),
            scalar.map(|s| ffi::elem(s)).unwrap_or(std::ptr::null()),
            exp,
            ffi::as_u64_mut(a),
            n as u64,
            F::FIELD_ID,
        )
    };
    assert!(rc == ffi::OK, "hodor_cuda_poly_op({}) failed: {}", op, ffi::last_error());
}

/// `add_assign` (:640-652 / :817-829), `sub_assign` (:671-683 / :860-872), `mul_assign` (:874-887).
pub fn cuda_add_assign<F: CudaField, P: PolynomialForm>(a: &mut Polynomial<F, P>, _worker: &Worker, other: &Polynomial<F, P>) {
    poly_op(OP_ADD, a.as_mut(), Some(other.as_ref()), None, 0)
}
pub fn cuda_sub_assign<F: CudaField, P: PolynomialForm>(a: &mut Polynomial<F, P>, _worker: &Worker, other: &Polynomial<F, P>) {
    poly_op(OP_SUB, a.as_mut(), Some(other.as_ref()), None, 0)
}
pub fn cuda_mul_assign<F: CudaField>(a: &mut Polynomial<F, Values>, _worker: &Worker, other: &Polynomial<F, Values>) {
    assert!(a.size() == other.size());
    poly_op(OP_MUL, a.as_mut(), Some(other.as_ref()), None, 0)
}
/// `scale` (:59-70): every element times `g`.
pub fn cuda_scale<F: CudaField, P: PolynomialForm>(a: &mut Polynomial<F, P>, _worker: &Worker, g: F) {
    if g == F::one() {
        return;
    }
    poly_op(OP_SCALE, a.as_mut(), Some(std::slice::from_ref(&g)), None, 0)
}
/// `negate` (:72-83).
pub fn cuda_negate<F: CudaField, P: PolynomialForm>(a: &mut Polynomial<F, P>, _worker: &Worker) {
    poly_op(OP_NEGATE, a.as_mut(), None, None, 0)
}
/// `add_assign_scaled` (:654-669 / :843-858): a += other * scaling.
pub fn cuda_add_assign_scaled<F: CudaField, P: PolynomialForm>(a: &mut Polynomial<F, P>, _worker: &Worker, other: &Polynomial<F, P>, scaling: &F) {
    poly_op(OP_ADD_SCALED, a.as_mut(), Some(other.as_ref()), Some(scaling), 0)
}
/// `add_constant` (:831-841), `square` (:760-771), `pow` (:744-758) of `Polynomial<F, Values>`.
pub fn cuda_add_constant<F: CudaField>(a: &mut Polynomial<F, Values>, _worker: &Worker, constant: &F) {
    poly_op(OP_ADD_CONST, a.as_mut(), None, Some(constant), 0)
}
pub fn cuda_square<F: CudaField>(a: &mut Polynomial<F, Values>, _worker: &Worker) {
    poly_op(OP_SQUARE, a.as_mut(), None, None, 0)
}
pub fn cuda_pow<F: CudaField>(a: &mut Polynomial<F, Values>, _worker: &Worker, exp: u64) {
    if exp == 2 {
        return cuda_square(a, _worker);
    }
    poly_op(OP_POW, a.as_mut(), None, None, exp)
}
