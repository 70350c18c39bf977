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
