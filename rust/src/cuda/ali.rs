//! Setup work of `Prover::new` on the GPU: the vectors of `PrecomputedOmegas::new_for_domain`
//! (src/precomputations/mod.rs:14-66) and the inverse divisors `ALIInstance::from_arp` precomputes
//! (src/ali/per_register/mod.rs:60-162 per dense constraint density, :214-227 per boundary row).
//! Each is one library call that fills a host `Vec<F>`; the values are the ones the CPU code produces
//! (field elements are canonical, so equal values are equal bits).
use crate::air::DenseConstraint;
use crate::domains::Domain;
use crate::fft::multicore::Worker;
use crate::polynomials::*;
use crate::precomputations::PrecomputedOmegas;
use crate::SynthesisError;

use super::ffi::{self, CudaField};

/// `PrecomputedOmegas::new_for_domain` (src/precomputations/mod.rs:14-66).
pub fn cuda_precomputed_omegas<F: CudaField>(domain: &Domain<F>, _worker: &Worker) -> PrecomputedOmegas<F> {
    ffi::init();
    let n = domain.size as usize;
    let mut omegas = vec![F::zero(); n];
    let mut coset = vec![F::zero(); n];
    let mut omegas_inv = vec![F::zero(); n / 2];
    let rc = unsafe {
        ffi::hodor_cuda_precomputed_omegas(
            ffi::as_u64_mut(&mut omegas),
            ffi::as_u64_mut(&mut coset),
            if n >= 2 { ffi::as_u64_mut(&mut omegas_inv) } else { std::ptr::null_mut() },
            domain.power_of_two as u32,
            F::FIELD_ID,
        )
    };
    assert!(rc == ffi::OK, "hodor_cuda_precomputed_omegas failed: {}", ffi::last_error());
    PrecomputedOmegas { omegas, coset, omegas_inv }
}

/// The inner fn `inverse_divisor_for_dense_constraint_in_coset` of `ALIInstance::from_arp`
/// (src/ali/per_register/mod.rs:60-162): same arguments, same `(values, divisor_degree)` result.
pub fn cuda_inverse_divisor_for_dense_constraint_in_coset<F: CudaField>(
    column_domain: &Domain<F>,
    evaluation_domain: &Domain<F>,
    dense_constraint: DenseConstraint,
    num_rows: u64,
    _worker: &Worker,
) -> Result<(Polynomial<F, Values>, usize), SynthesisError> {
    ffi::init();
    let mut out = vec![F::zero(); evaluation_domain.size as usize];
    let mut divisor_degree = 0u64;
    ffi::check(unsafe {
        ffi::hodor_cuda_ali_dense_inverse_divisor(
            ffi::as_u64_mut(&mut out),
            column_domain.power_of_two as u32,
            evaluation_domain.power_of_two as u32,
            dense_constraint.start_at as u64,
            dense_constraint.span as u64,
            num_rows,
            &mut divisor_degree,
            F::FIELD_ID,
        )
    })?;
    Ok((Polynomial::from_values(out)?, divisor_degree as usize))
}

/// The loop body of src/ali/per_register/mod.rs:214-227: 1 / (X - omega^row) on the coset of the
/// constraints domain (`coset_evaluate_at_domain_for_degree_one` + `batch_inversion`).
pub fn cuda_boundary_constraint_inverse_divisor<F: CudaField>(
    column_domain: &Domain<F>,
    constraints_domain: &Domain<F>,
    row: u64,
    _worker: &Worker,
) -> Result<Polynomial<F, Values>, SynthesisError> {
    ffi::init();
    let mut out = vec![F::zero(); constraints_domain.size as usize];
    ffi::check(unsafe {
        ffi::hodor_cuda_ali_boundary_inverse_divisor(
            ffi::as_u64_mut(&mut out),
            column_domain.power_of_two as u32,
            constraints_domain.power_of_two as u32,
            row,
            F::FIELD_ID,
        )
    })?;
    Polynomial::from_values(out)
}
