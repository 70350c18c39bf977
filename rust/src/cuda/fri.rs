//! Seam 3 (src/fri/mod.rs:36-61): `CudaFriIop<F>: FriIop<F>`.
//!
//! `proof_from_lde` is `NaiveFriIop::proof_from_lde_by_values` (src/fri/fri_on_values.rs:11-159) as
//! one call: l0 tree, then per layer fold -> tree -> root -> challenge, final iNTT, all enqueued on
//! one CUDA stream with no host round trip; the prototype stays in HBM behind a handle.
//! `prototype_into_proof` is `produce_proof` (src/fri/query_producer.rs:10-53) as one call,
//! `hodor_cuda_fri_produce_proof` (all openings gathered on the device, one copy).  `verify_proof` is the reference's CPU verifier, unchanged.
use std::marker::PhantomData;

use crate::domains::Domain;
use crate::fft::multicore::Worker;
use crate::fri::*;
use crate::iop::blake2s_trivial_iop::TrivialBlake2sIopQuery;
use crate::iop::trivial_coset_combiner::TrivialCombiner;
use crate::iop::*;
use crate::polynomials::*;
use crate::SynthesisError;

use super::ffi::{self, CudaField};
use super::iop::CudaBlake2sIOP;

type Digest = [u8; 32];

pub struct CudaFriPrototype<F: CudaField> {
    handle: *mut ffi::FriProto,
    roots: Vec<Digest>,      // l0 root, then every intermediate root
    pub challenges: Vec<F>,  // one per fold
    final_coefficients: Vec<F>,
    pub initial_degree_plus_one: usize,
    pub output_coeffs_at_degree_plus_one: usize,
    pub lde_factor: usize,
}

unsafe impl<F: CudaField> Send for CudaFriPrototype<F> {}

impl<F: CudaField> Drop for CudaFriPrototype<F> {
    fn drop(&mut self) {
        unsafe { ffi::hodor_cuda_fri_free(self.handle) }
    }
}

impl<F: CudaField> FriProofPrototype<F, CudaBlake2sIOP<F>> for CudaFriPrototype<F> {
    fn get_roots(&self) -> Vec<Digest> {
        self.roots.clone()
    }
    fn get_final_root(&self) -> Digest {
        *self.roots.last().expect("at least the l0 root")
    }
    fn get_final_coefficients(&self) -> Vec<F> {
        self.final_coefficients.clone()
    }
}

impl<F: CudaField> CudaFriPrototype<F> {
    /// One layer's tree and values copied to the host (layer 0 = l0 over the caller's LDE: values are
    /// the caller's own and are not copied).  For tests and for callers that insist on the
    /// reference's fully materialised `FRIProofPrototype`.
    pub fn layer(&self, layer: usize) -> (Vec<Digest>, Vec<F>) {
        let size = unsafe { ffi::hodor_cuda_fri_layer_size(self.handle, layer as u32) } as usize;
        let mut nodes = vec![[0u8; 32]; size];
        let mut values = if layer == 0 { vec![] } else { vec![F::zero(); size] };
        let vp = if layer == 0 { std::ptr::null_mut() } else { ffi::as_u64_mut(&mut values) };
        let rc = unsafe { ffi::hodor_cuda_fri_layer(self.handle, layer as u32, nodes.as_mut_ptr() as *mut u8, vp) };
        assert!(rc == ffi::OK, "hodor_cuda_fri_layer failed: {}", ffi::last_error());
        (nodes, values)
    }

    /// One opening of one layer (`iop.query(idx, values)` of the reference).
    pub fn query(&self, layer: usize, natural_index: usize, size: usize) -> TrivialBlake2sIopQuery<F> {
        let depth = size.trailing_zeros() as usize;
        let mut value = F::zero();
        let mut path = vec![[0u8; 32]; depth];
        let rc = unsafe {
            ffi::hodor_cuda_fri_query(self.handle, layer as u32, natural_index as u64, &mut value as *mut F as *mut u64,
                                      path.as_mut_ptr() as *mut u8)
        };
        assert!(rc == depth as i32, "hodor_cuda_fri_query failed: {}", ffi::last_error());
        TrivialBlake2sIopQuery::from_parts(natural_index, value, path)
    }

    /// `FRIProofPrototype::produce_proof` (src/fri/query_producer.rs:10-53) in ONE library call: for every
    /// committed layer the two members of the coset of the running index (sorted as
    /// `TrivialCombiner::get_coset_for_natural_index` does, src/iop/trivial_coset_combiner.rs:31-43), their
    /// values and authentication paths, gathered on the device and copied out once.
    pub fn produce_proof(self, natural_first_element_index: usize) -> Result<FRIProof<F, CudaBlake2sIOP<F>>, SynthesisError> {
        let layers = self.roots.len();
        let domain_size = self.initial_degree_plus_one * self.lde_factor;
        assert!(natural_first_element_index < domain_size); // query(): assert!(natural_index < self.size())
        let depth0 = domain_size.trailing_zeros() as usize;
        let total: usize = (0..layers).map(|l| 2 * (depth0 - l)).sum();
        let mut indices = vec![0u64; 2 * layers];
        let mut values = vec![F::zero(); 2 * layers];
        let mut paths = vec![[0u8; 32]; total];
        let got = ffi::check(unsafe {
            ffi::hodor_cuda_fri_produce_proof(self.handle, natural_first_element_index as u64, indices.as_mut_ptr(),
                                              ffi::as_u64_mut(&mut values), paths.as_mut_ptr() as *mut u8)
        })?;
        assert!(got as usize == total);
        let mut queries = Vec::with_capacity(2 * layers);
        let mut off = 0;
        for layer in 0..layers {
            let depth = depth0 - layer;
            for q in 0..2 {
                let i = 2 * layer + q;
                queries.push(TrivialBlake2sIopQuery::from_parts(indices[i] as usize, values[i], paths[off..off + depth].to_vec()));
                off += depth;
            }
        }
        Ok(FRIProof::<F, CudaBlake2sIOP<F>> {
            queries,
            roots: self.roots.clone(),
            final_coefficients: self.final_coefficients.clone(),
            initial_degree_plus_one: self.initial_degree_plus_one,
            output_coeffs_at_degree_plus_one: self.output_coeffs_at_degree_plus_one,
            lde_factor: self.lde_factor,
        })
    }
}

pub struct CudaFriIop<F: CudaField> {
    _marker: PhantomData<F>,
}

impl<F: CudaField> CudaFriIop<F> {
    /// The chain on a vector that already lives on the device (e.g. `CommittedOracle::device_values`).
    pub fn proof_from_device_lde(d_lde: *const u64, size: usize, lde_factor: usize, output_coeffs_at_degree_plus_one: usize)
        -> Result<CudaFriPrototype<F>, SynthesisError>
    {
        Self::commit(d_lde, size, lde_factor, output_coeffs_at_degree_plus_one, 1)
    }

    fn commit(lde: *const u64, size: usize, lde_factor: usize, out: usize, on_device: i32) -> Result<CudaFriPrototype<F>, SynthesisError> {
        // the reference's asserts (src/fri/fri_on_values.rs:42-46) are checked by the library: a NULL
        // handle covers them, a domain beyond the field's 2-adicity, and CUDA failures
        ffi::init();
        let h = unsafe { ffi::hodor_cuda_fri_commit(lde, size as u64, lde_factor as u32, out as u32, on_device, F::FIELD_ID) };
        if h.is_null() {
            return Err(SynthesisError::Error);
        }
        let steps = unsafe { ffi::hodor_cuda_fri_num_steps(h) } as usize;
        let mut roots = vec![[0u8; 32]; steps + 1];
        let mut challenges = vec![F::zero(); steps];
        let mut final_coefficients = vec![F::zero(); out];
        let rc = unsafe {
            ffi::hodor_cuda_fri_summary(h, roots.as_mut_ptr() as *mut u8, ffi::as_u64_mut(&mut challenges),
                                        ffi::as_u64_mut(&mut final_coefficients))
        };
        if rc != ffi::OK {
            unsafe { ffi::hodor_cuda_fri_free(h) };
            return Err(SynthesisError::Error);
        }
        Ok(CudaFriPrototype {
            handle: h,
            roots,
            challenges,
            final_coefficients,
            initial_degree_plus_one: size / lde_factor,
            output_coeffs_at_degree_plus_one: out,
            lde_factor,
        })
    }
}

impl<F: CudaField> FriIop<F> for CudaFriIop<F> {
    const DEGREE: usize = 2;

    type IopType = CudaBlake2sIOP<F>;
    type ProofPrototype = CudaFriPrototype<F>;
    type Proof = FRIProof<F, CudaBlake2sIOP<F>>;

    fn proof_from_lde(lde_values: &Polynomial<F, Values>, lde_factor: usize, output_coeffs_at_degree_plus_one: usize, _worker: &Worker)
        -> Result<Self::ProofPrototype, SynthesisError>
    {
        Self::commit(ffi::as_u64(lde_values.as_ref()), lde_values.size(), lde_factor, output_coeffs_at_degree_plus_one, 0)
    }

    fn prototype_into_proof(prototype: Self::ProofPrototype, _iop_values: &Polynomial<F, Values>, natural_first_element_index: usize)
        -> Result<Self::Proof, SynthesisError>
    {
        prototype.produce_proof(natural_first_element_index) // leaf values are read from HBM
    }

    fn verify_proof(proof: &Self::Proof, natural_element_index: usize, expected_value: F) -> Result<bool, SynthesisError> {
        NaiveFriIop::<F, CudaBlake2sIOP<F>>::verify_proof_queries(proof, natural_element_index, Self::DEGREE, expected_value)
    }
}
