//! Several GPUs, one process per GPU: the north star's "2^24 -> 2^28 coset LDE plus full FRI commit chain
//! on 8 x B200" and the four-step (Bailey) NTT, as the Rust host drives them.
//!
//! The reference's only parallelism is threads of one host (`Worker`, src/fft/multicore.rs); its
//! `parallel_fft` (src/fft/fft.rs:68-125) is the four-step decomposition over CPU threads, and its
//! multi-coset LDE runs the cosets on separate threads (src/polynomials/mod.rs:572-587).  Here a rank is
//! a process that owns one GPU; the exchange steps (NCCL, or NVLink peer stores fused into the last
//! pass of the local transform) happen inside the library.  The caller's only job is the rendezvous:
//! rank 0 makes a 128-byte id and ships it to the other ranks by whatever means the host program has
//! (MPI, a file on a shared disk, a TCP socket); every rank then builds its `Comm`.
//!
//! ```ignore
//! ffi::init();                                              // HODOR_CUDA_DEVICE = local rank
//! let id = if rank == 0 { Comm::unique_id()? } else { recv_from_rank0() };
//! let comm = Comm::init(rank, world, &id)?;
//! let coeffs = DeviceVec::from_slice(poly.as_ref())?;       // replicated on every rank
//! let chain = comm.lde_fri::<Fr>(&coeffs, poly.exp, 16, true, 1)?;
//! // chain.roots / chain.challenges / chain.final_coefficients: identical on every rank and bit-identical
//! // to CudaFriIop::proof_from_lde on the whole LDE
//! ```
use std::marker::PhantomData;
use std::os::raw::{c_int, c_void};

use crate::SynthesisError;

use super::ffi::{self, CudaField};

type Digest = [u8; 32];

/// A vector of field elements in HBM (32 bytes per element, the crate's own in-memory form).
pub struct DeviceVec<F: CudaField> {
    ptr: *mut c_void,
    len: usize,
    _marker: PhantomData<F>,
}

unsafe impl<F: CudaField> Send for DeviceVec<F> {}

impl<F: CudaField> Drop for DeviceVec<F> {
    fn drop(&mut self) {
        unsafe {
            ffi::hodor_cuda_stream_synchronize(std::ptr::null_mut()); // nothing enqueued may still use it
            ffi::hodor_cuda_free(self.ptr)
        }
    }
}

impl<F: CudaField> DeviceVec<F> {
    pub fn with_len(len: usize) -> Result<Self, SynthesisError> {
        ffi::init();
        let ptr = unsafe { ffi::hodor_cuda_malloc(len.max(1) * 32) };
        if ptr.is_null() {
            return Err(SynthesisError::Error);
        }
        Ok(Self { ptr, len, _marker: PhantomData })
    }

    pub fn from_slice(values: &[F]) -> Result<Self, SynthesisError> {
        let v = Self::with_len(values.len())?;
        ffi::check(unsafe { ffi::hodor_cuda_memcpy_h2d(v.ptr, values.as_ptr() as *const c_void, values.len() * 32, std::ptr::null_mut()) })?;
        ffi::check(unsafe { ffi::hodor_cuda_stream_synchronize(std::ptr::null_mut()) })?;
        Ok(v)
    }

    pub fn to_vec(&self) -> Result<Vec<F>, SynthesisError> {
        let mut out = vec![F::zero(); self.len];
        ffi::check(unsafe { ffi::hodor_cuda_memcpy_d2h(out.as_mut_ptr() as *mut c_void, self.ptr, self.len * 32, std::ptr::null_mut()) })?;
        ffi::check(unsafe { ffi::hodor_cuda_stream_synchronize(std::ptr::null_mut()) })?;
        Ok(out)
    }

    pub fn len(&self) -> usize {
        self.len
    }
    pub fn as_ptr(&self) -> *const c_void {
        self.ptr
    }
    pub fn as_mut_ptr(&mut self) -> *mut c_void {
        self.ptr
    }
}

/// Result of the sharded chain: what `FriProofPrototype::get_roots`, `.challenges` and
/// `get_final_coefficients` give for the unsharded chain (src/fri/mod.rs:107-117).
pub struct ShardedFriCommitment<F: CudaField> {
    pub roots: Vec<Digest>,
    pub challenges: Vec<F>,
    pub final_coefficients: Vec<F>,
}

/// This process's place among the ranks.  One per process (the library keeps one communicator).
pub struct Comm {
    pub rank: usize,
    pub world: usize,
}

impl Drop for Comm {
    fn drop(&mut self) {
        unsafe { ffi::hodor_cuda_comm_destroy() }
    }
}

impl Comm {
    /// Rank 0 only: the id every rank passes to `init`.
    pub fn unique_id() -> Result<[u8; 128], SynthesisError> {
        ffi::init();
        let mut id = [0u8; 128];
        ffi::check(unsafe { ffi::hodor_cuda_comm_unique_id(id.as_mut_ptr()) })?;
        Ok(id)
    }

    /// Collective: every rank calls it.  `world` a power of two <= 16; `world == 1` needs no NCCL.
    pub fn init(rank: usize, world: usize, id: &[u8; 128]) -> Result<Self, SynthesisError> {
        ffi::init();
        ffi::check(unsafe { ffi::hodor_cuda_comm_init(rank as c_int, world as c_int, id.as_ptr()) })?;
        Ok(Self { rank, world })
    }

    /// Payload this rank pushed through NCCL / stored straight into peers' buffers since `init`.
    pub fn traffic(&self) -> (u64, u64) {
        let (mut sent, mut stored) = (0u64, 0u64);
        unsafe { ffi::hodor_cuda_comm_info(std::ptr::null_mut(), std::ptr::null_mut(), &mut sent, &mut stored) };
        (sent, stored)
    }

    /// `best_fft(a, worker, omega, log_n, None)` (src/fft/fft.rs:5-19) on a vector dealt over the ranks.
    /// `local`: this rank's cyclic slice a[j * world + rank] (n / world elements).  Returns n / world elements:
    /// the rank-th (n / world^2)-element chunk of every length-(n / world) block of the natural-order result
    /// (`scatter_input` / `gather_output` below state the two layouts in code).  Collective.
    pub fn ntt<F: CudaField>(&self, local: &DeviceVec<F>, log_n: u32, omega: &F) -> Result<DeviceVec<F>, SynthesisError> {
        assert!(local.len() == (1usize << log_n) / self.world);
        let mut out = DeviceVec::<F>::with_len(local.len())?;
        ffi::check(unsafe {
            ffi::hodor_cuda_ntt_sharded(local.as_ptr(), out.as_mut_ptr(), log_n, ffi::elem(omega), F::FIELD_ID, std::ptr::null_mut())
        })?;
        ffi::check(unsafe { ffi::hodor_cuda_stream_synchronize(std::ptr::null_mut()) })?;
        Ok(out)
    }

    /// `poly.coset_lde(worker, factor)` (src/polynomials/mod.rs:349-352) followed by
    /// `NaiveFriIop::proof_from_lde_by_values(lde, factor, out, worker)` (src/fri/fri_on_values.rs:11-159) with
    /// cosets, folds and the bottom of every tree sharded over the ranks.  `coeffs` is replicated.  Collective;
    /// needs world <= factor.
    pub fn lde_fri<F: CudaField>(
        &self,
        coeffs: &DeviceVec<F>,
        log_n: u32,
        factor: usize,
        coset: bool,
        output_coeffs_at_degree_plus_one: usize,
    ) -> Result<ShardedFriCommitment<F>, SynthesisError> {
        assert!(factor.is_power_of_two() && output_coeffs_at_degree_plus_one.is_power_of_two());
        assert!(coeffs.len() == 1usize << log_n);
        let steps = (log_n - output_coeffs_at_degree_plus_one.trailing_zeros()) as usize;
        let mut roots = vec![[0u8; 32]; steps + 1];
        let mut challenges = vec![F::zero(); steps];
        let mut final_coefficients = vec![F::zero(); output_coeffs_at_degree_plus_one];
        let got = ffi::check(unsafe {
            ffi::hodor_cuda_lde_fri_sharded(
                coeffs.as_ptr(),
                log_n,
                factor.trailing_zeros(),
                coset as c_int,
                output_coeffs_at_degree_plus_one as u32,
                roots.as_mut_ptr() as *mut u8,
                ffi::as_u64_mut(&mut challenges),
                ffi::as_u64_mut(&mut final_coefficients),
                F::FIELD_ID,
            )
        })?;
        assert!(got as usize == steps);
        Ok(ShardedFriCommitment { roots, challenges, final_coefficients })
    }
}

/// The input layout of `Comm::ntt`: rank g gets a[j * world + g].
pub fn scatter_input<F: Copy>(a: &[F], world: usize) -> Vec<Vec<F>> {
    (0..world).map(|g| a.iter().skip(g).step_by(world).copied().collect()).collect()
}

/// The output layout of `Comm::ntt`, undone: parts[h][k2 * c + k] = A[k2 * m + h * c + k] with m = n / world,
/// c = m / world.
pub fn gather_output<F: Copy>(parts: &[Vec<F>]) -> Vec<F> {
    let world = parts.len();
    let m = parts[0].len();
    let c = m / world;
    let mut out = Vec::with_capacity(m * world);
    for k2 in 0..world {
        for part in parts.iter() {
            out.extend_from_slice(&part[k2 * c..(k2 + 1) * c]);
        }
    }
    out
}
