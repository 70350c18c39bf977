//! `feature = "cuda"`: the B200 hot path of hodor_b200 behind this crate's own seams.
//!
//!   * transforms   -- a third arm of the `cfg_if!` in src/fft/mod.rs:28-58 (see hodor_cuda.patch) and
//!                     `Polynomial`-level overrides in `poly.rs` (all L cosets of an LDE in one call);
//!   * oracle       -- `CudaBlake2sIOP<F>: IOP<F>` in `iop.rs` (tree built on the GPU, `nodes` in the
//!                     reference's heap layout) and `CommittedOracle<F>` (values + tree stay in HBM);
//!   * FRI          -- `CudaFriIop<F>: FriIop<F>` in `fri.rs` (whole commit chain on the device);
//!   * setup        -- `PrecomputedOmegas` and the ALI inverse divisors of `Prover::new` in `ali.rs`;
//!   * several GPUs -- `Comm`, the four-step NTT and the sharded LDE + FRI chain in `sharded.rs` (one process per GPU).
//!
//! `Prover<F, T, I, P, PR, FRI, A>` (src/prover/mod.rs:29) takes `I` and `FRI` as type parameters, so
//!
//! ```ignore
//! type I = CudaBlake2sIOP<Fr>;
//! type Fri = CudaFriIop<Fr>;
//! Prover::<Fr, Blake2sTranscript<Fr>, I, CudaFriPrototype<Fr>, FRIProof<Fr, I>, Fri, PerRegisterARP>::new(..)
//! ```
//!
//! plugs the GPU path in with no change to the prover.
pub mod ali;
pub mod ffi;
pub mod fri;
pub mod iop;
pub mod poly;
pub mod sharded;

pub use self::ffi::CudaField;
pub use self::fri::{CudaFriIop, CudaFriPrototype};
pub use self::iop::{CommittedOracle, CudaBlake2sIOP, CudaBlake2sIopTree};
pub use self::sharded::{Comm, DeviceVec, ShardedFriCommitment};
