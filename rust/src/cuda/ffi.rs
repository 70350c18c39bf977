//! Raw bindings of include/hodor_b200.h (the C ABI of libhodor_b200.so): every entry point the header declares
//! (tests/test_rust_ffi_in_sync.py holds the two files to each other by name, arity and type).
//!
//! A field element of every 25x-bit `#[derive(PrimeField)]` type of this crate is
//! `Fr(FrRepr([u64; 4]))`: 32 contiguous bytes, little-endian limbs, Montgomery form with
//! R = 2^256 -- exactly the bytes `Blake2sLeafEncoder::encode_leaf` hashes
//! (src/iop/blake2s_trivial_iop.rs:36-42).  A `&[F]` therefore crosses the boundary as
//! `*const u64` with no conversion.
use std::os::raw::{c_char, c_int, c_void};

use ff::PrimeField;

use crate::SynthesisError;

pub const FIELD_BLS12_381_FR: c_int = 0; // the modulus src/bn256.rs declares
pub const FIELD_BN254_FR: c_int = 1; // pairing_ce's bn256::Fr (not declared in this crate)
pub const FIELD_STARK252: c_int = 2; // src/experiments/mod.rs:18-21

pub const OK: c_int = 0;
pub const ERR_INVALID_ARG: c_int = -1;
pub const ERR_DOMAIN: c_int = -2;
pub const ERR_CUDA: c_int = -3;
pub const ERR_OOM: c_int = -4;
pub const ERR_NOT_A_ROOT: c_int = -5;
pub const ERR_NOT_INVERTIBLE: c_int = -6;

#[repr(C)]
pub struct FriProto {
    _private: [u8; 0],
}
#[repr(C)]
pub struct Tree {
    _private: [u8; 0],
}

extern "C" {
    pub fn hodor_cuda_device_count() -> c_int;
    pub fn hodor_cuda_init(device: c_int) -> c_int;
    pub fn hodor_cuda_shutdown();
    pub fn hodor_cuda_last_error() -> *const c_char;
    pub fn hodor_cuda_last_error_code() -> c_int;

    // seam 1: src/fft/mod.rs:28-58, :110-123 and the Polynomial methods built on them
    pub fn hodor_cuda_ntt(a: *mut u64, log_n: u32, omega: *const u64, field_id: c_int) -> c_int;
    pub fn hodor_cuda_fft(a: *mut u64, log_n: u32, coset: c_int, field_id: c_int) -> c_int;
    pub fn hodor_cuda_ifft(a: *mut u64, log_n: u32, coset: c_int, field_id: c_int) -> c_int;
    pub fn hodor_cuda_distribute_powers(a: *mut u64, n: u64, g: *const u64, field_id: c_int) -> c_int;
    pub fn hodor_cuda_lde(coeffs: *const u64, log_n: u32, log_factor: u32, coset: c_int, out: *mut u64, field_id: c_int) -> c_int;
    pub fn hodor_cuda_lde_batch(
        coeffs: *const *const u64,
        outs: *const *mut u64,
        count: u32,
        log_n: u32,
        log_factor: u32,
        coset: c_int,
        field_id: c_int,
    ) -> c_int;
    pub fn hodor_cuda_batch_inversion(a: *mut u64, n: u64, field_id: c_int) -> c_int;
    pub fn hodor_cuda_evaluate_at(coeffs: *const u64, n: u64, g: *const u64, out: *mut u64, field_id: c_int) -> c_int;

    // setup of Prover::new: src/precomputations/mod.rs:14-66, src/ali/per_register/mod.rs:60-162, 214-227
    pub fn hodor_cuda_precomputed_omegas(
        omegas: *mut u64,
        coset: *mut u64,
        omegas_inv: *mut u64,
        log_n: u32,
        field_id: c_int,
    ) -> c_int;
    pub fn hodor_cuda_ali_dense_inverse_divisor(
        out: *mut u64,
        log_column: u32,
        log_evaluation: u32,
        start_at: u64,
        span: u64,
        num_rows: u64,
        divisor_degree: *mut u64,
        field_id: c_int,
    ) -> c_int;
    pub fn hodor_cuda_ali_boundary_inverse_divisor(
        out: *mut u64,
        log_column: u32,
        log_evaluation: u32,
        row: u64,
        field_id: c_int,
    ) -> c_int;

    // seam 2: src/iop/mod.rs:58-92
    pub fn hodor_cuda_merkle_build(leaves: *const u64, n: u64, nodes: *mut u8, field_id: c_int) -> c_int;
    pub fn hodor_root_to_challenge(root: *const u8, out: *mut u64, field_id: c_int) -> c_int;
    pub fn hodor_cuda_lde_commit(
        coeffs: *const u64,
        log_n: u32,
        log_factor: u32,
        coset: c_int,
        coeffs_on_device: c_int,
        root: *mut u8,
        field_id: c_int,
    ) -> *mut Tree;
    pub fn hodor_cuda_lde_commit_batch(
        coeffs: *const *const u64,
        count: u32,
        log_n: u32,
        log_factor: u32,
        coset: c_int,
        coeffs_on_device: c_int,
        trees: *mut *mut Tree,
        roots: *mut u8,
        field_id: c_int,
    ) -> c_int;
    pub fn hodor_cuda_tree_commit(values: *const u64, n: u64, values_on_device: c_int, root: *mut u8, field_id: c_int) -> *mut Tree;
    pub fn hodor_cuda_tree_free(t: *mut Tree);
    pub fn hodor_cuda_tree_size(t: *const Tree) -> u64;
    pub fn hodor_cuda_tree_values(t: *const Tree) -> *const c_void;
    pub fn hodor_cuda_tree_root(t: *const Tree, root: *mut u8, challenge: *mut u64) -> c_int;
    pub fn hodor_cuda_tree_query(t: *const Tree, natural_index: u64, value: *mut u64, path: *mut u8) -> c_int;
    pub fn hodor_cuda_tree_read(t: *const Tree, first: u64, count: u64, values: *mut u64, nodes: *mut u8) -> c_int;

    // seam 3: src/fri/mod.rs:36-61
    pub fn hodor_cuda_fri_commit(
        lde: *const u64,
        n: u64,
        lde_factor: u32,
        out_coeffs: u32,
        lde_on_device: c_int,
        field_id: c_int,
    ) -> *mut FriProto;
    pub fn hodor_cuda_fri_free(p: *mut FriProto);
    pub fn hodor_cuda_fri_num_steps(p: *const FriProto) -> c_int;
    pub fn hodor_cuda_fri_summary(p: *const FriProto, roots: *mut u8, challenges: *mut u64, final_coeffs: *mut u64) -> c_int;
    pub fn hodor_cuda_fri_layer(p: *const FriProto, layer: u32, nodes: *mut u8, values: *mut u64) -> c_int;
    pub fn hodor_cuda_fri_layer_size(p: *const FriProto, layer: u32) -> u64;
    pub fn hodor_cuda_fri_query(p: *const FriProto, layer: u32, natural_index: u64, value: *mut u64, path: *mut u8) -> c_int;

    // ---- context housekeeping and diagnostics
    pub fn hodor_cuda_trim() -> c_int;
    pub fn hodor_cuda_workspace_bytes() -> usize;
    pub fn hodor_cuda_launch_count() -> u64;
    pub fn hodor_cuda_selftest_mul_pre(field_id: c_int) -> c_int;
    pub fn hodor_cuda_profile_begin() -> c_int;
    pub fn hodor_cuda_profile_end(json_out: *mut c_char, cap: usize) -> c_int;

    // ---- host scalar helpers.  A Rust caller has ff_ce for these; they are bound so that `debug_check` can compare the
    // library's view of a field with `F`'s, and for hosts without the crate's field types.
    pub fn hodor_field_constants(field_id: c_int, modulus: *mut u64, one: *mut u64, generator: *mut u64, root_of_unity: *mut u64, s: *mut u32, num_bits: *mut u32, capacity: *mut u32) -> c_int;
    pub fn hodor_domain_generator(field_id: c_int, log_n: u32, out: *mut u64) -> c_int;
    pub fn hodor_field_mul(field_id: c_int, a: *const u64, b: *const u64, out: *mut u64) -> c_int;
    pub fn hodor_field_add(field_id: c_int, a: *const u64, b: *const u64, out: *mut u64) -> c_int;
    pub fn hodor_field_sub(field_id: c_int, a: *const u64, b: *const u64, out: *mut u64) -> c_int;
    pub fn hodor_field_pow(field_id: c_int, a: *const u64, e: u64, out: *mut u64) -> c_int;
    pub fn hodor_field_inverse(field_id: c_int, a: *const u64, out: *mut u64) -> c_int;
    pub fn hodor_field_from_repr(field_id: c_int, plain: *const u64, out: *mut u64) -> c_int;
    pub fn hodor_field_into_repr(field_id: c_int, mont: *const u64, out: *mut u64) -> c_int;
    pub fn hodor_hash_leaf(leaf: *const u64, out: *mut u8) -> c_int;
    pub fn hodor_hash_node(left: *const u8, right: *const u8, out: *mut u8) -> c_int;

    // ---- device and pinned memory, copies (DeviceVec in sharded.rs; pinned buffers make the batch entry points overlap)
    pub fn hodor_cuda_malloc(bytes: usize) -> *mut c_void;
    pub fn hodor_cuda_free(dptr: *mut c_void);
    pub fn hodor_cuda_host_alloc(bytes: usize) -> *mut c_void;
    pub fn hodor_cuda_host_free(hptr: *mut c_void);
    pub fn hodor_cuda_memcpy_h2d(dptr: *mut c_void, hptr: *const c_void, bytes: usize, stream: *mut c_void) -> c_int;
    pub fn hodor_cuda_memcpy_d2h(hptr: *mut c_void, dptr: *const c_void, bytes: usize, stream: *mut c_void) -> c_int;
    pub fn hodor_cuda_stream_synchronize(stream: *mut c_void) -> c_int;

    // ---- elementwise Polynomial methods (src/polynomials/mod.rs:59-83, 640-683, 744-771, 817-887)
    pub fn hodor_cuda_elementwise(op: c_int, a: *const u64, b: *const u64, out: *mut u64, n: u64, field_id: c_int) -> c_int;
    pub fn hodor_cuda_poly_op(op: c_int, a: *const u64, b: *const u64, scalar: *const u64, exp: u64, out: *mut u64, n: u64, field_id: c_int) -> c_int;

    // ---- the rest of seams 2 and 3
    pub fn hodor_cuda_tree_nodes(t: *const Tree) -> *const c_void;
    pub fn hodor_cuda_tree_query_batch(t: *const Tree, natural_indices: *const u64, count: u32, values: *mut u64, paths: *mut u8) -> c_int;
    pub fn hodor_cuda_fri_produce_proof(p: *const FriProto, natural_first_element_index: u64, indices: *mut u64, values: *mut u64, paths: *mut u8) -> c_int;
    pub fn hodor_cuda_fri_commit_host(lde: *const u64, n: u64, lde_factor: u32, out_coeffs: u32, l0_nodes: *mut u8, layer_nodes: *mut *mut u8, layer_values: *mut *mut u64, challenges: *mut u64, final_root: *mut u8, final_coeffs: *mut u64, field_id: c_int) -> c_int;

    // ---- device-resident variants: device pointers + a cudaStream_t, enqueue and return
    pub fn hodor_cuda_ntt_dev(d_in: *const c_void, d_out: *mut c_void, log_n: u32, omega: *const u64, field_id: c_int, stream: *mut c_void) -> c_int;
    pub fn hodor_cuda_fft_dev(d_in: *const c_void, d_out: *mut c_void, log_n: u32, coset: c_int, field_id: c_int, stream: *mut c_void) -> c_int;
    pub fn hodor_cuda_ifft_dev(d_in: *const c_void, d_out: *mut c_void, log_n: u32, coset: c_int, field_id: c_int, stream: *mut c_void) -> c_int;
    pub fn hodor_cuda_lde_dev(d_coeffs: *const c_void, log_n: u32, log_factor: u32, coset: c_int, d_out: *mut c_void, field_id: c_int, stream: *mut c_void) -> c_int;
    pub fn hodor_cuda_distribute_powers_dev(d_a: *mut c_void, n: u64, g: *const u64, field_id: c_int, stream: *mut c_void) -> c_int;
    pub fn hodor_cuda_elementwise_dev(op: c_int, d_a: *const c_void, d_b: *const c_void, d_out: *mut c_void, n: u64, field_id: c_int, stream: *mut c_void) -> c_int;
    pub fn hodor_cuda_poly_op_dev(op: c_int, d_a: *const c_void, d_b: *const c_void, scalar: *const u64, exp: u64, d_out: *mut c_void, n: u64, field_id: c_int, stream: *mut c_void) -> c_int;
    pub fn hodor_cuda_batch_inversion_dev(d_a: *mut c_void, n: u64, d_status: *mut c_int, field_id: c_int, stream: *mut c_void) -> c_int;
    pub fn hodor_cuda_evaluate_at_dev(d_coeffs: *const c_void, n: u64, g: *const u64, d_out: *mut c_void, field_id: c_int, stream: *mut c_void) -> c_int;
    pub fn hodor_cuda_merkle_build_dev(d_leaves: *const c_void, n: u64, d_nodes: *mut c_void, d_root: *mut c_void, d_challenge: *mut c_void, field_id: c_int, stream: *mut c_void) -> c_int;
    pub fn hodor_cuda_merkle_top_dev(d_nodes: *mut c_void, w: u64, d_root: *mut c_void, d_challenge: *mut c_void, field_id: c_int, stream: *mut c_void) -> c_int;
    pub fn hodor_cuda_fri_fold_dev(d_in: *const c_void, n: u64, initial_domain_size: u64, layer: u32, d_challenge: *const c_void, d_out: *mut c_void, field_id: c_int, stream: *mut c_void) -> c_int;
    pub fn hodor_cuda_precomputed_omegas_dev(d_omegas: *mut c_void, d_coset: *mut c_void, d_omegas_inv: *mut c_void, log_n: u32, field_id: c_int, stream: *mut c_void) -> c_int;
    pub fn hodor_cuda_ali_dense_inverse_divisor_dev(d_out: *mut c_void, log_column: u32, log_evaluation: u32, start_at: u64, span: u64, num_rows: u64, divisor_degree: *mut u64, field_id: c_int, stream: *mut c_void) -> c_int;
    pub fn hodor_cuda_ali_boundary_inverse_divisor_dev(d_out: *mut c_void, log_column: u32, log_evaluation: u32, row: u64, field_id: c_int, stream: *mut c_void) -> c_int;

    // ---- several GPUs, one process per GPU (sharded.rs): the four-step NTT and the sharded LDE + FRI chain,
    // and the single-GPU building blocks they are made of
    pub fn hodor_cuda_comm_unique_id(id: *mut u8) -> c_int;
    pub fn hodor_cuda_comm_init(rank: c_int, world: c_int, id: *const u8) -> c_int;
    pub fn hodor_cuda_comm_destroy();
    pub fn hodor_cuda_comm_info(rank: *mut c_int, world: *mut c_int, bytes_sent: *mut u64, bytes_peer_stored: *mut u64) -> c_int;
    pub fn hodor_cuda_ntt_sharded(d_local: *const c_void, d_out: *mut c_void, log_n: u32, omega: *const u64, field_id: c_int, stream: *mut c_void) -> c_int;
    pub fn hodor_cuda_lde_fri_sharded(d_coeffs: *const c_void, log_n: u32, log_factor: u32, coset: c_int, out_coeffs: u32, roots: *mut u8, challenges: *mut u64, final_coeffs: *mut u64, field_id: c_int) -> c_int;
    pub fn hodor_cuda_lde_cosets_dev(d_coeffs: *const c_void, log_n: u32, log_factor: u32, coset: c_int, first_coset: u32, coset_stride: u32, log_count: u32, d_out: *mut c_void, field_id: c_int, stream: *mut c_void) -> c_int;
    pub fn hodor_cuda_fri_fold_shard_dev(d_in: *const c_void, n_local: u64, initial_domain_size: u64, layer: u32, log_g: u32, rank: u32, d_challenge: *const c_void, d_out: *mut c_void, field_id: c_int, stream: *mut c_void) -> c_int;
    pub fn hodor_cuda_ntt_shard_cols_dev(d_in: *const c_void, d_out: *mut c_void, log_n: u32, log_g: u32, rank: u32, omega: *const u64, field_id: c_int, stream: *mut c_void) -> c_int;
    pub fn hodor_cuda_ntt_shard_rows_dev(d_in: *const c_void, d_out: *mut c_void, log_n: u32, log_g: u32, rank: u32, omega: *const u64, field_id: c_int, stream: *mut c_void) -> c_int;
    pub fn hodor_cuda_merkle_build_shard_dev(d_chunks: *const c_void, n: u64, log_g: u32, d_nodes: *mut c_void, d_root: *mut c_void, d_challenge: *mut c_void, field_id: c_int, stream: *mut c_void) -> c_int;
}

/// Maps a `PrimeField` type of this crate to the library's field id.  The library compiles the
/// modulus in; `debug_check` (run once per type by `init`) compares `F::char()`, the generator and
/// the root of unity with the library's view so that a mismatch cannot go unnoticed.
pub trait CudaField: PrimeField {
    const FIELD_ID: c_int;
}
impl CudaField for crate::bn256::Fr {
    const FIELD_ID: c_int = FIELD_BLS12_381_FR;
}
impl CudaField for crate::experiments::Fr {
    const FIELD_ID: c_int = FIELD_STARK252;
}

/// The library's field id for `F`, by modulus (`F::char()`), or None for a field it does not carry.
/// Used by the `cfg_if!` arm of src/fft/mod.rs, whose callers are generic over every `PrimeField`.
pub fn field_id_of<F: PrimeField>() -> Option<c_int> {
    const MODULI: [(c_int, [u64; 4]); 3] = [
        (FIELD_BLS12_381_FR, [0xffffffff00000001, 0x53bda402fffe5bfe, 0x3339d80809a1d805, 0x73eda753299d7d48]),
        (FIELD_BN254_FR, [0x43e1f593f0000001, 0x2833e84879b97091, 0xb85045b68181585d, 0x30644e72e131a029]),
        (FIELD_STARK252, [0x0000000000000001, 0x0000000000000000, 0x0000000000000000, 0x0800000000000011]),
    ];
    if std::mem::size_of::<F>() != 32 {
        return None;
    }
    let p = F::char();
    let limbs: &[u64] = p.as_ref();
    MODULI.iter().find(|(_, m)| limbs == &m[..]).map(|(id, _)| *id)
}

/// `hodor_cuda_init` on first use (device from HODOR_CUDA_DEVICE, default 0).  The reference has no
/// such step; panicking here mirrors its `expect(..)` style for unrecoverable set-up failures.
pub fn init() {
    use std::sync::Once;
    static START: Once = Once::new();
    START.call_once(|| {
        let dev = std::env::var("HODOR_CUDA_DEVICE").ok().and_then(|s| s.parse().ok()).unwrap_or(0);
        let rc = unsafe { hodor_cuda_init(dev) };
        assert!(rc == OK, "hodor_cuda_init({}) failed: {}", dev, last_error());
    });
}

pub fn last_error() -> String {
    unsafe {
        let p = hodor_cuda_last_error();
        if p.is_null() {
            String::new()
        } else {
            std::ffi::CStr::from_ptr(p).to_string_lossy().into_owned()
        }
    }
}

/// Non-negative return codes pass through; every failure becomes `SynthesisError::Error`
/// (src/lib.rs:40-46), which is what `Domain::new_for_size` and `batch_inversion` return in the
/// reference for the two recoverable conditions (ERR_DOMAIN, ERR_NOT_INVERTIBLE).
pub fn check(rc: c_int) -> Result<c_int, SynthesisError> {
    if rc >= 0 {
        Ok(rc)
    } else {
        Err(SynthesisError::Error)
    }
}

#[inline]
pub fn as_u64<F: CudaField>(s: &[F]) -> *const u64 {
    debug_assert!(std::mem::size_of::<F>() == 32);
    s.as_ptr() as *const u64
}
#[inline]
pub fn as_u64_mut<F: CudaField>(s: &mut [F]) -> *mut u64 {
    debug_assert!(std::mem::size_of::<F>() == 32);
    s.as_mut_ptr() as *mut u64
}
#[inline]
pub fn elem<F: CudaField>(x: &F) -> *const u64 {
    x as *const F as *const u64
}
