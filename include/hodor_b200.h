/*
 * hodor_b200 -- C ABI of the B200 (sm_100a) STARK-prover hot path.
 *
 * Drop-in boundary for matter-labs/hodor (@76fc894).  The reference has no FFI; these entry points
 * are what a `feature = "cuda"` arm of its `cfg_if!` dispatch (src/fft/mod.rs:28-58) and new
 * `IOP` / `FriIop` type arguments of `Prover` (src/prover/mod.rs:29,44) would bind.  Each function
 * cites the reference item it replaces.  INTEGRATION.md shows the Rust side.
 *
 * Conventions
 *   - A field element is 4 little-endian u64 limbs in Montgomery form (R = 2^256), canonical (< p):
 *     the in-memory layout of ff_ce's derive(PrimeField) `Fr(FrRepr([u64; 4]))`, so `&[F]` can be
 *     passed as `*const u64` with no conversion.  Vectors are contiguous, 32 bytes per element.
 *   - Digests are 32 bytes.  A tree's `nodes` is n digests in heap order (nodes[0] zero,
 *     nodes[1] root, level with w nodes at [w, 2w)), exactly `Blake2sIopTree.nodes`.
 *   - Sizes are element counts.  All transforms are natural order in, natural order out.
 *   - Return value: 0 (or a non-negative count) on success, a negative HODOR_ERR_* otherwise;
 *     hodor_cuda_last_error() gives the message.  The library never aborts the process and never
 *     computes on the CPU: without a usable GPU every compute entry point fails with
 *     HODOR_ERR_CUDA.
 *   - One process drives one GPU (hodor_cuda_init(device)).  Entry points may be called from any
 *     host thread; calls are serialised on the context.  `_dev` variants take device pointers and a
 *     cudaStream_t (as void*), enqueue work and return without synchronising.  Device element and digest
 *     arrays must be 32-byte aligned (they are read with 256-bit loads): HODOR_ERR_INVALID_ARG otherwise.
 */
#ifndef HODOR_B200_H
#define HODOR_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* field_id.  HODOR_FIELD_BLS12_381_FR is the field that src/bn256.rs:4-7 declares under the name
 * bn256::Fr (its modulus is the BLS12-381 scalar field); HODOR_FIELD_BN254_FR is the curve usually
 * meant by "bn256"; HODOR_FIELD_STARK252 is experiments::Fr (src/experiments/mod.rs:18-21). */
#define HODOR_FIELD_BLS12_381_FR 0
#define HODOR_FIELD_BN254_FR 1
#define HODOR_FIELD_STARK252 2

#define HODOR_OK 0
#define HODOR_ERR_INVALID_ARG (-1)  /* reference: assert!/expect panic */
#define HODOR_ERR_DOMAIN (-2)       /* reference: Domain::new_for_size -> Err(SynthesisError::Error), src/domains/mod.rs:29-32 */
#define HODOR_ERR_CUDA (-3)         /* no device, launch or runtime failure */
#define HODOR_ERR_OOM (-4)
#define HODOR_ERR_NOT_A_ROOT (-5)   /* omega is not a primitive 2^log_n-th root of unity */
#define HODOR_ERR_NOT_INVERTIBLE (-6) /* batch_inversion met a zero: reference returns Err(SynthesisError::Error), src/polynomials/mod.rs:919 */

/* ---- context ------------------------------------------------------------------------------ */
int hodor_cuda_device_count(void);
int hodor_cuda_init(int device);
void hodor_cuda_shutdown(void);
const char* hodor_cuda_last_error(void);
/* the HODOR_ERR_* code that goes with hodor_cuda_last_error() (for entry points that return a handle or NULL) */
int hodor_cuda_last_error_code(void);
/* Handles (FRI prototypes, committed oracles) and staging buffers are carved from a cache of device blocks that
 * is kept across calls ($HODOR_POOL_CACHE_MB, default 96 GiB, oldest blocks released first); this hands every
 * cached block that is not in use back to the driver. */
int hodor_cuda_trim(void);
/* bytes of device workspace + cached twiddle tables currently held */
size_t hodor_cuda_workspace_bytes(void);

/* ---- host-side scalar helpers (what ff_ce / Domain give the Rust caller) --------------------- */
/* PrimeField consts: modulus (plain), one = R mod p, multiplicative_generator(), root_of_unity()
 * (all Montgomery), S, NUM_BITS, CAPACITY. */
int hodor_field_constants(int field_id, uint64_t modulus[4], uint64_t one[4], uint64_t generator[4],
                          uint64_t root_of_unity[4], uint32_t* s, uint32_t* num_bits, uint32_t* capacity);
/* Domain::new_for_size(2^log_n).generator  (src/domains/mod.rs:21-44) */
int hodor_domain_generator(int field_id, uint32_t log_n, uint64_t out[4]);
int hodor_field_mul(int field_id, const uint64_t a[4], const uint64_t b[4], uint64_t out[4]);
int hodor_field_add(int field_id, const uint64_t a[4], const uint64_t b[4], uint64_t out[4]);
int hodor_field_sub(int field_id, const uint64_t a[4], const uint64_t b[4], uint64_t out[4]);
int hodor_field_pow(int field_id, const uint64_t a[4], uint64_t e, uint64_t out[4]);
int hodor_field_inverse(int field_id, const uint64_t a[4], uint64_t out[4]); /* HODOR_ERR_INVALID_ARG on zero */
int hodor_field_from_repr(int field_id, const uint64_t plain[4], uint64_t out[4]); /* PrimeField::from_repr */
int hodor_field_into_repr(int field_id, const uint64_t mont[4], uint64_t out[4]);  /* PrimeField::into_repr */
/* Blake2sLeafEncoder::interpret_hash / IopTree::encode_root_into_challenge
 * (src/iop/blake2s_trivial_iop.rs:48-60, :226-228) */
int hodor_root_to_challenge(const uint8_t root[32], uint64_t out[4], int field_id);
/* IopTreeHasher::hash_leaf / hash_node for single items (src/iop/blake2s_trivial_iop.rs:81-104):
 * verifier-side O(log n) work (IopTree::verify, the leaf-pair hash of get_path) on the host. */
int hodor_hash_leaf(const uint64_t leaf[4], uint8_t out[32]);
int hodor_hash_node(const uint8_t left[32], const uint8_t right[32], uint8_t out[32]);

/* ---- memory helpers ---------------------------------------------------------------------- */
void* hodor_cuda_malloc(size_t bytes);
void hodor_cuda_free(void* dptr);
void* hodor_cuda_host_alloc(size_t bytes); /* pinned */
void hodor_cuda_host_free(void* hptr);
int hodor_cuda_memcpy_h2d(void* dptr, const void* hptr, size_t bytes, void* stream);
int hodor_cuda_memcpy_d2h(void* hptr, const void* dptr, size_t bytes, void* stream);
int hodor_cuda_stream_synchronize(void* stream);

/* ---- transforms, host memory ------------------------------------------------------------- */
/* best_fft / serial_fft (src/fft/fft.rs:5-66, dispatch src/fft/mod.rs:50-56):
 * a[k] <- sum_j a[j] * omega^(j*k), in place.  omega must be a primitive 2^log_n-th root. */
int hodor_cuda_ntt(uint64_t* a, uint32_t log_n, const uint64_t omega[4], int field_id);
/* Polynomial::fft / coset_fft (src/polynomials/mod.rs:611-631): omega = domain generator;
 * coset != 0 first scales a[j] by multiplicative_generator^j. */
int hodor_cuda_fft(uint64_t* a, uint32_t log_n, int coset, int field_id);
/* Polynomial::ifft / icoset_fft (src/polynomials/mod.rs:773-807): NTT with omega^-1, times n^-1;
 * coset != 0 then scales a[j] by generator^-j. */
int hodor_cuda_ifft(uint64_t* a, uint32_t log_n, int coset, int field_id);
/* distribute_powers (src/fft/mod.rs:110-123): a[j] <- a[j] * g^j */
int hodor_cuda_distribute_powers(uint64_t* a, uint64_t n, const uint64_t g[4], int field_id);
/* Polynomial::lde / coset_lde == (coset_)lde_using_multiple_cosets
 * (src/polynomials/mod.rs:343-352, 418-482, 544-609): out has n << log_factor elements,
 * out[i + L*k] = P(shift * w_{nL}^(i + L*k)), shift = 1 or multiplicative_generator. */
int hodor_cuda_lde(const uint64_t* coeffs, uint32_t log_n, uint32_t log_factor, int coset, uint64_t* out, int field_id);
/* The same LDE for `count` polynomials of one shape (the prover lifts every register:
 * `for w in witness { w.lde(&worker, lde_factor) }`, src/prover/mod.rs:73-76), pipelined so that the
 * host<->device copies of neighbouring polynomials overlap the transform of the current one.
 * coeffs[i]: 2^log_n elements, outs[i]: 2^(log_n+log_factor) elements; pinned host memory
 * (hodor_cuda_host_alloc) is needed for the overlap, pageable memory still gives correct results. */
int hodor_cuda_lde_batch(const uint64_t* const* coeffs, uint64_t* const* outs, uint32_t count, uint32_t log_n,
                         uint32_t log_factor, int coset, int field_id);
/* elementwise Polynomial ops (src/polynomials/mod.rs:640-683, 817-887): op 0 mul, 1 add, 2 sub,
 * 3 scale (b is one element).  out may alias a. */
int hodor_cuda_elementwise(int op, const uint64_t* a, const uint64_t* b, uint64_t* out, uint64_t n, int field_id);

/* The remaining elementwise Polynomial methods.  op: 0..3 as above, 4 add_assign_scaled (a + b * scalar,
 * :654-669, :843-858), 5 add_constant (a + scalar, :831-841), 6 negate (:72-83), 7 square (:760-771),
 * 8 pow (a^exp, :744-758).  b / scalar may be NULL for the ops that do not use them.  out may alias a. */
#define HODOR_OP_MUL 0
#define HODOR_OP_ADD 1
#define HODOR_OP_SUB 2
#define HODOR_OP_SCALE 3
#define HODOR_OP_ADD_SCALED 4
#define HODOR_OP_ADD_CONST 5
#define HODOR_OP_NEGATE 6
#define HODOR_OP_SQUARE 7
#define HODOR_OP_POW 8
int hodor_cuda_poly_op(int op, const uint64_t* a, const uint64_t* b, const uint64_t scalar[4], uint64_t exp, uint64_t* out,
                       uint64_t n, int field_id);
/* Polynomial<F, Values>::batch_inversion (src/polynomials/mod.rs:889-954): a[i] <- a[i]^-1 in place.
 * A zero element fails with HODOR_ERR_NOT_INVERTIBLE and leaves `a` untouched, like the reference,
 * which returns Err(SynthesisError::Error) before writing anything. */
int hodor_cuda_batch_inversion(uint64_t* a, uint64_t n, int field_id);
/* Polynomial<F, Coefficients>::evaluate_at (src/polynomials/mod.rs:685-711): out = sum_j coeffs[j] * g^j */
int hodor_cuda_evaluate_at(const uint64_t* coeffs, uint64_t n, const uint64_t g[4], uint64_t out[4], int field_id);

/* ---- Merkle oracle, host memory ----------------------------------------------------------- */
/* Blake2sIopTree::create (src/iop/blake2s_trivial_iop.rs:131-219).  n a power of two >= 2;
 * nodes receives n * 32 bytes. */
int hodor_cuda_merkle_build(const uint64_t* leaves, uint64_t n, uint8_t* nodes, int field_id);

/* ---- committed oracles: values + Merkle tree resident in HBM ------------------------------------ */
/* What Prover::prove does with every register and with g (src/prover/mod.rs:73-80, :91-95):
 *     let lde = w.lde(&worker, lde_factor)?;  let oracle = I::create(lde.as_ref());
 * in one call: coefficients in (host or device), the LDE and the tree stay on the device behind the handle,
 * only the root (32 B) comes back.  The handle serves IOP::query (src/iop/blake2s_trivial_iop.rs:251-279,
 * :324-338) without ever moving the 2n * 32 bytes of values and nodes over PCIe. */
typedef struct hodor_tree hodor_tree;
hodor_tree* hodor_cuda_lde_commit(const uint64_t* coeffs, uint32_t log_n, uint32_t log_factor, int coset,
                                  int coeffs_on_device, uint8_t root[32], int field_id);
/* The register loop of the prover: `count` polynomials of one shape; the host->device copy of polynomial
 * i+1 overlaps the LDE + tree build of polynomial i (pinned host memory needed for the overlap).
 * trees[i] receives the handles, roots (may be NULL) count * 32 bytes. */
int hodor_cuda_lde_commit_batch(const uint64_t* const* coeffs, uint32_t count, uint32_t log_n, uint32_t log_factor,
                                int coset, int coeffs_on_device, hodor_tree** trees, uint8_t* roots, int field_id);
/* IopTree::create / IOP::create (src/iop/blake2s_trivial_iop.rs:131-219, :291-297) on n values.  Host values
 * are copied in and owned by the handle; device values are borrowed (keep them alive until _free). */
hodor_tree* hodor_cuda_tree_commit(const uint64_t* values, uint64_t n, int values_on_device, uint8_t root[32],
                                   int field_id);
void hodor_cuda_tree_free(hodor_tree* t);
uint64_t hodor_cuda_tree_size(const hodor_tree* t);
/* device pointers of the committed values (n * 32 B, natural order) and of `nodes` (n * 32 B, heap order):
 * inputs for hodor_cuda_fri_commit(lde_on_device = 1) and the other `_dev` entry points */
const void* hodor_cuda_tree_values(const hodor_tree* t);
const void* hodor_cuda_tree_nodes(const hodor_tree* t);
/* get_root (:221-224) and get_challenge_scalar_from_root (:230-234); either pointer may be NULL */
int hodor_cuda_tree_root(const hodor_tree* t, uint8_t root[32], uint64_t challenge[4]);
/* IOP::query: value (4 u64) and path (log2(n) * 32 B, leaf-pair hash first).  Returns the path length. */
int hodor_cuda_tree_query(const hodor_tree* t, uint64_t natural_index, uint64_t value[4], uint8_t* path);
int hodor_cuda_tree_query_batch(const hodor_tree* t, const uint64_t* natural_indices, uint32_t count, uint64_t* values,
                                uint8_t* paths);
/* copies values[first, first+count) and / or nodes[first, first+count) to the host (either may be NULL) */
int hodor_cuda_tree_read(const hodor_tree* t, uint64_t first, uint64_t count, uint64_t* values, uint8_t* nodes);

/* ---- FRI commit chain ---------------------------------------------------------------------- */
/* NaiveFriIop::proof_from_lde_by_values (src/fri/fri_on_values.rs:11-159), result kept on the
 * device behind a handle (the FRIProofPrototype, src/fri/mod.rs:107-117).  lde may be a host or a
 * device pointer (lde_on_device).  Returns the handle or NULL. */
typedef struct hodor_fri_proto hodor_fri_proto;
hodor_fri_proto* hodor_cuda_fri_commit(const uint64_t* lde, uint64_t n, uint32_t lde_factor, uint32_t out_coeffs,
                                       int lde_on_device, int field_id);
void hodor_cuda_fri_free(hodor_fri_proto* p);
int hodor_cuda_fri_num_steps(const hodor_fri_proto* p);
/* roots: (num_steps + 1) * 32 bytes: l0 root, then every intermediate root (FriProofPrototype::
 * get_roots; the last one is also final_root).  challenges: num_steps * 4 u64.
 * final_coeffs: out_coeffs * 4 u64. */
int hodor_cuda_fri_summary(const hodor_fri_proto* p, uint8_t* roots, uint64_t* challenges, uint64_t* final_coeffs);
/* layer 0 = the l0 commitment over the caller's lde values; layer i >= 1 = intermediate i-1.
 * Copies the layer's nodes (size * 32 B) and, for i >= 1, values (size * 4 u64) to the host.
 * Either pointer may be NULL. */
int hodor_cuda_fri_layer(const hodor_fri_proto* p, uint32_t layer, uint8_t* nodes, uint64_t* values);
uint64_t hodor_cuda_fri_layer_size(const hodor_fri_proto* p, uint32_t layer);
/* IOP::query (src/iop/blake2s_trivial_iop.rs:251-279, 324-338) against layer `layer`:
 * value (4 u64) and path (log2(size) * 32 B, leaf-pair hash first).  Returns the path length. */
int hodor_cuda_fri_query(const hodor_fri_proto* p, uint32_t layer, uint64_t natural_index, uint64_t value[4],
                         uint8_t* path);
/* FRIProofPrototype::produce_proof / FriIop::prototype_into_proof (src/fri/query_producer.rs:10-53) in one call:
 * for every committed layer the two members of the coset of the running index (sorted), their values and paths.
 * indices: 2 * (steps + 1) u64; values: 2 * (steps + 1) * 4 u64; paths: for layer l (size n >> l) two paths of
 * log2(n >> l) digests, concatenated in layer order (steps + 1 layers).  Returns the number of digests written. */
int hodor_cuda_fri_produce_proof(const hodor_fri_proto* p, uint64_t natural_first_element_index, uint64_t* indices,
                                 uint64_t* values, uint8_t* paths);
/* Same as the reference signature: everything copied out to caller-allocated host buffers.
 * layer_nodes[i] / layer_values[i] hold (n >> (i+1)) entries.  Returns num_steps. */
int hodor_cuda_fri_commit_host(const uint64_t* lde, uint64_t n, uint32_t lde_factor, uint32_t out_coeffs,
                               uint8_t* l0_nodes, uint8_t** layer_nodes, uint64_t** layer_values,
                               uint64_t* challenges, uint8_t* final_root, uint64_t* final_coeffs, int field_id);

/* ---- device-resident variants (no host copies, stream ordered) ------------------------------ */
int hodor_cuda_ntt_dev(const void* d_in, void* d_out, uint32_t log_n, const uint64_t omega[4], int field_id,
                       void* stream);
int hodor_cuda_fft_dev(const void* d_in, void* d_out, uint32_t log_n, int coset, int field_id, void* stream);
int hodor_cuda_ifft_dev(const void* d_in, void* d_out, uint32_t log_n, int coset, int field_id, void* stream);
int hodor_cuda_lde_dev(const void* d_coeffs, uint32_t log_n, uint32_t log_factor, int coset, void* d_out,
                       int field_id, void* stream);
/* d_root (32 B) and d_challenge (32 B, Montgomery) may be NULL */
int hodor_cuda_merkle_build_dev(const void* d_leaves, uint64_t n, void* d_nodes, void* d_root, void* d_challenge,
                                int field_id, void* stream);
/* Top of a tree whose level of w nodes is already known (e.g. the all-gathered sub-roots of a tree built
 * as w subtrees on w GPUs): d_nodes holds 2w digests in heap order with [w, 2w) filled in; fills in
 * [1, w), and writes the root (32 B) and the challenge (32 B, Montgomery) if the pointers are non-NULL. */
int hodor_cuda_merkle_top_dev(void* d_nodes, uint64_t w, void* d_root, void* d_challenge, int field_id, void* stream);
/* one FRI layer: d_out[idx], idx < n/2, from d_in (n values); challenge read from d_challenge */
int hodor_cuda_fri_fold_dev(const void* d_in, uint64_t n, uint64_t initial_domain_size, uint32_t layer,
                            const void* d_challenge, void* d_out, int field_id, void* stream);
/* PrecomputedOmegas::new_for_domain (src/precomputations/mod.rs:14-66) as device vectors: d_omegas[i] = omega^i and
 * d_coset[i] = g * omega^i (n = 2^log_n elements each), d_omegas_inv[i] = omega^-i (n/2 elements); omega the generator of
 * the size-n domain, g the field's multiplicative generator.  Any pointer may be NULL to skip that vector. */
int hodor_cuda_precomputed_omegas_dev(void* d_omegas, void* d_coset, void* d_omegas_inv, uint32_t log_n, int field_id,
                                      void* stream);
/* inverse_divisor_for_dense_constraint_in_coset (src/ali/per_register/mod.rs:60-162): for x_j = g * w_E^j over the
 * evaluation domain of size E = 2^log_evaluation, d_out[j] = prod_root (x_j - root) / (x_j^T - 1), T = 2^log_column the
 * column domain, roots = w_T^k for k in [0, start_at) and [num_rows - span, T).  *divisor_degree (may be NULL) gets
 * T - start_at - (T - num_rows) - span (:71-75). */
int hodor_cuda_ali_dense_inverse_divisor_dev(void* d_out, uint32_t log_column, uint32_t log_evaluation, uint64_t start_at,
                                             uint64_t span, uint64_t num_rows, uint64_t* divisor_degree, int field_id,
                                             void* stream);
/* boundary-constraint divisor (src/ali/per_register/mod.rs:214-227): d_out[j] = 1 / (x_j - w_T^row) on the same coset;
 * returns after the batch inversion has completed (HODOR_ERR_NOT_INVERTIBLE if a point hits the root). */
int hodor_cuda_ali_boundary_inverse_divisor_dev(void* d_out, uint32_t log_column, uint32_t log_evaluation, uint64_t row,
                                                int field_id, void* stream);
/* The same three with host output vectors (a Rust Vec<F>); any vector of precomputed_omegas may be NULL. */
int hodor_cuda_precomputed_omegas(uint64_t* omegas, uint64_t* coset, uint64_t* omegas_inv, uint32_t log_n, int field_id);
int hodor_cuda_ali_dense_inverse_divisor(uint64_t* out, uint32_t log_column, uint32_t log_evaluation, uint64_t start_at,
                                         uint64_t span, uint64_t num_rows, uint64_t* divisor_degree, int field_id);
int hodor_cuda_ali_boundary_inverse_divisor(uint64_t* out, uint32_t log_column, uint32_t log_evaluation, uint64_t row,
                                            int field_id);
/* distribute_powers (src/fft/mod.rs:110-123) in place on a device vector: a[j] <- a[j] * g^j */
int hodor_cuda_distribute_powers_dev(void* d_a, uint64_t n, const uint64_t g[4], int field_id, void* stream);
int hodor_cuda_elementwise_dev(int op, const void* d_a, const void* d_b, void* d_out, uint64_t n, int field_id,
                               void* stream);
int hodor_cuda_poly_op_dev(int op, const void* d_a, const void* d_b, const uint64_t scalar[4], uint64_t exp, void* d_out,
                           uint64_t n, int field_id, void* stream);
/* d_status: one device int, set to 0 on success and to 1 (vector untouched) when an element is zero */
int hodor_cuda_batch_inversion_dev(void* d_a, uint64_t n, int* d_status, int field_id, void* stream);
/* d_out: one element (32 B) on the device */
int hodor_cuda_evaluate_at_dev(const void* d_coeffs, uint64_t n, const uint64_t g[4], void* d_out, int field_id,
                               void* stream);
/* Building blocks for an LDE + FRI commit sharded over G = 2^log_g GPUs (hodor_b200/sharded.py):
 * - the cosets i = first_coset + coset_stride * t (t < 2^log_count) of the L = 2^log_factor coset LDE,
 *   interleaved among themselves: d_out[t + 2^log_count * k].  With first = rank, stride = G this is
 *   exactly rank's cyclic slice v[rank + G * tau] of the full LDE (cosets are independent:
 *   src/polynomials/mod.rs:572-587), computed with no communication;
 * - one FRI layer on such a cyclic slice (n_local = layer size / G values): the fold pairs
 *   (idx, idx + M/2) of src/fri/fri_on_values.rs:74-101 stay on one rank. */
int hodor_cuda_lde_cosets_dev(const void* d_coeffs, uint32_t log_n, uint32_t log_factor, int coset, uint32_t first_coset,
                              uint32_t coset_stride, uint32_t log_count, void* d_out, int field_id, void* stream);
int hodor_cuda_fri_fold_shard_dev(const void* d_in, uint64_t n_local, uint64_t initial_domain_size, uint32_t layer,
                                  uint32_t log_g, uint32_t rank, const void* d_challenge, void* d_out, int field_id,
                                  void* stream);
/* Four-step building blocks for an NTT sharded over G = 2^log_g GPUs (multi-GPU):
 * step A on rank r: n/G-point column NTTs of the rank's slice + twiddle by omega^(j2 * k1);
 * after the all-to-all, step B: row NTTs.  See hodor_b200/sharded.py for the orchestration. */
int hodor_cuda_ntt_shard_cols_dev(const void* d_in, void* d_out, uint32_t log_n, uint32_t log_g, uint32_t rank,
                                  const uint64_t omega[4], int field_id, void* stream);
int hodor_cuda_ntt_shard_rows_dev(const void* d_in, void* d_out, uint32_t log_n, uint32_t log_g, uint32_t rank,
                                  const uint64_t omega[4], int field_id, void* stream);

/* Merkle tree over n leaves that arrived as G = 2^log_g chunks of n/G elements, chunk r holding the cyclic slice
 * v[r + G*t] of the natural-order leaf vector (the receive side of the sharded chain's all-to-all): the
 * re-blocking is index arithmetic inside the leaf kernel's loads, not a transposing copy.  n/G > 1024. */
int hodor_cuda_merkle_build_shard_dev(const void* d_chunks, uint64_t n, uint32_t log_g, void* d_nodes, void* d_root,
                                      void* d_challenge, int field_id, void* stream);

/* ---- several GPUs: one process per GPU, NCCL for the exchange steps (multi-GPU) -------------
 * The reference's only parallelism is threads on one host (src/fft/multicore.rs); these entry points are the
 * `hodor_cuda_ntt_sharded` of SURVEY.md 8(b) and the north star's "2^24 -> 2^28 coset LDE plus full FRI commit
 * chain on 8 x B200".  NCCL is loaded at run time (libnccl.so.2, or $HODOR_NCCL_LIB); nothing else needs it.
 *
 * Rendezvous: rank 0 calls hodor_cuda_comm_unique_id and ships the 128 bytes to the other ranks by whatever
 * means the caller has (MPI, a file, torch.distributed); every rank then calls hodor_cuda_comm_init after
 * hodor_cuda_init(device).  world must be a power of two <= 16; world == 1 needs no NCCL (id may be NULL). */
int hodor_cuda_comm_unique_id(uint8_t id[128]);
int hodor_cuda_comm_init(int rank, int world, const uint8_t id[128]);
void hodor_cuda_comm_destroy(void);
/* bytes_sent: payload this rank has pushed through NCCL since init; bytes_peer_stored: payload its kernels stored
 * straight into other ranks' buffers over NVLink (the fused last pass of the sharded NTT).  Any pointer may be NULL. */
int hodor_cuda_comm_info(int* rank, int* world, uint64_t* bytes_sent, uint64_t* bytes_peer_stored);
/* Four-step (Bailey) NTT of length 2^log_n over the G ranks; best_fft's result (src/fft/fft.rs:5-125),
 * distributed.  d_local: this rank's cyclic slice a[j*G + rank], n/G elements; d_out: n/G elements, the
 * rank-th (n/G^2)-element chunk of every length-(n/G) block of the natural-order result:
 * d_out[k2 * n/G^2 + k] = A[k2 * n/G + rank * n/G^2 + k].  Stream ordered.  The one exchange (each rank ships
 * (G-1)/G of its slice once) is fused into the last pass of the local transform: its stores go straight into
 * the owning rank's receive buffer over NVLink (CUDA IPC peer mappings), bracketed by two 32-byte collectives;
 * where peer mapping is unavailable ($HODOR_NO_PEER_STORES, or cudaIpc fails) it is an NCCL send/recv
 * all-to-all between the two local steps.  log_n >= 2 * log2(G). */
int hodor_cuda_ntt_sharded(const void* d_local, void* d_out, uint32_t log_n, const uint64_t omega[4], int field_id,
                           void* stream);
/* ONE (coset) LDE 2^log_n -> 2^(log_n + log_factor) and its whole FRI commit chain
 * (src/polynomials/mod.rs:544-609 then src/fri/fri_on_values.rs:11-159) over all ranks.  d_coeffs: 2^log_n
 * coefficients, replicated on every rank.  Cosets are sharded with no communication, folds are local, each
 * committed layer costs one all-to-all and a 32-byte all-gather.  Outputs in HOST memory, identical on every
 * rank and bit-identical to the single-GPU chain: roots (steps + 1) * 32 B, challenges steps * 4 u64,
 * final_coeffs out_coeffs * 4 u64 (any may be NULL).  Returns the number of folding steps.  G <= 2^log_factor. */
int hodor_cuda_lde_fri_sharded(const void* d_coeffs, uint32_t log_n, uint32_t log_factor, int coset, uint32_t out_coeffs,
                               uint8_t* roots, uint64_t* challenges, uint64_t* final_coeffs, int field_id);

/* Diagnostic: runs the fixed-operand multiplier behind every table multiply (Field::mul_pre)
 * against the Montgomery multiplier on the device, with its rare carry fix-up path forced on.
 * Returns the number of disagreeing threads (0 = pass). */
int hodor_cuda_selftest_mul_pre(int field_id);
/* kernels launched by this library since init (bench.py's gpu_launches) */
uint64_t hodor_cuda_launch_count(void);
/* Optional per-kernel timing: between begin and end every launch is bracketed by CUDA events on
 * its stream; end synchronises the device and writes a JSON array
 * [{"name": ..., "count": ..., "total_ms": ...}] into json_out.  Returns the number of entries. */
int hodor_cuda_profile_begin(void);
int hodor_cuda_profile_end(char* json_out, size_t cap);

#ifdef __cplusplus
}
#endif
#endif /* HODOR_B200_H */
