//! Seam 2 (src/iop/mod.rs:58-92): the Merkle oracle built on the GPU.
//!
//! `CudaBlake2sIopTree<F>` keeps `nodes: Vec<[u8; 32]>` in host memory in the reference's heap
//! layout (nodes[1] = root, level of w nodes at [w, 2w), nodes[0] zero), filled by
//! `hodor_cuda_merkle_build`; everything that is O(log n) stays the reference's host code.
//! `CommittedOracle<F>` keeps values AND tree in HBM behind a `hodor_tree` handle: `create` uploads
//! the leaves once, `query` extracts value + authentication path on the device.
use std::marker::PhantomData;

use crate::iop::blake2s_trivial_iop::*;
use crate::iop::trivial_coset_combiner::TrivialCombiner;
use crate::iop::*;

use super::ffi::{self, CudaField};

type Digest = [u8; 32];

pub struct CudaBlake2sIopTree<F: CudaField> {
    size: u64,
    nodes: Vec<Digest>,
    _marker: PhantomData<F>,
}

impl<F: CudaField> IopTree<F> for CudaBlake2sIopTree<F> {
    type Combiner = TrivialCombiner<F>;
    type Hasher = Blake2sTreeHasher<F>;

    fn create(leafs: &[F]) -> Self {
        let num_leafs = leafs.len();
        assert!(num_leafs == num_leafs.next_power_of_two()); // src/iop/blake2s_trivial_iop.rs:137
        ffi::init();
        let mut nodes = vec![[0u8; 32]; num_leafs];
        let rc = unsafe { ffi::hodor_cuda_merkle_build(ffi::as_u64(leafs), num_leafs as u64, nodes.as_mut_ptr() as *mut u8, F::FIELD_ID) };
        assert!(rc == ffi::OK, "hodor_cuda_merkle_build failed: {}", ffi::last_error());
        Self { size: num_leafs as u64, nodes, _marker: PhantomData }
    }

    fn size(&self) -> u64 {
        self.size
    }

    fn get_root(&self) -> Digest {
        self.nodes[1]
    }

    fn encode_root_into_challenge(root: &Digest) -> F {
        // O(1) host work: the reference's own interpret_hash (:48-60)
        <Blake2sIopTree<F> as IopTree<F>>::encode_root_into_challenge(root)
    }

    fn get_challenge_scalar_from_root(&self) -> F {
        Self::encode_root_into_challenge(&self.get_root())
    }

    fn verify(root: &Digest, leaf_value: &F, path: &[Digest], tree_index: usize) -> bool {
        <Blake2sIopTree<F> as IopTree<F>>::verify(root, leaf_value, path, tree_index)
    }

    /// Same path as the reference's get_path (:251-279), read off the heap indices: the hash of the
    /// sibling leaf, then the sibling of every ancestor below the root.
    fn get_path(&self, tree_index: usize, leafs_values: &[F]) -> Vec<Digest> {
        assert!(self.size == self.nodes.len() as u64);
        let sibling_leaf = <Self::Combiner as CosetCombiner<F>>::tree_index_into_natural_index(tree_index ^ 1);
        let mut path = vec![<Self::Hasher as IopTreeHasher<F>>::hash_leaf(&leafs_values[sibling_leaf])];
        let mut heap = (self.size as usize + tree_index) >> 1; // the leaf pair's parent
        while heap > 1 {
            path.push(self.nodes[heap ^ 1]);
            heap >>= 1;
        }
        path
    }
}

impl<F: CudaField> CudaBlake2sIopTree<F> {
    pub fn nodes(&self) -> &[Digest] {
        &self.nodes
    }
}

/// `TrivialBlake2sIOP` (src/iop/blake2s_trivial_iop.rs:282-341) over the GPU-built tree: the `I`
/// type argument of `Prover` / `NaiveFriIop`.
pub struct CudaBlake2sIOP<F: CudaField> {
    tree: CudaBlake2sIopTree<F>,
}

impl<F: CudaField> IOP<F> for CudaBlake2sIOP<F> {
    type Combiner = TrivialCombiner<F>;
    type Tree = CudaBlake2sIopTree<F>;
    type Query = TrivialBlake2sIopQuery<F>;

    fn create(leafs: &[F]) -> Self {
        Self { tree: CudaBlake2sIopTree::create(leafs) }
    }
    fn get_for_natural_index(leafs: &[F], natural_index: usize) -> &F {
        <Self::Combiner as CosetCombiner<F>>::get_for_natural_index(leafs, natural_index)
    }
    fn get_for_tree_index(leafs: &[F], tree_index: usize) -> &F {
        <Self::Combiner as CosetCombiner<F>>::get_for_tree_index(leafs, tree_index)
    }
    fn get_root(&self) -> Digest {
        self.tree.get_root()
    }
    fn encode_root_into_challenge(root: &Digest) -> F {
        <Self::Tree as IopTree<F>>::encode_root_into_challenge(root)
    }
    fn get_challenge_scalar_from_root(&self) -> F {
        self.tree.get_challenge_scalar_from_root()
    }
    fn verify_query(query: &Self::Query, root: &Digest) -> bool {
        <Self::Tree as IopTree<F>>::verify(root, &query.value(), query.path(), query.tree_index())
    }
    fn query(&self, natural_index: usize, leafs: &[F]) -> Self::Query {
        assert!(natural_index < self.tree.size() as usize);
        assert!(natural_index < leafs.len());
        let tree_index = <Self::Combiner as CosetCombiner<F>>::natural_index_into_tree_index(natural_index);
        // `from_parts` is the pub(crate) constructor hodor_cuda.patch adds (the fields are private)
        TrivialBlake2sIopQuery::from_parts(natural_index, leafs[natural_index], self.tree.get_path(tree_index, leafs))
    }
}

impl<F: CudaField> PartialEq for CudaBlake2sIOP<F> {
    fn eq(&self, other: &Self) -> bool {
        self.get_root() == other.get_root()
    }
}
impl<F: CudaField> Eq for CudaBlake2sIOP<F> {}

/// An oracle whose leaves and tree live in HBM (`hodor_tree`).  As `I: IOP<F>` it is a drop-in for
/// `Prover`: `I::create(lde.as_ref())` uploads the values once and returns after the root is known;
/// `query(idx, _)` ignores the host slice and reads value + path from the device.  `lde_commit` goes
/// one step further and never materialises the LDE on the host (src/prover/mod.rs:73-80 in one call).
pub struct CommittedOracle<F: CudaField> {
    handle: *mut ffi::Tree,
    size: u64,
    root: Digest,
    _marker: PhantomData<F>,
}

// the handle is only ever used through the library, which serialises calls on its context
unsafe impl<F: CudaField> Send for CommittedOracle<F> {}
unsafe impl<F: CudaField> Sync for CommittedOracle<F> {}

impl<F: CudaField> Drop for CommittedOracle<F> {
    fn drop(&mut self) {
        unsafe { ffi::hodor_cuda_tree_free(self.handle) }
    }
}

impl<F: CudaField> CommittedOracle<F> {
    fn from_handle(handle: *mut ffi::Tree, root: Digest) -> Self {
        assert!(!handle.is_null(), "hodor_cuda commit failed: {}", ffi::last_error());
        let size = unsafe { ffi::hodor_cuda_tree_size(handle) };
        Self { handle, size, root, _marker: PhantomData }
    }

    /// `let lde = w.lde(&worker, factor)?; let oracle = I::create(lde.as_ref());` without the LDE ever
    /// leaving the GPU.
    pub fn lde_commit(coeffs: &crate::polynomials::Polynomial<F, crate::polynomials::Coefficients>, factor: usize, coset: bool)
        -> Result<Self, crate::SynthesisError>
    {
        assert!(factor.is_power_of_two());
        ffi::init();
        let _ = crate::domains::Domain::<F>::new_for_size((coeffs.size() * factor) as u64)?;
        let mut root = [0u8; 32];
        let h = unsafe {
            ffi::hodor_cuda_lde_commit(ffi::as_u64(coeffs.as_ref()), coeffs.exp, factor.trailing_zeros(), coset as i32, 0,
                                       root.as_mut_ptr(), F::FIELD_ID)
        };
        if h.is_null() {
            return Err(crate::SynthesisError::Error);
        }
        Ok(Self::from_handle(h, root))
    }

    /// Copies the committed values back (e.g. for the DEEP step while that still runs on the host).
    pub fn values(&self) -> Vec<F> {
        let mut out = vec![F::zero(); self.size as usize];
        let rc = unsafe { ffi::hodor_cuda_tree_read(self.handle, 0, self.size, ffi::as_u64_mut(&mut out), std::ptr::null_mut()) };
        assert!(rc == ffi::OK, "hodor_cuda_tree_read failed: {}", ffi::last_error());
        out
    }

    /// Several openings in one launch and one pair of copies (the query phase of `Prover::prove` opens every
    /// register oracle at the same indices, src/prover/mod.rs:142-151).
    pub fn query_batch(&self, natural_indices: &[usize]) -> Vec<TrivialBlake2sIopQuery<F>> {
        let depth = self.size.trailing_zeros() as usize;
        let idx: Vec<u64> = natural_indices.iter().map(|&i| i as u64).collect();
        assert!(idx.iter().all(|&i| i < self.size));
        let mut values = vec![F::zero(); idx.len()];
        let mut paths = vec![[0u8; 32]; idx.len() * depth];
        let rc = unsafe {
            ffi::hodor_cuda_tree_query_batch(self.handle, idx.as_ptr(), idx.len() as u32, ffi::as_u64_mut(&mut values),
                                             paths.as_mut_ptr() as *mut u8)
        };
        assert!(rc == depth as i32, "hodor_cuda_tree_query_batch failed: {}", ffi::last_error());
        natural_indices.iter().enumerate()
            .map(|(k, &i)| TrivialBlake2sIopQuery::from_parts(i, values[k], paths[k * depth..(k + 1) * depth].to_vec()))
            .collect()
    }

    /// Device pointer of the values: input of `hodor_cuda_fri_commit(.., lde_on_device = 1, ..)`.
    pub fn device_values(&self) -> *const u64 {
        unsafe { ffi::hodor_cuda_tree_values(self.handle) as *const u64 }
    }
}

impl<F: CudaField> IOP<F> for CommittedOracle<F> {
    type Combiner = TrivialCombiner<F>;
    type Tree = CudaBlake2sIopTree<F>;
    type Query = TrivialBlake2sIopQuery<F>;

    fn create(leafs: &[F]) -> Self {
        assert!(leafs.len() == leafs.len().next_power_of_two());
        ffi::init();
        let mut root = [0u8; 32];
        let h = unsafe { ffi::hodor_cuda_tree_commit(ffi::as_u64(leafs), leafs.len() as u64, 0, root.as_mut_ptr(), F::FIELD_ID) };
        Self::from_handle(h, root)
    }
    fn get_for_natural_index(leafs: &[F], natural_index: usize) -> &F {
        <Self::Combiner as CosetCombiner<F>>::get_for_natural_index(leafs, natural_index)
    }
    fn get_for_tree_index(leafs: &[F], tree_index: usize) -> &F {
        <Self::Combiner as CosetCombiner<F>>::get_for_tree_index(leafs, tree_index)
    }
    fn get_root(&self) -> Digest {
        self.root
    }
    fn encode_root_into_challenge(root: &Digest) -> F {
        <Self::Tree as IopTree<F>>::encode_root_into_challenge(root)
    }
    fn get_challenge_scalar_from_root(&self) -> F {
        Self::encode_root_into_challenge(&self.root)
    }
    fn verify_query(query: &Self::Query, root: &Digest) -> bool {
        <Self::Tree as IopTree<F>>::verify(root, &query.value(), query.path(), query.tree_index())
    }
    fn query(&self, natural_index: usize, _leafs: &[F]) -> Self::Query {
        assert!((natural_index as u64) < self.size);
        let depth = self.size.trailing_zeros() as usize;
        let mut value = F::zero();
        let mut path = vec![[0u8; 32]; depth];
        let rc = unsafe {
            ffi::hodor_cuda_tree_query(self.handle, natural_index as u64, &mut value as *mut F as *mut u64, path.as_mut_ptr() as *mut u8)
        };
        assert!(rc == depth as i32, "hodor_cuda_tree_query failed: {}", ffi::last_error());
        TrivialBlake2sIopQuery::from_parts(natural_index, value, path)
    }
}

impl<F: CudaField> PartialEq for CommittedOracle<F> {
    fn eq(&self, other: &Self) -> bool {
        self.root == other.root
    }
}
impl<F: CudaField> Eq for CommittedOracle<F> {}
