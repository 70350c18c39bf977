"""Mirror of the reference's IOP / Merkle-oracle surface: traits in src/iop/mod.rs:22-92, the
trivial combiner src/iop/trivial_coset_combiner.rs, and Blake2sIopTree / TrivialBlake2sIOP in
src/iop/blake2s_trivial_iop.rs:107-339.

`create` -- the O(n) hashing -- runs on the GPU (hodor_cuda_merkle_build).  `get_path`, `verify` and
`query` are O(log n) host work exactly as in the reference; they use hashlib's RFC 7693 Blake2s with
the reference's key and personalisation, which is the verifier-side code a Rust caller would keep.
"""
from __future__ import annotations

import ctypes as C
import hashlib
import weakref
from dataclasses import dataclass
from typing import List, Optional, Sequence

import numpy as np

from . import field as fld
from ._ffi import check, ensure_init, lib, raise_last, u8p, vp
from .field import _p
from .polynomials import _as_elems

BLAKE2S_KEY = b"Squeamish Ossifrage"  # src/iop/blake2s_trivial_iop.rs:12
BLAKE2S_PERSONAL = b"Shaftoe"  # :13


def _H(data: bytes) -> bytes:
    return hashlib.blake2s(data, key=BLAKE2S_KEY, person=BLAKE2S_PERSONAL, digest_size=32).digest()


class TrivialCombiner:
    """src/iop/trivial_coset_combiner.rs:11-52: tree index == natural index."""

    EXPECTED_DEGREE = 2
    COSET_SIZE = 2

    @staticmethod
    def get_for_natural_index(leafs, natural_index: int):
        return leafs[natural_index]

    @staticmethod
    def get_for_tree_index(leafs, tree_index: int):
        return leafs[tree_index]

    @staticmethod
    def get_coset_for_natural_index(natural_index: int, domain_size: int) -> List[int]:
        assert natural_index < domain_size
        return sorted([natural_index, (natural_index + domain_size // 2) % domain_size])

    get_coset_for_tree_index = get_coset_for_natural_index

    @staticmethod
    def tree_index_into_natural_index(tree_index: int) -> int:
        return tree_index

    @staticmethod
    def natural_index_into_tree_index(natural_index: int) -> int:
        return natural_index


class Blake2sLeafEncoder:
    """:18-61"""

    @staticmethod
    def encode_leaf(value) -> bytes:
        return fld.limbs(value).tobytes()  # raw Montgomery limbs, little-endian (:36-42)

    @staticmethod
    def interpret_hash(field_id: int, digest: bytes) -> np.ndarray:
        d = np.frombuffer(bytes(digest), np.uint8).copy()
        out = np.zeros(4, np.uint64)
        check(lib.hodor_root_to_challenge(d.ctypes.data_as(u8p), _p(out), field_id))
        return out


class Blake2sTreeHasher:
    """:63-105 (host side; the GPU kernels implement the same function for `create`)."""

    @staticmethod
    def hash_leaf(value) -> bytes:
        return _H(Blake2sLeafEncoder.encode_leaf(value))

    @staticmethod
    def hash_encoded_leaf(value: bytes) -> bytes:
        return _H(value)

    @staticmethod
    def hash_node(values: Sequence[bytes], _level: int = 0) -> bytes:
        assert len(values) == 2
        return _H(bytes(values[0]) + bytes(values[1]))


class Blake2sIopTree:
    """:107-280.  `nodes` is the reference's Vec<[u8; 32]> in heap order."""

    Combiner = TrivialCombiner
    Hasher = Blake2sTreeHasher

    def __init__(self, field_id: int, size: int, nodes: np.ndarray):
        self.field_id = field_id
        self._size = size
        self.nodes = nodes

    @staticmethod
    def create(field_id: int, leafs) -> "Blake2sIopTree":
        leafs = _as_elems(leafs)
        n = leafs.shape[0]
        if n < 2 or n & (n - 1):
            raise AssertionError("assert!(num_leafs == num_leafs.next_power_of_two())")  # :137
        ensure_init()
        nodes = np.zeros((n, 32), np.uint8)
        check(lib.hodor_cuda_merkle_build(_p(leafs), C.c_uint64(n), nodes.ctypes.data_as(u8p), field_id))
        return Blake2sIopTree(field_id, n, nodes)

    def size(self) -> int:
        return self._size

    def get_root(self) -> bytes:
        return self.nodes[1].tobytes()  # :221-224

    @staticmethod
    def encode_root_into_challenge(field_id: int, root: bytes) -> np.ndarray:
        return Blake2sLeafEncoder.interpret_hash(field_id, root)

    def get_challenge_scalar_from_root(self) -> np.ndarray:
        return self.encode_root_into_challenge(self.field_id, self.get_root())

    @staticmethod
    def verify(root: bytes, leaf_value, path: Sequence[bytes], tree_index: int) -> bool:
        """:236-249"""
        h, idx = Blake2sTreeHasher.hash_leaf(leaf_value), tree_index
        for el in path:
            h = Blake2sTreeHasher.hash_node([h, el]) if idx & 1 == 0 else Blake2sTreeHasher.hash_node([el, h])
            idx >>= 1
        return h == bytes(root)

    def get_path(self, tree_index: int, leafs_values) -> List[bytes]:
        """:251-279"""
        leafs_values = _as_elems(leafs_values)
        assert self._size == self.nodes.shape[0]
        pair = TrivialCombiner.tree_index_into_natural_index(tree_index ^ 1)
        path = [Blake2sTreeHasher.hash_leaf(leafs_values[pair])]
        idx = (self._size + tree_index) >> 1
        while idx > 1:
            path.append(self.nodes[idx ^ 1].tobytes())
            idx >>= 1
        return path


@dataclass
class TrivialBlake2sIopQuery:
    """:343-368"""

    index: int
    _value: np.ndarray
    _path: List[bytes]

    def natural_index(self) -> int:
        return self.index

    def tree_index(self) -> int:
        return self.index

    def value(self) -> np.ndarray:
        return self._value

    def path(self) -> List[bytes]:
        return self._path

    def __eq__(self, other) -> bool:
        return (isinstance(other, TrivialBlake2sIopQuery) and self.index == other.index
                and np.array_equal(self._value, other._value) and self._path == other._path)


class TrivialBlake2sIOP:
    """:282-341"""

    Combiner = TrivialCombiner
    Tree = Blake2sIopTree
    Query = TrivialBlake2sIopQuery

    def __init__(self, tree: Blake2sIopTree):
        self.tree = tree

    @staticmethod
    def create(field_id: int, leafs) -> "TrivialBlake2sIOP":
        return TrivialBlake2sIOP(Blake2sIopTree.create(field_id, leafs))

    @staticmethod
    def get_for_natural_index(leafs, natural_index: int):
        return TrivialCombiner.get_for_natural_index(leafs, natural_index)

    @staticmethod
    def get_for_tree_index(leafs, tree_index: int):
        return TrivialCombiner.get_for_tree_index(leafs, tree_index)

    def get_root(self) -> bytes:
        return self.tree.get_root()

    @staticmethod
    def encode_root_into_challenge(field_id: int, root: bytes) -> np.ndarray:
        return Blake2sIopTree.encode_root_into_challenge(field_id, root)

    def get_challenge_scalar_from_root(self) -> np.ndarray:
        return self.tree.get_challenge_scalar_from_root()

    @staticmethod
    def verify_query(query: TrivialBlake2sIopQuery, root: bytes) -> bool:
        return Blake2sIopTree.verify(root, query.value(), query.path(), query.tree_index())

    def query(self, natural_index: int, leafs) -> TrivialBlake2sIopQuery:
        leafs = _as_elems(leafs)
        assert natural_index < self.tree.size()
        assert natural_index < leafs.shape[0]
        tree_index = TrivialCombiner.natural_index_into_tree_index(natural_index)
        return TrivialBlake2sIopQuery(natural_index, leafs[natural_index].copy(), self.tree.get_path(tree_index, leafs))

    def __eq__(self, other) -> bool:  # :336-340: equality is equality of roots
        return hasattr(other, "get_root") and self.get_root() == other.get_root()


class DeviceIOP:
    """An IOP whose tree stays in HBM behind a FRI prototype handle: roots come from
    the summary, queries are extracted on the device, `nodes` is fetched only if somebody asks."""

    def __init__(self, field_id: int, proto, layer: int, size: int, root: bytes):
        self.field_id = field_id
        self._proto_ref = weakref.ref(proto)  # no cycle: dropping the prototype frees its HBM at once
        self._layer = layer
        self._size = size
        self._root = root
        self._nodes: Optional[np.ndarray] = None

    def _proto(self):
        proto = self._proto_ref()
        if proto is None or not proto._handle:
            raise RuntimeError("the FRIProofPrototype that owns this commitment's device memory was freed")
        return proto

    def size(self) -> int:
        return self._size

    def get_root(self) -> bytes:
        return self._root

    def get_challenge_scalar_from_root(self) -> np.ndarray:
        return Blake2sIopTree.encode_root_into_challenge(self.field_id, self._root)

    @property
    def nodes(self) -> np.ndarray:
        if self._nodes is None:
            self._nodes = self._proto()._fetch_layer(self._layer, want_nodes=True)[0]
        return self._nodes

    def query(self, natural_index: int, leafs=None) -> TrivialBlake2sIopQuery:
        assert natural_index < self._size
        return self._proto()._query(self._layer, natural_index)

    verify_query = staticmethod(TrivialBlake2sIOP.verify_query)

    def __eq__(self, other) -> bool:
        return isinstance(other, (TrivialBlake2sIOP, DeviceIOP)) and self.get_root() == other.get_root()


class CommittedOracle:
    """`I::create(lde.as_ref())` with the leaves and the tree resident in HBM behind a `hodor_tree`
    handle -- the oracle a prover keeps per register between the commit phase and the query phase
    (src/prover/mod.rs:73-95 and :142-151).  Same surface as TrivialBlake2sIOP; `query` extracts the
    value and the authentication path on the device (log2(n) digests cross PCIe, not the tree)."""

    def __init__(self, field_id: int, handle: int, root: bytes):
        self.field_id = field_id
        self._handle = handle
        self._root = bytes(root)
        self._size = int(lib.hodor_cuda_tree_size(handle))
        self._keepalive = None

    # ---- constructors ---------------------------------------------------------------------------
    @staticmethod
    def create(field_id: int, leafs) -> "CommittedOracle":
        """IOP::create on host values (copied in once, then owned by the handle)."""
        leafs = _as_elems(leafs)
        n = leafs.shape[0]
        if n < 2 or n & (n - 1):
            raise AssertionError("assert!(num_leafs == num_leafs.next_power_of_two())")
        ensure_init()
        root = np.zeros(32, np.uint8)
        h = lib.hodor_cuda_tree_commit(leafs.ctypes.data, C.c_uint64(n), 0, root.ctypes.data_as(u8p), field_id)
        if not h:
            raise_last()
        return CommittedOracle(field_id, h, root.tobytes())

    @staticmethod
    def create_on_device(field_id: int, d_values) -> "CommittedOracle":
        """IOP::create on a device-resident vector (torch tensor (n, 4) int64); the tensor is borrowed."""
        n = int(d_values.shape[0])
        if n < 2 or n & (n - 1):
            raise AssertionError("assert!(num_leafs == num_leafs.next_power_of_two())")
        ensure_init()
        import torch
        torch.cuda.current_stream().synchronize()  # the commit runs on the library's stream
        root = np.zeros(32, np.uint8)
        h = lib.hodor_cuda_tree_commit(d_values.data_ptr(), C.c_uint64(n), 1, root.ctypes.data_as(u8p), field_id)
        if not h:
            raise_last()
        o = CommittedOracle(field_id, h, root.tobytes())
        o._keepalive = d_values
        return o

    @staticmethod
    def lde_commit_batch(polys: Sequence, factor: int, coset: bool = False) -> List["CommittedOracle"]:
        """`for w in witness { let lde = w.lde(..)?; I::create(lde.as_ref()) }` in one pipelined call."""
        from .polynomials import COEFFICIENTS
        if not polys:
            return []
        if factor < 1 or factor & (factor - 1):
            raise AssertionError("assert!(factor.is_power_of_two())")
        fid, exp = polys[0].field_id, polys[0].exp
        for p in polys:
            if p.form != COEFFICIENTS:
                raise TypeError("lde needs Polynomial<F, Coefficients>")
            if p.field_id != fid or p.exp != exp:
                raise ValueError("lde_commit_batch: polynomials must share field and size")
        ensure_init()
        count = len(polys)
        ins = (vp * count)(*[p.as_ref().ctypes.data for p in polys])
        outs = (vp * count)()
        roots = np.zeros((count, 32), np.uint8)
        check(lib.hodor_cuda_lde_commit_batch(ins, count, exp, factor.bit_length() - 1, int(coset), 0, outs,
                                              roots.ctypes.data_as(u8p), fid))
        return [CommittedOracle(fid, outs[i], roots[i].tobytes()) for i in range(count)]

    @staticmethod
    def lde_commit(poly, factor: int, coset: bool = False) -> "CommittedOracle":
        return CommittedOracle.lde_commit_batch([poly], factor, coset)[0]

    # ---- lifetime --------------------------------------------------------------------------------
    def free(self) -> None:
        h, self._handle = getattr(self, "_handle", None), None
        if h:
            lib.hodor_cuda_tree_free(h)

    def __del__(self):
        self.free()

    def _h(self) -> int:
        if not self._handle:
            raise RuntimeError("this CommittedOracle's device memory was freed")
        return self._handle

    # ---- IOP surface -----------------------------------------------------------------------------
    def size(self) -> int:
        return self._size

    def get_root(self) -> bytes:
        return self._root

    def get_challenge_scalar_from_root(self) -> np.ndarray:
        out = np.zeros(4, np.uint64)
        check(lib.hodor_cuda_tree_root(self._h(), None, _p(out)))  # computed on the device with the root
        return out

    def device_values_ptr(self) -> int:
        return int(lib.hodor_cuda_tree_values(self._h()))

    def values(self, first: int = 0, count: Optional[int] = None) -> np.ndarray:
        count = self._size - first if count is None else count
        out = np.zeros((count, 4), np.uint64)
        check(lib.hodor_cuda_tree_read(self._h(), C.c_uint64(first), C.c_uint64(count), _p(out), None))
        return out

    @property
    def nodes(self) -> np.ndarray:
        out = np.zeros((self._size, 32), np.uint8)
        check(lib.hodor_cuda_tree_read(self._h(), C.c_uint64(0), C.c_uint64(self._size), None, out.ctypes.data_as(u8p)))
        return out

    def query(self, natural_index: int, leafs=None) -> TrivialBlake2sIopQuery:
        return self.query_batch([natural_index])[0]

    def query_batch(self, natural_indices: Sequence[int]) -> List[TrivialBlake2sIopQuery]:
        for i in natural_indices:
            assert 0 <= i < self._size  # reference: assert!(natural_index < self.tree.size())
        k = len(natural_indices)
        depth = self._size.bit_length() - 1
        idx = np.asarray(list(natural_indices), dtype=np.uint64)
        values = np.zeros((k, 4), np.uint64)
        paths = np.zeros((k, depth, 32), np.uint8)
        n = check(lib.hodor_cuda_tree_query_batch(self._h(), _p(idx) if k else None, k, _p(values), paths.ctypes.data_as(u8p)))
        assert n == depth
        return [TrivialBlake2sIopQuery(int(idx[i]), values[i].copy(), [d.tobytes() for d in paths[i]]) for i in range(k)]

    verify_query = staticmethod(TrivialBlake2sIOP.verify_query)

    def __eq__(self, other) -> bool:
        return hasattr(other, "get_root") and self.get_root() == other.get_root()
