"""Coset LDE + FRI commit chain sharded over the G = 2^g GPUs of one box (BASELINE.json north star:
"2^24 -> 2^28 coset LDE plus full FRI commit chain on 8 x B200").  One process per GPU,
torch.distributed for the plumbing, hand-written CUDA through the C ABI for every compute step.

Distribution:

* LDE.  Cosets are independent (src/polynomials/mod.rs:572-587): rank r computes the cosets
  i = r, r+G, ... of the L-coset LDE from the (replicated) coefficient vector with no communication.
  Because out[i + L*k] and G | L, what rank r holds is exactly its CYCLIC slice v[r + G*tau] of the
  natural-order LDE.
* FRI fold.  The pairs (idx, idx + M/2) of src/fri/fri_on_values.rs:74-101 have equal residues mod G,
  so on cyclic slices every fold is local (hodor_cuda_fri_fold_shard_dev), layer after layer.
* Merkle trees want adjacent leaves together, i.e. natural-order BLOCKS.  The only exchange of the whole
  pipeline is therefore, per committed layer, one all-to-all that turns cyclic slices into blocks
  (each rank sends (G-1)/G of its slice once).  Rank q then builds the subtree over its block -- its
  local root is node G + q of the reference's heap layout -- the G sub-roots are all-gathered (32 B
  each, device to device) and every rank finishes the top log2(G) levels and the root -> challenge
  map redundantly on its GPU (hodor_cuda_merkle_top_dev).  The challenge stays in HBM for the next
  fold, so a committed layer costs no host round trip: roots and challenges are read back once, after
  the last layer.
* When a layer has shrunk below `gather_below` values the rest of the chain is tiny and strictly
  serial (root -> challenge -> fold), so the layer is all-gathered and finished on every rank with the
  single-GPU chain (hodor_cuda_fri_commit); its first tree is that layer's commitment.

This module is the torch.distributed statement of the algorithm with a pluggable compute backend: what the
world_size-2/4 gloo tests on the CPU run (oracle as compute double).  The product path is the same algorithm in
C++ behind the C ABI (`hodor_cuda_lde_fri_sharded`, csrc/sharded.cu; caller: hodor_b200/multigpu.py), which by
default uses the BLOCK-CYCLIC distribution (`blk_log` > 0 below): rank r computes the B = L/G adjacent cosets
r*B .. r*B+B-1, so it holds blocks of B adjacent leaves, v[(k*G + r)*B + c]; the bottom log2 B levels of every
tree are then local and the per-layer all-to-all moves the level of M/B digests instead of the M values.

The result is bit-identical to the single-GPU / reference chain: roots, challenges, final
coefficients (tests/test_sharded_cpu.py on gloo with the oracle as compute double;
tests/test_gpu_parity.py::test_sharded_lde_fri_single_gpu emulates all ranks on one GPU).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional, Protocol

import numpy as np
import torch
import torch.distributed as dist


class FriShardBackend(Protocol):
    def lde_cosets(self, coeffs, log_n, log_factor, coset, first, stride, log_count, field_id) -> torch.Tensor: ...

    def merkle_build(self, leaves, field_id) -> torch.Tensor: ...  # (n, 4) int64 nodes, heap order

    def fold_shard(self, values, initial_domain_size, layer, log_g, rank, challenge, field_id) -> torch.Tensor: ...
    # `challenge`: (1, 4) tensor as returned by top_tree

    def top_tree(self, sub_roots: torch.Tensor, field_id): ...  # (G, 4) -> (top nodes (2G, 4) heap order, challenge (1, 4))

    def fri_commit(self, values, lde_factor, out_coeffs, field_id): ...  # -> (roots, challenges, final_coeffs)

    def hash_node(self, left: bytes, right: bytes) -> bytes: ...

    def root_to_challenge(self, root: bytes, field_id) -> np.ndarray: ...


class CudaFriBackend:
    """The product path: libhodor_b200.so on torch's current stream."""

    def lde_cosets(self, coeffs, log_n, log_factor, coset, first, stride, log_count, field_id):
        from . import device as dev
        from ._ffi import check, ensure_init, lib

        ensure_init()
        out = dev.empty_elems((1 << log_n) << log_count, coeffs.device)
        check(lib.hodor_cuda_lde_cosets_dev(coeffs.data_ptr(), log_n, log_factor, int(coset), first, stride, log_count,
                                            out.data_ptr(), field_id, dev._stream()))
        return out

    def merkle_build(self, leaves, field_id):
        from . import device as dev

        nodes = torch.empty_like(leaves)
        dev.merkle_build(leaves, leaves.shape[0], nodes, field_id)
        return nodes

    def fold_shard(self, values, initial_domain_size, layer, log_g, rank, challenge, field_id):
        import ctypes as C
        from . import device as dev
        from ._ffi import check, ensure_init, lib

        ensure_init()
        if isinstance(challenge, torch.Tensor):
            d_chal = challenge
        else:
            d_chal = dev.to_device(np.ascontiguousarray(challenge, np.uint64).reshape(1, 4), values.device)
        out = dev.empty_elems(values.shape[0] // 2, values.device)
        check(lib.hodor_cuda_fri_fold_shard_dev(values.data_ptr(), C.c_uint64(values.shape[0]), C.c_uint64(initial_domain_size),
                                                layer, log_g, rank, d_chal.data_ptr(), out.data_ptr(), field_id, dev._stream()))
        return out

    def top_tree(self, sub_roots, field_id):
        from . import device as dev
        from ._ffi import check, ensure_init, lib

        ensure_init()
        w = sub_roots.shape[0]
        top = torch.zeros((2 * w, 4), dtype=torch.int64, device=sub_roots.device)
        top[w:] = sub_roots
        chal = torch.empty((1, 4), dtype=torch.int64, device=sub_roots.device)
        check(lib.hodor_cuda_merkle_top_dev(top.data_ptr(), w, None, chal.data_ptr(), field_id, dev._stream()))
        return top, chal

    def fri_commit(self, values, lde_factor, out_coeffs, field_id):
        from . import device as dev

        proto = dev.fri_commit(values.contiguous(), lde_factor, out_coeffs, field_id)
        res = (proto.get_roots(), proto.challenges.copy(), proto.final_coefficients.copy())
        proto.free()
        return res

    def hash_node(self, left, right):
        from .iop import Blake2sTreeHasher

        return Blake2sTreeHasher.hash_node([left, right])

    def root_to_challenge(self, root, field_id):
        from .iop import Blake2sLeafEncoder

        return Blake2sLeafEncoder.interpret_hash(field_id, root)


def _world(group):
    if dist.is_initialized():
        return dist.get_world_size(group), dist.get_rank(group)
    return 1, 0


def block_cyclic_index(t, rank: int, world: int, blk_log: int):
    """Natural index of local element t when blocks of 2^blk_log adjacent elements are dealt round-robin."""
    return (((t >> blk_log) * world + rank) << blk_log) | (t & ((1 << blk_log) - 1))


def lde_sharded(coeffs: torch.Tensor, log_n: int, log_factor: int, coset: bool, field_id: int, group=None,
                backend: Optional[FriShardBackend] = None, blk_log: int = 0) -> torch.Tensor:
    """Rank r's slice of the L-coset LDE; `coeffs` is replicated on every rank.  blk_log 0: the cyclic slice
    v[r + G*tau] (cosets r, r+G, ..); blk_log = log2(L/G): the block-cyclic slice (cosets r*B .. r*B+B-1)."""
    world, rank = _world(group)
    log_g = world.bit_length() - 1
    if world != 1 << log_g or log_g > log_factor:
        raise ValueError("world size must be a power of two not larger than the blowup factor")
    backend = backend or CudaFriBackend()
    if blk_log:
        if blk_log != log_factor - log_g:
            raise ValueError("block-cyclic sharding needs 2^blk_log == lde_factor / world")
        return backend.lde_cosets(coeffs, log_n, log_factor, coset, rank << blk_log, 1, blk_log, field_id)
    return backend.lde_cosets(coeffs, log_n, log_factor, coset, rank, world, log_factor - log_g, field_id)


def cyclic_to_block(local: torch.Tensor, group=None) -> torch.Tensor:
    """All-to-all: cyclic slice v[r + G*tau] -> natural-order block v[q*M/G .. (q+1)*M/G)."""
    world, _ = _world(group)
    if world == 1:
        return local
    m = local.shape[0]
    if m % world:
        raise ValueError("slice too short to re-block")
    recv = torch.empty_like(local)
    dist.all_to_all_single(recv, local.contiguous(), group=group)  # chunk q of rank r: tau in [q*m/G, (q+1)*m/G)
    # recv[r'][t'] = v[r' + G*(q*m/G + t')] = block element r' + G*t'  ->  interleave to (t', r')
    return recv.view(world, m // world, 4).transpose(0, 1).contiguous().view(m, 4)


@dataclass
class ShardedCommitment:
    """One committed layer: the local subtree (heap order, local root at [1] = global node G + rank),
    the top of the tree (global nodes 1 .. 2G-1, identical on every rank) and the layer's root."""

    size: int
    local_nodes: Optional[torch.Tensor]
    top_nodes: List[bytes]
    root: bytes

    def finalize(self) -> None:
        """Device -> host for the top of the tree (done once, after the chain)."""
        top = getattr(self, "_top", None)
        if top is not None:
            self.top_nodes = _digest_bytes(top)
            self.top_nodes[0] = b""
            self.root = self.top_nodes[1]
            self._top = None


@dataclass
class ShardedFriPrototype:
    roots: List[bytes] = field(default_factory=list)
    challenges: List[np.ndarray] = field(default_factory=list)
    final_root: bytes = b""
    final_coefficients: Optional[np.ndarray] = None
    commitments: List[ShardedCommitment] = field(default_factory=list)  # the layers committed while sharded
    layer_slices: List[torch.Tensor] = field(default_factory=list)     # this rank's cyclic slice of each of them
    num_steps: int = 0


def merkle_sharded(block_leaves: torch.Tensor, field_id: int, group=None,
                   backend: Optional[FriShardBackend] = None, from_digests: bool = False):
    """Commitment to a layer held as natural-order blocks.  Returns (commitment, challenge tensor); the
    commitment's `top_nodes` / `root` stay tensors until `finalize()` (no host synchronisation here)."""
    world, rank = _world(group)
    backend = backend or CudaFriBackend()
    nodes = backend.tree_from_digests(block_leaves, field_id) if from_digests else backend.merkle_build(block_leaves, field_id)
    sub_root = nodes[1:2].contiguous()
    if world > 1:
        gathered = torch.empty((world, 4), dtype=sub_root.dtype, device=sub_root.device)
        dist.all_gather_into_tensor(gathered, sub_root, group=group)
    else:
        gathered = sub_root
    top, challenge = backend.top_tree(gathered, field_id)
    com = ShardedCommitment(block_leaves.shape[0] * world, nodes, [], b"")
    com._top = top
    return com, challenge


def _digest_bytes(t: torch.Tensor) -> List[bytes]:
    a = t.cpu().numpy().view(np.uint8).reshape(-1, 32)
    return [bytes(r) for r in a]


def fri_commit_sharded(local_cyclic: torch.Tensor, domain_size: int, lde_factor: int, out_coeffs: int, field_id: int,
                       group=None, backend: Optional[FriShardBackend] = None, gather_below: int = 1 << 16,
                       keep_layers: bool = True, blk_log: int = 0) -> ShardedFriPrototype:
    """NaiveFriIop::proof_from_lde_by_values (src/fri/fri_on_values.rs:11-159) on an LDE held as
    cyclic slices.  Returns roots / challenges / final coefficients identical to the unsharded chain."""
    world, rank = _world(group)
    log_g = world.bit_length() - 1
    backend = backend or CudaFriBackend()
    steps = ((domain_size // lde_factor) // out_coeffs).bit_length() - 1
    if steps < 1:
        raise ValueError("zero folding steps (the reference panics here)")
    if local_cyclic.shape[0] * world != domain_size:
        raise ValueError("local slice has the wrong length")
    proto = ShardedFriPrototype(num_steps=steps)
    values, size, layer = local_cyclic, domain_size, 0
    gather_below = max(gather_below, (4 << blk_log) * world * world)
    pending = []  # (commitment, challenge tensor) of the layers committed while sharded
    while True:
        if size < gather_below or layer >= steps - 1:
            # finish on every rank with the single-GPU chain, starting at this layer's commitment
            # (at least one fold is always left for it, so it also produces the final coefficients)
            if world > 1:
                parts = [torch.empty_like(values) for _ in range(world)]
                dist.all_gather(parts, values.contiguous(), group=group)
                if blk_log:  # parts[r][(k << blk_log) | c] = v[((k*G + r) << blk_log) | c]
                    full = torch.stack([p.view(-1, 1 << blk_log, 4) for p in parts], dim=1).reshape(size, 4)
                else:
                    full = torch.stack(parts, dim=1).reshape(size, 4)  # v[r + G*tau] -> natural order
            else:
                full = values
            t_roots, t_chal, tail_final = backend.fri_commit(full, lde_factor, out_coeffs, field_id)
            proto.roots.extend(bytes(r) for r in t_roots)
            proto.challenges.extend(np.array(c, dtype=np.uint64) for c in t_chal)
            break
        if blk_log:
            # bottom log2 B levels locally (a block is a complete subtree), then re-block the DIGESTS
            digests = backend.leaf_blocks(values, blk_log, field_id)
            com, challenge = merkle_sharded(cyclic_to_block(digests, group), field_id, group, backend, from_digests=True)
            com.size = values.shape[0] * world
        else:
            com, challenge = merkle_sharded(cyclic_to_block(values, group), field_id, group, backend)
        pending.append((com, challenge))
        if not keep_layers:
            com.local_nodes = None  # let the allocator recycle the subtree
        if keep_layers:
            proto.commitments.append(com)
            proto.layer_slices.append(values)
        if blk_log:
            values = backend.fold_shard(values, domain_size, layer, log_g, rank, challenge, field_id, blk_log=blk_log)
        else:
            values = backend.fold_shard(values, domain_size, layer, log_g, rank, challenge, field_id)
        size //= 2
        layer += 1
    # one read-back for every sharded layer: roots and challenges, in order, before the tail's
    if pending:
        roots_t = torch.cat([c._top[1:2] for c, _ in pending])
        chals_t = torch.cat([ch for _, ch in pending])
        head_roots = _digest_bytes(roots_t)
        head_chals = [np.array(c, dtype=np.uint64) for c in chals_t.cpu().numpy().view(np.uint64).reshape(-1, 4)]
        for c, _ in pending:
            if keep_layers:
                c.finalize()
        proto.roots = head_roots + proto.roots
        proto.challenges = head_chals + proto.challenges
    proto.final_root = proto.roots[-1]
    proto.final_coefficients = np.array(tail_final, dtype=np.uint64)
    assert len(proto.roots) == steps + 1 and len(proto.challenges) == steps
    return proto
