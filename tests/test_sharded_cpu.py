"""world_size-2 (and 4) gloo test of the multi-GPU host logic in hodor_b200/sharded.py: the
distribution contract (scatter_input / gather_output) and the all-to-all plumbing of the four-step
NTT.  The two local compute steps are injected from the CPU oracle here, because this box has no
GPU; on the GPU box the same orchestration runs with CudaBackend (tests/test_gpu_parity.py checks
those kernels, bench.py runs the NCCL path)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class OracleBackend:
    """Test double for the two CUDA steps (TEST ONLY: uses oracle/)."""

    def __init__(self):
        from oracle import oracle as O
        self.O = O

    def shard_cols(self, src, log_n, log_g, rank, omega, field_id):
        O = self.O
        a = src.numpy().view(np.uint64)
        m = a.shape[0]
        wm = O.pow_(field_id, omega, 1 << log_g)
        b = O.serial_fft(field_id, a, wm, log_n - log_g)
        b = O.distribute_powers(field_id, b, O.pow_(field_id, omega, rank), cpus=1)
        assert b.shape[0] == m
        return torch.from_numpy(b.view(np.int64))

    def shard_rows(self, src, log_n, log_g, rank, omega, field_id):
        O = self.O
        G = 1 << log_g
        mat = src.numpy().view(np.uint64).reshape(G, -1, 4)
        cols = mat.shape[1]
        wg = O.pow_(field_id, omega, (1 << log_n) >> log_g)
        out = np.zeros_like(mat)
        for k in range(cols):
            out[:, k, :] = O.serial_fft(field_id, np.ascontiguousarray(mat[:, k, :]), wg, log_g)
        return torch.from_numpy(out.reshape(-1, 4).view(np.int64))


def _worker(rank, world, port, log_n, field_id, result_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from hodor_b200.sharded import ntt_sharded, scatter_input
        from oracle import oracle as O
        a = O.random_elements(field_id, 1 << log_n, seed=4321)
        omega = O.domain_generator(field_id, log_n)
        local = torch.from_numpy(scatter_input(a, world, rank).view(np.int64))
        out = ntt_sharded(local, log_n, omega, field_id, backend=OracleBackend())
        np.save(os.path.join(result_dir, f"out{rank}.npy"), out.numpy().view(np.uint64))
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


@pytest.mark.parametrize("world,log_n", [(2, 8), (4, 9)])
def test_four_step_over_gloo(oracle, tmp_path, world, log_n):
    from hodor_b200.sharded import gather_output
    field_id = 0
    mp.spawn(_worker, args=(world, _free_port(), log_n, field_id, str(tmp_path)), nprocs=world, join=True)
    parts = [np.load(tmp_path / f"out{r}.npy") for r in range(world)]
    a = oracle.random_elements(field_id, 1 << log_n, seed=4321)
    want = oracle.serial_fft(field_id, a, oracle.domain_generator(field_id, log_n), log_n)
    assert np.array_equal(gather_output(parts), want)


def test_sharded_argument_checks():
    from hodor_b200.sharded import gather_output, ntt_sharded, scatter_input
    a = np.arange(64, dtype=np.uint64).reshape(16, 4)
    assert np.array_equal(scatter_input(a, 4, 1), a[1::4])
    parts = [np.full((4, 4), h, np.uint64) for h in range(2)]
    g = gather_output(parts)  # m = 4, chunk = 2: [h0 h0 h1 h1 | h0 h0 h1 h1]
    assert [int(x) for x in g[:, 0]] == [0, 0, 1, 1, 0, 0, 1, 1]
    with pytest.raises(ValueError):
        ntt_sharded(torch.zeros((3, 4), dtype=torch.int64), 4, np.zeros(4, np.uint64), 0, backend=OracleBackend())


# --------------------------------------------------------------------------------------------------
# coset LDE + FRI commit chain sharded over ranks (hodor_b200/sharded_fri.py), gloo + oracle double
# --------------------------------------------------------------------------------------------------
class OracleFriBackend:
    """Test double for the CUDA steps of sharded_fri (TEST ONLY: uses oracle/)."""

    def __init__(self):
        from oracle import oracle as O
        self.O = O

    @staticmethod
    def _t(a):
        return torch.from_numpy(np.ascontiguousarray(a).view(np.int64).reshape(-1, 4))

    @staticmethod
    def _n(t):
        return t.contiguous().numpy().view(np.uint64).reshape(-1, 4)

    def lde_cosets(self, coeffs, log_n, log_factor, coset, first, stride, log_count, field_id):
        L, cnt, n = 1 << log_factor, 1 << log_count, 1 << log_n
        full = self.O.lde(field_id, self._n(coeffs), log_n, L, coset)
        out = np.zeros((n * cnt, 4), np.uint64)
        for t in range(cnt):
            out[t::cnt] = full[first + stride * t :: L]
        return self._t(out)

    def merkle_build(self, leaves, field_id):
        return self._t(self.O.merkle_create(field_id, self._n(leaves)).reshape(-1))

    def top_tree(self, sub_roots, field_id):
        O = self.O
        w = sub_roots.shape[0]
        top = [b""] * (2 * w)
        for q, r in enumerate(self._n(sub_roots)):
            top[w + q] = r.tobytes()
        for i in range(w - 1, 0, -1):
            top[i] = O.hash_node(top[2 * i], top[2 * i + 1])
        arr = np.frombuffer(b"".join(t if t else bytes(32) for t in top), np.uint64).reshape(-1, 4)
        return self._t(arr), self._t(O.interpret_hash(field_id, top[1]).reshape(1, 4))

    def leaf_blocks(self, values, blk_log, field_id):
        """One digest per block of 2^blk_log adjacent leaves: the root of that block's subtree."""
        v, B = self._n(values), 1 << blk_log
        out = [self.O.merkle_create(field_id, np.ascontiguousarray(v[i:i + B]))[1] for i in range(0, v.shape[0], B)]
        return self._t(np.frombuffer(b"".join(x.tobytes() for x in out), np.uint64).reshape(-1, 4))

    def tree_from_digests(self, digests, field_id):
        """Heap-ordered tree over a level of digests (root at [1]); the level itself is not stored."""
        d = [x.tobytes() for x in self._n(digests)]
        w = len(d)
        heap = [bytes(32)] * w + d
        for i in range(w - 1, 0, -1):
            heap[i] = self.O.hash_node(heap[2 * i], heap[2 * i + 1])
        return self._t(np.frombuffer(b"".join(heap[:w]), np.uint64).reshape(-1, 4))

    def fold_shard(self, values, initial_domain_size, layer, log_g, rank, challenge, field_id, blk_log=0):
        O, v = self.O, self._n(values)
        if isinstance(challenge, torch.Tensor):
            challenge = self._n(challenge)[0]
        half = v.shape[0] // 2
        log_n0 = initial_domain_size.bit_length() - 1
        winv = O.inverse(field_id, O.domain_generator(field_id, log_n0))
        from hodor_b200.sharded_fri import block_cyclic_index
        tw = np.stack([O.pow_(field_id, winv, block_cyclic_index(t, rank, 1 << log_g, blk_log) << layer) for t in range(half)])
        two_inv = O.inverse(field_id, O.to_mont(field_id, O.ints_to_array([2]))[0])
        f0, f1 = v[:half], v[half:]
        odd = O.mul(field_id, O.mul(field_id, O.sub(field_id, f0, f1), tw), np.tile(challenge, (half, 1)))
        return self._t(O.mul(field_id, O.add(field_id, odd, O.add(field_id, f0, f1)), np.tile(two_inv, (half, 1))))

    def fri_commit(self, values, lde_factor, out_coeffs, field_id):
        p = self.O.fri_commit(field_id, self._n(values), lde_factor, out_coeffs)
        return p.roots(), p.challenges, p.final_coefficients

    def hash_node(self, left, right):
        return self.O.hash_node(left, right)

    def root_to_challenge(self, root, field_id):
        return self.O.interpret_hash(field_id, root)


def _fri_worker(rank, world, port, log_n, log_factor, out_coeffs, gather_below, result_dir, blk_log=0):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from hodor_b200.sharded_fri import cyclic_to_block, fri_commit_sharded, lde_sharded
        from oracle import oracle as O
        be = OracleFriBackend()
        coeffs = O.random_elements(0, 1 << log_n, seed=99)
        local = lde_sharded(be._t(coeffs), log_n, log_factor, True, 0, backend=be, blk_log=blk_log)
        block = cyclic_to_block(local)
        proto = fri_commit_sharded(local, (1 << log_n) << log_factor, 1 << log_factor, out_coeffs, 0, backend=be,
                                   gather_below=gather_below, blk_log=blk_log)
        np.savez(os.path.join(result_dir, f"fri{rank}.npz"), local=be._n(local), block=be._n(block),
                 roots=np.frombuffer(b"".join(proto.roots), np.uint8), challenges=np.stack(proto.challenges),
                 final=proto.final_coefficients, sharded_layers=len(proto.commitments))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,log_n,log_factor,gather_below", [(2, 5, 3, 16), (4, 6, 2, 64), (2, 4, 1, 1 << 16)])
def test_sharded_lde_and_fri_chain_over_gloo(oracle, tmp_path, world, log_n, log_factor, gather_below):
    out_coeffs = 2
    mp.spawn(_fri_worker, args=(world, _free_port(), log_n, log_factor, out_coeffs, gather_below, str(tmp_path)),
             nprocs=world, join=True)
    coeffs = oracle.random_elements(0, 1 << log_n, seed=99)
    L = 1 << log_factor
    full = oracle.lde(0, coeffs, log_n, L, True)
    want = oracle.fri_commit(0, full, L, out_coeffs)
    m = full.shape[0] // world
    for r in range(world):
        res = np.load(tmp_path / f"fri{r}.npz")
        assert np.array_equal(res["local"], full[r::world])            # cyclic slice, no communication
        assert np.array_equal(res["block"], full[r * m : (r + 1) * m])  # after the all-to-all
        assert res["roots"].tobytes() == b"".join(want.roots())
        assert np.array_equal(res["challenges"], want.challenges)
        assert np.array_equal(res["final"], want.final_coefficients)
        if gather_below <= 64:
            assert int(res["sharded_layers"]) >= 2  # several layers really ran distributed


@pytest.mark.parametrize("world,log_n,log_factor,gather_below", [(2, 5, 3, 16), (4, 6, 3, 64), (2, 5, 1, 16)])
def test_block_cyclic_sharded_chain_over_gloo(oracle, tmp_path, world, log_n, log_factor, gather_below):
    """The distribution the C++ product path uses (csrc/sharded.cu): rank r owns the B = L/G adjacent cosets, holds
    blocks of B adjacent leaves, hashes the bottom log2 B tree levels locally and exchanges DIGESTS; folds stay local."""
    from hodor_b200.sharded_fri import block_cyclic_index
    out_coeffs = 2
    blk_log = log_factor - (world.bit_length() - 1)
    mp.spawn(_fri_worker, args=(world, _free_port(), log_n, log_factor, out_coeffs, gather_below, str(tmp_path), blk_log),
             nprocs=world, join=True)
    coeffs = oracle.random_elements(0, 1 << log_n, seed=99)
    L = 1 << log_factor
    full = oracle.lde(0, coeffs, log_n, L, True)
    want = oracle.fri_commit(0, full, L, out_coeffs)
    m = full.shape[0] // world
    for r in range(world):
        res = np.load(tmp_path / f"fri{r}.npz")
        idx = [block_cyclic_index(t, r, world, blk_log) for t in range(m)]
        assert np.array_equal(res["local"], full[idx])  # block-cyclic slice, no communication
        assert res["roots"].tobytes() == b"".join(want.roots())
        assert np.array_equal(res["challenges"], want.challenges)
        assert np.array_equal(res["final"], want.final_coefficients)
        assert int(res["sharded_layers"]) >= 1
    # fold pairs stay on one rank and land on the same distribution one layer down
    M = full.shape[0]
    for r in range(world):
        for t in range(m // 2):
            i = block_cyclic_index(t, r, world, blk_log)
            assert block_cyclic_index(t + m // 2, r, world, blk_log) == i + M // 2 and i < M // 2
