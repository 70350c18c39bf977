"""Multi-GPU parity + timing of the C-ABI sharded entry points (hodor_cuda_ntt_sharded, hodor_cuda_lde_fri_sharded).
    python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 tools/sharded_check.py [log_n] [log_factor] [ntt_log_n]
Every rank runs the sharded pipeline (NCCL send/recv issued by the library); rank 0 also runs the single-GPU
path on the same input and the results must be bit-identical: the gathered four-step NTT equals the single-GPU
NTT element for element, and roots / challenges / final coefficients of the sharded chain equal the single-GPU
chain's.  Prints one JSON line per check; exit code 1 on any mismatch.  Also imported by tests/ and bench.py."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist


def synthetic(count, seed):
    rng = np.random.default_rng(seed)
    a = rng.integers(0, 2**64, size=(count, 4), dtype=np.uint64)
    a[:, 3] = rng.integers(0, 0x73EDA753299D7D48, size=count, dtype=np.uint64)
    return a


def check_ntt(log_n, fid=0, reps=3):
    """Sharded NTT == single-GPU NTT (gathered on rank 0).  Returns dict (ok only meaningful on rank 0)."""
    import hodor_b200 as H
    from hodor_b200 import device as dev
    from hodor_b200 import multigpu as mg
    from hodor_b200.sharded import gather_output

    rank, world = mg.comm_init()
    n = 1 << log_n
    a = synthetic(n, 31 + log_n)  # the same vector on every rank; each uses its cyclic slice
    omega = H.Domain.new_for_size(fid, n).generator
    local = dev.to_device(np.ascontiguousarray(a[rank::world]))
    out = mg.ntt_sharded(local, log_n, omega, fid)
    torch.cuda.synchronize()
    times = []
    for _ in range(reps):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        mg.ntt_sharded(local, log_n, omega, fid, out)
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        times.append(float(t.item()))
    if world > 1:
        parts = [torch.empty_like(out) for _ in range(world)] if rank == 0 else None
        dist.gather(out, parts, dst=0)
    else:
        parts = [out]
    ok = True
    if rank == 0:
        got = gather_output([p.cpu().numpy().view(np.uint64) for p in parts])
        full = dev.to_device(a)
        ref = dev.empty_elems(n)
        dev.fft(full, ref, log_n, False, fid)
        ok = bool(np.array_equal(got, dev.to_host(ref)))
    return {"check": "four-step sharded NTT == single-GPU NTT", "ok": ok, "n_gpus": world, "log_n": log_n, "ms": min(times),
            "ms_all_reps": times}


def ntt_phases(log_n, fid=0):
    """Per-rank kernel times of one sharded NTT (the library's own per-launch events) beside the whole call: what is
    left over is the two barrier collectives and waiting for the slowest peer.  No data check (check_ntt does that)."""
    import ctypes as C

    import hodor_b200 as H
    from hodor_b200 import multigpu as mg
    from hodor_b200._ffi import lib

    rank, world = mg.comm_init()
    m = (1 << log_n) // world
    omega = H.Domain.new_for_size(fid, 1 << log_n).generator
    g = torch.Generator(device="cuda").manual_seed(1000 + rank)
    local = torch.randint(0, 1 << 62, (m, 4), dtype=torch.int64, device="cuda", generator=g)
    out = torch.empty_like(local)
    for _ in range(2):
        mg.ntt_sharded(local, log_n, omega, fid, out)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    lib.hodor_cuda_profile_begin()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    mg.ntt_sharded(local, log_n, omega, fid, out)
    e1.record()
    torch.cuda.synchronize()
    buf = C.create_string_buffer(1 << 16)
    lib.hodor_cuda_profile_end(buf, len(buf))
    mine = {"rank": rank, "call_ms": e0.elapsed_time(e1), "kernels": json.loads(buf.value.decode())}
    if world > 1:
        allr = [None] * world
        dist.all_gather_object(allr, mine)
    else:
        allr = [mine]
    return {"check": "phases of one sharded NTT (per rank)", "ok": True, "n_gpus": world, "log_n": log_n, "ranks": allr}


def lde_fri_phases(log_n, log_f, fid=0):
    """Per-rank kernel times of one sharded LDE + FRI chain beside the whole (synchronous) call."""
    import ctypes as C
    import time

    from hodor_b200 import device as dev
    from hodor_b200 import multigpu as mg
    from hodor_b200._ffi import lib

    rank, world = mg.comm_init()
    d_coeffs = dev.to_device(synthetic(1 << log_n, 77))
    for _ in range(2):
        mg.lde_fri_sharded(d_coeffs, log_n, log_f, True, 1, fid)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    lib.hodor_cuda_profile_begin()
    t0 = time.perf_counter()
    mg.lde_fri_sharded(d_coeffs, log_n, log_f, True, 1, fid)
    call_ms = (time.perf_counter() - t0) * 1e3
    buf = C.create_string_buffer(1 << 16)
    lib.hodor_cuda_profile_end(buf, len(buf))
    kernels = json.loads(buf.value.decode())
    mine = {"rank": rank, "call_ms": call_ms, "kernel_sum_ms": sum(k["total_ms"] for k in kernels), "kernels": kernels}
    if world > 1:
        allr = [None] * world
        dist.all_gather_object(allr, mine)
    else:
        allr = [mine]
    return {"check": "phases of one sharded LDE + FRI chain (per rank)", "ok": True, "n_gpus": world, "log_n": log_n,
            "lde_factor": 1 << log_f, "ranks": allr}


def check_lde_fri(log_n, log_f, fid=0, reps=3):
    from hodor_b200 import device as dev
    from hodor_b200 import multigpu as mg

    rank, world = mg.comm_init()
    n, L = 1 << log_n, 1 << log_f
    d_coeffs = dev.to_device(synthetic(n, 7))  # replicated coefficient vector
    res = mg.lde_fri_sharded(d_coeffs, log_n, log_f, True, 1, fid)  # warm-up (tables, NCCL channels)
    times = []
    for _ in range(reps):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        res = mg.lde_fri_sharded(d_coeffs, log_n, log_f, True, 1, fid)
        torch.cuda.synchronize()
        t = torch.tensor([(time.perf_counter() - t0) * 1e3], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        times.append(float(t.item()))
    ok = True
    if rank == 0:
        full = dev.empty_elems(n * L)
        dev.lde(d_coeffs, log_n, log_f, True, full, fid)
        ref = dev.fri_commit(full, L, 1, fid)
        ok = bool(res[0] == ref.get_roots() and np.array_equal(res[1], ref.challenges)
                  and np.array_equal(res[2], ref.final_coefficients))
        ref.free()
    return {"check": "sharded coset LDE + FRI commit == single-GPU chain", "ok": ok, "n_gpus": world, "log_n": log_n,
            "lde_factor": L, "domain": n * L, "layers": len(res[1]), "ms_lde_plus_fri": min(times), "ms_all_reps": times,
            "lde_elems_per_s": n * L / (min(times) * 1e-3)}


def main():
    import hodor_b200 as H

    log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    log_f = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    ntt_log_n = int(sys.argv[3]) if len(sys.argv) > 3 else log_n + 2
    rank, local_rank, world = (int(os.environ.get(k, 0)) for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE"))
    world = max(world, 1)
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    H.init(local_rank)
    if os.environ.get("HODOR_CHECK_PHASES"):  # e.g. "28,26": only the per-phase timing of the sharded NTT
        results = [ntt_phases(int(ln)) for ln in os.environ["HODOR_CHECK_PHASES"].split(",") if ln]
        for g in [x for x in os.environ.get("HODOR_CHECK_GATHER_SWEEP", "").split(",") if x]:  # "16,18,20,22"
            os.environ["HODOR_SHARD_GATHER_LOG2"] = g
            r = check_lde_fri(log_n, log_f)
            r["gather_log2"] = int(g)
            results.append(r)
            lp = lde_fri_phases(log_n, log_f)
            lp["gather_log2"] = int(g)
            results.append(lp)
        os.environ.pop("HODOR_SHARD_GATHER_LOG2", None)
        if os.environ.get("HODOR_CHECK_PHASES_LDE"):  # "24,4"
            a, b = (int(x) for x in os.environ["HODOR_CHECK_PHASES_LDE"].split(","))
            results.append(lde_fri_phases(a, b))
    elif os.environ.get("HODOR_CHECK_NTT_SIZES"):  # e.g. "14,18,22,20": the receive buffer grows (re-mapped by the peers) and shrinks
        results = [check_ntt(int(ln)) for ln in os.environ["HODOR_CHECK_NTT_SIZES"].split(",") if ln]
    else:
        results = [check_ntt(ntt_log_n), check_ntt(16), check_lde_fri(log_n, log_f), check_lde_fri(14, 4)]
    ok = all(r["ok"] for r in results)
    if rank == 0:
        for r in results:
            real_stdout.write(json.dumps(r) + "\n")
        real_stdout.flush()
    flag = torch.tensor([0 if ok else 1], device="cuda")
    if world > 1:
        dist.broadcast(flag, src=0)
        from hodor_b200 import multigpu as mg
        mg.comm_destroy()
        dist.destroy_process_group()
    sys.exit(int(flag.item()))


if __name__ == "__main__":
    main()
