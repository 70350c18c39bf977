"""Multi-GPU parity + timing of the sharded coset LDE + FRI commit chain (hodor_b200/sharded_fri.py).
    python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 tools/sharded_check.py [log_n] [log_factor]
Every rank computes its slice with the sharded pipeline (NCCL all-to-all per committed layer); rank 0
also runs the single-GPU chain on the same polynomial and the results must be bit-identical."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

import hodor_b200 as H
from hodor_b200 import device as dev
from hodor_b200.sharded_fri import fri_commit_sharded, lde_sharded

log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
log_f = int(sys.argv[2]) if len(sys.argv) > 2 else 3
rank, local_rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
real_stdout = os.fdopen(os.dup(1), "w")
os.dup2(2, 1)
torch.cuda.set_device(local_rank)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
H.init(local_rank)
fid, L = 0, 1 << log_f
n = 1 << log_n
rng = np.random.default_rng(7)  # same polynomial on every rank (the coefficient vector is replicated)
coeffs = rng.integers(0, 2**64, size=(n, 4), dtype=np.uint64)
coeffs[:, 3] = rng.integers(0, 0x73EDA753299D7D48, size=n, dtype=np.uint64)
d_coeffs = dev.to_device(coeffs)


def run():
    local = lde_sharded(d_coeffs, log_n, log_f, True, fid)
    return local, fri_commit_sharded(local, n * L, L, 1, fid, keep_layers=False)


local, proto = run()  # warm-up (tables, NCCL)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
reps, rep_ms = 3, []
for _ in range(reps):
    t0 = time.perf_counter()
    local, proto = run()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    rep_ms.append((time.perf_counter() - t0) * 1e3)
ms = min(rep_ms)  # every repetition is barrier-to-barrier; the best one is reported, all are listed
ok = True
if rank == 0:
    full = dev.empty_elems(n * L)
    dev.lde(d_coeffs, log_n, log_f, True, full, fid)
    ref = dev.fri_commit(full, L, 1, fid)
    ok = (proto.roots == ref.get_roots() and np.array_equal(np.stack(proto.challenges), ref.challenges)
          and np.array_equal(proto.final_coefficients, ref.final_coefficients)
          and torch.equal(local, full[rank::world]))
    real_stdout.write(json.dumps({"check": "sharded coset LDE + FRI commit == single-GPU chain", "ok": bool(ok), "n_gpus": world,
                                  "log_n": log_n, "lde_factor": L, "domain": n * L, "layers": proto.num_steps,
                                  "ms_lde_plus_fri": ms, "ms_all_reps": rep_ms, "lde_elems_per_s": n * L / (ms * 1e-3)}) + "\n")
    real_stdout.flush()
if world > 1:
    dist.destroy_process_group()
sys.exit(0 if ok else 1)
