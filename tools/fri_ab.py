"""Same-session A/B of the FRI commit chain on 2^24 values (blowup 8): prints ms per chain and the per-kernel
profile.  Variants are selected through the library's environment switches, one process each:
    HODOR_FUSE_FOLD_COMMIT=0 python tools/fri_ab.py      # fold and leaf hashing as separate kernels
    HODOR_FUSE_FOLD_COMMIT=1 python tools/fri_ab.py      # fold fused with the bottom three levels of the next tree"""
import ctypes as C
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import hodor_b200 as H
from hodor_b200 import _ffi
from hodor_b200 import device as dev

log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 24
L = int(sys.argv[2]) if len(sys.argv) > 2 else 8
H.init(0)
rng = np.random.default_rng(1)
a = rng.integers(0, 2**64, size=(1 << log_n, 4), dtype=np.uint64)
a[:, 3] = rng.integers(0, 0x73EDA753299D7D48, size=1 << log_n, dtype=np.uint64)
d = dev.to_device(a)
for _ in range(3):
    dev.fri_commit(d, L, 1, 0).free()
torch.cuda.synchronize()
reps = 10
t0 = time.perf_counter()
for _ in range(reps):
    dev.fri_commit(d, L, 1, 0).free()
torch.cuda.synchronize()
ms = (time.perf_counter() - t0) * 1e3 / reps
_ffi.check(_ffi.lib.hodor_cuda_profile_begin())
dev.fri_commit(d, L, 1, 0).free()
buf = C.create_string_buffer(1 << 16)
_ffi.check(_ffi.lib.hodor_cuda_profile_end(buf, len(buf)))
prof = {r["name"]: {"launches": r["count"], "total_ms": round(r["total_ms"], 4)} for r in json.loads(buf.value.decode())}
print(json.dumps({"bench": "fri_chain", "log_n": log_n, "lde_factor": L, "fuse_fold_commit": os.environ.get("HODOR_FUSE_FOLD_COMMIT", "default"),
                  "tail_max": os.environ.get("HODOR_MERKLE_TAIL_MAX", "default"), "ms_per_chain": ms, "kernels": prof}))
