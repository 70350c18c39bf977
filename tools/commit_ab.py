"""Same-session A/B of lift-and-commit (hodor_cuda_lde_commit_batch, 2^24 x 8, 8 polynomials per call, pinned host
coefficients): ms per polynomial.  Variants through the library's environment switches, one process each:
    HODOR_CONCURRENT_COMMIT=0 python tools/commit_ab.py   # tree of polynomial i after its LDE, on one stream
    python tools/commit_ab.py                             # tree of polynomial i beside the LDE of polynomial i+1
    HODOR_FUSE_LAST_COMMIT=0 python tools/commit_ab.py    # never hash the bottom three tree levels inside the last pass (ntt_commit.cuh; 3: always)"""
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

H.init(0)
lib = _ffi.lib
log_n = int(os.environ.get("COMMIT_AB_LOG_N", "24"))
n, count = 1 << log_n, 8
rng = np.random.default_rng(1)
a = rng.integers(0, 2**64, size=(n, 4), dtype=np.uint64)
a[:, 3] = rng.integers(0, 0x73EDA753299D7D48, size=n, dtype=np.uint64)
h_in = torch.empty((n, 4), dtype=torch.int64, pin_memory=True)
h_in.numpy().view(np.uint64)[:] = a


on_device = os.environ.get("COMMIT_AB_DEVICE", "0") == "1"  # coefficients already in HBM: no H2D in the pipeline
count = int(os.environ.get("COMMIT_AB_COUNT", count))         # 1: no neighbouring polynomial to overlap with
d_in = h_in.cuda() if on_device else None
call_s = 0.0  # time inside the batch call alone (the handles are freed outside it)


def chunk():
    global call_s
    src = d_in.data_ptr() if on_device else h_in.data_ptr()
    ins = (C.c_void_p * count)(*[src] * count)
    outs = (C.c_void_p * count)()
    roots = np.zeros((count, 32), np.uint8)
    t = time.perf_counter()
    _ffi.check(lib.hodor_cuda_lde_commit_batch(ins, count, log_n, 3, 1, int(on_device), outs, roots.ctypes.data_as(_ffi.u8p), 0))
    call_s += time.perf_counter() - t
    for i in range(count):
        lib.hodor_cuda_tree_free(outs[i])
    return roots


r0 = chunk()
call_s = 0.0
t0 = time.perf_counter()
reps = (3 if count > 1 else 12) * (1 if log_n >= 24 else 4)
for _ in range(reps):
    r = chunk()
ms = (time.perf_counter() - t0) * 1e3 / (reps * count)
ms_call = call_s * 1e3 / (reps * count)
assert all(r[i].tobytes() == r0[0].tobytes() for i in range(count))
_ffi.check(lib.hodor_cuda_profile_begin())  # per-kernel times of one more chunk (events around every launch)
chunk()
buf = C.create_string_buffer(1 << 16)
_ffi.check(lib.hodor_cuda_profile_end(buf, len(buf)))
kernels = {k["name"]: round(k["total_ms"] / count, 4) for k in json.loads(buf.value.decode())}
print(json.dumps({"bench": f"lde_commit_batch 2^{log_n} x 8", "concurrent_commit": os.environ.get("HODOR_CONCURRENT_COMMIT", "1"),
                  "fuse_last_commit": os.environ.get("HODOR_FUSE_LAST_COMMIT", "1 (default)"), "kernel_ms_per_polynomial": kernels,
                  "commit_priority": os.environ.get("HODOR_COMMIT_PRIORITY", "low"),
                  "backfill_persist": os.environ.get("HODOR_BACKFILL_PERSIST", "0"),
                  "backfill_block": os.environ.get("HODOR_BACKFILL_BLOCK", "128"),
                  "ms_per_polynomial": ms, "ms_per_polynomial_inside_the_call": ms_call, "polynomials_per_call": count,
                  "coefficients": "device" if on_device else "pinned host", "root": r0[0].tobytes().hex()}))
