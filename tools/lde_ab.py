"""Same-session A/B of the headline transforms: coset LDE 2^24 x 8 and forward NTT 2^24, ms per call and the
per-kernel profile.  The variant is whatever library HODOR_B200_LIB points at (default: the in-tree build):
    python tools/lde_ab.py
    HODOR_B200_LIB=build/libhodor_b200_pf.so python tools/lde_ab.py        # -DHODOR_PASS_PREFETCH=1"""
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import hodor_b200 as H
from hodor_b200 import _ffi
from hodor_b200 import device as dev

H.init(0)
rng = np.random.default_rng(1)
n = 1 << 24
a = rng.integers(0, 2**64, size=(n, 4), dtype=np.uint64)
a[:, 3] = rng.integers(0, 0x73EDA753299D7D48, size=n, dtype=np.uint64)
d_a = dev.to_device(a)
d_out = dev.empty_elems(n * 8)
d_tmp = dev.empty_elems(n)


def timed(fn, reps=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def profile(fn):
    _ffi.check(_ffi.lib.hodor_cuda_profile_begin())
    fn()
    buf = C.create_string_buffer(1 << 16)
    _ffi.check(_ffi.lib.hodor_cuda_profile_end(buf, len(buf)))
    return {r["name"]: round(r["total_ms"] / r["count"], 4) for r in json.loads(buf.value.decode())}


lde = lambda: dev.lde(d_a, 24, 3, True, d_out, 0)  # noqa: E731
ntt = lambda: dev.fft(d_a, d_tmp, 24, False, 0)  # noqa: E731
print(json.dumps({"bench": "lde_ntt_ab", "lib": os.environ.get("HODOR_B200_LIB", "in-tree"), "lde_2p24_x8_ms": timed(lde),
                  "ntt_2p24_ms": timed(ntt), "lde_kernels_ms": profile(lde), "ntt_kernels_ms": profile(ntt)}))
