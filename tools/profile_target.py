"""Small driver for ncu captures: a few passes of each hot kernel at the benchmark sizes.
    ncu ... python tools/profile_target.py [lde|merkle|fri|ntt|commit]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import hodor_b200 as H
from hodor_b200 import device as dev

what = sys.argv[1] if len(sys.argv) > 1 else "lde"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
H.init(0)
rng = np.random.default_rng(1)
n = 1 << 24
a = rng.integers(0, 2**64, size=(n, 4), dtype=np.uint64)
a[:, 3] = rng.integers(0, 0x73EDA753299D7D48, size=n, dtype=np.uint64)
d_a = dev.to_device(a)
if what == "lde":
    d_out = dev.empty_elems(n * 8)
    for _ in range(reps):
        dev.lde(d_a, 24, 3, True, d_out, 0)
elif what == "ntt":
    d_out = dev.empty_elems(n)
    for _ in range(reps):
        dev.fft(d_a, d_out, 24, False, 0)
elif what == "merkle":
    d_nodes = dev.empty_elems(n)
    for _ in range(reps):
        dev.merkle_build(d_a, n, d_nodes, 0)
elif what == "commit":  # lift-and-commit of one polynomial; HODOR_FUSE_LAST_COMMIT=1 -> ntt_last_commit_kernel
    import ctypes as C
    from hodor_b200 import _ffi
    for _ in range(reps):
        ins, outs = (C.c_void_p * 1)(d_a.data_ptr()), (C.c_void_p * 1)()
        _ffi.check(_ffi.lib.hodor_cuda_lde_commit_batch(ins, 1, 24, 3, 1, 1, outs, None, 0))
        _ffi.lib.hodor_cuda_tree_free(outs[0])
elif what == "fri":
    for _ in range(reps):
        p = dev.fri_commit(d_a, 8, 1, 0)
        p.free()
torch.cuda.synchronize()
print("done", what)
