#!/bin/bash
# forward NTT time per size under both pass-plan policies (HODOR_NTT_MAX_DIGIT = 8 | 9)
for md in 9 8; do
HODOR_NTT_MAX_DIGIT=$md python - <<PY
import os, sys, json
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import hodor_b200 as H
from hodor_b200 import device as dev
H.init(0)
rng = np.random.default_rng(1)
res = {}
for ln in (17, 18, 19, 22, 24, 25, 26, 27, 28):
    n = 1 << ln
    a = rng.integers(0, 2**64, size=(n, 4), dtype=np.uint64); a[:, 3] = rng.integers(0, 0x73EDA753299D7D48, size=n, dtype=np.uint64)
    d = dev.to_device(a); o = dev.empty_elems(n)
    for _ in range(2): dev.fft(d, o, ln, False, 0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 5
    e0.record()
    for _ in range(reps): dev.fft(d, o, ln, False, 0)
    e1.record(); torch.cuda.synchronize()
    res[ln] = round(e0.elapsed_time(e1) / reps, 4)
    del d, o
print(json.dumps({"max_digit": int(os.environ["HODOR_NTT_MAX_DIGIT"]), "ntt_ms": res}))
PY
done
