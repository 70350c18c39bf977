for tm in 4096 1024 512 256 128; do
HODOR_MERKLE_TAIL_MAX=$tm python - <<PY
import os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import hodor_b200 as H
from hodor_b200 import device as dev
H.init(0)
rng = np.random.default_rng(1); n = 1 << 24
a = rng.integers(0, 2**64, size=(n, 4), dtype=np.uint64); a[:, 3] = rng.integers(0, 0x73EDA753299D7D48, size=n, dtype=np.uint64)
d = dev.to_device(a)
for _ in range(3):
    p = dev.fri_commit(d, 8, 1, 0); p.free()
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(10):
    p = dev.fri_commit(d, 8, 1, 0); p.free()
torch.cuda.synchronize(); print("tail_max", os.environ["HODOR_MERKLE_TAIL_MAX"], "fri ms", (time.perf_counter() - t0) * 100)
PY
done
