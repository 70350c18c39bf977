"""Parity of lift-and-commit with the tree's bottom three levels hashed inside the last pass of the transform
(hodor_b200/csrc/ntt_commit.cuh) against the CPU oracle: values and every node, over every last-pass width
(6, 7, 8), two- to four-pass plans, blowups 1 .. 16, plain and coset, three fields.
    HODOR_FUSE_LAST_COMMIT=3 python tools/fused_commit_check.py      # 0: never fused, 1 (default): last digit 8, 2: also 7, 3: also 6
Prints one JSON line per case and a summary; exit code 1 on any mismatch, or if a case did not run the kernel its
plan and the level call for (the switch is read at hodor_cuda_init)."""
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import hodor_b200 as H
from hodor_b200 import _ffi
from oracle import oracle as O  # the checker

H.init(0)
level = int(os.environ.get("HODOR_FUSE_LAST_COMMIT", "1") or "1")  # 0 off, 1 (default) last digit 8, 2 also 7, 3 also 6


def last_digit(log_n, max_digit=8):
    """context.h make_plan: digits balanced in 6..max_digit, most significant first."""
    passes = -(-log_n // max_digit)
    if log_n // passes < 6 and passes > 2:
        passes -= 1
    base, rem = divmod(log_n, passes)
    return base + (1 if passes - 1 < rem else 0)


#        field, log_n, L, coset      plan (digits, last one is the fused kernel's width)
CASES = [(0, 12, 8, True),         # 6+6
         (0, 13, 1, False),        # 7+6, plain NTT commit: 8 adjacent outputs as the columns
         (0, 13, 2, True),         # 7+6, two cosets
         (1, 13, 4, False),
         (2, 14, 8, True),         # 7+7
         (0, 15, 16, True),        # 8+7, more cosets than columns
         (0, 16, 8, True),         # 8+8
         (1, 16, 8, False),
         (2, 16, 16, True),
         (0, 17, 4, True),         # 9+8 (digits below 6 are not built)
         (0, 18, 8, True),         # 6+6+6
         (0, 20, 8, True)]         # 7+7+6, expanded tables
max_log = int(os.environ.get("FUSED_CHECK_MAX_LOG", "99"))  # compute-sanitizer runs: keep the small shapes only
CASES = [c for c in CASES if c[1] <= max_log]
ok_all = True
for fid, log_n, L, coset in CASES:
    a = O.random_elements(fid, 1 << log_n, seed=7000 + 31 * log_n + L)
    _ffi.check(_ffi.lib.hodor_cuda_profile_begin())
    orc = H.CommittedOracle.lde_commit(H.Polynomial.from_coeffs(fid, a), L, coset)
    buf = C.create_string_buffer(1 << 16)
    _ffi.check(_ffi.lib.hodor_cuda_profile_end(buf, len(buf)))
    kernels = {r["name"]: r["count"] for r in json.loads(buf.value.decode())}
    lde = O.lde(fid, a, log_n, L, coset) if L > 1 else (O.fft(fid, a, log_n, coset=coset))
    nodes = O.merkle_create(fid, lde)
    fused = "ntt_pass_last_commit" in kernels
    want_fused = level > 0 and 9 - level <= last_digit(log_n) <= 8
    ok = bool(np.array_equal(orc.values(), lde) and np.array_equal(orc.nodes, nodes) and orc.get_root() == nodes[1].tobytes())
    ok = ok and fused == want_fused and ("merkle_levels_leaf" in kernels) != fused
    q = orc.query((1 << log_n) * L - 3)
    ok = ok and H.TrivialBlake2sIOP.verify_query(q, orc.get_root())
    orc.free()
    ok_all = ok_all and ok
    print(json.dumps({"field": fid, "log_n": log_n, "lde_factor": L, "coset": coset, "last_digit": last_digit(log_n), "fused": fused, "ok": ok,
                      "kernels": kernels}), flush=True)
print(json.dumps({"check": "lift-and-commit == oracle, last pass fused with the tree's bottom levels as the level says", "level": level,
                  "cases": len(CASES), "ok": ok_all}))
sys.exit(0 if ok_all else 1)
