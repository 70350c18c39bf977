"""Parity of the FRI commit chain with the fold fused with the bottom three levels of the next tree
(csrc/fri.cuh fri_fold_commit_kernel, HODOR_FUSE_FOLD_COMMIT) against the CPU oracle: every root, challenge, final
coefficient, every layer's values and nodes, over three fields, blowups 2..16 and 1..4 output coefficients.
    HODOR_FUSE_FOLD_COMMIT=1 python tools/fri_fused_check.py
Prints one JSON line per case and a summary; exit code 1 on any mismatch, or if the fused kernel ran (did not run)
against what the switch says (it is read at hodor_cuda_init)."""
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
want_fused = os.environ.get("HODOR_FUSE_FOLD_COMMIT", "0") not in ("", "0")
CASES = [(0, 13, 8, 1), (0, 14, 2, 1), (1, 14, 16, 2), (2, 15, 8, 4), (0, 16, 4, 1), (2, 16, 16, 1), (0, 18, 8, 1)]
max_log = int(os.environ.get("FUSED_CHECK_MAX_LOG", "99"))
ok_all = True
for fid, log_n, L, out in [c for c in CASES if c[1] <= max_log]:
    n = 1 << log_n
    vals = O.random_elements(fid, n, seed=8100 + 17 * log_n + L)
    want = O.fri_commit(fid, vals, L, out)
    _ffi.check(_ffi.lib.hodor_cuda_profile_begin())
    proto = H.NaiveFriIop.proof_from_lde(H.Polynomial.from_values(fid, vals), L, out, H.Worker())
    buf = C.create_string_buffer(1 << 16)
    _ffi.check(_ffi.lib.hodor_cuda_profile_end(buf, len(buf)))
    kernels = {r["name"]: r["count"] for r in json.loads(buf.value.decode())}
    fused = "fri_fold_commit" in kernels
    ok = proto.get_roots() == want.roots() and np.array_equal(proto.challenges, want.challenges)
    ok = ok and np.array_equal(proto.final_coefficients, want.final_coefficients)
    ok = ok and np.array_equal(proto.l0_commitment.nodes, want.l0_nodes)
    for i, (iop, v) in enumerate(zip(proto.intermediate_commitments, proto.intermediate_values)):
        ok = ok and np.array_equal(iop.nodes, want.layer_nodes[i]) and np.array_equal(v.as_ref(), want.layer_values[i])
    ok = bool(ok) and fused == want_fused
    proto.free()
    ok_all = ok_all and ok
    print(json.dumps({"field": fid, "log_n": log_n, "lde_factor": L, "out_coeffs": out, "fused": fused, "ok": ok, "kernels": kernels}),
          flush=True)
print(json.dumps({"check": "FRI commit chain == oracle", "want_fused": want_fused, "ok": ok_all}))
sys.exit(0 if ok_all else 1)
