#!/usr/bin/env python3
"""Times the local phases of the four-step sharded NTT on ONE GPU (no exchange), so that what the exchange costs on a
multi-GPU box can be read off as (measured sharded time) - (step A + step B):

    python tools/sharded_phases.py [log_n log_g ...]      # default: 28 3  26 3  28 1  24 3

Per (log_n, log_g): the plain NTT of the local length 2^(log_n - log_g), step A (the same transform with the per-rank
output scaling fused into its last pass), step B (the G-point DFT across the received slices), plus the per-launch
times of the library's own profiler for step A."""
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import hodor_b200 as H
from hodor_b200 import device as dev
from hodor_b200._ffi import lib
from hodor_b200.sharded import CudaBackend


def timed(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    out = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        out.append(a.elapsed_time(b))
    return min(out)


def main():
    args = [int(x) for x in sys.argv[1:]] or [28, 3, 26, 3, 28, 1, 24, 3]
    H.init(0)
    fid = 0
    be = CudaBackend()
    for log_n, log_g in zip(args[::2], args[1::2]):
        m = 1 << (log_n - log_g)
        omega = H.Domain.new_for_size(fid, 1 << log_n).generator
        omega_m = H.Domain.new_for_size(fid, m).generator
        g = torch.Generator(device="cuda").manual_seed(log_n)
        src = torch.randint(0, 1 << 62, (m, 4), dtype=torch.int64, device="cuda", generator=g)
        dst = torch.empty_like(src)
        plain = timed(lambda: dev.ntt(src, dst, log_n - log_g, omega_m, fid))
        step_a = timed(lambda: be.shard_cols(src, log_n, log_g, 1, omega, fid))
        step_b = timed(lambda: be.shard_rows(src, log_n, log_g, 1, omega, fid))
        lib.hodor_cuda_profile_begin()
        be.shard_cols(src, log_n, log_g, 1, omega, fid)
        buf = C.create_string_buffer(1 << 16)
        lib.hodor_cuda_profile_end(buf, len(buf))
        try:
            launches = json.loads(buf.value.decode())
        except Exception:
            launches = buf.value.decode()[:400]
        print(json.dumps({"log_n": log_n, "log_g": log_g, "local_len_log2": log_n - log_g, "plain_ntt_ms": plain, "step_a_ms": step_a,
                          "step_b_ms": step_b, "a_plus_b_ms": step_a + step_b, "step_a_launches": launches}), flush=True)
        del src, dst


if __name__ == "__main__":
    main()
