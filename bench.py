#!/usr/bin/env python
"""bench.py -- headline benchmark of the hodor_b200 hot path (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--sweep]

One "step" is one pass of the hot path over one batch of synthetic input: the coset LDE of a
2^24-coefficient polynomial over the `src/bn256.rs` field (= BLS12-381 Fr) with blowup 8
(BASELINE.json configs[1]: eight coset NTTs of 2^24 -> 2^27 values, 4 GiB), inputs resident in HBM.
The same line also carries the second half of BASELINE.json's metric, the FRI commit chain on a
2^24 domain (configs[2], Merkle leaves/s), a plain 2^24 NTT, the end-to-end number through the
host-pointer C ABI, the live roofline of the dominant kernel and the CPU baseline.

Under torchrun (N > 1) every rank runs the same per-GPU workload on its own polynomial (weak
scaling, no data-path collective -- registers / polynomials of a proof are independent), and the
four-step sharded NTT (the one path with a real exchange step, NCCL all-to-all) is timed as an
extra section.  Rank 0 prints ONE JSON line.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

FIELD = 0  # BLS12-381 Fr == the reference's src/bn256.rs
LOG_N = 24
LOG_L = 3
FRI_LOG_N = 24
FRI_L = 8
FRI_OUT = 1
WORKLOAD = (f"coset LDE 2^{LOG_N} -> 2^{LOG_N + LOG_L} (blowup {1 << LOG_L}: {1 << LOG_L} coset NTTs of 2^{LOG_N}) over "
            "bn256.rs Fr (= BLS12-381 Fr), natural order in/out; per GPU")
P_TOP_LIMB = 0x73EDA753299D7D48  # top u64 limb of the modulus: any element with a smaller top limb is canonical


_REAL_STDOUT = None


def emit(line: dict) -> None:
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def synthetic_elements(count: int, seed: int) -> np.ndarray:
    """i.i.d. canonical field elements used directly as Montgomery representations."""
    rng = np.random.default_rng(seed)
    a = rng.integers(0, 2**64, size=(count, 4), dtype=np.uint64)
    a[:, 3] = rng.integers(0, P_TOP_LIMB, size=count, dtype=np.uint64)
    return a


class ClockSampler:
    """Samples SM clock and throttle reasons while a timed region runs (pynvml, 100 ms period)."""

    REASONS = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
               0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}

    def __init__(self, index: int):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        while not self._stop.is_set():
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                mask = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for bit, name in self.REASONS.items():
                    if mask & bit and name != "gpu_idle":
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.1)

    def __enter__(self):
        if self.nv:
            self._stop.clear()
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()
        return self

    def __exit__(self, *exc):
        if self._thread:
            self._stop.set()
            self._thread.join()
            self._thread = None

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}
        return {"sm_mhz": statistics.median(self.samples), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


# --------------------------------------------------------------------------------------------------
# reference arm: the reference's own CPU algorithm (restated in oracle/, see its header) on the host cores
# --------------------------------------------------------------------------------------------------
def cpu_lde_sample(log_n: int, repeats: int):
    from oracle import oracle as O
    O.build()
    cores = O.default_cpus()
    coeffs = O.random_elements(FIELD, 1 << log_n, seed=11)
    O.lde(FIELD, coeffs[: 1 << 10], 10, 1 << LOG_L, True, cpus=cores)  # warm the library
    best = None
    for _ in range(repeats):
        t0 = time.perf_counter()
        O.lde(FIELD, coeffs, log_n, 1 << LOG_L, True, cpus=max(cores, 1 << LOG_L))
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return ((1 << log_n) << LOG_L) / best, cores, best


def cpu_fri_sample(log_n: int):
    from oracle import oracle as O
    cores = O.default_cpus()
    vals = O.random_elements(FIELD, 1 << log_n, seed=12)
    t0 = time.perf_counter()
    O.fri_commit(FIELD, vals, FRI_L, FRI_OUT, cpus=cores)
    dt = time.perf_counter() - t0
    steps = O.fri_num_steps(1 << log_n, FRI_L, FRI_OUT)
    leaves = sum((1 << log_n) >> i for i in range(steps + 1))
    return leaves / dt, cores, dt


def bench_config(world: int) -> dict:
    """`config` of the JSON line -- the SAME object for both arms (the driver compares them)."""
    return {"workload": WORKLOAD, "field": "bls12_381_fr", "log_n": LOG_N, "lde_factor": 1 << LOG_L, "passes": 3,
            "l2_policy": "inputs (512 MiB) and outputs (4 GiB) exceed the 126 MB L2; no flush between steps",
            "parallelism": f"{world} independent polynomials, one per GPU" if world > 1 else "single GPU"}


def run_reference(args, rank: int):
    """The reference's own CPU algorithm for the step (coset_lde_using_multiple_cosets, src/polynomials/mod.rs:
    544-609, restated in oracle/hodor_oracle.c -- the Rust crate cannot be built in this image) on the host cores,
    at the FULL size of the GPU arm's step: every timed step is one 2^24 -> 2^27 coset LDE.  One warm-up step is
    full size, the others run a 2^20 sample (warming a CPU path needs no more).  HODOR_REF_BUDGET_S (default 900)
    bounds the run: if the first full step projects past it, the arm says so and times the 2^20 sample instead."""
    if rank != 0:
        return
    from oracle import oracle as O
    O.build()
    cores = O.default_cpus()
    L = 1 << LOG_L
    threads = max(cores, L)
    budget = float(os.environ.get("HODOR_REF_BUDGET_S", "900"))
    small = O.random_elements(FIELD, 1 << 20, seed=11)
    t0 = time.perf_counter()
    O.lde(FIELD, small, 20, L, True, cpus=threads)
    small_dt = time.perf_counter() - t0
    log_n = LOG_N
    if small_dt * 16 * 1.3 * (args.steps + 1) > budget:
        log_n = 20  # a host this slow cannot run `steps` full-size steps inside the budget
    coeffs = O.random_elements(FIELD, 1 << log_n, seed=11) if log_n != 20 else small
    for i in range(args.warmup):
        if i == 0:
            O.lde(FIELD, coeffs, log_n, L, True, cpus=threads)
        else:
            O.lde(FIELD, small, 20, L, True, cpus=threads)
    times = []
    for i in range(args.steps):
        t0 = time.perf_counter()
        out = O.lde(FIELD, coeffs, log_n, L, True, cpus=threads)
        times.append(time.perf_counter() - t0)
    elems = (1 << log_n) << LOG_L
    total = sum(times)
    value = elems * len(times) / total
    # the committed form of the step (what `e2e` of the GPU arm does on top of the LDE): Blake2sIopTree::create
    t0 = time.perf_counter()
    O.merkle_create(FIELD, out, cpus=cores)
    tree_dt = time.perf_counter() - t0
    del out
    fri_value, _, fri_dt = cpu_fri_sample(FRI_LOG_N if log_n == LOG_N else 20)
    full = log_n == LOG_N
    sample = (f"every timed step = the full coset LDE 2^{log_n} x{L} ({'the GPU arm\'s workload' if full else 'a 1/16 SAMPLE: host too slow for the budget'}); "
              f"multi-coset path as the reference runs it on {cores} cores: one serial_fft thread per coset "
              f"(its num_cpus_hint rule, src/polynomials/mod.rs:549-558), distribute_powers and the interleave on all cores; "
              f"2^20 sample for comparison: {small_dt:.2f} s")
    cfg = bench_config(args.gpus)
    if not full:
        cfg["reference_sample_log_n"] = log_n
    line = {
        "impl": "reference", "metric": "ntt_field_elems_per_sec", "value": value, "unit": "field-elems/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u256 (4 x u64 Montgomery)",
        "data": "synthetic", "config": cfg,
        "cpu_baseline": {"value": value, "unit": "field-elems/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "field-elems/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "reference_impl": "C restatement of hodor's crossbeam path (Rust toolchain unavailable): oracle/hodor_oracle.c",
        "lde_commit": {"workload": "the same LDE followed by Blake2sIopTree::create on its 2^27 values (prover/mod.rs:73-80)",
                       "tree_s": tree_dt, "lde_s": total / len(times),
                       "value": elems / (total / len(times) + tree_dt), "unit": "field-elems/s"},
        "fri": {"metric": "fri_merkle_leaves_per_sec", "value": fri_value, "unit": "leaves/s",
                "sample": f"FRI commit chain on 2^{FRI_LOG_N if full else 20} values, blowup {FRI_L}, {cores} cores, {fri_dt:.2f} s"},
        "gpu_launches": 0,
    }
    emit(line)


# --------------------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------------------
def gpu_local_cpus(local_rank: int):
    """CPUs of the NUMA node the GPU hangs off (sysfs), or None.  Pinned staging buffers allocated from a
    thread bound there are local to the GPU's PCIe root: with one rank per GPU and unbound processes the
    buffers of all ranks can land on one socket and every copy then crosses the inter-socket link."""
    import torch
    try:
        pr = torch.cuda.get_device_properties(local_rank)
        bdf = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        text = open(f"/sys/bus/pci/devices/{bdf}/local_cpulist").read().strip()
        cpus = set()
        for part in text.split(","):
            if "-" in part:
                a, b = part.split("-")
                cpus.update(range(int(a), int(b) + 1))
            elif part:
                cpus.add(int(part))
        cpus &= os.sched_getaffinity(0)
        node = open(f"/sys/bus/pci/devices/{bdf}/numa_node").read().strip()
        return (cpus, node) if cpus else None
    except Exception:
        return None


def run_fib(args, dev, lib, _ffi, profile, cpu: bool, log_rows: int = 20, lde_factor: int = 16):
    import torch
    from hodor_b200 import fib_replay as R

    t0 = time.perf_counter()
    a, b = R.fibonacci_witness(FIELD, 1 << log_rows)
    witness_s = time.perf_counter() - t0
    t0 = time.perf_counter()
    prover = R.FibonacciProver(FIELD, log_rows, lde_factor, 1)
    setup_ms = (time.perf_counter() - t0) * 1e3
    prover.prove(a, b)  # warm-up: tables, pool blocks
    times = []
    for _ in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        pr = prover.prove(a, b)
        times.append((time.perf_counter() - t0) * 1e3)
    l0 = dev.launch_count()
    prof = profile(lambda: prover.prove(a, b))
    launches = dev.launch_count() - l0
    ksum = sum(r["total_ms"] for r in prof.values())
    ok = all(hodor_verify(q, root) for q, root in zip(pr.f_queries + [pr.g_query], pr.f_iop_roots + [pr.g_iop_root]))
    out = {"workload": f"Fibonacci AIR prove, trace 2^{log_rows} rows x 2 registers, blowup {lde_factor}, FRI to 1 coefficient: "
                       "2 iNTT, 3 LDE 2^20 -> 2^24 + 3 trees, 7 coset NTT + 1 icoset NTT, DEEP (5 evaluate_at, 3 batch inversions "
                       "over 2^24, elementwise passes), 2 FRI commit chains (21 trees each), queries",
           "ms_per_proof": min(times), "ms_all": times, "setup_ms (Prover::new: ALI divisors on device)": setup_ms,
           "witness_generation_s (host, not timed)": witness_s, "kernel_ms": ksum, "gpu_launches": launches,
           "h2d_bytes_per_proof": int(a.nbytes + b.nbytes), "openings_verify": bool(ok),
           "kernels": {k: {"launches": r["count"], "total_ms": round(r["total_ms"], 3)} for k, r in sorted(prof.items(), key=lambda kv: -kv[1]["total_ms"])},
           "timer": "host perf_counter around prove(): witness H2D, every kernel, transcript on the host, openings D2H",
           "parity": "bit-exact against the big-int model of the same call sequence at 2^2 / 2^5 / 2^8 rows (tests/test_fib_prove.py)"}
    if cpu:
        # the same proof's hot-path components on the host cores (the reference's algorithms, oracle/): a LOWER bound
        # of its prove() time -- the elementwise / DEEP passes and the AIR machinery are not included
        from oracle import oracle as O
        O.build()
        cores = O.default_cpus()
        n, N = 1 << log_rows, (1 << log_rows) * lde_factor
        x = O.random_elements(FIELD, n, seed=5)
        t0 = time.perf_counter()
        for _ in range(2):
            O.ifft(FIELD, x, log_rows)
        lde = None
        for _ in range(3):
            lde = O.lde(FIELD, x, log_rows, lde_factor, False, cpus=max(cores, lde_factor))
            O.merkle_create(FIELD, lde, cpus=cores)
        for _ in range(8):
            O.fft(FIELD, x, log_rows, coset=True)
        for _ in range(2):
            O.fri_commit(FIELD, lde, lde_factor, 1, cpus=cores)
        cpu_s = time.perf_counter() - t0
        out["cpu_hot_path_components_s"] = cpu_s
        out["cpu_cores"] = cores
        out["cpu_note"] = ("2 iNTT + 3 x (LDE + Merkle tree) + 8 coset NTT + 2 FRI chains with the reference's CPU algorithms "
                           "(oracle/hodor_oracle.c); a lower bound of the reference's prove(): DEEP / elementwise passes excluded")
        out["speedup_vs_cpu_lower_bound"] = cpu_s * 1e3 / min(times)
    del prover
    _ffi.check(lib.hodor_cuda_trim())
    return out


def hodor_verify(query, root) -> bool:
    import hodor_b200 as H
    return bool(H.TrivialBlake2sIOP.verify_query(query, root))


def run_ours(args, rank: int, local_rank: int, world: int):
    import torch
    import torch.distributed as dist

    import hodor_b200 as H
    from hodor_b200 import _ffi
    from hodor_b200 import device as dev
    from hodor_b200.field import _p

    torch.cuda.set_device(local_rank)
    H.init(local_rank)
    lib = _ffi.lib
    peak_gbs, peak_src = load_peaks()
    n, L = 1 << LOG_N, 1 << LOG_L

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms: float) -> float:
        if world == 1:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = dev.launch_count()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        timed.launches = dev.launch_count() - l0  # kernels of this library inside the timed region
        ms = e0.elapsed_time(e1)
        barrier()
        return max_over_ranks(ms)

    def profile(fn):
        _ffi.check(lib.hodor_cuda_profile_begin())
        fn()
        buf = C.create_string_buffer(1 << 16)
        _ffi.check(lib.hodor_cuda_profile_end(buf, len(buf)))
        return {r["name"]: r for r in json.loads(buf.value.decode())}

    # ---- inputs, resident in HBM before any timed region (larger than the 126 MB L2: no flush needed)
    coeffs = synthetic_elements(n, seed=1000 + rank)
    d_coeffs = dev.to_device(coeffs)
    d_out = dev.empty_elems(n * L)

    def lde_step():
        dev.lde(d_coeffs, LOG_N, LOG_L, True, d_out, FIELD)

    clocks = ClockSampler(local_rank)
    with clocks:
        total_ms = timed(lde_step, args.steps, args.warmup)
    launches_timed = timed.launches
    ms_per_step = total_ms / args.steps
    value = world * n * L / (ms_per_step * 1e-3)

    # ---- live per-kernel share and roofline of the dominant kernel (events around every launch)
    prof = profile(lambda: [lde_step() for _ in range(args.steps)])
    kern_total = sum(r["total_ms"] for r in prof.values())
    dom = max(prof.values(), key=lambda r: r["total_ms"])
    alg_bytes = {"ntt_pass_first_scaled": 32 * n + 32 * n * L, "ntt_pass": 64 * n * L, "ntt_pass_last": 64 * n * L}
    dom_ms = dom["total_ms"] / dom["count"]
    dom_bytes = alg_bytes.get(dom["name"], 64 * n * L)
    achieved = dom_bytes / (dom_ms * 1e-3) / 1e9
    # DRAM traffic of that kernel per launch from the committed `ncu --set full` capture (profiles/)
    traffic, pipes = None, None
    try:
        cap = json.load(open(os.path.join(ROOT, "profiles", "r02_traffic.json")))[dom["name"]]
        traffic = cap["traffic"]
        pipes = {"fmaheavy_pipe_cycles_active_pct": cap["fmaheavy_pipe_cycles_active_pct"],
                 "alu_pipe_cycles_active_pct": cap["alu_pipe_cycles_active_pct"], "issue_active_pct": cap["issue_active_pct"],
                 "source": "profiles/r02_traffic.json (ncu --set full of the same kernel; tools/capture_profiles.sh)"}
    except Exception:
        pass
    roofline = {
        "bound": "hbm", "kernel": dom["name"], "achieved": achieved, "peak": peak_gbs, "unit": "GB/s",
        "frac": achieved / peak_gbs, "traffic": traffic, "int_pipes": pipes, "peak_source": peak_src,
        "bytes_per_launch": dom_bytes, "ms_per_launch": dom_ms, "share_of_step": dom["total_ms"] / kern_total,
        "kernels": {k: {"launches": r["count"], "ms_per_launch": r["total_ms"] / r["count"],
                        "share": r["total_ms"] / kern_total} for k, r in prof.items()},
        "job": {"algorithmic_bytes": 32 * n + 32 * n * L, "achieved": (32 * n + 32 * n * L) / (ms_per_step * 1e-3) / 1e9,
                "frac": (32 * n + 32 * n * L) / (ms_per_step * 1e-3) / 1e9 / peak_gbs},
        "note": "INT32 multiplier-pipe bound by design: ~13 256-bit modular multiplies per 64 B moved (ncu: fmaheavy pipe "
                "72-78 % busy, DRAM 11-22 %); traffic exceeds the algorithmic bytes because every multiply streams a "
                "64-byte precomputed table entry instead of spending a second multiplication; see DESIGN.md",
    }

    # ---- plain forward NTT 2^24 (BASELINE metric, first half)
    d_tmp = dev.empty_elems(n)
    ntt_ms = timed(lambda: dev.fft(d_coeffs, d_tmp, LOG_N, False, FIELD), args.steps, args.warmup) / args.steps
    ntt = {"metric": "ntt_field_elems_per_sec", "workload": f"forward NTT 2^{LOG_N}", "value": world * n / (ntt_ms * 1e-3),
           "unit": "field-elems/s", "ms_per_step": ntt_ms, "algorithmic_bytes": 64 * n,
           "hbm_frac": 64 * n / (ntt_ms * 1e-3) / 1e9 / peak_gbs}
    del d_tmp

    # ---- FRI commit chain on a 2^24 domain incl. all Merkle trees (BASELINE metric, second half)
    fn_ = 1 << FRI_LOG_N
    d_vals = dev.to_device(synthetic_elements(fn_, seed=2000 + rank))
    steps_fri = ((fn_ // FRI_L) // FRI_OUT).bit_length() - 1
    leaves = sum(fn_ >> i for i in range(steps_fri + 1))

    def fri_step():
        proto = dev.fri_commit(d_vals, FRI_L, FRI_OUT, FIELD)
        proto.free()

    for _ in range(max(1, args.warmup - 1)):
        fri_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        fri_step()
    torch.cuda.synchronize()
    fri_ms = max_over_ranks((time.perf_counter() - t0) * 1e3) / args.steps
    fprof = profile(fri_step)
    f_total = sum(r["total_ms"] for r in fprof.values())
    leaf_k = fprof.get("merkle_levels_leaf")
    fri = {"metric": "fri_merkle_leaves_per_sec", "workload": f"FRI commit chain on 2^{FRI_LOG_N} values, blowup {FRI_L}, "
           f"{steps_fri} layers, {steps_fri + 1} Merkle trees", "value": world * leaves / (fri_ms * 1e-3), "unit": "leaves/s",
           "ms_per_step": fri_ms, "kernel_ms": f_total, "algorithmic_bytes": 128 * fn_,
           "hbm_frac": 128 * fn_ / (fri_ms * 1e-3) / 1e9 / peak_gbs,
           "kernels": {k: {"launches": r["count"], "total_ms": r["total_ms"], "share": r["total_ms"] / f_total}
                       for k, r in fprof.items()}}
    fri["timer"] = "host perf_counter around the synchronous commit call, max over ranks"
    # the same chain through the host-pointer entry point: pinned LDE values in (H2D inside the timed
    # region), roots / challenges / final coefficients out; trees and layer values stay in HBM behind the
    # prototype handle, as a prover that only opens a few queries needs them
    h_vals = torch.empty((fn_, 4), dtype=torch.int64, pin_memory=True)
    h_vals.numpy().view(np.uint64)[:] = d_vals.cpu().numpy().view(np.uint64).reshape(fn_, 4)
    roots = np.zeros(32 * (steps_fri + 1), np.uint8)
    chals = np.zeros(4 * steps_fri, np.uint64)
    fin = np.zeros(4 * FRI_OUT, np.uint64)

    def fri_e2e_step():
        h = lib.hodor_cuda_fri_commit(C.c_void_p(h_vals.data_ptr()), fn_, FRI_L, FRI_OUT, 0, FIELD)
        if not h:
            raise RuntimeError(_ffi.last_error())
        _ffi.check(lib.hodor_cuda_fri_summary(h, roots.ctypes.data_as(_ffi.u8p), chals.ctypes.data_as(_ffi.u64p),
                                              fin.ctypes.data_as(_ffi.u64p)))
        lib.hodor_cuda_fri_free(h)

    fri_e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        fri_e2e_step()
    torch.cuda.synchronize()
    fe_ms = max_over_ranks((time.perf_counter() - t0) * 1e3) / args.steps
    fri["e2e"] = {"value": world * leaves / (fe_ms * 1e-3), "unit": "leaves/s", "ms_per_step": fe_ms,
                  "h2d_bytes_per_step": 32 * fn_, "d2h_bytes_per_step": int(roots.nbytes + chals.nbytes + fin.nbytes),
                  "api": "hodor_cuda_fri_commit (host LDE values, pinned) + hodor_cuda_fri_summary"}
    del d_vals, h_vals

    # ---- end to end through the host-pointer C ABI, host buffers in, copies inside the timed region.
    # The call a prover makes per register is "lift and commit" (src/prover/mod.rs:73-80: `w.lde(..)` then
    # `I::create(&lde)`): hodor_cuda_lde_commit_batch takes the coefficients from (pinned) host memory, keeps the
    # LDE and its Merkle tree in HBM behind a handle and returns the 32-byte root -- MORE work per step than
    # `value` (the tree over the 2^27 values is built as well), with only the root crossing PCIe on the way
    # back.  The raw-LDE entry points (4 GiB D2H per step) are reported next to it as `raw_lde`.
    all_cpus = os.sched_getaffinity(0)
    near = gpu_local_cpus(local_rank)
    if near:
        os.sched_setaffinity(0, near[0])  # allocate (first-touch) the pinned buffers on the GPU's NUMA node
    h_in = torch.empty((n, 4), dtype=torch.int64, pin_memory=True)
    h_in.numpy().view(np.uint64)[:] = coeffs
    e2e_steps = max(2, args.steps)
    CH = 8  # handles alive at once: 8 x (4 GiB values + 4 GiB nodes)
    want_root = None

    def commit_chunk(src_ptr, count):
        ins = (C.c_void_p * count)(*[src_ptr] * count)
        outs = (C.c_void_p * count)()
        roots = np.zeros((count, 32), np.uint8)
        _ffi.check(lib.hodor_cuda_lde_commit_batch(ins, count, LOG_N, LOG_L, 1, 0, outs, roots.ctypes.data_as(_ffi.u8p), FIELD))
        for i in range(count):
            lib.hodor_cuda_tree_free(outs[i])
        return roots

    def e2e_commit(src_ptr, steps):
        done, r = 0, None
        while done < steps:
            m = min(CH, steps - done)
            r = commit_chunk(src_ptr, m)
            done += m
        return r

    want_root = e2e_commit(h_in.data_ptr(), min(CH, e2e_steps))[0].tobytes()  # warm-up: pool blocks, tables
    barrier()
    l0 = dev.launch_count()
    t0 = time.perf_counter()
    got_roots = e2e_commit(h_in.data_ptr(), e2e_steps)
    torch.cuda.synchronize()
    e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3) / e2e_steps
    e2e_launches = dev.launch_count() - l0
    # device-only time of the same lift-and-commit (coefficients already in HBM)
    d_ins = (C.c_void_p * 1)(d_coeffs.data_ptr())
    d_outs = (C.c_void_p * 1)()

    def commit_dev():
        _ffi.check(lib.hodor_cuda_lde_commit_batch(d_ins, 1, LOG_N, LOG_L, 1, 1, d_outs, None, FIELD))
        lib.hodor_cuda_tree_free(d_outs[0])

    torch.cuda.synchronize()
    commit_dev()
    t0 = time.perf_counter()
    for _ in range(3):
        commit_dev()
    commit_dev_ms = max_over_ranks((time.perf_counter() - t0) * 1e3) / 3
    cprof = profile(commit_dev)
    # the same call from PAGEABLE host memory (what a Rust Vec<F> is): the copy is staged by the driver
    pageable = np.array(coeffs, copy=True)
    commit_chunk(pageable.ctypes.data, 1)
    t0 = time.perf_counter()
    for _ in range(2):
        commit_chunk(pageable.ctypes.data, 2)
    pageable_ms = max_over_ranks((time.perf_counter() - t0) * 1e3) / 4
    # root of the committed LDE == root of a tree built over the device-timed LDE output
    d_nodes = dev.empty_elems(n * L)
    d_root = torch.zeros(32, dtype=torch.uint8, device="cuda")
    lde_step()
    dev.merkle_build(d_out, n * L, d_nodes, FIELD, root=d_root)
    same_root = bytes(d_root.cpu().numpy().tobytes()) == want_root == got_roots[-1].tobytes()
    del d_nodes

    # raw LDE through the host-pointer ABI (the reference's own signature: values back in host memory)
    h_outs = [torch.empty((n * L, 4), dtype=torch.int64, pin_memory=True) for _ in range(2)]
    for h in h_outs:
        h.zero_()
    raw_steps = min(e2e_steps, 4)

    def e2e_single():
        _ffi.check(lib.hodor_cuda_lde(C.cast(h_in.data_ptr(), _ffi.u64p), LOG_N, LOG_L, 1,
                                      C.cast(h_outs[0].data_ptr(), _ffi.u64p), FIELD))

    ins_arr = (_ffi.u64p * raw_steps)(*[C.cast(h_in.data_ptr(), _ffi.u64p)] * raw_steps)
    outs_arr = (_ffi.u64p * raw_steps)(*[C.cast(h_outs[i % 2].data_ptr(), _ffi.u64p) for i in range(raw_steps)])

    def e2e_batch():
        _ffi.check(lib.hodor_cuda_lde_batch(ins_arr, outs_arr, raw_steps, LOG_N, LOG_L, 1, FIELD))

    e2e_single()
    barrier()
    t0 = time.perf_counter()
    e2e_single()
    torch.cuda.synchronize()
    single_ms = max_over_ranks((time.perf_counter() - t0) * 1e3)
    e2e_batch()  # warm-up: staging buffers come from the pool afterwards
    barrier()
    t0 = time.perf_counter()
    e2e_batch()
    torch.cuda.synchronize()
    raw_ms = max_over_ranks((time.perf_counter() - t0) * 1e3) / raw_steps
    same = all(bool(np.array_equal(h.numpy()[: 1 << 16], d_out[: 1 << 16].cpu().numpy())) and
               bool(np.array_equal(h.numpy()[-(1 << 16):], d_out[-(1 << 16):].cpu().numpy())) for h in h_outs)
    del h_outs
    pg_out = np.zeros((n * L, 4), np.uint64)
    t0 = time.perf_counter()
    _ffi.check(lib.hodor_cuda_lde(_p(pageable), LOG_N, LOG_L, 1, _p(pg_out), FIELD))
    raw_pageable_ms = max_over_ranks((time.perf_counter() - t0) * 1e3)
    del pg_out, pageable
    e2e = {"value": world * n * L / (e2e_ms * 1e-3), "unit": "field-elems/s", "h2d_bytes_per_step": 32 * n,
           "d2h_bytes_per_step": 32, "ms_per_step": e2e_ms, "steps": e2e_steps,
           "api": "hodor_cuda_lde_commit_batch (host coefficients, pinned; per polynomial: H2D, coset LDE 2^24 -> 2^27, "
                  "Blake2s Merkle tree over the 2^27 values, root D2H; LDE and tree stay in HBM behind a handle; "
                  f"{CH} polynomials per call, H2D of the next one overlaps the kernels) + hodor_cuda_tree_free",
           "work_per_step": "the LDE of `value` PLUS the commitment to it (2^28 - 1 Blake2s compressions)",
           "device_only_ms_per_step": commit_dev_ms,
           "kernels": {k: {"launches": r["count"], "total_ms": r["total_ms"]} for k, r in cprof.items()},
           "gpu_launches": e2e_launches,
           "pageable_host_memory": {"ms_per_step": pageable_ms, "value": world * n * L / (pageable_ms * 1e-3),
                                    "note": "same call from an unpinned numpy buffer (what a Rust Vec<F> is)"},
           "timer": "host perf_counter around the synchronous C-ABI calls, max over ranks",
           "host_buffers": (f"pinned, allocated on NUMA node {near[1]} next to the GPU ({len(near[0])} CPUs)"
                            if near and near[1] not in ("-1", "") else "pinned; the box exposes no NUMA topology"),
           "root_matches_device_result": bool(same_root),
           "raw_lde": {"api": "hodor_cuda_lde_batch (values copied back: 4 GiB D2H per polynomial)", "ms_per_step": raw_ms,
                       "value": world * n * L / (raw_ms * 1e-3), "h2d_bytes_per_step": 32 * n, "d2h_bytes_per_step": 32 * n * L,
                       "single_call_ms": single_ms, "single_call_pageable_ms": raw_pageable_ms,
                       "matches_device_result": same}}
    del h_in
    os.sched_setaffinity(0, all_cpus)
    _ffi.check(lib.hodor_cuda_trim())  # the committed-oracle blocks of this leg go back to the driver

    # ---- the paths with a real exchange step, through the C ABI (hodor_cuda_ntt_sharded / _lde_fri_sharded:
    # NCCL send/recv issued by the library), at EVERY N including 1, each checked against the single-GPU result
    from hodor_b200 import multigpu as mg
    mg.comm_init()

    def device_elements(count, seed, start=0, step=1):
        """Deterministic canonical elements made on the device: element j depends on its global index only, so
        every rank can build its own slice (start + step * t) of one shared vector without host traffic."""
        idx = start + step * torch.arange(count, dtype=torch.int64, device="cuda")
        out = torch.empty((count, 4), dtype=torch.int64, device="cuda")
        for limb, mul in enumerate((0x2545F4914F6CDD1D, 0x5851F42D4C957F2D, 0x14057B7EF767814F, 0x27BB2EE687B0B0FD)):
            off = ((seed + 1) * 0x632BE59BD9B4E019) & (2**63 - 1)  # python int -> an int64 scalar
            x = (idx + off) * mul
            x ^= (x >> 29)
            out[:, limb] = x * 0x369DEA0F31A53F85
        out[:, 3] &= 0x3FFFFFFFFFFFFFFF  # top limb < 2^62 < the modulus' top limb: canonical
        return out

    def timed_max(fn, reps):
        fn()
        best = None
        for _ in range(reps):
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            ms = max_over_ranks(e0.elapsed_time(e1))
            best = ms if best is None else min(best, ms)
        return best

    def timed_local(fn, reps):
        """This rank only: no collective inside (rank 0 times its single-GPU reference alone)."""
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    log_g = world.bit_length() - 1
    sharded = []
    sizes = list(range(18, 29, 2)) if not args.sweep else list(range(18, 29))  # configs[4]: 2^18 .. 2^28
    for ln in sizes:
        nn = 1 << ln
        m = nn >> log_g
        omega = H.Domain.new_for_size(FIELD, nn).generator
        local = device_elements(m, ln, start=rank, step=world)  # cyclic slice a[j * G + rank]
        out = dev.empty_elems(m)
        ms = timed_max(lambda: mg.ntt_sharded(local, ln, omega, FIELD, out), 3)
        row = {"log_n": ln, "ms": ms, "value": nn / (ms * 1e-3), "unit": "field-elems/s",
               "hbm_frac_aggregate": 64 * nn / (ms * 1e-3) / 1e9 / (peak_gbs * world)}
        # parity + the single-GPU time of the same transform, on rank 0 (its own share of the output: the
        # rank-th n/G^2 chunk of every n/G block; every rank's share at sizes <= 2^24 via gather)
        if world > 1:
            mg.ntt_sharded(local, ln, omega, FIELD, out)
            ok = torch.ones(1, dtype=torch.int64, device="cuda")
            if rank == 0:
                full = device_elements(nn, ln)
                ref = dev.empty_elems(nn)
                t1 = timed_local(lambda: dev.fft(full, ref, ln, False, FIELD), 3)
                del full
                chunk = m >> log_g
                mine = ref.view(world, world, chunk, 4)[:, 0].reshape(m, 4)  # A[k2 * m + 0 * chunk + k]
                good = bool(torch.equal(mine, out))
                row["single_gpu_ms"] = t1
                row["strong_efficiency"] = t1 / (world * ms)
            if ln <= 24:
                parts = [torch.empty_like(out) for _ in range(world)] if rank == 0 else None
                dist.gather(out, parts, dst=0)
                if rank == 0:
                    for h in range(world):
                        good = good and bool(torch.equal(ref.view(world, world, m >> log_g, 4)[:, h].reshape(m, 4), parts[h]))
                    row["checked"] = "every rank's output == single-GPU NTT (gathered)"
                    del parts
            elif rank == 0:
                row["checked"] = "rank 0's output share == single-GPU NTT"
            if rank == 0:
                row["matches_single_gpu"] = good
                del ref
        else:
            row["matches_single_gpu"] = True  # world 1: the same kernels, nothing exchanged
            row["strong_efficiency"] = 1.0
        sharded.append(row)
        del local, out
        torch.cuda.empty_cache()

    exchange_kind = ("single GPU: nothing exchanged" if world == 1 else
                     (f"peer stores over NVLink ({mg.bytes_peer_stored()} B stored into peers' buffers by rank 0's kernels, "
                      f"{mg.bytes_sent()} B through NCCL)" if mg.bytes_peer_stored() > 0 else
                      f"NCCL send/recv all-to-all ({mg.bytes_sent()} B sent by rank 0)"))

    # ---- the north-star target: ONE 2^24 -> 2^28 coset LDE + full FRI commit chain over all ranks
    s_log_f = 4
    d_shared = device_elements(n, 4000)  # replicated coefficient vector

    def lde_fri():
        return mg.lde_fri_sharded(d_shared, LOG_N, s_log_f, True, 1, FIELD)

    res = lde_fri()
    barrier()
    reps = 3
    t0 = time.perf_counter()
    for _ in range(reps):
        res = lde_fri()
    torch.cuda.synchronize()
    s_ms = max_over_ranks((time.perf_counter() - t0) * 1e3) / reps
    sharded_fri = {"workload": f"coset LDE 2^{LOG_N} -> 2^{LOG_N + s_log_f} + FRI commit chain ({len(res[1])} layers, "
                               f"{len(res[1]) + 1} Merkle trees) sharded over {world} GPU(s)", "ms": s_ms,
                   "lde_elems_per_s": (n << s_log_f) / (s_ms * 1e-3), "scaling": "strong",
                   "api": "hodor_cuda_lde_fri_sharded (C ABI; rank r owns L/G adjacent cosets, folds and the bottom tree levels are local; per "
                          "committed layer one NCCL send/recv all-to-all of DIGESTS + a 32-byte all-gather of sub-roots)",
                   "nccl_payload_bytes_per_rank": None,
                   "timer": "host perf_counter around the synchronous C-ABI call, barrier + synchronize both sides, max over ranks"}
    b0 = mg.bytes_sent()
    lde_fri()
    sharded_fri["nccl_payload_bytes_per_rank"] = mg.bytes_sent() - b0
    if rank == 0:
        # the same polynomial through the single-GPU entry points
        full = dev.empty_elems(n << s_log_f)
        dev.lde(d_shared, LOG_N, s_log_f, True, full, FIELD)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        dev.lde(d_shared, LOG_N, s_log_f, True, full, FIELD)
        ref = dev.fri_commit(full, 1 << s_log_f, 1, FIELD)
        torch.cuda.synchronize()
        one_ms = (time.perf_counter() - t0) * 1e3
        sharded_fri["matches_single_gpu"] = bool(res[0] == ref.get_roots() and np.array_equal(res[1], ref.challenges)
                                                 and np.array_equal(res[2], ref.final_coefficients))
        sharded_fri["single_gpu_ms"] = one_ms
        sharded_fri["strong_efficiency"] = one_ms / (world * s_ms)
        ref.free()
        del full
    del d_shared
    torch.cuda.empty_cache()

    # ---- configs[4], single-GPU column: forward NTT 2^18 .. 2^28 (the sharded columns are `sharded_ntt`)
    sweep = None
    if world == 1 and not args.no_sweep:
        sweep = []
        for ln in range(18, 29):
            a = dev.to_device(synthetic_elements(1 << ln, seed=ln))
            b = dev.empty_elems(1 << ln)
            reps = 5 if ln <= 26 else 3
            ms = timed(lambda: dev.fft(a, b, ln, False, FIELD), reps, 3) / reps
            sweep.append({"log_n": ln, "ms": ms, "value": (1 << ln) / (ms * 1e-3),
                          "hbm_frac": 64 * (1 << ln) / (ms * 1e-3) / 1e9 / peak_gbs})
            del a, b
        if rank == 0 and not args.no_cpu:
            # the reference's best_fft (parallel_fft over all host cores, src/fft/fft.rs:68-125) per size:
            # measured up to 2^24, extrapolated with n log n beyond (labelled)
            from oracle import oracle as O
            O.build()
            cpu_ms = {}
            for ln in (18, 20, 22, 24):
                x = O.random_elements(FIELD, 1 << ln, seed=ln)
                t0 = time.perf_counter()
                O.best_fft(FIELD, x, O.domain_generator(FIELD, ln), ln)
                cpu_ms[ln] = (time.perf_counter() - t0) * 1e3
            for row in sweep:
                ln = row["log_n"]
                if ln in cpu_ms:
                    row["cpu_ms"], row["cpu_kind"] = cpu_ms[ln], "measured"
                else:
                    base = max(k for k in cpu_ms if k <= ln) if ln > 18 else 18
                    row["cpu_ms"] = cpu_ms[base] * ((1 << ln) * ln) / ((1 << base) * base)
                    row["cpu_kind"] = f"extrapolated from 2^{base} by n log n"
                row["speedup_vs_cpu"] = row["cpu_ms"] / row["ms"]
                row["cpu_cores"] = O.default_cpus()

    # ---- configs[3]: Fibonacci AIR end-to-end prove, trace 2^20, blowup 16 (hot-path call sequence of
    # src/prover/mod.rs:66-174 replayed on device-resident polynomials; hodor_b200/fib_replay.py)
    fib = None
    if world == 1 and not args.no_fib:
        fib = run_fib(args, dev, lib, _ffi, profile, cpu=not args.no_cpu)

    # ---- the headline rows repeated for the other two fields the library carries (true BN254 Fr, the
    # Stark-252 prime of src/experiments): same kernels, other template constants
    other_fields = None
    if world == 1 and not args.no_fields:
        other_fields = {}
        tops = {1: 0x30644E72E131A029, 2: 0x0800000000000011}  # top limb of p: keeps the synthetic elements canonical
        for fid, name in ((1, "bn254_fr"), (2, "stark252")):
            rng = np.random.default_rng(3000 + fid)
            a = rng.integers(0, 2**64, size=(n, 4), dtype=np.uint64)
            a[:, 3] = rng.integers(0, tops[fid], size=n, dtype=np.uint64)
            d_a = dev.to_device(a)
            d_o = dev.empty_elems(n * L)
            d_t = dev.empty_elems(n)
            lde_ms = timed(lambda: dev.lde(d_a, LOG_N, LOG_L, True, d_o, fid), 3, 2) / 3
            ntt_ms_f = timed(lambda: dev.fft(d_a, d_t, LOG_N, False, fid), 3, 2) / 3

            def chain():
                dev.fri_commit(d_a, FRI_L, FRI_OUT, fid).free()

            chain()
            barrier()
            t0 = time.perf_counter()
            for _ in range(3):
                chain()
            torch.cuda.synchronize()
            fri_ms_f = (time.perf_counter() - t0) * 1e3 / 3
            other_fields[name] = {"lde_2p24_x8_ms": lde_ms, "lde_elems_per_s": n * L / (lde_ms * 1e-3), "ntt_2p24_ms": ntt_ms_f,
                                  "ntt_elems_per_s": n / (ntt_ms_f * 1e-3), "fri_chain_2p24_ms": fri_ms_f,
                                  "fri_leaves_per_s": leaves / (fri_ms_f * 1e-3)}
            del d_a, d_o, d_t
            torch.cuda.empty_cache()

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu:
        v, cores, dt = cpu_lde_sample(20, 2)
        fv, _, fdt = cpu_fri_sample(20)
        cpu_baseline = {"value": v, "unit": "field-elems/s", "cores": cores, "kind": "port",
                        "sample": f"coset LDE 2^20 x{L} (1/16 of the workload), best of 2, {dt:.2f} s; C restatement of the "
                                  "reference's multi-coset crossbeam path (Rust toolchain unavailable)",
                        "fri_leaves_per_sec": fv, "fri_sample": f"FRI chain on 2^20 values, {fdt:.2f} s"}

    if rank == 0:
        line = {
            "metric": "ntt_field_elems_per_sec", "value": value, "unit": "field-elems/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u256 (8 x u32 Montgomery limbs, INT32 pipe)",
            "data": "synthetic",
            "config": bench_config(world),
            "clocks": clocks.summary(), "e2e": e2e, "gpu_launches": launches_timed, "roofline": roofline,
            "cpu_baseline": cpu_baseline, "ntt": ntt, "fri": fri,
        }
        line["sharded_ntt"] = {"api": "hodor_cuda_ntt_sharded (C ABI; four-step; the exchange is fused into the last pass of the local "
                                      "transform as NVLink peer stores, NCCL send/recv where peer mapping is unavailable)",
                               "exchange": exchange_kind, "scaling": "strong", "sizes": sharded}
        line["sharded_lde_fri"] = sharded_fri
        if sweep is not None:
            line["ntt_sweep"] = sweep
        if fib is not None:
            line["fib_prove"] = fib
        if other_fields is not None:
            line["other_fields"] = other_fields
        emit(line)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--sweep", action="store_true", help="sharded legs: every size 2^18..2^28 instead of the even ones")
    ap.add_argument("--no-sweep", action="store_true", help="skip the single-GPU NTT size sweep 2^18..2^28 (configs[4])")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-fields", action="store_true", help="skip the BN254 / Stark-252 repeats of the headline rows")
    ap.add_argument("--no-fib", action="store_true", help="skip the Fibonacci prove leg (configs[3])")
    ap.add_argument("--config", choices=["default", "fib"], default="default", help="fib: only the Fibonacci prove leg")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    # The contract is ONE JSON line on stdout.  Libraries write banners to fd 1 behind Python's back
    # (NCCL prints its version there), so everything but the final line goes to stderr.
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)

    if args.impl == "reference":
        run_reference(args, rank)
        return

    if args.config == "fib":
        if rank == 0:
            import torch
            import hodor_b200 as H
            from hodor_b200 import _ffi
            from hodor_b200 import device as dev
            torch.cuda.set_device(local_rank)
            H.init(local_rank)

            def profile(fn):
                _ffi.check(_ffi.lib.hodor_cuda_profile_begin())
                fn()
                buf = C.create_string_buffer(1 << 16)
                _ffi.check(_ffi.lib.hodor_cuda_profile_end(buf, len(buf)))
                return {r["name"]: r for r in json.loads(buf.value.decode())}

            emit({"metric": "fib_prove_ms", "config": {"workload": "configs[3]"}, **run_fib(args, dev, _ffi.lib, _ffi, profile, not args.no_cpu)})
        return

    if world > 1:
        import torch
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        run_ours(args, rank, local_rank, world)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
