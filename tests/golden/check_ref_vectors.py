#!/usr/bin/env python
"""Compares a vectors file produced by the REAL reference (rust/src/gen_vectors.rs run inside
matter-labs/hodor with cargo) with the committed tests/golden/vectors.json, case by case.

    python tests/golden/check_ref_vectors.py vectors_ref.json

Exit code 0 and "PINNED" means every known answer this repository tests against (C oracle on the CPU,
CUDA path on the GPU: tests/test_golden.py) is an output of the reference itself.  Nothing in this
repository can run that generator (no Rust toolchain in the image); until somebody does, parity
stays "unpinned" and every header says so."""
import json
import os
import sys


def key(c):
    return (c["kind"], c["field"], c["log_n"], c["seed"])


def main():
    if len(sys.argv) != 2:
        sys.exit(__doc__)
    here = os.path.dirname(os.path.abspath(__file__))
    ours = {key(c): c for c in json.load(open(os.path.join(here, "vectors.json")))["cases"]}
    ref = {key(c): c for c in json.load(open(sys.argv[1]))["cases"]}
    bad = 0
    for k, c in sorted(ours.items()):
        r = ref.get(k)
        if r is None:
            print("MISSING in the reference file:", k)
            bad += 1
            continue
        for field, want in c.items():
            if r.get(field) != want:
                print(f"MISMATCH {k} field {field!r}:\n   reference {r.get(field)!r}\n   ours      {want!r}")
                bad += 1
    extra = sorted(set(ref) - set(ours))
    if extra:
        print("cases only in the reference file (ignored):", extra)
    print(f"{len(ours)} cases, {bad} problems ->", "PINNED" if bad == 0 else "NOT PINNED")
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
