"""Generates tests/golden/vectors.json from the independent big-integer model (oracle/pymodel.py,
Python ints + hashlib.blake2s).  The Rust reference cannot be executed in this environment, so these
are known answers of the *published algorithms* on seeded inputs, not outputs of the reference
binary ("parity unpinned", see oracle/hodor_oracle.c).  Both the C oracle (CPU tests) and the CUDA
path (GPU tests) are compared with this file.

    python tests/golden/make_golden.py
"""
import hashlib
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import pymodel as M  # noqa: E402

FIELDS = [(0, M.BLS12_381_FR), (1, M.BN254_FR), (2, M.STARK252)]


def mont_bytes(F, plain_values):
    return b"".join(F.to_mont(v).to_bytes(32, "little") for v in plain_values)


def digest(b: bytes) -> str:
    return hashlib.sha256(b).hexdigest()


def main():
    out = {"generator": "splitmix64 seed 0x3DBE62598D313D76 (+case index), limbs used directly as Montgomery form",
           "cases": []}
    for fid, F in FIELDS:
        for ci, ln in enumerate([0, 1, 2, 5, 8, 11, 12, 13]):
            seed = 0x3DBE62598D313D76 + ci
            a_m = M.random_mont_elements(F, 1 << ln, seed)
            a = [F.from_mont(x) for x in a_m]
            res = M.serial_fft(F, a, F.domain_generator(ln), ln)
            out["cases"].append({"kind": "ntt", "field": fid, "log_n": ln, "seed": seed,
                                 "sha256": digest(mont_bytes(F, res)), "first": hex(F.to_mont(res[0])),
                                 "last": hex(F.to_mont(res[-1]))})
        for ci, (ln, L, coset) in enumerate([(3, 2, False), (4, 8, True), (6, 8, True), (9, 8, True), (5, 16, False)]):
            seed = 0x1000 + ci
            a = [F.from_mont(x) for x in M.random_mont_elements(F, 1 << ln, seed)]
            res = M.lde(F, a, ln, L, coset)
            out["cases"].append({"kind": "lde", "field": fid, "log_n": ln, "factor": L, "coset": coset, "seed": seed,
                                 "sha256": digest(mont_bytes(F, res))})
        for ci, ln in enumerate([1, 2, 5, 9, 13]):
            seed = 0x2000 + ci
            a = [F.from_mont(x) for x in M.random_mont_elements(F, 1 << ln, seed)]
            nodes = M.merkle_create(F, a)
            out["cases"].append({"kind": "merkle", "field": fid, "log_n": ln, "seed": seed, "root": nodes[1].hex(),
                                 "nodes_sha256": digest(b"".join(nodes)),
                                 "challenge": hex(F.to_mont(M.interpret_hash(F, nodes[1])))})
        for ci, (ln, L, oc) in enumerate([(4, 4, 2), (8, 8, 1), (10, 16, 4)]):
            seed = 0x3000 + ci
            a = [F.from_mont(x) for x in M.random_mont_elements(F, 1 << ln, seed)]
            pr = M.fri_commit(F, a, L, oc)
            out["cases"].append({"kind": "fri", "field": fid, "log_n": ln, "lde_factor": L, "out_coeffs": oc, "seed": seed,
                                 "roots": [pr.l0_nodes[1].hex()] + [n[1].hex() for n in pr.layer_nodes],
                                 "challenges": [hex(F.to_mont(c)) for c in pr.challenges],
                                 "final_root": pr.final_root.hex(),
                                 "final_coefficients": [hex(F.to_mont(c)) for c in pr.final_coefficients],
                                 "values_sha256": [digest(mont_bytes(F, v)) for v in pr.layer_values]})
        # Polynomial::batch_inversion / evaluate_at (src/polynomials/mod.rs:889-954, 685-711)
        for ci, ln in enumerate([0, 3, 8, 12]):
            seed = 0x4000 + ci
            a = [F.from_mont(x) for x in M.random_mont_elements(F, 1 << ln, seed)]
            inv = [pow(v, -1, F.p) for v in a]
            out["cases"].append({"kind": "batch_inversion", "field": fid, "log_n": ln, "seed": seed,
                                 "sha256": digest(mont_bytes(F, inv)), "first": hex(F.to_mont(inv[0]))})
            z = F.from_mont(M.random_mont_elements(F, 1, seed + 0x100)[0])
            out["cases"].append({"kind": "evaluate_at", "field": fid, "log_n": ln, "seed": seed, "point_seed": seed + 0x100,
                                 "value": hex(F.to_mont(M.evaluate(F, a, z)))})
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "vectors.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", path, len(out["cases"]), "cases")


if __name__ == "__main__":
    main()
