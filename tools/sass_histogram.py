#!/usr/bin/env python
"""SASS opcode histogram of the hot kernels in hodor_b200/libhodor_b200.so (cuobjdump -sass; no GPU needed).
    python tools/sass_histogram.py > profiles/r02_sass_histogram.md
Shows, per kernel: instruction count, code size, and the mnemonics that decide which pipe bounds it
(IMAD.WIDE / IMAD / IMAD.HI on the multiplier pipe; IADD3 / LOP3 / SHF / SEL on the ALU pipe), plus the
Blackwell-specific mnemonics the round-1 verdict asked about (UTMALDG / UBLKCP / SHFL / LDGSTS / CCTL)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "hodor_b200", "libhodor_b200.so")
KERNELS = [
    ("ntt_pass_kernel<BlsFr, 8, SCALE_IN, !LAST>  (pass 1 of the coset LDE: the dominant kernel)", "ntt_pass_kernelINS_5BlsFrELi8ELb1ELb0"),
    ("ntt_pass_kernel<BlsFr, 8, !SCALE_IN, !LAST> (middle pass)", "ntt_pass_kernelINS_5BlsFrELi8ELb0ELb0"),
    ("ntt_pass_kernel<BlsFr, 8, !SCALE_IN, LAST>  (last pass: natural-order / interleaved / peer stores)", "ntt_pass_kernelINS_5BlsFrELi8ELb0ELb1"),
    ("merkle_levels_kernel<3, LEAF>", "merkle_levels_kernelILi3ELb1"),
    ("fri_fold_kernel<BlsFr, FLAT>", "fri_fold_kernelINS_5BlsFrELb1"),
    ("fri_fold_commit_kernel<BlsFr, FLAT> (fold into a shared tile + leaf subtrees from it; off by default)", "fri_fold_commit_kernelINS_5BlsFrELb1"),
    ("ntt_last_commit_kernel<BlsFr, 8> (last pass + bottom three tree levels: what lift-and-commit runs at 2^24)", "ntt_last_commit_kernelINS_5BlsFrELi8"),
]
PIPE = {"fma (multiplier) pipe": ("IMAD.WIDE", "IMAD.HI", "IMAD.X", "IMAD.IADD", "IMAD.MOV", "IMAD.SHL", "IMAD", "HFMA2", "FFMA"),
        "alu pipe": ("IADD3", "LOP3", "SHF", "SEL", "PRMT", "MOV", "ISETP", "LEA", "VIADD", "IABS", "FMNMX"),
        "memory": ("LDG", "STG", "LDS", "STS", "LDC", "LDCU", "LDGSTS", "UBLKCP", "UTMALDG", "UTMASTG", "LDL", "STL"),
        "control": ("BRA", "BAR", "CALL", "RET", "EXIT", "BSSY", "BSYNC", "NOP", "MEMBAR", "CCTL", "SHFL", "WARPSYNC")}


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    funcs, cur = {}, None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            funcs[cur] = []
        elif cur is not None:
            m = re.match(r"\s+/\*([0-9a-f]+)\*/\s+(?:@!?U?P\w+\s+)?([A-Z][A-Z0-9_.]*)", line)
            if m:
                funcs[cur].append(m.group(2))
    print(f"# SASS opcode histograms (`cuobjdump -sass {os.path.relpath(LIB, ROOT)}`, sm_100a)\n")
    print("Made by `tools/sass_histogram.py`.  16 bytes per instruction.  No `UTMALDG` / `UBLKCP` (TMA) and no `SHFL` appear: the tiles are "
          "staged by per-thread 256-bit global accesses (`LDG.E.256` / `STG.E.256`, one 32-byte element = one sector per instruction, sm_100 only) and `STS.128` / `LDS.128` because each thread multiplies its own elements by its own streamed table "
          "entries (nothing to broadcast or bulk-copy into a shared tile that fits: a 2^8 x 8 tile's table entries are 128 KiB), and the "
          "exchange between butterfly groups goes through shared memory, not shuffles (a warp-autonomous shuffle variant was measured "
          "in round 1: -13 %).  See profiles/r02_experiments.md.\n")
    for title, pat in KERNELS:
        names = [f for f in funcs if pat in f]
        if not names:
            print(f"## {title}\n\n(not in this build)\n")
            continue
        ops = funcs[names[0]]
        hist = collections.Counter(ops)
        print(f"## {title}\n\n`{names[0]}`: {len(ops)} instructions, {16 * len(ops) / 1024:.1f} KiB\n")
        print("| pipe | instructions | top mnemonics |\n|---|---|---|")
        seen = set()
        for pipe, prefixes in PIPE.items():
            sel = {k: v for k, v in hist.items() if any(k == p or k.startswith(p + ".") for p in prefixes) and k not in seen}
            seen |= set(sel)
            top = ", ".join(f"{k} {v}" for k, v in sorted(sel.items(), key=lambda kv: -kv[1])[:8])
            print(f"| {pipe} | {sum(sel.values())} | {top} |")
        rest = {k: v for k, v in hist.items() if k not in seen}
        print(f"| other | {sum(rest.values())} | " + ", ".join(f"{k} {v}" for k, v in sorted(rest.items(), key=lambda kv: -kv[1])[:8]) + " |")
        wide = sum(v for k, v in hist.items() if k.startswith("IMAD.WIDE"))
        single = sum(v for k, v in hist.items() if k.startswith("IMAD") and not k.startswith("IMAD.WIDE"))
        print(f"\nmultiplier-pipe issue slots (IMAD.WIDE = 2.25 slots measured by tools/mixbench.cu, other IMAD = 1): "
              f"{wide} x 2.25 + {single} = {wide * 2.25 + single:.0f} (static count over the whole kernel body)\n")
        tma = [k for k in hist if k.startswith(("UTMA", "UBLKCP", "SHFL", "LDGSTS"))]
        print(f"TMA / shuffle / cp.async mnemonics present: {', '.join(tma) if tma else 'none'}\n")


if __name__ == "__main__":
    main()
