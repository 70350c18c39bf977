#!/usr/bin/env python3
"""Generates hodor_b200/csrc/shoup_rows.cuh: the carry chains of the fixed-operand ("Shoup")
multiplier as straight-line PTX, one asm statement per chain.

    python tools/gen_shoup.py > hodor_b200/csrc/shoup_rows.cuh

Two products of two 8 x u32 numbers are emitted, both by rows with two carry-save accumulators
(products whose low word lands on an even / odd word index), like the Montgomery multiplier in
field.cuh, so that every 32x32->64 product is one mad.lo.cc / madc.hi.cc pair on an aligned pair:

  hi_trunc(a, b)   words 7..15 of  sum_{i+j >= 7} a_i b_j 2^(32(i+j)) + sum_{i+j = 6} hi32(a_i b_j) 2^(32*7)
                   (the exact product minus a dropped part D < 14 * 2^224; word 7 is the guard word)
  lo_acc(x, y)     words 0..7 of x*y, accumulated (mod 2^256) into running accumulators; `y` either a
                   register vector or compile-time words (template constants of the field)
"""
import os
import sys

IMM_FIRST = os.environ.get('IMM_FIRST', '1') == '1'
IMM_ROLE = os.environ.get('IMM_ROLE', 'y')
SWAP_HI = os.environ.get('SWAP_HI', '0') == '1'


class Acc:
    """Two absolute-word-indexed accumulators e[w], o[w]; tracks which words are initialised."""

    def __init__(self, lo, hi):
        self.init = {"e": set(), "o": set()}
        self.lo, self.hi = lo, hi


def chain(acc, par, items, yname, xname, imm=False):
    """items: list of (kind, i, word) in ascending word order; kind in lo/hi.  Returns one asm
    statement that adds x_i * y (low or high half) into acc[par][word] with a carry chain."""
    lines, outs, ins = [], [], []
    out_idx, in_idx = {}, {}

    def out_op(w, fresh):
        key = (par, w)
        if key not in out_idx:
            out_idx[key] = len(outs)
            outs.append(('"=r"' if fresh else '"+r"') + f"({par}[{w}])")
        return f"%{out_idx[key]}"

    # operand numbering: outputs first, then inputs -- resolved in a second pass
    plan = []
    first = True
    for n, (kind, i, w) in enumerate(items):
        fresh = w not in acc.init[par]
        last = n == len(items) - 1
        plan.append((kind, i, w, fresh, first, last))
        first = False
    last_w = items[-1][2]
    last_fresh = last_w not in acc.init[par]
    carry_word = None
    if not last_fresh and last_w + 1 <= acc.hi:
        carry_word = last_w + 1
        assert carry_word not in acc.init[par], (par, carry_word)
    for kind, i, w, fresh, _, _ in plan:
        out_op(w, fresh)
    if carry_word is not None:
        out_op(carry_word, True)
    nout = len(outs)

    def in_op(name):
        if name not in in_idx:
            in_idx[name] = nout + len(ins)
            ins.append(name)
        return f"%{in_idx[name]}"

    for kind, i, w, fresh, is_first, is_last in plan:
        d = f"%{out_idx[(par, w)]}"
        addend = "0" if fresh else d
        if imm and IMM_ROLE == 'x':   # immediates vary along the chain, the common factor is a register
            x = in_op(f'"r"({xname.format(i=i)})')
            y = in_op(f'"r"({yname})')
        else:
            x = in_op(f'"r"({xname}[{i}])')
            y = in_op(f'"n"({yname})' if imm else f'"r"({yname})')
        cc_out = ".cc" if (not is_last or carry_word is not None) else ""
        op = ("mad" if is_first else "madc") + f".{kind}{cc_out}.u32"
        if imm and IMM_FIRST:
            x, y = y, x
        elif not imm and SWAP_HI and kind == "hi":
            x, y = y, x
        lines.append(f"{op} {d}, {x}, {y}, {addend};")
        acc.init[par].add(w)
    if carry_word is not None:
        lines.append(f"addc.u32 %{out_idx[(par, carry_word)]}, 0, 0;")
        acc.init[par].add(carry_word)
    body = '\\n\\t"\n        "'.join(lines)
    return f'    asm("{body}"\n        : {", ".join(outs)}\n        : {", ".join(ins)});\n'


def gen_hi_trunc():
    acc = Acc(7, 15)
    out = []
    for j in range(8):
        per = {"e": [], "o": []}
        for i in range(8):
            c = i + j
            par = "e" if c % 2 == 0 else "o"
            if c == 6:
                per[par].append(("hi", i, 7))
            elif c >= 7:
                per[par].append(("lo", i, c))
                per[par].append(("hi", i, c + 1))
        for par in ("e", "o"):
            items = sorted(per[par], key=lambda t: t[2])
            if items:
                out.append(chain(acc, par, items, f"b[{j}]", "a"))
    return "".join(out), acc


def gen_lo(xname, yfmt, imm, acc):
    out = []
    for j in range(8):
        per = {"e": [], "o": []}
        for i in range(8 - j):
            c = i + j
            par = "e" if c % 2 == 0 else "o"
            per[par].append(("lo", i, c))
            if c + 1 <= 7:
                per[par].append(("hi", i, c + 1))
        body = ""
        for par in ("e", "o"):
            items = sorted(per[par], key=lambda t: t[2])
            if items:
                body += chain(acc, par, items, yfmt.format(j=j), xname, imm)
        if imm and IMM_ROLE != 'x':
            # a word of NP equal to 2^32 - 1 is handled after the combine with two add/sub chains
            # (ALU pipe) instead of 8 - j products; a zero word contributes nothing
            y = yfmt.format(j=j)
            out.append(f"    if constexpr ({y} != 0u && {y} != 0xffffffffu) {{\n{body}    }}\n")
        else:
            out.append(body)
    return "".join(out)


def main():
    w = sys.stdout.write
    w("// GENERATED by tools/gen_shoup.py -- do not edit.  Carry chains of the fixed-operand multiplier\n")
    w("// (Field<F>::mul_pre in field.cuh); device only.\n#pragma once\n#include <stdint.h>\n\n")
    w("namespace hodor {\n\n")
    w("#ifdef __CUDA_ARCH__\n")
    w("// q[0..7] = words 8..15, guard = word 7 of the truncated product a*b (see tools/gen_shoup.py)\n")
    w("__device__ __forceinline__ void shoup_hi_trunc(uint32_t (&q)[8], uint32_t& guard, const uint32_t (&a)[8],\n"
      "                                               const uint32_t (&b)[8]) {\n")
    w("    uint32_t e[16], o[16];\n")
    body, acc = gen_hi_trunc()
    w(body)
    for par in ("e", "o"):
        for k in range(7, 16):
            if k not in acc.init[par]:
                w(f"    {par}[{k}] = 0;\n")
    w('    asm("add.cc.u32 %0, %9, %18;\\n\\t"\n')
    for k in range(1, 8):
        w(f'        "addc.cc.u32 %{k}, %{9 + k}, %{18 + k};\\n\\t"\n')
    w('        "addc.u32 %8, %17, %26;"\n')
    w("        : \"=r\"(guard), " + ", ".join(f'"=r"(q[{k}])' for k in range(8)) + "\n")
    w("        : " + ", ".join(f'"r"(e[{k}])' for k in range(7, 16)) + ", " + ", ".join(f'"r"(o[{k}])' for k in range(7, 16)) + ");\n")
    w("}\n\n")
    w("// r = (a*w + q*NP) mod 2^256, NP = 2^256 - p as compile-time words F::NP(j)\n")
    w("template <class F>\n")
    w("__device__ __forceinline__ void shoup_lo2(uint32_t (&r)[8], const uint32_t (&a)[8], const uint32_t (&w)[8],\n"
      "                                          const uint32_t (&q)[8]) {\n")
    w("    uint32_t e[8], o[8];\n")
    acc = Acc(0, 7)
    w(gen_lo("a", "w[{j}]", False, acc))
    if IMM_ROLE == 'x':
        w(gen_lo("F::NP({i})", "q[{j}]", True, acc))
    else:
        w(gen_lo("q", "F::NP({j})", True, acc))
    assert acc.init["e"] == set(range(8)) and acc.init["o"] == set(range(1, 8)), acc.init
    w('    asm("add.cc.u32 %0, %8, 0;\\n\\t"\n')
    for k in range(1, 7):
        w(f'        "addc.cc.u32 %{k}, %{8 + k}, %{15 + k};\\n\\t"\n')
    w('        "addc.u32 %7, %15, %22;"\n')
    w("        : " + ", ".join(f'"=r"(r[{k}])' for k in range(8)) + "\n")
    w("        : " + ", ".join(f'"r"(e[{k}])' for k in range(8)) + ", " + ", ".join(f'"r"(o[{k}])' for k in range(1, 8)) + ");\n")
    if IMM_ROLE != 'x':
        for j in range(8):
            # r += q * (2^32 - 1) * 2^(32 j)  =  r + (q << 32(j+1)) - (q << 32 j)   (mod 2^256)
            w(f"    if constexpr (F::NP({j}) == 0xffffffffu) {{\n")
            if j + 1 <= 7:
                ks = list(range(j + 1, 8))
                lines = []
                for n, k in enumerate(ks):
                    op = "add.cc.u32" if n == 0 else ("addc.cc.u32" if n < len(ks) - 1 else "addc.u32")
                    if len(ks) == 1:
                        op = "add.u32"
                    lines.append(f"{op} %{n}, %{n}, %{len(ks) + n};")
                w('        asm("' + '\\n\\t"\n            "'.join(lines) + '"\n')
                w("            : " + ", ".join(f'"+r"(r[{k}])' for k in ks) + "\n")
                w("            : " + ", ".join(f'"r"(q[{k - j - 1}])' for k in ks) + ");\n")
            ks = list(range(j, 8))
            lines = []
            for n, k in enumerate(ks):
                op = "sub.cc.u32" if n == 0 else ("subc.cc.u32" if n < len(ks) - 1 else "subc.u32")
                if len(ks) == 1:
                    op = "sub.u32"
                lines.append(f"{op} %{n}, %{n}, %{len(ks) + n};")
            w('        asm("' + '\\n\\t"\n            "'.join(lines) + '"\n')
            w("            : " + ", ".join(f'"+r"(r[{k}])' for k in ks) + "\n")
            w("            : " + ", ".join(f'"r"(q[{k - j}])' for k in ks) + ");\n")
            w("    }\n")
    w("}\n#endif  // __CUDA_ARCH__\n\n}  // namespace hodor\n")


if __name__ == "__main__":
    main()
