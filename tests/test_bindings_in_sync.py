"""The three statements of the drop-in boundary must agree: include/hodor_b200.h (the C ABI), the Rust `extern "C"`
block a maintainer adds to the reference crate (rust/src/cuda/ffi.rs -- uncompiled here: no cargo in the image) and the
ctypes table the Python mirror binds (hodor_b200/_ffi.py).  Checked by name, arity and type of every parameter and
of the return value, so a signature cannot drift in one of them unnoticed."""
import ctypes as C
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def c_prototypes():
    s = open(os.path.join(ROOT, "include", "hodor_b200.h")).read()
    s = re.sub(r"/\*.*?\*/", "", s, flags=re.S)
    s = re.sub(r"^\s*#.*$", "", s, flags=re.M)
    out = {}
    for m in re.finditer(r"([A-Za-z_][\w\s\*]*?)\b(hodor_\w+)\s*\(([^;{]*?)\)\s*;", s, flags=re.S):
        ret, name, args = " ".join(m.group(1).split()), m.group(2), " ".join(m.group(3).split())
        if ret.startswith("typedef"):
            continue
        out[name] = (ret, [] if args in ("void", "") else [a.strip() for a in args.split(",")])
    return out


def c_param_type(decl: str) -> str:
    """'const uint64_t omega[4]' -> 'const uint64_t*'; 'void* stream' -> 'void*'."""
    m = re.match(r"^(const\s+)?(\w+)\s+\w+\[\d+\]$", decl)
    if m:
        return f"{m.group(1) or ''}{m.group(2)}*".replace("  ", " ")
    t = re.match(r"^(.*?)(\w+)$", decl).group(1).strip()
    return " ".join(t.split())


RUST_OF_C = {
    "int": "c_int", "uint32_t": "u32", "uint64_t": "u64", "size_t": "usize",
    "const uint64_t*": "*const u64", "uint64_t*": "*mut u64", "const uint8_t*": "*const u8", "uint8_t*": "*mut u8",
    "const void*": "*const c_void", "void*": "*mut c_void", "int*": "*mut c_int", "uint32_t*": "*mut u32",
    "char*": "*mut c_char", "const char*": "*const c_char",
    "const uint64_t* const*": "*const *const u64", "uint64_t* const*": "*const *mut u64",
    "uint8_t**": "*mut *mut u8", "uint64_t**": "*mut *mut u64",
    "hodor_tree**": "*mut *mut Tree", "const hodor_tree*": "*const Tree", "hodor_tree*": "*mut Tree",
    "const hodor_fri_proto*": "*const FriProto", "hodor_fri_proto*": "*mut FriProto",
}

u64p, u8p, u32p, vp = C.POINTER(C.c_uint64), C.POINTER(C.c_uint8), C.POINTER(C.c_uint32), C.c_void_p
# ctypes spellings the Python table may use for a C type (void* stands in for opaque handles and device pointers)
CTYPES_OF_C = {
    "int": {C.c_int}, "uint32_t": {C.c_uint32}, "uint64_t": {C.c_uint64}, "size_t": {C.c_size_t},
    "const uint64_t*": {u64p, vp}, "uint64_t*": {u64p, vp}, "const uint8_t*": {u8p}, "uint8_t*": {u8p},
    "const void*": {vp}, "void*": {vp}, "int*": {C.POINTER(C.c_int), vp}, "uint32_t*": {u32p},
    "char*": {C.c_char_p}, "const char*": {C.c_char_p},
    "const uint64_t* const*": {C.POINTER(u64p), C.POINTER(vp)}, "uint64_t* const*": {C.POINTER(u64p)},
    "uint8_t**": {C.POINTER(u8p)}, "uint64_t**": {C.POINTER(u64p)},
    "hodor_tree**": {C.POINTER(vp)}, "const hodor_tree*": {vp}, "hodor_tree*": {vp},
    "const hodor_fri_proto*": {vp}, "hodor_fri_proto*": {vp}, "void": {None},
}


def rust_prototypes():
    s = open(os.path.join(ROOT, "rust", "src", "cuda", "ffi.rs")).read()
    block = s[s.index('extern "C" {'):]
    block = block[:block.index("\n}\n")]
    block = re.sub(r"//.*$", "", block, flags=re.M)
    out = {}
    for m in re.finditer(r"pub fn (hodor_\w+)\s*\((.*?)\)\s*(?:->\s*([^;]+?))?\s*;", block, flags=re.S):
        params = [p.strip() for p in m.group(2).split(",") if p.strip()]
        out[m.group(1)] = (m.group(3).strip() if m.group(3) else None, [p.split(":", 1)[1].strip() for p in params])
    return out


def test_header_parses_to_the_expected_number_of_entry_points():
    protos = c_prototypes()
    assert len(protos) >= 90 and "hodor_cuda_lde_fri_sharded" in protos and "hodor_cuda_init" in protos
    assert protos["hodor_cuda_ntt"] == ("int", ["uint64_t* a", "uint32_t log_n", "const uint64_t omega[4]", "int field_id"])


def test_rust_extern_block_matches_the_header():
    c, r = c_prototypes(), rust_prototypes()
    assert set(c) == set(r), (sorted(set(c) - set(r)), sorted(set(r) - set(c)))
    for name, (ret, params) in c.items():
        r_ret, r_params = r[name]
        want_ret = None if ret == "void" else RUST_OF_C[ret]
        assert r_ret == want_ret, (name, r_ret, want_ret)
        want = [RUST_OF_C[c_param_type(p)] for p in params]
        assert r_params == want, (name, r_params, want)


def test_rust_constants_match_the_header():
    h = open(os.path.join(ROOT, "include", "hodor_b200.h")).read()
    r = open(os.path.join(ROOT, "rust", "src", "cuda", "ffi.rs")).read()
    defines = {m.group(1): int(m.group(2).strip("()")) for m in re.finditer(r"#define (HODOR_\w+) (\(?-?\d+\)?)", h)}
    consts = {m.group(1): int(m.group(2)) for m in re.finditer(r"pub const (\w+): c_int = (-?\d+);", r)}
    for name, value in defines.items():
        short = name[len("HODOR_"):]
        if short.startswith("OP_"):
            continue  # op codes are spelled in poly.rs
        assert consts.get(short) == value, (name, value, consts.get(short))


def test_python_ctypes_table_matches_the_header():
    from hodor_b200 import _ffi

    c = c_prototypes()
    assert set(c) == set(_ffi._SIGS), (sorted(set(c) - set(_ffi._SIGS)), sorted(set(_ffi._SIGS) - set(c)))
    for name, (ret, params) in c.items():
        res, args = _ffi._SIGS[name]
        ok_ret = CTYPES_OF_C[ret] if ret in CTYPES_OF_C else {vp}
        if ret in ("void*", "const void*", "hodor_tree*", "hodor_fri_proto*"):
            ok_ret = {vp}
        assert res in ok_ret, (name, res, ret)
        assert len(args) == len(params), (name, len(args), len(params))
        for a, p in zip(args, params):
            assert a in CTYPES_OF_C[c_param_type(p)], (name, p, a)


def _rust_sources():
    d = os.path.join(ROOT, "rust", "src")
    for dirpath, _, files in os.walk(d):
        for f in sorted(files):
            if f.endswith(".rs"):
                yield os.path.join(dirpath, f)


def _strip_rust_comments_and_strings(s: str) -> str:
    s = re.sub(r"//[^\n]*", "", s)
    s = re.sub(r"/\*.*?\*/", "", s, flags=re.S)
    s = re.sub(r'"(?:[^"\\]|\\.)*"', '""', s, flags=re.S)  # `\` + newline continues a string literal
    s = re.sub(r"'(?:[^'\\]|\\.)'", "' '", s)  # char literals ('S'), not lifetimes
    return s


def test_rust_files_are_lexically_balanced():
    """No compiler here: at least every bracket of every committed .rs file closes."""
    pairs = {")": "(", "]": "[", "}": "{"}
    for path in _rust_sources():
        stack = []
        for ch in _strip_rust_comments_and_strings(open(path).read()):
            if ch in "([{":
                stack.append(ch)
            elif ch in pairs:
                assert stack and stack.pop() == pairs[ch], path
        assert not stack, path


def test_rust_calls_of_the_c_abi_have_the_declared_arity():
    """Every `hodor_*(..)` call in the shim passes as many arguments as ffi.rs (== the header) declares."""
    declared = {name: len(params) for name, (_, params) in rust_prototypes().items()}
    seen = set()
    for path in _rust_sources():
        if path.endswith("ffi.rs"):
            text = _strip_rust_comments_and_strings(open(path).read())
            text = text[:text.index('extern "" {')] + text[text.index("\n}\n", text.index('extern "" {')):]  # skip the declarations
        else:
            text = _strip_rust_comments_and_strings(open(path).read())
        for m in re.finditer(r"\b(hodor_\w+)\s*\(", text):
            name, i, depth, commas, any_arg = m.group(1), m.end(), 1, 0, False
            while depth:
                ch = text[i]
                if ch in "([{":
                    depth += 1
                elif ch in ")]}":
                    depth -= 1
                elif ch == "," and depth == 1:
                    commas += 1
                if depth >= 1 and not ch.isspace():
                    any_arg = True
                i += 1
            args = text[m.end():i - 1]
            if args.rstrip().endswith(","):
                commas -= 1  # trailing comma of a multi-line call
            n = commas + 1 if any_arg else 0
            assert name in declared, (path, name)
            assert n == declared[name], (path, name, n, declared[name])
            seen.add(name)
    assert {"hodor_cuda_lde_fri_sharded", "hodor_cuda_ntt_sharded", "hodor_cuda_fri_produce_proof", "hodor_cuda_poly_op",
            "hodor_cuda_lde_commit", "hodor_cuda_merkle_build"} <= seen


def test_every_environment_switch_is_documented():
    """INTEGRATION.md's table of environment switches == the HODOR_* variables the library, the Python mirror and
    bench.py actually read."""
    read = set()
    csrc = os.path.join(ROOT, "hodor_b200", "csrc")
    for f in os.listdir(csrc):
        if os.path.isfile(os.path.join(csrc, f)):
            read |= set(re.findall(r'getenv\("(HODOR_[A-Z0-9_]+)"\)', open(os.path.join(csrc, f)).read()))
    for path in [os.path.join(ROOT, "bench.py")] + [os.path.join(ROOT, "hodor_b200", f) for f in os.listdir(os.path.join(ROOT, "hodor_b200"))
                                                     if f.endswith(".py")]:
        read |= set(re.findall(r'environ(?:\.get)?[\[(]\s*"(HODOR_[A-Z0-9_]+)"', open(path).read()))
    documented = set(re.findall(r"`(HODOR_[A-Z0-9_]+)`", open(os.path.join(ROOT, "INTEGRATION.md")).read()))
    assert read <= documented, sorted(read - documented)
    assert documented <= read | {"HODOR_B200_LIB_DIR", "HODOR_CUDA_DEVICE"}, sorted(documented - read)  # the two are read by rust/
