"""Builds and runs tests/cpp/test_mirror.cpp: the reference-shaped parity tests written against the
C++ host mirror (include/hodor_b200.hpp), the host language for a compiled reference whose own
toolchain (Rust) is absent.  The compile-and-link check runs on the CPU box; the run needs a B200."""
import os
import subprocess

import pytest

from conftest import ROOT


def build_binary(tmpdir):
    from oracle import oracle as O
    O.build()
    exe = os.path.join(str(tmpdir), "test_mirror")
    subprocess.check_call([
        "g++", "-std=c++17", "-O2", "-Wall", "-I", os.path.join(ROOT, "include"),
        os.path.join(ROOT, "tests", "cpp", "test_mirror.cpp"), "-o", exe,
        "-L", os.path.join(ROOT, "hodor_b200"), "-lhodor_b200",
        "-L", os.path.join(ROOT, "oracle", "_build"), "-lhodor_oracle",
        "-Wl,-rpath," + os.path.join(ROOT, "hodor_b200"), "-Wl,-rpath," + os.path.join(ROOT, "oracle", "_build"),
    ])
    return exe


def test_cpp_mirror_compiles_and_links(tmp_path):
    exe = build_binary(tmp_path)
    assert os.path.exists(exe)
    # without a GPU the binary must refuse loudly (exit 2: init throws), never compute on the CPU
    from hodor_b200 import _ffi
    if _ffi.lib.hodor_cuda_device_count() == 0:
        r = subprocess.run([exe], capture_output=True, text=True)
        assert r.returncode == 2 and "no CUDA device" in r.stdout


@pytest.mark.gpu
def test_cpp_mirror_parity(tmp_path):
    exe = build_binary(tmp_path)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    print(r.stdout)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "0 failure(s)" in r.stdout
