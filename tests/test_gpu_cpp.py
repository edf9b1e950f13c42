"""Builds tests/cpp/test_tokenizer.cpp (the reference's src/tests.rs restated against the C++ host
mirror include/kanpyo_b200.hpp) with g++ and runs it.  CPU: compile + link only, and the binary must
fail with KP_ERR_CUDA (exit 77) instead of falling back.  GPU: all assertions must hold."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def cpp_binary(tmp_path_factory):
    from kanpyo_b200 import _lib
    _lib.load()
    out = str(tmp_path_factory.mktemp("cpp") / "test_tokenizer")
    libdir = os.path.dirname(_lib.lib_path())
    cmd = ["g++", "-std=c++17", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "tests", "cpp", "test_tokenizer.cpp"), "-o", out, "-L", libdir, "-lkanpyo_b200",
           "-Wl,-rpath," + libdir]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return out


def test_cpp_mirror_compiles_and_refuses_cpu(cpp_binary):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by test_cpp_reference_tests")
    r = subprocess.run([cpp_binary], capture_output=True, text=True)
    assert r.returncode == 77, (r.returncode, r.stdout, r.stderr)     # KP_ERR_CUDA, no CPU fallback


@pytest.mark.gpu
def test_cpp_reference_tests(cpp_binary):
    r = subprocess.run([cpp_binary, "0"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "cpp tests ok" in r.stdout, (r.returncode, r.stdout, r.stderr)
