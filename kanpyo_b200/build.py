"""Builds libkanpyo_b200.so (the C-ABI library of include/kanpyo_b200.h) in-tree with nvcc for sm_100a.

    python -m kanpyo_b200.build [--force]

The library links the CUDA runtime statically, so it has no dependency on PyTorch or on the
libcudart that PyTorch ships.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libkanpyo_b200.so")
# tuning experiments: KP_VARIANT=name KP_DEFINES="-DKP_X=1 ..." python -m kanpyo_b200.build builds
# kanpyo_b200/_variants/libkanpyo_b200.<name>.so; KANPYO_B200_LIB=<path> makes _lib.load() use it.
VARIANT_DIR = os.path.join(HERE, "_variants")
SOURCES = ["kp_dict.cu", "kp_kernels.cu", "kp_fused.cu", "kp_api.cu", "kp_queue.cu", "kp_dictbuild.cpp"]
HEADERS = ["kp_common.cuh", "kp_kernels.cuh", "kp_exports.map", os.path.join("..", "..", "include", "kanpyo_b200.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC,-Wall,-Wno-unused-function", "-Xptxas", "-v", "-shared", "-cudart", "static", "-ldl", "-lpthread",
              "-Xlinker", "--version-script=" + os.path.join(CSRC, "kp_exports.map")]


def nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; kanpyo_b200 cannot be built (there is no CPU fallback)")


def stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.exists(d) and os.path.getmtime(d) > t for d in deps)


def build_variant(name: str, defines) -> str:
    os.makedirs(VARIANT_DIR, exist_ok=True)
    out = os.path.join(VARIANT_DIR, "libkanpyo_b200.%s.so" % name)
    srcs = [os.path.join(CSRC, f) for f in SOURCES]
    r = subprocess.run([nvcc()] + NVCC_FLAGS + list(defines) + ["-o", out] + srcs, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc failed building variant %s" % name)
    return out


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not stale():
        return LIB
    srcs = [os.path.join(CSRC, f) for f in SOURCES if os.path.exists(os.path.join(CSRC, f))]
    cmd = [nvcc()] + NVCC_FLAGS + ["-o", LIB + ".tmp"] + srcs
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc failed building libkanpyo_b200.so")
    os.replace(LIB + ".tmp", LIB)
    log = os.path.join(HERE, "_build_ptxas.log")
    with open(log, "w") as f:
        f.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
    if verbose:
        sys.stderr.write(r.stdout + r.stderr)
    return LIB


if __name__ == "__main__":
    if os.environ.get("KP_VARIANT"):
        print(build_variant(os.environ["KP_VARIANT"], os.environ.get("KP_DEFINES", "").split()))
    else:
        print(build(force="--force" in sys.argv, verbose=True))
