"""In-tree build of libinmf_b200.so with nvcc for sm_100a (cross-compiles without a GPU)."""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB_PATH = os.path.join(CSRC, "libinmf_b200.so")
SOURCES = ["inmf_kernels.cu"]
HEADERS = ["inmf_common.cuh", "bpp.cuh", "blocked.cuh", "sell.cuh", "bpp16.cuh", "quantile.cuh", "bpp_fast.cuh", "tc_common.cuh", "tc_pass.cuh", "preprocess.cuh", "blockfmt.cuh", os.path.join(ROOT, "include", "inmf_b200.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
    "-I", os.path.join(ROOT, "include"), "-I", CSRC,
]


def _nvcc():
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found; cannot build libinmf_b200.so")
    return exe


def is_stale():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, s) for s in SOURCES]
    deps += [h if os.path.isabs(h) else os.path.join(CSRC, h) for h in HEADERS]
    return any(os.path.getmtime(d) > t for d in deps)


def build_library(force=False, verbose=False, extra_flags=()):
    if not force and not is_stale():
        return LIB_PATH
    cmd = [_nvcc()] + NVCC_FLAGS + list(extra_flags) + ["-o", LIB_PATH]
    cmd += [os.path.join(CSRC, s) for s in SOURCES]
    if verbose:
        print(" ".join(cmd), file=sys.stderr)
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose and (res.stdout or res.stderr):
        print(res.stdout + res.stderr, file=sys.stderr)
    return LIB_PATH


if __name__ == "__main__":
    build_library(force=True, verbose=True, extra_flags=sys.argv[1:])
    print(LIB_PATH)
