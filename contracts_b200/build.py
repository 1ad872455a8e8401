"""In-tree build of libssd_b200.so with nvcc for sm_100a (no JIT cache: the .so travels with the tree)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC = os.path.join(HERE, "csrc", "ssd_b200.cu")
OUT = os.path.join(HERE, "libssd_b200.so")

NVCC_FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-fmad=false",                      # float64 reward/transfer arithmetic must round like the reference
    "-Xcompiler", "-fPIC", "-shared", "-I", os.path.join(ROOT, "include"),
]


def _sources():
    d = os.path.join(HERE, "csrc")
    return [os.path.join(d, f) for f in os.listdir(d)] + [os.path.join(ROOT, "include", "ssd_b200.h")]


def build(force=False, verbose=False, extra_flags=()):
    if not force and os.path.exists(OUT) and all(os.path.getmtime(OUT) >= os.path.getmtime(s) for s in _sources()):
        return OUT
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    if not os.path.exists(nvcc):
        nvcc = "nvcc"
    cmd = [nvcc] + NVCC_FLAGS + list(extra_flags) + ["-o", OUT, SRC]
    if verbose:
        print(" ".join(cmd))
    env = dict(os.environ)
    env.pop("CC", None)                 # the image's $CC (/opt/gcc) is not a usable nvcc host compiler setting
    env.pop("CXX", None)
    subprocess.run(cmd, check=True, env=env)
    return OUT
