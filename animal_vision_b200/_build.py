"""Build libavb200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m animal_vision_b200._build [--force] [--verbose]
"""
from __future__ import annotations

import glob
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "lib", "libavb200.so")
STAMP = LIB + ".srchash"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xptxas=-v",
    "-Xcompiler", "-fPIC,-O2,-fvisibility=hidden",
    "-shared",
]


def _extra_flags():
    """Extra nvcc flags for experiments (e.g. AVB_NVCC_EXTRA="-DUV_MINB=2"); part of the source hash."""
    return os.environ.get("AVB_NVCC_EXTRA", "").split()


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libavb200.so cannot be built (there is no CPU fallback)")


HEADER = os.path.join(os.path.dirname(HERE), "include", "avb200.h")


def header_sha() -> str:
    """What avb_header_sha() of a library built from the current header returns."""
    with open(HEADER, "rb") as fh:
        return hashlib.sha256(fh.read()).hexdigest()[:16]


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def _src_hash() -> str:
    h = hashlib.sha256()
    for f in sources() + sorted(glob.glob(os.path.join(CSRC, "*.cuh"))) + \
            [os.path.join(os.path.dirname(HERE), "include", "avb200.h"), os.path.abspath(__file__)]:
        h.update(os.path.relpath(f, HERE).encode())
        with open(f, "rb") as fh:
            h.update(fh.read())
    h.update(" ".join(_extra_flags()).encode())
    return h.hexdigest()


def is_current() -> bool:
    try:
        return os.path.exists(LIB) and open(STAMP).read().strip() == _src_hash()
    except OSError:
        return False


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every csrc/*.cu into one shared library. Returns the library path."""
    if not force and is_current():
        return LIB
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    objs = []
    objdir = os.path.join(HERE, "lib", "obj")
    os.makedirs(objdir, exist_ok=True)
    nvcc = _nvcc()
    procs = []
    for src in sources():
        obj = os.path.join(objdir, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        cmd = [nvcc, "-c", src, "-o", obj, f'-DAVB_HEADER_SHA="{header_sha()}"'] + [f for f in NVCC_FLAGS if f != "-shared"] + _extra_flags()
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    log = []
    for src, p in procs:
        out, _ = p.communicate()
        log.append(f"== {os.path.basename(src)}\n{out}")
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{out}")
    cmd = [nvcc, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}")
    with open(os.path.join(HERE, "lib", "build.log"), "w") as fh:
        fh.write("\n".join(log))
    with open(STAMP, "w") as fh:
        fh.write(_src_hash())
    if verbose:
        print("\n".join(log))
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
