"""Build libgptq_b200.so in-tree with nvcc for sm_100a (no torch extension, no JIT cache).

    python -m gptq_b200.build [--force]
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
BUILD = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libgptq_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xptxas", "-v"]
# per-file extra flags
EXTRA_ENV = os.environ.get("GPTQ_B200_NVCC_EXTRA", "").split()
EXTRA = {
    "colloop.cu": ["--fmad=false"],      # error feedback must not be contracted (see colloop.cu)
    "find_params.cu": ["--fmad=false"],
    "prepare.cu": ["--fmad=false"],
    "pack.cu": ["--fmad=false"],
}


def sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def _headers():
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hs.append(os.path.join(HERE, "..", "include", "gptq_b200.h"))
    return hs


def compile_one(src, force=False):
    obj = os.path.join(BUILD, src[:-3] + ".o")
    path = os.path.join(CSRC, src)
    if not force and not _stale(obj, [path] + _headers()):
        return obj, ""
    cmd = [NVCC, *ARCH, *COMMON, *EXTRA.get(src, []), *EXTRA_ENV, "-c", path, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src}:\n{' '.join(cmd)}\n{r.stdout}\n{r.stderr}")
    with open(obj + ".log", "w") as f:
        f.write(r.stdout + r.stderr)
    return obj, r.stderr


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(BUILD, exist_ok=True)
    srcs = sources()
    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        results = list(ex.map(lambda s: compile_one(s, force), srcs))
    objs = [o for o, _ in results]
    if verbose:
        for _, log in results:
            if log:
                print(log)
    if force or _stale(LIB, objs):
        tmp = LIB + ".tmp"               # link beside the target, then rename: a reader never sees a half-written library
        cmd = [NVCC, *ARCH, "-shared", "-o", tmp, *objs, "-lcudart"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
        os.replace(tmp, LIB)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
