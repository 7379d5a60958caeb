"""Builds libqbm_b200.so in-tree with nvcc for sm_100a (no torch extension machinery: the library
is a plain C-ABI shared object loaded through ctypes)."""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libqbm_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
]


def _sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _deps():
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hdrs.append(os.path.join(HERE, "..", "include", "qbm_b200.h"))
    return hdrs


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    files = [os.path.join(CSRC, f) for f in _sources()] + _deps() + [__file__]
    return any(os.path.getmtime(f) > t for f in files)


def build(force: bool = False, verbose: bool = False, variant: str = "", extra: tuple = ()) -> str:
    """`variant` / `extra`: an instrumented build (own object directory and library name, extra nvcc flags)."""
    if variant:
        return _build(True, verbose, os.path.join(HERE, "build_" + variant), os.path.join(HERE, f"libqbm_b200_{variant}.so"), tuple(extra))
    if not force and not needs_build():
        return LIB
    return _build(force, verbose, OBJ, LIB, ())


def _build(force, verbose, OBJ, LIB, extra) -> str:
    nvcc = os.environ.get("NVCC", "nvcc")
    os.makedirs(OBJ, exist_ok=True)
    hdr_t = max(os.path.getmtime(f) for f in _deps() + [__file__])

    def compile_one(src):
        obj = os.path.join(OBJ, src[:-3] + ".o")
        s = os.path.join(CSRC, src)
        if not force and os.path.exists(obj) and os.path.getmtime(obj) > max(os.path.getmtime(s), hdr_t):
            return obj
        cmd = [nvcc, *NVCC_FLAGS, *extra, *os.environ.get("QBM_NVCC_EXTRA", "").split(), "-c", s, "-o", obj]
        if verbose:
            print(" ".join(cmd), file=sys.stderr)
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        return obj

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        objs = list(ex.map(compile_one, _sources()))
    cmd = [nvcc, "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
