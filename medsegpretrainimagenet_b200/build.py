"""In-tree build of libmsp_b200.so (hand-written sm_100a kernels + C ABI) with nvcc.

`python -m medsegpretrainimagenet_b200.build` or `build_lib()`.  The shared object is written next to
this file so that it travels to the GPU box with the source snapshot; nothing is JIT-compiled at
import time and there is no fallback when it is missing (see _lib.py).
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
BUILD = PKG.parent / "build" / "msp"
LIB = PKG / "libmsp_b200.so"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
    "-Xptxas", "-v",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.sep not in cand or os.path.exists(cand)):
            return cand
    raise RuntimeError("nvcc not found")


def _digest(paths) -> str:
    h = hashlib.sha256()
    for p in sorted(paths):
        h.update(p.name.encode())
        h.update(p.read_bytes())
    h.update((" ".join(NVCC_FLAGS) + " cudart=shared").encode())
    return h.hexdigest()


def sources():
    return sorted(CSRC.glob("*.cu"))


def build_lib(force: bool = False, verbose: bool = False) -> Path:
    srcs = sources()
    deps = srcs + sorted(CSRC.glob("*.cuh")) + sorted((PKG.parent / "include").glob("*.h"))
    stamp = BUILD / "digest.txt"
    dig = _digest(deps)
    if not force and LIB.exists() and stamp.exists() and stamp.read_text() == dig:
        return LIB
    BUILD.mkdir(parents=True, exist_ok=True)
    nvcc = _nvcc()

    def compile_one(src: Path):
        obj = BUILD / (src.stem + ".o")
        cmd = [nvcc, *NVCC_FLAGS, "-c", str(src), "-o", str(obj)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        (BUILD / (src.stem + ".ptxas.log")).write_text(r.stderr)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src.name}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            sys.stderr.write(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(compile_one, srcs))
    # SHARED cudart (the process's one runtime: torch's libcudart.so.12 when torch is imported first, else the
    # toolkit's through the rpath) - a statically linked runtime would carry every runtime entry point's name
    # into the shipped .so whether or not anything calls it.  libcuda is resolved lazily through
    # cudaGetDriverEntryPoint and is NOT linked, so the library still loads (symbol checks) without a driver.
    cmd = [nvcc, "-shared", "--cudart", "shared", "-Xlinker", "-rpath,/usr/local/cuda/lib64",
           "-o", str(LIB), *map(str, objs)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    stamp.write_text(dig)
    return LIB


if __name__ == "__main__":
    print(build_lib(force="--force" in sys.argv, verbose=True))
