"""In-tree build of the sm_100a C-ABI library (``mvuld_b200/libmvuld_b200.so``) with nvcc.

nvcc cross-compiles without a GPU; the .so travels to the GPU box with the repo snapshot.
"""
from __future__ import annotations

import concurrent.futures as cf
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJDIR = os.path.join(HERE, "csrc", "build")
LIB = os.path.join(HERE, "libmvuld_b200.so")
PROBE_LIB = os.path.join(HERE, "libmvuld_probe.so")          # test fixture (csrc/testlib), not part of the product library

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "--expt-relaxed-constexpr", "-Xcompiler", "-fPIC", "-Xcompiler", "-Wall", "-Xcompiler", "-Wno-unused-function",
    "-I", os.path.join(os.path.dirname(HERE), "include"),
] + (os.environ.get("MVULD_NVCC_EXTRA", "").split())


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _deps_digest() -> str:
    h = hashlib.sha1()
    for root in (CSRC, os.path.join(os.path.dirname(HERE), "include")):
        for f in sorted(os.listdir(root)):
            if f.endswith((".cuh", ".h")):
                with open(os.path.join(root, f), "rb") as fh:
                    h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def _compile_one(src: str, digest: str, verbose: bool) -> str:
    obj = os.path.join(OBJDIR, os.path.basename(src)[:-3] + ".o")
    stamp = obj + ".stamp"
    with open(src, "rb") as fh:
        key = hashlib.sha1(fh.read()).hexdigest() + digest
    if os.path.exists(obj) and os.path.exists(stamp) and open(stamp).read() == key:
        return obj
    cmd = [_nvcc(), *NVCC_FLAGS, "-c", src, "-o", obj]
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
    if verbose:
        sys.stderr.write(r.stderr)
    with open(stamp, "w") as fh:
        fh.write(key)
    return obj


def build(verbose: bool = False, force: bool = False) -> str:
    os.makedirs(OBJDIR, exist_ok=True)
    if force:
        for f in os.listdir(OBJDIR):
            if os.path.isfile(os.path.join(OBJDIR, f)):       # (build/variants/ holds experiment builds: left alone)
                os.remove(os.path.join(OBJDIR, f))
    digest = _deps_digest()
    srcs = sources()
    with cf.ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(lambda s: _compile_one(s, digest, verbose), srcs))
    newest = max(os.path.getmtime(o) for o in objs)
    if not os.path.exists(LIB) or os.path.getmtime(LIB) < newest:
        cmd = [_nvcc(), "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    _build_probe(digest, verbose, os.path.join(OBJDIR, "host_util.o"))
    return LIB


def _build_probe(digest: str, verbose: bool, host_util_obj: str) -> str:
    """The UMMA / TMA layout probe the kernel tests use: one more .so next to the product library."""
    obj = _compile_one(os.path.join(CSRC, "testlib", "probe.cu"), digest, verbose)
    if not os.path.exists(PROBE_LIB) or os.path.getmtime(PROBE_LIB) < max(os.path.getmtime(obj), os.path.getmtime(host_util_obj)):
        cmd = [_nvcc(), "-shared", "-o", PROBE_LIB, obj, host_util_obj, "-gencode", "arch=compute_100a,code=sm_100a"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return PROBE_LIB


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv, force="-f" in sys.argv))
