"""Build libfnssl_b200.so (the C-ABI library of include/fnssl_b200.h) in-tree with nvcc for sm_100a.

    python -m fn_ssl_b200.build          # or __graft_entry__.build()

No torch linkage: the library only depends on the (statically linked) CUDA runtime, so it loads on a
CPU-only box too (symbol checks) and travels to the GPU box with the repository snapshot.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
BUILD = os.path.join(HERE, "csrc", "build")
LIB = os.path.join(HERE, "libfnssl_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-Xptxas", "-v",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: cannot build libfnssl_b200.so")


def have_nvcc() -> bool:
    try:
        _nvcc()
        return True
    except RuntimeError:
        return False


def stale_sources():
    """Sources whose object stamp does not match their current digest (empty list = the library matches the tree)."""
    out = []
    for src in _sources():
        stamp = os.path.join(BUILD, src[:-3] + ".o.sha")
        if not os.path.exists(stamp) or open(stamp).read() != _digest(os.path.join(CSRC, src)):
            out.append(src)
    return out


def _sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _digest(path: str) -> str:
    h = hashlib.sha256()
    for dep in [path] + [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith(".cuh")] + [
            os.path.join(HERE, "..", "include", "fnssl_b200.h")]:
        with open(dep, "rb") as fh:
            h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(BUILD, exist_ok=True)
    nvcc = _nvcc()
    objs, jobs = [], []
    for src in _sources():
        path = os.path.join(CSRC, src)
        obj = os.path.join(BUILD, src[:-3] + ".o")
        stamp = obj + ".sha"
        dig = _digest(path)
        objs.append(obj)
        if (not force and os.path.exists(obj) and os.path.exists(stamp) and open(stamp).read() == dig):
            continue
        jobs.append((src, path, obj, stamp, dig))

    def compile_one(job):
        src, path, obj, stamp, dig = job
        cmd = [nvcc] + NVCC_FLAGS + ["-c", path, "-o", obj]
        res = subprocess.run(cmd, capture_output=True, text=True)
        with open(obj + ".log", "w") as fh:
            fh.write(" ".join(cmd) + "\n" + res.stdout + res.stderr)
        if res.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{res.stdout}\n{res.stderr}")
        with open(stamp, "w") as fh:
            fh.write(dig)
        return src, res.stderr

    if jobs:
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            for src, log in ex.map(compile_one, jobs):
                if verbose:
                    print(f"[fnssl build] {src}\n{log}")
    if jobs or force or not os.path.exists(LIB):
        cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + objs
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError(f"link failed:\n{res.stdout}\n{res.stderr}")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
