"""Build libunit_b200.so (all CUDA kernels + the C ABI) in-tree with nvcc for sm_100a.

    python -m unit_b200.build [--force] [--verbose]

The library has no torch / Python dependency: plain `nvcc -shared`, static cudart.  It is git-ignored but
travels to the GPU box with the repo snapshot, so nothing is compiled there.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libunit_b200.so")
STAMP = os.path.join(HERE, ".libunit_b200.stamp")  # git-ignored convenience only; the .so carries its own digest
SOURCES = ["api.cu", "roi_align.cu", "roi_align_fwd.cu", "roi_align_fwd_band.cu", "roi_align_bwd.cu", "roi_align_bwd_cl.cu", "roi_align_bwd_cl2.cu", "match.cu", "detect.cu", "transfer.cu", "gemm.cu", "weak.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr",
]


def _nvcc() -> str:
    cand = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(cand):
        raise RuntimeError("nvcc not found; libunit_b200.so cannot be built")
    return cand


def source_digest() -> str:
    """sha256 over csrc/, the public header and the compile flags.  It is compiled into the library
    (``unit_source_digest()``), so ``_lib.lib()`` can tell a stale .so from a fresh one without any side file."""
    h = hashlib.sha256()
    files = sorted(os.listdir(CSRC)) + ["../../include/unit_b200.h"]
    for f in files:
        p = os.path.join(CSRC, f)
        if os.path.isfile(p):
            h.update(f.encode())
            h.update(open(p, "rb").read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def embedded_digest(path: str = LIB):
    """Digest compiled into an existing library, or None (missing / pre-digest build).  Read from the file's bytes
    (marker string), not through dlopen: a stale library must not be mapped before it is replaced."""
    if not os.path.exists(path):
        return None
    marker = b"UNIT_SOURCE_DIGEST="
    with open(path, "rb") as f:
        blob = f.read()
    i = blob.find(marker)
    if i < 0:
        return None
    return blob[i + len(marker): i + len(marker) + 64].decode("ascii", "replace")


def build(force: bool = False, verbose: bool = False) -> str:
    digest = source_digest()
    if not force and embedded_digest() == digest:
        return LIB
    nvcc = _nvcc()
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    procs = []
    for src in SOURCES:
        sp = os.path.join(CSRC, src)
        if not os.path.exists(sp):
            continue
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        extra = [f'-DUNIT_SOURCE_DIGEST="{digest}"'] if src == "api.cu" else []
        cmd = [nvcc] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-c", sp, "-o", obj]
        procs.append((src, obj, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    objs = []
    for src, obj, pr in procs:
        out, _ = pr.communicate()
        if verbose or pr.returncode != 0:
            sys.stderr.write(out)
        if pr.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}")
        objs.append(obj)
    cmd = [nvcc, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout)
        raise RuntimeError("link of libunit_b200.so failed")
    with open(STAMP, "w") as f:
        f.write(digest)
    return LIB


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(path)
