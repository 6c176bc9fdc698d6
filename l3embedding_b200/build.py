"""In-tree build of libl3b200.so (sm_100a only) with plain nvcc; no torch involved.

    python -m l3embedding_b200.build [--force]

The shared library lands next to this file so that it travels with the source tree.
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libl3b200.so")
OBJ_DIR = os.path.join(HERE, "build")
SOURCES = ["api.cu", "frontend.cu", "elementwise.cu", "conv_simt.cu", "head.cu", "conv_tc.cu", "dp.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
         "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr"]


def _digest(paths):
    h = hashlib.sha256()
    for p in sorted(paths):
        with open(p, "rb") as f:
            h.update(p.encode())
            h.update(f.read())
    h.update(" ".join(FLAGS).encode())
    return h.hexdigest()


def _deps():
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
    hdrs.append(os.path.join(HERE, "..", "include", "l3b200.h"))
    return hdrs


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every .cu for sm_100a (per-file objects, cached by content hash) and link the .so."""
    os.makedirs(OBJ_DIR, exist_ok=True)
    hdr_digest = _digest(_deps())
    objs, jobs = [], []
    for src in SOURCES:
        sp = os.path.join(CSRC, src)
        op = os.path.join(OBJ_DIR, src.replace(".cu", ".o"))
        stamp = op + ".sha"
        dg = _digest([sp]) + hdr_digest
        objs.append(op)
        if force or not os.path.exists(op) or not os.path.exists(stamp) or open(stamp).read() != dg:
            jobs.append((sp, op, stamp, dg))

    def compile_one(job):
        sp, op, stamp, dg = job
        cmd = [NVCC, *FLAGS, "-c", sp, "-o", op]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (sp, r.stdout, r.stderr))
        if verbose:
            sys.stderr.write(r.stderr)
        with open(stamp, "w") as f:
            f.write(dg)

    if jobs:
        with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 1)) as ex:
            list(ex.map(compile_one, jobs))
    if jobs or force or not os.path.exists(LIB):
        r = subprocess.run([NVCC, "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-ldl"],
                           capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
