"""Compile libgolfer_b200.so in-tree with nvcc for sm_100a (no torch extension, no
torch types in the ABI).  Output: <package>/lib/libgolfer_b200.so — git-ignored,
but it travels to the GPU box with the gpurun snapshot."""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT_DIR = os.path.join(HERE, "lib")
OBJ_DIR = os.path.join(HERE, "build")
LIB = os.path.join(OUT_DIR, "libgolfer_b200.so")
SOURCES = ["api.cu", "align.cu", "align_embed.cu", "pose.cu", "segment_fp32.cu", "segment_bf16.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr",
    # instrumentation builds only (e.g. GOLFER_NVCC_EXTRA=-DGOLFER_TCN_TRACE with --force; tools/trace_tcn.py)
    *os.environ.get("GOLFER_NVCC_EXTRA", "").split(),
]


def _deps():
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hdrs.append(os.path.join(os.path.dirname(HERE), "include", "golfer_b200.h"))
    return hdrs


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OUT_DIR, exist_ok=True)
    os.makedirs(OBJ_DIR, exist_ok=True)
    hdr_mtime = max(os.path.getmtime(h) for h in _deps())

    def compile_one(src):
        s = os.path.join(CSRC, src)
        o = os.path.join(OBJ_DIR, src.replace(".cu", ".o"))
        if (not force and os.path.exists(o)
                and os.path.getmtime(o) >= max(os.path.getmtime(s), hdr_mtime)):
            return o, False
        cmd = ["nvcc", *NVCC_FLAGS, "-c", s, "-o", o]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            sys.stderr.write(r.stderr)
        return o, True

    with ThreadPoolExecutor(max_workers=4) as ex:
        results = list(ex.map(compile_one, SOURCES))
    objs = [o for o, _ in results]
    if force or any(ch for _, ch in results) or not os.path.exists(LIB):
        r = subprocess.run(["nvcc", "-shared", "-gencode", "arch=compute_100a,code=sm_100a", *objs,
                            "-o", LIB, "-cudart", "static"], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
