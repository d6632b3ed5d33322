"""Build the oracle's own plain-C restatement (oracle/align_c.c) with gcc.

There is no `oracle/_ref`: the reference ships no source to compile
(SURVEY.md section 8c), so `cpu_baseline.kind` is always "port".
"""
from __future__ import annotations

import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
OUT_DIR = os.path.join(HERE, "_build")
LIB = os.path.join(OUT_DIR, "liboracle_align.so")


def build(force: bool = False) -> str:
    src = os.path.join(HERE, "align_c.c")
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(src):
        return LIB
    os.makedirs(OUT_DIR, exist_ok=True)
    cmd = ["gcc", "-O2", "-ffp-contract=off", "-fno-fast-math", "-fopenmp", "-shared", "-fPIC",
           src, "-o", LIB, "-lm"]
    subprocess.run(cmd, check=True)
    return LIB


if __name__ == "__main__":
    print(build(force=True))
