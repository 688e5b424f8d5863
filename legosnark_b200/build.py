"""Builds legosnark_b200/libb200msm.so (the C-ABI library of include/b200_msm.h)
in-tree with nvcc for sm_100a.  Five translation units compiled in parallel."""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libb200msm.so")
OBJ = os.path.join(HERE, "build")
UNITS = ["engine_core", "engine_g1", "engine_g2", "engine_fr", "engine_wire", "engine_sort"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
         "-Xcompiler", "-fPIC", "-Xcompiler", "-O2"]


def _deps():
    return [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "b200_msm.h")]


def needs_build() -> bool:
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    return any(os.path.getmtime(d) > t for d in _deps())


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return OUT
    os.makedirs(OBJ, exist_ok=True)
    extra = ["-Xptxas", "-v"] if verbose else []

    def compile_unit(u):
        cmd = [NVCC] + FLAGS + extra + ["-c", os.path.join(CSRC, u + ".cu"), "-o", os.path.join(OBJ, u + ".o")]
        r = subprocess.run(cmd, capture_output=True, text=True)
        return u, r

    with ThreadPoolExecutor(len(UNITS)) as ex:
        results = list(ex.map(compile_unit, UNITS))
    for u, r in results:
        if verbose or r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {u}.cu")
    cmd = [NVCC, "-shared", "-o", OUT] + [os.path.join(OBJ, u + ".o") for u in UNITS] + ["-lcudart_static", "-lpthread", "-ldl", "-lrt"]
    subprocess.run(cmd, check=True)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
