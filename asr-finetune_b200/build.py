"""In-tree build of libwfe.so for sm_100a (nvcc cross-compiles without a GPU)."""
from __future__ import annotations

import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libwfe.so")
SOURCES = ["wfe_api.cu", "wfe_codelets.cuh", "wfe_logmel.cuh", "wfe_logmel_tc.cuh", "wfe_tc_epilogue_gen.inc", "wfe_collate.cuh", "wfe_copy_pool.h",
           os.path.join("..", "..", "include", "wfe.h")]


def is_stale() -> bool:
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    return any(os.path.getmtime(os.path.join(CSRC, s)) > t for s in SOURCES)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not is_stale():
        return OUT
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: cannot build the sm_100a extension")
    cmd = [nvcc, "-O3", "-std=c++17", "-lineinfo", "-gencode", "arch=compute_100a,code=sm_100a",
           "-Xcompiler", "-fPIC,-O3", "-Xptxas", "-v", "-shared", "-o", OUT, "wfe_api.cu", "-lcudart"]
    res = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        print(res.stdout + res.stderr)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed building libwfe.so")
    return OUT


if __name__ == "__main__":
    print(build(force=True, verbose=True))
