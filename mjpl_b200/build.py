"""Build ``libmjpl_b200.so`` in-tree with nvcc for sm_100a (no JIT cache, no pip install)."""

from __future__ import annotations

import os
import shutil
import subprocess
from pathlib import Path

_PKG = Path(__file__).resolve().parent
SOURCES = [_PKG / "csrc" / "mjpl_b200.cu"]
HEADERS = sorted((_PKG / "csrc").glob("*.cuh")) + sorted((_PKG / "csrc").glob("*.h")) + [_PKG.parent / "include" / "mjpl_b200.h"]
OUT = _PKG / "lib" / "libmjpl_b200.so"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-shared", "-Xcompiler", "-fPIC",
    "-Xptxas", "-v",
]


def nvcc_path() -> str:
    for c in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and Path(c).exists():
            return c
    raise RuntimeError("nvcc not found")


def build(force: bool = False, verbose: bool = False) -> Path:
    newest = max(p.stat().st_mtime for p in SOURCES + HEADERS)
    if not force and OUT.exists() and OUT.stat().st_mtime >= newest:
        return OUT
    OUT.parent.mkdir(exist_ok=True)
    cmd = [nvcc_path(), *NVCC_FLAGS, "-o", str(OUT), *map(str, SOURCES)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or r.returncode != 0:
        print(r.stdout + r.stderr)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed")
    # registers / spills / stack per kernel for the record; compile times left out so that the file only changes
    # when the code does
    log = "\n".join(ln for ln in (r.stdout + r.stderr).splitlines() if "Compile time" not in ln)
    (OUT.parent / "ptxas.log").write_text(log + "\n")
    return OUT
