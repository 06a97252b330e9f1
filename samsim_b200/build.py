"""In-tree nvcc build of the CUDA library (sm_100a only).  Used by __graft_entry__.build()."""
from __future__ import annotations

import os
import subprocess
from pathlib import Path

ROOT = Path(__file__).resolve().parent
SRC = ROOT / "csrc" / "samsim_b200.cu"
HOST_SRC = ROOT / "csrc" / "host" / "grotz_host.cpp"
LIBDIR = ROOT / "_lib"
LIB = LIBDIR / "libsamsim_b200.so"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-fmad=false",            # arithmetic contract: no contraction, see csrc/physics.cuh
    "-Xcompiler", "-fPIC", "-Xcompiler", "-ffp-contract=off",
    "-shared",
]


def _deps():
    d = list((ROOT / "csrc").rglob("*.cu")) + list((ROOT / "csrc").rglob("*.cuh")) + list((ROOT / "csrc").rglob("*.h")) + \
        list((ROOT / "csrc").rglob("*.cpp")) + [ROOT.parent / "include" / "samsim_b200.h"]
    return [p for p in d if p.exists()]


def needs_build() -> bool:
    if not LIB.exists():
        return True
    t = LIB.stat().st_mtime
    return any(p.stat().st_mtime > t for p in _deps())


def build(force: bool = False, verbose: bool = False) -> Path:
    if not force and not needs_build():
        return LIB
    LIBDIR.mkdir(exist_ok=True)
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    srcs = [str(SRC)] + ([str(HOST_SRC)] if HOST_SRC.exists() else [])
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", str(LIB)] + srcs
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + r.stdout + r.stderr)
    if verbose:
        print(r.stderr)
    return LIB


if __name__ == "__main__":
    import sys
    print(build(force="--force" in sys.argv, verbose=True))
