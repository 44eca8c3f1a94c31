"""In-tree build of the CUDA engine: libfw25.so (C-ABI, include/fw25.h) and the fw25_engine executable
(the .dat-directory drop-in for the reference's pre-compiled binaries).  sm_100a only.

    python -m fullwave25_b200.build [--force]
"""

from __future__ import annotations

import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
ROOT = PKG.parent
LIB = PKG / "libfw25.so"
CLI = PKG / "fw25_engine"

NVCC_FLAGS = [
    "-O3", "-std=c++17",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo",
    "-fmad=false",            # no implicit FMA contraction: every FFMA is an explicit __fmaf_rn
    "-Xcompiler", "-fPIC",
    "-Xptxas", "-v",
    "-I", str(ROOT / "include"),
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found: the fw25 engine needs the CUDA 12.9 toolkit to build")


def _host_cxx() -> list[str]:
    # the image's default CC/CXX wrappers lack parts of the toolchain; pin the distro compiler
    for cand in ("/usr/bin/g++-13", "/usr/bin/g++"):
        if Path(cand).exists():
            return ["-ccbin", cand]
    return []


def sources() -> list[Path]:
    return sorted(p for p in CSRC.glob("*.cu") if p.name != "fw25_cli.cu")


def _stale(target: Path, deps: list[Path]) -> bool:
    if not target.exists():
        return True
    t = target.stat().st_mtime
    return any(d.stat().st_mtime > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> Path:
    deps = list(CSRC.glob("*")) + [ROOT / "include" / "fw25.h", Path(__file__)]
    nvcc = _nvcc()
    if force or _stale(LIB, deps):
        cmd = [nvcc, *NVCC_FLAGS, *_host_cxx(), "-shared", "-o", str(LIB), *map(str, sources())]
        r = subprocess.run(cmd, capture_output=True, text=True)
        (PKG / "build.log").write_text(" ".join(cmd) + "\n" + r.stdout + r.stderr)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("nvcc failed building libfw25.so")
        if verbose:
            print(r.stderr)
    cli_src = CSRC / "fw25_cli.cu"
    if cli_src.exists() and (force or _stale(CLI, deps)):
        cmd = [nvcc, *NVCC_FLAGS, *_host_cxx(), "-o", str(CLI), str(cli_src), *map(str, sources())]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("nvcc failed building fw25_engine")
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True)
    print(LIB)
