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


def _tuning_defines() -> list[str]:
    # kernel tuning knobs (defaults live in the .cu files); e.g. FW25_WS_TY=16 FW25_WS_MINB=2 python -m ...build --force
    return [f"-D{k}={v}" for k, v in sorted(os.environ.items()) if k.startswith("FW25_WS_") and k != "FW25_WS_WAVES"]


def _compile(nvcc: str, src: Path, obj: Path) -> tuple[int, str]:
    cmd = [nvcc, *NVCC_FLAGS, *_tuning_defines(), *_host_cxx(), "-c", "-o", str(obj), str(src)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    return r.returncode, " ".join(cmd) + "\n" + r.stdout + r.stderr


def build(force: bool = False, verbose: bool = False) -> Path:
    """Compile every .cu to an object (in parallel), link libfw25.so and the fw25_engine executable."""
    from concurrent.futures import ThreadPoolExecutor
    hdrs = list(CSRC.glob("*.h")) + list(CSRC.glob("*.cuh")) + [ROOT / "include" / "fw25.h", Path(__file__)]
    nvcc = _nvcc()
    objdir = PKG / "build"
    objdir.mkdir(exist_ok=True)
    srcs = sorted(CSRC.glob("*.cu"))
    todo = [(s, objdir / (s.stem + ".o")) for s in srcs]
    stale = [(s, o) for s, o in todo if force or _stale(o, [s, *hdrs])]
    log = []
    if stale:
        with ThreadPoolExecutor(max_workers=min(8, len(stale))) as ex:
            for (s, o), (rc, out) in zip(stale, ex.map(lambda so: _compile(nvcc, *so), stale)):
                log.append(out)
                if rc != 0:
                    sys.stderr.write(out)
                    raise RuntimeError(f"nvcc failed on {s.name}")
        (PKG / "build.log").write_text("\n".join(log))
        if verbose:
            print("\n".join(log))
    lib_objs = [str(o) for s, o in todo if s.name != "fw25_cli.cu"]
    all_objs = [str(o) for _, o in todo]
    if stale or not LIB.exists():
        r = subprocess.run([nvcc, *_host_cxx(), "-shared", "-o", str(LIB), *lib_objs], capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("linking libfw25.so failed")
    if stale or not CLI.exists():
        r = subprocess.run([nvcc, *_host_cxx(), "-o", str(CLI), *all_objs], capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("linking fw25_engine failed")
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True)
    print(LIB)
