"""In-tree build of the C-ABI CUDA library (sm_100a only). No JIT cache, no torch extension:
`nvcc -shared` straight into swarm_robotics_pheromones_b200/lib/libpherofield.so so the built file
travels with the repo snapshot to the GPU box."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
LIB_DIR = PKG / "lib"
LIB = LIB_DIR / "libpherofield.so"
HEADER = PKG.parent / "include" / "pherofield.h"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-fmad=false",            # belt and braces: the kernels use explicit __f*_rn intrinsics
    "-prec-div=true", "-prec-sqrt=true", "-ftz=false",
    "-Xcompiler", "-fPIC,-Wno-deprecated-declarations", "-shared", "-diag-suppress", "1444",
    "--expt-relaxed-constexpr",
    "-Xptxas", "-v",
]


def nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: the pherofield library cannot be built")
    return exe


def sources():
    return [CSRC / "pherofield.cu", CSRC / "pf_trail.cu", CSRC / "pf_heatmap.cu"]


def deps():
    return list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + [HEADER]


def is_stale() -> bool:
    if not LIB.exists():
        return True
    t = LIB.stat().st_mtime
    return any(p.stat().st_mtime > t for p in deps())


def build(force: bool = False, verbose: bool = False) -> Path:
    if not force and not is_stale():
        return LIB
    LIB_DIR.mkdir(exist_ok=True)
    # one builder at a time (every torchrun rank may get here at once on a fresh clone), and the library
    # appears atomically: nvcc writes a temporary file that is renamed into place
    import fcntl
    with open(LIB_DIR / ".build.lock", "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        if not force and not is_stale():      # another process built it while this one waited
            return LIB
        return _build_locked(verbose)


def _build_locked(verbose: bool) -> Path:
    tmp = LIB_DIR / f".libpherofield.{os.getpid()}.so"
    cmd = [nvcc(), *NVCC_FLAGS, "-o", str(tmp), *map(str, sources())]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    (LIB_DIR / "build.log").write_text(" ".join(cmd) + "\n" + proc.stdout + proc.stderr)
    if proc.returncode == 0:
        os.replace(tmp, LIB)
    elif tmp.exists():
        tmp.unlink()
    if proc.returncode != 0:
        errs = [ln for ln in (proc.stdout + proc.stderr).splitlines() if "error" in ln.lower()]
        sys.stderr.write("\n".join(errs[:40]) + f"\n(full log: {LIB_DIR / 'build.log'})\n")
        raise RuntimeError("nvcc failed building libpherofield.so")
    if verbose:
        print(proc.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
