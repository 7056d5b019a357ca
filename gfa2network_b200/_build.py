"""Build libg2n.so (hand-written sm_100a CUDA kernels + C ABI) in-tree with nvcc."""
from __future__ import annotations

import os
import shutil
import subprocess
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
LIB = PKG / "libg2n.so"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17", "-shared", "-Xcompiler", "-fPIC,-pthread",
]


def sources() -> list[Path]:
    return sorted(CSRC.glob("*.cu")) + sorted(CSRC.glob("*.cuh")) + sorted(CSRC.glob("*.h")) + [PKG.parent / "include" / "g2n.h"]


def needs_build() -> bool:
    if not LIB.exists():
        return True
    t = LIB.stat().st_mtime
    return any(s.stat().st_mtime > t for s in sources())


def build(force: bool = False, verbose: bool = False) -> Path:
    """Several ranks of one job may get here at the same time (torchrun): one of them builds under a file
    lock into a temporary name and renames it into place, the others wait and find it up to date."""
    if not force and not needs_build():
        return LIB
    import fcntl

    with open(PKG / ".build.lock", "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        if not force and not needs_build():
            return LIB
        nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
        if not Path(nvcc).exists():
            raise RuntimeError("nvcc not found: libg2n.so cannot be built (there is no CPU fallback)")
        extra = os.environ.get("G2N_NVCC_EXTRA", "").split()
        tmp = LIB.with_name(f".libg2n.{os.getpid()}.so")
        cmd = [nvcc, *NVCC_FLAGS, *extra, "-o", str(tmp), str(CSRC / "g2n.cu"), "-lz"]  # zlib: *.gz input (g2n_build_gz)
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            tmp.unlink(missing_ok=True)
            raise RuntimeError("nvcc failed:\n" + r.stdout + r.stderr)
        os.replace(tmp, LIB)
        if verbose:
            print(r.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force=True, verbose=bool(os.environ.get("G2N_VERBOSE"))))
