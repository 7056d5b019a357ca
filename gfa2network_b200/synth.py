"""Deterministic synthetic GFA text for the benchmark configurations (BASELINE.json configs,
SURVEY.md 8(d)).  Thin ctypes wrapper over csrc/synth.c (host C, built with gcc); it only makes
input bytes -- both the GPU path and the CPU oracle read the same buffer."""
from __future__ import annotations

import ctypes as C
import shutil
import subprocess
from pathlib import Path

import numpy as np

PKG = Path(__file__).resolve().parent
SRC = PKG / "csrc" / "synth.c"
LIB = PKG / "libg2nsynth.so"


class _SP(C.Structure):
    _fields_ = [("n_seg", C.c_uint64), ("n_link", C.c_uint64), ("seed", C.c_uint64), ("id_base", C.c_uint64),
                ("kind", C.c_int32), ("seq_mean", C.c_int32), ("n_paths", C.c_int32), ("n_walks", C.c_int32),
                ("interleave", C.c_int32), ("header", C.c_int32), ("uni_base", C.c_uint64), ("uni_n", C.c_uint64)]


def build(force: bool = False) -> Path:
    stale = lambda: not LIB.exists() or LIB.stat().st_mtime < SRC.stat().st_mtime  # noqa: E731
    if force or stale():
        import fcntl
        import os

        with open(PKG / ".build.lock", "w") as lock:  # several ranks of one job may get here together
            fcntl.flock(lock, fcntl.LOCK_EX)
            if force or stale():
                cc = shutil.which("gcc") or "cc"
                tmp = LIB.with_name(f".libg2nsynth.{os.getpid()}.so")
                subprocess.run([cc, "-O2", "-fPIC", "-shared", "-std=c11", "-o", str(tmp), str(SRC)], check=True)
                os.replace(tmp, LIB)
    return LIB


_lib = None


def _load():
    global _lib
    if _lib is None:
        _lib = C.CDLL(str(build()))
        _lib.g2n_synth_bound.argtypes = [C.POINTER(_SP)]
        _lib.g2n_synth_bound.restype = C.c_uint64
        _lib.g2n_synth.argtypes = [C.c_void_p, C.c_uint64, C.POINTER(_SP)]
        _lib.g2n_synth.restype = C.c_uint64
    return _lib


def synth_gfa(n_seg: int, n_link: int, *, seed: int = 2, kind: int = 1, seq_mean: int = 0, n_paths: int = 0,
              n_walks: int = 0, interleave: int = 0, header: bool = True, id_base: int = 0, uniform_range: tuple[int, int] = (0, 0),
              out: np.ndarray | None = None) -> np.ndarray:
    """Returns a uint8 array holding the GFA text (a view into *out* when given)."""
    lib = _load()
    sp = _SP(n_seg, n_link, seed, id_base, kind, seq_mean, n_paths, n_walks, interleave, int(header), uniform_range[0], uniform_range[1])
    bound = int(lib.g2n_synth_bound(C.byref(sp)))
    buf = out if out is not None else np.empty(bound, dtype=np.uint8)
    if buf.size < bound:
        raise ValueError(f"buffer too small: {buf.size} < {bound}")
    n = int(lib.g2n_synth(buf.ctypes.data, buf.size, C.byref(sp)))
    return buf[:n]


# the named configurations of BASELINE.json (index = config number)
CONFIGS = {
    "C2": dict(n_seg=1_000_000, n_link=3_000_000, seed=2, kind=1, mode=dict(directed=False), fmt="csr"),
    "C3": dict(n_seg=10_000_000, n_link=30_000_000, seed=3, kind=2, mode=dict(bidirected=True, weight_tag="RC"), fmt="csr"),
    # BASELINE "directed COO": by the reference's quirk Q6 (builders.py:282-283, utils.py:47-48) `--matrix-format coo` in default
    # directed mode saves max(S, S^T) as CSR (C4d); only with --asymmetric is the result the raw COO (C4)
    "C4": dict(n_seg=20_000_000, n_link=60_000_000, seed=4, kind=1, n_paths=25, n_walks=25, mode=dict(asymmetric=True), fmt="coo"),
    "C4d": dict(n_seg=20_000_000, n_link=60_000_000, seed=4, kind=1, n_paths=25, n_walks=25, mode=dict(), fmt="coo"),
    "C5": dict(n_seg=100_000_000, n_link=400_000_000, seed=5, kind=1, seq_mean=270, mode=dict(), fmt="csr"),
}
