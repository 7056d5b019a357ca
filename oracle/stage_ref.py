"""oracle/stage_ref.py -- TEST / MEASUREMENT INFRASTRUCTURE ONLY.  NOT PART OF THE PRODUCT PATH.

Stages the UNMODIFIED reference package next to the oracle so that it travels to the GPU box:

    /root/reference/gfa2network/*.py  ->  oracle/_ref/gfa2network/      (git-ignored, NOT gpurun-ignored)

The reference is pure Python; nothing is compiled and nothing is edited -- the copy is
byte-identical (`python oracle/stage_ref.py --check` compares sha256 per file).  `bench.py --impl reference`
times it (`cpu_baseline_reference`, kind "reference") beside the C port of its algorithm (`cpu_baseline`, kind
"port"); `__graft_entry__.build()` calls stage() when /root/reference is present (the build container) and the
GPU box only uses the staged files.  No reference source is ever committed to this repository.
"""
from __future__ import annotations

import hashlib
import shutil
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
SRC = Path("/root/reference/gfa2network")
DST = HERE / "_ref" / "gfa2network"


def _sha(p: Path) -> str:
    return hashlib.sha256(p.read_bytes()).hexdigest()


def stage(verbose: bool = False) -> bool:
    """Copy the reference package when its source tree is present.  Returns True if oracle/_ref is usable."""
    if SRC.is_dir():
        DST.parent.mkdir(parents=True, exist_ok=True)
        if DST.exists():
            shutil.rmtree(DST)
        shutil.copytree(SRC, DST, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
        (DST.parent / "MANIFEST.sha256").write_text("".join(f"{_sha(p)}  {p.relative_to(DST.parent)}\n" for p in sorted(DST.rglob("*.py"))))
        if verbose:
            print(f"staged {len(list(DST.rglob('*.py')))} files -> {DST}")
    return (DST / "builders.py").exists()


def check() -> bool:
    man = DST.parent / "MANIFEST.sha256"
    if not man.exists():
        return False
    for line in man.read_text().splitlines():
        h, rel = line.split("  ", 1)
        if _sha(DST.parent / rel) != h:
            return False
        if SRC.is_dir() and _sha(SRC.parent / rel) != h:
            return False
    return True


def import_reference():
    """The staged reference as a module object (its own public API: parse_gfa, convert_format), or None."""
    if not (DST / "builders.py").exists():
        return None
    import importlib

    root = str(DST.parent)
    if root not in sys.path:
        sys.path.insert(0, root)
    return importlib.import_module("gfa2network")


if __name__ == "__main__":
    if "--check" in sys.argv:
        ok = check()
        print("oracle/_ref matches the manifest" + (" and /root/reference" if SRC.is_dir() else "") if ok else "oracle/_ref is missing or differs")
        sys.exit(0 if ok else 1)
    print("ok" if stage(verbose=True) else "no reference source here and nothing staged")
