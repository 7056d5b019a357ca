#!/usr/bin/env python
"""Generate tests/golden/export.json by running the REAL reference's `export --format edge-list`
(cli.py:267-281, imported from /root/reference) in the build container: for every golden input, plain
and --bidirected, the bytes it wrote (also when it raised half-way), the exception and the warnings.

    python tools/gen_golden_export.py
"""
from __future__ import annotations

import base64
import hashlib
import json
import sys
import tempfile
import warnings
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, "/root/reference")
sys.path.insert(0, str(ROOT / "tests"))

from gfa2network.cli import main as ref_main  # noqa: E402  (the reference)

import golden_inputs as gi  # noqa: E402


def run_one(path: Path, bidirected: bool, full: bool) -> dict:
    out_path = path.parent / "edges.tsv"
    if out_path.exists():
        out_path.unlink()
    res: dict = {}
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        try:
            ref_main(["export", str(path), "--format", "edge-list", "--output", str(out_path)] + (["--bidirected"] if bidirected else []))
        except Exception as exc:  # noqa: BLE001
            res["raises"] = {"type": type(exc).__name__, "msg": str(exc)}
    res["warnings"] = [str(x.message) for x in w if issubclass(x.category, RuntimeWarning)]
    data = out_path.read_bytes() if out_path.exists() else b""
    res["nbytes"] = len(data)
    res["sha"] = hashlib.sha256(data).hexdigest()[:16]
    if full:
        res["out_b64"] = base64.b64encode(data).decode()
    return res


def main() -> None:
    gold = ROOT / "tests" / "golden"
    tmp = Path(tempfile.mkdtemp())
    out = {"cases": [], "fuzz": [], "drb1": []}
    for name, text in gi.LITERAL_CASES:
        p = tmp / "case.gfa"
        p.write_bytes(text)
        out["cases"].append({"name": name, "runs": [{"bidirected": b, "expect": run_one(p, b, True)} for b in (False, True)]})
    for seed in gi.FUZZ_SEEDS:
        p = tmp / "fuzz.gfa"
        p.write_bytes(gi.fuzz_text(seed))
        out["fuzz"].append({"seed": seed, "runs": [{"bidirected": b, "expect": run_one(p, b, False)} for b in (False, True)]})
    drb1 = gold / "DRB1-3123_unsorted.gfa"
    out["drb1"] = [{"bidirected": b, "expect": run_one(drb1, b, False)} for b in (False, True)]
    (gold / "export.json").write_text(json.dumps(out, indent=0, sort_keys=True))
    n_raise = sum(1 for c in out["cases"] for r in c["runs"] if "raises" in r["expect"])
    print("export.json:", len(out["cases"]), "literal inputs,", n_raise, "raising runs,", len(out["fuzz"]), "fuzz inputs, drb1")


if __name__ == "__main__":
    main()
