#!/usr/bin/env python
"""Generate tests/golden/distance.json by running the REAL reference's load_paths / genome_distance_matrix
(analysis.py, imported from /root/reference) in the build container.

    python tools/gen_golden_distance.py
"""
from __future__ import annotations

import base64
import json
import sys
import tempfile
import warnings
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, "/root/reference")
sys.path.insert(0, str(ROOT / "tests"))

from gfa2network.analysis import genome_distance_matrix, load_paths  # noqa: E402  (the reference)

import distance_inputs as di  # noqa: E402


def run(path: Path, method: str) -> dict:
    res: dict = {}
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        try:
            paths = load_paths(str(path))
            res["paths"] = {k: v for k, v in paths.items()}
            M = genome_distance_matrix(str(path), method=method)
            arr = M.values if hasattr(M, "values") else np.asarray(M)
            res["labels"] = list(M.index) if hasattr(M, "index") else None
            res["matrix"] = [[("inf" if np.isinf(x) else float(x)) for x in row] for row in arr.tolist()]
        except Exception as exc:  # noqa: BLE001
            res["raises"] = {"type": type(exc).__name__, "msg": str(exc)}
    res["warnings"] = sorted({str(x.message) for x in w if issubclass(x.category, RuntimeWarning)})
    return res


def main() -> None:
    tmp = Path(tempfile.mkdtemp())
    out = []
    for name, text in di.CASES:
        p = tmp / "case.gfa"
        p.write_bytes(text)
        out.append({"name": name, "text_b64": base64.b64encode(text).decode() if len(text) < 100_000 else None,
                    "runs": [{"method": m, "expect": run(p, m)} for m in ("min", "mean")]})
    (ROOT / "tests" / "golden" / "distance.json").write_text(json.dumps(out, indent=0, sort_keys=True))
    print("distance.json:", len(out), "inputs;", sum(1 for c in out for r in c["runs"] if "raises" in r["expect"]), "raising runs")


if __name__ == "__main__":
    main()
