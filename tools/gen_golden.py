#!/usr/bin/env python
"""Generate tests/golden/*.json|npz by running the REAL reference (imported from
/root/reference) in the build container.  The reference cannot travel to the GPU box, so
its outputs are committed as fixtures together with this script.

    python tools/gen_golden.py            # rewrites tests/golden/cases.json, fuzz.npz, drb1.json

For every (input, mode) it records what ``parse_gfa(..., build_matrix=True,
return_node_list=True)`` returned (format class, dtype, index dtype, shape, arrays, node
list, warnings) or the exception it raised, plus ``convert_format(A, "csr"|"csc")``.
"""
from __future__ import annotations

import base64
import hashlib
import json
import sys
import tempfile
import warnings
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, "/root/reference")
sys.path.insert(0, str(ROOT / "tests"))

from gfa2network import parse_gfa, convert_format  # noqa: E402  (the reference)

import golden_inputs as gi  # noqa: E402


def enc(a: np.ndarray) -> dict:
    a = np.ascontiguousarray(a)
    return {"dtype": a.dtype.str, "b64": base64.b64encode(a.tobytes()).decode()}


def describe(A) -> dict:
    d = {"format": A.format, "dtype": A.dtype.str, "shape": list(A.shape), "nnz": int(A.nnz)}
    if A.format == "coo":
        d["row"] = enc(A.row)
        d["col"] = enc(A.col)
        d["data"] = enc(A.data)
    else:
        d["indptr"] = enc(A.indptr)
        d["indices"] = enc(A.indices)
        d["data"] = enc(A.data)
    return d


def sha(*arrays) -> str:
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()[:16]


def run_one(path: Path, mode: dict, full: bool = True) -> dict:
    out: dict = {}
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        try:
            A, nodes = parse_gfa(path, build_graph=False, build_matrix=True,
                                 return_node_list=True, raw_bytes_id=True, **mode)
        except Exception as exc:  # noqa: BLE001
            out["raises"] = {"type": type(exc).__name__, "msg": str(exc)}
            A = None
    out["warnings"] = [str(x.message) for x in w if issubclass(x.category, RuntimeWarning)]
    if A is None:
        return out
    csr = convert_format(A, "csr")
    csc = convert_format(A, "csc")
    if full:
        out["raw"] = describe(A)
        out["csr"] = describe(csr)
        out["csc"] = describe(csc)
        out["nodes_b64"] = [base64.b64encode(x).decode() for x in nodes]
    else:
        out["raw"] = {"format": A.format, "dtype": A.dtype.str, "shape": list(A.shape), "nnz": int(A.nnz)}
        if A.format == "coo":
            out["raw"]["sha"] = sha(A.row, A.col, A.data)
            out["raw"]["idx_dtype"] = A.row.dtype.str
        else:
            out["raw"]["sha"] = sha(A.indptr, A.indices, A.data)
            out["raw"]["idx_dtype"] = A.indices.dtype.str
        out["csr"] = {"nnz": int(csr.nnz), "sha": sha(csr.indptr, csr.indices, csr.data),
                      "sum": float(csr.data.astype(np.float64).sum())}
        out["csc"] = {"nnz": int(csc.nnz), "sha": sha(csc.indptr, csc.indices, csc.data)}
        out["nodes_sha"] = hashlib.sha256(b"\n".join(nodes)).hexdigest()[:16]
    return out


def main() -> None:
    gold = ROOT / "tests" / "golden"
    gold.mkdir(parents=True, exist_ok=True)
    tmp = Path(tempfile.mkdtemp())

    # ---- literal cases: full arrays ------------------------------------------------
    cases = []
    for name, text in gi.LITERAL_CASES:
        p = tmp / "case.gfa"
        p.write_bytes(text)
        entry = {"name": name, "text_b64": base64.b64encode(text).decode(), "runs": []}
        for mode in gi.MODES:
            entry["runs"].append({"mode": mode, "expect": run_one(p, mode)})
        cases.append(entry)
    (gold / "cases.json").write_text(json.dumps(cases, indent=0, sort_keys=True))
    print("cases.json:", len(cases), "inputs x", len(gi.MODES), "modes")

    # ---- fuzz cases: bigger inputs, hashes only (inputs regenerated from the seed) -------
    fuzz = []
    for seed in gi.FUZZ_SEEDS:
        text = gi.fuzz_text(seed)
        p = tmp / "fuzz.gfa"
        p.write_bytes(text)
        entry = {"seed": seed, "text_sha": hashlib.sha256(text).hexdigest()[:16], "nbytes": len(text), "runs": []}
        for mode in gi.FUZZ_MODES:
            entry["runs"].append({"mode": mode, "expect": run_one(p, mode, full=False)})
        fuzz.append(entry)
    (gold / "fuzz.json").write_text(json.dumps(fuzz, indent=0, sort_keys=True))
    print("fuzz.json:", len(fuzz))

    # ---- DRB1 fixture (config C1) ------------------------------------------------------
    drb1 = gold / "DRB1-3123_unsorted.gfa"
    if not drb1.exists():
        drb1.write_bytes(Path("/root/reference/tests/data/DRB1-3123_unsorted.gfa").read_bytes())
    entry = {"runs": []}
    for mode in gi.MODES:
        entry["runs"].append({"mode": mode, "expect": run_one(drb1, mode, full=False)})
    (gold / "drb1.json").write_text(json.dumps(entry, indent=0, sort_keys=True))
    print("drb1.json done")


if __name__ == "__main__":
    main()
