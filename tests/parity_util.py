"""Shared comparison helpers: golden-vector decoding and bit-exact matrix checks."""
from __future__ import annotations

import base64
import hashlib
import json
import warnings
from pathlib import Path

import numpy as np

GOLD = Path(__file__).resolve().parent / "golden"

# integer-dtype overflow inside NumPy's list->array cast: declared outside parity scope
# (SURVEY 8d "NaN weights, int8 overflow and ints beyond 2^63 are declared outside parity scope")
OUT_OF_SCOPE_MSG = ("out of bounds for int", "cannot convert float infinity to integer", "cannot convert float NaN")


def dec(d) -> np.ndarray:
    return np.frombuffer(base64.b64decode(d["b64"]), dtype=np.dtype(d["dtype"]))


def sha(*arrays) -> str:
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()[:16]


def load_json(name):
    return json.loads((GOLD / name).read_text())


def out_of_scope(expect) -> bool:
    r = expect.get("raises")
    return bool(r) and any(m in r["msg"] for m in OUT_OF_SCOPE_MSG)


def arrays_of(A):
    if A.format == "coo":
        return A.row, A.col, A.data
    return A.indptr, A.indices, A.data


def check_full(A, exp, what=""):
    """Bit-exact comparison against a fully stored golden matrix description."""
    assert A.format == exp["format"], (what, A.format, exp["format"])
    assert A.dtype.str == exp["dtype"], (what, A.dtype, exp["dtype"])
    assert list(A.shape) == exp["shape"], what
    keys = ("row", "col", "data") if exp["format"] == "coo" else ("indptr", "indices", "data")
    for k, got in zip(keys, arrays_of(A)):
        want = dec(exp[k])
        assert got.dtype == want.dtype, (what, k, got.dtype, want.dtype)
        assert got.shape == want.shape, (what, k, got.shape, want.shape)
        # bit-exact (NaN-safe, signed-zero-safe): compare raw bytes
        assert got.tobytes() == want.tobytes(), (what, k, got[:20], want[:20])


def check_sha(A, exp, what=""):
    assert A.format == exp["format"], (what, A.format)
    assert A.dtype.str == exp["dtype"], what
    assert list(A.shape) == exp["shape"], what
    assert int(A.nnz) == exp["nnz"], (what, A.nnz, exp["nnz"])
    arrs = arrays_of(A)
    assert arrs[1].dtype.str == exp["idx_dtype"], what
    assert sha(*arrs) == exp["sha"], what


def run_and_compare(parse, convert, text_or_path, mode, expect, full=True, what=""):
    """parse(text_or_path, return_node_list=True, raw_bytes_id=True, **mode) -> (A, nodes)."""
    if out_of_scope(expect):
        return "skipped"
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        try:
            A, nodes = parse(text_or_path, return_node_list=True, raw_bytes_id=True, **mode)
            raised = None
        except Exception as exc:  # noqa: BLE001
            raised = exc
    got_w = [str(x.message) for x in w if issubclass(x.category, RuntimeWarning)]
    assert got_w == expect["warnings"], (what, got_w, expect["warnings"])
    if "raises" in expect:
        assert raised is not None, (what, "expected", expect["raises"])
        assert type(raised).__name__ == expect["raises"]["type"], (what, repr(raised))
        assert str(raised) == expect["raises"]["msg"], (what, str(raised))
        return "raised"
    assert raised is None, (what, repr(raised))
    if full:
        check_full(A, expect["raw"], what + " raw")
        check_full(convert(A, "csr"), expect["csr"], what + " csr")
        check_full(convert(A, "csc"), expect["csc"], what + " csc")
        want_nodes = [base64.b64decode(x) for x in expect["nodes_b64"]]
        assert list(nodes) == want_nodes, (what, nodes[:5], want_nodes[:5])
    else:
        check_sha(A, expect["raw"], what + " raw")
        csr = convert(A, "csr")
        assert int(csr.nnz) == expect["csr"]["nnz"], what
        assert sha(csr.indptr, csr.indices, csr.data) == expect["csr"]["sha"], what + " csr"
        csc = convert(A, "csc")
        assert sha(csc.indptr, csc.indices, csc.data) == expect["csc"]["sha"], what + " csc"
        assert hashlib.sha256(b"\n".join(nodes)).hexdigest()[:16] == expect["nodes_sha"], what + " nodes"
    return "ok"
