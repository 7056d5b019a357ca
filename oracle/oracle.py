"""oracle/oracle.py -- TEST INFRASTRUCTURE ONLY.  NOT PART OF THE PRODUCT PATH.

ctypes wrapper over oracle/_build/liboracle.so (the C restatement in gfa_oracle.c) that
finishes the reference's ``parse_gfa(build_matrix=True)`` exactly where the reference hands
over to SciPy:

* ``sp.coo_matrix((data, (rows, cols)), shape=(n, n), dtype=dt)``   builders.py:279-281
* ``out_mat.maximum(out_mat.T)`` when ``not asymmetric and graph_directed``   builders.py:282-283
* node list in ID order, decoded unless ``raw_bytes_id``   builders.py:284-288
* ``convert_format``   utils.py:40-63

Only tests/, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs may import this module.  Parity status: pinned (see gfa_oracle.c header).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import warnings
from pathlib import Path

import numpy as np
import scipy.sparse as sp

_HERE = Path(__file__).resolve().parent
_LIB_PATH = _HERE / "_build" / "liboracle.so"

ERR_MESSAGES = {
    1: (ValueError, "Malformed L record"),
    2: (ValueError, "Malformed E record"),
    3: (ValueError, "Malformed C record"),
    4: (ValueError, "Malformed P record"),
    5: (ValueError, "Malformed O record"),
    6: (IndexError, "list index out of range"),
    7: (IndexError, "index out of range"),
    9: (OverflowError, "int too large to convert to float"),
    10: (NotImplementedError, "non-ASCII numeric weight tag value"),
}


class _Params(C.Structure):
    _fields_ = [
        ("directed", C.c_int32),
        ("bidirected", C.c_int32),
        ("keep_directed_bidir", C.c_int32),
        ("strip_orientation", C.c_int32),
        ("weight_tag", C.c_char_p),
        ("weight_tag_len", C.c_int32),
    ]


class _Result(C.Structure):
    _fields_ = [
        ("n_nodes", C.c_int64),
        ("n_triplets", C.c_int64),
        ("rows", C.POINTER(C.c_int32)),
        ("cols", C.POINTER(C.c_int32)),
        ("data", C.POINTER(C.c_double)),
        ("names", C.POINTER(C.c_uint8)),
        ("name_off", C.POINTER(C.c_int64)),
        ("n_records", C.c_int64),
        ("n_edge_records", C.c_int64),
        ("err_kind", C.c_int32),
        ("err_offset", C.c_int64),
        ("err_aux_off", C.c_int64),
        ("err_aux_len", C.c_int64),
        ("unknown_byte", C.c_int32),
        ("unknown_offset", C.c_int64),
    ]


def build_oracle() -> Path:
    """Compile the C restatement with gcc (idempotent)."""
    src = _HERE / "gfa_oracle.c"
    if not _LIB_PATH.exists() or _LIB_PATH.stat().st_mtime < src.stat().st_mtime:
        subprocess.run(["make", "-C", str(_HERE), "-s"], check=True, capture_output=True)
    return _LIB_PATH


_lib = None


def _load():
    global _lib
    if _lib is None:
        build_oracle()
        _lib = C.CDLL(str(_LIB_PATH))
        _lib.ora_parse.argtypes = [C.c_void_p, C.c_int64, C.POINTER(_Params), C.POINTER(_Result)]
        _lib.ora_parse.restype = C.c_int
        _lib.ora_free.argtypes = [C.POINTER(_Result)]
        _lib.ora_free.restype = None
    return _lib


def _as_u8(text) -> np.ndarray:
    if isinstance(text, np.ndarray):
        return np.ascontiguousarray(text, dtype=np.uint8)
    if isinstance(text, (bytes, bytearray, memoryview)):
        return np.frombuffer(text, dtype=np.uint8)
    p = str(text)
    if p.endswith(".gz"):
        import gzip

        with gzip.open(p, "rb") as fh:
            return np.frombuffer(fh.read(), dtype=np.uint8)
    return np.fromfile(p, dtype=np.uint8)


def oracle_triplets(text, *, directed=True, bidirected=False, keep_directed_bidir=False,
                    strip_orientation=False, weight_tag=None):
    """Run the C restatement; return dict(rows, cols, data, names(list[bytes]), diag...)."""
    lib = _load()
    buf = _as_u8(text)
    wt = weight_tag.encode() if weight_tag else None
    p = _Params(int(bool(directed)), int(bool(bidirected)), int(bool(keep_directed_bidir)),
                int(bool(strip_orientation)), wt, len(wt) if wt else 0)
    r = _Result()
    lib.ora_parse(buf.ctypes.data if buf.size else None, buf.size, C.byref(p), C.byref(r))
    try:
        nt, nn = r.n_triplets, r.n_nodes
        rows = np.ctypeslib.as_array(r.rows, shape=(nt,)).copy() if nt else np.zeros(0, np.int32)
        cols = np.ctypeslib.as_array(r.cols, shape=(nt,)).copy() if nt else np.zeros(0, np.int32)
        data = np.ctypeslib.as_array(r.data, shape=(nt,)).copy() if nt else np.zeros(0, np.float64)
        off = np.ctypeslib.as_array(r.name_off, shape=(nn + 1,)).copy()
        nb = int(off[-1])
        names = np.ctypeslib.as_array(r.names, shape=(nb,)).copy() if nb else np.zeros(0, np.uint8)
        out = dict(rows=rows, cols=cols, data=data, name_bytes=names, name_off=off, n_nodes=nn,
                   n_records=r.n_records, n_edge_records=r.n_edge_records, err_kind=r.err_kind,
                   err_offset=r.err_offset, err_aux=(r.err_aux_off, r.err_aux_len),
                   unknown_byte=r.unknown_byte, unknown_offset=r.unknown_offset)
    finally:
        lib.ora_free(C.byref(r))
    return out


def _raise_for(res, buf):
    # parser.py:125-131 -- the one-shot warning precedes any later error
    if res["unknown_byte"] >= 0:
        warnings.warn(
            "Skipping unsupported record: " + bytes([res["unknown_byte"]]).decode(),
            RuntimeWarning, stacklevel=3)
    k = res["err_kind"]
    if k == 0:
        return
    if k == 8:
        o, l = res["err_aux"]
        bytes(buf[o:o + l]).decode()  # raises the same UnicodeDecodeError
    exc, msg = ERR_MESSAGES[k]
    raise exc(msg)


def oracle_parse_gfa(text, *, directed=True, weight_tag=None, strip_orientation=False,
                     bidirected=False, keep_directed_bidir=False, dtype="float64",
                     asymmetric=False, raw_bytes_id=False, return_node_list=False):
    """The matrix path of the reference's parse_gfa (builders.py:129-299), CPU oracle."""
    buf = _as_u8(text)
    res = oracle_triplets(buf, directed=directed, bidirected=bidirected,
                          keep_directed_bidir=keep_directed_bidir,
                          strip_orientation=strip_orientation, weight_tag=weight_tag)
    _raise_for(res, buf)
    graph_directed = keep_directed_bidir or (not bidirected and directed)  # builders.py:143
    n = res["n_nodes"]
    dt = np.dtype(dtype)
    with np.errstate(invalid="ignore"):
        data = res["data"].astype(dt)
    out = sp.coo_matrix((data, (res["rows"], res["cols"])), shape=(n, n), dtype=dt)  # builders.py:281
    if not asymmetric and graph_directed:
        out = out.maximum(out.T)  # builders.py:283
    if not return_node_list:
        return out
    off = res["name_off"]
    nb = res["name_bytes"].tobytes()
    nodes = [nb[off[i]:off[i + 1]] for i in range(n)]
    if not raw_bytes_id:
        nodes = [x.decode() for x in nodes]  # builders.py:287
    return out, nodes


def oracle_convert_format(A, fmt: str):
    """utils.py:40-63."""
    fmt = fmt.lower()
    if fmt not in {"csr", "csc", "coo", "dok"}:
        raise ValueError("matrix-format must be csr|csc|coo|dok")
    if fmt == "coo":
        return A
    return A.asformat(fmt)


def oracle_edge_list(text, *, bidirected=False):
    """``export --format edge-list`` (cli.py:264-281), CPU oracle: (bytes written, exception or None, warnings).

    The endpoint strings the reference writes (``from_segment[:orientation]``, ``to_segment[:orientation]``) are
    the node keys of the matrix builder in its one-triplet-per-record mode (builders.py:211-212, 222-226), so
    the restated tokenizer/builder is reused: row/col of the k-th triplet name the k-th line.  On a malformed
    record the reference has already written every earlier line (the loop of cli.py:269-279 raises inside
    ``GFAParser``): those bytes are the edge list of the text in front of the offending line."""
    buf = _as_u8(text)
    res = oracle_triplets(buf, directed=True, bidirected=bidirected, keep_directed_bidir=True)
    warns = []
    if res["unknown_byte"] >= 0:
        warns.append("Skipping unsupported record: " + bytes([res["unknown_byte"]]).decode())
    exc = None
    if res["err_kind"]:
        try:
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                _raise_for(res, buf)
        except Exception as e:  # noqa: BLE001
            exc = e
        res = oracle_triplets(buf[: res["err_offset"]], directed=True, bidirected=bidirected, keep_directed_bidir=True)
        assert res["err_kind"] == 0
    off = res["name_off"]
    nb = res["name_bytes"].tobytes()
    names = [nb[off[i]:off[i + 1]] for i in range(res["n_nodes"])]
    lines = []
    for r, c in zip(res["rows"].tolist(), res["cols"].tolist()):
        u, v = names[r], names[c]
        if exc is None:
            try:
                u.decode(), v.decode()  # cli.py:278
            except UnicodeDecodeError as e:
                exc = e
                break
        lines.append(u + b"\t" + v + b"\n")
    return b"".join(lines), exc, warns


def oracle_load_paths(text, *, raw_bytes=False):
    """analysis.py:164-177 over parser.py:114-134, 229-247, 343-361: P / O records, later names overwrite."""
    buf = _as_u8(text).tobytes()
    paths = {}
    for line in buf.split(b"\n"):
        if not line or line[0] not in b"PO":
            continue
        fields = line.split(b"\t")
        if fields[0] not in (b"P", b"O"):
            continue
        if len(fields) < 3:
            raise ValueError(f"Malformed {fields[0].decode()} record")
        segs = []
        for entry in fields[2].split(b","):
            seg = entry[:-1] if entry.endswith((b"+", b"-")) else entry
            segs.append(seg if raw_bytes else seg.decode("ascii"))
        paths[fields[1] if raw_bytes else fields[1].decode("ascii")] = segs
    return paths


def oracle_distance_matrix(text, method="min"):
    """analysis.py:180-272, CPU oracle: (labels, matrix).  The reference's graph is the DiGraph of
    parse_gfa(build_graph=True) without weights (builders.py:141, 246-256): nodes and out-edges are those of
    the asymmetric COO the matrix half emits, every edge counts 1 -- multi-source BFS with SciPy's csgraph."""
    from scipy.sparse.csgraph import dijkstra

    buf = _as_u8(text)
    A, nodes = oracle_parse_gfa(buf, asymmetric=True, return_node_list=True)  # raises / warns like the reference's parser
    paths = oracle_load_paths(buf)
    index = {n: i for i, n in enumerate(nodes)}
    S = A.tocsr()
    S.data[:] = 1.0
    names = list(paths)
    n = len(names)
    lengths = []
    for name in names:
        src = []
        for s in paths[name]:
            if s not in index:
                import networkx as nx

                raise nx.NodeNotFound(f"Node {s} not found in graph")
            src.append(index[s])
        if S.shape[0] and src:
            lengths.append(dijkstra(S, directed=True, indices=sorted(set(src)), unweighted=True, min_only=True))
        else:
            lengths.append(np.full(S.shape[0], np.inf))
    M = np.zeros((n, n))
    for i in range(n):
        for j in range(i, n):
            if i == j:
                d = 0.0
            elif method == "min":
                ds = [lengths[i][index[v]] for v in paths[names[j]] if v in index and np.isfinite(lengths[i][index[v]])]
                d = min(ds) if ds else float("inf")
            else:
                tot, cnt = 0.0, 0
                for u in paths[names[i]]:
                    if u in index and np.isfinite(lengths[j][index[u]]):
                        tot += lengths[j][index[u]]
                        cnt += 1
                for v in paths[names[j]]:
                    if v in index and np.isfinite(lengths[i][index[v]]):
                        tot += lengths[i][index[v]]
                        cnt += 1
                d = tot / cnt if cnt else float("inf")
            M[i, j] = M[j, i] = d
    return names, M
