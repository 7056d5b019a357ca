"""Path extraction and path-to-path distances over the GPU path (gfa2network/analysis.py:116-272).

* ``load_paths``: the P / O records of the file as ``{name: [segment, ...]}`` (analysis.py:164-177).  The
  records are split on the host (they are a few lines of the file, however long), after the device build
  has validated the whole file -- so malformed records raise exactly what the reference's parser raises.
* ``genome_distance_matrix`` / ``genome_distance(method="min")``: the reference runs
  ``networkx.multi_source_dijkstra_path_length`` per path on the graph ``parse_gfa(build_graph=True)`` builds
  without a weight tag -- hop counts along out-edges.  Here the adjacency stays in HBM as the CSR the matrix
  path builds (asymmetric for the default DiGraph, undirected for ``directed=False``) and every search is one
  cooperative multi-level BFS kernel (csrc/bfs.cuh); the per-pair minima / means are reduced on the device too.
Not provided (they need node sequences / per-pair Dijkstra on the object graph): ``sequence_distance``,
``genome_distance(method="mean")``, ``compute_stats``, the igraph backend."""
from __future__ import annotations

import ctypes as C
import warnings

import numpy as np

from . import _capi
from .builders import _FileSource, _read_source, parse_gfa


def _node_not_found(name):
    try:
        import networkx as nx

        return nx.NodeNotFound(f"Node {name} not found in graph")
    except Exception:  # pragma: no cover - networkx is a dependency of the reference, not of this package
        return KeyError(f"Node {name} not found in graph")


def _text_bytes(path) -> bytes:
    host, dev_ptr, nbytes, keep = _read_source(path)
    if isinstance(keep, _FileSource):
        return keep.read_at(0, nbytes)
    if dev_ptr is not None:
        return bytes(keep.cpu().numpy().tobytes())
    return host.tobytes()


def _split_paths(text: bytes, raw_bytes: bool):
    """analysis.py:164-177 over parser.py:229-247 / 343-361: P and O records in file order, later names overwrite."""
    paths = {}
    buf = np.frombuffer(text, dtype=np.uint8)
    # a record line starts at byte 0 or after '\n' with b"P\t" / b"O\t" (parser.py:117-134: the first FIELD must be P / O)
    starts = np.flatnonzero(buf[:-1] == 10) + 1 if buf.size > 1 else np.zeros(0, np.int64)
    starts = np.concatenate([[0], starts]) if buf.size else starts
    starts = starts[starts + 1 < buf.size]
    first, second = buf[starts], buf[starts + 1]
    sel = starts[((first == ord("P")) | (first == ord("O"))) & (second == 9)]
    for s in sel.tolist():
        e = text.find(b"\n", s)
        fields = text[s: e if e >= 0 else len(text)].split(b"\t")
        if len(fields) < 3:  # the device build has raised "Malformed P/O record" for these already
            raise ValueError(f"Malformed {fields[0].decode()} record")
        segs = []
        for entry in fields[2].split(b","):
            seg = entry[:-1] if entry.endswith((b"+", b"-")) else entry
            segs.append(seg if raw_bytes else seg.decode("ascii"))
        paths[fields[1] if raw_bytes else fields[1].decode("ascii")] = segs
    return paths


def load_paths(path, *, raw_bytes: bool = False):
    """Return mapping of path/walk names to their node lists (analysis.py:164-177)."""
    # GFAParser streams the WHOLE file: unsupported-record warning and malformed-record errors come from there
    parse_gfa(path, build_graph=False, build_matrix=True, asymmetric=True)
    return _split_paths(_text_bytes(path), raw_bytes)


class _Graph:
    """The device-resident adjacency of one file plus the name -> node ID map (the reference's ``G``)."""

    def __init__(self, path, *, directed: bool = True, raw_bytes_id: bool = False, verbose: bool = False, device=None):
        # DiGraph: rows = out-neighbours = the asymmetric CSR; Graph: the undirected build (builders.py:138-142, 226-228)
        kw = dict(asymmetric=True) if directed else dict(directed=False)
        self.A, nodes = parse_gfa(path, build_graph=False, build_matrix=True, return_node_list=True, raw_bytes_id=raw_bytes_id,
                                  matrix_format="csr", verbose=verbose, device=device, **kw)
        self.handle = self.A._g2n_session.handle
        self.index = {n: i for i, n in enumerate(nodes)}
        self.n_slots = 0

    def ids(self, names, *, strict: bool) -> np.ndarray:
        out = []
        for nm in names:
            i = self.index.get(nm)
            if i is None:
                if strict:
                    raise _node_not_found(nm)
                continue
            out.append(i)
        return np.asarray(out, dtype=np.int32)

    def bfs(self, sources: np.ndarray, slot: int, n_slots: int):
        h = self.handle
        h.check(h.lib.g2n_bfs(h.h, C.c_void_p(sources.ctypes.data if sources.size else 0), int(sources.size), slot, n_slots))

    def reduce(self, slot: int, nodes: np.ndarray):
        h = self.handle
        out = (C.c_int64 * 3)()
        h.check(h.lib.g2n_levels_reduce(h.h, slot, C.c_void_p(nodes.ctypes.data if nodes.size else 0), int(nodes.size), out))
        return int(out[0]), int(out[1]), int(out[2])

    def levels(self, slot: int) -> np.ndarray:
        h = self.handle
        out = np.empty(self.A.shape[0], dtype=np.int32)
        h.check(h.lib.g2n_fetch_levels(h.h, slot, C.c_void_p(out.ctypes.data)))
        return out


def genome_distance(gfa_path, nodes_a, nodes_b, *, method: str = "min", directed: bool = True, raw_bytes_id: bool = False, device=None) -> float:
    """analysis.py:116-161 for ``method="min"``: the hop distance between two node sets of the graph of *gfa_path*
    (the reference takes the NetworkX graph; here the graph is the device-resident CSR of the file)."""
    if method == "mean":
        raise NotImplementedError("genome_distance(method='mean') runs one Dijkstra per node pair on the object graph (analysis.py:147-159)")
    if method != "min":
        raise ValueError(f"unknown method: {method}")
    G = _Graph(gfa_path, directed=directed, raw_bytes_id=raw_bytes_id, device=device)
    G.bfs(G.ids(list(nodes_a), strict=True), 0, 1)
    mn, _, cnt = G.reduce(0, G.ids(list(nodes_b), strict=False))
    if cnt == 0:
        try:
            import networkx as nx

            raise nx.NetworkXNoPath("no path between node sets")
        except ImportError:  # pragma: no cover
            raise ValueError("no path between node sets") from None
    return mn


def genome_distance_matrix(gfa_path, method: str = "min", *, raw_bytes_id: bool = False, backend: str = "networkx", verbose: bool = False, device=None):
    """Return pairwise distances between all paths in *gfa_path* (analysis.py:180-272)."""
    if backend != "networkx":
        raise NotImplementedError("backend='igraph' is outside the B200 path")
    G = _Graph(gfa_path, directed=True, raw_bytes_id=raw_bytes_id, verbose=verbose, device=device)
    paths = _split_paths(_text_bytes(gfa_path), raw_bytes_id)  # analysis.py:217 (the build above already validated the file)
    names = list(paths)
    n = len(names)
    M = np.zeros((n, n), dtype=float)
    # one multi-source search per path (analysis.py:236-240), all level arrays kept in HBM
    lists = []
    for k, name in enumerate(names):
        G.bfs(G.ids(paths[name], strict=True), k, max(n, 1))
        lists.append(G.ids(paths[name], strict=False))
    for i in range(n):
        for j in range(i, n):
            if i == j:
                dist = 0.0
            elif method == "min":
                mn, _, cnt = G.reduce(i, lists[j])  # analysis.py:250-251
                dist = float(mn) if cnt else float("inf")
            else:  # mean of node-to-path distances, analysis.py:252-264
                _, s1, c1 = G.reduce(j, lists[i])
                _, s2, c2 = G.reduce(i, lists[j])
                dist = (float(s1) + float(s2)) / (c1 + c2) if (c1 + c2) else float("inf")
            M[i, j] = M[j, i] = dist
    try:
        import pandas as pd
    except Exception:  # pragma: no cover - optional dependency, as in the reference
        return M
    labels = [x.decode() if isinstance(x, bytes) else str(x) for x in names]
    return pd.DataFrame(M, index=labels, columns=labels)
