"""Path extraction and path-to-path distances over the GPU path (gfa2network/analysis.py:116-272).

* ``load_paths``: the P / O records of the file as ``{name: [segment, ...]}`` (analysis.py:164-177).  The
  records are split on the host (they are a few lines of the file, however long), after the device build
  has validated the whole file -- so malformed records raise exactly what the reference's parser raises.
* ``genome_distance_matrix`` / ``genome_distance(method="min")``: the reference runs
  ``networkx.multi_source_dijkstra_path_length`` per path on the graph ``parse_gfa(build_graph=True)`` builds
  without a weight tag -- hop counts along out-edges.  Here the adjacency stays in HBM as the CSR the matrix
  path builds (asymmetric for the default DiGraph, undirected for ``directed=False``) and every search is one
  cooperative multi-level BFS kernel (csrc/bfs.cuh); the per-pair minima / means are reduced on the device too.
Not provided (they need node sequences / the object graph): ``sequence_distance``, ``compute_stats``, the igraph backend."""
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
        return keep.read_all()
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
    # GFAParser streams the WHOLE file: unsupported-record warning and malformed-record errors come from there.
    # The source is consumed ONCE (stdin, a FIFO or a file object cannot be read twice): the device build and the
    # host split of the P / O lines see the same bytes.
    host, dev_ptr, _nbytes, keep = _read_source(path)
    if isinstance(keep, _FileSource):
        parse_gfa(keep.path, build_graph=False, build_matrix=True, asymmetric=True)
        text = keep.read_all()
    elif dev_ptr is not None:
        parse_gfa(path, build_graph=False, build_matrix=True, asymmetric=True)
        text = _text_bytes(path)
    else:
        parse_gfa(host, build_graph=False, build_matrix=True, asymmetric=True)
        text = host.tobytes() if not isinstance(keep, (bytes, bytearray)) else bytes(keep)
    return _split_paths(text, raw_bytes)


class _Graph:
    """The device-resident adjacency of one file plus the name -> node ID map (the reference's ``G``)."""

    def __init__(self, path, *, directed: bool = True, raw_bytes_id: bool = False, verbose: bool = False, device=None, want_index: bool = True):
        # DiGraph: rows = out-neighbours = the asymmetric CSR; Graph: the undirected build (builders.py:138-142, 226-228)
        kw = dict(asymmetric=True) if directed else dict(directed=False)
        res = parse_gfa(path, build_graph=False, build_matrix=True, return_node_list=want_index, raw_bytes_id=raw_bytes_id,
                        matrix_format="csr", verbose=verbose, device=device, **kw)
        self.A, nodes = res if want_index else (res, [])
        self.handle = self.A._g2n_session.handle
        self._keep = path  # a device tensor given as the source must outlive the searches (the text is read in place)
        self.index = {n: i for i, n in enumerate(nodes)}  # name -> node ID (callers that bring their own node names)

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


def _no_path(msg: str):
    try:
        import networkx as nx

        return nx.NetworkXNoPath(msg)
    except ImportError:  # pragma: no cover
        return ValueError(msg)


def genome_distance(gfa_path, nodes_a, nodes_b, *, method: str = "min", directed: bool = True, raw_bytes_id: bool = False, device=None) -> float:
    """analysis.py:116-161: the hop distance between two node sets of the graph of *gfa_path* (the reference takes the
    NetworkX graph; here the graph is the device-resident CSR of the file).  ``min``: one multi-source search from
    *nodes_a*; ``mean``: one search per node of *nodes_a*, averaged over the reachable pairs (analysis.py:141-159)."""
    import os

    nodes_a, nodes_b = list(nodes_a), list(nodes_b)
    if method not in ("min", "mean"):
        raise ValueError(f"unknown method: {method}")
    G = _Graph(gfa_path, directed=directed, raw_bytes_id=raw_bytes_id, device=device)
    if method == "min":
        G.bfs(G.ids(nodes_a, strict=True), 0, 1)
        mn, _, cnt = G.reduce(0, G.ids(nodes_b, strict=False))
        if cnt == 0:
            raise _no_path("no path between node sets")
        return mn
    if len(nodes_a) * len(nodes_b) > 1000 and os.getenv("GFANET_DISABLE_WARNINGS") != "1":
        warnings.warn("Mean distance scales quadratically; this may be very slow on large sets", RuntimeWarning)
    targets = G.ids(nodes_b, strict=False)  # a target that is not a node is "no path" for nx.shortest_path_length: skipped
    total, count = 0.0, 0
    for u in nodes_a:
        if not nodes_b:
            break
        G.bfs(G.ids([u], strict=True), 0, 1)
        _, s, c = G.reduce(0, targets)
        total += s
        count += c
    if count == 0:
        raise _no_path("no path between node sets")
    return total / count


class _DevicePaths:
    """The P / O records of the file behind a `_Graph`, resolved to node IDs on the device (csrc/paths.cuh): names and
    entry counts come to the host, the node lists do not."""

    def __init__(self, G: "_Graph", raw_bytes: bool):
        self.G, h = G, G.handle
        n = C.c_uint64()
        h.check(h.lib.g2n_paths_load(h.h, C.byref(n)))
        self.infos = []
        for i in range(n.value):
            info = _capi.PathInfo()
            h.check(h.lib.g2n_path_info(h.h, i, C.byref(info)))
            self.infos.append(info)
        # dict semantics of analysis.py:170-176: a later record of the same name replaces the list, not the position
        self.index: dict = {}
        for i, info in enumerate(self.infos):
            name = self._text(info.name_offset, info.name_len)
            self.index[name if raw_bytes else name.decode("ascii")] = i
        self.raw_bytes = raw_bytes

    def _text(self, off: int, n: int) -> bytes:
        h = self.G.handle
        buf = (C.c_uint8 * max(n, 1))()
        h.check(h.lib.g2n_fetch_text(h.h, off, n, buf))
        return bytes(buf[:n])

    def check_sources(self, i: int):
        """nx.multi_source_dijkstra raises for the first source that is not a node (analysis.py:237-239)."""
        info = self.infos[i]
        if info.missing_entry >= 0:
            name = self._text(info.missing_offset, info.missing_len)
            raise _node_not_found(name if self.raw_bytes else name.decode("ascii"))

    def bfs(self, i: int, slot: int, n_slots: int):
        h = self.G.handle
        h.check(h.lib.g2n_path_bfs(h.h, i, slot, n_slots))

    def reduce(self, slot: int, i: int):
        h = self.G.handle
        out = (C.c_int64 * 3)()
        h.check(h.lib.g2n_path_reduce(h.h, slot, i, out))
        return int(out[0]), int(out[1]), int(out[2])

    def nodes(self, i: int) -> np.ndarray:
        h = self.G.handle
        out = np.empty(int(self.infos[i].n_entries), dtype=np.int32)
        h.check(h.lib.g2n_fetch_path_nodes(h.h, i, C.c_void_p(out.ctypes.data)))
        return out


def genome_distance_matrix(gfa_path, method: str = "min", *, raw_bytes_id: bool = False, backend: str = "networkx", verbose: bool = False, device=None):
    """Return pairwise distances between all paths in *gfa_path* (analysis.py:180-272).  Nothing per path entry
    happens on the host: the records' node lists are resolved, searched from and reduced over on the device."""
    if backend != "networkx":
        raise NotImplementedError("backend='igraph' is outside the B200 path")
    G = _Graph(gfa_path, directed=True, raw_bytes_id=raw_bytes_id, verbose=verbose, device=device, want_index=False)
    P = _DevicePaths(G, raw_bytes_id)  # analysis.py:217 (the build above already validated the file)
    names = list(P.index)
    recs = [P.index[k] for k in names]
    n = len(names)
    M = np.zeros((n, n), dtype=float)
    # one multi-source search per path (analysis.py:236-240), all level arrays kept in HBM
    for k, r in enumerate(recs):
        P.check_sources(r)
        P.bfs(r, k, max(n, 1))
    for i in range(n):
        for j in range(i, n):
            if i == j:
                dist = 0.0
            elif method == "min":
                mn, _, cnt = P.reduce(i, recs[j])  # analysis.py:250-251
                dist = float(mn) if cnt else float("inf")
            else:  # mean of node-to-path distances, analysis.py:252-264
                _, s1, c1 = P.reduce(j, recs[i])
                _, s2, c2 = P.reduce(i, recs[j])
                dist = (float(s1) + float(s2)) / (c1 + c2) if (c1 + c2) else float("inf")
            M[i, j] = M[j, i] = dist
    try:
        import pandas as pd
    except Exception:  # pragma: no cover - optional dependency, as in the reference
        return M
    labels = [x.decode() if isinstance(x, bytes) else str(x) for x in names]
    return pd.DataFrame(M, index=labels, columns=labels)
