"""Host-side mirror of ``gfa2network.builders.parse_gfa`` for the matrix path.

Same name, keyword arguments, return convention, exceptions, warnings and verbose strings as
the reference (``gfa2network/builders.py:30-299``); the work itself -- tokenizer
(``parser.py:114-361``), node-ID assignment and triplet emission (``builders.py:163-234``),
``coo_matrix`` + ``maximum(A, A.T)`` (``builders.py:279-283``) -- runs in libg2n.so on the GPU.
Combinations outside the hot path (graph building, igraph backend, split-on-alignment) fail
loudly instead of silently computing something else.  There is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import gzip
import os
import sys
import warnings
from pathlib import Path

import numpy as np
import scipy.sparse as sp

from . import _capi

_ERRORS = {
    1: (ValueError, "Malformed L record"),  # parser.py:209
    2: (ValueError, "Malformed E record"),  # parser.py:252
    3: (ValueError, "Malformed C record"),  # parser.py:300
    4: (ValueError, "Malformed P record"),  # parser.py:232
    5: (ValueError, "Malformed O record"),  # parser.py:346
    6: (IndexError, "list index out of range"),  # parser.py:163 fields[1]
    7: (IndexError, "index out of range"),  # parser.py:220-221 u_field[-1]
    9: (OverflowError, "int too large to convert to float"),  # builders.py:209
    10: (NotImplementedError, "non-ASCII numeric weight tag value is outside the supported scope"),
}


def _default_device() -> int:
    for k in ("G2N_DEVICE", "LOCAL_RANK"):
        v = os.environ.get(k)
        if v is not None and v.isdigit():
            return int(v)
    return 0


class _Session:
    """Ties a returned SciPy matrix to the device-resident result it was fetched from (node map, edge list,
    distances and the CLI's format conversion are made from that result).  The tie holds only while the result is
    still the handle's current one: ``Handle.generation`` is bumped by every call that replaces it (``build``,
    ``build_file``, ``coo_to_compressed``)."""

    def __init__(self, handle: _capi.Handle):
        self.handle = handle
        self.generation = handle.generation

    def live(self) -> bool:
        return self.handle.h is not None and self.handle.generation == self.generation


class _FileSource:
    gz = False

    def __init__(self, path: str):
        self.path = path

    def read_at(self, off: int, n: int) -> bytes:
        with open(self.path, "rb") as fh:
            fh.seek(off)
            return fh.read(n)

    def read_all(self) -> bytes:
        with open(self.path, "rb") as fh:
            return fh.read()


class _GzSource(_FileSource):
    """A *.gz path: inflated by the library (windowed, BGZF block-parallel) on its way to the device."""
    gz = True

    def read_at(self, off: int, n: int) -> bytes:
        with gzip.open(self.path, "rb") as fh:  # (only to re-split one offending line for an exception message)
            fh.seek(off)
            return fh.read(n)

    def read_all(self) -> bytes:
        with gzip.open(self.path, "rb") as fh:
            return fh.read()


def _read_source(path):
    """Returns (host uint8 array | None, device pointer | None, nbytes, keepalive)."""
    if hasattr(path, "is_cuda") and hasattr(path, "data_ptr"):  # torch tensor (extension)
        t = path
        if not t.is_cuda:
            arr = t.numpy()
            return np.ascontiguousarray(arr.view(np.uint8).reshape(-1)), None, arr.nbytes, t
        if not t.is_contiguous():
            raise ValueError("device text tensor must be contiguous")
        return None, t.data_ptr(), t.numel() * t.element_size(), t
    if isinstance(path, np.ndarray):
        arr = np.ascontiguousarray(path).view(np.uint8).reshape(-1)
        return arr, None, arr.size, arr
    if isinstance(path, (bytes, bytearray, memoryview)):
        arr = np.frombuffer(path, dtype=np.uint8)
        return arr, None, arr.size, path
    if isinstance(path, (str, Path)):
        p = str(path) or "-"  # parser.py:104
        if p == "-":
            raw = sys.stdin.buffer.read()  # parser.py:105-106
        elif p.endswith(".gz"):
            st = os.stat(p)  # FileNotFoundError as gzip.open() (parser.py:108-109)
            import stat as _stat

            if _stat.S_ISREG(st.st_mode) and os.access(p, os.R_OK) and not os.environ.get("G2N_HOST_GZIP"):
                return None, None, st.st_size, _GzSource(p)  # g2n_build_gz inflates on the way to the device
            with gzip.open(p, "rb") as fh:
                raw = fh.read()
        else:
            # parser.py:111 open(path, "rb"): the library reads the file itself (g2n_build_file: reader threads ->
            # pinned staging buffers -> device, tokenized piece by piece behind the copy)
            import stat as _stat

            st = os.stat(p)  # FileNotFoundError as the reference's open()
            if not _stat.S_ISREG(st.st_mode):
                # FIFOs, /dev/stdin, <(zcat x.gz): read once on the host like the reference's buffered open()
                # (IsADirectoryError for a directory); never opened twice -- a second open of a FIFO would SIGPIPE the writer
                with open(p, "rb") as fh:
                    raw = fh.read()
                arr = np.frombuffer(raw, dtype=np.uint8)
                return arr, None, arr.size, raw
            if not os.access(p, os.R_OK):
                raise PermissionError(13, "Permission denied", p)
            return None, None, st.st_size, _FileSource(p)
        arr = np.frombuffer(raw, dtype=np.uint8)
        return arr, None, arr.size, raw
    if hasattr(path, "read"):  # binary file object, parser.py:90-92
        raw = path.read()
        arr = np.frombuffer(raw, dtype=np.uint8)
        return arr, None, arr.size, raw
    raise TypeError(f"unsupported GFA source: {type(path)!r}")


def _raise_parse_error(diag, host):
    kind = diag.err_kind
    if kind == 8:
        # UnicodeDecodeError from an orientation field: re-split the one offending line on the
        # host to raise the identical exception (parser.py:214, 291-293, 337-339)
        if host is not None:
            off = int(diag.err_offset)
            tail = host.read_at(off, 1 << 20) if isinstance(host, _FileSource) else host[off:off + (1 << 20)].tobytes()
            line = tail.split(b"\n", 1)[0]
            f = line.split(b"\t")
            cand = {b"L": (4,), b"E": (3, 5), b"C": (2, 4)}.get(f[0], ())
            for i in cand:
                if i < len(f):
                    f[i].decode()
        raise UnicodeDecodeError("utf-8", b"", 0, 1, "invalid orientation field")
    exc, msg = _ERRORS[kind]
    raise exc(msg)


def _node_list(handle, raw_bytes_id):
    names, offs = handle.fetch_names()
    blob = names.tobytes()
    o = offs.tolist()
    if raw_bytes_id:
        return [blob[o[i]:o[i + 1]] for i in range(len(o) - 1)]
    if blob.isascii():
        s = blob.decode("ascii")
        return [s[o[i]:o[i + 1]] for i in range(len(o) - 1)]
    return [blob[o[i]:o[i + 1]].decode() for i in range(len(o) - 1)]  # builders.py:287


def _matrix_from_handle(handle):
    """SciPy matrix over the fetched arrays.  The arrays come straight from the device build (sorted,
    in range, right dtypes), so SciPy's O(nnz) constructor validation is skipped by filling an empty
    matrix object."""
    s, a0, a1, data = handle.fetch_matrix()
    n = s.n_nodes
    if s.format == _capi.FMT_COO:
        A = sp.coo_matrix((n, n), dtype=data.dtype)
        A.data = data
        if hasattr(A, "coords"):
            A.coords = (a0, a1)
        else:  # SciPy < 1.13
            A.row, A.col = a0, a1
        A.has_canonical_format = False
        return A
    cls = sp.csr_matrix if s.format == _capi.FMT_CSR else sp.csc_matrix
    A = cls((n, n), dtype=data.dtype)
    A.data, A.indices, A.indptr = data, a1, a0
    return A


def parse_gfa(
    path,
    *,
    build_graph: bool,
    build_matrix: bool,
    directed: bool = True,
    weight_tag: str | None = None,
    store_seq: bool = False,
    store_tags: bool = False,
    strip_orientation: bool = False,
    verbose: bool = False,
    bidirected: bool = False,
    keep_directed_bidir: bool = False,
    backend: str = "networkx",
    dtype: str | object = "float64",
    asymmetric: bool = False,
    raw_bytes_id: bool = False,
    return_node_list: bool = False,
    max_tag_mb: float = 100.0,
    split_on_alignment: bool = False,
    matrix_format: str | None = None,
    device: int | None = None,
    devices: list[int] | None = None,
):
    """Parse *path* on the GPU and return the adjacency matrix (and node list).

    Drop-in for ``gfa2network.parse_gfa(..., build_graph=False, build_matrix=True)``
    (``builders.py:30-50``).  Extensions, all optional: ``matrix_format`` ("csr"/"csc") fuses
    ``convert_format`` into the device build; ``device`` selects the CUDA device; ``devices=[...]``
    (two to eight GPUs of this host) splits the FILE at newline boundaries, one byte range per GPU, and
    builds one graph over all of them (``dist.MultiGpuBuilder``: each GPU ends with the CSR/CSC slab of its
    row block; the slabs and name ranges are concatenated for the caller).  ``path`` may also be a
    bytes-like object, a uint8 NumPy array or a CUDA uint8 torch tensor (single device only).
    """
    if backend == "igraph":
        raise NotImplementedError("backend='igraph' is outside the B200 GFA->matrix path (builders.py:95-109)")
    if backend != "networkx":
        raise ValueError(f"unknown backend {backend!r}")
    if split_on_alignment:
        raise NotImplementedError("split_on_alignment is outside the B200 GFA->matrix path (builders.py:110-128)")
    if return_node_list and not build_matrix:
        raise ValueError("return_node_list requires build_matrix=True")  # builders.py:130
    if build_graph:
        raise NotImplementedError(
            "build_graph=True (NetworkX object graph) is outside the B200 GFA->matrix path; "
            "use the reference for graphs (builders.py:138-142, 166-189, 235-256)")
    dt = np.dtype(dtype)  # builders.py:280
    if dt.name not in _capi.DTYPES:
        raise NotImplementedError(f"dtype {dt.name!r}: the device path supports {sorted(_capi.DTYPES)}")
    want = _capi.FMT_NATIVE
    if matrix_format is not None:
        mf = matrix_format.lower()
        if mf not in {"csr", "csc", "coo", "dok"}:
            raise ValueError("matrix-format must be csr|csc|coo|dok")  # utils.py:46
        want = {"csr": _capi.FMT_CSR, "csc": _capi.FMT_CSC}.get(mf, _capi.FMT_NATIVE)

    if devices is not None and len(devices) > 1:
        return _parse_gfa_multi(path, list(devices), build_matrix=build_matrix, directed=directed, weight_tag=weight_tag,
                                strip_orientation=strip_orientation, verbose=verbose, bidirected=bidirected,
                                keep_directed_bidir=keep_directed_bidir, dtype=dt.name, asymmetric=asymmetric,
                                raw_bytes_id=raw_bytes_id, return_node_list=return_node_list, matrix_format=matrix_format)
    if devices is not None and len(devices) == 1 and device is None:
        device = devices[0]
    host, dev_ptr, nbytes, keep = _read_source(path)
    handle = _capi.default_handle(_default_device() if device is None else device)
    wt = weight_tag.encode() if weight_tag else None  # builders.py:206 "if weight_tag and ..."
    params = _capi.Params(
        int(bool(directed)), int(bool(bidirected)), int(bool(keep_directed_bidir)), int(bool(asymmetric)),
        int(bool(strip_orientation)), _capi.DTYPES[dt.name], want, 0 if dev_ptr is None else 1,
        wt, len(wt) if wt else 0, 0)
    if isinstance(keep, _GzSource):
        host = keep
        rc = handle.build_gz(keep.path, params)
        if rc == _capi.G2N_ERR_INVALID:
            # damaged or not-gzip container: the reference's own inflate raises the reference's exception
            # (gzip.BadGzipFile / EOFError / zlib.error); if it does not, its bytes are built from the host
            with gzip.open(keep.path, "rb") as fh:
                raw = fh.read()
            host = np.frombuffer(raw, dtype=np.uint8)
            keep = raw
            nbytes = host.size
            rc = handle.build(host.ctypes.data if nbytes else 0, nbytes, params)
    elif isinstance(keep, _FileSource):
        host = keep
        rc = handle.build_file(keep.path, params)
    else:
        ptr = dev_ptr if dev_ptr is not None else (host.ctypes.data if nbytes else 0)
        rc = handle.build(ptr, nbytes, params)
    diag = handle.status()
    if rc in (_capi.G2N_OK, _capi.G2N_ERR_PARSE):
        if diag.unknown_byte >= 0:
            # parser.py:125-131, once per parse, before any error of a later line
            warnings.warn(
                f"Skipping unsupported record: {bytes([diag.unknown_byte]).decode()}",
                RuntimeWarning, stacklevel=2)
        if verbose:
            for k in range(500_000, int(diag.n_records) + 1, 500_000):  # builders.py:257-258
                print(f"\r[{k:,} lines]", end="", file=sys.stderr)
    if rc == _capi.G2N_ERR_PARSE:
        _raise_parse_error(diag, host)
    handle.check(rc)
    if verbose:
        print("\r[parse_gfa] done")  # builders.py:261
    if build_matrix and diag.warn_flags & 1:
        # NumPy's float64 -> float32 list cast inside sp.coo_matrix (builders.py:281)
        warnings.warn("overflow encountered in cast", RuntimeWarning, stacklevel=2)
    if not build_matrix:
        return None
    A = _matrix_from_handle(handle)
    A._g2n_session = _Session(handle)
    if return_node_list:
        return A, _node_list(handle, raw_bytes_id)
    return A


_multi: dict[tuple, object] = {}


def _parse_gfa_multi(path, devices, *, build_matrix, directed, weight_tag, strip_orientation, verbose, bidirected,
                     keep_directed_bidir, dtype, asymmetric, raw_bytes_id, return_node_list, matrix_format):
    """parse_gfa over several GPUs of this host (one process): the file is split at newline boundaries, one byte range
    per GPU (north_star; reference surface cli.py:206-225).  Returns what parse_gfa returns on one GPU for the modes whose
    result is compressed: the CSR of max(S, S^T) in the default directed mode (builders.py:282-283), or the CSR / CSC asked
    for with ``matrix_format``.  A raw COO in emission order (non-symmetric modes without ``matrix_format``) is a single-GPU result."""
    from .dist import MultiGpuBuilder

    if not isinstance(path, (str, Path)) or str(path) in ("", "-") or str(path).endswith(".gz"):
        raise NotImplementedError("devices=[...] reads a plain file by byte ranges: give a path to an uncompressed GFA file")
    p = str(path)
    st = os.stat(p)
    import stat as _stat

    if not _stat.S_ISREG(st.st_mode):
        raise NotImplementedError("devices=[...] needs a regular file (byte ranges are read by offset)")
    graph_directed = keep_directed_bidir or (not bidirected and directed)  # builders.py:143
    symmax = graph_directed and not asymmetric
    fmt = (matrix_format or "").lower()
    if fmt and fmt not in {"csr", "csc", "coo", "dok"}:
        raise ValueError("matrix-format must be csr|csc|coo|dok")
    if fmt not in ("csr", "csc"):
        if not symmax:
            raise NotImplementedError("a multi-GPU build ends in CSR / CSC slabs: pass matrix_format='csr' or 'csc' "
                                      "(the raw COO in emission order is the single-GPU result)")
        fmt = "csr"
    key = tuple(devices)
    mg = _multi.get(key)
    if mg is None:
        mg = _multi[key] = MultiGpuBuilder(devices)
    with warnings.catch_warnings(record=True) as caught:
        warnings.simplefilter("always")
        try:
            mg.build_file(p, directed=directed, bidirected=bidirected, keep_directed_bidir=keep_directed_bidir, asymmetric=asymmetric,
                          strip_orientation=strip_orientation, dtype=dtype, matrix_format=fmt, weight_tag=weight_tag)
            exc = None
        except Exception as e:  # noqa: BLE001 - re-raised below, after the warning that precedes it in the file
            exc = e
    for w in caught:
        warnings.warn(str(w.message), w.category, stacklevel=3)
    if exc is not None:
        raise exc
    if verbose:
        for k in range(500_000, mg.diag_records() + 1, 500_000):  # builders.py:257-258
            print(f"\r[{k:,} lines]", end="", file=sys.stderr)
        print("\r[parse_gfa] done")  # builders.py:261
    if not build_matrix:
        return None
    A = mg.matrix(fmt)
    if return_node_list:
        return A, mg.node_list(raw_bytes_id)
    return A
