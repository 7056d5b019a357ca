"""Host-side mirror of ``gfa2network/utils.py``: ``convert_format`` (``:40-63``) finishes
COO -> CSR/CSC on the GPU (stage K4); ``save_matrix`` (``:66-105``) writes the same ``.npz`` members with a
multi-threaded deflate (``writers.py``); ``save_node_map`` (``:108-114``) is the reference's host writer for
arbitrary node lists -- the CLI writes the node map of a device build from GPU-made bytes instead."""
from __future__ import annotations

import ctypes as C
import sys
import time
from pathlib import Path
from typing import Sequence

import numpy as np
import scipy.sparse as sp

from . import _capi


def _np_dtype_code(dt: np.dtype):
    return _capi.DTYPES.get(dt.name)


def convert_format(A, fmt: str, *, verbose: bool = False, _untouched: bool = False):
    """Convert COO -> *fmt* (``utils.py:40-63``).  ``coo`` returns *A* untouched even when *A*
    is CSR.  COO -> csr/csc runs on the device (stage K4) on the triplets *A* holds NOW, so
    edits made to ``A.data`` / ``A.row`` / ``A.col`` after ``parse_gfa`` are honoured exactly like the
    reference's ``A.asformat`` honours them.  ``_untouched`` (internal: the CLI, which converts the matrix it
    has just parsed) lets the conversion start from the device-resident build instead of re-uploading."""
    fmt = fmt.lower()
    if fmt not in {"csr", "csc", "coo", "dok"}:
        raise ValueError("matrix-format must be csr|csc|coo|dok")
    if fmt == "coo":
        return A
    if verbose:
        start = time.perf_counter()
        print(f"[convert] -> {fmt} …", end="", file=sys.stderr, flush=True)
    out = _convert(A, fmt, _untouched)
    if verbose:
        print(f" done in {time.perf_counter() - start:,.1f}s", file=sys.stderr)
    return out


def _convert(A, fmt: str, untouched: bool = False):
    if A.format == fmt:
        return A  # scipy/_base.py:471-501 asformat: same format -> self
    want = {"csr": _capi.FMT_CSR, "csc": _capi.FMT_CSC}.get(fmt)
    session = getattr(A, "_g2n_session", None)
    if untouched and want is not None and session is not None and A.format in ("coo", "csr") and session.live():
        from .builders import _matrix_from_handle

        session.handle.convert(want)
        out = _matrix_from_handle(session.handle)
        out._g2n_session = session
        return out
    code = _np_dtype_code(A.dtype)
    if want is not None and A.format == "coo" and code is not None and A.shape[0] == A.shape[1] \
            and A.row.dtype == np.int32 and A.shape[0] > 0 and A.nnz > 0:
        # any other COO matrix: stage K4 alone (g2n_coo_to_compressed)
        from .builders import _default_device

        h = _capi.default_handle(_default_device())
        n, nnz = A.shape[0], A.nnz
        row = np.ascontiguousarray(A.row)
        col = np.ascontiguousarray(A.col)
        data = np.ascontiguousarray(A.data)
        indptr = np.empty(n + 1, np.int32)
        indices = np.empty(nnz, np.int32)
        dout = np.empty(nnz, A.dtype)
        k = h.coo_to_compressed(row, col, data, nnz, n, code, want, indptr, indices, dout)
        cls = sp.csr_matrix if fmt == "csr" else sp.csc_matrix
        return cls((dout[:k].copy(), indices[:k].copy(), indptr), shape=A.shape)
    # host-only object formats (dok) and conversions that are not on the GFA->matrix path
    return A.asformat(fmt)


def save_matrix(A, dest: Path, *, verbose: bool = False, max_dense_gb: float = 5.0):
    """Write *A* to *dest* (.npz sparse, .npy/.csv dense with the size guard of utils.py:70-77)."""
    dest = Path(dest)
    limit = max_dense_gb * 1_000_000_000
    if dest.suffix in {".csv", ".npy"}:
        nnz = A.nnz if sp.issparse(A) else A.size
        itemsize = A.dtype.itemsize if hasattr(A, "dtype") else 8
        if nnz * itemsize > limit:
            raise MemoryError(
                f"dense export would allocate {nnz*itemsize/1e9:.1f} GB; choose a sparse .npz or write an edge list instead")
    if verbose:
        start = time.perf_counter()
        print(f"[save] {dest.suffix[1:]} → {dest}", "...", end="", file=sys.stderr, flush=True)
    if dest.suffix == ".npz":
        from .writers import save_npz_parallel

        save_npz_parallel(dest, A)  # same members as sp.save_npz(dest, A) (utils.py:86), deflated on all host cores
    elif dest.suffix == ".npy":
        np.save(dest, A.toarray() if sp.issparse(A) else A)
    elif dest.suffix == ".csv":
        np.savetxt(dest, A.toarray() if sp.issparse(A) else A, delimiter=",", fmt="%.6g")
    else:
        raise ValueError("matrix path must end with .npz, .npy, or .csv")
    if verbose:
        print(f" done in {time.perf_counter() - start:,.1f}s", file=sys.stderr)


def save_node_map(nodes: Sequence[bytes | str], dest: Path) -> None:
    """``<index>\\t<name>`` per line (utils.py:108-114)."""
    with open(dest, "w") as fh:
        fh.write("".join(f"{i}\t{n.decode() if isinstance(n, (bytes, bytearray)) else n}\n" for i, n in enumerate(nodes)))
