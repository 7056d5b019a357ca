"""``gfa2network export --format edge-list`` over the GPU path (gfa2network/cli.py:267-281).

The reference streams the records of ``GFAParser`` and writes ``<from>\\t<to>\\n`` for every L / E / C record
(``<from>:<orientation>\\t<to>:<orientation>`` with ``--bidirected``).  Those endpoint strings are exactly the
node keys of a matrix build with the same ``--bidirected`` flag, so the same tokenizer kernel runs, and two
small kernels turn the stored edge records into the file's bytes on the device (csrc/ids.cuh:
k_edge_line_len / k_edge_line_write).  Exceptions, the one-shot warning and the partially written file of
a malformed input match the reference: everything before the offending line is written, then it raises."""
from __future__ import annotations

import sys
import warnings

import numpy as np

from . import _capi
from .builders import _FileSource, _default_device, _raise_parse_error, _read_source


def _edge_list_bytes(handle, host, dev_ptr, nbytes, keep, bidirected: bool):
    """Returns (bytes of the edge list, diag, rc) for the given source."""
    params = _capi.Params(1, int(bool(bidirected)), 1, 1, 0, _capi.DTYPES["float64"], _capi.FMT_NATIVE, 0 if dev_ptr is None else 1, None, 0, 0)
    if isinstance(keep, _FileSource) and keep.gz:
        rc = handle.build_gz(keep.path, params)
        if rc == _capi.G2N_ERR_INVALID:  # damaged container: the reference's own inflate raises its exception
            raw = np.frombuffer(keep.read_all(), dtype=np.uint8)
            rc = handle.build(raw.ctypes.data if raw.size else 0, raw.size, params)
    elif isinstance(keep, _FileSource):
        rc = handle.build_file(keep.path, params)
    else:
        ptr = dev_ptr if dev_ptr is not None else (host.ctypes.data if nbytes else 0)
        rc = handle.build(ptr, nbytes, params)
    diag = handle.status()
    data = handle.fetch_edge_list() if rc == _capi.G2N_OK else None
    return data, diag, rc


def edge_list_bytes(gfa, *, bidirected: bool = False, device: int | None = None) -> tuple[np.ndarray, BaseException | None, list[str]]:
    """The bytes ``export --format edge-list`` writes for *gfa*, the exception it raises afterwards (or
    None) and the warnings it issues.  *gfa* is anything ``parse_gfa`` accepts."""
    host, dev_ptr, nbytes, keep = _read_source(gfa)
    handle = _capi.default_handle(_default_device() if device is None else device)
    data, diag, rc = _edge_list_bytes(handle, host, dev_ptr, nbytes, keep, bidirected)
    warns = []
    if rc in (_capi.G2N_OK, _capi.G2N_ERR_PARSE) and diag.unknown_byte >= 0:
        warns.append(f"Skipping unsupported record: {bytes([diag.unknown_byte]).decode(errors='replace')}")  # parser.py:125-131
    exc: BaseException | None = None
    if rc == _capi.G2N_ERR_PARSE:
        try:
            _raise_parse_error(diag, keep if isinstance(keep, _FileSource) else host)
        except Exception as e:  # noqa: BLE001 - handed to the caller, who raises it after writing
            exc = e
        # the reference has already written the records before the offending line (cli.py:269-279):
        # the same bytes come from the text in front of that line
        cut = int(diag.err_offset)
        if isinstance(keep, _FileSource):
            prefix = np.frombuffer(keep.read_at(0, cut), dtype=np.uint8)
            data, _, rc2 = _edge_list_bytes(handle, prefix, None, cut, prefix, bidirected)
        elif dev_ptr is not None:
            data, _, rc2 = _edge_list_bytes(handle, None, dev_ptr, cut, keep, bidirected)
        else:
            data, _, rc2 = _edge_list_bytes(handle, host[:cut], None, cut, keep, bidirected)
        handle.check(rc2)
    else:
        handle.check(rc)
    if exc is None:
        # `u.decode()` / `v.decode()` per line (cli.py:278): the first undecodable name raises after the earlier lines
        try:
            data.tobytes().decode()
        except UnicodeDecodeError as e:
            raw = data.tobytes()
            keep_to = raw.rfind(b"\n", 0, e.start) + 1
            data, exc = data[:keep_to], e
    return data, exc, warns


def export_edge_list(gfa, *, bidirected: bool = False, output: str = "-", device: int | None = None) -> None:
    """cli.py:264-281 for ``--format edge-list``: write the edge list to *output* ("-" = stdout)."""
    data, exc, warns = edge_list_bytes(gfa, bidirected=bidirected, device=device)
    for w in warns:
        warnings.warn(w, RuntimeWarning, stacklevel=2)
    if output != "-":
        with open(output, "wb") as fh:
            fh.write(memoryview(data))
    else:
        sys.stdout.flush()
        sys.stdout.buffer.write(memoryview(data))
        sys.stdout.buffer.flush()
    if exc is not None:
        raise exc
