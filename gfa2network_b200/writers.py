"""Result writers of the `convert` command, mirrors of utils.py:66-114.

`save_npz_parallel` writes what `scipy.sparse.save_npz(dest, A)` (utils.py:86) writes -- a zip of the
`.npy` members `indices/indptr` (or `row/col`), `format`, `shape`, `data`, deflate-compressed -- but
compresses every member in independent 1 MiB blocks on all host cores (zlib releases the GIL).  The blocks
are raw-deflate streams ended with a full flush, so their concatenation is one valid deflate stream (the
pigz construction); `scipy.sparse.load_npz` / `numpy.load` / `zipfile` read the file like any other.
Once the matrix build takes milliseconds the single-threaded zlib of `save_npz` is the wall clock of
`gfa2network convert` (about 1 s per 1.2 M stored entries).

`write_node_map` produces `<index>\\t<name>\\n` per node (utils.py:108-114) on the GPU: line lengths,
prefix sum and the bytes themselves are device kernels (csrc/ids.cuh k_tsv_*); the host only writes the
buffer to the file."""
from __future__ import annotations

import io
import os
import struct
import zlib
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

import numpy as np
import scipy.sparse as sp

_BLOCK = 1 << 20


def _npy_bytes_header(arr: np.ndarray) -> bytes:
    """The .npy header numpy.save would write for *arr* (version chosen by numpy)."""
    buf = io.BytesIO()
    np.lib.format.write_array_header_1_0(buf, np.lib.format.header_data_from_array_1_0(arr))
    return buf.getvalue()


def _members(A) -> list[tuple[str, np.ndarray]]:
    """Member order and contents of scipy.sparse.save_npz (scipy/sparse/_matrix_io.py)."""
    if A.format in ("csr", "csc"):
        out = [("indices", A.indices), ("indptr", A.indptr)]
    elif A.format == "coo":
        out = [("row", A.row), ("col", A.col)]
    else:
        raise NotImplementedError(f"Save is not implemented for sparse matrix of format {A.format}.")
    out += [("format", np.asanyarray(A.format.encode("ascii"))), ("shape", np.asanyarray(A.shape)), ("data", A.data)]
    if isinstance(A, sp.sparray):
        out.append(("_is_array", np.asanyarray(True)))
    return [(k, v if v.flags.c_contiguous else np.ascontiguousarray(v)) for k, v in ((k, np.asanyarray(v)) for k, v in out)]


# ---- crc32(A || B) from crc32(A), crc32(B), len(B): zlib's crc32_combine (GF(2) matrix squaring), which
# Python's zlib module does not export.  Lets every block's CRC be computed by the worker that deflates it.
def _gf2_times(mat, vec):
    s, i = 0, 0
    while vec:
        if vec & 1:
            s ^= mat[i]
        vec >>= 1
        i += 1
    return s


def _gf2_square(mat):
    return [_gf2_times(mat, mat[n]) for n in range(32)]


def crc32_combine(crc1: int, crc2: int, len2: int) -> int:
    if len2 <= 0:
        return crc1
    odd = [0xEDB88320] + [1 << n for n in range(31)]  # operator for one zero bit
    even = _gf2_square(odd)   # two zero bits
    odd = _gf2_square(even)   # four zero bits
    while True:
        even = _gf2_square(odd)  # first pass: one zero byte
        if len2 & 1:
            crc1 = _gf2_times(even, crc1)
        len2 >>= 1
        if not len2:
            break
        odd = _gf2_square(even)
        if len2 & 1:
            crc1 = _gf2_times(odd, crc1)
        len2 >>= 1
        if not len2:
            break
    return crc1 ^ crc2


def _deflate_block(args):
    parts, last, level = args  # parts: buffers that make up this block, in order
    c = zlib.compressobj(level, zlib.DEFLATED, -15)
    out, crc, n = [], 0, 0
    for p in parts:
        out.append(c.compress(p))
        crc = zlib.crc32(p, crc)
        n += len(p)
    out.append(c.flush(zlib.Z_FINISH if last else zlib.Z_FULL_FLUSH))
    return b"".join(out), crc, n


def save_npz_parallel(dest, A, *, level: int = 6, threads: int | None = None) -> None:
    """Write *A* as a compressed .npz that `scipy.sparse.load_npz` reads back unchanged."""
    dest = Path(dest)
    members = _members(A)
    # no ZIP64 records here: every member AND every offset (so the whole archive; deflate never grows a block by more
    # than a few bytes per 16 KiB) must stay below 4 GiB -- otherwise leave the file to numpy's zipfile writer
    if sum(v.nbytes + v.nbytes // 1000 + 4096 for _, v in members) >= 0xFFFF0000:
        sp.save_npz(dest, A)
        return
    threads = threads or os.cpu_count() or 1
    with ThreadPoolExecutor(max_workers=threads) as pool:
        # every block of every member is queued at once; the array bytes are never copied on the host
        futures = []
        for name, arr in members:
            header = _npy_bytes_header(arr)
            body = memoryview(arr.reshape(-1).view(np.uint8)) if arr.nbytes else memoryview(b"")
            room = _BLOCK - len(header)
            cuts = [0] + list(range(room, len(body), _BLOCK)) if len(body) > room else [0]
            jobs = []
            for j, a in enumerate(cuts):
                b = cuts[j + 1] if j + 1 < len(cuts) else len(body)
                parts = ([header] if j == 0 else []) + [body[a:b]]
                jobs.append(pool.submit(_deflate_block, (parts, j + 1 == len(cuts), level)))
            futures.append((name, jobs))
        central = []
        tmp = dest.with_name(f".{dest.name}.{os.getpid()}.tmp")  # a failure never leaves a truncated archive under the final name
        with open(tmp, "wb") as fh:
            for name, jobs in futures:
                fname = (name + ".npy").encode()
                blocks, crc, usize = [], 0, 0
                for f in jobs:
                    data, c, n = f.result()
                    blocks.append(data)
                    crc = crc32_combine(crc, c, n) if usize else c
                    usize += n
                csize = sum(len(b) for b in blocks)
                offset = fh.tell()
                if max(usize, csize, offset) >= 0xFFFFFFFF:
                    raise OverflowError("member needs ZIP64")  # (unreachable: guarded by the total-size check above)
                # local file header: version 2.0, no flags, deflate, DOS date 1980-01-01 (the timestamp is not part of the data)
                fh.write(struct.pack("<IHHHHHIIIHH", 0x04034B50, 20, 0, 8, 0, 0x0021, crc, csize, usize, len(fname), 0))
                fh.write(fname)
                for b in blocks:
                    fh.write(b)
                central.append((fname, crc, csize, usize, offset))
            cd_start = fh.tell()
            for fname, crc, csize, usize, offset in central:
                fh.write(struct.pack("<IHHHHHHIIIHHHHHII", 0x02014B50, 20, 20, 0, 8, 0, 0x0021, crc, csize, usize, len(fname), 0, 0, 0, 0,
                                     0o600 << 16, offset))
                fh.write(fname)
            cd_size = fh.tell() - cd_start
            fh.write(struct.pack("<IHHHHIIH", 0x06054B50, 0, 0, len(central), len(central), cd_size, cd_start, 0))
        os.replace(tmp, dest)


def write_node_map(handle, dest) -> int:
    """`<index>\\t<name>\\n` for every node of the handle's last build, bytes made on the GPU; returns the byte count."""
    buf = handle.fetch_nodes_tsv()
    with open(dest, "wb") as fh:
        fh.write(memoryview(buf))
    return int(buf.size)
