"""Multi-GPU GFA -> CSR build: one process per GPU, `torch.distributed` (NCCL) for the plumbing,
libg2n.so for every device phase (include/g2n.h "multi-GPU phases").

    rank r holds its own newline-aligned byte range of the text
    1. g2n_dist_scan      tokenize the shard (local table, local first-appearance order)
    2. all_gather         distinct keys (32 B each) + per-tile record prefix of every rank
       g2n_dist_merge     same global dictionary on every rank -> global node IDs (the reference's
                          numbering over the concatenated shards, builders.py:194-198, 219-221)
    3. g2n_dist_entries   row entries bucketed by owner(row) = row // rows_per_rank
       all_to_all_single  entries travel to the rank that owns their row block
       g2n_dist_slab      duplicate sum / max(S, S^T) -> this rank's CSR slab (builders.py:279-283)

The exchange steps are real NCCL collectives over NVLink; phase 1 has no collective.
`LocalRank` holds the device work of one rank, `DistBuilder` adds the collectives; tests drive several
`LocalRank`s on one GPU with the exchanges done by tensor slicing (logical shards).
Restrictions of this version: unweighted builds, node names <= 15 bytes, <= 8 ranks."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np
import scipy.sparse as sp

from . import _capi

PAIR_WORDS = 1  # int64 words per exchanged row entry {entry u32 = col << 1 | dir, row u32}
KEY_WORDS = 4   # int64 words per exchanged key {k0, k1, order, pad}
META_WORDS = 8  # int64 words of per-rank counts at the head of an exchanged block
MAX_WORLD = 8


def block_layout(key_stride: int, tile_stride: int) -> tuple[int, int, int]:
    """One all-gathered block per rank: [meta | keys | tile prefix]; returns (block words, key offset, tile offset)."""
    key_off = META_WORDS
    tile_off = key_off + key_stride * KEY_WORDS
    return tile_off + tile_stride, key_off, tile_off


# ------------------------------------------------------------------ host logic (pure, CPU-testable)
def shard_range(nbytes: int, rank: int, world: int, find_newline) -> tuple[int, int]:
    """Byte range of `rank`: [rank*N/world, (rank+1)*N/world) with both cuts moved forward to the byte
    after the next '\\n' (`find_newline(pos)` = index of the first newline at or after pos, or -1).
    Every line belongs to exactly one rank; file order = (rank, offset)."""
    def cut(i: int) -> int:
        if i <= 0:
            return 0
        if i >= world:
            return nbytes
        pos = (nbytes * i) // world
        if pos == 0:
            return 0
        j = find_newline(pos - 1)  # a cut that already follows a newline stays put
        return nbytes if j < 0 else j + 1
    return cut(rank), cut(rank + 1)


def exclusive_prefix(values: list[int]) -> list[int]:
    out, run = [], 0
    for v in values:
        out.append(run)
        run += v
    return out


def rows_per_rank(n_global: int, world: int) -> int:
    return max(1, -(-n_global // world))


def slab_bounds(n_global: int, rank: int, world: int) -> tuple[int, int]:
    rpr = rows_per_rank(n_global, world)
    row0 = min(n_global, rank * rpr)
    return row0, min(n_global, row0 + rpr) - row0


def assemble_slabs(slabs, n_global: int, fmt: str = "csr"):
    """Concatenate per-rank (indptr, indices, data) slabs (rank order = row order) into one matrix."""
    parts, base = [np.zeros(1, np.int64)], 0
    for ip, _, _ in slabs:
        parts.append(np.asarray(ip[1:], dtype=np.int64) + base)
        base += int(ip[-1])
    indptr = np.concatenate(parts)
    assert len(indptr) == n_global + 1, (len(indptr), n_global)
    indices = np.concatenate([np.asarray(s[1]) for s in slabs])
    data = np.concatenate([np.asarray(s[2]) for s in slabs])
    cls = sp.csr_matrix if fmt == "csr" else sp.csc_matrix
    A = cls((n_global, n_global), dtype=data.dtype)
    A.data, A.indices, A.indptr = data, indices.astype(np.int32), indptr.astype(np.int32)
    return A


# ------------------------------------------------------------------ one rank's device work
class LocalRank:
    def __init__(self, device_index: int, rank: int, world: int, stream_ptr: int | None = None):
        import torch

        if not 1 <= world <= MAX_WORLD:
            raise ValueError(f"world size {world} not in 1..{MAX_WORLD}")
        self.torch = torch
        self.rank, self.world = rank, world
        self.dev = torch.device("cuda", device_index)
        self.h = _capi.Handle(device_index)
        self.h.set_stream(stream_ptr if stream_ptr is not None else torch.cuda.current_stream(self.dev).cuda_stream)

    def scan(self, text_dev, *, directed=True, bidirected=False, keep_directed_bidir=False, asymmetric=False,
             strip_orientation=False, dtype="float64", matrix_format="csr") -> list[int]:
        want = {"csr": _capi.FMT_CSR, "csc": _capi.FMT_CSC}[matrix_format]
        self.params = _capi.Params(int(directed), int(bidirected), int(keep_directed_bidir), int(asymmetric), int(strip_orientation),
                                   _capi.DTYPES[np.dtype(dtype).name], want, 1, None, 0, 0)
        self.text = text_dev
        nbytes = int(text_dev.numel())
        info = _capi.DistInfo()
        rc = self.h.lib.g2n_dist_scan(self.h.h, C.c_void_p(text_dev.data_ptr() if nbytes else 0), nbytes, C.byref(self.params), C.byref(info))
        self.scan_rc = rc
        self.h.check(rc)
        self.info = info
        return [int(info.n_keys), int(info.n_tiles), int(info.n_records), int(info.n_edge_records), int(info.n_entries)]

    def export_block(self, mine: list[int], key_stride: int, tile_stride: int):
        """This rank's block [meta | keys | tile prefix] for the all-gather.  If the shard holds more keys or
        tiles than the strides allow, only the meta words are valid (every rank sees that and re-plans)."""
        t = self.torch
        words, key_off, tile_off = block_layout(key_stride, tile_stride)
        block = t.empty(words, dtype=t.int64, device=self.dev)
        meta = list(mine) + [key_stride, tile_stride] + [0] * (META_WORDS - len(mine) - 2)
        block[:META_WORDS] = t.tensor(meta, dtype=t.int64)  # one small H2D copy, ordered on the stream
        if mine[0] <= key_stride and mine[1] + 1 <= tile_stride:
            self.h.check(self.h.lib.g2n_dist_export(self.h.h, C.c_void_p(block.data_ptr() + 8 * key_off), C.c_void_p(block.data_ptr() + 8 * tile_off)))
        return block

    def merge(self, blocks_all, key_stride, tile_stride, meta) -> int:
        """blocks_all: the `world` blocks of export_block(), rank order, one buffer."""
        W = self.world
        words, key_off, tile_off = block_layout(key_stride, tile_stride)
        u64a = C.c_uint64 * MAX_WORLD
        n_keys = [m[0] for m in meta] + [0] * (MAX_WORLD - W)
        rec_base = exclusive_prefix([m[2] for m in meta]) + [0] * (MAX_WORLD - W)
        n_global = C.c_uint64()
        base = blocks_all.data_ptr()
        self.h.check(self.h.lib.g2n_dist_merge(self.h.h, C.c_void_p(base + 8 * key_off), 8 * words, u64a(*n_keys),
                                               C.c_void_p(base + 8 * tile_off), 8 * words, u64a(*rec_base), sum(m[2] for m in meta), W,
                                               C.byref(n_global)))
        self.n_global = int(n_global.value)
        return self.n_global

    def entries(self, meta):
        t = self.torch
        edge_base = exclusive_prefix([m[3] for m in meta])[self.rank]
        n_ent = int(self.info.n_entries)
        send = t.empty(max(1, n_ent) * PAIR_WORDS, dtype=t.int64, device=self.dev)
        dest = (C.c_uint64 * MAX_WORLD)()
        self.h.check(self.h.lib.g2n_dist_entries(self.h.h, self.world, rows_per_rank(self.n_global, self.world), edge_base,
                                                 C.c_void_p(send.data_ptr()), n_ent, dest))
        return send, [int(dest[d]) for d in range(self.world)]

    def slab(self, recv, n_recv: int):
        row0, n_rows = slab_bounds(self.n_global, self.rank, self.world)
        self.h.check(self.h.lib.g2n_dist_slab(self.h.h, C.c_void_p(recv.data_ptr()), n_recv, row0, n_rows))
        self._recv = recv
        return row0, n_rows

    def fetch_slab(self):
        _, indptr, indices, data = self.h.fetch_matrix()
        return indptr, indices, data

    def node_list(self, raw_bytes_id: bool = False):
        from .builders import _node_list

        return _node_list(self.h, raw_bytes_id)


@dataclass
class DistResult:
    n_global: int
    row0: int
    n_rows: int
    nnz_local: int
    info: dict


# ------------------------------------------------------------------ collectives
class DistBuilder:
    """Per-rank driver: LocalRank + NCCL collectives (`torch.distributed`, default or given group)."""

    def __init__(self, device_index: int, group=None):
        import torch
        import torch.distributed as dist

        self.torch, self.dist, self.group = torch, dist, group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.local = LocalRank(device_index, self.rank, self.world)
        self.dev = self.local.dev
        self.strides = None  # (key_stride, tile_stride) of the exchanged blocks, remembered between builds

    def build(self, text_dev, **mode) -> DistResult:
        torch, dist, W, L = self.torch, self.dist, self.world, self.local
        import os
        import time
        dbg = os.environ.get("G2N_DIST_DEBUG")
        marks = []

        def mark(name):
            if dbg:
                torch.cuda.synchronize()
                marks.append((name, time.perf_counter()))

        mark("start")
        mine = L.scan(text_dev, **mode)
        mark("scan")
        # ---- phase 2: dictionary merge.  ONE all-gather of [meta | keys | tile prefix] blocks; the block
        # strides are those of the previous build of this builder (+ slack) -- if any rank's shard does not
        # fit, every rank sees it in the gathered meta words and the exchange is repeated with exact strides.
        meta = None
        for attempt in range(2):
            if self.strides is None:
                if W > 1:
                    t = torch.tensor(mine, dtype=torch.int64, device=self.dev)
                    out = torch.empty(W * len(mine), dtype=torch.int64, device=self.dev)
                    dist.all_gather_into_tensor(out, t, group=self.group)
                    meta = out.view(W, len(mine)).tolist()
                else:
                    meta = [mine]
                ks, ts = max(max(m[0] for m in meta), 1), max(m[1] for m in meta) + 1
                self.strides = (ks + ks // 16 + 64, ts + ts // 16 + 8)
            key_stride, tile_stride = self.strides
            block = L.export_block(mine, key_stride, tile_stride)
            mark("export")
            if W > 1:
                blocks_all = torch.empty(W * block.numel(), dtype=torch.int64, device=self.dev)
                dist.all_gather_into_tensor(blocks_all, block, group=self.group)
            else:
                blocks_all = block
            got = blocks_all.view(W, -1)[:, :META_WORDS].tolist()
            meta = [g[:len(mine)] for g in got]
            if all(m[0] <= key_stride and m[1] + 1 <= tile_stride for m in meta):
                break
            self.strides = None  # some shard outgrew the remembered strides: plan again (same decision on every rank)
        mark("allgather")
        ng = L.merge(blocks_all, key_stride, tile_stride, meta)
        mark("merge")
        # ---- phase 3: edge exchange by owner row block
        send, send_counts = L.entries(meta)
        mark("entries")
        if W > 1:
            sc = torch.tensor(send_counts, dtype=torch.int64, device=self.dev)
            rcnt = torch.empty(W, dtype=torch.int64, device=self.dev)
            dist.all_to_all_single(rcnt, sc, group=self.group)
            recv_counts = rcnt.tolist()
            n_recv = sum(recv_counts)
            recv = torch.empty(max(1, n_recv) * PAIR_WORDS, dtype=torch.int64, device=self.dev)
            dist.all_to_all_single(recv[: n_recv * PAIR_WORDS], send[: sum(send_counts) * PAIR_WORDS],
                                   [c * PAIR_WORDS for c in recv_counts], [c * PAIR_WORDS for c in send_counts], group=self.group)
        else:
            recv, n_recv = send, send_counts[0]
        mark("alltoall")
        row0, n_rows = L.slab(recv, n_recv)
        mark("slab")
        if dbg and self.rank == 0:
            import sys
            print("dist phases (ms): " + ", ".join(f"{b[0]} {1e3 * (b[1] - a[1]):.3f}" for a, b in zip(marks, marks[1:])), file=sys.stderr)
        s = L.h.sizes()
        return DistResult(ng, row0, n_rows, int(s.nnz), dict(meta=meta, send_counts=send_counts, n_recv=n_recv,
                                                            key_bytes=W * key_stride * KEY_WORDS * 8,
                                                            pair_bytes=sum(send_counts) * PAIR_WORDS * 8))

    def fetch_slab(self):
        return self.local.fetch_slab()

    def gather_matrix(self, result: DistResult, fmt: str = "csr"):
        """Assemble the full matrix on every rank (parity tests; production keeps the slabs)."""
        slab = tuple(np.array(a) for a in self.fetch_slab())
        if self.world == 1:
            return assemble_slabs([slab], result.n_global, fmt)
        objs = [None] * self.world
        self.dist.all_gather_object(objs, slab, group=self.group)
        return assemble_slabs(objs, result.n_global, fmt)

    def node_list(self, raw_bytes_id: bool = False):
        return self.local.node_list(raw_bytes_id)
