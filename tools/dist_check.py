#!/usr/bin/env python
"""Multi-GPU build of a named configuration at scale, with checks that do not need the CPU oracle
(run under torchrun, one rank per GPU):

    python -m torch.distributed.run --nproc-per-node 8 tools/dist_check.py --config C5 --scale 0.125

* every rank generates its shard (bench.make_text: `--scale` is per GPU, the shards form ONE graph), builds
  it once host-planned and `--steps` times speculatively (CUDA events, max over ranks);
* size-independent properties of the result, checked on the device: slab indptr monotone and consistent
  with nnz, column indices strictly increasing inside every row, row blocks and name ranges partition
  the node set, entries sent == entries received, node count == segments generated, and -- for the
  max(S, S^T) modes -- symmetry through an order-independent checksum of (row, col) against (col, row);
* `--verify-single`: the shards are gathered on rank 0, built there by the single-GPU path (which the
  parity suite pins against the oracle), and every rank compares its slab and its node names bit for bit.
Prints one JSON line on rank 0."""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))


class _DevArr:
    def __init__(self, ptr: int, n: int, typestr: str):
        self.__cuda_array_interface__ = dict(shape=(int(n),), typestr=typestr, data=(int(ptr), False), version=2)


def device_result(h, torch, dev):
    """(indptr, indices, data) of the handle's resident CSR result as torch views (no copy)."""
    s = h.sizes()
    a0, a1, d = C.c_void_p(), C.c_void_p(), C.c_void_p()
    h.check(h.lib.g2n_device_result(h.h, C.byref(a0), C.byref(a1), C.byref(d)))
    rows, nnz = int(s.slab_rows), int(s.nnz)
    ip = torch.as_tensor(_DevArr(a0.value, rows + 1, "<i4"), device=dev)
    ix = torch.as_tensor(_DevArr(a1.value, max(nnz, 1), "<i4"), device=dev)[:nnz]
    dt = torch.as_tensor(_DevArr(d.value, max(nnz, 1), "<f8"), device=dev)[:nnz]
    return ip, ix, dt


def main():
    import torch
    import torch.distributed as dist

    import bench
    from gfa2network_b200 import _capi
    from gfa2network_b200.dist import DistBuilder

    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="C5")
    ap.add_argument("--scale", type=float, default=0.01)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--verify-single", action="store_true")
    ap.add_argument("--sha", action="store_true", help="sha256 of every rank's slab (indptr | indices | data) and of its node names: compared "
                    "offline with tools/oracle_slab_sha.py (the CPU oracle over the concatenated shards)")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def allsum(x: int) -> int:
        t = torch.tensor([x], dtype=torch.int64, device=dev)
        if world > 1:
            dist.all_reduce(t)
        return int(t.item())

    def allmax(x: float) -> float:
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    t0 = time.perf_counter()
    cfg, text_np, n_seg, n_link = bench.make_text(args.config, args.scale, rank=rank, world=world)
    nbytes = int(text_np.size)
    text_dev = torch.from_numpy(text_np).to(dev)
    del text_np
    gen_s = time.perf_counter() - t0
    mode = dict(cfg["mode"])
    sym = bool(mode.get("keep_directed_bidir") or (not mode.get("bidirected") and mode.get("directed", True))) and not mode.get("asymmetric")
    b = DistBuilder(local)
    stream = torch.cuda.current_stream()
    t0 = time.perf_counter()
    res = b.build(text_dev, matrix_format="csr", **mode)
    torch.cuda.synchronize()
    first_s = time.perf_counter() - t0
    for _ in range(2):  # the first speculative build still grows a few buffers (slack over the exact sizes)
        res = b.build(text_dev, matrix_format="csr", **mode)
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    spec = []
    for a, e in ev:
        if world > 1:
            dist.barrier()
        a.record(stream)
        res = b.build(text_dev, matrix_format="csr", **mode)
        e.record(stream)
        spec.append(bool(res.info["speculative"]))
    torch.cuda.synchronize()
    ms = allmax(sum(a.elapsed_time(e) for a, e in ev) / max(1, args.steps)) if args.steps else float("nan")

    # ---- per-kernel times of this rank (event pair around every launch; not part of the timed builds above)
    h = b.local.h
    kern = {}
    h.set_profile(True)
    for _ in range(2):
        if world > 1:
            dist.barrier()
        res = b.build(text_dev, matrix_format="csr", **mode)
        torch.cuda.synchronize()
        for k, (m, c) in h.kernel_times().items():
            kern[k] = round(kern.get(k, 0.0) + m / 2, 4)
    h.set_profile(False)
    res = b.build(text_dev, matrix_format="csr", **mode)
    torch.cuda.synchronize()

    # ---- properties of the resident result
    ip, ix, dt = device_result(h, torch, dev)
    nnz = int(res.nnz_local)
    checks = {}
    ok_local = bool(ip[0].item() == 0 and ip[-1].item() == nnz and bool((ip[1:] >= ip[:-1]).all().item()))
    if nnz > 1:
        inc = ix[1:] > ix[:-1]
        starts = ip[1:-1].long()
        starts = starts[(starts > 0) & (starts < nnz)]
        inc[starts - 1] = True  # the first entry of a row may be smaller than the last one of the row before
        ok_local = ok_local and bool(inc.all().item())
        del inc, starts
    if nnz:
        ok_local = ok_local and bool((ix >= 0).all().item()) and bool((ix < res.n_global).all().item())
    checks["slabs_sorted_and_consistent"] = allsum(int(ok_local)) == world
    checks["rows_partition_nodes"] = allsum(res.n_rows) == res.n_global
    checks["names_partition_nodes"] = allsum(res.info["n_first"]) == res.n_global
    checks["entries_sent_eq_received"] = allsum(sum(res.info["pairs_to"])) == allsum(res.info["n_recv"])
    checks["node_count_eq_segments"] = res.n_global == world * n_seg * (2 if mode.get("bidirected") else 1)  # S registers id:+ and id:-
    total_nnz = allsum(nnz)
    if sym:
        rows = torch.repeat_interleave(torch.arange(res.row0, res.row0 + res.n_rows, device=dev, dtype=torch.int64), (ip[1:] - ip[:-1]).long())
        cols = ix.long()
        A_, B_ = 0x9E3779B97F4A7C15 - (1 << 64), 0x2545F4914F6CDD1D  # odd 64-bit multipliers (two's complement)
        f = lambda r, c: int(((r * A_) ^ (c * B_ + (r << 7))).sum().item())  # noqa: E731 - order-independent over entries
        s_rc, s_cr = f(rows, cols), f(cols, rows)
        tot = lambda s: ((allsum(s >> 31) << 31) + allsum(s & 0x7FFFFFFF)) % (1 << 64)  # noqa: E731 - exact sum over ranks, mod 2^64
        checks["symmetric_checksum"] = tot(s_rc) == tot(s_cr)
        if not mode.get("weight_tag"):
            checks["weights_are_positive_integers"] = allsum(int(bool(((dt >= 1.0) & (dt == dt.floor())).all().item()))) == world
        del rows, cols

    # ---- bit-exact comparison with the single-GPU build of the concatenated shards
    verified = None
    if args.verify_single:
        sizes = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
        if world > 1:
            dist.all_gather(sizes, torch.tensor([nbytes], dtype=torch.int64, device=dev))
        else:
            sizes[0][0] = nbytes
        sizes = [int(s.item()) for s in sizes]
        pad = max(sizes)
        mine = torch.zeros(pad, dtype=torch.uint8, device=dev)
        mine[:nbytes] = text_dev
        parts = [torch.empty(pad, dtype=torch.uint8, device=dev) for _ in range(world)] if rank == 0 else None
        if world > 1:
            dist.gather(mine, parts, dst=0)
        else:
            parts = [mine]
        _, my_names = b.local.node_list()
        if rank == 0:
            full = torch.cat([p[:n] for p, n in zip(parts, sizes)])
            del parts
            hs = _capi.Handle(local)
            hs.set_stream(stream.cuda_stream)
            wtb = mode["weight_tag"].encode() if mode.get("weight_tag") else None
            params = _capi.Params(int(mode.get("directed", True)), int(mode.get("bidirected", False)), int(mode.get("keep_directed_bidir", False)),
                                  int(mode.get("asymmetric", False)), 0, _capi.DTYPES["float64"], _capi.FMT_CSR, 1, wtb, len(wtb) if wtb else 0, 0)
            hs.check(hs.build(full.data_ptr(), int(full.numel()), params))
            sip, six, sdt = device_result(hs, torch, dev)
            from gfa2network_b200.builders import _node_list

            all_names = _node_list(hs, False)
        # every rank tells rank 0 its row block and name range, rank 0 answers with the expected pieces
        meta = [(res.row0, res.n_rows, res.info["id0"], res.info["n_first"])]
        if world > 1:
            meta = [None] * world
            dist.all_gather_object(meta, (res.row0, res.n_rows, res.info["id0"], res.info["n_first"]))
        same = True
        for r in range(world):
            row0, n_rows, id0, n_first = meta[r]
            if rank == 0:
                lo, hi = int(sip[row0].item()), int(sip[row0 + n_rows].item())
                exp = ((sip[row0:row0 + n_rows + 1] - lo).contiguous(), six[lo:hi].contiguous(), sdt[lo:hi].contiguous())
                names = all_names[id0:id0 + n_first]
                if r == 0:
                    got = exp
                    exp_names = names
                else:
                    dist.send(torch.tensor([hi - lo], dtype=torch.int64, device=dev), dst=r)
                    for t in exp:
                        dist.send(t, dst=r)
                    dist.send_object_list([names], dst=r)
            if rank == r:
                if r != 0:
                    n = torch.zeros(1, dtype=torch.int64, device=dev)
                    dist.recv(n, src=0)
                    n = int(n.item())
                    got = (torch.empty(n_rows + 1, dtype=torch.int32, device=dev), torch.empty(n, dtype=torch.int32, device=dev),
                           torch.empty(n, dtype=torch.float64, device=dev))
                    for t in got:
                        dist.recv(t, src=0)
                    box = [None]
                    dist.recv_object_list(box, src=0)
                    exp_names = box[0]
                same = (got[0].numel() == ip.numel() and bool(torch.equal(got[0], ip)) and got[1].numel() == ix.numel() and bool(torch.equal(got[1], ix))
                        and bool(torch.equal(got[2].view(torch.int64), dt.view(torch.int64))) and exp_names == my_names)
        verified = allsum(int(same)) == world

    # ---- fingerprints of this rank's slab and names, for the offline comparison with the CPU oracle
    slabs = None
    if args.sha:
        import hashlib

        hip, hix, hdt = b.fetch_slab()
        hsh = hashlib.sha256()
        for a in (hip, hix, hdt):
            hsh.update(np.ascontiguousarray(a).tobytes())
        _, my_names2 = b.local.node_list(raw_bytes_id=True)
        mine = dict(rank=rank, row0=int(res.row0), n_rows=int(res.n_rows), nnz=int(nnz), id0=int(res.info["id0"]), n_first=int(res.info["n_first"]),
                    sha_slab=hsh.hexdigest(), sha_names=hashlib.sha256(b"\n".join(my_names2)).hexdigest(), text_bytes=nbytes)
        del hip, hix, hdt, my_names2
        slabs = [mine]
        if world > 1:
            slabs = [None] * world
            dist.all_gather_object(slabs, mine)

    total_bytes = allsum(nbytes)
    line = {
        "tool": "dist_check", "config": args.config, "scale_per_gpu": args.scale, "n_gpus": world,
        "segments": world * n_seg, "links": world * n_link, "text_bytes": total_bytes, "nodes": res.n_global, "nnz": total_nnz,
        "ms_per_build": ms, "GBps": total_bytes / (ms * 1e6) if args.steps else None, "edges_per_s": world * n_link / (ms / 1e3) if args.steps else None,
        "hbm_frac_of_aggregate_peak": (total_bytes / (ms * 1e6)) / (world * bench.peaks()[0]) if args.steps else None,
        "speculative_steps": spec, "first_build_s": first_s, "generate_s": gen_s, "checks": checks, "verified_against_single_gpu": verified,
        "mode": mode or "directed (default): CSR of max(S, S^T)",
    }
    line["kernels_rank0_ms"] = kern
    if slabs is not None:
        line["slabs"] = slabs
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    bad = [k for k, v in checks.items() if not v] + ([] if verified in (None, True) else ["verify_single"])
    if bad:
        raise SystemExit(f"rank {rank}: FAILED {bad}")


if __name__ == "__main__":
    main()
