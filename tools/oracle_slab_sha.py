#!/usr/bin/env python
"""CPU-oracle side of the bit-exact check of a multi-GPU build at full size.

    python tools/oracle_slab_sha.py <dist_check.json line file> [--out profiles/...json]

Reads the JSON line `tools/dist_check.py --sha` printed (config, per-GPU scale, world size and, per rank, the row
block, the name range and the sha256 of its slab / names), regenerates the SAME shards on the host (bench.make_text is
deterministic), runs the CPU oracle (oracle/: C restatement of parser.py / builders.py + SciPy, pinned against the
real reference) over the concatenated text, cuts the oracle's CSR into the same row blocks and compares the sha256 of
every block and of every name range.  Needs no GPU; memory ~ 3x the text (C5 in full: ~40 GB text, ~120 GB peak)."""
from __future__ import annotations

import argparse
import hashlib
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))


def main():
    import bench
    from oracle.oracle import oracle_convert_format, oracle_parse_gfa

    ap = argparse.ArgumentParser()
    ap.add_argument("line_file")
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    d = next(json.loads(ln) for ln in open(args.line_file) if ln.strip().startswith("{") and '"slabs"' in ln)
    world, cfg_name, scale = d["n_gpus"], d["config"], d["scale_per_gpu"]
    t0 = time.perf_counter()
    parts = []
    for r in range(world):
        cfg, text, _, _ = bench.make_text(cfg_name, scale, rank=r, world=world)
        assert text.size == d["slabs"][r]["text_bytes"], (r, text.size, d["slabs"][r]["text_bytes"])
        parts.append(text)
    full = np.concatenate(parts)
    del parts
    gen_s = time.perf_counter() - t0
    mode = dict(cfg["mode"])
    t0 = time.perf_counter()
    A, names = oracle_parse_gfa(full, return_node_list=True, raw_bytes_id=True, **mode)
    parse_s = time.perf_counter() - t0
    del full
    t0 = time.perf_counter()
    A = oracle_convert_format(A, "csr")
    conv_s = time.perf_counter() - t0
    assert A.shape[0] == d["nodes"], (A.shape, d["nodes"])
    out = []
    ok = True
    for s in d["slabs"]:
        row0, n_rows = s["row0"], s["n_rows"]
        lo, hi = int(A.indptr[row0]), int(A.indptr[row0 + n_rows])
        h = hashlib.sha256()
        h.update((A.indptr[row0:row0 + n_rows + 1] - lo).astype(np.int32).tobytes())
        h.update(np.ascontiguousarray(A.indices[lo:hi]).astype(np.int32, copy=False).tobytes())
        h.update(np.ascontiguousarray(A.data[lo:hi]).tobytes())
        hn = hashlib.sha256(b"\n".join(names[s["id0"]:s["id0"] + s["n_first"]])).hexdigest()
        same = h.hexdigest() == s["sha_slab"] and hn == s["sha_names"] and hi - lo == s["nnz"]
        ok = ok and same
        out.append(dict(rank=s["rank"], row0=row0, n_rows=n_rows, nnz_oracle=hi - lo, nnz_gpu=s["nnz"], slab_equal=h.hexdigest() == s["sha_slab"],
                        names_equal=hn == s["sha_names"], sha_slab_oracle=h.hexdigest(), sha_names_oracle=hn))
    res = dict(tool="oracle_slab_sha", config=cfg_name, scale_per_gpu=scale, n_gpus=world, nodes=int(A.shape[0]), nnz=int(A.nnz), text_bytes=d["text_bytes"],
               bit_exact=bool(ok), peak_rss_gb=round(__import__("resource").getrusage(__import__("resource").RUSAGE_SELF).ru_maxrss / 1e6, 1), generate_s=round(gen_s, 1), oracle_parse_s=round(parse_s, 1), oracle_convert_s=round(conv_s, 1), slabs=out)
    print(json.dumps(res))
    if args.out:
        Path(args.out).write_text(json.dumps(res) + "\n")
    if not ok:
        raise SystemExit("MISMATCH")


if __name__ == "__main__":
    main()
