#!/usr/bin/env python
"""The small fixtures under compute-sanitizer:

    compute-sanitizer --tool memcheck  python tools/sanitize_run.py
    compute-sanitizer --tool racecheck python tools/sanitize_run.py

* the reference's DRB1 fixture (tests/golden/DRB1-3123_unsorted.gfa) in five modes incl. weighted / bidirected,
* a fuzz text with deferred lines, long keys, errors and unknown records (tests/golden_inputs.fuzz_text(11)),
* a 3-logical-shard multi-GPU build on one GPU (peer flags, owner table, entry exchange) incl. a speculative repeat,
* the partitioned row build (rowsort.cuh: RowBuckets / SubBuckets) forced on the fixture, single-GPU and slab,
* a *.gz source, the edge-list export and a path-distance matrix.
Every result is compared with the CPU oracle, so a run that passes did the real work."""
import gzip
import os
import sys
import tempfile
import warnings
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
warnings.simplefilter("ignore")

import torch  # noqa: E402

import golden_inputs as gi  # noqa: E402
from gfa2network_b200 import _capi, convert_format, parse_gfa  # noqa: E402
from gfa2network_b200 import dist as D  # noqa: E402
from gfa2network_b200.export import edge_list_bytes  # noqa: E402
from oracle.oracle import oracle_convert_format, oracle_edge_list, oracle_parse_gfa  # noqa: E402


def same(A, B, what):
    assert A.format == B.format and A.shape == B.shape and A.dtype == B.dtype, what
    if A.format == "coo":
        a, b = (A.row, A.col, A.data), (B.row, B.col, B.data)
    else:
        a, b = (A.indptr, A.indices, A.data), (B.indptr, B.indices, B.data)
    for x, y in zip(a, b):
        assert np.array_equal(x, y, equal_nan=True), what


drb1 = (ROOT / "tests" / "golden" / "DRB1-3123_unsorted.gfa").read_bytes()
n = 0
for mode in (dict(), dict(directed=False), dict(bidirected=True), dict(asymmetric=True, dtype="float32"), dict(bidirected=True, keep_directed_bidir=True, weight_tag="RC")):
    A, nodes = parse_gfa(drb1, build_graph=False, build_matrix=True, return_node_list=True, **mode)
    B, onodes = oracle_parse_gfa(drb1, return_node_list=True, **mode)
    same(A, B, mode)
    assert nodes == onodes
    same(convert_format(A, "csr", _untouched=True), oracle_convert_format(B, "csr"), mode)
    n += 1
fz = gi.fuzz_text(11, 1500)
for mode in (dict(), dict(weight_tag="RC", directed=False), dict(bidirected=True)):
    try:
        A = parse_gfa(fz, build_graph=False, build_matrix=True, **mode)
        B = oracle_parse_gfa(fz, **mode)
        same(A, B, mode)
    except Exception as e:  # noqa: BLE001 - the same exception on both sides
        try:
            oracle_parse_gfa(fz, **mode)
            raise AssertionError(f"only the device path raised: {e!r}")
        except type(e):
            pass
    n += 1
# row arrays "far larger than L2" (forced): both bucket levels, the overflow path of a sub-bucket, one level only
for variant in (dict(), dict(G2N_DBG_SUBCAP="8"), dict(G2N_DBG_NOSUB="1")):
    os.environ["G2N_DBG_ROWPASS"] = "5"
    os.environ.update(variant)
    for mode in (dict(), dict(bidirected=True, weight_tag="RC"), dict(directed=False)):
        A = parse_gfa(drb1, build_graph=False, build_matrix=True, **mode)
        same(A, oracle_parse_gfa(drb1, **mode), (variant, mode))
        n += 1
    for k in ("G2N_DBG_ROWPASS", *variant):
        del os.environ[k]
# gz source, edge list
with tempfile.TemporaryDirectory() as d:
    f = os.path.join(d, "x.gfa.gz")
    open(f, "wb").write(gzip.compress(drb1))
    same(parse_gfa(f, build_graph=False, build_matrix=True), oracle_parse_gfa(drb1), "gz")
el, exc, _ = edge_list_bytes(drb1)
assert exc is None and el.tobytes() == oracle_edge_list(drb1)[0]
n += 2
# three logical shards on one GPU: host-planned, then speculative
G = 3
text = np.frombuffer(drb1, dtype=np.uint8)
shards = []
for r in range(G):
    lo, hi = D.shard_range(len(drb1), r, G, lambda p: drb1.find(b"\n", p))
    shards.append(torch.from_numpy(text[lo:hi].copy()).cuda())
ranks = [D.LocalRank(0, r, G) for r in range(G)]
caps = None
for spec in (False, True, "buckets"):
    if spec == "buckets":  # the slab's partitioned receive (forced), speculative
        os.environ["G2N_DBG_ROWPASS"] = "5"
        spec = True
    for r, t in zip(ranks, shards):
        r.set_input(t)
    if not spec:
        infos = [r.probe() for r in ranks]
        D.raise_agreed(infos)
        caps = D.plan_caps(infos, G)
        for r in ranks:
            r.plan(*caps)
        mems = [r.local_mem() for r in ranks]
        for r in ranks:
            r.set_peers([m[0] for m in mems], [m[1] for m in mems])
    for k in range(D.N_STAGES):
        for r in ranks:
            r.stage(k, spec)
    out = [r.finish() for r in ranks]
    assert all(rc == _capi.G2N_OK for rc, _ in out)
    slabs = [tuple(np.array(a) for a in r.fetch_slab()) for r in ranks]
    A = D.assemble_slabs(slabs, int(out[0][1].n_global), "csr")
    same(A, oracle_convert_format(oracle_parse_gfa(drb1), "csr"), f"dist spec={spec}")
    for r, (_, res) in zip(ranks, out):
        r.remember(res, *caps)
    n += 1
print(f"sanitize_run: {n} device builds checked against the oracle")
