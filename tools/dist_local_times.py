#!/usr/bin/env python
"""Per-kernel times of a multi-rank build whose ranks all live on ONE GPU (logical shards: the peers' exchange arenas are plain
device pointers, nothing crosses NVLink).  Compared with the per-kernel lines of `bench.py --gpus N` this separates what the
dictionary / entry exchange costs as COMPUTE from what it costs as peer traffic:

    python tools/dist_local_times.py 2        # two C2 shards of one graph on cuda:0
    python tools/dist_local_times.py 8 0.25   # eight shards at a quarter of the C2 size each

(measurement scaffolding; bench.py is the benchmark)"""
import json
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch  # noqa: E402

from bench import make_text  # noqa: E402
from gfa2network_b200 import _capi  # noqa: E402
from gfa2network_b200 import dist as D  # noqa: E402

G = int(sys.argv[1]) if len(sys.argv) > 1 else 2
scale = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
steps = 10
mode = dict(directed=False)
shards = [torch.from_numpy(make_text("C2", scale, rank=r, world=G)[1]).cuda() for r in range(G)]
ranks = [D.LocalRank(0, r, G) for r in range(G)]


def build(spec, caps=None):
    for r, t in zip(ranks, shards):
        r.set_input(t, **mode)
    if not spec:
        infos = [r.probe() for r in ranks]
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
    assert all(rc == _capi.G2N_OK for rc, _ in out), [rc for rc, _ in out]
    for r, (_, res) in zip(ranks, out):
        r.remember(res, *caps)
    return caps


caps = build(False)
for _ in range(3):
    build(True, caps)
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
for r in ranks:
    r.h.set_profile(True)
tot = {}
for _ in range(steps):
    flush.zero_()
    build(True, caps)
    torch.cuda.synchronize()
    for k, (m, c) in ranks[0].h.kernel_times().items():
        tot[k] = tot.get(k, 0.0) + m / steps
print(json.dumps({"tool": "dist_local_times", "logical_ranks": G, "scale": scale, "text_mb_per_rank": round(shards[0].numel() / 1e6, 1),
                  "rank0_kernels_ms": {k: round(v, 4) for k, v in tot.items()}, "sum_ms": round(sum(tot.values()), 4)}))
