"""Host-side logic of the multi-GPU path under a real 2-process `gloo` group on CPU: newline-aligned
sharding, prefix bookkeeping, owner row blocks and slab assembly (device phases are emulated with
SciPy here -- the oracle side of the house; the CUDA phases are covered by tests/test_gpu_dist.py)."""
import os
import socket

import numpy as np
import pytest
import scipy.sparse as sp
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gfa2network_b200 import dist as D


def test_shard_range_covers_text_on_line_boundaries():
    import golden_inputs as gi

    text = gi.fuzz_text(11, 3000)
    for world in (1, 2, 3, 5, 8):
        cuts = [D.shard_range(len(text), r, world, lambda p: text.find(b"\n", p)) for r in range(world)]
        assert cuts[0][0] == 0 and cuts[-1][1] == len(text)
        for (a, b), (c, d) in zip(cuts, cuts[1:]):
            assert b == c
        for a, b in cuts:
            assert a == 0 or text[a - 1:a] == b"\n"
        assert b"".join(text[a:b] for a, b in cuts) == text


def test_slab_bounds_partition_rows():
    for n in (0, 1, 7, 8, 9, 1000):
        for w in (1, 2, 3, 8):
            b = [D.slab_bounds(n, r, w) for r in range(w)]
            assert sum(x[1] for x in b) == n
            assert all(b[i][0] + b[i][1] == b[i + 1][0] or b[i + 1][1] == 0 for i in range(w - 1))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(3)
        n = 257
        r = rng.integers(0, n, 4000)
        c = rng.integers(0, n, 4000)
        full = sp.coo_matrix((np.ones(4000), (r, c)), shape=(n, n)).tocsr()
        full.sum_duplicates()
        # every rank "owns" a row block and sends each entry to its owner with all_to_all
        rpr = D.rows_per_rank(n, world)
        mine = np.arange(rank, 4000, world)  # the entries this rank produced
        pairs = np.stack([c[mine], r[mine]], 1).astype(np.int64)
        dest = pairs[:, 1] // rpr
        order = np.argsort(dest, kind="stable")
        pairs = pairs[order]
        send_counts = [int((dest == d).sum()) for d in range(world)]
        sc = torch.tensor(send_counts)
        rc = torch.empty(world, dtype=torch.int64)
        dist.all_to_all_single(rc, sc)
        recv = torch.empty(int(rc.sum()) * 2, dtype=torch.int64)
        dist.all_to_all_single(recv, torch.from_numpy(pairs.reshape(-1).copy()), [int(x) * 2 for x in rc], [x * 2 for x in send_counts])
        got = recv.view(-1, 2).numpy()
        row0, n_rows = D.slab_bounds(n, rank, world)
        slab = sp.coo_matrix((np.ones(len(got)), (got[:, 1] - row0, got[:, 0])), shape=(n_rows, n)).tocsr()
        slab.sum_duplicates()
        objs = [None] * world
        dist.all_gather_object(objs, (slab.indptr, slab.indices, slab.data))
        A = D.assemble_slabs(objs, n)
        ok = np.array_equal(A.indptr, full.indptr) and np.array_equal(A.indices, full.indices) and np.array_equal(A.data, full.data)
        meta = [None] * world
        dist.all_gather_object(meta, len(mine))
        ok = ok and D.exclusive_prefix(meta)[rank] == sum(len(np.arange(k, 4000, world)) for k in range(rank))
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_two_rank_gloo_exchange_and_assembly():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=100) for _ in procs)
    for p in procs:
        p.join(30)
    assert res == [(0, True), (1, True)]
