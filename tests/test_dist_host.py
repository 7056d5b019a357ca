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


def _info(rc=0, **kw):
    d = dict(rc=rc, msg="", n_keys=10, n_tiles=1, n_records=20, n_edge_records=10, n_entries=20, err_kind=0, err_offset=0,
             unknown_byte=-1, unknown_offset=0, nbytes=100)
    d.update(kw)
    return d


def test_agreed_exceptions_follow_file_order():
    """Every rank raises what parse_gfa raises for the concatenated shards: the lowest rank's error,
    the unsupported-record warning only if it does not come after that error."""
    import warnings

    from gfa2network_b200 import _capi

    with warnings.catch_warnings():
        warnings.simplefilter("error")
        D.raise_agreed([_info(), _info()])  # nothing to report
    with pytest.raises(ValueError, match="Malformed E record"):
        D.raise_agreed([_info(), _info(rc=_capi.G2N_ERR_PARSE, err_kind=2), _info(rc=_capi.G2N_ERR_PARSE, err_kind=1)])
    with pytest.warns(RuntimeWarning, match="Skipping unsupported record: #"):
        D.raise_agreed([_info(unknown_byte=ord("#")), _info(unknown_byte=ord("W"))])  # the first one in file order
    with pytest.warns(RuntimeWarning, match="Skipping unsupported record: W"):
        with pytest.raises(IndexError):
            D.raise_agreed([_info(unknown_byte=ord("W")), _info(rc=_capi.G2N_ERR_PARSE, err_kind=6)])
    with warnings.catch_warnings():
        warnings.simplefilter("error")  # the unsupported record sits in a LATER shard than the error: never reported
        with pytest.raises(ValueError, match="Malformed P record"):
            D.raise_agreed([_info(rc=_capi.G2N_ERR_PARSE, err_kind=4), _info(unknown_byte=ord("W"))])
    with pytest.raises(NotImplementedError, match="hash collision"):
        D.raise_agreed([_info(), _info(rc=_capi.G2N_ERR_UNSUPPORTED, msg="multi-GPU build: hash collision among long node names")])
    with pytest.raises(_capi.G2NError):
        D.raise_agreed([_info(rc=_capi.G2N_ERR_CUDA, msg="out of memory")])


def test_plan_caps_grow_with_retries_and_cover_one_sided_shards():
    infos = [_info(n_keys=1_000_000, n_entries=6_000_000), _info(n_keys=1_300_000, n_entries=100)]
    for world in (2, 4, 8):
        k0, p0 = D.plan_caps(infos, world, 0)
        k1, p1 = D.plan_caps(infos, world, 1)
        assert k0 >= 1_300_000 / world * 1.2 and k1 > k0  # hash partition + slack, more slack after a miss
        assert p0 >= 6_000_000 and p1 == p0               # a shard may send every entry to ONE owner
    assert D.plan_caps([_info(n_keys=0, n_entries=0)], 1)[0] > 0


def test_dist_result_reports_what_moved():
    import ctypes

    from gfa2network_b200 import _capi

    res = _capi.DistResult()
    res.n_global, res.row0, res.n_rows, res.nnz, res.n_recv, res.n_first, res.id0 = 100, 25, 25, 60, 70, 30, 20
    for d in range(4):
        res.keys_to[d], res.pairs_to[d] = 10 + d, 20 + d
    r = D._result(res, 4, True)
    assert (r.n_global, r.row0, r.n_rows, r.nnz_local) == (100, 25, 25, 60)
    assert r.info["keys_to"] == [10, 11, 12, 13] and r.info["pairs_to"] == [20, 21, 22, 23] and r.info["speculative"] is True
    assert ctypes.sizeof(_capi.PathInfo) == 48


def test_file_cuts_partition_the_file_at_newlines(tmp_path):
    """dist.file_cuts: the byte ranges parse_gfa(path, devices=[...]) hands to the GPUs."""
    import random

    from gfa2network_b200.dist import file_cuts

    r = random.Random(4)
    for trial in range(30):
        lines = [bytes(r.choice(b"SLPabc\t+-0123") for _ in range(r.choice([0, 1, 3, 17, 200, 5000]))) + b"\n" for _ in range(r.randrange(0, 40))]
        data = b"".join(lines)
        if trial % 3 == 0 and data:
            data = data[:-1]  # no final newline
        f = tmp_path / f"t{trial}.gfa"
        f.write_bytes(data)
        for world in (1, 2, 3, 8):
            cuts = file_cuts(str(f), world)
            assert cuts[0][0] == 0 and cuts[-1][1] == len(data)
            for (a, b), (c, d) in zip(cuts, cuts[1:]):
                assert b == c and a <= b
            for a, b in cuts:
                assert a == 0 or a == len(data) or data[a - 1:a] == b"\n"
