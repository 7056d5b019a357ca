"""Multi-GPU path on ONE GPU: G logical ranks (one LocalRank/handle each) with the NCCL exchanges
replaced by tensor slicing -- the same kernels and the same host logic as `DistBuilder`, compared
bit-for-bit with the single-GPU build and the CPU oracle."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _run_logical(text: np.ndarray, G: int, **mode):
    import torch

    from gfa2network_b200 import dist as D

    dev = torch.device("cuda", 0)
    tb = text.tobytes()
    ranks, shards = [], []
    for r in range(G):
        lo, hi = D.shard_range(len(tb), r, G, lambda p: tb.find(b"\n", p))
        shards.append(torch.from_numpy(text[lo:hi].copy()).to(dev))
        ranks.append(D.LocalRank(0, r, G))
    meta = [ranks[r].scan(shards[r], **mode) for r in range(G)]
    key_stride = max(max(m[0] for m in meta), 1)
    tile_stride = max(m[1] for m in meta) + 1
    blocks_all = torch.cat([ranks[r].export_block(meta[r], key_stride, tile_stride) for r in range(G)])
    ng = {ranks[r].merge(blocks_all, key_stride, tile_stride, meta) for r in range(G)}
    assert len(ng) == 1
    ng = ng.pop()
    sends = [ranks[r].entries(meta) for r in range(G)]
    slabs = []
    for dst in range(G):
        parts = []
        for src in range(G):
            send, counts = sends[src]
            off = sum(counts[:dst]) * D.PAIR_WORDS
            parts.append(send[off: off + counts[dst] * D.PAIR_WORDS])
        recv = torch.cat(parts) if parts else torch.empty(0, dtype=torch.int64, device=dev)
        n_recv = recv.numel() // D.PAIR_WORDS
        if n_recv == 0:
            recv = torch.zeros(2, dtype=torch.int64, device=dev)
        ranks[dst].slab(recv, n_recv)
        slabs.append(tuple(np.array(a) for a in ranks[dst].fetch_slab()))
    A = D.assemble_slabs(slabs, ng, mode.get("matrix_format", "csr"))
    nodes = ranks[0].node_list()
    return A, nodes


MODES = [dict(), dict(directed=False), dict(asymmetric=True), dict(bidirected=True), dict(bidirected=True, keep_directed_bidir=True),
         dict(directed=False, matrix_format="csc"), dict(dtype="bool")]


@pytest.mark.parametrize("G", [1, 2, 3, 8])
@pytest.mark.parametrize("mode", MODES, ids=[str(m) for m in MODES])
def test_logical_shards_match_oracle(G, mode):
    from gfa2network_b200.synth import synth_gfa
    from oracle.oracle import oracle_convert_format, oracle_parse_gfa

    text = synth_gfa(30_000, 90_000, seed=9, kind=1, interleave=2048, n_paths=1)
    A, nodes = _run_logical(text, G, **mode)
    omode = {k: v for k, v in mode.items() if k != "matrix_format"}
    fmt = mode.get("matrix_format", "csr")
    B, onodes = oracle_parse_gfa(text, return_node_list=True, **omode)
    B = oracle_convert_format(B, fmt)
    assert A.format == B.format and A.shape == B.shape and A.dtype == B.dtype
    assert np.array_equal(A.indptr, B.indptr) and np.array_equal(A.indices, B.indices)
    assert A.data.tobytes() == B.data.tobytes()
    assert nodes == onodes


def test_dist_refuses_weights_and_long_names():
    import torch

    from gfa2network_b200 import _capi
    from gfa2network_b200 import dist as D

    r = D.LocalRank(0, 0, 1)
    long_text = torch.from_numpy(np.frombuffer(b"S\tthis_name_is_longer_than_15_bytes\t*\n", dtype=np.uint8).copy()).cuda()
    with pytest.raises(NotImplementedError):
        r.scan(long_text)
    params = _capi.Params(1, 0, 0, 0, 0, 0, 1, 1, b"RC", 2, 0)
    info = _capi.DistInfo()
    import ctypes as C

    rc = r.h.lib.g2n_dist_scan(r.h.h, C.c_void_p(long_text.data_ptr()), long_text.numel(), C.byref(params), C.byref(info))
    assert rc == _capi.G2N_ERR_UNSUPPORTED
