"""Multi-GPU path on ONE GPU: G logical ranks (one LocalRank / handle / exchange arena each) whose peers
are plain device pointers -- the same kernels, protocol and host logic as `DistBuilder` (the stages are
queued rank by rank, so every signal is already there when its consumer runs) -- compared bit-for-bit
with the CPU oracle.  Both kinds of build are covered: host-planned (first build of a shape) and
speculative (no host round trip), and the capacity-miss path between them."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _shards(text: np.ndarray, G: int):
    import torch

    from gfa2network_b200 import dist as D

    tb = text.tobytes()
    out = []
    for r in range(G):
        lo, hi = D.shard_range(len(tb), r, G, lambda p: tb.find(b"\n", p))
        out.append(torch.from_numpy(text[lo:hi].copy()).cuda())
    return out


def _connect(ranks):
    mems = [r.local_mem() for r in ranks]
    for r in ranks:
        r.set_peers([m[0] for m in mems], [m[1] for m in mems])


def _build(ranks, shards, speculative, caps=None, attempt=0, **mode):
    """One build over all logical ranks; returns ([(rc, result)], caps)."""
    from gfa2network_b200 import dist as D

    G = len(ranks)
    for r, t in zip(ranks, shards):
        r.set_input(t, **mode)
    if not speculative:
        infos = [r.probe() for r in ranks]
        D.raise_agreed(infos)
        if caps is None:
            caps = D.plan_caps(infos, G, attempt)
        for r in ranks:
            r.plan(*caps)
        _connect(ranks)
    for k in range(D.N_STAGES):
        for r in ranks:
            r.stage(k, speculative)
    out = [r.finish() for r in ranks]
    return out, caps


def _assemble(ranks, out, fmt):
    from gfa2network_b200 import dist as D

    ng = {int(res.n_global) for _, res in out}
    assert len(ng) == 1
    ng = ng.pop()
    slabs = [tuple(np.array(a) for a in r.fetch_slab()) for r in ranks]
    A = D.assemble_slabs(slabs, ng, fmt)
    nodes = []
    for r in ranks:
        id0, names = r.node_list()
        assert id0 == len(nodes)
        nodes.extend(names)
    return A, nodes


def _check(A, nodes, text, mode):
    from oracle.oracle import oracle_convert_format, oracle_parse_gfa

    omode = {k: v for k, v in mode.items() if k != "matrix_format"}
    fmt = mode.get("matrix_format", "csr")
    B, onodes = oracle_parse_gfa(text, return_node_list=True, **omode)
    B = oracle_convert_format(B, fmt)
    assert A.format == B.format and A.shape == B.shape and A.dtype == B.dtype
    assert np.array_equal(A.indptr, B.indptr) and np.array_equal(A.indices, B.indices)
    assert A.data.tobytes() == B.data.tobytes()
    assert nodes == onodes


MODES = [dict(), dict(directed=False), dict(asymmetric=True), dict(bidirected=True), dict(bidirected=True, keep_directed_bidir=True),
         dict(directed=False, matrix_format="csc"), dict(dtype="bool")]


@pytest.mark.parametrize("G", [1, 2, 3, 8])
@pytest.mark.parametrize("mode", MODES, ids=[str(m) for m in MODES])
def test_logical_shards_match_oracle(G, mode):
    from gfa2network_b200 import _capi
    from gfa2network_b200 import dist as D
    from gfa2network_b200.synth import synth_gfa

    text = synth_gfa(30_000, 90_000, seed=9, kind=1, interleave=2048, n_paths=1)
    shards = _shards(text, G)
    ranks = [D.LocalRank(0, r, G) for r in range(G)]
    fmt = mode.get("matrix_format", "csr")
    out, caps = _build(ranks, shards, False, **mode)  # host-planned
    assert all(rc == _capi.G2N_OK for rc, _ in out)
    _check(*_assemble(ranks, out, fmt), text, mode)
    for r, (_, res) in zip(ranks, out):
        r.remember(res, *caps)
    out, _ = _build(ranks, shards, True, **mode)  # speculative: same kernels, sizes read on the device
    assert all(rc == _capi.G2N_OK for rc, _ in out), [hex(int(res.bad)) for _, res in out]
    assert all(r.h.status().speculative == 1 for r in ranks)
    _check(*_assemble(ranks, out, fmt), text, mode)


def test_capacity_miss_repeats_everywhere():
    """A speculative build whose input outgrew the remembered plan ends with the same verdict on every
    rank (G2N_ERR_RETRY), and the host-planned build that follows is right."""
    from gfa2network_b200 import _capi
    from gfa2network_b200 import dist as D
    from gfa2network_b200.synth import synth_gfa

    G = 3
    small = synth_gfa(4_000, 12_000, seed=5, kind=1)
    big = synth_gfa(20_000, 60_000, seed=6, kind=1, interleave=512)
    ranks = [D.LocalRank(0, r, G) for r in range(G)]
    out, caps = _build(ranks, _shards(small, G), False)
    assert all(rc == _capi.G2N_OK for rc, _ in out)
    for r, (_, res) in zip(ranks, out):
        r.remember(res, *caps)
    out, _ = _build(ranks, _shards(big, G), True)
    assert all(rc == _capi.G2N_ERR_RETRY for rc, _ in out)
    assert len({int(res.bad) for _, res in out}) == 1 and int(out[0][1].bad) != 0
    out, caps = _build(ranks, _shards(big, G), False)
    assert all(rc == _capi.G2N_OK for rc, _ in out)
    _check(*_assemble(ranks, out, "csr"), big, {})
    # only ONE shard changes: the others still fit their remembered sizes, the verdict is shared all the same
    for r, (_, res) in zip(ranks, out):
        r.remember(res, *caps)
    import torch

    extra = np.frombuffer(b"".join(b"S\textra%d\t*\n" % i for i in range(30_000)), dtype=np.uint8)
    grown = np.concatenate([big, extra])
    sh = _shards(big, G)
    sh[G - 1] = torch.from_numpy(np.concatenate([sh[G - 1].cpu().numpy(), extra])).cuda()
    out, _ = _build(ranks, sh, True)
    assert all(rc == _capi.G2N_ERR_RETRY for rc, _ in out)
    out, _ = _build(ranks, sh, False)
    assert all(rc == _capi.G2N_OK for rc, _ in out)
    _check(*_assemble(ranks, out, "csr"), grown, {})


def test_key_segment_overflow_is_a_retry():
    from gfa2network_b200 import _capi
    from gfa2network_b200 import dist as D
    from gfa2network_b200.synth import synth_gfa

    G = 2
    text = synth_gfa(20_000, 40_000, seed=3, kind=1)
    ranks = [D.LocalRank(0, r, G) for r in range(G)]
    out, _ = _build(ranks, _shards(text, G), False, caps=(1024, 200_000))  # 1024 keys per segment cannot hold ~5000
    assert all(rc == _capi.G2N_ERR_RETRY for rc, _ in out)
    assert all(int(res.bad) & 2 for _, res in out)  # DXB_KEYS, known to every rank
    out, _ = _build(ranks, _shards(text, G), False)
    assert all(rc == _capi.G2N_OK for rc, _ in out)
    _check(*_assemble(ranks, out, "csr"), text, {})


def test_errors_and_warnings_are_agreed():
    import torch

    from gfa2network_b200 import dist as D

    G = 2
    ranks = [D.LocalRank(0, r, G) for r in range(G)]

    def probe(parts):
        for r, p in zip(ranks, parts):
            r.set_input(torch.from_numpy(np.frombuffer(p, dtype=np.uint8).copy()).cuda())
        return [r.probe() for r in ranks]

    # the malformed record sits in the second shard, the unsupported record before it in the first
    infos = probe([b"S\ta\t*\nW\tz\n", b"S\tb\t*\nL\ta\t+\n"])
    with pytest.warns(RuntimeWarning, match="Skipping unsupported record: W"):
        with pytest.raises(ValueError, match="Malformed L record"):
            D.raise_agreed(infos)
    # an unsupported record AFTER the error line is never reported
    infos = probe([b"S\ta\t*\nP\tonly\n", b"#\tcomment\nS\tb\t*\n"])
    import warnings

    with warnings.catch_warnings():
        warnings.simplefilter("error")
        with pytest.raises(ValueError, match="Malformed P record"):
            D.raise_agreed(infos)


def _weighted_text(seed: int, n: int = 6000) -> np.ndarray:
    """Chain links with inexact float weights, duplicates and reverse duplicates in shuffled order: duplicate
    sums depend on the summation order, rows stay below SciPy's 16-entry insertion-sort regime."""
    rng = np.random.default_rng(seed)
    lines = [b"S\ts%d\t*" % i for i in range(n)]
    links = []
    for i in range(n - 1):
        links.append(b"L\ts%d\t+\ts%d\t-\t0M\tRC:f:%.3f" % (i, i + 1, rng.integers(1, 9000) / 7))
        if rng.random() < 0.5:
            links.append(b"L\ts%d\t+\ts%d\t-\t0M\tRC:f:%.3f" % (i, i + 1, rng.integers(1, 9000) / 11))
        if rng.random() < 0.3:
            links.append(b"L\ts%d\t-\ts%d\t+\t0M\tXX:i:3\tRC:i:%d" % (i + 1, i, rng.integers(-5, 50)))
        if rng.random() < 0.1:
            links.append(b"L\ts%d\t+\ts%d\t-\t0M" % (i, i + 1))  # no tag: weight 1.0
    order = rng.permutation(len(links))
    lines += [links[k] for k in order]
    return np.frombuffer(b"\n".join(lines) + b"\n", dtype=np.uint8)


WMODES = [dict(), dict(directed=False), dict(asymmetric=True), dict(bidirected=True), dict(bidirected=True, keep_directed_bidir=True),
          dict(dtype="float32"), dict(dtype="int32", directed=False, matrix_format="csc")]


@pytest.mark.parametrize("G", [2, 5])
@pytest.mark.parametrize("mode", WMODES, ids=[str(m) for m in WMODES])
def test_weighted_logical_shards_match_oracle(G, mode):
    """With a weight tag the row entries travel in emission order with their weight: duplicate sums are bit-identical
    to the reference's (SciPy's) order although the duplicates sit in different shards."""
    from gfa2network_b200 import _capi
    from gfa2network_b200 import dist as D

    text = _weighted_text(17)
    mode = dict(mode, weight_tag="RC")
    shards = _shards(text, G)
    ranks = [D.LocalRank(0, r, G) for r in range(G)]
    fmt = mode.get("matrix_format", "csr")
    out, caps = _build(ranks, shards, False, **mode)
    assert all(rc == _capi.G2N_OK for rc, _ in out)
    _check(*_assemble(ranks, out, fmt), text, mode)
    for r, (_, res) in zip(ranks, out):
        r.remember(res, *caps)
    out, _ = _build(ranks, shards, True, **mode)
    assert all(rc == _capi.G2N_OK for rc, _ in out), [hex(int(res.bad)) for _, res in out]
    _check(*_assemble(ranks, out, fmt), text, mode)


def test_weighted_c3_dialect_shards():
    """The C3 shape (reference E-dialect with RC:f tags, bidirected) over 3 logical ranks."""
    from gfa2network_b200 import _capi
    from gfa2network_b200 import dist as D
    from gfa2network_b200.synth import synth_gfa

    text = synth_gfa(20_000, 60_000, seed=3, kind=2)
    mode = dict(bidirected=True, weight_tag="RC")
    G = 3
    ranks = [D.LocalRank(0, r, G) for r in range(G)]
    out, _ = _build(ranks, _shards(text, G), False, **mode)
    assert all(rc == _capi.G2N_OK for rc, _ in out)
    _check(*_assemble(ranks, out, "csr"), text, mode)


def test_empty_and_lopsided_shards():
    """Shards without any record, shards without segments (links only) and a world larger than the node count."""
    from gfa2network_b200 import _capi
    from gfa2network_b200 import dist as D

    import torch

    parts = [b"H\tVN:Z:1.0\n", b"S\ta\t*\nS\tb\t*\nS\tc\t*\n", b"", b"L\tc\t+\ta\t-\t0M\nL\ta\t+\tb\t+\t0M\nL\td\t+\ta\t+\t0M\n"]
    text = np.frombuffer(b"".join(parts), dtype=np.uint8)
    G = len(parts)
    ranks = [D.LocalRank(0, r, G) for r in range(G)]
    shards = [torch.from_numpy(np.frombuffer(p, dtype=np.uint8).copy()).cuda() if p else torch.empty(0, dtype=torch.uint8, device="cuda") for p in parts]
    for mode in (dict(), dict(directed=False), dict(bidirected=True)):
        out, caps = _build(ranks, shards, False, **mode)
        assert all(rc == _capi.G2N_OK for rc, _ in out)
        _check(*_assemble(ranks, out, "csr"), text, mode)


@pytest.mark.parametrize("mode", [dict(), dict(bidirected=True), dict(directed=False)], ids=str)
def test_long_names_across_shards(mode):
    """Names beyond the 15-byte inline key (hashed keys, bytes kept by the shard that saw them first),
    mentioned from several shards, next to short ones."""
    from gfa2network_b200 import _capi
    from gfa2network_b200 import dist as D

    rng = np.random.default_rng(4)
    names = [b"s%d" % i for i in range(300)] + [b"chromosome_segment_with_a_long_name_%06d" % i for i in range(300)] + [b"exactly15bytes_%d" % (i % 10) for i in range(10)] + [b"sixteen_bytes_ab%d" % (i % 10) for i in range(10)]
    lines = [b"H\tVN:Z:1.0"]
    for k in rng.permutation(len(names))[:400]:
        lines.append(b"S\t" + names[k] + b"\t*")
    for _ in range(3000):
        u, v = rng.integers(0, len(names), 2)
        lines.append(b"L\t" + names[u] + b"\t" + b"+-"[rng.integers(0, 2):][:1] + b"\t" + names[v] + b"\t" + b"+-"[rng.integers(0, 2):][:1] + b"\t0M")
    text = np.frombuffer(b"\n".join(lines) + b"\n", dtype=np.uint8)
    for G in (2, 5):
        ranks = [D.LocalRank(0, r, G) for r in range(G)]
        shards = _shards(text, G)
        out, caps = _build(ranks, shards, False, **mode)
        assert all(rc == _capi.G2N_OK for rc, _ in out)
        _check(*_assemble(ranks, out, "csr"), text, mode)
        for r, (_, res) in zip(ranks, out):
            r.remember(res, *caps)
        out, _ = _build(ranks, shards, True, **mode)
        assert all(rc == _capi.G2N_OK for rc, _ in out)
        _check(*_assemble(ranks, out, "csr"), text, mode)


@pytest.mark.parametrize("variant", [dict(), dict(G2N_DBG_SUBCAP="8"), dict(G2N_DBG_NOSUB="1"), dict(G2N_DBG_NOBUCKET="1")],
                         ids=["sub-buckets", "sub-buckets-overflow", "buckets", "row-range-passes"])
@pytest.mark.parametrize("mode", [dict(), dict(weight_tag="RC", bidirected=True)], ids=str)
def test_slab_row_range_passes(monkeypatch, mode, variant):
    """A slab far larger than L2 (forced): received entries partitioned by row bucket (dist.cuh: k_pairs_bucket_*), or
    -- G2N_DBG_NOBUCKET -- the bucketing kernels in several row-range passes."""
    for k, v in variant.items():
        monkeypatch.setenv(k, v)
    from gfa2network_b200 import _capi
    from gfa2network_b200 import dist as D
    from gfa2network_b200.synth import synth_gfa

    monkeypatch.setenv("G2N_DBG_ROWPASS", "5")
    text = synth_gfa(20_000, 60_000, seed=8, kind=2 if mode else 1)
    G = 3
    ranks = [D.LocalRank(0, r, G) for r in range(G)]
    out, _ = _build(ranks, _shards(text, G), False, **mode)
    assert all(rc == _capi.G2N_OK for rc, _ in out)
    _check(*_assemble(ranks, out, "csr"), text, mode)


@pytest.mark.parametrize("variant", [dict(), dict(G2N_DBG_SUBCAP="8"), dict(G2N_DBG_NOBUCKET="1")], ids=["sub-buckets", "sub-buckets-overflow", "row-range-passes"])
@pytest.mark.parametrize("mode", [dict(), dict(weight_tag="RC", bidirected=True)], ids=str)
def test_slab_partitioned_with_hub_rows(monkeypatch, mode, variant):
    """Rows of hundreds of entries (hub segments, repeated links) in a slab whose receive side is partitioned (forced):
    left unsorted by the placement, sorted by k_rows_big, counted by k_rows_sort_rest."""
    import golden_inputs as gi
    from gfa2network_b200 import _capi
    from gfa2network_b200 import dist as D

    monkeypatch.setenv("G2N_DBG_ROWPASS", "3")
    for k, v in variant.items():
        monkeypatch.setenv(k, v)
    text = np.frombuffer(gi.hub_text(), dtype=np.uint8)
    G = 3
    ranks = [D.LocalRank(0, r, G) for r in range(G)]
    out, _ = _build(ranks, _shards(text, G), False, **mode)
    assert all(rc == _capi.G2N_OK for rc, _ in out)
    _check(*_assemble(ranks, out, "csr"), text, mode)


def _n_gpus():
    import torch

    return torch.cuda.device_count()


@pytest.mark.parametrize("mode", [dict(), dict(directed=False, matrix_format="csr"), dict(bidirected=True, matrix_format="csc", return_node_list=True),
                                  dict(weight_tag="RC", matrix_format="csr")], ids=str)
def test_parse_gfa_devices_file_split(tmp_path, mode):
    """parse_gfa(path, devices=[...]): ONE process, one byte range of the FILE per GPU (newline-aligned cuts, each rank
    loads its own range), same matrix and node list as the oracle over the whole file.  Needs >= 2 GPUs (skipped on
    the single-GPU test box; run with `gpurun --gpus 2`)."""
    if _n_gpus() < 2:
        pytest.skip("needs two GPUs")
    from gfa2network_b200 import parse_gfa
    from gfa2network_b200.synth import synth_gfa
    from oracle.oracle import oracle_convert_format, oracle_parse_gfa

    kind = 2 if mode.get("weight_tag") else 1
    text = synth_gfa(40_000, 120_000, seed=11, kind=kind, interleave=1024, n_paths=1, n_walks=1)
    f = tmp_path / "g.gfa"
    f.write_bytes(text.tobytes())
    devices = list(range(min(_n_gpus(), 4)))
    omode = {k: v for k, v in mode.items() if k not in ("matrix_format", "return_node_list")}
    fmt = mode.get("matrix_format", "csr")
    import warnings

    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        got = parse_gfa(str(f), build_graph=False, build_matrix=True, devices=devices, **mode)
    assert [str(x.message) for x in w] == ["Skipping unsupported record: W"]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        B, onodes = oracle_parse_gfa(text, return_node_list=True, **omode)
    B = oracle_convert_format(B, fmt)
    A, nodes = got if mode.get("return_node_list") else (got, None)
    assert A.format == B.format and A.shape == B.shape and A.dtype == B.dtype
    assert np.array_equal(A.indptr, B.indptr) and np.array_equal(A.indices, B.indices) and A.data.tobytes() == B.data.tobytes()
    if nodes is not None:
        assert nodes == onodes
    # the second build of the same shape is speculative (no host round trip) and gives the same result
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        A2 = parse_gfa(str(f), build_graph=False, build_matrix=True, devices=devices, **{k: v for k, v in mode.items() if k != "return_node_list"})
    assert np.array_equal(A2.indptr, B.indptr) and np.array_equal(A2.indices, B.indices) and A2.data.tobytes() == B.data.tobytes()


def test_parse_gfa_devices_errors(tmp_path):
    if _n_gpus() < 2:
        pytest.skip("needs two GPUs")
    from gfa2network_b200 import parse_gfa

    f = tmp_path / "bad.gfa"
    f.write_bytes(b"S\ta\t*\n" * 50 + b"W\tx\n" + b"S\tb\t*\n" * 50 + b"L\tonly\n" + b"S\tc\t*\n" * 50)
    with pytest.warns(RuntimeWarning, match="Skipping unsupported record: W"):
        with pytest.raises(ValueError, match="Malformed L record"):
            parse_gfa(str(f), build_graph=False, build_matrix=True, devices=[0, 1])
    with pytest.raises(NotImplementedError):
        parse_gfa(str(f), build_graph=False, build_matrix=True, devices=[0, 1], asymmetric=True)


def test_cli_convert_devices(tmp_path):
    if _n_gpus() < 2:
        pytest.skip("needs two GPUs")
    import scipy.sparse as sp

    from gfa2network_b200.cli import main
    from gfa2network_b200.synth import synth_gfa
    from oracle.oracle import oracle_parse_gfa

    text = synth_gfa(5_000, 15_000, seed=12)
    f = tmp_path / "g.gfa"
    f.write_bytes(text.tobytes())
    out = tmp_path / "m.npz"
    main(["convert", str(f), "--matrix", str(out), "--devices", "0,1"])
    A = sp.load_npz(out)
    B, onodes = oracle_parse_gfa(text, return_node_list=True)
    assert A.format == "csr" and np.array_equal(A.indptr, B.indptr) and np.array_equal(A.indices, B.indices) and np.array_equal(A.data, B.data)
    lines = (tmp_path / "m.npz.nodes.tsv").read_text().splitlines()
    assert lines == [f"{i}\t{n}" for i, n in enumerate(onodes)]
