"""Parity of the CUDA path (through the C ABI: gfa2network_b200 -> ctypes -> libg2n.so) against
(a) golden vectors recorded from the real reference and (b) the CPU oracle on seeded synthetic
inputs.  Bit-exact: structure, node list, integer and exactly-representable float weights."""
import base64

import numpy as np
import pytest

import golden_inputs as gi
import parity_util as pu

pytestmark = pytest.mark.gpu

CASES = pu.load_json("cases.json")
FUZZ = pu.load_json("fuzz.json")
DRB1 = pu.load_json("drb1.json")


def _api():
    """parse + convert.  convert() checks BOTH conversion paths against each other: from the device-resident build
    (what the CLI does, g2n_convert) while the matrix's session is live, and from the triplets the matrix holds
    (the public convert_format: upload + stage K4, g2n_coo_to_compressed)."""
    from gfa2network_b200 import convert_format, parse_gfa

    def parse(text, **kw):
        return parse_gfa(text, build_graph=False, build_matrix=True, **kw)

    def convert(A, fmt):
        sess = getattr(A, "_g2n_session", None)
        resident = {f: convert_format(A, f, _untouched=True) for f in ("csr", "csc")} if sess is not None and sess.live() else {}
        out = convert_format(A, fmt)
        if fmt in resident:
            r = resident[fmt]
            assert r.format == out.format and r.dtype == out.dtype and r.shape == out.shape
            for x, y in ((r.indptr, out.indptr), (r.indices, out.indices), (r.data, out.data)):
                assert x.dtype == y.dtype and np.array_equal(x, y, equal_nan=True), fmt
        return out

    return parse, convert


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_literal_cases(case):
    parse, convert = _api()
    text = base64.b64decode(case["text_b64"])
    for run in case["runs"]:
        pu.run_and_compare(parse, convert, text, run["mode"], run["expect"], full=True, what=f"{case['name']} {run['mode']}")


@pytest.mark.parametrize("entry", FUZZ, ids=[str(e["seed"]) for e in FUZZ])
def test_fuzz(entry):
    parse, convert = _api()
    text = gi.fuzz_text(entry["seed"])
    for run in entry["runs"]:
        pu.run_and_compare(parse, convert, text, run["mode"], run["expect"], full=False, what=f"fuzz{entry['seed']} {run['mode']}")


def test_drb1_fixture_c1():
    """BASELINE.json configs[0]: the reference's own CPU-runnable case."""
    parse, convert = _api()
    path = pu.GOLD / "DRB1-3123_unsorted.gfa"
    for run in DRB1["runs"]:
        pu.run_and_compare(parse, convert, path, run["mode"], run["expect"], full=False, what=f"drb1 {run['mode']}")


def _same(A, B, what):
    assert A.format == B.format and A.dtype == B.dtype and A.shape == B.shape, what
    for x, y in zip(pu.arrays_of(A), pu.arrays_of(B)):
        assert x.dtype == y.dtype and x.shape == y.shape, what
        assert x.tobytes() == y.tobytes(), what


SYN_MODES = [
    dict(directed=False),
    dict(),
    dict(asymmetric=True),
    dict(bidirected=True),
    dict(bidirected=True, keep_directed_bidir=True),
    dict(dtype="float32", directed=False),
    dict(dtype="bool"),
]


@pytest.mark.parametrize("mode", SYN_MODES, ids=[str(m) for m in SYN_MODES])
def test_synthetic_gfa1_vs_oracle(mode):
    from gfa2network_b200.synth import synth_gfa
    from oracle.oracle import oracle_convert_format, oracle_parse_gfa

    parse, convert = _api()
    text = synth_gfa(100_000, 300_000, seed=2, kind=1, n_paths=2, n_walks=2, interleave=4096)
    A, nodes = parse(text, return_node_list=True, **mode)
    B, onodes = oracle_parse_gfa(text, return_node_list=True, **mode)
    _same(A, B, f"raw {mode}")
    assert nodes == onodes
    for fmt in ("csr", "csc"):
        _same(convert(A, fmt), oracle_convert_format(B, fmt), f"{fmt} {mode}")
    # fused build (want_format) gives the same arrays as parse + convert
    from gfa2network_b200 import parse_gfa

    C = parse_gfa(text, build_graph=False, build_matrix=True, matrix_format="csr", **mode)
    _same(C, oracle_convert_format(B, "csr"), f"fused csr {mode}")


def test_synthetic_e_dialect_weighted_vs_oracle():
    """C3 shape at 1/100 scale: reference E dialect, RC:f weights k/8, bidirected."""
    from gfa2network_b200.synth import synth_gfa
    from oracle.oracle import oracle_convert_format, oracle_parse_gfa

    parse, convert = _api()
    text = synth_gfa(100_000, 300_000, seed=3, kind=2)
    for mode in (dict(bidirected=True, weight_tag="RC"), dict(weight_tag="RC"), dict(weight_tag="RC", asymmetric=True)):
        A, nodes = parse(text, return_node_list=True, **mode)
        B, onodes = oracle_parse_gfa(text, return_node_list=True, **mode)
        _same(A, B, f"raw {mode}")
        assert nodes == onodes
        _same(convert(A, "csr"), oracle_convert_format(B, "csr"), f"csr {mode}")


def test_segments_with_sequences_vs_oracle():
    """C5 shape at small scale: S lines carry sequences (mean 270 B), lines cross tile boundaries."""
    from gfa2network_b200.synth import synth_gfa
    from oracle.oracle import oracle_parse_gfa

    parse, _ = _api()
    text = synth_gfa(50_000, 200_000, seed=5, kind=1, seq_mean=270)
    A, nodes = parse(text, return_node_list=True)
    B, onodes = oracle_parse_gfa(text, return_node_list=True)
    _same(A, B, "c5-small")
    assert nodes == onodes


def test_hub_rows_vs_oracle():
    """Rows far above the one-lane limit (64) and above the shared-memory bitonic limit (4096):
    a hub linked to 12 000 spokes, with duplicate and reverse links and integer weights."""
    import random

    from oracle.oracle import oracle_convert_format, oracle_parse_gfa

    parse, convert = _api()
    r = random.Random(5)
    lines = ["S\thub\t*", "S\tmid\t*"]
    for i in range(12000):
        lines.append(f"L\thub\t+\tn{i % 9000}\t-\t0M\tRC:i:{r.randrange(1, 50)}")
    for i in range(300):
        lines.append(f"L\tn{r.randrange(9000)}\t+\tmid\t+\t0M\tRC:i:{r.randrange(1, 9)}")
    for i in range(3000):
        lines.append(f"L\tn{r.randrange(9000)}\t-\thub\t+\t0M")
    text = ("\n".join(lines) + "\n").encode()
    for mode in (dict(weight_tag="RC"), dict(weight_tag="RC", directed=False), dict(asymmetric=True), dict(bidirected=True, weight_tag="RC")):
        A, nodes = parse(text, return_node_list=True, **mode)
        B, onodes = oracle_parse_gfa(text, return_node_list=True, **mode)
        _same(A, B, f"hub raw {mode}")
        assert nodes == onodes
        for fmt in ("csr", "csc"):
            _same(convert(A, fmt), oracle_convert_format(B, fmt), f"hub {fmt} {mode}")


def test_full_size_c2_properties():
    """BASELINE.json configs[1] at full size: size-independent properties + oracle check."""
    from gfa2network_b200.synth import CONFIGS, synth_gfa
    from oracle.oracle import oracle_convert_format, oracle_parse_gfa
    from gfa2network_b200 import parse_gfa

    cfg = CONFIGS["C2"]
    text = synth_gfa(cfg["n_seg"], cfg["n_link"], seed=cfg["seed"], kind=cfg["kind"])
    A = parse_gfa(text, build_graph=False, build_matrix=True, matrix_format="csr", **cfg["mode"])
    assert A.format == "csr" and A.shape == (cfg["n_seg"],) * 2
    assert A.data.sum() == 2 * cfg["n_link"]            # undirected: every link emitted twice, weights 1.0
    assert (A != A.T).nnz == 0                          # symmetric
    ind = A.indices
    rows = np.repeat(np.arange(A.shape[0]), np.diff(A.indptr))
    assert np.all((np.diff(ind) > 0) | (np.diff(rows) > 0))  # canonical: sorted, no duplicates
    B = oracle_convert_format(oracle_parse_gfa(text, **cfg["mode"]), "csr")
    _same(A, B, "C2 full")


def test_cli_convert_drb1(tmp_path, capsys):
    import scipy.sparse as sp
    from gfa2network_b200.cli import main

    out = tmp_path / "drb1.npz"
    main(["convert", str(pu.GOLD / "DRB1-3123_unsorted.gfa"), "--matrix", str(out)])
    assert "Using backend: networkx" in capsys.readouterr().out
    A = sp.load_npz(out)
    exp = DRB1["runs"][0]["expect"]["csr"]
    assert pu.sha(A.indptr, A.indices, A.data) == exp["sha"]
    lines = (tmp_path / "drb1.npz.nodes.tsv").read_text().splitlines()
    assert len(lines) == A.shape[0] and lines[0].split("\t")[0] == "0"


def test_standalone_coo_to_compressed():
    import scipy.sparse as sp
    from gfa2network_b200 import convert_format

    rng = np.random.default_rng(7)
    n, nnz = 5000, 60000
    r = rng.integers(0, n, nnz).astype(np.int32)
    c = rng.integers(0, 40, nnz).astype(np.int32)
    d = rng.integers(-3, 9, nnz).astype(np.float64)
    A = sp.coo_matrix((d, (r, c)), shape=(n, n))
    for fmt in ("csr", "csc"):
        got, want = convert_format(A, fmt), A.asformat(fmt)
        _same(got, want, fmt)


def test_c5_shaped_shard_vs_oracle():
    """BASELINE.json configs[4] shape (segments with sequences, mean 270 B; 4 links per segment; directed
    default = max(S, S^T) CSR) at a size the oracle finishes in seconds: full comparison + properties."""
    from gfa2network_b200 import parse_gfa
    from gfa2network_b200.synth import CONFIGS, synth_gfa
    from oracle.oracle import oracle_convert_format, oracle_parse_gfa

    cfg = CONFIGS["C5"]
    n_seg, n_link = 400_000, 1_600_000
    text = synth_gfa(n_seg, n_link, seed=cfg["seed"], kind=cfg["kind"], seq_mean=cfg["seq_mean"])
    assert text.size > 100_000_000
    A, nodes = parse_gfa(text, build_graph=False, build_matrix=True, return_node_list=True, **cfg["mode"])
    B, onodes = oracle_parse_gfa(text, return_node_list=True, **cfg["mode"])
    _same(A, B, "C5 shard")
    assert nodes == onodes
    assert A.format == "csr" and (A != A.T).nnz == 0 and A.data.min() == 1.0
    from gfa2network_b200 import convert_format

    _same(convert_format(A, "csc"), oracle_convert_format(B, "csc"), "C5 shard csc")


def test_conditional_first_appearance_variant(monkeypatch):
    """The tokenizer specialisation used when the table is far larger than L2 (TM_COND), forced on small
    inputs: same results in every mode."""
    from gfa2network_b200 import parse_gfa
    from gfa2network_b200.synth import synth_gfa
    from oracle.oracle import oracle_parse_gfa

    monkeypatch.setenv("G2N_DBG_COND", "1")
    plain = synth_gfa(30_000, 90_000, seed=41, kind=1, n_paths=1, n_walks=1, interleave=1)
    wtd = synth_gfa(10_000, 30_000, seed=42, kind=2)
    for text, mode in ((plain, dict()), (plain, dict(directed=False)), (plain, dict(bidirected=True)),
                       (wtd, dict(bidirected=True, weight_tag="RC")), (wtd, dict(weight_tag="RC", asymmetric=True))):
        A, nodes = parse_gfa(text, build_graph=False, build_matrix=True, return_node_list=True, **mode)
        B, onodes = oracle_parse_gfa(text, return_node_list=True, **mode)
        _same(A, B, f"cond {mode}")
        assert nodes == onodes


@pytest.mark.parametrize("variant", [dict(), dict(G2N_DBG_SUBCAP="8"), dict(G2N_DBG_NOSUB="1"), dict(G2N_DBG_NOBUCKET="1")],
                         ids=["sub-buckets", "sub-buckets-overflow", "buckets", "row-range-passes"])
@pytest.mark.parametrize("passes", ["3", "64"])
def test_row_range_passes_variant(monkeypatch, passes, variant):
    """Row arrays far larger than L2: the entries are partitioned by row bucket before the histogram / scatter passes
    (rowsort.cuh: RowBuckets) -- or, G2N_DBG_NOBUCKET, the bucketing kernels run once per slice of rows (RowRange).
    Forced on small inputs: same results in every mode, on repeated (speculative) builds and after a convert."""
    for k, v in variant.items():
        monkeypatch.setenv(k, v)
    from gfa2network_b200 import convert_format, parse_gfa
    from gfa2network_b200.synth import synth_gfa
    from oracle.oracle import oracle_convert_format, oracle_parse_gfa

    monkeypatch.setenv("G2N_DBG_ROWPASS", passes)
    plain = synth_gfa(30_000, 90_000, seed=44, kind=1, n_paths=1, interleave=7)
    wtd = synth_gfa(10_000, 30_000, seed=45, kind=2)
    for text, mode in ((plain, dict()), (plain, dict(directed=False)), (plain, dict(bidirected=True)), (plain, dict(asymmetric=True)),
                       (wtd, dict(bidirected=True, weight_tag="RC")), (wtd, dict(weight_tag="RC")), (wtd, dict(weight_tag="RC", asymmetric=True))):
        for rep in range(2):
            A = parse_gfa(text, build_graph=False, build_matrix=True, **mode)
            B = oracle_parse_gfa(text, **mode)
            _same(A, B, f"rowpass {mode} rep {rep}")
            for fmt in ("csr", "csc"):
                _same(convert_format(A, fmt), oracle_convert_format(B, fmt), f"rowpass {mode} {fmt}")


@pytest.mark.parametrize("variant", [dict(), dict(G2N_DBG_SUBCAP="8"), dict(G2N_DBG_NOSUB="1"), dict(G2N_DBG_NOBUCKET="1")],
                         ids=["sub-buckets", "sub-buckets-overflow", "buckets", "row-range-passes"])
def test_partitioned_build_with_hub_rows(monkeypatch, variant):
    """Rows far longer than RS_SMALL (a hub segment linked to hundreds of others, with repeated links) inside the partitioned
    build: the placement leaves them unsorted (SB_UNSORTED), k_rows_big sorts them, k_rows_sort_rest counts their stored
    entries; duplicates are summed (scipy/_coo.py tocsr + sum_duplicates, builders.py:279-283)."""
    from gfa2network_b200 import convert_format, parse_gfa
    from oracle.oracle import oracle_convert_format, oracle_parse_gfa

    monkeypatch.setenv("G2N_DBG_ROWPASS", "3")
    for k, v in variant.items():
        monkeypatch.setenv(k, v)
    text = gi.hub_text()
    for mode in (dict(), dict(directed=False), dict(bidirected=True), dict(asymmetric=True), dict(weight_tag="RC"),
                 dict(weight_tag="RC", bidirected=True), dict(weight_tag="RC", directed=False, dtype="int32")):
        for rep in range(2):
            A = parse_gfa(text, build_graph=False, build_matrix=True, **mode)
            B = oracle_parse_gfa(text, **mode)
            _same(A, B, f"hub {variant} {mode} rep {rep}")
            for fmt in ("csr", "csc"):
                _same(convert_format(A, fmt), oracle_convert_format(B, fmt), f"hub {variant} {mode} {fmt}")


def test_unaligned_device_text():
    """A device pointer that is not 16-byte aligned (a slice of a larger tensor) is accepted."""
    import torch

    from gfa2network_b200 import parse_gfa
    from gfa2network_b200.synth import synth_gfa
    from oracle.oracle import oracle_parse_gfa

    text = synth_gfa(5_000, 15_000, seed=43)
    want = oracle_parse_gfa(text)
    big = torch.zeros(text.size + 64, dtype=torch.uint8, device="cuda")
    for shift in (1, 7, 16, 33):
        big[shift:shift + text.size] = torch.from_numpy(text).cuda()
        _same(parse_gfa(big[shift:shift + text.size], build_graph=False, build_matrix=True), want, f"shift {shift}")
