"""load_paths / genome_distance_matrix (analysis.py:164-272).  CPU: the oracle restatement
against goldens recorded from the real reference (tools/gen_golden_distance.py).  GPU: the device path (CSR in
HBM + cooperative multi-level BFS) against the same goldens and against the oracle on a synthetic input."""
import warnings

import numpy as np
import pytest

import distance_inputs as di
import parity_util as pu

GOLD = pu.load_json("distance.json")
CASES = dict(di.CASES)


def _run(fn_paths, fn_matrix, text, method):
    res = {}
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        try:
            res["paths"] = fn_paths(text)
            labels, M = fn_matrix(text, method)
            res["labels"] = labels
            res["matrix"] = [[("inf" if np.isinf(x) else float(x)) for x in row] for row in np.asarray(M).tolist()]
        except Exception as exc:  # noqa: BLE001
            res["raises"] = {"type": type(exc).__name__, "msg": str(exc)}
    res["warnings"] = sorted({str(x.message) for x in w if issubclass(x.category, RuntimeWarning)})
    return res


def _check(got, expect, what):
    if "raises" in expect:
        assert got.get("raises") == expect["raises"], (what, got.get("raises"))
    else:
        assert "raises" not in got, (what, got.get("raises"))
        assert got["paths"] == expect["paths"], what
        assert list(got["labels"]) == list(expect["labels"] or []), what
        assert got["matrix"] == expect["matrix"], what
    assert got["warnings"] == expect["warnings"], what


@pytest.mark.parametrize("case", GOLD, ids=[c["name"] for c in GOLD])
def test_oracle_distance_matches_reference_goldens(case):
    from oracle.oracle import oracle_distance_matrix, oracle_load_paths

    for run in case["runs"]:
        got = _run(oracle_load_paths, oracle_distance_matrix, CASES[case["name"]], run["method"])
        _check(got, run["expect"], f"{case['name']} {run['method']}")


def _dev_paths(text):
    from gfa2network_b200.analysis import load_paths

    return load_paths(text)


def _dev_matrix(text, method):
    from gfa2network_b200.analysis import genome_distance_matrix

    M = genome_distance_matrix(text, method=method)
    return (list(M.index), M.values) if hasattr(M, "index") else ([], M)


@pytest.mark.gpu
@pytest.mark.parametrize("case", GOLD, ids=[c["name"] for c in GOLD])
def test_device_distance_matches_reference_goldens(case):
    for run in case["runs"]:
        got = _run(_dev_paths, _dev_matrix, CASES[case["name"]], run["method"])
        _check(got, run["expect"], f"{case['name']} {run['method']}")


@pytest.mark.gpu
def test_device_distance_matches_oracle_on_synthetic(tmp_path):
    from gfa2network_b200.analysis import genome_distance, genome_distance_matrix, load_paths
    from oracle.oracle import oracle_distance_matrix, oracle_load_paths

    text = di.random_graph(11, 30_000, 45_000, 10, 40)
    p = tmp_path / "g.gfa"
    p.write_bytes(text)
    assert load_paths(str(p)) == oracle_load_paths(text)
    for method in ("min", "mean"):
        names, want = oracle_distance_matrix(text, method)
        M = genome_distance_matrix(str(p), method=method)
        assert list(M.index) == names and np.array_equal(M.values, want)
    paths = oracle_load_paths(text)
    names, want = oracle_distance_matrix(text, "min")
    if np.isfinite(want[0, 1]):
        assert genome_distance(str(p), paths[names[0]], paths[names[1]]) == want[0, 1]


@pytest.mark.gpu
def test_bfs_levels_match_scipy():
    """The raw level array of a search (g2n_fetch_levels) against SciPy's csgraph on the same CSR."""
    from scipy.sparse.csgraph import dijkstra

    from gfa2network_b200.analysis import _Graph

    text = di.random_graph(12, 5_000, 9_000, 1, 3)
    G = _Graph(text)
    rng = np.random.default_rng(0)
    src = rng.integers(0, G.A.shape[0], 7).astype(np.int32)
    G.bfs(src, 0, 1)
    got = G.levels(0).astype(float)
    got[got < 0] = np.inf
    S = G.A.copy()
    S.data[:] = 1.0
    want = dijkstra(S, directed=True, indices=np.unique(src), unweighted=True, min_only=True)
    assert np.array_equal(got, want)


@pytest.mark.gpu
def test_cli_distance_commands(tmp_path):
    """The reference's CLI tests (tests/test_distance.py:66-131) against this package's CLI."""
    import subprocess
    import sys

    gfa = tmp_path / "paths.gfa"
    gfa.write_bytes(CASES["ref_sample"])
    r = subprocess.run([sys.executable, "-m", "gfa2network_b200", "distance", str(gfa), "--path", "p1", "p2"], capture_output=True, text=True, check=True)
    assert r.stdout.strip() == "0"
    r = subprocess.run([sys.executable, "-m", "gfa2network_b200", "distance", str(gfa), "--path", "p1", "nope"], capture_output=True, text=True)
    assert r.returncode != 0 and "unknown path: nope" in r.stderr
    for extra in ([], ["--backend", "networkx", "--verbose"]):
        out = tmp_path / "dist.csv"
        subprocess.run([sys.executable, "-m", "gfa2network_b200", "distance-matrix", str(gfa), "-o", str(out)] + extra, check=True)
        arr = np.loadtxt(out, delimiter=",")
        assert arr.shape == (2, 2) and np.allclose(arr, [[0, 0], [0, 0]])
    chain = tmp_path / "chain.gfa"
    chain.write_bytes(CASES["chain_directed"])
    r = subprocess.run([sys.executable, "-m", "gfa2network_b200", "distance", str(chain), "--path", "head", "tail"], capture_output=True, text=True, check=True)
    assert r.stdout.strip() == "396"
    r = subprocess.run([sys.executable, "-m", "gfa2network_b200", "distance", str(chain), "--path", "tail", "head", "--undirected"], capture_output=True, text=True, check=True)
    assert r.stdout.strip() == "396"


@pytest.mark.gpu
def test_device_path_nodes_match_host_split():
    """The node lists resolved on the device (csrc/paths.cuh) against the host split of the same records: lines far
    longer than one 4 KiB block, O records, unsigned / empty / unknown entries, a trailing comma, long names."""
    from gfa2network_b200.analysis import _DevicePaths, _Graph
    from gfa2network_b200.synth import synth_gfa
    from oracle.oracle import oracle_load_paths

    big = synth_gfa(60_000, 120_000, seed=31, kind=1, n_paths=3, n_walks=1).tobytes()
    extra = (b"S\ta_segment_name_longer_than_fifteen_bytes\t*\nL\ta_segment_name_longer_than_fifteen_bytes\t+\ts7\t-\t0M\n"
             b"O\two\ts5-,s6,a_segment_name_longer_than_fifteen_bytes+,s5--\textra\n"
             b"P\todd\ts1+,,nope+,s2+,\t*\n"
             b"P\thap0\ts9+\t*\n"      # replaces the list of the first record, keeps its position
             b"P\tlast\ts3+")          # no newline at the end of the file
    text = big + extra
    G = _Graph(text)
    P = _DevicePaths(G, False)
    want = oracle_load_paths(text)
    assert list(P.index) == list(want)
    for name, segs in want.items():
        got = P.nodes(P.index[name])
        exp = np.array([G.index.get(s, -1) for s in segs], dtype=np.int32)
        assert np.array_equal(got, exp), name
    info = P.infos[P.index["odd"]]
    assert info.missing_entry == 1 and info.missing_len == 0  # the empty entry comes first
    with pytest.raises(Exception, match="Node  not found in graph"):
        P.check_sources(P.index["odd"])


@pytest.mark.gpu
def test_genome_distance_min_and_mean_semantics():
    """analysis.py:116-161 on a small graph; the expected values were produced by the real reference
    (genome_distance on the NetworkX graph of the same text)."""
    from gfa2network_b200.analysis import genome_distance

    text = b"S\ta\t*\nS\tb\t*\nS\tc\t*\nS\td\t*\nL\ta\t+\tb\t+\t0M\nL\tb\t+\tc\t+\t0M\nL\td\t+\ta\t+\t0M\n"

    def run(A, B, method):
        try:
            return genome_distance(text, A, B, method=method)
        except Exception as e:  # noqa: BLE001
            return (type(e).__name__, str(e))

    assert run(["a"], ["c"], "mean") == 2.0 and run(["a", "b"], ["c", "c"], "mean") == 1.5 and run(["a"], ["a"], "mean") == 0.0
    assert run(["c"], ["a"], "mean") == ("NetworkXNoPath", "no path between node sets")
    assert run(["a", "zz"], ["c"], "mean") == ("NodeNotFound", "Node zz not found in graph")
    assert run(["a"], ["zz", "c"], "mean") == 2.0  # a target that is not a node is skipped
    assert run(["a"], ["c"], "min") == 2 and run(["c"], ["a"], "min") == ("NetworkXNoPath", "no path between node sets")
    assert run(["zz"], ["a"], "min") == ("NodeNotFound", "Node zz not found in graph")
    assert run(["a"], ["zz"], "min") == ("NetworkXNoPath", "no path between node sets")
    assert run(["a"], ["c"], "bogus") == ("ValueError", "unknown method: bogus")
