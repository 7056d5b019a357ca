"""`export --format edge-list` (cli.py:264-281).  CPU: the oracle restatement against the
golden vectors recorded from the real reference (tools/gen_golden_export.py).  GPU: the device path
against the same vectors, against the oracle on a larger synthetic input, and through the CLI."""
import base64
import hashlib
import subprocess
import sys

import numpy as np
import pytest

import golden_inputs as gi
import parity_util as pu

GOLD = pu.load_json("export.json")


def _compare(got, expect, what):
    data, exc, warns = got
    data = bytes(memoryview(np.ascontiguousarray(data))) if not isinstance(data, bytes) else data
    assert len(data) == expect["nbytes"], what
    assert hashlib.sha256(data).hexdigest()[:16] == expect["sha"], what
    if "out_b64" in expect:
        assert data == base64.b64decode(expect["out_b64"]), what
    if "raises" in expect:
        assert exc is not None and type(exc).__name__ == expect["raises"]["type"], (what, exc)
        if expect["raises"]["type"] != "UnicodeDecodeError":
            assert str(exc) == expect["raises"]["msg"], what
    else:
        assert exc is None, (what, exc)
    assert warns == expect["warnings"], what


def _all_inputs():
    for c in GOLD["cases"]:
        text = dict(gi.LITERAL_CASES)[c["name"]]
        for r in c["runs"]:
            yield f"{c['name']} bidirected={r['bidirected']}", text, r["bidirected"], r["expect"]
    for f in GOLD["fuzz"]:
        text = gi.fuzz_text(f["seed"])
        for r in f["runs"]:
            yield f"fuzz{f['seed']} bidirected={r['bidirected']}", text, r["bidirected"], r["expect"]
    text = (pu.GOLD / "DRB1-3123_unsorted.gfa").read_bytes()
    for r in GOLD["drb1"]:
        yield f"drb1 bidirected={r['bidirected']}", text, r["bidirected"], r["expect"]


def test_oracle_edge_list_matches_reference_goldens():
    from oracle.oracle import oracle_edge_list

    n = 0
    for what, text, bidir, expect in _all_inputs():
        _compare(oracle_edge_list(text, bidirected=bidir), expect, what)
        n += 1
    assert n >= 100


@pytest.mark.gpu
def test_device_edge_list_matches_reference_goldens():
    from gfa2network_b200.export import edge_list_bytes

    for what, text, bidir, expect in _all_inputs():
        _compare(edge_list_bytes(text, bidirected=bidir), expect, what)


@pytest.mark.gpu
@pytest.mark.parametrize("bidir", [False, True])
def test_device_edge_list_matches_oracle_on_synthetic(bidir, tmp_path):
    from gfa2network_b200.export import edge_list_bytes
    from gfa2network_b200.synth import synth_gfa
    from oracle.oracle import oracle_edge_list

    text = synth_gfa(50_000, 150_000, seed=21, kind=1, interleave=4096, n_paths=1, n_walks=1)
    # long names (beyond the 15-byte inline key) and an E / C record mixed in
    extra = b"S\ta_segment_name_longer_than_fifteen_bytes\t*\nL\ta_segment_name_longer_than_fifteen_bytes\t+\ts7\t-\t0M\n" \
            b"E\t*\ts3+\t0\t10\ts4-\t0\t10\t10M\nC\ts5\t+\ts6\t-\t0\t5M\n"
    text = np.concatenate([text, np.frombuffer(extra, dtype=np.uint8)])
    want, wexc, wwarn = oracle_edge_list(text, bidirected=bidir)
    for src in (text, None):
        if src is None:  # the same through a file (g2n_build_file)
            p = tmp_path / "x.gfa"
            p.write_bytes(text.tobytes())
            src = str(p)
        got, exc, warns = edge_list_bytes(src, bidirected=bidir)
        assert got.tobytes() == want and exc is None and wexc is None and warns == wwarn


@pytest.mark.gpu
def test_cli_export_edge_list(tmp_path):
    """The reference's own test (tests/test_export_edge_list.py) against this package's CLI."""
    gfa = tmp_path / "e.gfa"
    gfa.write_bytes(b"S\ts1\t4\nS\ts2\t4\nL\ts1\t+\ts2\t+\t0M\n")
    out = tmp_path / "edges.tsv"
    subprocess.run([sys.executable, "-m", "gfa2network_b200", "export", str(gfa), "--format", "edge-list", "--output", str(out)], check=True)
    assert out.read_text().strip() == "s1\ts2"
    r = subprocess.run([sys.executable, "-m", "gfa2network_b200", "export", str(gfa), "--bidirected"], check=True, capture_output=True)
    assert r.stdout == b"s1:+\ts2:+\n"
