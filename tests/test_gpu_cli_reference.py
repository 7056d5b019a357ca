"""The reference's own CLI tests of the `convert --matrix` path, run against this package's CLI
(`python -m gfa2network_b200 …` / `gfa2network_b200.cli.main`): tests/test_matrix_asym.py, test_matrix_dtype.py,
test_matrix_nodes_map.py, test_limits.py, test_bidirected.py (matrix side) of the reference, plus the flags
cli.py:53-135 adds around them.  Same inputs, same assertions."""
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest
import scipy.sparse as sp

pytestmark = pytest.mark.gpu

SAMPLE_GFA = b"S\ts1\t4\nS\ts2\t4\nL\ts1\t+\ts2\t-\t0M\n"


def _write(tmp_path: Path, content: bytes = SAMPLE_GFA, name: str = "sample.gfa") -> Path:
    p = tmp_path / name
    p.write_bytes(content)
    return p


def _cli(*args):
    subprocess.run([sys.executable, "-m", "gfa2network_b200", *map(str, args)], check=True)


def test_matrix_asymmetric(tmp_path):  # tests/test_matrix_asym.py
    out = tmp_path / "adj.npz"
    _cli("convert", _write(tmp_path), "--matrix", out, "--asymmetric")
    arr = sp.load_npz(out).toarray()
    assert not (arr == arr.T).all()


def test_matrix_dtype(tmp_path):  # tests/test_matrix_dtype.py
    out = tmp_path / "adj.npz"
    _cli("convert", _write(tmp_path), "--matrix", out, "--dtype", "bool")
    assert sp.load_npz(out).dtype == bool


def test_matrix_node_map(tmp_path):  # tests/test_matrix_nodes_map.py
    out = tmp_path / "adj.npz"
    _cli("convert", _write(tmp_path), "--matrix", out)
    A = sp.load_npz(out)
    lines = Path(str(out) + ".nodes.tsv").read_text().strip().splitlines()
    assert len(lines) == A.shape[0]
    for i, line in enumerate(lines):
        idx, node = line.split("\t")
        assert int(idx) == i
    assert [ln.split("\t")[1] for ln in lines] == ["s1", "s2"]


def _chain(tmp_path: Path, n: int) -> Path:  # tests/test_limits.py:7-15
    text = "".join(f"S\t{i}\t*\n" for i in range(n)) + "".join(f"L\t{i}\t+\t{i + 1}\t+\t0M\n" for i in range(n - 1))
    return _write(tmp_path, text.encode(), "big.gfa")


def test_dense_matrix_limit(tmp_path):  # tests/test_limits.py
    from gfa2network_b200.cli import main

    out = tmp_path / "dense.npy"
    with pytest.raises(SystemExit):
        main(["convert", str(_chain(tmp_path, 400)), "--matrix", str(out), "--max-dense-gb", "0.001"])


def test_dense_matrix_limit_respects_dtype(tmp_path):  # tests/test_limits.py
    from gfa2network_b200.cli import main

    out = tmp_path / "dense.npy"
    main(["--max-dense-gb", "0.001", "convert", str(_chain(tmp_path, 400)), "--matrix", str(out), "--dtype", "float32"])
    assert out.exists()
    arr = np.load(out)
    assert arr.shape == (400, 400) and arr.dtype == np.float32 and arr.sum() == 2 * 399  # symmetrised by default (Q7)


def test_flags_around_the_matrix_path(tmp_path):
    """--undirected, --bidirected, --weight-tag, --matrix-format, --no-node-map, the hidden --save-matrix alias (cli.py:53-135)."""
    gfa = _write(tmp_path, b"S\ta\t*\nS\tb\t*\nL\ta\t+\tb\t-\t0M\tRC:i:7\nL\ta\t+\tb\t-\t0M\tRC:i:5\n")
    out = tmp_path / "m.npz"
    _cli("convert", gfa, "--save-matrix", out, "--undirected", "--weight-tag", "RC", "--matrix-format", "csc", "--no-node-map")
    A = sp.load_npz(out)
    assert A.format == "csc" and not Path(str(out) + ".nodes.tsv").exists()
    assert A.toarray().tolist() == [[0.0, 12.0], [12.0, 0.0]]
    _cli("convert", gfa, "--matrix", out, "--bidirected", "--dtype", "int32")
    B = sp.load_npz(out)
    names = [ln.split("\t")[1] for ln in Path(str(out) + ".nodes.tsv").read_text().splitlines()]
    assert names == ["a:+", "a:-", "b:+", "b:-"] and B.shape == (4, 4) and B.dtype == np.int32
    # tests/test_bidirected.py:17-19: the reverse-complement edge b:+ -> a:- exists next to a:+ -> b:-
    arr = B.toarray()
    assert arr[0, 3] == 2 and arr[2, 1] == 2
    r = subprocess.run([sys.executable, "-m", "gfa2network_b200", "convert", str(gfa)], capture_output=True, text=True)
    assert r.returncode != 0 and "convert requires --graph or --matrix" in r.stderr  # cli.py:194-195
