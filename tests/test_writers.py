"""Result writers: the parallel .npz writer against scipy.sparse.save_npz (CPU), the
GPU-made node map against the reference's save_node_map bytes (GPU)."""
import hashlib
import zipfile

import numpy as np
import pytest
import scipy.sparse as sp


def _rand(fmt, n, nnz, dtype, seed=0):
    rng = np.random.default_rng(seed)
    if nnz == 0:
        return sp.coo_matrix((n, n), dtype=dtype).asformat(fmt)
    A = sp.coo_matrix((rng.integers(1, 5, nnz).astype(dtype), (rng.integers(0, n, nnz).astype(np.int32), rng.integers(0, n, nnz).astype(np.int32))), shape=(n, n))
    return A.asformat(fmt)


@pytest.mark.parametrize("fmt", ["csr", "csc", "coo"])
@pytest.mark.parametrize("shape", [(0, 0), (5, 7), (3, 0), (40_000, 500_000)])
@pytest.mark.parametrize("dtype", [np.float64, np.bool_, np.int8])
def test_npz_members_equal_scipy(tmp_path, fmt, shape, dtype):
    from gfa2network_b200.writers import save_npz_parallel

    A = _rand(fmt, shape[0], shape[1], dtype)
    ours, ref = tmp_path / "a.npz", tmp_path / "b.npz"
    save_npz_parallel(ours, A)
    sp.save_npz(ref, A)  # what utils.py:86 calls
    assert zipfile.ZipFile(ours).testzip() is None  # CRCs of the hand-written container
    za, zb = np.load(ours), np.load(ref)
    assert list(za.files) == list(zb.files)
    for k in za.files:
        assert za[k].dtype == zb[k].dtype and za[k].shape == zb[k].shape and za[k].tobytes() == zb[k].tobytes(), k
    B = sp.load_npz(ours)
    assert B.format == A.format and B.dtype == A.dtype and B.shape == A.shape and (B != A).nnz == 0
    assert all(i.compress_type == zipfile.ZIP_DEFLATED for i in zipfile.ZipFile(ours).infolist())


def _ref_node_map_bytes(nodes):
    """utils.py:108-114 verbatim semantics: text file, "{i}\\t{node}\\n"."""
    return "".join(f"{i}\t{n.decode() if isinstance(n, (bytes, bytearray)) else n}\n" for i, n in enumerate(nodes)).encode()


@pytest.mark.gpu
def test_node_map_bytes_match_reference_writer():
    from gfa2network_b200 import _capi, parse_gfa
    from gfa2network_b200.synth import synth_gfa

    long_a, long_b = b"chr1_" + b"x" * 40, "séég_ü".encode()  # > 15 bytes (hashed key) and non-ASCII UTF-8
    odd = b"S\t" + long_a + b"\t*\nS\t" + long_b + b"\t*\nS\ta\t*\nL\t" + long_a + b"\t+\ta\t-\t0M\nL\tzz\t+\t" + long_b + b"\t+\t0M\n"
    cases = [(bytes(synth_gfa(120_000, 300_000, seed=51)), dict()), (bytes(synth_gfa(2_000, 5_000, seed=52)), dict(bidirected=True)),
             (odd, dict()), (odd, dict(bidirected=True)), (b"", dict()), (b"S\tonly\t*\n", dict())]
    h = _capi.default_handle(0)
    for text, mode in cases:
        _, nodes = parse_gfa(text, build_graph=False, build_matrix=True, return_node_list=True, **mode)
        got = h.fetch_nodes_tsv().tobytes()
        assert got == _ref_node_map_bytes(nodes), (mode, got[:80])
        _, raw = parse_gfa(text, build_graph=False, build_matrix=True, return_node_list=True, raw_bytes_id=True, **mode)
        assert h.fetch_nodes_tsv().tobytes() == _ref_node_map_bytes(raw)


@pytest.mark.gpu
def test_cli_writes_reference_node_map_and_loadable_npz(tmp_path):
    import parity_util as pu
    from gfa2network_b200.cli import main

    out = tmp_path / "drb1.npz"
    main(["convert", str(pu.GOLD / "DRB1-3123_unsorted.gfa"), "--matrix", str(out), "--matrix-format", "csr"])
    tsv = (tmp_path / "drb1.npz.nodes.tsv").read_bytes()
    assert hashlib.sha256(tsv).hexdigest().startswith("5642957c2c7df506")  # recorded from the reference CLI
    z = np.load(out)
    assert sorted(z.files) == ["data", "format", "indices", "indptr", "shape"] and z["format"].item() == b"csr"
