"""API-surface parity of the host mirror on a GPU: sources (path, .gz, file object, bytes), verbose
strings, return conventions, warm-handle reuse, error/warning fuzz against the oracle, float-tag
tolerance."""
import gzip
import io
import random
import warnings

import numpy as np
import pytest

import golden_inputs as gi
import parity_util as pu

pytestmark = pytest.mark.gpu

SAMPLE = b"S\ts1\tACGT\nS\ts2\tTTTT\nL\ts1\t+\ts2\t-\t0M\nP\tp1\ts1+,s2-\t*\n"


def _same(A, B, what=""):
    assert A.format == B.format and A.dtype == B.dtype and A.shape == B.shape, what
    for x, y in zip(pu.arrays_of(A), pu.arrays_of(B)):
        assert x.dtype == y.dtype and x.shape == y.shape and x.tobytes() == y.tobytes(), what


def test_sources_and_return_conventions(tmp_path):
    from gfa2network_b200 import parse_gfa
    from oracle.oracle import oracle_parse_gfa

    want, wnodes = oracle_parse_gfa(SAMPLE, return_node_list=True)
    plain = tmp_path / "t.gfa"
    plain.write_bytes(SAMPLE)
    gz = tmp_path / "t.gfa.gz"
    gz.write_bytes(gzip.compress(SAMPLE))  # tests/test_parser.py:98-104
    for src in (plain, str(plain), gz, io.BytesIO(SAMPLE), SAMPLE, bytearray(SAMPLE), np.frombuffer(SAMPLE, np.uint8)):
        A, nodes = parse_gfa(src, build_graph=False, build_matrix=True, return_node_list=True)
        _same(A, want, str(type(src)))
        assert nodes == wnodes == ["s1", "s2"]
    A = parse_gfa(plain, build_graph=False, build_matrix=True)  # matrix only (builders.py:296-299)
    _same(A, want)
    _, raw = parse_gfa(plain, build_graph=False, build_matrix=True, return_node_list=True, raw_bytes_id=True)
    assert raw == [b"s1", b"s2"]  # tests/test_parser.py:107-110
    assert parse_gfa(plain, build_graph=False, build_matrix=False) is None


def test_device_tensor_source():
    import torch

    from gfa2network_b200 import parse_gfa
    from oracle.oracle import oracle_parse_gfa

    t = torch.from_numpy(np.frombuffer(SAMPLE, np.uint8).copy()).cuda()
    _same(parse_gfa(t, build_graph=False, build_matrix=True), oracle_parse_gfa(SAMPLE))


def test_verbose_strings(capsys):
    from gfa2network_b200 import parse_gfa
    from gfa2network_b200.synth import synth_gfa

    text = synth_gfa(300_000, 800_000, seed=1)
    parse_gfa(text, build_graph=False, build_matrix=True, verbose=True)
    out = capsys.readouterr()
    assert "[parse_gfa] done" in out.out  # builders.py:261, tests/test_large_graph.py
    assert "\r[500,000 lines]" in out.err and "\r[1,000,000 lines]" in out.err  # builders.py:257-258
    assert "1,500,000" not in out.err


def test_warm_handle_reuse_across_shapes():
    """Same handle, alternating large / tiny / weighted / bidirected inputs: no stale state."""
    from gfa2network_b200 import convert_format, parse_gfa
    from gfa2network_b200.synth import synth_gfa
    from oracle.oracle import oracle_convert_format, oracle_parse_gfa

    big = synth_gfa(60_000, 200_000, seed=21)
    wtd = synth_gfa(5_000, 20_000, seed=22, kind=2)
    seq = [(big, dict()), (SAMPLE, dict(bidirected=True)), (wtd, dict(weight_tag="RC", directed=False)), (b"", dict()),
           (big, dict(directed=False)), (SAMPLE, dict(asymmetric=True)), (wtd, dict(weight_tag="RC")), (big, dict(bidirected=True))]
    for text, mode in seq * 2:
        A, nodes = parse_gfa(text, build_graph=False, build_matrix=True, return_node_list=True, **mode)
        B, onodes = oracle_parse_gfa(text, return_node_list=True, **mode)
        _same(A, B, str(mode))
        assert nodes == onodes
        _same(convert_format(A, "csc"), oracle_convert_format(B, "csc"), str(mode))


def test_convert_format_honours_the_matrix_it_is_given():
    """convert_format must convert what A holds NOW (utils.py:55 A.asformat), not a device result that a
    later build, export or convert has replaced, and not the build's triplets when A was edited in place."""
    from gfa2network_b200 import convert_format, parse_gfa
    from gfa2network_b200.export import edge_list_bytes
    from gfa2network_b200.synth import synth_gfa

    a = synth_gfa(3_000, 9_000, seed=31)
    b = synth_gfa(2_000, 5_000, seed=32)
    for mode in (dict(directed=False), dict(asymmetric=True), dict()):
        A = parse_gfa(a, build_graph=False, build_matrix=True, **mode)
        want = A.copy().asformat("csr")
        edge_list_bytes(b)  # rebuilds the shared default handle with another text
        _same(convert_format(A, "csr"), want, f"after export {mode}")
        A = parse_gfa(a, build_graph=False, build_matrix=True, **mode)
        parse_gfa(b, build_graph=False, build_matrix=False)  # build_matrix=False also replaces the device result
        _same(convert_format(A, "csc"), A.copy().asformat("csc"), f"after parse {mode}")
        A = parse_gfa(a, build_graph=False, build_matrix=True, **mode)
        convert_format(A, "csr")  # (the upload replaces the handle's resident result)
        _same(convert_format(A, "csc"), A.copy().asformat("csc"), f"second convert {mode}")
        # in-place edits between parse and convert are part of the matrix
        A = parse_gfa(a, build_graph=False, build_matrix=True, **mode)
        A.data[:] = 3.0
        A.data[::7] = 0.5
        _same(convert_format(A, "csr"), A.copy().asformat("csr"), f"edited {mode}")
    with pytest.raises(Exception):
        parse_gfa(b"L\tonly\n", build_graph=False, build_matrix=True)
    A = parse_gfa(a, build_graph=False, build_matrix=True, directed=False)
    with pytest.raises(Exception):
        parse_gfa(b"L\tonly\n", build_graph=False, build_matrix=True)  # a failed parse leaves no result behind
    _same(convert_format(A, "csr"), A.copy().asformat("csr"), "after a parse error")


def _bgzf(data: bytes, block: int = 0xFF00) -> bytes:
    """A BGZF file (bgzip's container: independent deflate blocks, size announced in the 'BC' extra field) + EOF block."""
    import struct
    import zlib

    out = []
    for a in list(range(0, len(data), block)) + [None]:
        chunk = data[a:a + block] if a is not None else b""
        c = zlib.compressobj(6, zlib.DEFLATED, -15)
        body = c.compress(chunk) + c.flush()
        bsize = 12 + 6 + len(body) + 8
        out.append(b"\x1f\x8b\x08\x04" + b"\0" * 4 + b"\x00\xff" + struct.pack("<H", 6) + b"BC" + struct.pack("<HH", 2, bsize - 1) + body
                   + struct.pack("<II", zlib.crc32(chunk), len(chunk)))
    return b"".join(out)


def test_gz_sources_inflated_by_the_library(tmp_path):
    """*.gz paths (parser.py:108-109, tests/test_parser.py:98-104): plain gzip, several members, BGZF (block-parallel),
    a stream larger than one 64 MiB window; damaged containers raise what gzip.open().read() raises."""
    import gzip

    from gfa2network_b200 import parse_gfa
    from gfa2network_b200.export import edge_list_bytes
    from gfa2network_b200.synth import synth_gfa
    from oracle.oracle import oracle_parse_gfa

    small = synth_gfa(3_000, 9_000, seed=41).tobytes()
    big = synth_gfa(700_000, 2_100_000, seed=42, seq_mean=40).tobytes()  # > 64 MiB inflated
    assert len(big) > (64 << 20)
    cases = {
        "plain.gfa.gz": (gzip.compress(small, 6), small),
        "members.gfa.gz": (gzip.compress(small[: len(small) // 2], 1) + gzip.compress(small[len(small) // 2:], 9), small),
        "bgzf.gfa.gz": (_bgzf(small), small),
        "bgzf_big.gfa.gz": (_bgzf(big), big),
        "plain_big.gfa.gz": (gzip.compress(big, 1), big),
        "empty_member.gfa.gz": (gzip.compress(b""), b""),
    }
    for name, (blob, text) in cases.items():
        f = tmp_path / name
        f.write_bytes(blob)
        assert gzip.open(f).read() == text
        A, nodes = parse_gfa(str(f), build_graph=False, build_matrix=True, return_node_list=True)
        B, onodes = oracle_parse_gfa(text, return_node_list=True)
        _same(A, B, name)
        assert nodes == onodes, name
    el, exc, _ = edge_list_bytes(str(tmp_path / "bgzf.gfa.gz"))
    assert exc is None and el.tobytes() == edge_list_bytes(small)[0].tobytes()
    # damaged containers: the exception of the reference's own gzip.open(...).read()
    good = cases["plain.gfa.gz"][0]
    bz = cases["bgzf.gfa.gz"][0]
    bad = {"trunc.gfa.gz": good[: len(good) // 2], "notgz.gfa.gz": small[:5000], "crc.gfa.gz": good[:-8] + b"\0\0\0\0" + good[-4:],
           "bgzf_crc.gfa.gz": bz[:200] + bytes([bz[200] ^ 0x55]) + bz[201:], "zero.gfa.gz": b""}
    for name, blob in bad.items():
        f = tmp_path / name
        f.write_bytes(blob)
        try:
            want_text = gzip.open(f).read()
            want_exc = None
        except Exception as e:  # noqa: BLE001
            want_text, want_exc = None, e
        if want_exc is None:
            A = parse_gfa(str(f), build_graph=False, build_matrix=True)
            _same(A, oracle_parse_gfa(want_text), name)
        else:
            with pytest.raises(type(want_exc)):
                parse_gfa(str(f), build_graph=False, build_matrix=True)


BAD_LINES = [b"L\ta\n", b"L\ta+\tb-\t0M\n", b"S\n", b"P\tonly\n", b"O\tx\n", b"E\t*\ta\t+\tb\n", b"C\ta\t+\tb\n", b"L\t\tb+\t0M\tx\n",
             b"L\ta\t+\tb\t+\t0M\tRC:i:" + b"9" * 400 + b"\n", b"W\tw\t1\n", b"# c\n", b"\n", b"x\ty\n", b"L\ta\t+\tb\t\xff\t0M\n"]


@pytest.mark.parametrize("seed", range(12))
def test_error_and_warning_fuzz_vs_oracle(seed):
    """Random malformed / unsupported lines dropped into a valid file: the first error in file order,
    its exception type and message, and the one-shot warning must match the oracle."""
    from gfa2network_b200 import parse_gfa
    from oracle.oracle import oracle_parse_gfa

    r = random.Random(seed)
    lines = gi.fuzz_text(11 + seed % 4, 1500).split(b"\n")
    lines = [ln + b"\n" for ln in lines if ln]
    for _ in range(r.randrange(0, 4)):
        lines.insert(r.randrange(len(lines)), r.choice(BAD_LINES))
    text = b"".join(lines)
    mode = r.choice([dict(), dict(weight_tag="RC"), dict(bidirected=True), dict(directed=False, weight_tag="RC")])

    def run(fn):
        with warnings.catch_warnings(record=True) as w:
            warnings.simplefilter("always")
            try:
                out = fn()
                exc = None
            except Exception as e:  # noqa: BLE001
                out, exc = None, e
        return out, exc, [str(x.message) for x in w if issubclass(x.category, RuntimeWarning)]

    A, e1, w1 = run(lambda: parse_gfa(text, build_graph=False, build_matrix=True, **mode))
    B, e2, w2 = run(lambda: oracle_parse_gfa(text, **mode))
    assert w1 == w2
    assert type(e1) is type(e2) and str(e1) == str(e2), (e1, e2)
    if e1 is None:
        _same(A, B)


def test_inexact_float_weights_within_stated_tolerance():
    """Float-tag weights that are not exactly representable (x = '%.3f' % (k/7)), with duplicate links.
    Structure is bit-exact.  Values: SciPy sums duplicates in emission order for rows of <= 16 raw
    entries (insertion-sort regime of std::sort) -- the device path does the same for every row, so
    those rows must be bit-exact; longer rows are only required to agree within
    |delta| <= 4 * eps * sum|w| per stored entry (SciPy's own order is unspecified there)."""
    from gfa2network_b200 import convert_format, parse_gfa
    from oracle.oracle import oracle_convert_format, oracle_parse_gfa

    r = random.Random(3)
    n = 4000
    lines = [b"S\ts%d\t*\n" % i for i in range(n)]
    for _ in range(20000):
        u = r.randrange(n)
        v = min(n - 1, u + r.randrange(1, 4))
        lines.append(b"L\ts%d\t+\ts%d\t+\t0M\tRC:f:%s\n" % (u, v, ("%.3f" % (r.randrange(1, 8000) / 7)).encode()))
    text = b"".join(lines)
    eps = 2.0 ** -53
    for mode in (dict(weight_tag="RC", asymmetric=True), dict(weight_tag="RC"), dict(weight_tag="RC", directed=False)):
        raw = parse_gfa(text, build_graph=False, build_matrix=True, **mode)
        A = convert_format(raw, "csr")
        Braw = oracle_parse_gfa(text, **mode)
        B = oracle_convert_format(Braw, "csr")
        assert np.array_equal(A.indptr, B.indptr) and np.array_equal(A.indices, B.indices)
        # sum|w| per stored entry is bounded by |value| here (all weights are positive)
        assert np.all(np.abs(A.data - B.data) <= 4 * eps * np.abs(B.data))
        if Braw.format == "coo":
            raw_per_row = np.bincount(Braw.row, minlength=n)
            rows = np.repeat(np.arange(n), np.diff(B.indptr))
            short = raw_per_row[rows] <= 16
            assert short.sum() > 0.9 * len(rows)
            assert np.array_equal(A.data[short], B.data[short])


def test_speculative_builds_hit_and_miss():
    """A repeat build of the same input size and mode skips the host round trip after the tokenizer
    (g2n_set_speculation); a same-sized input that does not fit the learnt sizes, or that holds an
    error record, falls back to the host-checked path.  Results are identical either way."""
    from gfa2network_b200 import _capi, parse_gfa
    from gfa2network_b200.synth import synth_gfa
    from oracle.oracle import oracle_parse_gfa

    h = _capi.default_handle(0)
    a = bytes(synth_gfa(20_000, 60_000, seed=31, kind=1, n_paths=1, n_walks=1))
    line = b"L\ta\t+\tb\t+\n"
    b = (line * (len(a) // len(line) + 1))[: len(a) - len(a) % len(line)]
    b = b + b"#" * (len(a) - len(b) - 1) + b"\n" if len(a) - len(b) >= 1 else b  # same size, 3x the edge records
    assert len(b) == len(a)
    bad = bytearray(a)
    pos = a.index(b"\nL\t", len(a) // 2)  # a link in the middle of the file loses its fields
    end = a.index(b"\n", pos + 1)
    bad[pos + 1:end] = b"L\tx" + b" " * (end - pos - 4)
    bad = bytes(bad)
    for mode in (dict(), dict(directed=False), dict(bidirected=True)):
        want, wnodes = oracle_parse_gfa(a, return_node_list=True, **mode)
        wantb = oracle_parse_gfa(b, **mode)
        flags = []
        for text, exp in ((a, want), (a, want), (a, want), (b, wantb), (b, wantb), (a, want), (a, want)):
            A, nodes = parse_gfa(text, build_graph=False, build_matrix=True, return_node_list=True, **mode)
            flags.append(int(h.status().speculative))
            _same(A, exp, str(mode))
            if text is a:
                assert nodes == wnodes
        assert flags[1:3] == [1, 1] and flags[3] == 0 and flags[4] == 1 and flags[5] == 0 and flags[6] == 1, flags
        with pytest.raises(ValueError, match="Malformed L record"):
            parse_gfa(bad, build_graph=False, build_matrix=True, **mode)
        A = parse_gfa(a, build_graph=False, build_matrix=True, **mode)
        _same(A, want, str(mode))
    h.set_speculation(False)
    try:
        for _ in range(2):
            _same(parse_gfa(a, build_graph=False, build_matrix=True), oracle_parse_gfa(a))
            assert h.status().speculative == 0
    finally:
        h.set_speculation(True)


def test_file_source_streams_in_pieces(tmp_path):
    """A path is read by the library itself (g2n_build_file: reader threads -> pinned staging -> device, the
    tokenizer launched behind every 8 MiB piece): same result as the bytes, also for an error in a late
    piece, an empty file, and when the same handle repeats the build (speculative path)."""
    from gfa2network_b200 import parse_gfa
    from gfa2network_b200.synth import synth_gfa
    from oracle.oracle import oracle_parse_gfa

    text = synth_gfa(400_000, 1_200_000, seed=61, kind=1, n_paths=1, n_walks=1)
    assert text.size > 4 * (8 << 20)  # several pieces, every reader thread gets work
    p = tmp_path / "big.gfa"
    p.write_bytes(text.tobytes())
    want, wnodes = oracle_parse_gfa(text, return_node_list=True)
    for _ in range(3):
        A, nodes = parse_gfa(p, build_graph=False, build_matrix=True, return_node_list=True)
        _same(A, want, "file")
        assert nodes == wnodes
    _same(parse_gfa(str(p), build_graph=False, build_matrix=True, directed=False), oracle_parse_gfa(text, directed=False), "file undirected")
    bad = bytearray(text.tobytes())
    pos = bytes(bad).rindex(b"\nL\t")  # the last link: several pieces into the file
    assert pos > 3 * (8 << 20)
    end = bytes(bad).index(b"\n", pos + 1)
    bad[pos + 1:end] = b"L\tx" + b" " * (end - pos - 4)
    q = tmp_path / "bad.gfa"
    q.write_bytes(bytes(bad))
    with pytest.raises(ValueError, match="Malformed L record"):
        parse_gfa(q, build_graph=False, build_matrix=True)
    e = tmp_path / "empty.gfa"
    e.write_bytes(b"")
    _same(parse_gfa(e, build_graph=False, build_matrix=True), oracle_parse_gfa(b""), "empty file")
    with pytest.raises(FileNotFoundError):
        parse_gfa(tmp_path / "missing.gfa", build_graph=False, build_matrix=True)
