"""Inputs of the distance goldens (tools/gen_golden_distance.py runs the real reference on them)."""
from pathlib import Path

import numpy as np

GOLD = Path(__file__).resolve().parent / "golden"


def chain(n: int, paths: list[tuple[str, list[int]]], extra: bytes = b"") -> bytes:
    """tests/bench_distance.py:11-19 style: a chain s0 -> s1 -> ... with P lines over chosen segments."""
    lines = [b"S\ts%d\t*" % i for i in range(n)] + [b"L\ts%d\t+\ts%d\t+\t0M" % (i, i + 1) for i in range(n - 1)]
    for name, segs in paths:
        lines.append(b"P\t" + name.encode() + b"\t" + b",".join(b"s%d+" % i for i in segs) + b"\t*")
    return b"\n".join(lines) + b"\n" + extra


def random_graph(seed: int, n: int, m: int, n_paths: int, plen: int) -> bytes:
    rng = np.random.default_rng(seed)
    lines = [b"S\tn%d\t*" % i for i in range(n)]
    for _ in range(m):
        u, v = rng.integers(0, n, 2)
        lines.append(b"L\tn%d\t%s\tn%d\t%s\t0M" % (u, b"+-"[rng.integers(0, 2):][:1], v, b"+-"[rng.integers(0, 2):][:1]))
    for k in range(n_paths):
        segs = rng.integers(0, n, plen)
        lines.insert(int(rng.integers(0, len(lines))), b"P\tpath%d\t" % k + b",".join(b"n%d%s" % (i, b"+-"[rng.integers(0, 2):][:1]) for i in segs) + b"\t*")
    return b"\n".join(lines) + b"\n"


CASES = [
    # tests/test_distance.py:13
    ("ref_sample", b"S\ts1\t*\nS\ts2\t*\nS\ts3\t*\nL\ts1\t+\ts2\t+\t0M\nL\ts2\t+\ts3\t+\t0M\nP\tp1\ts1+,s2+\t*\nP\tp2\ts3+,s2+\t*\n"),
    ("chain_directed", chain(400, [("head", [0, 1, 2]), ("mid", [200, 201]), ("tail", [398, 399]), ("one", [57])])),
    ("deep_chain", chain(5000, [("a", [0]), ("b", [4999]), ("c", [2500, 10])])),
    ("unreachable", b"S\ta\t*\nS\tb\t*\nS\tc\t*\nS\td\t*\nL\ta\t+\tb\t+\t0M\nL\tc\t+\td\t+\t0M\nP\tp1\ta+\t*\nP\tp2\td+\t*\nP\tp3\tb-,c\t*\n"),
    ("o_records_and_overwrite", b"S\ta\t*\nS\tb\t*\nS\tc\t*\nL\ta\t+\tb\t+\t0M\nL\tb\t+\tc\t-\t0M\nO\tw1\ta+,b-\nP\tp\ta+\t*\nP\tp\tc+\t*\nO\tw2\tc\textra\n"),
    ("duplicates_in_lists", b"S\ta\t*\nS\tb\t*\nS\tc\t*\nL\ta\t+\tb\t+\t0M\nL\ta\t+\tb\t-\t0M\nL\tb\t+\tc\t+\t0M\nP\tp1\ta+,a-,a+\t*\nP\tp2\tc+,c+,b+\t*\n"),
    ("missing_node", b"S\ta\t*\nS\tb\t*\nL\ta\t+\tb\t+\t0M\nP\tp1\ta+,zzz+\t*\nP\tp2\tb+\t*\n"),
    ("no_paths", b"S\ta\t*\nS\tb\t*\nL\ta\t+\tb\t+\t0M\n"),
    ("malformed_after_paths", b"S\ta\t*\nP\tp1\ta+\t*\nL\ta\t+\n"),
    ("unknown_record_warning", b"S\ta\t*\nS\tb\t*\nW\tx\t1\tchr\t0\t5\t>a>b\nL\ta\t+\tb\t+\t0M\nP\tp1\ta+\t*\nP\tp2\tb+\t*\n"),
    ("random_60", random_graph(5, 60, 150, 6, 5)),
    ("random_2000", random_graph(6, 2000, 5000, 8, 12)),
    ("drb1", (GOLD / "DRB1-3123_unsorted.gfa").read_bytes()),
]
