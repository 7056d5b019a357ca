"""Inputs and modes shared by tools/gen_golden.py (which runs the real reference on them) and
the parity tests (which run the oracle / the CUDA path on the same bytes).

The literals follow the reference's own tests (tests/test_parser.py:11,66,
tests/test_bidirected.py:5, tests/test_split_alignment.py:5-7, tests/test_limits.py:7-15)
and the quirk list of SURVEY.md section 8(a) (Q1-Q12)."""
from __future__ import annotations

import random

T = b"\t"


def L(*f) -> bytes:
    return T.join(x if isinstance(x, bytes) else str(x).encode() for x in f) + b"\n"


def chain(n: int) -> bytes:
    # tests/test_limits.py:7-15
    return b"".join(L("S", i, "*") for i in range(n)) + b"".join(L("L", i, "+", i + 1, "+", "0M") for i in range(n - 1))


def hub(n: int) -> bytes:
    # one row with > 16 stored entries incl. duplicate columns (integer weights)
    out = [L("S", "hub", "*")]
    for i in range(n):
        out.append(L("L", "hub", "+", "n%d" % (i % (n // 2)), "-", "0M", "RC:i:%d" % (i + 1)))
    for i in range(n):
        out.append(L("L", "n%d" % (i % 7), "+", "hub", "+", "0M"))
    return b"".join(out)


LONG_A = b"EDGE_1234_length_5678_cov_12.345"
LONG_B = b"EDGE_1234_length_5678_cov_12.346"
LONG_C = b"contig_with_a_really_long_name_that_exceeds_sixty_four_bytes_in_total_length_0001"

LITERAL_CASES: list[tuple[str, bytes]] = [
    ("sample", b"S\ts1\tACGT\nS\ts2\tTTTT\nL\ts1\t+\ts2\t-\t0M\nP\tp1\ts1+,s2-\t*\n"),
    ("tag_parsing", b"S\ts1\t4\tRC:i:5\nS\ts2\t4\t\nL\ts1\t+\ts2\t+\t0M\tRC:i:2\n"),
    ("bidirected_sample", b"S\ts1\t4\nS\ts2\t4\nL\ts1\t+\ts2\t-\t0M\n"),
    ("e_coord", b"S\ts1\t6\nS\ts2\t10\nE\t*\ts1+\t0\t6\ts2+\t0\t6\t6M\n"),
    ("e_orient", b"S\ts1\t6\nS\ts2\t10\nE\t*\ts1\t+\ts2\t+\n"),
    ("e_and_l", b"S\ts1\t6\nS\ts2\t10\nL\ts1\t+\ts2\t-\t0M\nE\t*\ts1+\t0\t3\ts2+\t0\t3\t3M\n"),
    ("empty", b""),
    ("header_only", b"H\tVN:Z:1.0\n"),
    ("s_only", b"S\ta\t*\nS\tb\t*\nS\ta\t*\n"),
    ("no_trailing_newline", b"S\ta\t*\nL\ta\t+\tb\t-\t0M"),
    ("only_newlines_then_s", b"\n\nS\ta\t*\n"),
    ("crlf", b"S\tx\r\nS\ty\t*\r\nL\tx\t+\ty\t+\r\nL\ty\t-\tx\t-\t0M\r\n"),
    ("compact_l", b"L\ta+\tb-\t0M\tRC:i:2\nL\tb\tc+-\t*\tzz\nL\tc+-\ta\t*\t\n"),
    ("malformed_l", b"S\ta\t*\nL\ta+\tb-\t0M\n"),
    ("malformed_l_short", b"L\ta\n"),
    ("s_no_id", b"S\ta\t*\nS\n"),
    ("malformed_p", b"S\ta\t*\nP\tonly\n"),
    ("malformed_o", b"O\tx\nS\ta\n"),
    ("malformed_e", b"E\t*\ta\t+\tb\n"),
    ("malformed_c", b"C\ta\t+\tb\n"),
    ("compact_empty_u", b"L\t\tb+\t0M\tx\n"),
    ("compact_empty_v", b"L\ta+\t\t0M\tx\n"),
    ("warn_then_error", b"W\tz\nS\ta\t*\nP\tonly\n"),
    ("error_then_warn", b"S\ta\t*\nP\tonly\nW\tz\n"),
    ("unknown_records", b"H\tVN:Z:1.0\nF\tx\n#c\nW\ts\t1\tc\t0\t5\t>a<b\n\ns\tlow\nU\tu\ta b\nS\ta\t*\nL\ta\t+\tb\t+\t0M\n"),
    ("first_field_longer", b"Sx\tfoo\nS\ta\t*\nLL\ta\t+\tb\t+\t0M\nL\ta\t+\tc\t+\t0M\nL\n"),
    ("first_appearance", b"L\tb\t+\ta\t+\t0M\nS\ta\t*\nS\tb\t*\nS\tc\t*\nL\tc\t-\td\t+\t0M\n"),
    ("duplicates", b"S\ta\t*\nS\tb\t*\nL\ta\t+\tb\t+\t0M\nL\ta\t-\tb\t-\t0M\nL\tb\t+\ta\t+\t0M\nL\ta\t+\ta\t+\t0M\nL\ta\t+\tb\t+\t0M\n"),
    ("self_loops", b"L\ta\t+\ta\t-\t0M\nL\ta\t+\ta\t+\t0M\nL\tb\t+\tb\t+\t0M\n"),
    ("weights_basic", b"L\ta\t+\tb\t+\t0M\tRC:i:5\nL\tb\t+\tc\t+\t0M\tRC:f:2.5\nL\tc\t+\ta\t+\t0M\tXX:i:7\nL\ta\t+\tb\t+\t0M\tRC:i:3\n"
                      b"L\tc\t+\tb\t+\t0M\tRC:i:2\tRC:i:9\nL\td\t+\ta\t+\t0M\tRC:i:4\tRC:Z:str\nL\td\t+\tb\t+\t0M\tRC:i:4\tRC:i:abc\n"
                      b"L\td\t+\tc\t+\t0M\tRC:i:4\tRC:B:1,2\nL\te\t+\ta\t+\t0M\tRC:f:7\tRC::5\nL\te\t+\tb\t+\t0M\tRC:i\nL\te\t+\tc\t+\t0M\tRCX:i:3\n"
                      b"L\te\t+\td\t+\t0M\tRC:ii:3\tRC:f:1.25\n"),
    ("weights_lenient", b"L\ta\t+\tb\t+\t0M\tRC:i: 12 \nL\ta\t+\tc\t+\t0M\tRC:i:1_0\nL\ta\t+\td\t+\t0M\tRC:f:1_0.5\nL\ta\t+\te\t+\t0M\tRC:i:0x10\n"
                        b"L\ta\t+\tf\t+\t0M\tRC:i:1.5\nL\ta\t+\tg\t+\t0M\tRC:i:+7\nL\ta\t+\th\t+\t0M\tRC:i:-3\nL\ta\t+\ti\t+\t0M\tRC:i:0\n"
                        b"L\ta\t+\tj\t+\t0M\tRC:f:.5\nL\ta\t+\tk\t+\t0M\tRC:f:5.\nL\ta\t+\tl\t+\t0M\tRC:f:+.5e-3\nL\ta\t+\tm\t+\t0M\tRC:f:1e5\n"
                        b"L\ta\t+\tn\t+\t0M\tRC:f:1E+5\nL\ta\t+\to\t+\t0M\tRC:f:1e1_0\nL\ta\t+\tp\t+\t0M\tRC:f:1_e5\nL\ta\t+\tq\t+\t0M\tRC:f:\x0b2.5\x0c\n"
                        b"L\ta\t+\tr\t+\t0M\tRC:i:1__0\nL\ta\t+\ts\t+\t0M\tRC:i:_1\nL\ta\t+\tt\t+\t0M\tRC:i:1_\nL\ta\t+\tu\t+\t0M\tRC:i:\n"
                        b"L\ta\t+\tv\t+\t0M\tRC:f:\nL\ta\t+\tw\t+\t0M\tRC:f:1.5.2\nL\ta\t+\tx\t+\t0M\tRC:f:--1\nL\ta\t+\ty\t+\t0M\tRC:i:007\n"
                        b"L\ta\t+\tz\t+\t0M\tRC:f:1d5\nL\ta\t+\taa\t+\t0M\tRC:f:0x1p3\nL\ta\t+\tab\t+\t0M\tRC:f:1e\nL\ta\t+\tac\t+\t0M\tRC:f:-0.0\n"),
    ("weights_rounding", b"L\ta\t+\tb\t+\t0M\tRC:i:9007199254740993\nL\ta\t+\tc\t+\t0M\tRC:i:9007199254740995\n"
                         b"L\ta\t+\td\t+\t0M\tRC:f:0.1000000000000000055511151231257827021181583404541015625\n"
                         b"L\ta\t+\te\t+\t0M\tRC:f:1e23\nL\ta\t+\tf\t+\t0M\tRC:f:8.5e-324\nL\ta\t+\tg\t+\t0M\tRC:f:2.2250738585072011e-308\n"
                         b"L\ta\t+\th\t+\t0M\tRC:f:1.7976931348623157e308\nL\ta\t+\ti\t+\t0M\tRC:f:1e400\nL\ta\t+\tj\t+\t0M\tRC:f:inf\n"
                         b"L\ta\t+\tk\t+\t0M\tRC:f:-Infinity\nL\ta\t+\tl\t+\t0M\tRC:f:123456789012345678901234567890\n"
                         b"L\ta\t+\tm\t+\t0M\tRC:f:4.9406564584124654e-324\nL\ta\t+\tn\t+\t0M\tRC:f:2.4703282292062327e-324\n"
                         b"L\ta\t+\to\t+\t0M\tRC:f:2.4703282292062328e-324\nL\ta\t+\tp\t+\t0M\tRC:f:9007199254740993\n"
                         b"L\ta\t+\tq\t+\t0M\tRC:f:1.00000000000000011102230246251565404236316680908203125\n"
                         b"L\ta\t+\tr\t+\t0M\tRC:f:1.00000000000000011102230246251565404236316680908203124\n"
                         b"L\ta\t+\ts\t+\t0M\tRC:f:1.00000000000000011102230246251565404236316680908203126\n"
                         b"L\ta\t+\tt\t+\t0M\tRC:f:3.14159\nL\ta\t+\tu\t+\t0M\tRC:f:0.001\nL\ta\t+\tv\t+\t0M\tRC:f:12345.678\n"
                         b"L\ta\t+\tw\t+\t0M\tRC:f:1e-400\nL\ta\t+\tx\t+\t0M\tRC:i:18446744073709551616\nL\ta\t+\ty\t+\t0M\tRC:f:7.038531e-26\n"
                         b"L\ta\t+\tz\t+\t0M\tRC:f:1e22\nL\ta\t+\taa\t+\t0M\tRC:f:1e-22\nL\ta\t+\tab\t+\t0M\tRC:f:9e15\n"),
    ("weights_neg_zero", b"L\ta\t+\tb\t+\t0M\tRC:i:-0\nL\ta\t+\tc\t+\t0M\tRC:f:-0\nL\ta\t+\td\t+\t0M\tRC:f:-0.0\nL\ta\t+\te\t+\t0M\tRC:i:-00\n"
                         b"L\ta\t+\tf\t+\t0M\tA:B:i:3\tRC:f:+2.50\nL\ta\t+\tg\t+\t0M\tRC:f:00012.5000\nL\ta\t+\th\t+\t0M\tRC:f:123456789012345.6\n"),
    ("weights_signs", b"L\ta\t+\tb\t+\t0M\tRC:f:-2.5\nL\tb\t+\ta\t+\t0M\tRC:f:1.5\nL\tc\t+\td\t+\t0M\tRC:i:0\nL\td\t+\te\t+\t0M\tRC:i:-4\n"
                      b"L\te\t+\tf\t+\t0M\tRC:i:2\nL\te\t+\tf\t+\t0M\tRC:i:-2\nL\tf\t+\te\t+\t0M\tRC:i:-7\n"),
    ("weight_overflow", b"S\ta\t*\nL\ta\t+\tb\t+\t0M\tRC:i:" + b"9" * 400 + b"\n"),
    ("weight_overflow_overwritten", b"L\ta\t+\tb\t+\t0M\tRC:i:" + b"9" * 400 + b"\tRC:i:5\nL\ta\t+\tc\t+\t0M\tRC:i:-" + b"9" * 400 + b"\tRC:Z:x\n"),
    ("weight_int_digit_limit", b"L\ta\t+\tb\t+\t0M\tRC:i:3\tRC:i:" + b"1" * 4301 + b"\n"),
    ("weight_float_long", b"L\ta\t+\tb\t+\t0M\tRC:f:" + b"1" * 400 + b"\nL\ta\t+\tc\t+\t0M\tRC:f:0." + b"0" * 400 + b"7\n"),
    ("weight_bad_utf8_tag", b"L\ta\t+\tb\t+\t0M\tRC:i:5\tRC:i:\xff9\nL\ta\t+\tc\t+\t0M\tR\xc3\xa9:i:5\tRC:f:3\n"),
    ("e_int_lenient", b"E\t*\ta+\t 0 \t+6\tb-\t0\t1_0\t6M\tRC:f:2.5\nE\t*\ta+\t0\t15$\tb+\t0\t6\t6M\nE\t*\tc-\t0\t1\td\t0x1\t2\t*\n"
                      b"E\te1\tx+-\t1\t2\ty-+\t3\t4\t*\tRC:i:6\tRC:i:8\n"),
    ("c_records", b"C\ta\t+\tb\t-\t10\t5M\nC\t*\ta+\t0\t6\tb-\t0\t6\t6M\tRC:i:3\nC\tp\tq\tr\ts\tRC:i:9\n"),
    ("strip_quirk", b"S\ta+\t*\nL\ta+\t+\tb\t+\t0M\nL\tb-\t+\ta\t-\t0M\n"),
    ("long_names", L("S", LONG_A, "*") + L("S", LONG_B, "*") + L("L", LONG_A, "+", LONG_B, "-", "0M") + L("L", LONG_C, "+", LONG_A, "-", "0M")
                   + L("L", LONG_B, "-", LONG_C, "+", "0M") + L("S", b"exactly15bytes_", "*") + L("S", b"exactly16bytes__", "*")
                   + L("S", b"exactly14bytes", "*") + L("S", b"exactly13byte", "*") + L("L", b"exactly13byte", "+", b"exactly14bytes", "-", "0M")
                   + L("L", b"exactly15bytes_", "+", b"exactly16bytes__", "-", "0M")),
    ("odd_names", b"S\t\t*\nS\tna me\t*\nS\tn:a\t*\nS\t\xc3\xa9t\xc3\xa9\t*\nL\t\t+\tna me\t-\t0M\nL\tn:a\t+\t\xc3\xa9t\xc3\xa9\t+\t0M\nL\tn:a:+\t+\tn\t-\t0M\n"),
    ("odd_orientations", b"L\ta\t+\tb\tX\t0M\nL\ta\t+\tb\t\t0M\nL\ta\t-\tb\t++\t0M\nE\t*\ta\tfwd\tb\trev\nC\ta\t?\tb\t+\t0\t1M\n"),
    ("chain400", chain(400)),
    ("hub40", hub(40)),
    ("many_tags", L("L", "a", "+", "b", "+", "0M", *["T%d:i:%d" % (i, i) for i in range(30)], "RC:i:77", *["U%d:Z:x" % i for i in range(30)])),
    ("long_seq_lines", L("S", "a", b"ACGT" * 3000) + L("S", "b", b"G" * 70000, "LN:i:70000") + L("L", "a", "+", "b", "+", "0M")
                       + L("P", "p", b",".join(b"a+" for _ in range(20000)), "*") + L("W", "smp", 1, "chr", 0, 9, b">a<b" * 9000) + L("L", "b", "-", "a", "-", "0M")),
]

MODES: list[dict] = [
    {},
    {"asymmetric": True},
    {"directed": False},
    {"bidirected": True},
    {"bidirected": True, "keep_directed_bidir": True},
    {"bidirected": True, "keep_directed_bidir": True, "asymmetric": True},
    {"strip_orientation": True},
    {"strip_orientation": True, "bidirected": True},
    {"weight_tag": "RC"},
    {"weight_tag": "RC", "asymmetric": True},
    {"weight_tag": "RC", "directed": False},
    {"weight_tag": "RC", "bidirected": True},
    {"weight_tag": "RC", "dtype": "float32"},
    {"weight_tag": "RC", "dtype": "int32", "directed": False},
    {"weight_tag": "RC", "dtype": "int8", "asymmetric": True},
    {"dtype": "bool"},
    {"dtype": "bool", "directed": False},
    {"dtype": "int32", "directed": False},
    {"weight_tag": "A:B", "asymmetric": True},
]

# ---------------------------------------------------------------------------- fuzz inputs
FUZZ_SEEDS = [11, 12, 13, 14]
FUZZ_MODES: list[dict] = [
    {},
    {"directed": False},
    {"bidirected": True},
    {"bidirected": True, "keep_directed_bidir": True},
    {"weight_tag": "RC", "asymmetric": True},
    {"weight_tag": "RC"},
    {"weight_tag": "RC", "directed": False, "dtype": "float32"},
    {"weight_tag": "RC", "bidirected": True, "strip_orientation": True},
]


def fuzz_text(seed: int, n_lines: int = 12000) -> bytes:
    """Deterministic (random.Random) mix of every record shape the tokenizer accepts, with a
    name pool small enough that duplicates, reverse duplicates and self loops are common.
    Weights are k/8 (exact in binary) so duplicate sums do not depend on summation order."""
    r = random.Random(seed)
    pool = []
    for i in range(r.choice([50, 700, 3000])):
        kind = r.random()
        if kind < 0.6:
            pool.append(b"s%d" % r.randrange(10 ** r.randrange(1, 9)))
        elif kind < 0.8:
            pool.append(b"utg%06dl" % i)
        elif kind < 0.9:
            pool.append(b"EDGE_%d_length_%d_cov_%d.%d" % (i, r.randrange(10 ** 6), r.randrange(100), r.randrange(1000)))
        else:
            pool.append(bytes(r.choice(b"abcXYZ_.|:") for _ in range(r.randrange(1, 24))))
    ori = [b"+", b"-"]

    def tags() -> list[bytes]:
        t = []
        for _ in range(r.choice([0, 0, 1, 1, 2, 4])):
            c = r.random()
            if c < 0.35:
                t.append(b"RC:i:%d" % r.randrange(-3, 60))
            elif c < 0.6:
                t.append(b"RC:f:%s" % repr(r.randrange(-16, 800) / 8).encode())
            elif c < 0.7:
                t.append(b"RC:Z:abc")
            elif c < 0.8:
                t.append(b"RC:i:x%d" % r.randrange(9))
            else:
                t.append(b"FC:i:%d" % r.randrange(100))
        return t

    out = [b"H\tVN:Z:1.0\n"]
    for _ in range(n_lines):
        c = r.random()
        a, b = r.choice(pool), r.choice(pool)
        if r.random() < 0.3:
            j = pool.index(a) if a in pool[:64] else r.randrange(len(pool))
            b = pool[min(len(pool) - 1, j + r.randrange(0, 3))]
        if c < 0.25:
            seq = r.choice([b"*", b"ACGT" * r.randrange(1, 40), str(r.randrange(1000)).encode()])
            out.append(L("S", a, seq, *tags()))
        elif c < 0.70:
            out.append(L("L", a, r.choice(ori), b, r.choice(ori), r.choice([b"0M", b"*", b"12M3D"]), *tags()))
        elif c < 0.76:
            out.append(L("L", a + r.choice(ori), b + r.choice([b"+", b"-", b""]), b"0M", *(tags() or [b""])))
        elif c < 0.86:
            out.append(L("E", "*", a + r.choice(ori), 0, r.randrange(100), b + r.choice(ori), 0, r.randrange(100), b"5M", *tags()))
        elif c < 0.89:
            out.append(L("E", "e%d" % r.randrange(99), a, r.choice(ori), b, r.choice(ori), *tags()))
        elif c < 0.93:
            out.append(L("C", a, r.choice(ori), b, r.choice(ori), r.randrange(50), b"4M", *tags()))
        elif c < 0.95:
            out.append(L("C", "*", a + r.choice(ori), 1, 2, b + r.choice(ori), 3, 4, b"1M", *tags()))
        elif c < 0.97:
            out.append(L("P", "p%d" % r.randrange(99), b",".join(r.choice(pool) + r.choice(ori) for _ in range(r.randrange(1, 600))), "*"))
        elif c < 0.98:
            out.append(L("O", "o%d" % r.randrange(99), b",".join(r.choice(pool) + r.choice(ori) for _ in range(r.randrange(1, 50)))))
        elif c < 0.99:
            out.append(L("W", "smp", 1, "chr1", 0, 100, b"".join(r.choice([b">", b"<"]) + r.choice(pool) for _ in range(r.randrange(1, 3000)))))
        else:
            out.append(r.choice([b"\n", b"# comment\n", b"H\tx\n", b"Sx\ty\n", b"F\tfrag\n"]))
    text = b"".join(out)
    if seed % 2 == 0:
        text = text[:-1]  # final line without "\n"
    return text


def hub_text(seed: int = 9) -> bytes:
    """1200 ordinary segments, two hub segments linked to hundreds of them (rows of > 500 stored entries), the same link
    repeated many times (duplicates to sum), integer RC weights: rows far longer than what one lane sorts (rowsort.cuh: RS_SMALL)."""
    import numpy as np

    rng = np.random.default_rng(seed)
    lines = ["H\tVN:Z:1.0"] + [f"S\ts{i}\t*" for i in range(1200)] + ["S\thub\t*", "S\thub2\tACGT"]
    for i in range(700):
        j = int(rng.integers(0, 1200))
        o1, o2 = "+-"[i % 2], "+-"[(i // 2) % 2]
        lines.append(f"L\thub\t{o1}\ts{j}\t{o2}\t0M\tRC:i:{1 + i % 7}")
        if i % 3 == 0:
            lines.append(f"L\ts{j}\t{o2}\thub2\t{o1}\t*\tRC:i:{2 + i % 5}")
        if i % 5 == 0:
            lines.append("L\thub\t+\thub2\t-\t0M\tRC:i:3")  # the same link again and again
    for i in range(3000):
        a, b = rng.integers(0, 1200, 2)
        lines.append(f"L\ts{a}\t+\ts{b}\t-\t0M\tRC:i:{1 + i % 3}")
    return ("\n".join(lines) + "\n").encode()
