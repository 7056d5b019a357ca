#!/usr/bin/env python
"""Bit-exact comparison of a named configuration at FULL size with the CPU oracle:

    python tools/parity_full.py C4      # --asymmetric: raw COO in emission order
    python tools/parity_full.py C4d     # default directed: `--matrix-format coo` keeps the CSR of max(S, S^T)

Generates the configuration's text, builds it through the public API on the GPU (host text in, host arrays out), runs
the oracle (oracle/: C restatement of the reference's tokenizer / builder + SciPy, pinned against the real reference)
on the same bytes and compares format, shape, dtypes, every array and the node list.  Prints one JSON line."""
import hashlib
import json
import sys
import time
import warnings
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
warnings.simplefilter("ignore")

from bench import make_text  # noqa: E402
from gfa2network_b200 import convert_format, parse_gfa  # noqa: E402
from oracle.oracle import oracle_convert_format, oracle_parse_gfa  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "C4"
scale = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
cfg, text, n_seg, n_link = make_text(name, scale)
mode, fmt = dict(cfg["mode"]), cfg["fmt"]
t = time.perf_counter()
A, nodes = parse_gfa(text, build_graph=False, build_matrix=True, return_node_list=True, **mode)
A = convert_format(A, fmt, _untouched=True)
gpu_s = time.perf_counter() - t
t = time.perf_counter()
B, onodes = oracle_parse_gfa(text, return_node_list=True, **mode)
B = oracle_convert_format(B, fmt)
cpu_s = time.perf_counter() - t


def arrays(M):
    return (M.row, M.col, M.data) if M.format == "coo" else (M.indptr, M.indices, M.data)


def sha(M):
    h = hashlib.sha256()
    for a in arrays(M):
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


same_meta = A.format == B.format and A.shape == B.shape and A.dtype == B.dtype and all(x.dtype == y.dtype for x, y in zip(arrays(A), arrays(B)))
same_arrays = same_meta and all(np.array_equal(x, y) for x, y in zip(arrays(A), arrays(B)))
same_nodes = nodes == onodes
line = dict(tool="parity_full", config=name, scale=scale, mode=mode or "directed (default)", matrix_format=fmt, text_bytes=int(text.size), segments=n_seg, links=n_link,
            result_format=A.format, nodes=int(A.shape[0]), nnz=int(A.nnz), bit_exact=bool(same_arrays and same_nodes), arrays_equal=bool(same_arrays),
            node_list_equal=bool(same_nodes), sha256_gpu=sha(A), sha256_oracle=sha(B), gpu_call_s=round(gpu_s, 3), oracle_s=round(cpu_s, 1))
print(json.dumps(line))
if not line["bit_exact"]:
    raise SystemExit("MISMATCH")
