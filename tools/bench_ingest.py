"""Wall clock of parse_gfa(path) on the C2 text from the page cache: library-side streaming ingest
(g2n_build_file) against np.fromfile + g2n_build, and the reference-style CPU port.   python tools/bench_ingest.py [C2|C4]"""
import os
import sys
import tempfile
import time

import numpy as np

sys.path.insert(0, os.getcwd())
from gfa2network_b200 import _capi, parse_gfa  # noqa: E402
from gfa2network_b200.synth import CONFIGS, synth_gfa  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "C2"
cfg = CONFIGS[name]
text = synth_gfa(cfg["n_seg"], cfg["n_link"], seed=cfg["seed"], kind=cfg["kind"], n_paths=cfg.get("n_paths", 0), n_walks=cfg.get("n_walks", 0))
d = tempfile.mkdtemp()
path = os.path.join(d, "in.gfa")
text.tofile(path)
mode = {k: v for k, v in cfg["mode"].items()}
for rep in range(3):
    t = time.perf_counter()
    A = parse_gfa(path, build_graph=False, build_matrix=True, matrix_format="csr", **mode)
    t_file = time.perf_counter() - t
    t = time.perf_counter()
    arr = np.fromfile(path, dtype=np.uint8)
    t_read = time.perf_counter() - t
    B = parse_gfa(arr, build_graph=False, build_matrix=True, matrix_format="csr", **mode)
    t_arr = time.perf_counter() - t
    assert A.nnz == B.nnz
    print(f"{name} {text.size/1e6:.0f} MB: parse_gfa(path) {1e3*t_file:.1f} ms = {text.size/t_file/1e9:.2f} GB/s | np.fromfile {1e3*t_read:.1f} ms + parse_gfa(array) = {1e3*t_arr:.1f} ms")
os.remove(path)
