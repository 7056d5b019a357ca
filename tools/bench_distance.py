#!/usr/bin/env python
"""Path-to-path distance matrix on a C4-shaped synthetic GFA: segments + links + 25 P lines
(and 25 W lines, which the reference skips) that each walk every segment.

    python tools/bench_distance.py [--scale 0.25] [--oracle-scale 0.01]

Times `gfa2network_b200.analysis.genome_distance_matrix` end to end (file in the page cache -> matrix) and its
device phases, and -- at a scale the CPU finishes -- the oracle restatement (SciPy csgraph) next to it with an
exact comparison of the two matrices.  Prints one JSON line."""
import argparse
import ctypes as C
import json
import sys
import tempfile
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))


def main():
    import bench
    from gfa2network_b200 import _capi
    from gfa2network_b200.analysis import genome_distance_matrix

    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=float, default=0.25)
    ap.add_argument("--oracle-scale", type=float, default=0.01)
    args = ap.parse_args()
    out = {"tool": "bench_distance", "config": "C4 shape"}
    tmp = Path(tempfile.mkdtemp())
    for label, scale, with_oracle in (("small", args.oracle_scale, True), ("large", args.scale, False)):
        if scale <= 0:
            continue
        cfg, text, n_seg, n_link = bench.make_text("C4", scale)
        p = tmp / f"{label}.gfa"
        text.tofile(p)
        res = {"segments": n_seg, "links": n_link, "text_bytes": int(text.size)}
        import warnings

        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            genome_distance_matrix(str(p))  # warm (device scratch, pinned pools)
            t0 = time.perf_counter()
            M = genome_distance_matrix(str(p))
            res["gpu_s"] = time.perf_counter() - t0
            h = _capi.default_handle(0)
            h.set_profile(True)
            genome_distance_matrix(str(p))
            res["kernels_ms"] = {k: round(v[0], 3) for k, v in h.kernel_times().items()}
            h.set_profile(False)
        res["paths"] = int(M.shape[0])
        res["matrix_max"] = float(np.nanmax(np.where(np.isfinite(M.values), M.values, np.nan))) if M.size else None
        if with_oracle:
            from oracle.oracle import oracle_distance_matrix

            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                t0 = time.perf_counter()
                names, W = oracle_distance_matrix(text)
                res["cpu_oracle_s"] = time.perf_counter() - t0
            res["identical_to_oracle"] = bool(list(M.index) == names and np.array_equal(M.values, W))
        out[label] = res
        del text
    print(json.dumps(out))


if __name__ == "__main__":
    main()
