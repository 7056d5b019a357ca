#!/usr/bin/env python
"""Print the interesting parts of a bench.py JSON line.   python tools/show_bench.py <file.json> [...]"""
import json
import sys

for f in sys.argv[1:]:
    d = json.loads(open(f).read().strip().splitlines()[-1])
    print(f"{f}: {d['ms_per_step']:.4f} ms/step  {d['value']:.1f} GB/s  n_gpus={d['n_gpus']}  e2e {d['e2e']['ms_per_step']:.3f} ms = {d['e2e']['value']:.2f} GB/s")
    ks = d.get("kernels", {})
    print("  kernels:", {k: round(v["ms_per_step"], 4) for k, v in ks.items()}, " sum=%.4f" % sum(v["ms_per_step"] for v in ks.values()))
    print("  kernel_timing:", d.get("kernel_timing"))
    r = d.get("roofline") or {}
    print("  roofline:", {k: (round(v, 4) if isinstance(v, float) else v) for k, v in r.items()})
    print("  path_roofline frac: %.4f" % d["path_roofline"]["frac"], " clocks:", d.get("clocks"), " launches:", d.get("gpu_launches"))
    print("  cpu:", d.get("cpu_baseline"))
