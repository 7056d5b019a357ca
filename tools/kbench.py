#!/usr/bin/env python
"""Kernel-variant timing: per-kernel CUDA-event times of the device-resident build for one configuration.
    G2N_LIB=build/var/libg2n_X.so python tools/kbench.py C5 0.05 [steps]
(measurement scaffolding; bench.py is the benchmark)"""
import json
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch  # noqa: E402

from bench import make_text  # noqa: E402
from gfa2network_b200 import _capi  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "C2"
scale = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
if name == "custom":  # custom n_seg n_link seq_mean [steps]: GFA-1 text of that shape, default directed CSR
    from gfa2network_b200.synth import synth_gfa

    n_seg, n_link, seq_mean = int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
    steps = int(sys.argv[5]) if len(sys.argv) > 5 else 5
    text = synth_gfa(n_seg, n_link, seed=5, kind=1, seq_mean=seq_mean)
    cfg = dict(mode=dict(), fmt="csr")
    scale = 0.0
else:
    cfg, text, _, _ = make_text(name, scale)
t = torch.from_numpy(text).cuda()
mode = cfg["mode"]
h = _capi.Handle(0)
h.set_stream(torch.cuda.current_stream().cuda_stream)
want = {"csr": _capi.FMT_CSR, "csc": _capi.FMT_CSC, "coo": _capi.FMT_NATIVE}[cfg["fmt"]]
wt = mode.get("weight_tag")
wtb = wt.encode() if wt else None
p = _capi.Params(int(mode.get("directed", True)), int(mode.get("bidirected", False)), int(mode.get("keep_directed_bidir", False)),
                 int(mode.get("asymmetric", False)), 0, _capi.DTYPES["float64"], want, 1, wtb, len(wtb) if wtb else 0, 0)
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
for _ in range(3):
    flush.zero_()
    h.check(h.build(t.data_ptr(), t.numel(), p))
torch.cuda.synchronize()
ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
for a, b in ev:
    flush.zero_()
    a.record()
    h.check(h.build(t.data_ptr(), t.numel(), p))
    b.record()
torch.cuda.synchronize()
ms = sum(a.elapsed_time(b) for a, b in ev) / steps
h.set_profile(True)
tot = {}
for _ in range(steps):
    flush.zero_()
    h.check(h.build(t.data_ptr(), t.numel(), p))
    torch.cuda.synchronize()
    for k, (m, c) in h.kernel_times().items():
        tot[k] = tot.get(k, 0.0) + m / steps
print(json.dumps({"lib": os.environ.get("G2N_LIB", "default"), "env": {k: v for k, v in os.environ.items() if k.startswith("G2N_DBG")}, "config": name if name != "custom" else " ".join(sys.argv[1:5]), "scale": scale, "text_mb": round(text.size / 1e6, 1),
                  "ms_step": round(ms, 4), "kernels": {k: round(v, 4) for k, v in tot.items()}}))
