import time, sys, ctypes as C
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from gfa2network_b200 import _capi, parse_gfa
from gfa2network_b200.synth import synth_gfa
from gfa2network_b200 import builders
text = synth_gfa(1_000_000, 3_000_000, seed=2)
pinned = torch.empty(text.size, dtype=torch.uint8).pin_memory(); pinned.numpy()[:] = text
host = pinned.numpy()
for _ in range(3): parse_gfa(host, build_graph=False, build_matrix=True, matrix_format="csr", directed=False)
h = _capi.default_handle(0)
params = _capi.Params(0,0,0,0,0,0,_capi.FMT_CSR,0,None,0,0)
def T(): torch.cuda.synchronize(); return time.perf_counter()
for rep in range(3):
    t0=T(); rc = h.build(host.ctypes.data, host.size, params); t1=T()
    d = h.status()
    s,a0,a1,data = h.fetch_matrix(); t2=T()
    A = builders._matrix_from_handle(h); t3=T()
    print(f"build {1e3*(t1-t0):.2f} ms (h2d {d.ms_h2d:.2f} total_dev {d.ms_total:.2f}) fetch {1e3*(t2-t1):.2f} ms  matrix(fetch+wrap) {1e3*(t3-t2):.2f} ms")
t0=T(); A = parse_gfa(host, build_graph=False, build_matrix=True, matrix_format="csr", directed=False); t1=T(); print("parse_gfa total", 1e3*(t1-t0))
import cProfile, pstats
pr = cProfile.Profile(); pr.enable(); A = parse_gfa(host, build_graph=False, build_matrix=True, matrix_format="csr", directed=False); pr.disable()
pstats.Stats(pr).sort_stats('cumtime').print_stats(12)
