"""Wall clock of the result writers on the C2 result: scipy.sparse.save_npz / save_node_map (what the reference calls)
against writers.save_npz_parallel and the GPU-made node map.   python tools/bench_writers.py"""
import time, numpy as np, scipy.sparse as sp, os, tempfile, sys
sys.path.insert(0, os.getcwd())
from gfa2network_b200 import parse_gfa, _capi
from gfa2network_b200.synth import synth_gfa, CONFIGS
from gfa2network_b200.writers import save_npz_parallel
from gfa2network_b200.utils import save_node_map
cfg=CONFIGS["C2"]; text=synth_gfa(cfg["n_seg"],cfg["n_link"],seed=cfg["seed"],kind=cfg["kind"])
h=_capi.default_handle(0)
for rep in range(2):
    A,nodes=parse_gfa(text,build_graph=False,build_matrix=True,return_node_list=True,matrix_format="csr",**cfg["mode"])
    d=tempfile.mkdtemp()
    t=time.perf_counter(); sp.save_npz(d+"/ref.npz",A); t_ref=time.perf_counter()-t
    t=time.perf_counter(); save_npz_parallel(d+"/ours.npz",A); t_ours=time.perf_counter()-t
    t=time.perf_counter(); save_node_map(nodes,d+"/ref.tsv"); t_tsv_ref=time.perf_counter()-t
    t=time.perf_counter(); buf=h.fetch_nodes_tsv(); t1=time.perf_counter()-t
    t=time.perf_counter(); open(d+"/ours.tsv","wb").write(memoryview(buf)); t2=time.perf_counter()-t
    assert open(d+"/ref.tsv","rb").read()==open(d+"/ours.tsv","rb").read()
    print("C2 writers: npz scipy %.3fs  parallel %.3fs (%d cores) | nodes.tsv python %.3fs  gpu fetch %.4fs + write %.4fs | sizes %d %d"%(t_ref,t_ours,os.cpu_count(),t_tsv_ref,t1,t2,os.path.getsize(d+"/ref.npz"),os.path.getsize(d+"/ours.npz")))
    print(h.kernel_times())
