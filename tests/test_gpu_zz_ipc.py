"""Multi-GPU path with REAL process boundaries: two ranks in two processes, exchange arenas mapped through
CUDA IPC, kernels of one process writing into the other's memory and waiting on its flags.  Uses one
GPU per rank when the box has two, else both ranks share cuda:0 (IPC and the signalling protocol are the
same; the processes are time-sliced).  Bootstrap over gloo so that it also runs on a single GPU."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from gfa2network_b200 import dist as D
        from gfa2network_b200.synth import synth_gfa
        from oracle.oracle import oracle_convert_format, oracle_parse_gfa

        dev = rank % torch.cuda.device_count()
        torch.cuda.set_device(dev)
        text = synth_gfa(40_000, 120_000, seed=12, kind=1, interleave=4096)
        tb = text.tobytes()
        lo, hi = D.shard_range(len(tb), rank, world, lambda p: tb.find(b"\n", p))
        shard = torch.from_numpy(text[lo:hi].copy()).cuda()
        b = D.DistBuilder(dev)
        ok = True
        for mode in (dict(), dict(directed=False), dict(bidirected=True)):
            B, onodes = oracle_parse_gfa(text, return_node_list=True, **mode)
            B = oracle_convert_format(B, "csr")
            spec = []
            for _ in range(3):  # host-planned (a new mode), then twice without any host round trip
                res = b.build(shard, **mode)
                spec.append(res.info["speculative"])
                A = b.gather_matrix(res)
                nodes = b.node_list()
                ok = ok and A.shape == B.shape and np.array_equal(A.indptr, B.indptr) and np.array_equal(A.indices, B.indices)
                ok = ok and A.data.tobytes() == B.data.tobytes() and nodes == onodes
            ok = ok and spec == [False, True, True]
        q.put((rank, bool(ok)))
    except Exception as e:  # noqa: BLE001 - reported to the parent
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_two_processes_over_cuda_ipc():
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=500) for _ in procs)
    for p in procs:
        p.join(60)
    assert res == [(0, True), (1, True)], res
