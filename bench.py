#!/usr/bin/env python
"""bench.py -- GFA -> CSR parse+build throughput on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config C2] [--impl ours|reference]

A step = one pass of the hot path (tokenize -> node IDs -> triplets -> sort/dedup -> CSR) over one
synthetic GFA text of the named configuration.  `value` is whole-job input GB/s with
the text already resident in HBM and the CSR left resident in HBM (CUDA events on the launching
stream, L2 flushed between steps, max over ranks); `e2e` is the same metric through the public
Python API with a pinned HOST buffer in and host NumPy arrays out (H2D and D2H inside the timed
region).  `roofline` is for the kernel with the largest share of the step, from per-launch CUDA
events; `cpu_baseline` is the CPU oracle (C port of the reference's tokenizer/builder + SciPy)
timed on this host.  `--impl reference` times that CPU path alone.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

HBM_FALLBACK_GBS = 6650.0  # B200_PROFILING.md fallback, used only if MEASURED_PEAKS.json is absent
# dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the committed `ncu --set full` captures
# (profiles/r2b_ncu_full.md: C2, C3; profiles/r2_ncu_full.md: C5 shape at 5 %), keyed by (config, scale, kernel)
TRAFFIC = {("C2", 1.0, "k_tokenize"): 205_775_616, ("C3", 1.0, "k_tokenize"): 19_774_601_000, ("C5", 0.05, "k_tokenize"): 7_395_001_616}


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        return float(json.loads(p.read_text())["hbm_gbs"]), "measured"
    return HBM_FALLBACK_GBS, "fallback"


def make_text(cfg_name: str, scale: float = 1.0, out=None, rank: int = 0, world: int = 1):
    """Rank `rank`'s shard of the workload: the named configuration per GPU (weak scaling).  With
    world > 1 the shards form ONE graph: rank r holds segments [r*n_seg, (r+1)*n_seg) and its links; the
    10 % "uniform" link targets are drawn over all ranks' segments, so dictionary merge and edge
    exchange are exercised."""
    from gfa2network_b200.synth import CONFIGS, synth_gfa

    cfg = CONFIGS[cfg_name]
    n_seg = max(2, int(cfg["n_seg"] * scale))
    n_link = max(1, int(cfg["n_link"] * scale))
    text = synth_gfa(n_seg, n_link, seed=cfg["seed"] + 1000 * rank, kind=cfg["kind"], seq_mean=cfg.get("seq_mean", 0),
                     n_paths=cfg.get("n_paths", 0), n_walks=cfg.get("n_walks", 0), out=out, id_base=rank * n_seg,
                     uniform_range=(0, world * n_seg) if world > 1 else (0, 0), header=(rank == 0))
    return cfg, text, n_seg, n_link


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""

    def __init__(self, index: int):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self, wait_s: float = 5.0):
        """nvidia-smi needs a few hundred ms to print its first row: start before the warm-up and wait
        for that row, so that the timed region itself is covered by samples."""
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,power.draw")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
            t0 = time.perf_counter()
            while not self.rows and time.perf_counter() - t0 < wait_s:
                time.sleep(0.02)
        except Exception:
            self.proc = None

    def mark(self):
        """number of rows seen so far (to cut the sample list at region boundaries)"""
        return len(self.rows)

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self, lo: int = 0, hi: int | None = None):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows[lo:hi]:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                for k, nme in enumerate(names):
                    if r[2 + k].lower().startswith("active"):
                        reasons.add(nme)
            except Exception:
                pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def bind_near_gpu(index: int):
    """Run this process (and first-touch its pinned buffers) on the CPUs of the GPU's NUMA node, as `numactl` would:
    host<->device copies of the e2e path then do not cross the socket interconnect.  Returns the CPU count or None."""
    try:
        import pynvml

        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = (os.cpu_count() + 63) // 64
        masks = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = [64 * i + b for i, m in enumerate(masks) for b in range(64) if (int(m) >> b) & 1]
        allowed = os.sched_getaffinity(0)
        cpus = [c for c in cpus if c in allowed]
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return None


def cpu_oracle_run(text, mode, fmt):
    from oracle.oracle import oracle_convert_format, oracle_parse_gfa

    t = time.perf_counter()
    A = oracle_parse_gfa(text, **mode)
    A = oracle_convert_format(A, fmt)
    return time.perf_counter() - t, A


REFERENCE_BUDGET_S = 150.0  # CPU seconds the whole --impl reference run may take (both legs together)


def _mode_text(cfg):
    return ",".join(f"{k}={v}" for k, v in cfg["mode"].items()) or "directed(default)"


def _time_passes(fn, n_warm, n_steps):
    for _ in range(n_warm):
        fn()
    ts = []
    for _ in range(n_steps):
        t = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t)
    return sum(ts) / len(ts)


def _scaled_text(cfg_name, target_bytes, full_scale):
    """Text of the configuration's shape with about target_bytes (same generator, same mix of records)."""
    frac = min(1.0, target_bytes / (CONFIG_BYTES_HINT[cfg_name] * full_scale))
    return make_text(cfg_name, full_scale * frac) + (frac,)


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path on this host.

    Leg 1 (the line's `value`, cpu_baseline.kind = "reference"): the UNMODIFIED reference package, staged from
    /root/reference into oracle/_ref by oracle/stage_ref.py (pure Python: `gfa2network.parse_gfa(path,
    build_graph=False, build_matrix=True, ...)` + `convert_format`, file in the page cache).  It is single-
    threaded: cores = 1.
    Leg 2 (`cpu_baseline_port`, kind "port"): the C restatement of its algorithm + SciPy (oracle/), the stronger
    baseline -- about 25x the reference's own speed on the same core.
    Every step is one pass over a text of the configuration's shape, sized so that W + K passes of each leg fit the
    time budget (GB/s is size-normalised; `sample` says what was run).  If oracle/_ref is absent the port is the line."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n_warm = max(0, min(args.warmup, 1))
    n_pass = args.steps + n_warm
    cfg0 = make_text(args.config, min(args.scale, 0.002))[0]
    mode, fmt = cfg0["mode"], cfg0["fmt"]
    full_bytes = CONFIG_BYTES_HINT[args.config] * args.scale
    # ---- leg 2: the port (rate estimate on a small text, then the largest text whose passes fit a third of the budget)
    _, probe, _, _ = make_text(args.config, min(args.scale, 0.02))
    dt, _A = cpu_oracle_run(probe, mode, fmt)
    rate = probe.size / dt
    cfg, text, n_seg, n_link, frac = _scaled_text(args.config, rate * (REFERENCE_BUDGET_S / 3) / n_pass, args.scale)
    port_s = _time_passes(lambda: cpu_oracle_run(text, mode, fmt), n_warm, args.steps)
    port = {"value": text.size / (port_s * 1e9), "unit": "GB/s", "cores": 1, "kind": "port", "host_cores_available": os.cpu_count(),
            "sample": (f"{args.config} shape at {frac:.4f} of the workload ({text.size} B per pass)" if frac < 1 else f"full {args.config} text ({text.size} B)")
                      + f", {args.steps} passes, C port of parser.py/builders.py + SciPy",
            "ms_per_step": port_s * 1e3, "edges_per_s": n_link / port_s}
    line_leg, ms, n_link_leg = port, port_s * 1e3, n_link
    # ---- leg 1: the reference itself
    ref_leg = None
    try:
        from oracle.stage_ref import check as ref_check, import_reference

        ref = import_reference()
    except Exception as exc:  # noqa: BLE001
        ref, ref_err = None, repr(exc)
    if ref is not None:
        import tempfile

        tmpdir = "/dev/shm" if os.path.isdir("/dev/shm") else None
        _, probe, _, _ = make_text(args.config, min(args.scale, 0.002))
        with tempfile.NamedTemporaryFile(suffix=".gfa", dir=tmpdir) as fh:
            fh.write(probe.tobytes())
            fh.flush()

            def ref_run(path=fh.name):
                A = ref.parse_gfa(path, build_graph=False, build_matrix=True, **mode)
                return ref.convert_format(A, fmt)
            t = time.perf_counter()
            ref_run()
            rrate = probe.size / (time.perf_counter() - t)
        cfg, rtext, r_seg, r_link, rfrac = _scaled_text(args.config, rrate * (REFERENCE_BUDGET_S * 2 / 3) / n_pass, args.scale)
        with tempfile.NamedTemporaryFile(suffix=".gfa", dir=tmpdir) as fh:
            fh.write(rtext.tobytes())
            fh.flush()

            def ref_run2(path=fh.name):
                A = ref.parse_gfa(path, build_graph=False, build_matrix=True, **mode)
                return ref.convert_format(A, fmt)
            ref_s = _time_passes(ref_run2, n_warm, args.steps)
        ref_leg = {"value": rtext.size / (ref_s * 1e9), "unit": "GB/s", "cores": 1, "kind": "reference", "host_cores_available": os.cpu_count(),
                   "sample": f"{args.config} shape at {rfrac:.5f} of the workload ({rtext.size} B per pass, file in /dev/shm), {args.steps} passes of the unmodified "
                             f"reference (oracle/_ref/gfa2network {getattr(ref, '__version__', '?')}: parse_gfa + convert_format, manifest ok: {ref_check()})",
                   "ms_per_step": ref_s * 1e3, "edges_per_s": r_link / ref_s}
        line_leg, ms, n_link_leg = ref_leg, ref_s * 1e3, r_link
    gbs = line_leg["value"]
    line = {
        "impl": "reference", "metric": "gfa_to_csr_parse_build_GBps", "value": gbs, "unit": "GB/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8/int32/f64", "data": "synthetic",
        "config": {"workload": workload_name(args.config, args.scale, *make_sizes(args.config, args.scale), cfg), "text_bytes_per_gpu": int(full_bytes)},
        "edges_per_s": n_link_leg / (ms / 1e3),
        "cpu_baseline": {k: v for k, v in line_leg.items() if k not in ("ms_per_step", "edges_per_s")},
        "cpu_baseline_port": port,
        "e2e": {"value": gbs, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    if ref_leg is None:
        line["reference_package"] = "oracle/_ref is absent (oracle/stage_ref.py stages it where /root/reference exists): the line is the port"
    print(json.dumps(line))


# approximate text bytes of the full configurations (only to decide whether generating them is affordable)
CONFIG_BYTES_HINT = {"C2": 86e6, "C3": 1.67e9, "C4": 11.9e9, "C4d": 11.9e9, "C5": 39.8e9}


def make_sizes(cfg_name: str, scale: float):
    from gfa2network_b200.synth import CONFIGS

    cfg = CONFIGS[cfg_name]
    return max(2, int(cfg["n_seg"] * scale)), max(1, int(cfg["n_link"] * scale))


def workload_name(cfg_name, scale, n_seg, n_link, cfg):
    mode = ",".join(f"{k}={v}" for k, v in cfg["mode"].items()) or "directed(default)"
    kind = "GFA-1 S/L" if cfg["kind"] == 1 else "reference E-dialect S/E + RC:f"
    return f"{cfg_name} synthetic {kind}: {n_seg} segments / {n_link} links, {mode}, {cfg['fmt'].upper()}" + ("" if scale == 1.0 else f" (scale {scale})")


# algorithmic bytes per launch of each kernel: compulsory reads + writes only
def algo_bytes(name, st):
    N, E, spe, M, n, nnz, weighted, world = (st[k] for k in ("N", "E", "spe", "M", "n", "nnz", "weighted", "world"))
    ent = 8 if weighted else 4  # bytes per row entry (rowsort.cuh: Ent64 / Ent32)
    table = {
        "k_tokenize": N + 4 * spe * E + (8 * E if weighted else 0),
        "k_mark_first": 24 * 2 * n,
        "k_assign_ids": 24 * 2 * n + 12 * n,
        "k_rows_count": 2 * 4 * spe * E + 4 * M + (16 * E if weighted else 0),
        "k_rows_scatter": 4 * spe * E + 4 * M + ent * M,           # ids in, cursors, entries out
        "k_edges_count_flat": 2 * 4 * spe * E + 4 * M,             # slots in, ids out, one histogram update per entry
        "k_edges_scatter_flat": 4 * spe * E + 4 * M + 4 * M,       # ids in, cursors, 32-bit entries out
        "k_rows_sort": 2 * ent * M + 8 * n,                        # entries in and (sorted) out, rowptr in, counts out
        "k_rows_write": ent * M + 8 * n + 12 * nnz,                # entries in, rowptr + indptr in, indices/data out
        "k_emit_coo": 4 * spe * E + 16 * M,
        # bucketed row build (row arrays far larger than L2): records -> bucket-major (major, entry) pairs -> rows
        "k_bucket_count": 2 * 4 * spe * E,                         # slots in, ids out
        "k_bucket_scatter": 4 * spe * E + (4 * E if weighted else 0) + (4 + ent) * M,
        "k_bucket_rows_count": 4 * M + 4 * M,
        "k_bucket_rows_scatter": (4 + ent) * M + 4 * M + ent * M,
        # multi-GPU (dist.cuh); n = keys of this shard, M = row entries of this shard / slab
        "k_dx_export": 24 * 2 * n + 24 * n,   # local table in (keys + first at load <= 0.5), key + order + position out
        "k_dx_insert": 20 * n + 24 * n + 4 * n,
        "k_dx_entries": 2 * 4 * spe * E + 8 * M,
        "k_pairs_count": 8 * M + 4 * M,
        "k_pairs_scatter": 8 * M + 4 * M + 4 * M,
    }
    return table.get(name)


def run_ours(args):
    import torch
    import torch.distributed as dist

    from gfa2network_b200 import _capi, parse_gfa

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    near = None if os.environ.get("G2N_BENCH_NO_BIND") else bind_near_gpu(local)
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)

    # weak scaling: every rank holds one shard of the configuration's shape
    cfg, text_np, n_seg, n_link = make_text(args.config, args.scale, rank=rank, world=world)
    nbytes = int(text_np.size)
    pinned = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    pinned.numpy()[:] = text_np
    text_dev = pinned.to(dev, non_blocking=True)
    torch.cuda.synchronize()

    stream = torch.cuda.current_stream()
    mode = cfg["mode"]
    builder = None
    if world > 1:
        from gfa2network_b200.dist import DistBuilder

        builder = DistBuilder(local)
        h = builder.local.h
    else:
        h = _capi.Handle(local)
        h.set_stream(stream.cuda_stream)
    want = {"csr": _capi.FMT_CSR, "csc": _capi.FMT_CSC, "coo": _capi.FMT_NATIVE}[cfg["fmt"]]
    wt = mode.get("weight_tag")
    wtb = wt.encode() if wt else None
    params = _capi.Params(int(mode.get("directed", True)), int(mode.get("bidirected", False)), int(mode.get("keep_directed_bidir", False)),
                          int(mode.get("asymmetric", False)), 0, _capi.DTYPES["float64"], want, 1, wtb, len(wtb) if wtb else 0, 0)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    dmode = dict(mode)
    state = {}

    def step():
        if builder is not None:
            state["res"] = builder.build(text_dev, matrix_format=cfg["fmt"] if cfg["fmt"] != "coo" else "csr", **dmode)
        else:
            h.check(h.build(text_dev.data_ptr(), nbytes, params))

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()  # before the warm-up: nvidia-smi's first row takes a few hundred ms
    # the very first build of this process: buffer allocation, capacity discovery (no size hints yet), host round trips
    torch.cuda.synchronize()
    t_first = time.perf_counter()
    step()
    torch.cuda.synchronize()
    first_call_ms = (time.perf_counter() - t_first) * 1e3
    for _ in range(max(args.warmup, 3) - 1):
        flush.zero_()
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    mark0 = sampler.mark()
    # timed region: K steps, per-step CUDA events on the launching stream, no per-kernel events
    h.set_profile(False)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    launches = 0
    torch.cuda.synchronize()
    for i in range(args.steps):
        flush.zero_()  # evict the text and the tables from L2 between steps (not timed)
        ev[i][0].record(stream)
        step()
        ev[i][1].record(stream)
        launches += h.status().gpu_launches
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    mark1 = sampler.mark()
    total_ms = sum(a.elapsed_time(b) for a, b in ev)
    t = torch.tensor([total_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    ms_step = total_ms / args.steps
    # ---- the same steps with speculation off: sizes are read back after the tokenizer and every buffer is sized from
    # them, as for a text this handle has not seen before (a `convert` sees each file once)
    ms_nonspec = None
    if builder is None:
        h.set_speculation(False)
        nsteps = max(1, min(args.steps, 20))
        nev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(nsteps)]
        for i in range(nsteps):
            flush.zero_()
            nev[i][0].record(stream)
            step()
            nev[i][1].record(stream)
        torch.cuda.synchronize()
        ms_nonspec = sum(a.elapsed_time(b) for a, b in nev) / nsteps
        h.set_speculation(True)
        step()  # (re-learn the hints)
        torch.cuda.synchronize()
    # ---- H2D ingest alone: the pinned text copied to the device, nothing else (reported separately, BASELINE north_star)
    iev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(5)]
    for a, b in iev:
        a.record(stream)
        text_dev.copy_(pinned, non_blocking=True)
        b.record(stream)
    torch.cuda.synchronize()
    ti = torch.tensor([min(a.elapsed_time(b) for a, b in iev)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ti, op=dist.ReduceOp.MAX)
    ingest_ms = float(ti.item())
    # per-kernel durations: a second pass of the same steps with one CUDA-event pair around every launch
    # (g2n_set_profile); kept out of the timed region because the event records themselves cost time
    h.set_profile(True)
    ktot: dict[str, list[float]] = {}
    ksteps = max(1, min(args.steps, 20))
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(ksteps)]
    for i in range(ksteps):
        flush.zero_()
        kev[i][0].record(stream)
        step()
        kev[i][1].record(stream)
        torch.cuda.synchronize()
        for k, (ms, cnt) in h.kernel_times().items():
            a = ktot.setdefault(k, [0.0, 0])
            a[0] += ms
            a[1] += cnt
    ms_step_profiled = sum(a.elapsed_time(b) for a, b in kev) / ksteps
    h.set_profile(False)
    diag = h.status()
    nb = C.c_uint64()
    h.check(h.lib.g2n_names_bytes(h.h, C.byref(nb)))  # untimed: only sizes the name table for the byte accounting
    sz = h.sizes()
    n_edges_rank = int(diag.n_edge_records)

    # ---- e2e through the public API: pinned host text in, host CSR arrays out
    fmt = cfg["fmt"]
    e2e_steps = max(2, min(args.steps, 10))
    host_in = pinned.numpy()
    if builder is None:
        def e2e_call():
            return parse_gfa(host_in, build_graph=False, build_matrix=True, matrix_format=fmt, device=local, **mode)
    else:
        class _Slab:  # the public multi-GPU call: DistBuilder.build on this rank's shard + fetch of its CSR slab
            pass

        def e2e_call():
            dev_text = pinned.to(dev, non_blocking=True)
            builder.build(dev_text, matrix_format=fmt if fmt != "coo" else "csr", **dmode)
            ip, ix, dt = builder.fetch_slab()
            o = _Slab()
            o.format, o.indptr, o.indices, o.data = "csr", ip, ix, dt
            return o
    A = None
    for _ in range(6):  # warm: device scratch and the pinned-buffer pool (two result generations) reach steady state
        A = e2e_call()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e2e_times = []
    for _ in range(e2e_steps):
        t0 = time.perf_counter()
        A = e2e_call()
        torch.cuda.synchronize()
        e2e_times.append(time.perf_counter() - t0)
    e2e_s = sum(e2e_times) / e2e_steps
    if rank == 0 and os.environ.get("G2N_BENCH_DEBUG"):
        print("e2e per-call ms:", [round(1e3 * x, 2) for x in e2e_times], file=sys.stderr)
    te = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_s = float(te.item())
    if A.format == "coo":
        d2h = A.row.nbytes + A.col.nbytes + A.data.nbytes
    else:
        d2h = A.indptr.nbytes + A.indices.nbytes + A.data.nbytes

    # ---- the host<->device copy floor of the e2e call on this box: the same bytes in (pinned -> device) and out
    # (device -> pinned), nothing else, all ranks at once.  e2e cannot be faster than this; with N processes sharing the
    # host's PCIe root complexes and memory controllers it is what bends the e2e scaling curve.
    out_dev = torch.empty(int(d2h), dtype=torch.uint8, device=dev)
    out_host = torch.empty(int(d2h), dtype=torch.uint8).pin_memory()
    copy_times = []
    for i in range(6):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        text_dev.copy_(pinned, non_blocking=True)
        out_host.copy_(out_dev, non_blocking=True)
        torch.cuda.synchronize()
        if i:
            copy_times.append(time.perf_counter() - t0)
    tcp = torch.tensor([sum(copy_times) / len(copy_times)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tcp, op=dist.ReduceOp.MAX)
    copy_floor_ms = float(tcp.item()) * 1e3
    del out_dev, out_host

    # ---- N > 1: the north-star workload as well -- every GPU holds 1/8 of C5 (12.5 M segments with sequences / 50 M
    # links, ~5 GB of text per GPU; at N = 8 that is BASELINE's config 5 in full: 100 M segments / 400 M links / ~40 GB).
    # Device-resident build of ONE graph over the N shards, CUDA events, max over ranks; reported beside the line's
    # metric (which stays the C2-shaped shard per GPU, so that the N = 1 .. 8 values form one weak-scaling curve).
    ns = None
    if world > 1 and args.config == "C2" and args.scale == 1.0 and not os.environ.get("G2N_BENCH_NO_C5"):
        try:
            text_dev = pinned = flush = None  # (referenced by closures above: rebinding, not del)
            torch.cuda.empty_cache()
            c5cfg, c5np, c5seg, c5link = make_text("C5", 0.125, rank=rank, world=world)
            c5bytes = int(c5np.size)
            c5dev = torch.from_numpy(c5np).to(dev)
            del c5np
            c5mode = dict(c5cfg["mode"])
            for _ in range(3):
                builder.build(c5dev, matrix_format="csr", **c5mode)
            torch.cuda.synchronize()
            nstep = 5
            cev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(nstep)]
            for a, b in cev:
                dist.barrier()
                a.record(stream)
                c5res = builder.build(c5dev, matrix_format="csr", **c5mode)
                b.record(stream)
            torch.cuda.synchronize()
            tc = torch.tensor([sum(a.elapsed_time(b) for a, b in cev) / nstep, float(c5bytes), float(c5res.nnz_local)], device=dev, dtype=torch.float64)
            tmax = tc.clone()
            dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
            dist.all_reduce(tc, op=dist.ReduceOp.SUM)
            c5ms, c5total, c5nnz = float(tmax[0].item()), float(tc[1].item()), int(tc[2].item())
            ns = {"workload": f"C5 shape, 1/8 of it per GPU: {world * c5seg} segments / {world * c5link} links, directed (default), CSR of max(S, S^T)",
                  "text_bytes": int(c5total), "nodes": int(c5res.n_global), "nnz": c5nnz, "ms_per_build": c5ms, "value": c5total / (c5ms * 1e6), "unit": "GB/s",
                  "edges_per_s": world * c5link / (c5ms / 1e3), "hbm_frac_of_aggregate_peak": c5total / (c5ms * 1e6) / (world * peaks()[0]), "steps": nstep,
                  "parity": "profiles/r2_dist_c5full_8gpu_oracle_sha.json: every slab and name range of this build at N = 8 is sha256-identical to the CPU oracle"}
            del c5dev
        except Exception as exc:  # noqa: BLE001 - the extra measurement must never take the line down
            ns = {"error": repr(exc)[:300]}
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    # clocks: samples taken inside the timed region when it was long enough to hold some, else the
    # whole loaded window (timed region + e2e loop); `window` says which
    if mark1 - mark0 >= 3:
        clocks = sampler.stop(mark0, mark1)
        clocks["window"] = "timed region"
    else:
        clocks = sampler.stop(mark0, None)
        clocks["window"] = "timed region + e2e loop (timed region shorter than 3 sampling periods)"
    # ---- roofline of the dominant kernel
    peak, peak_src = peaks()
    graph_directed = bool(params.keep_directed_bidir or (not params.bidirected and params.directed))
    spe = 4 if (params.bidirected and not params.keep_directed_bidir) else 2
    tpe = 1 if graph_directed else (4 if spe == 4 else 2)
    E_rank = int(diag.n_edge_records)
    M = E_rank * tpe * (2 if (graph_directed and not params.asymmetric) else 1)
    st = dict(N=nbytes, E=E_rank, spe=spe, M=M, n=int(sz.n_nodes), nnz=int(sz.nnz), weighted=bool(wtb), world=world)
    kern = {k: {"ms_per_step": v[0] / ksteps, "launches_per_step": v[1] / ksteps} for k, v in ktot.items()}
    dom = max(kern, key=lambda k: kern[k]["ms_per_step"]) if kern else None
    roof = None
    if dom:
        per_launch_ms = ktot[dom][0] / max(1, ktot[dom][1])
        ab = algo_bytes(dom, st)
        if ab:
            ach = ab / (per_launch_ms * 1e6)
            roof = {"bound": "hbm", "kernel": dom, "achieved": ach, "peak": peak, "peak_source": peak_src, "unit": "GB/s",
                    "frac": ach / peak, "traffic": TRAFFIC.get((args.config, float(args.scale), dom)) if world == 1 else None, "algorithmic_bytes_per_launch": ab, "ms_per_launch": per_launch_ms,
                    "share_of_step": kern[dom]["ms_per_step"] / ms_step}
    # whole-path algorithmic bytes: text in + result arrays out.  The node-name table (names + offsets,
    # SURVEY's last term) is NOT counted: it is gathered on demand (k_gather_names, when the caller asks for the node
    # list) and not inside the timed step.
    idx = 4
    dsz = 8
    if sz.format == _capi.FMT_COO:
        out_bytes = (2 * idx + dsz) * sz.nnz
    else:
        out_bytes = idx * (sz.n_nodes + 1) + (idx + dsz) * sz.nnz
    path_ach = (nbytes + out_bytes) / (ms_step * 1e6)
    # ---- CPU baseline on this host (rank 0, N=1 only; bounded: one full pass of the same text)
    cpu_base = None
    if world == 1 and not os.environ.get("G2N_BENCH_NO_CPU"):  # (kernel experiments skip the CPU pass)
        cpu_dt, _B = cpu_oracle_run(text_np, mode, fmt)
        cpu_base = {"value": nbytes / (cpu_dt * 1e9), "unit": "GB/s", "cores": 1, "kind": "port",
                    "sample": f"full {args.config} text ({nbytes} B), 1 run: C port of parser.py/builders.py + SciPy tocsr/maximum",
                    "host_cores_available": os.cpu_count()}
    gbs = world * nbytes / (ms_step * 1e6)
    stage_ms = {k: float(diag.ms_stage[i]) for i, k in ((0, "tokenize+hash"), (1, "ids"), (3, "emit"), (4, "sort"), (5, "reduce"))}
    if world > 1:
        # the multi-GPU build queues its stages without the single-GPU stage events: group the per-kernel times (rank 0)
        groups = {"tokenize+hash": ("k_tokenize", "k_tokenize_slow"), "dictionary": ("k_dx_export", "k_dx_owner", "k_dx_insert", "k_dx_reply_first", "k_dx_mark", "k_dx_send_rank",
                                                                                       "k_dx_reply_ids", "k_dx_localmap", "k_dx_ids", "k_dx_rank"),
                  "entries": ("k_dx_entries", "k_dxw_count", "k_dxw_scatter", "k_dx_slab_sizes"), "slab": ("k_pairs_count", "k_pairs_scatter", "k_pairsw_count", "k_pairsw_scatter",
                                                                                                           "k_rows_big", "k_rows_sort", "k_rows_write", "k_dx_final")}
        stage_ms = {g: sum(kern[k]["ms_per_step"] for k in ks if k in kern) for g, ks in groups.items()}
        stage_ms["scans"] = sum(v["ms_per_step"] for k, v in kern.items() if k.startswith("k_scan"))
    line = {
        "metric": "gfa_to_csr_parse_build_GBps", "value": gbs, "unit": "GB/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8/int32/f64", "data": "synthetic",
        "config": {"workload": workload_name(args.config, args.scale, n_seg, n_link, cfg), "text_bytes_per_gpu": nbytes,
                   "l2": "flushed between steps (512 MiB write)", "nodes": int(sz.n_nodes), "nnz": int(sz.nnz),
                   "sharding": ("one shard of this shape per GPU, ONE graph; hash-partitioned node dictionary and row entries by owner row "
                                "block, both written straight into peer memory over NVLink by the kernels (no collective library, no "
                                "host round trip inside a step); per-GPU CSR slab") if world > 1 else "single GPU"},
        "edges_per_s": world * n_link / (ms_step / 1e3),
        "path_roofline": {"algorithmic_bytes": int(nbytes + out_bytes), "achieved": path_ach, "peak": peak, "frac": path_ach / peak, "unit": "GB/s"},
        "roofline": roof,
        "kernels": kern,
        "kernel_timing": {"how": "second pass with a CUDA-event pair around every launch", "steps": ksteps, "ms_per_step_with_events": ms_step_profiled},
        "stage_ms": stage_ms,
        "value_nonspeculative": (world * nbytes / (ms_nonspec * 1e6)) if ms_nonspec else None,
        "ms_per_step_nonspeculative": ms_nonspec,
        "first_call_ms": first_call_ms,
        "speculation": "the timed steps rebuild a text of the same size and mode: buffers are sized from the previous build and the host looks at the "
                       "counters once, at the end (g2n_set_speculation); value_nonspeculative = the same steps with the host round trip after the "
                       "tokenizer, first_call_ms = the first build of the process (allocation + capacity discovery, wall clock)",
        "ingest": {"value": world * nbytes / (ingest_ms * 1e6), "unit": "GB/s", "ms": ingest_ms, "what": "pinned host text -> device copy alone (best of 5, max over ranks)"},
        "cpu_baseline": cpu_base,
        "e2e": {"value": world * nbytes / (e2e_s * 1e9), "unit": "GB/s", "h2d_bytes_per_step": nbytes, "d2h_bytes_per_step": int(d2h),
                "ms_per_step": e2e_s * 1e3, "host_cpus_near_gpu": near,
                "copy_floor_ms": copy_floor_ms, "copy_floor": "the same bytes pinned -> device and device -> pinned with no build in between, all ranks at once (max over ranks)",
                "api": "gfa2network_b200.parse_gfa(pinned uint8 buffer, matrix_format=...)" if world == 1 else
                       "gfa2network_b200.dist.DistBuilder.build(shard) + fetch_slab() per rank (bytes are per rank)"},
        "gpu_launches": launches,
        "clocks": clocks,
    }
    if ns is not None:
        line["north_star_shards"] = ns
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--config", default="C2", choices=["C2", "C3", "C4", "C4d", "C5"])
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
