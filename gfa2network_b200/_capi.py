"""ctypes binding of include/g2n.h (libg2n.so).  Thin: structs, prototypes, error mapping.

The library is the only compute path.  If it cannot be loaded, or no CUDA device is usable,
every entry point raises -- there is no CPU fallback."""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import numpy as np

from . import _build

G2N_OK, G2N_ERR_CUDA, G2N_ERR_INVALID, G2N_ERR_PARSE, G2N_ERR_UNSUPPORTED, G2N_ERR_INTERNAL, G2N_ERR_RETRY = range(7)
ABI_VERSION = 2
FMT_NATIVE, FMT_CSR, FMT_CSC, FMT_COO = 0, 1, 2, 3
FMT_NAMES = {FMT_CSR: "csr", FMT_CSC: "csc", FMT_COO: "coo"}
DTYPES = {"float64": 0, "float32": 1, "int32": 2, "int8": 3, "bool": 4}
DTYPE_NP = {0: np.float64, 1: np.float32, 2: np.int32, 3: np.int8, 4: np.bool_}

EXPORTS = [
    "g2n_abi_version", "g2n_create", "g2n_destroy", "g2n_set_stream", "g2n_host_alloc", "g2n_host_free",
    "g2n_build", "g2n_build_file", "g2n_build_gz", "g2n_convert", "g2n_sizes", "g2n_fetch_matrix", "g2n_names_bytes", "g2n_fetch_names", "g2n_device_result",
    "g2n_status", "g2n_last_error", "g2n_coo_to_compressed", "g2n_set_profile", "g2n_set_speculation", "g2n_kernel_times", "g2n_nodes_tsv_bytes", "g2n_fetch_nodes_tsv",
    "g2n_edge_list_bytes", "g2n_fetch_edge_list", "g2n_bfs", "g2n_levels_reduce", "g2n_fetch_levels",
    "g2n_paths_load", "g2n_path_info", "g2n_path_bfs", "g2n_path_reduce", "g2n_fetch_path_nodes", "g2n_fetch_text",
    "g2n_dist_init", "g2n_dist_probe", "g2n_dist_plan", "g2n_dist_local_mem", "g2n_dist_set_peers", "g2n_dist_open_peers",
    "g2n_dist_close_peers", "g2n_dist_stage", "g2n_dist_finish", "g2n_dist_enable_peer", "g2n_load_file_range", "g2n_plan_row_buckets",
]


class Params(C.Structure):
    _fields_ = [
        ("directed", C.c_int32), ("bidirected", C.c_int32), ("keep_directed_bidir", C.c_int32),
        ("asymmetric", C.c_int32), ("strip_orientation", C.c_int32), ("dtype", C.c_int32),
        ("want_format", C.c_int32), ("text_on_device", C.c_int32),
        ("weight_tag", C.c_char_p), ("weight_tag_len", C.c_int32), ("reserved", C.c_int32),
    ]


class Sizes(C.Structure):
    _fields_ = [
        ("n_nodes", C.c_uint64), ("nnz", C.c_uint64), ("names_bytes", C.c_uint64),
        ("format", C.c_int32), ("index_bytes", C.c_int32), ("dtype", C.c_int32), ("reserved", C.c_int32),
        ("slab_rows", C.c_uint64), ("names_count", C.c_uint64), ("names_id0", C.c_uint64),
    ]


class DistInfo(C.Structure):
    _fields_ = [("n_keys", C.c_uint64), ("n_tiles", C.c_uint64), ("n_records", C.c_uint64),
                ("n_edge_records", C.c_uint64), ("n_entries", C.c_uint64), ("reserved", C.c_uint64)]


class DistResult(C.Structure):
    _fields_ = [("bad", C.c_uint64), ("n_global", C.c_uint64), ("row0", C.c_uint64), ("n_rows", C.c_uint64), ("nnz", C.c_uint64),
                ("n_recv", C.c_uint64), ("n_first", C.c_uint64), ("id0", C.c_uint64), ("n_keys", C.c_uint64), ("n_records", C.c_uint64),
                ("n_edge_records", C.c_uint64), ("keys_to", C.c_uint64 * 8), ("pairs_to", C.c_uint64 * 8)]


class PathInfo(C.Structure):
    _fields_ = [("line_offset", C.c_uint64), ("name_offset", C.c_uint64), ("n_entries", C.c_uint64), ("missing_entry", C.c_int64),
                ("missing_offset", C.c_uint64), ("name_len", C.c_uint32), ("missing_len", C.c_uint32)]


class Diag(C.Structure):
    _fields_ = [
        ("err_kind", C.c_int32), ("unknown_byte", C.c_int32),
        ("err_offset", C.c_uint64), ("unknown_offset", C.c_uint64),
        ("n_records", C.c_uint64), ("n_edge_records", C.c_uint64), ("n_triplets", C.c_uint64),
        ("n_long_keys", C.c_uint64), ("retries", C.c_uint32), ("gpu_launches", C.c_uint32),
        ("ms_total", C.c_float), ("ms_h2d", C.c_float), ("ms_stage", C.c_float * 8),
        ("warn_flags", C.c_uint32), ("speculative", C.c_uint32),
    ]


class KTime(C.Structure):
    _fields_ = [("name", C.c_char * 40), ("ms", C.c_float), ("launches", C.c_uint32)]


_lib = None


def lib_path() -> Path:
    return _build.LIB


def load():
    """dlopen libg2n.so (building it first if the sources are newer) and set prototypes."""
    global _lib
    if _lib is not None:
        return _lib
    import os

    path = os.environ.get("G2N_LIB") or _build.build()  # G2N_LIB: a specific build of the same sources (kernel-variant experiments)
    lib = C.CDLL(str(path))
    vp, u64, i32 = C.c_void_p, C.c_uint64, C.c_int32
    lib.g2n_abi_version.restype = C.c_int
    lib.g2n_create.argtypes = [C.c_int, C.POINTER(vp)]
    lib.g2n_destroy.argtypes = [vp]
    lib.g2n_destroy.restype = None
    lib.g2n_set_stream.argtypes = [vp, vp]
    lib.g2n_host_alloc.argtypes = [u64]
    lib.g2n_host_alloc.restype = vp
    lib.g2n_host_free.argtypes = [vp]
    lib.g2n_host_free.restype = None
    lib.g2n_build.argtypes = [vp, vp, u64, C.POINTER(Params)]
    lib.g2n_build_file.argtypes = [vp, C.c_char_p, C.POINTER(Params)]
    lib.g2n_build_gz.argtypes = [vp, C.c_char_p, C.POINTER(Params)]
    lib.g2n_convert.argtypes = [vp, i32]
    lib.g2n_sizes.argtypes = [vp, C.POINTER(Sizes)]
    lib.g2n_fetch_matrix.argtypes = [vp, vp, vp, vp]
    lib.g2n_names_bytes.argtypes = [vp, C.POINTER(u64)]
    lib.g2n_fetch_names.argtypes = [vp, vp, vp]
    lib.g2n_device_result.argtypes = [vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp)]
    lib.g2n_status.argtypes = [vp, C.POINTER(Diag)]
    lib.g2n_last_error.argtypes = [vp]
    lib.g2n_last_error.restype = C.c_char_p
    lib.g2n_coo_to_compressed.argtypes = [vp, vp, vp, vp, u64, u64, i32, i32, vp, vp, vp, C.POINTER(u64)]
    pu64 = C.POINTER(u64)
    lib.g2n_dist_init.argtypes = [vp, C.c_int, C.c_int]
    lib.g2n_dist_probe.argtypes = [vp, vp, u64, C.POINTER(Params), C.POINTER(DistInfo)]
    lib.g2n_dist_plan.argtypes = [vp, u64, u64, u64, u64, C.c_int, C.POINTER(C.c_int)]
    lib.g2n_dist_local_mem.argtypes = [vp, C.POINTER(vp), C.POINTER(vp), vp]
    lib.g2n_dist_set_peers.argtypes = [vp, C.POINTER(vp), C.POINTER(vp)]
    lib.g2n_dist_open_peers.argtypes = [vp, C.c_char_p]
    lib.g2n_dist_close_peers.argtypes = [vp]
    lib.g2n_dist_stage.argtypes = [vp, C.c_int, vp, u64, C.POINTER(Params), C.c_int]
    lib.g2n_dist_finish.argtypes = [vp, C.POINTER(DistResult)]
    lib.g2n_dist_enable_peer.argtypes = [vp, C.c_int]
    lib.g2n_load_file_range.argtypes = [vp, C.c_char_p, u64, u64, C.POINTER(vp)]
    lib.g2n_set_profile.argtypes = [vp, C.c_int]
    lib.g2n_set_speculation.argtypes = [vp, C.c_int]
    lib.g2n_nodes_tsv_bytes.argtypes = [vp, C.POINTER(u64)]
    lib.g2n_fetch_nodes_tsv.argtypes = [vp, vp]
    lib.g2n_kernel_times.argtypes = [vp, C.POINTER(KTime), C.c_int]
    lib.g2n_paths_load.argtypes = [vp, C.POINTER(u64)]
    lib.g2n_path_info.argtypes = [vp, u64, C.POINTER(PathInfo)]
    lib.g2n_path_bfs.argtypes = [vp, u64, i32, i32]
    lib.g2n_path_reduce.argtypes = [vp, i32, u64, vp]
    lib.g2n_fetch_path_nodes.argtypes = [vp, u64, vp]
    lib.g2n_fetch_text.argtypes = [vp, u64, u64, vp]
    lib.g2n_bfs.argtypes = [vp, vp, u64, i32, i32]
    lib.g2n_levels_reduce.argtypes = [vp, i32, vp, u64, vp]
    lib.g2n_fetch_levels.argtypes = [vp, i32, vp]
    lib.g2n_edge_list_bytes.argtypes = [vp, C.POINTER(u64)]
    lib.g2n_fetch_edge_list.argtypes = [vp, vp]
    _lib = lib
    return lib


class G2NError(RuntimeError):
    pass


class Handle:
    """One libg2n handle = one device, one stream, warm device scratch."""

    def __init__(self, device: int = 0):
        self.lib = load()
        h = C.c_void_p()
        rc = self.lib.g2n_create(device, C.byref(h))
        if rc != G2N_OK or not h:
            raise G2NError(
                "g2n_create failed: no usable CUDA device (the GFA->matrix path has no CPU fallback)")
        self.h = h
        self.device = device
        self.generation = 0  # bumped by every call that replaces the device-resident result (build, build_file, coo_to_compressed)

    def close(self):
        if getattr(self, "h", None):
            self.lib.g2n_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def last_error(self) -> str:
        return (self.lib.g2n_last_error(self.h) or b"").decode(errors="replace")

    def check(self, rc: int):
        if rc == G2N_OK:
            return
        msg = self.last_error()
        if rc == G2N_ERR_UNSUPPORTED:
            raise NotImplementedError(msg)
        if rc == G2N_ERR_INVALID:
            raise ValueError(msg or "invalid argument")
        raise G2NError(f"libg2n status {rc}: {msg}")

    def set_stream(self, stream_ptr: int | None):
        """cudaStream_t as an integer (0 = the CUDA default stream); None = the handle's own stream."""
        self.check(self.lib.g2n_set_stream(self.h, C.c_void_p(-1 if stream_ptr is None else stream_ptr)))

    def set_profile(self, on: bool):
        self.check(self.lib.g2n_set_profile(self.h, int(on)))

    def fetch_nodes_tsv(self) -> np.ndarray:
        """The node map file of the last build as bytes: "<index>\\t<name>\\n" per node (utils.py:108-114)."""
        nb = C.c_uint64()
        self.check(self.lib.g2n_nodes_tsv_bytes(self.h, C.byref(nb)))
        out = np.empty(nb.value, dtype=np.uint8)
        self.check(self.lib.g2n_fetch_nodes_tsv(self.h, C.c_void_p(out.ctypes.data)))
        return out

    def fetch_edge_list(self) -> np.ndarray:
        """The edge list of the last build as bytes: "<from>\\t<to>\\n" per L / E / C record in file order (cli.py:267-281)."""
        nb = C.c_uint64()
        self.check(self.lib.g2n_edge_list_bytes(self.h, C.byref(nb)))
        out = np.empty(nb.value, dtype=np.uint8)
        self.check(self.lib.g2n_fetch_edge_list(self.h, C.c_void_p(out.ctypes.data)))
        return out

    def set_speculation(self, on: bool):
        """Repeat builds of the same input size and mode skip the host round trip after the tokenizer
        (include/g2n.h: g2n_set_speculation); on by default."""
        self.check(self.lib.g2n_set_speculation(self.h, int(on)))

    def kernel_times(self) -> dict[str, tuple[float, int]]:
        buf = (KTime * 32)()
        n = self.lib.g2n_kernel_times(self.h, buf, 32)
        return {buf[i].name.decode(): (float(buf[i].ms), int(buf[i].launches)) for i in range(max(n, 0))}

    def status(self) -> Diag:
        d = Diag()
        self.lib.g2n_status(self.h, C.byref(d))
        return d

    def sizes(self) -> Sizes:
        s = Sizes()
        self.check(self.lib.g2n_sizes(self.h, C.byref(s)))
        return s

    def build(self, text_ptr: int, nbytes: int, params: Params) -> int:
        self.generation += 1
        return self.lib.g2n_build(self.h, C.c_void_p(text_ptr), nbytes, C.byref(params))

    def build_file(self, path: str, params: Params) -> int:
        self.generation += 1
        return self.lib.g2n_build_file(self.h, os.fsencode(path), C.byref(params))

    def build_gz(self, path: str, params: Params) -> int:
        self.generation += 1
        return self.lib.g2n_build_gz(self.h, os.fsencode(path), C.byref(params))

    def coo_to_compressed(self, row, col, data, nnz: int, n: int, code: int, want: int, indptr, indices, dout) -> int:
        """Stage K4 alone on caller-provided triplets (g2n_coo_to_compressed); it reuses the handle's device
        scratch, so whatever build was resident is gone afterwards."""
        self.generation += 1
        nnz_out = C.c_uint64()
        self.check(self.lib.g2n_coo_to_compressed(self.h, row.ctypes.data, col.ctypes.data, data.ctypes.data, nnz, n, code, want,
                                                  indptr.ctypes.data, indices.ctypes.data, dout.ctypes.data, C.byref(nnz_out)))
        return int(nnz_out.value)

    def convert(self, fmt: int):
        self.check(self.lib.g2n_convert(self.h, fmt))

    def fetch_matrix(self):
        """Result arrays as NumPy arrays backed by pinned host memory (DMA target of the D2H copy)."""
        s = self.sizes()
        dt = DTYPE_NP[s.dtype]
        n, nnz = s.n_nodes, s.nnz
        a0 = pinned_empty(nnz if s.format == FMT_COO else s.slab_rows + 1, np.int32)
        a1 = pinned_empty(nnz, np.int32)
        data = pinned_empty(nnz, dt)
        self.check(self.lib.g2n_fetch_matrix(self.h, a0.ctypes.data, a1.ctypes.data, data.ctypes.data))
        return s, a0, a1, data

    def fetch_names(self):
        nb = C.c_uint64()
        self.check(self.lib.g2n_names_bytes(self.h, C.byref(nb)))
        s = self.sizes()
        names = np.empty(max(1, nb.value), dtype=np.uint8)
        offs = np.empty(s.names_count + 1, dtype=np.uint64)
        self.check(self.lib.g2n_fetch_names(self.h, names.ctypes.data, offs.ctypes.data))
        return names[: nb.value], offs


class _PinnedPool:
    """Size-classed cache of page-locked host blocks.  cudaHostAlloc / cudaFreeHost cost tens of
    milliseconds for result-sized buffers, so blocks are recycled when their NumPy views die."""

    MAX_CACHED = 8 << 30

    def __init__(self):
        self.free: dict[int, list[int]] = {}
        self.cached = 0

    @staticmethod
    def size_class(nbytes: int) -> int:
        n = max(nbytes, 4096)
        k = max(n.bit_length() - 4, 0)
        return ((n + (1 << k) - 1) >> k) << k  # <= 12.5 % above the request

    def take(self, nbytes: int) -> tuple[int, int]:
        cls = self.size_class(nbytes)
        lst = self.free.get(cls)
        if lst:
            self.cached -= cls
            return lst.pop(), cls
        ptr = load().g2n_host_alloc(cls)
        if not ptr:
            self.trim(0)
            ptr = load().g2n_host_alloc(cls)
            if not ptr:
                raise MemoryError(f"g2n_host_alloc({cls}) failed")
        return ptr, cls

    def give(self, ptr: int, cls: int):
        if self.cached + cls > self.MAX_CACHED:
            load().g2n_host_free(ptr)
            return
        self.free.setdefault(cls, []).append(ptr)
        self.cached += cls

    def trim(self, keep_bytes: int = 0):
        for cls, lst in list(self.free.items()):
            while lst and self.cached > keep_bytes:
                load().g2n_host_free(lst.pop())
                self.cached -= cls


_pool = _PinnedPool()


def _release(key: int):
    ent = _keepalive.pop(key, None)
    if ent is not None:
        try:
            _pool.give(*ent)
        except Exception:
            pass


def pinned_empty(count: int, dtype) -> np.ndarray:
    """np.empty(count, dtype) in page-locked host memory (DMA target / source); raises on failure."""
    import weakref

    dt = np.dtype(dtype)
    nbytes = int(count) * dt.itemsize
    if nbytes == 0:
        return np.empty(0, dtype=dt)
    ptr, cls = _pool.take(nbytes)
    buf = (C.c_uint8 * nbytes).from_address(ptr)
    arr = np.frombuffer(buf, dtype=dt, count=int(count))
    # the block goes back to the pool when the last view of the array is gone
    _keepalive[id(buf)] = (ptr, cls)
    weakref.finalize(buf, _release, id(buf))
    return arr


_keepalive: dict[int, tuple[int, int]] = {}
_default: dict[int, Handle] = {}


def default_handle(device: int = 0) -> Handle:
    h = _default.get(device)
    if h is None or h.h is None:
        h = _default[device] = Handle(device)
    return h
