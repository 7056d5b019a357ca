"""Wall clock of parse_gfa("x.gfa.gz"): the library's windowed / block-parallel inflate (g2n_build_gz) against the
reference-style host inflate (gzip.open().read(), one core) followed by the same build.
    python tools/bench_gz.py [C2|C5] [scale]"""
import gzip
import os
import struct
import sys
import tempfile
import time
import zlib
from concurrent.futures import ThreadPoolExecutor

sys.path.insert(0, os.getcwd())
from bench import make_text  # noqa: E402
from gfa2network_b200 import parse_gfa  # noqa: E402


def bgzf(data: memoryview, block: int = 0xFF00) -> bytes:
    def one(a):
        chunk = bytes(data[a:a + block]) if a is not None else b""
        c = zlib.compressobj(6, zlib.DEFLATED, -15)
        body = c.compress(chunk) + c.flush()
        return (b"\x1f\x8b\x08\x04" + b"\0" * 4 + b"\x00\xff" + struct.pack("<H", 6) + b"BC" + struct.pack("<HH", 2, 12 + 6 + len(body) + 8 - 1) + body
                + struct.pack("<II", zlib.crc32(chunk), len(chunk)))
    with ThreadPoolExecutor(os.cpu_count()) as pool:
        return b"".join(pool.map(one, list(range(0, len(data), block)) + [None]))


name = sys.argv[1] if len(sys.argv) > 1 else "C2"
scale = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
cfg, text, _, _ = make_text(name, scale)
mode = dict(cfg["mode"])
d = tempfile.mkdtemp(dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
raw = memoryview(text)
files = {"gzip -6 (one stream)": os.path.join(d, "a.gfa.gz"), "BGZF (bgzip blocks)": os.path.join(d, "b.gfa.gz")}
with open(files["gzip -6 (one stream)"], "wb") as fh:
    fh.write(gzip.compress(raw, 6))
with open(files["BGZF (bgzip blocks)"], "wb") as fh:
    fh.write(bgzf(raw))
print(f"{name} x {scale}: {text.size/1e6:.0f} MB of text, host cores {os.cpu_count()}")
for label, path in files.items():
    for rep in range(3):
        os.environ.pop("G2N_HOST_GZIP", None)
        t = time.perf_counter()
        A = parse_gfa(path, build_graph=False, build_matrix=True, matrix_format="csr", **mode)
        t_lib = time.perf_counter() - t
        os.environ["G2N_HOST_GZIP"] = "1"  # the reference's way: gzip.open(path).read() on one core, then the build
        t = time.perf_counter()
        B = parse_gfa(path, build_graph=False, build_matrix=True, matrix_format="csr", **mode)
        t_host = time.perf_counter() - t
        assert A.nnz == B.nnz and (A.indices == B.indices).all()
    print(f"  {label:22s} {os.path.getsize(path)/1e6:7.1f} MB compressed | library inflate + build {1e3*t_lib:8.1f} ms = {text.size/t_lib/1e9:5.2f} GB/s of text | "
          f"host gzip.open().read() + build {1e3*t_host:8.1f} ms = {text.size/t_host/1e9:5.2f} GB/s | x{t_host/t_lib:.1f}")
for p in files.values():
    os.remove(p)
