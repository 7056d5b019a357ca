"""CPU-only checks of the boundary: the C-ABI library builds, loads and exports every symbol
include/g2n.h declares; the product path fails loudly without a GPU (no CPU fallback)."""
import ctypes
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]


def _declared():
    text = (ROOT / "include" / "g2n.h").read_text()
    return sorted(set(re.findall(r"\b(g2n_[a-z_0-9]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from gfa2network_b200 import _capi

    lib = _capi.load()
    names = _declared()
    assert len(names) >= 14
    for name in names:
        assert hasattr(lib, name), name
    assert sorted(_capi.EXPORTS) == names
    assert lib.g2n_abi_version() == _capi.ABI_VERSION


def test_header_is_plain_c():
    """include/g2n.h is the boundary a C / cgo / JNI binding would include: it must compile as C99 on its own."""
    import shutil
    import subprocess

    gcc = shutil.which("gcc")
    if not gcc:
        pytest.skip("no gcc")
    for std, lang in (("-std=c99", "c"), ("-std=c++17", "c++")):
        r = subprocess.run([gcc, std, "-Wall", "-Wextra", "-pedantic", "-fsyntax-only", "-x", lang, str(ROOT / "include" / "g2n.h")], capture_output=True, text=True)
        assert r.returncode == 0 and not r.stderr.strip(), r.stderr


def test_struct_layouts_match_header():
    from gfa2network_b200 import _capi

    assert ctypes.sizeof(_capi.Params) == 48
    assert ctypes.sizeof(_capi.Sizes) == 64
    assert ctypes.sizeof(_capi.DistResult) == 11 * 8 + 2 * 64
    assert ctypes.sizeof(_capi.DistInfo) == 48
    assert ctypes.sizeof(_capi.Diag) == 112
    assert ctypes.sizeof(_capi.PathInfo) == 48
    # the C compiler's view of the same structs
    import shutil
    import subprocess
    import tempfile

    gcc = shutil.which("gcc")
    if gcc:
        src = '#include <stdio.h>\n#include "g2n.h"\nint main(void){printf("%zu %zu %zu %zu %zu %zu\\n", sizeof(g2n_params), sizeof(g2n_sizes_t), sizeof(g2n_dist_info), sizeof(g2n_diag), sizeof(g2n_dist_result), sizeof(g2n_path_info_t));return 0;}\n'
        with tempfile.TemporaryDirectory() as d:
            c = Path(d) / "s.c"
            c.write_text(src)
            subprocess.run([gcc, "-I", str(ROOT / "include"), "-o", str(Path(d) / "s"), str(c)], check=True)
            out = subprocess.run([str(Path(d) / "s")], capture_output=True, text=True, check=True).stdout.split()
        assert [int(x) for x in out] == [ctypes.sizeof(t) for t in (_capi.Params, _capi.Sizes, _capi.DistInfo, _capi.Diag, _capi.DistResult, _capi.PathInfo)]


def test_no_cpu_fallback():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from gfa2network_b200 import _capi, parse_gfa

    with pytest.raises(_capi.G2NError, match="no CPU fallback"):
        parse_gfa(b"S\ta\t*\n", build_graph=False, build_matrix=True)


def test_out_of_scope_arguments_fail_loudly():
    from gfa2network_b200 import parse_gfa

    with pytest.raises(ValueError, match="return_node_list requires build_matrix=True"):
        parse_gfa(b"", build_graph=False, build_matrix=False, return_node_list=True)
    for kw in (dict(backend="igraph"), dict(split_on_alignment=True)):
        with pytest.raises(NotImplementedError):
            parse_gfa(b"", build_graph=False, build_matrix=True, **kw)
    with pytest.raises(NotImplementedError):
        parse_gfa(b"", build_graph=True, build_matrix=True)


def test_product_never_imports_oracle():
    for p in (ROOT / "gfa2network_b200").rglob("*"):
        if p.suffix in {".py", ".cu", ".cuh", ".h", ".c"}:
            assert "oracle" not in p.read_text().replace("CPU oracle", "").replace("the oracle", "").lower() or p.name in {"synth.c", "synth.py"}, p


def test_plain_c_example_links_against_the_library(tmp_path):
    """examples/convert.c uses the ABI from C exactly as a cgo / JNI binding would; it must compile, link against
    libg2n.so and -- without a GPU -- fail loudly at g2n_create (no CPU fallback)."""
    import shutil
    import subprocess

    import torch

    from gfa2network_b200 import _capi

    gcc = shutil.which("gcc")
    if not gcc:
        pytest.skip("no gcc")
    _capi.load()
    lib_dir = _capi.lib_path().parent
    exe = tmp_path / "convert"
    subprocess.run([gcc, "-std=c99", "-Wall", "-Wextra", "-I", str(ROOT / "include"), "-o", str(exe), str(ROOT / "examples" / "convert.c"),
                    "-L", str(lib_dir), "-lg2n", f"-Wl,-rpath,{lib_dir}"], check=True)
    r = subprocess.run([str(exe), str(ROOT / "tests" / "golden" / "DRB1-3123_unsorted.gfa")], capture_output=True, text=True)
    if torch.cuda.is_available():
        assert r.returncode == 0 and "nodes 3214  nnz 8751" in r.stdout, (r.stdout, r.stderr)  # C1 default directed CSR
    else:
        assert r.returncode == 1 and "no CPU fallback" in r.stderr
