"""``gfa2network convert`` (gfa2network/cli.py:22-135, 193-250), ``export --format edge-list`` (cli.py:137-150,
264-281), ``distance --path`` (cli.py:161-175, 302-334) and ``distance-matrix`` (cli.py:177-190, 335-350) over the
GPU path.  Same flags, defaults, stdout/stderr strings and exit behaviour; the sub-commands and formats that need
the NetworkX object graph (``stats``, ``distance --seq``, graph exports) are not provided here."""
from __future__ import annotations

import argparse
from pathlib import Path

from .builders import parse_gfa
from .utils import convert_format, save_matrix, save_node_map
from .version import __version__


def _parser() -> argparse.ArgumentParser:
    ap = argparse.ArgumentParser(prog="gfa2network")
    ap.add_argument("--version", action="version", version=f"gfa2network {__version__}")
    ap.add_argument("--raw-bytes-id", action="store_true", help="Use raw bytes for node identifiers (legacy)")
    ap.add_argument("--max-dense-gb", type=float, default=5.0, help="Abort dense matrix saves over N GB (default 5)")
    ap.add_argument("--max-tag-mb", type=float, default=100.0, help="Warn when stored tags exceed N MB (default 100)")
    sub = ap.add_subparsers(dest="cmd", required=True)
    c = sub.add_parser("convert", help="Convert GFA to graph or matrix")
    c.add_argument("gfa", help="Input *.gfa* file or - for stdin")
    c.add_argument("--backend", choices=["networkx", "igraph"], default="networkx", help="Graph backend to use")
    g = c.add_mutually_exclusive_group()
    g.add_argument("--directed", dest="directed", action="store_true", default=True, help="Treat graph as directed")
    g.add_argument("--undirected", dest="directed", action="store_false", help="Treat graph as undirected")
    c.add_argument("--graph", action="store_true", help="Build a NetworkX object")
    c.add_argument("--matrix", metavar="PATH", help="Write adjacency matrix to PATH (.npz|.npy|.csv)")
    c.add_argument("--save-matrix", dest="matrix", metavar="PATH", help=argparse.SUPPRESS)
    c.add_argument("--matrix-format", default="csr", help="Sparse format for .npz (csr|csc|coo|dok)")
    c.add_argument("--dtype", choices=["bool", "int8", "int32", "float32", "float64"], default="float64",
                   help="Data type for adjacency matrix")
    c.add_argument("--asymmetric", action="store_true", help="Do not mirror upper triangle")
    c.add_argument("--no-node-map", action="store_true", help="Do not write <matrix>.nodes.tsv sidecar")
    c.add_argument("--weight-tag")
    c.add_argument("--store-seq", action="store_true")
    c.add_argument("--store-tags", action="store_true")
    c.add_argument("--split-on-alignment", action="store_true", help="Split segments at alignment boundaries")
    c.add_argument("--strip-orientation", action="store_true", help="Strip +/- from IDs (v0.1 behaviour)")
    c.add_argument("--bidirected", action="store_true", help="Use bidirected representation")
    c.add_argument("--keep-directed-bidir", action="store_true", help="Keep original directed bidirected behaviour")
    c.add_argument("--verbose", action="store_true")
    c.add_argument("--devices", metavar="I,J,...", help="Build on these GPUs of this host: the file is split at newline boundaries, one byte "
                   "range per GPU (B200 extension; default: one GPU)")
    c.add_argument("-o", "--output", metavar="PATH", help="Write graph pickle to PATH")
    e = sub.add_parser("export", help="Stream edges in simple formats")
    e.add_argument("gfa")
    e.add_argument("--format", default="edge-list", choices=["edge-list", "graphml", "gexf", "json"])
    e.add_argument("--bidirected", action="store_true")
    e.add_argument("--keep-directed-bidir", action="store_true", help="Keep original directed bidirected behaviour")
    e.add_argument("--output", help="Output path", default="-")
    d = sub.add_parser("distance", help="Compute distances")
    d.add_argument("gfa", help="Input *.gfa* file")
    g3 = d.add_mutually_exclusive_group(required=True)
    g3.add_argument("--seq", nargs=2, metavar=("SEQ_A", "SEQ_B"))
    g3.add_argument("--path", nargs=2, metavar=("PATH_A", "PATH_B"))
    g4 = d.add_mutually_exclusive_group()
    g4.add_argument("--directed", dest="directed", action="store_true", default=True)
    g4.add_argument("--undirected", dest="directed", action="store_false")
    d.add_argument("--backend", choices=["networkx", "igraph"], default="networkx", help="Graph backend to use")
    d.add_argument("--verbose", action="store_true")
    m = sub.add_parser("distance-matrix", help="Pairwise distances between paths")
    m.add_argument("gfa", help="Input *.gfa* file")
    m.add_argument("-o", "--output", required=True, help="Write matrix to PATH (.csv|.npy|.npz)")
    m.add_argument("--method", choices=["min", "mean"], default="min")
    m.add_argument("--backend", choices=["networkx", "igraph"], default="networkx", help="Graph backend to use")
    m.add_argument("--verbose", action="store_true")
    return ap


def main(argv: list[str] | None = None) -> None:
    ap = _parser()
    args = ap.parse_args(argv)
    if args.cmd == "export":
        if args.format != "edge-list":  # graphml / gexf / json go through a NetworkX object graph (cli.py:282-300)
            raise NotImplementedError(f"export --format {args.format} needs the NetworkX graph half of the reference")
        from .export import export_edge_list

        export_edge_list(args.gfa, bidirected=args.bidirected, output=args.output)
        return
    if args.cmd == "distance":
        if args.seq:  # needs the node sequences of the object graph (cli.py:303-316)
            raise NotImplementedError("distance --seq needs the NetworkX graph half of the reference (store_seq)")
        if args.backend != "networkx":
            raise NotImplementedError("backend='igraph' is outside the B200 path")
        from .analysis import genome_distance, load_paths

        paths = load_paths(args.gfa, raw_bytes=args.raw_bytes_id)  # cli.py:318
        name_a, name_b = args.path
        try:
            nodes_a = paths[name_a if not args.raw_bytes_id else name_a.encode()]
            nodes_b = paths[name_b if not args.raw_bytes_id else name_b.encode()]
        except KeyError as exc:  # cli.py:325-329
            msg = exc.args[0]
            raise SystemExit(f"unknown path: {msg.decode() if isinstance(msg, bytes) else msg}") from exc
        print(genome_distance(args.gfa, nodes_a, nodes_b, directed=args.directed, raw_bytes_id=args.raw_bytes_id))  # cli.py:330-334
        return
    if args.cmd == "distance-matrix":
        from .analysis import genome_distance_matrix

        M = genome_distance_matrix(args.gfa, method=args.method, raw_bytes_id=args.raw_bytes_id, backend=args.backend, verbose=args.verbose)
        try:
            save_matrix(M, Path(args.output), verbose=args.verbose, max_dense_gb=args.max_dense_gb)  # cli.py:343-350
        except MemoryError as exc:
            raise SystemExit(str(exc)) from exc
        return
    if not args.graph and not args.matrix:
        ap.error("convert requires --graph or --matrix")  # cli.py:194-195
    print(f"Using backend: {args.backend}")  # cli.py:198
    want_nodes = bool(args.matrix) and not args.no_node_map  # cli.py:222
    devs = [int(x) for x in args.devices.split(",")] if getattr(args, "devices", None) else None
    if devs and len(devs) > 1:
        # one graph over several GPUs: the result is already in its final compressed format (utils.py:47-48: `coo` keeps
        # whatever parse_gfa returned -- the CSR of max(S, S^T) in the default directed mode)
        from .utils import save_node_map

        mf = args.matrix_format.lower()
        if mf not in {"csr", "csc", "coo", "dok"}:
            raise ValueError("matrix-format must be csr|csc|coo|dok")
        res = parse_gfa(
            args.gfa, build_graph=args.graph, build_matrix=bool(args.matrix), directed=args.directed, weight_tag=args.weight_tag,
            strip_orientation=args.strip_orientation, verbose=args.verbose, bidirected=args.bidirected,
            keep_directed_bidir=args.keep_directed_bidir, backend=args.backend, dtype=args.dtype, asymmetric=args.asymmetric,
            raw_bytes_id=args.raw_bytes_id, return_node_list=want_nodes, split_on_alignment=args.split_on_alignment,
            matrix_format=mf if mf in ("csr", "csc") else None, devices=devs)
        A, nodes = res if want_nodes else (res, None)
        if mf == "dok":
            A = A.asformat("dok")
        try:
            save_matrix(A, Path(args.matrix), verbose=args.verbose, max_dense_gb=args.max_dense_gb)
        except MemoryError as exc:  # cli.py:247-248
            raise SystemExit(str(exc)) from exc
        if want_nodes:
            save_node_map(nodes, str(args.matrix) + ".nodes.tsv")  # cli.py:249-250
        return
    result = parse_gfa(
        args.gfa, build_graph=args.graph, build_matrix=bool(args.matrix), directed=args.directed,
        weight_tag=args.weight_tag, store_seq=args.store_seq, store_tags=args.store_tags,
        strip_orientation=args.strip_orientation, verbose=args.verbose, bidirected=args.bidirected,
        keep_directed_bidir=args.keep_directed_bidir, backend=args.backend, dtype=args.dtype,
        asymmetric=args.asymmetric, raw_bytes_id=args.raw_bytes_id, return_node_list=False,
        max_tag_mb=args.max_tag_mb, split_on_alignment=args.split_on_alignment)
    A = result
    tsv = None
    if want_nodes:
        # the node map file is made on the GPU ("<index>\t<name>\n", utils.py:108-114) instead of building a
        # Python list of names first; names are validated as UTF-8 where the reference decodes them
        sess = getattr(A, "_g2n_session", None)
        if sess is None or not sess.live():
            raise RuntimeError("the matrix is not backed by a live device build")
        tsv = sess.handle.fetch_nodes_tsv()
        if not args.raw_bytes_id:
            tsv.tobytes().decode()  # builders.py:287 node.decode() raises UnicodeDecodeError here
    A = convert_format(A, args.matrix_format, verbose=args.verbose, _untouched=True)  # cli.py:239 (A is exactly what parse_gfa returned)
    try:
        save_matrix(A, Path(args.matrix), verbose=args.verbose, max_dense_gb=args.max_dense_gb)
    except MemoryError as exc:  # cli.py:247-248
        raise SystemExit(str(exc)) from exc
    if want_nodes:
        if args.raw_bytes_id:
            tsv.tobytes().decode()  # utils.py:112 node.decode() raises here
        with open(str(args.matrix) + ".nodes.tsv", "wb") as fh:  # cli.py:249-250
            fh.write(memoryview(tsv))
