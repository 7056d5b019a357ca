/*
 * g2n.h -- C ABI of the B200-native GFA -> sparse adjacency path (libg2n.so).
 *
 * The reference (sclipman/gfa2network) is pure Python and has no FFI of its own; the
 * boundary it exposes for this path is
 *     gfa2network/builders.py:30-50   parse_gfa(path, *, build_matrix=True, directed, weight_tag,
 *                                     strip_orientation, bidirected, keep_directed_bidir, dtype,
 *                                     asymmetric, raw_bytes_id, return_node_list, ...)
 *     gfa2network/utils.py:40-63      convert_format(A, fmt)
 *     gfa2network/cli.py:193-250      `gfa2network convert --matrix ... --matrix-format ...`
 * Each entry point below names the reference lines it replaces.  The Python host mirror
 * (gfa2network_b200/builders.py, utils.py, cli.py) binds these with ctypes; INTEGRATION.md
 * shows the stub a maintainer of the reference would add.
 *
 * Conventions: extern "C", plain pointers and sizes, int status returns (0 = G2N_OK).
 * The caller allocates every output buffer after g2n_sizes(); the library never frees
 * caller memory and owns all device scratch for the handle's lifetime.  A handle is not
 * thread-safe; separate handles may be used from separate threads.  There is no CPU
 * fallback: without a CUDA device g2n_create fails with G2N_ERR_CUDA.
 */
#ifndef G2N_H
#define G2N_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define G2N_ABI_VERSION 2

/* ---- status codes ------------------------------------------------------------------- */
#define G2N_OK 0
#define G2N_ERR_CUDA 1        /* CUDA runtime failure: g2n_last_error() has the driver string */
#define G2N_ERR_INVALID 2     /* bad argument / call order */
#define G2N_ERR_PARSE 3       /* the input holds a record the reference raises on: see g2n_diag */
#define G2N_ERR_UNSUPPORTED 4 /* declared out of scope (e.g. > 2^31-1 nodes, weight tag > 64 bytes) */
#define G2N_ERR_INTERNAL 5
#define G2N_ERR_RETRY 6       /* multi-GPU: a remembered capacity was exceeded on some rank (same verdict on every
                                 rank): repeat the build with g2n_dist_probe + a host-planned pass */

/* ---- first-error kinds (g2n_diag.err_kind); map 1:1 to the reference's exceptions ------ */
#define G2N_PE_NONE 0
#define G2N_PE_MALFORMED_L 1      /* ValueError("Malformed L record")   parser.py:208-209 */
#define G2N_PE_MALFORMED_E 2      /* ValueError("Malformed E record")   parser.py:251-252 */
#define G2N_PE_MALFORMED_C 3      /* ValueError("Malformed C record")   parser.py:299-300 */
#define G2N_PE_MALFORMED_P 4      /* ValueError("Malformed P record")   parser.py:231-232 */
#define G2N_PE_MALFORMED_O 5      /* ValueError("Malformed O record")   parser.py:345-346 */
#define G2N_PE_S_NO_ID 6          /* IndexError  fields[1]              parser.py:163 */
#define G2N_PE_COMPACT_EMPTY 7    /* IndexError  u_field[-1] on b""     parser.py:220-221 */
#define G2N_PE_ORI_UTF8 8         /* UnicodeDecodeError  ori.decode()   parser.py:214,291,293,337,339 */
#define G2N_PE_WEIGHT_OVERFLOW 9  /* OverflowError float(int)           builders.py:209 */
#define G2N_PE_UNSUPPORTED_NUM 10 /* non-ASCII numeric weight value: outside parity scope, loud */

/* ---- enums -------------------------------------------------------------------------- */
#define G2N_DTYPE_F64 0
#define G2N_DTYPE_F32 1
#define G2N_DTYPE_I32 2
#define G2N_DTYPE_I8 3
#define G2N_DTYPE_BOOL 4

#define G2N_FMT_NATIVE 0 /* what parse_gfa returns: CSR for sym-max modes, raw COO otherwise (builders.py:281-283) */
#define G2N_FMT_CSR 1    /* convert_format(parse_gfa(...), "csr")  (utils.py:55) */
#define G2N_FMT_CSC 2    /* convert_format(parse_gfa(...), "csc") */
#define G2N_FMT_COO 3    /* result format code only: raw COO triplets in emission order */

typedef struct g2n_handle g2n_handle;

/* Mirrors the keyword arguments of parse_gfa that reach the matrix path (builders.py:30-50). */
typedef struct g2n_params {
    int32_t directed;            /* builders.py:35 */
    int32_t bidirected;          /* builders.py:41 */
    int32_t keep_directed_bidir; /* builders.py:42 */
    int32_t asymmetric;          /* builders.py:45 */
    int32_t strip_orientation;   /* builders.py:39 */
    int32_t dtype;               /* G2N_DTYPE_*  (builders.py:44, cli.py:92-97) */
    int32_t want_format;         /* G2N_FMT_NATIVE | G2N_FMT_CSR | G2N_FMT_CSC */
    int32_t text_on_device;      /* 1: `text` is a device pointer on the handle's device */
    const uint8_t *weight_tag;   /* UTF-8 bytes of weight_tag or NULL (builders.py:36, 205-209) */
    int32_t weight_tag_len;
    int32_t reserved;
} g2n_params;

typedef struct g2n_sizes_t {
    uint64_t n_nodes;
    uint64_t nnz;          /* stored entries of the result (raw triplets for COO) */
    uint64_t names_bytes;  /* total bytes of all node names */
    int32_t format;        /* G2N_FMT_CSR | G2N_FMT_CSC | G2N_FMT_COO */
    int32_t index_bytes;   /* 4 (SciPy picks int32 below 2^31-1; larger is G2N_ERR_UNSUPPORTED) */
    int32_t dtype;         /* G2N_DTYPE_* of the data array */
    int32_t reserved;
    uint64_t slab_rows;    /* rows held by this handle (== n_nodes except for a multi-GPU slab) */
    uint64_t names_count;  /* node names held by this handle (== n_nodes except for a multi-GPU slab: the nodes
                              whose first appearance is in this rank's shard) */
    uint64_t names_id0;    /* ID of the first of them (0 on one GPU) */
} g2n_sizes_t;

typedef struct g2n_diag {
    int32_t err_kind;        /* G2N_PE_* of the first offending line in file order, 0 if none */
    int32_t unknown_byte;    /* first byte of the first unsupported record (parser.py:125-131) that
                                precedes the error line, -1 if none */
    uint64_t err_offset;     /* byte offset of the start of the offending line */
    uint64_t unknown_offset; /* byte offset of that unsupported record */
    uint64_t n_records;      /* records the reference's parser would have yielded (S L E C P O) */
    uint64_t n_edge_records; /* L + E + C */
    uint64_t n_triplets;     /* COO triplets emitted (builders.py:222-228) */
    uint64_t n_long_keys;    /* node keys longer than the 15-byte inline form */
    uint32_t retries;        /* capacity / hash-seed retries of the last build */
    uint32_t gpu_launches;   /* kernels launched by the last build */
    float ms_total;          /* device time of the last build, CUDA events on the handle's stream */
    float ms_h2d;            /* host->device copy of the text (0 when text_on_device) */
    float ms_stage[8];       /* scan+hash, ids, names, emit, sort, reduce, (spare) */
    uint32_t warn_flags;     /* G2N_WARN_* */
    uint32_t speculative;    /* 1: the build ran without a host round trip after the tokenizer (buffers
                                sized from the previous build of the same input size and mode) */
} g2n_diag;

#define G2N_WARN_CAST_OVERFLOW 1u /* a finite float64 weight became inf in the float32 cast: NumPy's
                                     RuntimeWarning("overflow encountered in cast") at builders.py:281 */

/* Create a handle on CUDA device `device`.  Fails (G2N_ERR_CUDA) without a usable GPU. */
int g2n_create(int device, g2n_handle **out);
void g2n_destroy(g2n_handle *h);

/* Use an existing CUDA stream (cudaStream_t) for all work of this handle.  NULL is the CUDA default
 * stream (e.g. PyTorch's default stream); (void*)-1 selects the handle's own non-blocking stream,
 * which is also the initial state. */
int g2n_set_stream(g2n_handle *h, void *cuda_stream);

/* Pinned host staging memory (so callers can read files straight into DMA-able buffers). */
void *g2n_host_alloc(uint64_t nbytes);
void g2n_host_free(void *p);

/* The hot path.  Replaces the tokenizer loop (parser.py:114-176), node-ID assignment and triplet
 * emission (builders.py:163-234), coo_matrix + maximum(A, A.T) (builders.py:279-283) and, when
 * want_format is CSR/CSC, convert_format (utils.py:40-63).  Results stay resident on the device
 * until the next build or g2n_destroy.  On G2N_ERR_PARSE consult g2n_status(). */
int g2n_build(g2n_handle *h, const uint8_t *text, uint64_t nbytes, const g2n_params *p);

/* Same build with the text read from a regular file (the reference opens the path itself,
 * parser.py:100-112): reader threads pread() 8 MiB pieces into pinned staging buffers, the pieces are
 * copied to the device as they arrive and tokenized behind the copy.  `p->text_on_device` is ignored.
 * stdin stays with the caller (read on the host, then g2n_build). */
int g2n_build_file(g2n_handle *h, const char *path, const g2n_params *p);

/* Same build from a gzip-compressed file (parser.py:108-109 gzip.open): the compressed file is mapped and inflated
 * in 64 MiB windows into two pinned staging buffers; every finished window is copied to the device while the next one
 * is inflated.  BGZF files (bgzip: independent blocks that announce their size in a gzip extra field) are inflated
 * by all host cores, any other gzip stream (one or several members) by one core.  CRC32 / ISIZE of BGZF blocks and
 * zlib's own checks of plain streams are honoured.  G2N_ERR_INVALID: the container is damaged or not gzip -- the caller
 * falls back to the reference's own host inflate, which raises the reference's exception. */
int g2n_build_gz(g2n_handle *h, const char *path, const g2n_params *p);

/* Convert the device-resident raw COO of the last build to CSR/CSC in place of calling
 * convert_format(A, fmt) on the host (utils.py:55).  No-op if already in that format. */
int g2n_convert(g2n_handle *h, int32_t want_format);

int g2n_sizes(g2n_handle *h, g2n_sizes_t *out);

/* Copy the result to caller memory.  CSR/CSC: a0 = indptr (n_nodes+1), a1 = indices (nnz);
 * COO: a0 = row (nnz), a1 = col (nnz); data = nnz elements of the result dtype. */
int g2n_fetch_matrix(g2n_handle *h, void *a0, void *a1, void *data);

/* Total bytes of all node names.  The name table is sized lazily (callers that never ask for the node
 * list pay nothing); g2n_sizes().names_bytes is valid after this call. */
int g2n_names_bytes(g2n_handle *h, uint64_t *out);

/* Node names in ID order (builders.py:284-288): `names` = names_bytes bytes,
 * `offsets` = n_nodes+1 uint64 offsets into it. */
int g2n_fetch_names(g2n_handle *h, uint8_t *names, uint64_t *offsets);

/* The node map as text, "<index>\t<name>\n" per node in ID order -- the bytes save_node_map writes
 * (utils.py:108-114) -- produced on the device.  g2n_nodes_tsv_bytes sizes it, g2n_fetch_nodes_tsv copies it. */
int g2n_nodes_tsv_bytes(g2n_handle *h, uint64_t *out);
int g2n_fetch_nodes_tsv(g2n_handle *h, uint8_t *out);

/* `export --format edge-list` (cli.py:267-281): "<from>\t<to>\n" per L / E / C record in file order, produced on
 * the device from the last single-GPU build.  The endpoints are that build's node keys, so build with the same
 * `bidirected` flag as the export (from:orientation / to:orientation) and without strip_orientation. */
int g2n_edge_list_bytes(g2n_handle *h, uint64_t *out);
int g2n_fetch_edge_list(g2n_handle *h, uint8_t *out);

/* Distances on the resident CSR/CSC result of the last single-GPU build (analysis.py:116-161, 180-272: the
 * reference runs networkx.multi_source_dijkstra_path_length on the graph it builds WITHOUT a weight tag, i.e.
 * hop counts along out-edges).  g2n_bfs: multi-source breadth-first search from `sources` (node IDs) into
 * level slot `slot` of `n_slots` (all levels of one search run inside one cooperative kernel);
 * g2n_levels_reduce: out3 = {min level or -1, sum of levels, count} over the reachable nodes of a node list
 * (duplicates count as often as they occur); g2n_fetch_levels: the n_nodes hop counts of a slot, -1 = unreachable. */
int g2n_bfs(g2n_handle *h, const int32_t *sources, uint64_t n_sources, int32_t slot, int32_t n_slots);
int g2n_levels_reduce(g2n_handle *h, int32_t slot, const int32_t *nodes, uint64_t n_nodes, int64_t *out3);
int g2n_fetch_levels(g2n_handle *h, int32_t slot, int32_t *out);

/* The P / O records of the last build's text (analysis.py:164-177 over parser.py:229-247, 343-361), resolved on the
 * device: g2n_paths_load finds the records (file order) and turns every "<segment>[+-]" entry of their lists into the
 * node ID of that segment (-1: not a node); g2n_path_info describes record i (name = bytes of the text; the first
 * entry that is not a node, which the reference answers with NodeNotFound); g2n_path_bfs / g2n_path_reduce are g2n_bfs /
 * g2n_levels_reduce with record i's node list, which never leaves the device; g2n_fetch_path_nodes copies it out. */
typedef struct g2n_path_info_t {
    uint64_t line_offset;    /* first byte of the record */
    uint64_t name_offset;    /* fields[1] */
    uint64_t n_entries;      /* entries of fields[2].split(",") */
    int64_t missing_entry;   /* index of the first entry that is not a node, -1 if all are */
    uint64_t missing_offset; /* that entry's name (sign stripped) in the text */
    uint32_t name_len;
    uint32_t missing_len;
} g2n_path_info_t;
int g2n_paths_load(g2n_handle *h, uint64_t *n_paths);
int g2n_path_info(g2n_handle *h, uint64_t i, g2n_path_info_t *out);
int g2n_path_bfs(g2n_handle *h, uint64_t i, int32_t slot, int32_t n_slots);
int g2n_path_reduce(g2n_handle *h, int32_t slot, uint64_t i, int64_t *out3);
int g2n_fetch_path_nodes(g2n_handle *h, uint64_t i, int32_t *out);
/* `len` bytes of the device-resident text of the last build (names of records, offending entries) */
int g2n_fetch_text(g2n_handle *h, uint64_t offset, uint64_t len, uint8_t *out);

/* Device pointers of the resident result (for device-side consumers / benchmarks). */
int g2n_device_result(g2n_handle *h, void **a0, void **a1, void **data);

/* Per-kernel timing of the last build (CUDA event pair around every launch, on the handle's
 * stream).  g2n_set_profile(h, 1) turns it on; g2n_kernel_times fills up to `cap` entries
 * (one per kernel name: summed milliseconds and launch count) and returns how many. */
typedef struct g2n_ktime {
    char name[40];
    float ms;
    uint32_t launches;
} g2n_ktime;
int g2n_set_profile(g2n_handle *h, int on);

/* A handle remembers the sizes (nodes, records, edge records) of its last build.  A build of the same
 * input size and mode then runs speculatively: every buffer is sized up front, size-dependent kernels
 * read the actual sizes from device memory, and the host synchronises once, at the end; if the sizes
 * do not fit the build is repeated with the usual host round trip after the tokenizer.  The result
 * is identical either way.  On by default; g2n_set_speculation(h, 0) turns it off. */
int g2n_set_speculation(g2n_handle *h, int on);
int g2n_kernel_times(g2n_handle *h, g2n_ktime *out, int cap);

int g2n_status(g2n_handle *h, g2n_diag *out);
const char *g2n_last_error(g2n_handle *h);
int g2n_abi_version(void);

/* How a build with `entries` row entries of `entry_bytes` (4: no weights, 8: with) over `rows` rows would organise stage 4
 * (scipy/_coo.py tocsr's counting sort by row, builders.py:283 / utils.py:55): host logic only, needs no device.
 *   out[0] passes of the flat kernels (1: row arrays fit L2)   out[1] 1: entries are partitioned by row bucket first
 *   out[2] buckets, out[3] log2(rows per bucket)               out[4] 1: second level (sub-buckets placed in shared memory)
 *   out[5] sub-buckets, out[6] log2(rows per sub-bucket), out[7] entries a sub-bucket may hold in shared memory */
int g2n_plan_row_buckets(uint64_t entries, uint64_t rows, int entry_bytes, uint32_t out[8]);

/* Standalone stage 4 (SciPy COO -> CSR/CSC with duplicate summing, utils.py:55 on a user matrix):
 * host COO arrays in, host compressed arrays out.  *nnz_out receives the stored-entry count;
 * indptr has n+1 entries; indices/data need capacity nnz_in. */
int g2n_coo_to_compressed(g2n_handle *h, const int32_t *row, const int32_t *col, const void *data,
                          uint64_t nnz_in, uint64_t n, int32_t dtype, int32_t want_format,
                          int32_t *indptr, int32_t *indices, void *data_out, uint64_t *nnz_out);

/* ---- multi-GPU build -------------------------------------------------------
 * One process and one handle per GPU; each rank holds a newline-aligned byte range of the text (file
 * order = rank order).  The result on every rank is the CSR/CSC slab of its row block
 * [rank * ceil(n / world), ...) of the matrix parse_gfa + convert_format return on one GPU
 * (builders.py:190-234, 279-283; utils.py:55), and the names of the nodes that first appear in its shard
 * (consecutive IDs from g2n_sizes().names_id0).  The library moves the data itself: kernels write into
 * the peers' exchange arenas over NVLink and signal through flags in the peers' control blocks
 * (gfa2network_b200/csrc/dist.cuh); the caller only bootstraps -- it agrees on capacities, passes the
 * CUDA IPC handles around (any out-of-band channel; dist.py uses torch.distributed) and queues the stages.
 *
 *   g2n_dist_init        rank / world of this handle
 *   g2n_dist_probe       host-planned pass: tokenize the shard, report its counts (or its parse error)
 *   g2n_dist_plan        capacities every rank agreed on: keys per (source, owner) segment, row entries per
 *                        (source, row owner) segment, rows and received entries of the slab (the last two
 *                        are only used by speculative builds).  dry_run: only report whether the arena
 *                        would have to grow (peers must close their mappings first)
 *   g2n_dist_local_mem   this rank's arena / control block: pointers (ranks sharing a process) and the
 *                        two 64-byte CUDA IPC handles (ranks in other processes)
 *   g2n_dist_set_peers / g2n_dist_open_peers / g2n_dist_close_peers
 *   g2n_dist_stage       queue stage 0..6 on the handle's stream (no host synchronisation in a speculative
 *                        build; `text`, `nbytes`, `p` are read by stage 0 only).  speculative = 1: the shard is
 *                        tokenized inside stage 0 with the sizes remembered from the previous build;
 *                        0: the tokenizer results of g2n_dist_probe are used
 *   g2n_dist_finish      the one host round trip: G2N_OK, or G2N_ERR_RETRY on every rank
 * Ranks that share a process (tests: several logical ranks on one GPU) must queue stage k on every rank
 * before stage k+1 on any.  Restriction of this version: <= 8 ranks.  With a weight tag the row entries
 * travel in emission order together with their weight, so duplicate sums match the single-GPU result bit for bit.
 * Node names longer than
 * the 15-byte inline key are exchanged as their 128-bit tagged hash (equality of two DIFFERENT long names on
 * different ranks is decided by that hash alone; inside a shard the bytes are compared as on one GPU). */
typedef struct g2n_dist_info {
    uint64_t n_keys;         /* distinct node keys in this shard */
    uint64_t n_tiles;
    uint64_t n_records;
    uint64_t n_edge_records;
    uint64_t n_entries;      /* row entries this rank will send (triplets, x2 in the max(S,S^T) mode) */
    uint64_t reserved;
} g2n_dist_info;

typedef struct g2n_dist_result {
    uint64_t bad;            /* 0, or the DXB_* status bits (dist.cuh) that made the build repeat */
    uint64_t n_global;       /* nodes of the whole graph */
    uint64_t row0, n_rows;   /* this rank's row block */
    uint64_t nnz;            /* stored entries of the slab */
    uint64_t n_recv;         /* row entries received */
    uint64_t n_first, id0;   /* nodes first seen in this shard; ID of the first of them */
    uint64_t n_keys, n_records, n_edge_records; /* this shard */
    uint64_t keys_to[8];     /* keys sent to each owner rank */
    uint64_t pairs_to[8];    /* row entries sent to each row owner */
} g2n_dist_result;

int g2n_dist_init(g2n_handle *h, int rank, int world);
/* ranks of one process on different GPUs: peer access from this handle's device to peer_device (their arenas are then
 * exchanged as plain pointers, g2n_dist_local_mem / g2n_dist_set_peers) */
int g2n_dist_enable_peer(g2n_handle *h, int peer_device);
/* bytes [offset, offset + nbytes) of a file -> this handle's device text buffer (*dev_text: 16-byte aligned, valid until
 * the handle loads another text).  The multi-GPU counterpart of g2n_build_file: every rank loads its own newline-aligned
 * range (parser.py:111 reads the file; SURVEY 8e step 1 splits it) */
int g2n_load_file_range(g2n_handle *h, const char *path, uint64_t offset, uint64_t nbytes, void **dev_text);
int g2n_dist_probe(g2n_handle *h, const uint8_t *text, uint64_t nbytes, const g2n_params *p, g2n_dist_info *out);
int g2n_dist_plan(g2n_handle *h, uint64_t key_cap, uint64_t pair_cap, uint64_t rows_cap, uint64_t recv_cap,
                  int dry_run, int *will_realloc);
int g2n_dist_local_mem(g2n_handle *h, void **arena, void **ctl, uint8_t *ipc128);
int g2n_dist_set_peers(g2n_handle *h, void *const *arenas, void *const *ctls);
int g2n_dist_open_peers(g2n_handle *h, const uint8_t *ipc_all /* world x 128 bytes, rank order */);
int g2n_dist_close_peers(g2n_handle *h);
int g2n_dist_stage(g2n_handle *h, int stage, const uint8_t *text, uint64_t nbytes, const g2n_params *p, int speculative);
int g2n_dist_finish(g2n_handle *h, g2n_dist_result *out);

#ifdef __cplusplus
}
#endif
#endif /* G2N_H */
