// g2n.cu -- host orchestration and the C ABI (include/g2n.h) of the B200-native GFA -> sparse
// adjacency path.  Stages on one CUDA stream:
//   K1  k_tokenize        line scan + field split + node-key hashing + edge-record emission (tokenize.cuh)
//   K2  k_mark_first / k_assign_ids / k_gather_names   first-appearance ranking -> node IDs   (ids.cuh)
//   K3  k_emit_coo | k_emit_keys                        COO triplets / sort keys               (ids.cuh)
//   K4  k_rows_count / k_rows_scatter / k_rows_big / k_rows_sort / k_rows_write   row bucketing + in-row sort + dedup/sum (+max) (rowsort.cuh)
// No CPU fallback exists: every entry point needs a CUDA device.
#include <cuda_runtime.h>
#include <stdio.h>
#include <string.h>

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>

#include <atomic>
#include <memory>
#include <string>
#include <thread>
#include <vector>

#include "../../include/g2n.h"
#include "dist.cuh"
#include "bfs.cuh"
#include "paths.cuh"

#include <algorithm>

using namespace g2n;

namespace {

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes)
    {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) { e = cudaMalloc(&p, bytes); want = bytes; }
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release()
    {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    template <typename T> T* as() const { return (T*)p; }
};

struct KTimer {
    const char* name;
    cudaEvent_t a, b;
};

enum { EV_START = 0, EV_H2D, EV_TOKENIZE, EV_IDS, EV_EMIT, EV_SORT, EV_REDUCE, EV_COUNT };

}  // namespace

struct g2n_handle {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t own_stream = nullptr;
    cudaStream_t copy_stream = nullptr;  // host -> device copy of the text in pieces, overlapped with the tokenizer
    std::vector<cudaEvent_t> copy_ev;
    std::string err;
    cudaEvent_t ev[EV_COUNT];
    // device buffers (kept between builds: a warm handle allocates nothing)
    DevBuf text, defer, edge_slots, edge_w, longs, tile_info, tile_base, wprefix, slot_id, id2slot, name_off, names;
    DevBuf rowptr, cursor, entries, w_emit, biglist, ucnt, indptr, indices, data, row, col, scan_state;
    DevBuf up_row, up_col, up_data, tsv, tsv_off, el_len, el_off, el_text, emit_t0;
    DevBuf pair_major, pair_ent, bucket_ctl, pair2_major, pair2_ent, sub_ctl, sub_off;  // bucketed row build (rowsort.cuh: RowBuckets, SubBuckets)
    bool rows_presorted = false;  // k_sub_rows_place sorted the short rows and wrote ucnt: rows_finalize runs k_rows_sort_rest
    u32 place_attr = 0;  // k_sub_rows_place instantiations whose dynamic shared memory opt-in was set on this device
    DevBuf bfs_levels, bfs_q0, bfs_q1, bfs_ctl, bfs_nodes, bfs_out;  // distances on the resident CSR (bfs.cuh)
    int bfs_slots = 0;
    u64 bfs_n = 0;
    // P / O records of the last build's text, resolved to node IDs on the device (paths.cuh)
    DevBuf path_starts, path_recs, path_cnt, path_off, path_ids, path_misc;
    std::vector<PathRec> h_paths;
    std::vector<u64> h_path_entry0;  // n_paths + 1
    std::vector<u64> h_path_blkoff;
    bool paths_ready = false;
    u64 seed_used = 0;
    bool el_ready = false;
    u64 el_bytes = 0;
    bool tsv_ready = false;
    u64 tsv_bytes = 0;
    // zero-initialised state, one memset per arena and build:
    //   zearly  hash table (keys | first | rep), counters + DevSizes, look-back state of the tile scan
    //   zids    first-appearance bitmap, look-back state of its scan
    //   zrows   row histogram, long-row counter, look-back state of the two row scans
    DevBuf zearly, zids, zrows;
    Slot* d_slots = nullptr;
    u32* d_slot_cnt = nullptr;  // row counters bumped by the tokenizer (NULL: the build counts with a pass over the edge records)
    int count_mode = CM_NONE;
    Ctl* d_ctl = nullptr;
    Counters* d_cnt = nullptr;
    DevSizes* d_ds = nullptr;
    u64* d_scan_tiles = nullptr;
    u32* d_bitmap = nullptr;
    u64* d_scan_words = nullptr;
    u32* d_rowcnt = nullptr;
    u32* d_bigcount = nullptr;
    u64* d_scan_rows[2] = {nullptr, nullptr};
    size_t zids_bytes = 0, zrows_bytes = 0;
    Ctl* h_ctl = nullptr;       // pinned copy of the counters + device-side sizes
    Counters* h_cnt = nullptr;  // = &h_ctl->c
    u64* h_tail = nullptr;      // pinned: {-, names_bytes, tile_base total, -}
    // capacity hints learnt from previous builds
    u64 hint_keys = 0, hint_edges = 0, hint_long = 0, hint_defer = 0, hint_records = 0;
    // A build whose shape (input size + mode) equals the previous one runs speculatively: buffers are
    // sized from the hints, every size-dependent kernel reads the actual sizes from DevSizes, and the
    // host looks at the counters once, at the end.  A miss (DevSizes.ok == 0) re-runs the build with
    // the host round trip after the tokenizer.
    u64 hint_sig = 0;
    bool hint_valid = false;
    bool speculate = true;  // g2n_set_option("speculate", 0) turns it off
    bool spec = false;      // the current build is speculative
    bool slow_ran = false;
    bool reseeded = false;  // the last tokenizer pass changed the hash seed of long keys (collision retry)
    u32 tk_attr = 0;     // tokenizer specialisations whose dynamic shared memory opt-in was set on this device
    u32 n_pieces = 0;    // host text of the current build: copy pieces still to be waited for (0: text is on the device)
    u64 piece_bytes = 0;
    // file source (g2n_build_file): reader threads pread() pieces into pinned staging buffers and queue the copies
    int src_fd = -1;
    bool src_resident = false;  // the file is already in h->text (second pass of the same g2n_build_file call)
    std::vector<void*> stage;
    std::vector<cudaEvent_t> stage_ev;
    void* gz_stage[2] = {nullptr, nullptr};  // pinned 64 MiB windows of inflated text (g2n_build_gz)
    cudaEvent_t gz_ev[2] = {nullptr, nullptr};
    std::vector<std::thread> readers;
    std::unique_ptr<std::atomic<int>[]> piece_issued;
    std::atomic<int> reader_err{0};
    bool gang_scan = true;  // scans run as one co-resident gang (cooperative launch); cleared if the launch is refused
    u64 cap_n = 0, cap_E = 0, cap_R = 0;  // what this build's buffers were sized for
    // state of the last build
    g2n_params params;
    uint8_t weight_tag[64];
    const uint8_t* d_text = nullptr;
    u64 nbytes = 0;
    bool built = false;
    bool have_edges = false;
    u64 n_nodes = 0, nnz = 0, names_bytes = 0, n_edges = 0, n_triplets = 0, n_records = 0, n_long = 0;
    // multi-GPU state
    bool slab_mode = false;
    u64 slab_rows = 0, n_global = 0;
    u64 names_n = 0, names_id0 = 0;  // slab mode: this rank names the IDs [names_id0, names_id0 + names_n)
    bool dx_inited = false, dx_probed = false, dx_spec = false;
    DxPeers dxp;   // rank, world, epoch, peer arenas / control blocks
    DxLayout dxl;  // layout of every rank's exchange arena
    DevBuf dx_arena, dx_ctl, dx_loc, dx_zg, dx_gslot, dx_gpos, dx_sent, dx_tcnt, dx_toff;
    DxLocal* h_loc = nullptr;  // pinned copy of the build's local status
    bool dx_peer_open[DX_MAXW] = {false, false, false, false, false, false, false, false};
    u32 dx_gcap = 0;
    u64 dx_rows_cap = 0, dx_recv_cap = 0;
    u32 table_cap = 0;
    u32 n_tiles = 0;
    int tpe = 1, spe = 2;
    bool symmax = false;
    int result_format = G2N_FMT_COO;
    bool names_ready = false;
    bool names_sized = false;
    bool edges_are_ids = false; // edge_slots were translated to node IDs in place
    g2n_diag diag;
    u32 launches = 0;
    // optional per-kernel timing (g2n_set_profile): one event pair per launch
    bool profile = false;
    std::vector<KTimer> ktimers;
    size_t kt_used = 0;
};

namespace {

#define CK(call)                                                                      \
    do {                                                                              \
        cudaError_t _e = (call);                                                      \
        if (_e != cudaSuccess) {                                                      \
            char _b[512];                                                             \
            snprintf(_b, sizeof _b, "%s failed at %s:%d: %s", #call, __FILE__, __LINE__, cudaGetErrorString(_e)); \
            h->err = _b;                                                              \
            return G2N_ERR_CUDA;                                                      \
        }                                                                             \
    } while (0)

// Brackets one kernel launch: counts it and, in profile mode, times it with an event pair.
struct KScope {
    g2n_handle* h;
    bool on;
    KScope(g2n_handle* h_, const char* name) : h(h_), on(h_->profile)
    {
        h->launches++;
        if (!on) return;
        if (h->kt_used == h->ktimers.size()) {
            KTimer t;
            t.name = name;
            cudaEventCreate(&t.a);
            cudaEventCreate(&t.b);
            h->ktimers.push_back(t);
        }
        h->ktimers[h->kt_used].name = name;
        cudaEventRecord(h->ktimers[h->kt_used].a, h->stream);
    }
    ~KScope()
    {
        if (!on) return;
        cudaEventRecord(h->ktimers[h->kt_used].b, h->stream);
        h->kt_used++;
    }
};

inline u32 grid_for(u64 work_items, u32 per_block, u32 waves = 8)
{
    u64 blocks = (work_items + per_block - 1) / per_block;
    u64 cap = (u64)G2N_SM_COUNT * waves;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (u32)blocks;
}

inline u32 next_pow2(u64 x)
{
    u64 p = 1;
    while (p < x) p <<= 1;
    return (u32)p;
}

// exclusive scan launcher: out[0..n], out[n] = total (also into out2 if given).  n_cap bounds n on the
// host (grid, look-back state); n_dev, if given, holds the actual n on the device.  `state` is a
// zeroed region of n_cap / SCAN_TILE + 3 words out of an arena, or NULL: memset a private one.
inline size_t scan_state_bytes(u64 n_cap) { return ((n_cap + SCAN_TILE - 1) / SCAN_TILE + 3) * sizeof(u64); }

template <typename Tout, class LoadOp>
int launch_scan(g2n_handle* h, LoadOp load, Tout* out, Tout* out2, u64 n_cap, const u32* n_dev, u64* state_region)
{
    const u64 n_tiles = (n_cap + SCAN_TILE - 1) / SCAN_TILE;
    if (!state_region) {
        CK(h->scan_state.ensure(scan_state_bytes(n_cap)));
        CK(cudaMemsetAsync(h->scan_state.p, 0, scan_state_bytes(n_cap), h->stream));
        state_region = h->scan_state.as<u64>();
    }
    const u32 grid = grid_for(n_tiles, 1, 4);
    if (h->gang_scan && getenv("G2N_DBG_NOGANG")) h->gang_scan = false;  // timing experiments: look-back scan
    if (h->gang_scan) {
        // all CTAs co-resident (cooperative launch): reduce, grid barrier, scan -- no look-back chain
        KScope ks(h, "k_scan_gang");
        u64 n_host = n_cap;
        void* args[] = {(void*)&load, (void*)&out, (void*)&out2, (void*)&n_host, (void*)&n_dev, (void*)&state_region};
        cudaError_t e = cudaLaunchCooperativeKernel((const void*)k_scan_gang<Tout, LoadOp>, dim3(grid), dim3(256), args, 0, h->stream);
        if (e == cudaSuccess) return G2N_OK;
        (void)cudaGetLastError();
        h->gang_scan = false;  // not launchable as a gang on this device / context: use the look-back scan from now on
    }
    u64* state = state_region + 1;
    u32* ticket = (u32*)state_region;
    { KScope ks(h, "k_scan_exclusive"); k_scan_exclusive<Tout, LoadOp><<<grid, 256, 0, h->stream>>>(load, out, out2, n_cap, n_dev, state, ticket); }
    CK(cudaGetLastError());
    return G2N_OK;
}

// zrows arena: row histogram | long-row counter | look-back state of the rowptr and indptr scans
int layout_zrows(g2n_handle* h, u64 n_cap)
{
    const size_t a = ((n_cap + 2) * sizeof(u32) + 255) & ~(size_t)255;
    const size_t b = 256;
    const size_t c = (scan_state_bytes(n_cap) + 255) & ~(size_t)255;
    const size_t d = c;  // look-back state of the indptr scan
    CK(h->zrows.ensure(a + b + c + d));
    uint8_t* base = h->zrows.as<uint8_t>();
    h->d_rowcnt = (u32*)base;
    h->d_bigcount = (u32*)(base + a);
    h->d_scan_rows[0] = (u64*)(base + a + b);
    h->d_scan_rows[1] = (u64*)(base + a + b + c);
    h->zrows_bytes = a + b + c + d;
    return G2N_OK;
}

size_t dtype_size(int dt)
{
    switch (dt) {
        case G2N_DTYPE_F64: return 8;
        case G2N_DTYPE_F32: return 4;
        case G2N_DTYPE_I32: return 4;
        default: return 1;
    }
}

// rowptr (u32, rows+1) and entries (ENT, M) are ready, long rows are listed: sort every row, sum
// duplicates, write the result.  M and n are host-side bounds (exact or capacities); the kernels read the
// actual row count from DevSizes.
template <typename T, class ENT>
int rows_finalize_typed(g2n_handle* h, u64 M, u64 n, int sym, const WEmit w_emit, const T* w_typed)
{
    typedef typename ENT::type E;
    const u32* n_dev = &h->d_ds->rows;
    CK(h->indptr.ensure((n + 2) * sizeof(int32_t)));
    CK(h->indices.ensure((M + 1) * sizeof(int32_t)));
    CK(h->data.ensure((M + 1) * sizeof(T)));
    { KScope ks(h, "k_rows_big"); k_rows_big<ENT><<<G2N_SM_COUNT, 256, 0, h->stream>>>(h->rowptr.as<u32>(), h->biglist.as<u32>(), h->d_bigcount, h->entries.as<E>()); }
    const u64 n_chunks = (n + RF_ROWS - 1) / RF_ROWS;
    CK(h->ucnt.ensure((n + 2) * sizeof(u32)));
    if (h->rows_presorted) { KScope ks(h, "k_rows_sort_rest"); k_rows_sort_rest<T, ENT><<<grid_for(n, 256), 256, 0, h->stream>>>(h->rowptr.as<u32>(), h->entries.as<E>(), n_dev, sym, w_emit, w_typed, h->ucnt.as<u32>()); }
    else { KScope ks(h, "k_rows_sort"); k_rows_sort<T, ENT><<<grid_for(n_chunks, 1, 8), RF_ROWS, 0, h->stream>>>(h->rowptr.as<u32>(), h->entries.as<E>(), n_dev, sym, w_emit, w_typed, h->ucnt.as<u32>()); }
    h->rows_presorted = false;
    CK(cudaGetLastError());
    LoadArray<u32> ldu{h->ucnt.as<u32>()};
    int rc = launch_scan<int32_t>(h, ldu, h->indptr.as<int32_t>(), nullptr, n, n_dev, h->d_scan_rows[1]);
    if (rc) return rc;
    { KScope ks(h, "k_rows_write"); k_rows_write<T, ENT><<<grid_for(n_chunks, 1, 8), RF_ROWS, 0, h->stream>>>(h->rowptr.as<u32>(), h->entries.as<E>(), n_dev, sym, w_emit, w_typed, h->indptr.as<int32_t>(), h->indices.as<int32_t>(), h->data.as<T>(), &h->d_ds->nnz); }
    CK(cudaGetLastError());
    return G2N_OK;
}

// weighted: 64-bit entries that carry the emission index; unweighted: 32-bit entries
int rows_finalize(g2n_handle* h, int dtype, bool weighted, u64 M, u64 n, int sym, const WEmit w_emit, const void* w_typed)
{
#define G2N_FIN(T) (weighted ? rows_finalize_typed<T, Ent64>(h, M, n, sym, w_emit, (const T*)w_typed) : rows_finalize_typed<T, Ent32>(h, M, n, sym, WEmit{nullptr, 0u}, nullptr))
    switch (dtype) {
        case G2N_DTYPE_F64: return G2N_FIN(double);
        case G2N_DTYPE_F32: return G2N_FIN(float);
        case G2N_DTYPE_I32: return G2N_FIN(int32_t);
        case G2N_DTYPE_I8: return G2N_FIN(int8_t);
        case G2N_DTYPE_BOOL: return G2N_FIN(BoolT);
    }
#undef G2N_FIN
    h->err = "unknown dtype";
    return G2N_ERR_INVALID;
}

// Bucketing passes (rowsort.cuh: RowRange): one when the random-access working set (entries + row histogram +
// cursors) is near the L2 size, else one per slice of rows that is.  G2N_DBG_ROWPASS forces a count (tests).
struct RowPasses {
    u32 count, width;
    bool bucketed;  // partition the entries by row bucket first (rowsort.cuh: RowBuckets) instead of one pass per row range
    RowRange at(u32 p, u64 n) const
    {
        RowRange r;
        r.lo = p * width;
        r.n = p + 1 == count ? 0xFFFFFFFFu - r.lo : width;  // the last pass takes everything above
        (void)n;
        return r;
    }
};
static const RowRange ROW_RANGE_ALL = {0u, 0xFFFFFFFFu};
static RowPasses row_passes(u64 M, size_t ent_bytes, u64 n)
{
    // Row arrays up to 96 MB: one flat pass (C2).  Beyond: partitioned into buckets of about 32 MB and sub-buckets that
    // fit shared memory (rowsort.cuh) -- C5 shape at 5 % (200 MB): three row-range passes 1.00 ms, partitioned 0.82 ms;
    // C4d (800 MB): 5.84 -> 3.03 ms; C3 (1.1 GB, weighted): 7.0 -> 3.75 ms (profiles/r2b_buckets.md).  G2N_DBG_NOBUCKET:
    // the older scheme, one pass of the flat kernels per 96 MB slice of rows.
    const u64 bytes = M * ent_bytes + n * 8;
    const bool forced = getenv("G2N_DBG_ROWPASS") != nullptr;
    const bool bucketed = (forced || bytes > (96ull << 20)) && !getenv("G2N_DBG_NOBUCKET");
    const u64 budget = bucketed ? 32ull << 20 : 96ull << 20;
    u64 R = (bytes + budget - 1) / budget;
    if (const char* e = getenv("G2N_DBG_ROWPASS")) R = (u64)atoll(e);
    if (R < 1) R = 1;
    if (R > 64) R = 64;
    if (R > n) R = n ? n : 1;
    RowPasses rp;
    rp.bucketed = bucketed && R > 1;
    rp.count = (u32)R;
    rp.width = (u32)((n + R - 1) / R);
    if (rp.width == 0) rp.width = 1;
    return rp;
}

int rows_scan(g2n_handle* h, u64 n_cap, const u32* n_dev);

// Row buckets of the partitioned build: a power-of-two number of rows each, about as many buckets as row_passes() asks
// for passes (never more than RB_MAX).
static RowBuckets row_buckets(const RowPasses& rp, u64 n)
{
    RowBuckets rb;
    rb.shift = 0;
    while (rb.shift < 31 && (2ull << rb.shift) <= rp.width) rb.shift++;  // largest power of two <= width
    while (rb.shift < 31 && ((n ? n - 1 : 0) >> rb.shift) + 1 > RB_MAX) rb.shift++;
    rb.count = (u32)(((n ? n - 1 : 0) >> rb.shift) + 1);
    return rb;
}

// Both levels of the partitioned build (rowsort.cuh).  Sub-buckets: a power-of-two number of rows with about half the
// shared-memory capacity in entries on average; buckets: at most SB_FAN sub-buckets each, at most RB_MAX buckets.
struct BucketPlan {
    RowBuckets rb;
    SubBuckets sb;
    bool two_level;
    u32 smem_entries;  // staged entries the placement kernel's shared memory is sized for
};
static BucketPlan plan_buckets(const RowPasses& rp, u64 M, u64 n, size_t ent_bytes)
{
    BucketPlan P;
    memset(&P, 0, sizeof(P));
    P.rb = row_buckets(rp, n);
    if (getenv("G2N_DBG_NOSUB")) return P;
    const u32 nominal = ent_bytes == 8 ? 12288u : 16384u;  // 96 KB / 64 KB of staged entries + 16 KB of cursors
    const double avg = (n && M > n) ? (double)M / (double)n : 1.0;
    u32 shift2 = 0;
    while (shift2 < 12 && (double)(2u << shift2) * avg <= nominal / 2) shift2++;
    u32 shift1 = P.rb.shift;
    if (shift1 < shift2) shift2 = shift1;  // small inputs (forced in tests): one sub-bucket per bucket
    if (shift1 - shift2 > 8) shift1 = shift2 + 8;
    const u64 last = n ? n - 1 : 0;
    while ((last >> shift1) + 1 > RB_MAX) {
        shift1++;
        if (shift1 - shift2 > 8) shift2++;
    }
    if (shift2 > 12) return P;  // more than SB_ROWS_MAX rows per sub-bucket: one level only
    P.rb.shift = shift1;
    P.rb.count = (u32)((last >> shift1) + 1);
    P.sb.shift2 = shift2;
    P.sb.fan_shift = shift1 - shift2;
    P.sb.n_sub = P.rb.count << P.sb.fan_shift;
    P.sb.cap = nominal;
    if (const char* e = getenv("G2N_DBG_SUBCAP")) { const u32 v = (u32)atoll(e); if (v < nominal) P.sb.cap = v; }  // tests: force the fallback
    P.smem_entries = nominal;
    P.two_level = true;
    return P;
}

// From the bucket-major pair list (pair_major / pair_ent, counts in bucket_ctl) to rowptr + entries grouped by row.
// `M` and `n` are host-side capacities; `counted`: d_rowcnt already holds the row histogram.
template <class ENT>
static int bucketed_tail(g2n_handle* h, const BucketPlan& P, u64 M, u64 n, bool counted, int sym)
{
    h->rows_presorted = false;
    typedef typename ENT::type EV;
    BucketCtl* ctl = h->bucket_ctl.as<BucketCtl>();
    const u32 pgrid = grid_for((M + 3) / 4, 256);
    if (!P.two_level) {
        if (!counted) { KScope ks(h, "k_bucket_rows_count"); k_bucket_rows_count<<<pgrid, 256, 0, h->stream>>>(h->pair_major.as<u32>(), h->d_ds, ctl, P.rb.count, h->d_rowcnt); }
        CK(cudaGetLastError());
        int rc = rows_scan(h, n, &h->d_ds->rows);
        if (rc) return rc;
        CK(cudaEventRecord(h->ev[EV_EMIT], h->stream));
        { KScope ks(h, "k_bucket_rows_scatter"); k_bucket_rows_scatter<ENT><<<pgrid, 256, 0, h->stream>>>(h->pair_major.as<u32>(), h->pair_ent.as<EV>(), h->d_ds, ctl, P.rb.count, h->cursor.as<u32>(), h->entries.as<EV>()); }
        CK(cudaGetLastError());
        return G2N_OK;
    }
    const SubBuckets sb = P.sb;
    CK(h->pair2_major.ensure((M + 1) * sizeof(u32)));
    CK(h->pair2_ent.ensure((M + 1) * sizeof(EV)));
    CK(h->sub_ctl.ensure(2 * (size_t)sb.n_sub * sizeof(u32)));
    CK(h->sub_off.ensure(((size_t)sb.n_sub + 2) * sizeof(u32)));
    CK(cudaMemsetAsync(h->sub_ctl.p, 0, 2 * (size_t)sb.n_sub * sizeof(u32), h->stream));
    u32* sub_cnt = h->sub_ctl.as<u32>();
    u32* sub_cur = sub_cnt + sb.n_sub;
    { KScope ks(h, "k_sub_count"); k_sub_count<<<grid_for((M + SB_CHUNK - 1) / SB_CHUNK, 1, 8), 256, 0, h->stream>>>(h->pair_major.as<u32>(), h->d_ds, ctl, P.rb, sb, sub_cnt); }
    CK(cudaGetLastError());
    LoadArray<u32> lsc{sub_cnt};
    int rc = launch_scan<u32>(h, lsc, h->sub_off.as<u32>(), nullptr, sb.n_sub, nullptr, nullptr);
    if (rc) return rc;
    { KScope ks(h, "k_sub_scatter"); k_sub_scatter<ENT><<<grid_for((M + RB_ROUND - 1) / RB_ROUND, 1, 8), 256, 0, h->stream>>>(h->pair_major.as<u32>(), h->pair_ent.as<EV>(), h->d_ds, ctl, P.rb, sb, h->sub_off.as<u32>(), sub_cur, h->pair2_major.as<u32>(), h->pair2_ent.as<EV>()); }
    const u32 cgrid = grid_for(sb.n_sub, 1, 8);
    if (!counted) { KScope ks(h, "k_sub_rows_count"); k_sub_rows_count<<<cgrid, 256, 0, h->stream>>>(h->pair2_major.as<u32>(), h->d_ds, sb, h->sub_off.as<u32>(), h->d_rowcnt); }
    CK(cudaGetLastError());
    rc = rows_scan(h, n, &h->d_ds->rows);
    if (rc) return rc;
    CK(cudaEventRecord(h->ev[EV_EMIT], h->stream));
    const size_t smem = SB_ROWS_MAX * sizeof(u32) + (size_t)P.smem_entries * sizeof(EV);
    const u32 bit = sizeof(EV) == 8 ? 2u : 1u;
    if (!(h->place_attr & bit)) {
        CK(cudaFuncSetAttribute(k_sub_rows_place<ENT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        h->place_attr |= bit;
    }
    // the short rows are sorted while their sub-bucket sits in shared memory (not weighted max(S, S^T): k_rows_sort counts
    // with the summed weights there)
    const bool fuse_sort = !(ENT::kWeighted && sym) && !getenv("G2N_DBG_NOFUSESORT");
    if (fuse_sort) CK(h->ucnt.ensure((n + 2) * sizeof(u32)));
    { KScope ks(h, "k_sub_rows_place"); k_sub_rows_place<ENT><<<grid_for(sb.n_sub, 1, 2), SB_PT, smem, h->stream>>>(h->pair2_major.as<u32>(), h->pair2_ent.as<EV>(), h->d_ds, sb, h->sub_off.as<u32>(), h->rowptr.as<u32>(), h->cursor.as<u32>(), h->entries.as<EV>(), fuse_sort ? h->ucnt.as<u32>() : nullptr); }
    h->rows_presorted = fuse_sort;
    CK(cudaGetLastError());
    return G2N_OK;
}

// bucketed passes shared by the unweighted and the weighted build: partition the entries of the stored edge records
// (node IDs in place unless `translate`) by row bucket, then bucketed_tail()
template <class ENT>
static int bucketed_rows(g2n_handle* h, const BucketPlan& P, u64 M, u64 n, int sym, int csc, bool translate, bool counted, const u32* emit_t0)
{
    typedef typename ENT::type EV;
    const RowBuckets rb = P.rb;
    CK(h->pair_major.ensure((M + 1) * sizeof(u32)));
    CK(h->pair_ent.ensure((M + 1) * sizeof(EV)));
    CK(h->bucket_ctl.ensure(sizeof(BucketCtl)));
    CK(cudaMemsetAsync(h->bucket_ctl.p, 0, sizeof(BucketCtl), h->stream));
    BucketCtl* ctl = h->bucket_ctl.as<BucketCtl>();
    u32* es = h->edge_slots.as<u32>();
    const u32 fgrid = grid_for((h->cap_E + EF_BATCH - 1) / EF_BATCH, 256);
    const u32 sgrid = grid_for((h->cap_E + 511) / 512, 1, 8);  // a CTA round takes 512 - 1024 records (RbRecs)
    const u32* sid = translate ? h->slot_id.as<u32>() : nullptr;
    {
        KScope ks(h, "k_bucket_count");
        switch (h->tpe) {
            case 1: k_bucket_count<1><<<fgrid, 256, 0, h->stream>>>(es, sid, h->d_ds, sym, csc, rb, ctl); break;
            case 2: k_bucket_count<2><<<fgrid, 256, 0, h->stream>>>(es, sid, h->d_ds, sym, csc, rb, ctl); break;
            default: k_bucket_count<4><<<fgrid, 256, 0, h->stream>>>(es, sid, h->d_ds, sym, csc, rb, ctl); break;
        }
    }
    h->edges_are_ids = true;
    {
        KScope ks(h, "k_bucket_scatter");
        switch (h->tpe) {
            case 1: k_bucket_scatter<1, ENT><<<sgrid, 256, 0, h->stream>>>(es, emit_t0, h->d_ds, sym, csc, rb, ctl, h->pair_major.as<u32>(), h->pair_ent.as<EV>()); break;
            case 2: k_bucket_scatter<2, ENT><<<sgrid, 256, 0, h->stream>>>(es, emit_t0, h->d_ds, sym, csc, rb, ctl, h->pair_major.as<u32>(), h->pair_ent.as<EV>()); break;
            default: k_bucket_scatter<4, ENT><<<sgrid, 256, 0, h->stream>>>(es, emit_t0, h->d_ds, sym, csc, rb, ctl, h->pair_major.as<u32>(), h->pair_ent.as<EV>()); break;
        }
    }
    CK(cudaGetLastError());
    return bucketed_tail<ENT>(h, P, M, n, counted, sym);
}

// rowcnt -> rowptr + cursors; rows longer than RS_SMALL are listed on the way
int rows_scan(g2n_handle* h, u64 n_cap, const u32* n_dev)
{
    CK(h->rowptr.ensure((n_cap + 2) * sizeof(u32)));
    CK(h->cursor.ensure((n_cap + 2) * sizeof(u32)));
    CK(h->biglist.ensure((n_cap + 2) * sizeof(u32)));
    LoadRowCounts ldc{h->d_rowcnt, h->biglist.as<u32>(), h->d_bigcount};
    return launch_scan<u32>(h, ldc, h->rowptr.as<u32>(), h->cursor.as<u32>(), n_cap, n_dev, h->d_scan_rows[0]);
}

struct LoadTileCounts {
    const TileInfo* p;
    __device__ __forceinline__ u64 operator()(u64 i) const { return ((u64)p[i].n_rec << 32) | (u64)p[i].n_edge; }
    __device__ __forceinline__ u64 peek(u64 i) const { return (*this)(i); }
};

EmitParams emit_params(g2n_handle* h)
{
    EmitParams E;
    E.edge_slots = h->edge_slots.as<u32>();
    E.edge_w = h->params.weight_tag_len > 0 ? h->edge_w.as<double>() : nullptr;
    E.slot_id = h->slot_id.as<u32>();
    E.tile_info = h->tile_info.as<TileInfo>();
    E.tile_base = h->tile_base.as<u64>();
    E.n_tiles = h->n_tiles;
    E.slots_per_edge = h->spe;
    E.tpe = h->tpe;
    E.ids_ready = h->edges_are_ids ? 1 : 0;
    E.write_ids = 0;
    E.ds = h->d_ds;
    return E;
}

// K3 + K4 for a compressed result of the current build.  fmt: G2N_FMT_CSR | G2N_FMT_CSC.
// Sizes on the host are this build's capacities (cap_n nodes, cap_E edge records): exact after a host
// round trip, hints in a speculative build; the kernels read the actual ones from DevSizes.
// `zeroed`: the zrows arena was already laid out and cleared for this build.
int build_compressed(g2n_handle* h, int fmt, bool zeroed, bool counted)
{
    const u64 n = h->cap_n;
    const u64 T = h->cap_E * (u64)h->tpe;
    const int sym = h->symmax ? 1 : 0;
    const u64 M = sym ? 2 * T : T;
    h->result_format = fmt;
    h->rows_presorted = false;
    if (M >= 0xFFFFFFF0ull) { h->err = "more than 2^32 triplets in one build"; return G2N_ERR_UNSUPPORTED; }
    const bool weighted = h->params.weight_tag_len > 0;
    // for a symmetric result CSC arrays equal CSR arrays; bucket by row either way
    const int csc = (!sym && fmt == G2N_FMT_CSC) ? 1 : 0;
    if (!zeroed) {
        int rc = layout_zrows(h, n);
        if (rc) return rc;
        CK(cudaMemsetAsync(h->zrows.p, 0, h->zrows_bytes, h->stream));
    }
    CK(h->entries.ensure((M + 1) * (weighted ? sizeof(u64) : sizeof(u32))));
    if (weighted) CK(h->w_emit.ensure((h->cap_E + 1) * sizeof(double)));
    EmitParams E = emit_params(h);
    int rc;
    if (!weighted) {
        // flat passes over the stored edge records (rowsort.cuh): the emission index is not needed
        const u32 fgrid = grid_for((h->cap_E + EF_BATCH - 1) / EF_BATCH, 256);
        u32* es = h->edge_slots.as<u32>();
        // the histogram (4 bytes per row) stays L2-resident by itself: one pass; the scatter below, whose entries do
        // not, runs once per row range
        const RowPasses rp = row_passes(M, sizeof(u32), n);
        const bool bucketed = rp.bucketed;
        if (bucketed) {
            // row arrays far larger than L2: partition the entries by row bucket first (rowsort.cuh: RowBuckets)
            rc = bucketed_rows<Ent32>(h, plan_buckets(rp, M, n, sizeof(u32)), M, n, sym, csc, !h->edges_are_ids, counted, nullptr);
            if (rc) return rc;
        } else {
        if (!counted) {  // (the tokenizer did not count the rows: table far larger than L2, or a later convert)
            const RowRange rr = ROW_RANGE_ALL;
            // IDs are in place already after an earlier convert of the same build
            const int translate = h->edges_are_ids ? 0 : 1;
            KScope ks(h, "k_edges_count_flat");
            switch (h->tpe) {
                case 1: k_edges_count_flat<1><<<fgrid, 256, 0, h->stream>>>(es, h->slot_id.as<u32>(), h->d_ds, sym, csc, h->d_rowcnt, rr, translate); break;
                case 2: k_edges_count_flat<2><<<fgrid, 256, 0, h->stream>>>(es, h->slot_id.as<u32>(), h->d_ds, sym, csc, h->d_rowcnt, rr, translate); break;
                default: k_edges_count_flat<4><<<fgrid, 256, 0, h->stream>>>(es, h->slot_id.as<u32>(), h->d_ds, sym, csc, h->d_rowcnt, rr, translate); break;
            }
            CK(cudaGetLastError());
            h->edges_are_ids = true;
        }
        rc = rows_scan(h, n, &h->d_ds->rows);
        if (rc) return rc;
        CK(cudaEventRecord(h->ev[EV_EMIT], h->stream));
        for (u32 ps = 0; ps < rp.count; ps++) {
            const RowRange rr = rp.at(ps, n);
            const u32* sid = h->edges_are_ids ? nullptr : h->slot_id.as<u32>();  // first pass after a counting tokenizer: slots -> IDs here
            KScope ks(h, "k_edges_scatter_flat");
            switch (h->tpe) {
                case 1: k_edges_scatter_flat<1><<<fgrid, 256, 0, h->stream>>>(es, sid, h->d_ds, sym, csc, h->cursor.as<u32>(), h->entries.as<u32>(), rr); break;
                case 2: k_edges_scatter_flat<2><<<fgrid, 256, 0, h->stream>>>(es, sid, h->d_ds, sym, csc, h->cursor.as<u32>(), h->entries.as<u32>(), rr); break;
                default: k_edges_scatter_flat<4><<<fgrid, 256, 0, h->stream>>>(es, sid, h->d_ds, sym, csc, h->cursor.as<u32>(), h->entries.as<u32>(), rr); break;
            }
            h->edges_are_ids = true;
        }
        CK(cudaGetLastError());
        }
    } else {
        const u32 egrid = grid_for((u64)h->n_tiles * 32, 256);
        const RowPasses rp = row_passes(M, sizeof(u64), n);
        const bool bucketed = rp.bucketed;
        CK(h->emit_t0.ensure((h->cap_E + 1) * sizeof(u32)));
        {
            // leaves node IDs in edge_slots for the scatter passes (and later converts) and lays out the weights
            E.ids_ready = h->edges_are_ids ? 1 : 0;
            E.write_ids = E.ids_ready ? 0 : 1;
            KScope ks(h, "k_rows_count");
            k_rows_count<<<egrid, 256, 0, h->stream>>>(E, sym, csc, (counted || bucketed) ? nullptr : h->d_rowcnt, h->w_emit.as<double>(), ROW_RANGE_ALL, h->emit_t0.as<u32>());
        }
        CK(cudaGetLastError());
        h->edges_are_ids = true;
        E.ids_ready = 1;
        E.write_ids = 0;
        if (bucketed) {
            rc = bucketed_rows<Ent64>(h, plan_buckets(rp, M, n, sizeof(u64)), M, n, sym, csc, false, counted, h->emit_t0.as<u32>());
            if (rc) return rc;
        } else {
        rc = rows_scan(h, n, &h->d_ds->rows);
        if (rc) return rc;
        CK(cudaEventRecord(h->ev[EV_EMIT], h->stream));
        const u32 fgrid = grid_for((h->cap_E + EF_BATCH - 1) / EF_BATCH, 256);
        const u32* es = h->edge_slots.as<u32>();
        for (u32 ps = 0; ps < rp.count; ps++) {
            const RowRange rr = rp.at(ps, n);
            KScope ks(h, "k_rows_scatter_flat");
            switch (h->tpe) {
                case 1: k_rows_scatter_flat<1><<<fgrid, 256, 0, h->stream>>>(es, h->emit_t0.as<u32>(), h->d_ds, sym, csc, h->cursor.as<u32>(), h->entries.as<u64>(), rr); break;
                case 2: k_rows_scatter_flat<2><<<fgrid, 256, 0, h->stream>>>(es, h->emit_t0.as<u32>(), h->d_ds, sym, csc, h->cursor.as<u32>(), h->entries.as<u64>(), rr); break;
                default: k_rows_scatter_flat<4><<<fgrid, 256, 0, h->stream>>>(es, h->emit_t0.as<u32>(), h->d_ds, sym, csc, h->cursor.as<u32>(), h->entries.as<u64>(), rr); break;
            }
        }
        CK(cudaGetLastError());
        }
    }
    CK(cudaEventRecord(h->ev[EV_SORT], h->stream));
    // one weight per edge record: the record's tpe triplets find it at (emission index) >> log2(tpe)
    rc = rows_finalize(h, h->params.dtype, weighted, M, n, sym, WEmit{weighted ? h->w_emit.as<double>() : nullptr, h->tpe == 4 ? 2u : (h->tpe == 2 ? 1u : 0u)}, nullptr);
    if (rc) return rc;
    CK(cudaEventRecord(h->ev[EV_REDUCE], h->stream));
    return G2N_OK;
}

template <typename T>
int emit_coo_typed(g2n_handle* h, u64 T_)
{
    CK(h->data.ensure((T_ + 1) * sizeof(T)));
    { KScope ks(h, "k_emit_coo"); k_emit_coo<T><<<grid_for((u64)h->n_tiles * 32, 256), 256, 0, h->stream>>>(emit_params(h), h->row.as<int32_t>(), h->col.as<int32_t>(), h->data.as<T>()); }
    CK(cudaGetLastError());
    return G2N_OK;
}

int build_coo(g2n_handle* h)
{
    const u64 T = h->cap_E * (u64)h->tpe;
    h->result_format = G2N_FMT_COO;
    CK(h->row.ensure((T + 1) * sizeof(int32_t)));
    CK(h->col.ensure((T + 1) * sizeof(int32_t)));
    int rc = G2N_OK;
    switch (h->params.dtype) {
        case G2N_DTYPE_F64: rc = emit_coo_typed<double>(h, T); break;
        case G2N_DTYPE_F32: rc = emit_coo_typed<float>(h, T); break;
        case G2N_DTYPE_I32: rc = emit_coo_typed<int32_t>(h, T); break;
        case G2N_DTYPE_I8: rc = emit_coo_typed<int8_t>(h, T); break;
        case G2N_DTYPE_BOOL: rc = emit_coo_typed<BoolT>(h, T); break;
        default: h->err = "unknown dtype"; return G2N_ERR_INVALID;
    }
    CK(cudaEventRecord(h->ev[EV_EMIT], h->stream));
    CK(cudaEventRecord(h->ev[EV_SORT], h->stream));
    CK(cudaEventRecord(h->ev[EV_REDUCE], h->stream));
    return rc;
}

// first error / first unknown record in file order, counts
void collect_diag(g2n_handle* h, const Counters& hc, u64 n_records)
{
    h->diag.n_records = n_records;
    h->diag.n_edge_records = hc.edge_alloc;
    h->diag.n_long_keys = hc.n_long;
    if (hc.flags & CF_CAST_OVERFLOW) h->diag.warn_flags |= G2N_WARN_CAST_OVERFLOW;
    const u64 first_error = ~hc.first_error_inv, first_unknown = ~hc.first_unknown_inv;  // ~0 if none
    if (first_error != ~0ull) {
        h->diag.err_kind = (int32_t)(first_error & 0xFF);
        h->diag.err_offset = first_error >> 8;
    }
    if (first_unknown != ~0ull && (first_error == ~0ull || (first_unknown >> 8) < (first_error >> 8))) {
        h->diag.unknown_byte = (int32_t)(first_unknown & 0xFF);
        h->diag.unknown_offset = first_unknown >> 8;
    }
}

#define G2N_H2D_PIECE ((u64)8 << 20)  // bytes per host -> device copy piece
#define G2N_READERS 4                  // file reader threads (g2n_build_file)
#define G2N_SPEC_MISS 1000  // internal: the speculative build has to be repeated with a host round trip

// The one host round trip of a build: counters + device-side sizes come back, timings are read.
int finish_result(g2n_handle* h)
{
    CK(cudaMemcpyAsync(h->h_ctl, h->d_ctl, sizeof(Ctl), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    const DevSizes& ds = h->h_ctl->s;
    if (!ds.ok) {
        if (h->spec) return G2N_SPEC_MISS;
        h->err = "device-side size check failed after a host-checked tokenizer pass";
        return G2N_ERR_INTERNAL;
    }
    if (h->spec) {
        const Counters& hc = *h->h_cnt;
        collect_diag(h, hc, ds.R);
        h->n_nodes = ds.n;
        h->n_edges = ds.E;
        h->n_records = ds.R;
        h->n_long = hc.n_long;
    }
    h->nnz = ds.nnz;
    float ms = 0;
    cudaEventElapsedTime(&ms, h->ev[EV_START], h->ev[EV_REDUCE]);
    h->diag.ms_total = ms;
    cudaEventElapsedTime(&h->diag.ms_h2d, h->ev[EV_START], h->ev[EV_H2D]);
    cudaEventElapsedTime(&h->diag.ms_stage[0], h->ev[EV_H2D], h->ev[EV_TOKENIZE]);
    cudaEventElapsedTime(&h->diag.ms_stage[1], h->ev[EV_TOKENIZE], h->ev[EV_IDS]);
    cudaEventElapsedTime(&h->diag.ms_stage[3], h->ev[EV_IDS], h->ev[EV_EMIT]);
    cudaEventElapsedTime(&h->diag.ms_stage[4], h->ev[EV_EMIT], h->ev[EV_SORT]);
    cudaEventElapsedTime(&h->diag.ms_stage[5], h->ev[EV_SORT], h->ev[EV_REDUCE]);
    h->diag.gpu_launches = h->launches;
    h->diag.n_triplets = h->n_edges * (u64)h->tpe;
    h->diag.speculative = h->spec ? 1 : 0;
    return G2N_OK;
}

}  // namespace

// =====================================================================================
extern "C" {

int g2n_abi_version(void) { return G2N_ABI_VERSION; }

int g2n_plan_row_buckets(uint64_t entries, uint64_t rows, int entry_bytes, uint32_t out[8])
{
    if (!out || (entry_bytes != 4 && entry_bytes != 8)) return G2N_ERR_INVALID;
    const RowPasses rp = row_passes(entries, (size_t)entry_bytes, rows);
    memset(out, 0, 8 * sizeof(uint32_t));
    out[0] = rp.count;
    out[1] = rp.bucketed ? 1u : 0u;
    if (rp.bucketed) {
        const BucketPlan P = plan_buckets(rp, entries, rows, (size_t)entry_bytes);
        out[2] = P.rb.count; out[3] = P.rb.shift;
        out[4] = P.two_level ? 1u : 0u;
        if (P.two_level) { out[5] = P.sb.n_sub; out[6] = P.sb.shift2; out[7] = P.sb.cap; }
    }
    return G2N_OK;
}
int g2n_dist_close_peers(g2n_handle* h);

int g2n_create(int device, g2n_handle** out)
{
    if (!out) return G2N_ERR_INVALID;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) return G2N_ERR_CUDA;
    if (cudaSetDevice(device) != cudaSuccess) return G2N_ERR_CUDA;
    g2n_handle* h = new g2n_handle();
    h->device = device;
    if (cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking) != cudaSuccess) { delete h; return G2N_ERR_CUDA; }
    if (cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking) != cudaSuccess) { delete h; return G2N_ERR_CUDA; }
    h->stream = h->own_stream;
    for (int i = 0; i < EV_COUNT; i++) cudaEventCreate(&h->ev[i]);
    if (cudaHostAlloc((void**)&h->h_ctl, sizeof(Ctl), cudaHostAllocDefault) != cudaSuccess ||
        cudaHostAlloc((void**)&h->h_tail, 8 * sizeof(u64), cudaHostAllocDefault) != cudaSuccess) {
        delete h;
        return G2N_ERR_CUDA;
    }
    h->h_cnt = &h->h_ctl->c;
    memset(&h->diag, 0, sizeof(h->diag));
    h->diag.unknown_byte = -1;
    *out = h;
    return G2N_OK;
}

void g2n_destroy(g2n_handle* h)
{
    if (!h) return;
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    DevBuf* bufs[] = {&h->text, &h->zearly, &h->zids, &h->zrows, &h->defer, &h->edge_slots, &h->edge_w, &h->longs, &h->tile_info, &h->tile_base, &h->wprefix,
                      &h->slot_id, &h->id2slot, &h->name_off, &h->names, &h->rowptr, &h->cursor, &h->entries, &h->w_emit, &h->biglist, &h->ucnt, &h->indptr, &h->indices, &h->data, &h->row, &h->col,
                      &h->scan_state, &h->up_row, &h->up_col, &h->up_data, &h->tsv, &h->tsv_off, &h->el_len, &h->el_off, &h->el_text, &h->emit_t0, &h->bfs_levels, &h->bfs_q0, &h->bfs_q1, &h->bfs_ctl, &h->bfs_nodes, &h->bfs_out, &h->path_starts, &h->path_recs, &h->path_cnt, &h->path_off, &h->path_ids, &h->path_misc, &h->dx_arena, &h->dx_ctl, &h->dx_loc, &h->dx_zg, &h->dx_gslot, &h->dx_gpos, &h->dx_sent, &h->dx_tcnt, &h->dx_toff, &h->pair_major, &h->pair_ent, &h->bucket_ctl, &h->pair2_major, &h->pair2_ent, &h->sub_ctl, &h->sub_off};
    if (h->dx_inited) g2n_dist_close_peers(h);
    for (DevBuf* b : bufs) b->release();
    if (h->h_loc) cudaFreeHost(h->h_loc);
    for (int i = 0; i < EV_COUNT; i++) cudaEventDestroy(h->ev[i]);
    for (KTimer& t : h->ktimers) { cudaEventDestroy(t.a); cudaEventDestroy(t.b); }
    if (h->h_ctl) cudaFreeHost(h->h_ctl);
    if (h->h_tail) cudaFreeHost(h->h_tail);
    for (cudaEvent_t e : h->copy_ev) cudaEventDestroy(e);
    for (cudaEvent_t e : h->stage_ev) cudaEventDestroy(e);
    for (void* b : h->stage) cudaFreeHost(b);
    for (int k = 0; k < 2; k++) { if (h->gz_stage[k]) cudaFreeHost(h->gz_stage[k]); if (h->gz_ev[k]) cudaEventDestroy(h->gz_ev[k]); }
    if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
    if (h->own_stream) cudaStreamDestroy(h->own_stream);
    delete h;
}

int g2n_set_stream(g2n_handle* h, void* cuda_stream)
{
    if (!h) return G2N_ERR_INVALID;
    // (void*)-1: the handle's own non-blocking stream; anything else, including NULL (the CUDA default
    // stream), is used as given so that work is ordered with the caller's stream
    h->stream = cuda_stream == (void*)-1 ? h->own_stream : (cudaStream_t)cuda_stream;
    return G2N_OK;
}

void* g2n_host_alloc(uint64_t nbytes)
{
    void* p = nullptr;
    if (cudaHostAlloc(&p, nbytes ? nbytes : 1, cudaHostAllocDefault) != cudaSuccess) return nullptr;
    return p;
}

void g2n_host_free(void* p)
{
    if (p) cudaFreeHost(p);
}

int g2n_set_speculation(g2n_handle* h, int on)
{
    if (!h) return G2N_ERR_INVALID;
    h->speculate = on != 0;
    return G2N_OK;
}

int g2n_set_profile(g2n_handle* h, int on)
{
    if (!h) return G2N_ERR_INVALID;
    h->profile = on != 0;
    return G2N_OK;
}

int g2n_kernel_times(g2n_handle* h, g2n_ktime* out, int cap)
{
    if (!h || (!out && cap > 0)) return -1;
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    int n = 0;
    for (size_t i = 0; i < h->kt_used; i++) {
        float ms = 0;
        if (cudaEventElapsedTime(&ms, h->ktimers[i].a, h->ktimers[i].b) != cudaSuccess) continue;
        int j = 0;
        for (; j < n; j++)
            if (strcmp(out[j].name, h->ktimers[i].name) == 0) break;
        if (j == n) {
            if (n >= cap) continue;
            memset(&out[n], 0, sizeof(out[n]));
            strncpy(out[n].name, h->ktimers[i].name, sizeof(out[n].name) - 1);
            n++;
        }
        out[j].ms += ms;
        out[j].launches++;
    }
    return n;
}

const char* g2n_last_error(g2n_handle* h) { return h ? h->err.c_str() : "null handle"; }

int g2n_status(g2n_handle* h, g2n_diag* out)
{
    if (!h || !out) return G2N_ERR_INVALID;
    h->diag.gpu_launches = h->launches;
    h->diag.n_triplets = h->n_edges * (u64)h->tpe;
    *out = h->diag;
    return G2N_OK;
}

// zearly arena: table slots | per-slot row counters (optional) | counters + DevSizes | look-back state of the tile scan
static int layout_zearly(g2n_handle* h, u32 cap, u32 n_tiles, bool slot_counters, size_t* bytes)
{
    auto up = [](size_t x) { return (x + 255) & ~(size_t)255; };
    const size_t a = up((size_t)cap * sizeof(Slot)), c = slot_counters ? up((size_t)cap * sizeof(u32)) : 0;
    const size_t d = up(sizeof(Ctl)), e = up(scan_state_bytes(n_tiles));
    CK(h->zearly.ensure(a + c + d + e));
    uint8_t* base = h->zearly.as<uint8_t>();
    h->d_slots = (Slot*)base;
    h->d_slot_cnt = slot_counters ? (u32*)(base + a) : nullptr;
    h->d_ctl = (Ctl*)(base + a + c);
    h->d_cnt = &h->d_ctl->c;
    h->d_ds = &h->d_ctl->s;
    h->d_scan_tiles = (u64*)(base + a + c + d);
    *bytes = a + c + d + e;
    return G2N_OK;
}

// zids arena: first-appearance bitmap | look-back state of its popcount scan
static int layout_zids(g2n_handle* h, u64 R_cap)
{
    const u64 words = ((4 * R_cap + 31) / 32 + 1 + BM_GROUP) / BM_GROUP * BM_GROUP;  // whole groups of BM_GROUP words
    const size_t a = (words * sizeof(u32) + 255) & ~(size_t)255;
    const size_t b = (scan_state_bytes(words / BM_GROUP + 1) + 255) & ~(size_t)255;
    CK(h->zids.ensure(a + b));
    h->d_bitmap = h->zids.as<u32>();
    h->d_scan_words = (u64*)(h->zids.as<uint8_t>() + a);
    h->zids_bytes = a + b;
    return G2N_OK;
}

// Everything downstream of the tokenizer that can be sized from (cap_n, cap_R): allocate and clear.
static int prepare_late(g2n_handle* h, bool want_rows)
{
    int rc = layout_zids(h, h->cap_R);
    if (rc) return rc;
    CK(cudaMemsetAsync(h->zids.p, 0, h->zids_bytes, h->stream));
    if (want_rows) {
        rc = layout_zrows(h, h->cap_n);
        if (rc) return rc;
        CK(cudaMemsetAsync(h->zrows.p, 0, h->zrows_bytes, h->stream));
    }
    return G2N_OK;
}

static u64 shape_signature(const g2n_params* p, u64 nbytes, int wt_len)
{
    u64 s = nbytes * 0x9e3779b97f4a7c15ull;
    s ^= (u64)(p->directed != 0) | (u64)(p->bidirected != 0) << 1 | (u64)(p->keep_directed_bidir != 0) << 2 | (u64)(p->asymmetric != 0) << 3 |
         (u64)(p->strip_orientation != 0) << 4 | (u64)(wt_len > 0) << 5 | (u64)(p->want_format & 7) << 6;
    return s | 1;
}

// Phase 1 of every build: text -> hash table (min order per key), tile_info / tile_base, edge_slots,
// DevSizes.  spec: no host round trip -- buffers come from the hints of the previous build of this shape.
static int tokenize_phase(g2n_handle* h, const uint8_t* text, uint64_t nbytes, const g2n_params* p, bool spec, bool late_rows)
{
    if (!h || !p || (!text && nbytes)) return G2N_ERR_INVALID;
    h->err.clear();
    h->built = false;
    h->have_edges = false;
    h->names_ready = false;
    h->tsv_ready = false;
    h->el_ready = false;
    h->paths_ready = false;
    h->edges_are_ids = false;
    h->spec = spec;
    if (p->dtype < G2N_DTYPE_F64 || p->dtype > G2N_DTYPE_BOOL) { h->err = "unknown dtype"; return G2N_ERR_INVALID; }
    if (p->want_format < G2N_FMT_NATIVE || p->want_format > G2N_FMT_CSC) { h->err = "unknown want_format"; return G2N_ERR_INVALID; }
    if (p->weight_tag_len > 64) { h->err = "weight tag longer than 64 bytes"; return G2N_ERR_UNSUPPORTED; }
    if (nbytes >= (1ull << 46)) { h->err = "input larger than 2^46 bytes"; return G2N_ERR_UNSUPPORTED; }
    h->params = *p;
    if (p->weight_tag && p->weight_tag_len > 0) memcpy(h->weight_tag, p->weight_tag, p->weight_tag_len);
    else h->params.weight_tag_len = 0;
    for (int i = 0; i < h->params.weight_tag_len; i++)
        if (h->weight_tag[i] == ':') { h->params.weight_tag_len = 0; break; }  // tag names never hold ':' (parser.py:184)
    h->params.weight_tag = h->weight_tag;
    CK(cudaSetDevice(h->device));
    h->launches = 0;
    h->kt_used = 0;
    memset(&h->diag, 0, sizeof(h->diag));
    h->diag.unknown_byte = -1;

    CK(cudaEventRecord(h->ev[EV_START], h->stream));
    if (p->text_on_device && ((uintptr_t)text & 15) == 0) {
        h->d_text = text;
    } else if (p->text_on_device) {
        // the tokenizer reads 16-byte pieces (TMA windows): an unaligned device range is copied once
        CK(h->text.ensure(nbytes + 64));
        if (nbytes) CK(cudaMemcpyAsync(h->text.p, text, nbytes, cudaMemcpyDeviceToDevice, h->stream));
        h->d_text = h->text.as<uint8_t>();
    } else if (h->src_fd >= 0 && h->src_resident) {
        h->d_text = h->text.as<uint8_t>();
        h->n_pieces = 0;
    } else if (h->src_fd >= 0) {
        // file source: G2N_READERS threads pread() 8 MiB pieces into pinned staging buffers (two per thread) and queue
        // their copies on the copy stream; the tokenizer is launched piece by piece behind them (as for a host text)
        CK(h->text.ensure(nbytes + 64));
        h->d_text = h->text.as<uint8_t>();
        h->n_pieces = 0;
        if (nbytes) {
            const u64 piece = G2N_H2D_PIECE;
            const u32 np = (u32)((nbytes + piece - 1) / piece);
            const u32 R = np < G2N_READERS ? np : G2N_READERS;
            while (h->copy_ev.size() < np + 1) {
                cudaEvent_t e;
                CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
                h->copy_ev.push_back(e);
            }
            while (h->stage.size() < 2 * G2N_READERS) {
                void* b = nullptr;
                CK(cudaHostAlloc(&b, piece, cudaHostAllocDefault));
                h->stage.push_back(b);
                cudaEvent_t e;
                CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
                h->stage_ev.push_back(e);
            }
            h->piece_issued.reset(new std::atomic<int>[np]);
            for (u32 k = 0; k < np; k++) h->piece_issued[k].store(0);
            h->reader_err.store(0);
            CK(cudaEventRecord(h->copy_ev[np], h->stream));
            CK(cudaStreamWaitEvent(h->copy_stream, h->copy_ev[np], 0));
            uint8_t* dtext = h->text.as<uint8_t>();
            for (u32 r = 0; r < R; r++) {
                h->readers.emplace_back([h, r, R, np, piece, nbytes, dtext]() {
                    cudaSetDevice(h->device);
                    for (u32 k = r, turn = 0; k < np; k += R, turn++) {
                        const u32 sb = 2 * r + (turn & 1);
                        cudaEventSynchronize(h->stage_ev[sb]);  // the previous copy out of this staging buffer is done
                        const u64 a = (u64)k * piece, len = (a + piece < nbytes ? a + piece : nbytes) - a;
                        u64 got = 0;
                        while (got < len) {
                            const ssize_t n = pread(h->src_fd, (uint8_t*)h->stage[sb] + got, len - got, (off_t)(a + got));
                            if (n <= 0) { h->reader_err.store(1); break; }
                            got += (u64)n;
                        }
                        if (got == len) {
                            if (cudaMemcpyAsync(dtext + a, h->stage[sb], len, cudaMemcpyHostToDevice, h->copy_stream) != cudaSuccess ||
                                cudaEventRecord(h->copy_ev[k], h->copy_stream) != cudaSuccess ||
                                cudaEventRecord(h->stage_ev[sb], h->copy_stream) != cudaSuccess)
                                h->reader_err.store(2);
                        }
                        h->piece_issued[k].store(1, std::memory_order_release);
                    }
                });
            }
            h->n_pieces = np;
            h->piece_bytes = piece;
        }
    } else {
        // host text: copied in pieces on a second stream; the tokenizer is launched piece by piece behind it
        CK(h->text.ensure(nbytes + 64));
        h->d_text = h->text.as<uint8_t>();
        h->n_pieces = 0;
        if (nbytes) {
            const u64 piece = nbytes <= G2N_H2D_PIECE * 2 ? nbytes : G2N_H2D_PIECE;
            const u32 np = (u32)((nbytes + piece - 1) / piece);
            while (h->copy_ev.size() < np + 1) {
                cudaEvent_t e;
                CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
                h->copy_ev.push_back(e);
            }
            CK(cudaEventRecord(h->copy_ev[np], h->stream));  // the copies start after whatever the caller queued before
            CK(cudaStreamWaitEvent(h->copy_stream, h->copy_ev[np], 0));
            for (u32 k = 0; k < np; k++) {
                const u64 a = (u64)k * piece, b = a + piece < nbytes ? a + piece : nbytes;
                CK(cudaMemcpyAsync(h->text.as<uint8_t>() + a, text + a, b - a, cudaMemcpyHostToDevice, h->copy_stream));
                CK(cudaEventRecord(h->copy_ev[k], h->copy_stream));
            }
            h->n_pieces = np;
            h->piece_bytes = piece;
        }
    }
    h->nbytes = nbytes;
    CK(cudaEventRecord(h->ev[EV_H2D], h->stream));

    const bool graph_directed = p->keep_directed_bidir || (!p->bidirected && p->directed);  // builders.py:143
    h->symmax = graph_directed && !p->asymmetric;                                           // builders.py:282
    h->spe = (p->bidirected && !p->keep_directed_bidir) ? 4 : 2;
    h->tpe = graph_directed ? 1 : (h->spe == 4 ? 4 : 2);
    const bool weighted = h->params.weight_tag_len > 0;

    const u64 n_tiles64 = (nbytes + WT_TILE - 1) / WT_TILE;
    const u32 n_tiles = (u32)n_tiles64;
    h->n_tiles = n_tiles;
    u64 keys_cap = h->hint_keys ? h->hint_keys + 64 : nbytes / 24 + 1024;
    u64 edge_cap = h->hint_edges ? h->hint_edges + 64 : nbytes / 20 + 1024;
    u64 long_cap = h->hint_long ? h->hint_long + h->hint_long / 4 + 1024 : 65536;
    u64 defer_cap = h->hint_defer ? h->hint_defer + h->hint_defer / 4 + 1024 : (nbytes / 2048 > 65536 ? nbytes / 2048 : 65536);
    if (spec) edge_cap += edge_cap / 16;
    if (const char* kc = getenv("G2N_DBG_KEYS")) keys_cap = (u64)atoll(kc);  // timing experiments: table size of a warm handle
    u64 seed = 0x51ed270b7a2d4c1full;
    h->reseeded = false;
    Counters& hc = *h->h_cnt;
    for (u32 attempt = 0;; attempt++) {
        if (attempt > 12) { h->err = "capacity retry limit exceeded"; return G2N_ERR_INTERNAL; }
        h->diag.retries = attempt;
        if (edge_cap > 0xFFFFFFF0ull) { h->err = "more than 2^32 edge records"; return G2N_ERR_UNSUPPORTED; }
        // load factor in (0.35, 0.7]: the smaller the table, the more of it stays resident in L2
        const u64 want_slots = keys_cap + keys_cap / 2 - keys_cap / 16;
        const u32 cap = next_pow2(want_slots < 1024 ? 1024 : want_slots);
        if (want_slots > (1ull << 31)) { h->err = "more than 2^30 distinct node keys"; return G2N_ERR_UNSUPPORTED; }
        h->table_cap = cap;
        size_t zbytes = 0;
        {
            // The tokenizer counts the rows itself (one fire-and-forget atomic per counted mention into a 4-byte-per-slot
            // side array) while table + counters stay L2-resident; beyond that the flat count pass over the edge records,
            // whose histogram is 4 bytes per NODE, is cheaper (tools/ubench/randmem.cu: atomics on a working set >> L2 run
            // at 20 G/s).  Weighted builds walk the tiles for the emission order anyway and count there.
            bool tokcnt = late_rows && !weighted && (size_t)cap * (sizeof(Slot) + sizeof(u32)) <= ((size_t)80 << 20);
            if (const char* tc = getenv("G2N_DBG_TOKCNT")) tokcnt = late_rows && !weighted && atoi(tc) != 0;
            h->count_mode = CM_NONE;
            if (tokcnt) {
                const bool sym = h->symmax;
                if (sym || h->tpe >= 2) h->count_mode = CM_ALL;
                else h->count_mode = p->want_format == G2N_FMT_CSC ? CM_DST : CM_SRC;
            }
            int rc = layout_zearly(h, cap, n_tiles, tokcnt, &zbytes);
            if (rc) return rc;
        }
        CK(cudaMemsetAsync(h->zearly.p, 0, zbytes, h->stream));
        CK(h->edge_slots.ensure((edge_cap + 1) * h->spe * sizeof(u32)));
        if (weighted) CK(h->edge_w.ensure((edge_cap + 1) * sizeof(double)));
        CK(h->longs.ensure((long_cap + 1) * sizeof(LongDesc)));
        CK(h->defer.ensure((defer_cap + 1) * sizeof(DeferEnt)));
        CK(h->tile_info.ensure(((size_t)n_tiles + 1) * sizeof(TileInfo)));
        CK(h->tile_base.ensure(((size_t)n_tiles + 2) * sizeof(u64)));
        if (spec) {
            // everything downstream is sized now, so that the kernels run back to back
            h->cap_n = cap / 2 + cap / 4;  // the table refuses more keys than this
            h->cap_E = edge_cap;
            h->cap_R = h->hint_records + h->hint_records / 16 + 1024;
            int rc = prepare_late(h, late_rows);
            if (rc) return rc;
        }
        ScanParams P;
        memset(&P, 0, sizeof(P));
        h->slow_ran = false;
        if (n_tiles > 0) {
            P.text = h->d_text;
            P.nbytes = nbytes;
            P.slots = h->d_slots;
            P.slot_cnt = h->d_slot_cnt;
            P.count_mode = h->count_mode;
            P.table_mask = cap - 1;
            P.table_max_keys = (u32)(cap / 2 + cap / 4);
            P.n_tiles = n_tiles;
            P.edge_slots = h->edge_slots.as<u32>();
            P.edge_w = weighted ? h->edge_w.as<double>() : nullptr;
            P.edge_cap = (u32)edge_cap;
            P.longs = h->longs.as<LongDesc>();
            P.long_cap = (u32)long_cap;
            P.defer = h->defer.as<DeferEnt>();
            P.defer_cap = (u32)defer_cap;
            P.tile_info = h->tile_info.as<TileInfo>();
            P.cnt = h->d_cnt;
            P.n_tiles = n_tiles;
            P.bidirected = p->bidirected ? 1 : 0;
            P.slots_per_edge = h->spe;
            P.strip_orientation = p->strip_orientation ? 1 : 0;
            P.wt_len = h->params.weight_tag_len;
            P.dtype = p->dtype;
            P.seed = seed;
            h->seed_used = seed;
            memcpy(P.wt, h->weight_tag, sizeof(P.wt));
            P.tile_begin = 0;
            P.tile_end = n_tiles;
            const int tm = (P.bidirected ? TM_BIDIR : 0) | (P.slots_per_edge == 4 ? TM_FOUR : 0) | (P.wt_len > 0 ? TM_WEIGHT : 0);
            // one launch per arrived piece of a host text (the copy of the next piece overlaps this launch)
            const u32 n_launch = h->n_pieces > 1 ? h->n_pieces : 1;
            u32 t_begin = 0;
            for (u32 k = 0; k < n_launch; k++) {
                if (h->n_pieces && !h->readers.empty()) {
                    while (!h->piece_issued[k].load(std::memory_order_acquire)) std::this_thread::yield();  // its copy is queued
                }
                if (h->n_pieces) CK(cudaStreamWaitEvent(h->stream, h->copy_ev[k], 0));
                u32 t_end = n_tiles;
                if (k + 1 < n_launch) {
                    const u64 copied = (u64)(k + 1) * h->piece_bytes;  // windows of these tiles end inside the copied prefix
                    t_end = (u32)((copied - WT_LOOK) / WT_TILE);
                }
                if (t_end <= t_begin) continue;
                P.tile_begin = t_begin;
                P.tile_end = t_end;
                KScope ks(h, "k_tokenize");
                const dim3 block(WT_WARPS * 32);
#define G2N_TK(M)                                                                                                         \
    case M: {                                                                                                             \
        if (!(h->tk_attr & (1u << (M)))) { /* per handle = per device: the opt-in is a per-device function attribute */ \
            CK(cudaFuncSetAttribute(k_tokenize<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tk_smem_bytes<M>())); \
            h->tk_attr |= 1u << (M);                                                                                      \
        }                                                                                                                 \
        k_tokenize<M><<<grid_for(t_end - t_begin, WT_WARPS, tk_min_blocks<M>()), block, tk_smem_bytes<M>(), h->stream>>>(P); \
    } break;
                switch (tm) {  // bidirected keys x four slots x weights (the combinations parse_gfa can ask for)
                    G2N_TK(0) G2N_TK(TM_WEIGHT) G2N_TK(TM_BIDIR) G2N_TK(TM_BIDIR | TM_WEIGHT) G2N_TK(TM_BIDIR | TM_FOUR) G2N_TK(TM_BIDIR | TM_FOUR | TM_WEIGHT)
                    default: h->err = "no tokenizer specialisation for this mode"; return G2N_ERR_INTERNAL;
                }
#undef G2N_TK
                t_begin = t_end;
            }
            for (std::thread& t : h->readers) t.join();
            h->readers.clear();
            if (h->reader_err.load()) { h->err = h->reader_err.load() == 1 ? "reading the input file failed" : "queueing a copy of the input file failed"; return G2N_ERR_CUDA; }
            if (h->src_fd >= 0) h->src_resident = true;
            h->n_pieces = 0;  // a capacity retry finds the whole text on the device
            P.tile_begin = 0;
            P.tile_end = n_tiles;
            CK(cudaGetLastError());
            if (getenv("G2N_DBG_TOKENIZE_ONLY")) {  // kernel-variant timing experiments: stop after the hot kernel
                cudaEvent_t e0, e1;
                cudaEventCreate(&e0); cudaEventCreate(&e1);
                CK(cudaStreamSynchronize(h->stream));
                float best = 1e9f;
                for (int rep = 0; rep < 5; rep++) {
                    CK(cudaMemsetAsync(h->zearly.p, 0, zbytes, h->stream));
                    cudaEventRecord(e0, h->stream);
                    k_tokenize<0><<<grid_for(n_tiles, WT_WARPS, tk_min_blocks<0>()), WT_WARPS * 32, tk_smem_bytes<0>(), h->stream>>>(P);
                    cudaEventRecord(e1, h->stream);
                    CK(cudaStreamSynchronize(h->stream));
                    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
                    if (ms < best) best = ms;
                }
                fprintf(stderr, "DBG k_tokenize<0> best of 5: %.4f ms (L2 warm)\n", best);
                h->err = "G2N_DBG_TOKENIZE_ONLY";
                return G2N_ERR_INTERNAL;
            }
            // (records, edge records) before every tile; the grand totals come back with the counters
            LoadTileCounts ltc{h->tile_info.as<TileInfo>()};
            int rc = launch_scan<u64>(h, ltc, h->tile_base.as<u64>(), nullptr, n_tiles, nullptr, h->d_scan_tiles);
            if (rc) return rc;
            if (spec) {
                if (h->hint_defer > 0) {
                    // the previous build of this shape deferred lines: same launch, the count is read on the device
                    { KScope ks(h, "k_tokenize_slow"); k_tokenize_slow<<<grid_for(h->hint_defer + h->hint_defer / 4 + 1024, 128), 128, 0, h->stream>>>(P); }
                    CK(cudaGetLastError());
                    h->slow_ran = true;
                }
                break;
            }
            CK(cudaMemcpyAsync(&h->h_tail[2], h->tile_base.as<u64>() + n_tiles, sizeof(u64), cudaMemcpyDeviceToHost, h->stream));
        } else {
            if (spec) break;
            h->h_tail[2] = 0;
        }
        CK(cudaEventRecord(h->ev[EV_TOKENIZE], h->stream));
        CK(cudaMemcpyAsync(&hc, h->d_cnt, sizeof(Counters), cudaMemcpyDeviceToHost, h->stream));
        CK(cudaStreamSynchronize(h->stream));
        const u32 fatal = CF_TABLE_FULL | CF_EDGE_FULL | CF_LONG_FULL | CF_DEFER_FULL;
        if (!(hc.flags & fatal) && hc.n_defer > 0) {
            // lines the hot kernel handed over: generic byte-wise parser, one line per thread
            { KScope ks(h, "k_tokenize_slow"); k_tokenize_slow<<<grid_for(hc.n_defer, 128), 128, 0, h->stream>>>(P); }
            CK(cudaGetLastError());
            h->slow_ran = true;
            CK(cudaEventRecord(h->ev[EV_TOKENIZE], h->stream));
            CK(cudaMemcpyAsync(&hc, h->d_cnt, sizeof(Counters), cudaMemcpyDeviceToHost, h->stream));
            CK(cudaStreamSynchronize(h->stream));
        }
        bool retry = false;
        if (hc.edge_alloc > edge_cap) hc.flags |= CF_EDGE_FULL;
        if (n_tiles > 0 && hc.n_keys > (u32)(cap / 2 + cap / 4)) hc.flags |= CF_TABLE_FULL;
        if (hc.flags & CF_DEFER_FULL) { defer_cap = defer_cap * 8 > (u64)hc.n_defer + 1024 ? defer_cap * 8 : (u64)hc.n_defer + 1024; retry = true; }
        if (hc.flags & CF_TABLE_FULL) { keys_cap = keys_cap * 4 > hc.n_keys * 2ull ? keys_cap * 4 : hc.n_keys * 2ull; retry = true; }
        if (hc.flags & CF_EDGE_FULL) { edge_cap = (u64)hc.edge_alloc + 64; retry = true; }
        if (hc.flags & CF_LONG_FULL) { long_cap = (u64)hc.n_long * 2 + 1024; retry = true; }
        if (!retry && hc.collision) { seed = seed * 6364136223846793005ull + 1442695040888963407ull; retry = true; h->reseeded = true; }
        if (!retry) break;
    }
    SizeCaps caps;
    caps.n_tiles = n_tiles;
    caps.tpe = h->tpe;
    caps.sym = h->symmax ? 1 : 0;
    caps.slow_ran = h->slow_ran ? 1 : 0;
    if (spec) {
        caps.n_cap = (u32)h->cap_n; caps.E_cap = (u32)h->cap_E; caps.R_cap = (u32)h->cap_R;
        k_sizes<<<1, 32, 0, h->stream>>>(h->d_cnt, h->tile_base.as<u64>(), caps, h->d_ds);
        CK(cudaGetLastError());
        CK(cudaEventRecord(h->ev[EV_TOKENIZE], h->stream));
        return G2N_OK;
    }
    h->hint_keys = hc.n_keys;
    h->hint_edges = hc.edge_alloc;
    h->hint_long = hc.n_long;
    h->hint_defer = hc.n_defer;
    const u64 R = h->h_tail[2] >> 32;
    h->hint_records = R;
    collect_diag(h, hc, R);
    if (h->diag.err_kind) {
        h->diag.gpu_launches = h->launches;
        h->err = "input holds a record the reference raises on";
        return G2N_ERR_PARSE;
    }
    if (R >= (1ull << 30) - 1) { h->err = "more than 2^30 records in one build"; return G2N_ERR_UNSUPPORTED; }
    const u64 n = hc.n_keys;
    const u64 E = hc.edge_alloc;
    if (n > 0x7FFFFFFFull) { h->err = "more than 2^31-1 nodes (int64 indices) is out of scope"; return G2N_ERR_UNSUPPORTED; }
    if (E * (u64)h->tpe * (h->symmax ? 2 : 1) >= 0xFFFFFFF0ull) { h->err = "more than 2^32 triplets in one build"; return G2N_ERR_UNSUPPORTED; }
    h->n_nodes = n;
    h->n_edges = E;
    h->n_records = R;
    h->n_long = hc.n_long;
    h->cap_n = n;
    h->cap_E = E;
    h->cap_R = R;
    caps.n_cap = (u32)n; caps.E_cap = (u32)(E > edge_cap ? E : edge_cap); caps.R_cap = (u32)R;
    k_sizes<<<1, 32, 0, h->stream>>>(h->d_cnt, h->tile_base.as<u64>(), caps, h->d_ds);
    CK(cudaGetLastError());
    int rc = prepare_late(h, late_rows);
    if (rc) return rc;
    return G2N_OK;
}

// Phase 2 on one GPU: first-appearance ranking -> node IDs (slot_id, id2slot, name lengths).
static int ids_phase(g2n_handle* h)
{
    const u64 n = h->cap_n, R = h->cap_R;
    const u32 cap = h->table_cap;
    const u64 words = (4 * R + 31) / 32 + 1;
    CK(h->wprefix.ensure((words + 2) * sizeof(u32)));
    CK(h->slot_id.ensure((size_t)cap * sizeof(u32)));
    CK(h->id2slot.ensure((n + 1) * sizeof(u32)));
    CK(h->name_off.ensure((n + 2) * sizeof(u64)));
    if (n > 0) {
        { KScope ks(h, "k_mark_first"); k_mark_first<<<grid_for(cap, 256), 256, 0, h->stream>>>(h->d_slots, cap, h->tile_base.as<u64>(), h->d_bitmap, h->d_ds); }
        LoadPopc8 lp{h->d_bitmap};
        int rc = launch_scan<u32>(h, lp, h->wprefix.as<u32>(), nullptr, words / BM_GROUP + 1, &h->d_ds->wgroups, h->d_scan_words);
        if (rc) return rc;
        { KScope ks(h, "k_assign_ids"); k_assign_ids<<<grid_for(cap, 256), 256, 0, h->stream>>>(h->d_slots, cap, h->tile_base.as<u64>(), h->d_bitmap, h->wprefix.as<u32>(),
                                                                 h->slot_id.as<u32>(), h->id2slot.as<u32>(), h->d_ds,
                                                                 h->count_mode != CM_NONE ? h->d_slot_cnt : nullptr, h->d_rowcnt); }
        CK(cudaGetLastError());
    }
    h->names_sized = false;  // name offsets are scanned on demand (g2n_names_bytes / g2n_fetch_names)
    CK(cudaEventRecord(h->ev[EV_IDS], h->stream));
    h->have_edges = true;
    return G2N_OK;
}

// names held by this handle: all nodes, or -- multi-GPU slab -- the nodes that first appear in this rank's shard
static u64 names_count(const g2n_handle* h) { return h->slab_mode ? h->names_n : h->n_nodes; }

// name lengths (from the table slots, ids.cuh: LoadNameLen) -> name_off (exclusive scan) and the total; only when somebody asks for the node names
static int size_names(g2n_handle* h)
{
    if (h->names_sized) return G2N_OK;
    const u64 n = names_count(h);
    if (n > 0) {
        LoadNameLen ln{h->d_slots, h->id2slot.as<u32>()};
        int rc = launch_scan<u64>(h, ln, (u64*)h->name_off.as<u64>(), nullptr, n, nullptr, nullptr);
        if (rc) return rc;
        CK(cudaMemcpyAsync(&h->h_tail[1], h->name_off.as<u64>() + n, sizeof(u64), cudaMemcpyDeviceToHost, h->stream));
    } else {
        CK(cudaMemsetAsync(h->name_off.p, 0, 2 * sizeof(u64), h->stream));
        h->h_tail[1] = 0;
    }
    CK(cudaStreamSynchronize(h->stream));
    h->names_bytes = h->h_tail[1];
    h->names_sized = true;
    return G2N_OK;
}

static int build_once(g2n_handle* h, const uint8_t* text, uint64_t nbytes, const g2n_params* p, bool spec)
{
    const bool graph_directed = p->keep_directed_bidir || (!p->bidirected && p->directed);
    const bool rows = (graph_directed && !p->asymmetric) || p->want_format != G2N_FMT_NATIVE;
    int rc = tokenize_phase(h, text, nbytes, p, spec, rows);
    if (rc) return rc;
    h->slab_mode = false;
    rc = ids_phase(h);
    if (rc) return rc;
    // ---- K3 + K4
    const bool counted = h->count_mode != CM_NONE;  // the rows were counted by the tokenizer (rowcnt filled by k_assign_ids)
    if (h->symmax) rc = build_compressed(h, p->want_format == G2N_FMT_CSC ? G2N_FMT_CSC : G2N_FMT_CSR, true, counted);
    else if (p->want_format == G2N_FMT_NATIVE) rc = build_coo(h);
    else rc = build_compressed(h, p->want_format, true, counted);
    if (rc) return rc;
    return finish_result(h);
}

int g2n_build(g2n_handle* h, const uint8_t* text, uint64_t nbytes, const g2n_params* p)
{
    if (!h || !p) return G2N_ERR_INVALID;
    const int wt_len = (p->weight_tag && p->weight_tag_len > 0) ? p->weight_tag_len : 0;
    const u64 sig = shape_signature(p, nbytes, wt_len);
    int rc = G2N_SPEC_MISS;
    if (h->speculate && h->hint_valid && h->hint_sig == sig && h->hint_keys > 0 && h->hint_edges > 0)
        rc = build_once(h, text, nbytes, p, true);
    if (rc == G2N_SPEC_MISS) {
        h->hint_valid = false;
        rc = build_once(h, text, nbytes, p, false);
    }
    if (rc) return rc;
    h->hint_sig = sig;
    h->hint_valid = true;
    h->built = true;
    return G2N_OK;
}

int g2n_build_file(g2n_handle* h, const char* path, const g2n_params* p)
{
    if (!h || !path || !p) return G2N_ERR_INVALID;
    const int fd = open(path, O_RDONLY);
    if (fd < 0) { h->err = std::string("cannot open ") + path; return G2N_ERR_INVALID; }
    struct stat st;
    if (fstat(fd, &st) != 0 || !S_ISREG(st.st_mode)) { close(fd); h->err = "not a regular file"; return G2N_ERR_INVALID; }
    g2n_params q = *p;
    q.text_on_device = 0;
    h->src_fd = fd;
    h->src_resident = false;
    // any non-null pointer: the bytes come from the file, never from this address
    const int rc = g2n_build(h, (const uint8_t*)h, (uint64_t)st.st_size, &q);
    for (std::thread& t : h->readers) t.join();  // (only after an early error return)
    h->readers.clear();
    h->src_fd = -1;
    h->src_resident = false;
    close(fd);
    return rc;
}

// ---- *.gz input (parser.py:108-109 gzip.open).  The reference inflates the whole stream on one core before
// anything else happens; here the compressed file is mapped, inflated in windows of 64 MiB into two pinned staging
// buffers and every finished window is copied to the device while the next one is being inflated.  A BGZF file
// (bgzip: independent <= 64 KiB deflate blocks that announce their size in a gzip extra field) is inflated by all
// host cores, a plain gzip stream (one or several members) by one, as fast as zlib goes.  The build then runs on the
// device-resident text.  Any irregularity in the container -> G2N_ERR_INVALID: the caller falls back to the
// reference's own host inflate, which raises the reference's own exception.
namespace {

#define G2N_GZ_WINDOW ((u64)64 << 20)

struct BgzfBlock { u64 in_off; u32 in_len, out_len; u64 out_off; };

// walks the members of a BGZF file; false if any member is not a BGZF block
bool bgzf_index(const uint8_t* z, u64 n, std::vector<BgzfBlock>& blocks, u64& total)
{
    u64 p = 0;
    total = 0;
    while (p < n) {
        if (n - p < 28 || z[p] != 0x1f || z[p + 1] != 0x8b || z[p + 2] != 8 || !(z[p + 3] & 4)) return false;
        const u32 xlen = z[p + 10] | (z[p + 11] << 8);
        if (n - p < 12 + (u64)xlen + 8) return false;
        u32 bsize = 0;
        bool found = false;
        for (u32 q = 0; q + 4 <= xlen;) {
            const uint8_t* f = z + p + 12 + q;
            const u32 slen = f[2] | (f[3] << 8);
            if (f[0] == 'B' && f[1] == 'C' && slen == 2 && q + 6 <= xlen) { bsize = (f[4] | (f[5] << 8)) + 1u; found = true; }
            q += 4 + slen;
        }
        if (!found || (z[p + 3] & ~4) || bsize < 12 + xlen + 8 || p + bsize > n) return false;
        BgzfBlock b;
        b.in_off = p + 12 + xlen;
        b.in_len = bsize - 12 - xlen - 8;
        const uint8_t* tr = z + p + bsize - 4;
        b.out_len = tr[0] | (tr[1] << 8) | (tr[2] << 16) | ((u32)tr[3] << 24);
        b.out_off = total;
        if (b.out_len > (1u << 16)) return false;
        total += b.out_len;
        blocks.push_back(b);
        p += bsize;
    }
    return true;
}

bool bgzf_inflate(const uint8_t* z, const BgzfBlock& b, uint8_t* out)
{
    z_stream zs;
    memset(&zs, 0, sizeof zs);
    if (inflateInit2(&zs, -15) != Z_OK) return false;
    zs.next_in = const_cast<Bytef*>(z + b.in_off);
    zs.avail_in = b.in_len;
    zs.next_out = out;
    zs.avail_out = b.out_len;
    const int rc = b.out_len || b.in_len ? inflate(&zs, Z_FINISH) : Z_STREAM_END;
    const bool ok = rc == Z_STREAM_END && zs.avail_out == 0;
    inflateEnd(&zs);
    if (!ok) return false;
    const uint8_t* tr = z + b.in_off + b.in_len;  // CRC32 | ISIZE
    const u32 want = tr[0] | (tr[1] << 8) | (tr[2] << 16) | ((u32)tr[3] << 24);
    return (u32)crc32(crc32(0L, Z_NULL, 0), out, b.out_len) == want;
}

}  // namespace

int g2n_build_gz(g2n_handle* h, const char* path, const g2n_params* p)
{
    if (!h || !path || !p) return G2N_ERR_INVALID;
    CK(cudaSetDevice(h->device));
    const int fd = open(path, O_RDONLY);
    if (fd < 0) { h->err = std::string("cannot open ") + path; return G2N_ERR_INVALID; }
    struct stat st;
    if (fstat(fd, &st) != 0 || !S_ISREG(st.st_mode)) { close(fd); h->err = "not a regular file"; return G2N_ERR_INVALID; }
    const u64 zn = (u64)st.st_size;
    const uint8_t* z = zn ? (const uint8_t*)mmap(nullptr, zn, PROT_READ, MAP_PRIVATE, fd, 0) : nullptr;
    close(fd);
    if (zn && z == (const uint8_t*)MAP_FAILED) { h->err = "cannot map the compressed file"; return G2N_ERR_INVALID; }
    int rc = G2N_OK;
    u64 total = 0;
    auto fail = [&](const char* msg) { h->err = msg; rc = G2N_ERR_INVALID; };
    do {
        if (cudaStreamSynchronize(h->stream) != cudaSuccess) { rc = G2N_ERR_CUDA; break; }
        for (int k = 0; k < 2 && !rc && !h->gz_stage[1]; k++) {  // two pinned staging windows (the file reader's 8 MiB pieces are too small)
            if (h->gz_stage[k]) continue;
            if (cudaHostAlloc(&h->gz_stage[k], G2N_GZ_WINDOW, cudaHostAllocDefault) != cudaSuccess ||
                cudaEventCreateWithFlags(&h->gz_ev[k], cudaEventDisableTiming) != cudaSuccess) { h->err = "pinned staging for gz input"; rc = G2N_ERR_CUDA; }
        }
        if (rc) break;
        std::vector<BgzfBlock> blocks;
        const bool bgzf = zn > 0 && bgzf_index(z, zn, blocks, total);
        if (bgzf) {
            // ---- block-parallel: windows of whole blocks, all host cores per window
            if (h->text.ensure(total + 64) != cudaSuccess) { h->err = "out of device memory for the text"; rc = G2N_ERR_CUDA; break; }
            const unsigned T = std::max(1u, std::min(32u, std::thread::hardware_concurrency()));
            size_t b0 = 0;
            u32 w = 0;
            while (b0 < blocks.size() && !rc) {
                size_t b1 = b0;
                while (b1 < blocks.size() && blocks[b1].out_off + blocks[b1].out_len - blocks[b0].out_off <= G2N_GZ_WINDOW) b1++;
                const u32 sb = w & 1;
                cudaEventSynchronize(h->gz_ev[sb]);
                uint8_t* stage = (uint8_t*)h->gz_stage[sb];
                const u64 base = blocks[b0].out_off;
                std::atomic<size_t> next{b0};
                std::atomic<int> bad{0};
                std::vector<std::thread> pool;
                for (unsigned t = 0; t < T; t++)
                    pool.emplace_back([&]() {
                        for (;;) {
                            const size_t i = next.fetch_add(8);
                            if (i >= b1) break;
                            for (size_t j = i; j < std::min(i + 8, b1); j++)
                                if (!bgzf_inflate(z, blocks[j], stage + (blocks[j].out_off - base))) bad.store(1);
                        }
                    });
                for (std::thread& t : pool) t.join();
                if (bad.load()) { fail("corrupt BGZF block"); break; }
                const u64 len = blocks[b1 - 1].out_off + blocks[b1 - 1].out_len - base;
                if (len && (cudaMemcpyAsync(h->text.as<uint8_t>() + base, stage, len, cudaMemcpyHostToDevice, h->copy_stream) != cudaSuccess ||
                            cudaEventRecord(h->gz_ev[sb], h->copy_stream) != cudaSuccess)) { h->err = "queueing a copy of the inflated text failed"; rc = G2N_ERR_CUDA; }
                b0 = b1;
                w++;
            }
        } else {
            // ---- one gzip stream (possibly several concatenated members): streaming inflate, window by window
            z_stream zs;
            memset(&zs, 0, sizeof zs);
            if (inflateInit2(&zs, 15 + 16) != Z_OK) { fail("zlib init"); break; }
            zs.next_in = const_cast<Bytef*>(z);
            u64 in_left = zn;
            u64 cap = 0;
            {
                // size guess: ISIZE of the last member (exact for one member below 4 GiB), at least 3x the compressed size
                u64 guess = zn * 3;
                if (zn >= 18) { const uint8_t* tr = z + zn - 4; const u64 isz = tr[0] | (tr[1] << 8) | (tr[2] << 16) | ((u64)tr[3] << 24); if (isz > guess) guess = isz; }
                cap = guess + (1 << 20);
                if (h->text.ensure(cap + 64) != cudaSuccess) { h->err = "out of device memory for the text"; rc = G2N_ERR_CUDA; inflateEnd(&zs); break; }
                cap = h->text.cap - 64;
            }
            bool done = zn == 0;
            u32 w = 0;
            if (zn == 0) fail("empty gz file");
            while (!done && !rc) {
                const u32 sb = w & 1;
                cudaEventSynchronize(h->gz_ev[sb]);
                uint8_t* stage = (uint8_t*)h->gz_stage[sb];
                u64 filled = 0;
                while (filled < G2N_GZ_WINDOW && !done && !rc) {
                    const u64 room = G2N_GZ_WINDOW - filled;
                    zs.next_out = stage + filled;
                    zs.avail_out = (uInt)std::min<u64>(room, 1u << 30);
                    if (zs.avail_in == 0 && in_left) { const u64 take = std::min<u64>(in_left, 1u << 30); zs.avail_in = (uInt)take; in_left -= take; }
                    const uInt before = zs.avail_out;
                    const int zr = inflate(&zs, Z_NO_FLUSH);
                    filled += before - zs.avail_out;
                    if (zr == Z_STREAM_END) {
                        if (zs.avail_in == 0 && in_left == 0) done = true;
                        else if (inflateReset(&zs) != Z_OK) fail("zlib reset");  // next member
                    } else if (zr != Z_OK) {
                        fail(zr == Z_BUF_ERROR ? "truncated gz stream" : "corrupt gz stream");
                    } else if (zs.avail_in == 0 && in_left == 0 && before == zs.avail_out) {
                        fail("truncated gz stream");
                    }
                }
                if (rc) break;
                if (total + filled > cap) {
                    // the guess was too small: a larger buffer, what is on the device already moves over
                    if (cudaStreamSynchronize(h->copy_stream) != cudaSuccess) { rc = G2N_ERR_CUDA; break; }
                    void* bigger = nullptr;
                    const u64 want = (total + filled) * 2 + (64 << 20);
                    if (cudaMalloc(&bigger, want + 64) != cudaSuccess) { h->err = "out of device memory for the text"; rc = G2N_ERR_CUDA; break; }
                    if (total && cudaMemcpy(bigger, h->text.p, total, cudaMemcpyDeviceToDevice) != cudaSuccess) { cudaFree(bigger); rc = G2N_ERR_CUDA; break; }
                    cudaFree(h->text.p);
                    h->text.p = bigger;
                    h->text.cap = want + 64;
                    cap = want;
                }
                if (filled && (cudaMemcpyAsync(h->text.as<uint8_t>() + total, stage, filled, cudaMemcpyHostToDevice, h->copy_stream) != cudaSuccess ||
                               cudaEventRecord(h->gz_ev[sb], h->copy_stream) != cudaSuccess)) { h->err = "queueing a copy of the inflated text failed"; rc = G2N_ERR_CUDA; }
                total += filled;
                w++;
            }
            inflateEnd(&zs);
        }
        if (!rc && cudaStreamSynchronize(h->copy_stream) != cudaSuccess) rc = G2N_ERR_CUDA;
    } while (0);
    if (zn) munmap((void*)z, zn);
    if (rc) return rc;
    g2n_params q = *p;
    q.text_on_device = 1;
    return g2n_build(h, h->text.as<uint8_t>(), total, &q);
}

int g2n_convert(g2n_handle* h, int32_t want_format)
{
    if (!h || !h->built || !h->have_edges) return G2N_ERR_INVALID;
    if (want_format != G2N_FMT_CSR && want_format != G2N_FMT_CSC) { h->err = "convert target must be CSR or CSC"; return G2N_ERR_INVALID; }
    CK(cudaSetDevice(h->device));
    if (h->result_format == want_format) return G2N_OK;
    if (h->symmax) { h->result_format = want_format; return G2N_OK; }  // symmetric: same arrays
    CK(cudaEventRecord(h->ev[EV_START], h->stream));
    CK(cudaEventRecord(h->ev[EV_H2D], h->stream));
    CK(cudaEventRecord(h->ev[EV_TOKENIZE], h->stream));
    CK(cudaEventRecord(h->ev[EV_IDS], h->stream));
    h->spec = false;
    h->cap_n = h->n_nodes;
    h->cap_E = h->n_edges;
    int rc = build_compressed(h, want_format, false, false);
    if (rc) return rc;
    return finish_result(h);
}

int g2n_sizes(g2n_handle* h, g2n_sizes_t* out)
{
    if (!h || !out || !h->built) return G2N_ERR_INVALID;
    out->n_nodes = h->n_nodes;  /* in slab mode: nodes of the whole graph (= number of columns) */
    out->nnz = h->nnz;
    out->names_bytes = h->names_sized ? h->names_bytes : 0;  /* filled once g2n_names_bytes ran */
    out->format = h->result_format;
    out->index_bytes = 4;
    out->dtype = h->params.dtype;
    out->reserved = 0;
    out->slab_rows = h->slab_mode ? h->slab_rows : h->n_nodes;
    out->names_count = names_count(h);
    out->names_id0 = h->slab_mode ? h->names_id0 : 0;
    return G2N_OK;
}

int g2n_device_result(g2n_handle* h, void** a0, void** a1, void** data)
{
    if (!h || !h->built) return G2N_ERR_INVALID;
    const bool coo = h->result_format == G2N_FMT_COO;
    if (a0) *a0 = coo ? h->row.p : h->indptr.p;
    if (a1) *a1 = coo ? h->col.p : h->indices.p;
    if (data) *data = h->data.p;
    return G2N_OK;
}

int g2n_fetch_matrix(g2n_handle* h, void* a0, void* a1, void* data)
{
    if (!h || !h->built) return G2N_ERR_INVALID;
    CK(cudaSetDevice(h->device));
    const size_t ds = dtype_size(h->params.dtype);
    if (h->result_format == G2N_FMT_COO) {
        if (h->nnz) {
            CK(cudaMemcpyAsync(a0, h->row.p, h->nnz * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
            CK(cudaMemcpyAsync(a1, h->col.p, h->nnz * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
            CK(cudaMemcpyAsync(data, h->data.p, h->nnz * ds, cudaMemcpyDeviceToHost, h->stream));
        }
    } else {
        const u64 rows = h->slab_mode ? h->slab_rows : h->n_nodes;
        CK(cudaMemcpyAsync(a0, h->indptr.p, (rows + 1) * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
        if (h->nnz) {
            CK(cudaMemcpyAsync(a1, h->indices.p, h->nnz * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
            CK(cudaMemcpyAsync(data, h->data.p, h->nnz * ds, cudaMemcpyDeviceToHost, h->stream));
        }
    }
    CK(cudaStreamSynchronize(h->stream));
    return G2N_OK;
}

int g2n_names_bytes(g2n_handle* h, uint64_t* out)
{
    if (!h || !out || !h->built) return G2N_ERR_INVALID;
    CK(cudaSetDevice(h->device));
    int rc = size_names(h);
    if (rc) return rc;
    *out = h->names_bytes;
    return G2N_OK;
}

int g2n_fetch_names(g2n_handle* h, uint8_t* names, uint64_t* offsets)
{
    if (!h || !h->built) return G2N_ERR_INVALID;
    CK(cudaSetDevice(h->device));
    {
        int rc = size_names(h);
        if (rc) return rc;
    }
    if (!h->names_ready) {
        CK(h->names.ensure(h->names_bytes + 16));
        const u64 nn = names_count(h);
        if (nn > 0) {
            { KScope ks(h, "k_gather_names"); k_gather_names<<<grid_for(nn, 256), 256, 0, h->stream>>>(h->d_slots, h->id2slot.as<u32>(), h->name_off.as<u64>(),
                                                                              (u32)nn, h->d_text, h->longs.as<LongDesc>(),
                                                                              h->names.as<uint8_t>()); }
            CK(cudaGetLastError());
        }
        h->names_ready = true;
    }
    CK(cudaMemcpyAsync(offsets, h->name_off.p, (names_count(h) + 1) * sizeof(u64), cudaMemcpyDeviceToHost, h->stream));
    if (h->names_bytes) CK(cudaMemcpyAsync(names, h->names.p, h->names_bytes, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return G2N_OK;
}

// "<index>\t<name>\n" per node (utils.py:108-114), made on the device
static int build_tsv(g2n_handle* h)
{
    if (h->tsv_ready) return G2N_OK;
    if (h->slab_mode) { h->err = "the node map of a multi-GPU slab is assembled by the caller (g2n_fetch_names per rank)"; return G2N_ERR_INVALID; }
    const u64 n = h->n_nodes;
    CK(h->tsv_off.ensure((n + 2) * sizeof(u64)));
    if (n > 0) {
        LoadTsvLen ll{LoadNameLen{h->d_slots, h->id2slot.as<u32>()}};
        int rc = launch_scan<u64>(h, ll, h->tsv_off.as<u64>(), nullptr, n, nullptr, nullptr);
        if (rc) return rc;
        CK(cudaMemcpyAsync(&h->h_tail[4], h->tsv_off.as<u64>() + n, sizeof(u64), cudaMemcpyDeviceToHost, h->stream));
        CK(cudaStreamSynchronize(h->stream));
    } else {
        h->h_tail[4] = 0;
    }
    h->tsv_bytes = h->h_tail[4];
    CK(h->tsv.ensure(h->tsv_bytes + 16));
    if (n > 0) {
        KScope ks(h, "k_tsv_write");
        k_tsv_write<<<grid_for(n, 256), 256, 0, h->stream>>>(h->d_slots, h->id2slot.as<u32>(), h->tsv_off.as<u64>(), (u32)n, h->d_text,
                                                              h->longs.as<LongDesc>(), h->tsv.as<uint8_t>());
        CK(cudaGetLastError());
    }
    h->tsv_ready = true;
    return G2N_OK;
}

int g2n_nodes_tsv_bytes(g2n_handle* h, uint64_t* out)
{
    if (!h || !out || !h->built) return G2N_ERR_INVALID;
    CK(cudaSetDevice(h->device));
    int rc = build_tsv(h);
    if (rc) return rc;
    *out = h->tsv_bytes;
    return G2N_OK;
}

int g2n_fetch_nodes_tsv(g2n_handle* h, uint8_t* out)
{
    if (!h || !h->built) return G2N_ERR_INVALID;
    CK(cudaSetDevice(h->device));
    int rc = build_tsv(h);
    if (rc) return rc;
    if (h->tsv_bytes) CK(cudaMemcpyAsync(out, h->tsv.p, h->tsv_bytes, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return G2N_OK;
}

// "<u>\t<v>\n" per edge record in file order (cli.py:267-281), made on the device from the last build
static int build_edge_list(g2n_handle* h)
{
    if (h->el_ready) return G2N_OK;
    if (h->slab_mode || !h->have_edges) { h->err = "the edge list needs a single-GPU build"; return G2N_ERR_INVALID; }
    const u64 E = h->n_edges;
    CK(h->el_len.ensure((E + 2) * sizeof(u32)));
    CK(h->el_off.ensure((E + 2) * sizeof(u64)));
    h->el_bytes = 0;
    if (E > 0) {
        h->spec = false;
        EmitParams EP = emit_params(h);
        EP.ids_ready = 1;  // hand the stored words through untouched; NameSrc maps IDs back to slots if need be
        EP.write_ids = 0;
        NameSrc N;
        N.slots = h->d_slots; N.longs = h->longs.as<LongDesc>(); N.text = h->d_text;
        N.id2slot = h->edges_are_ids ? h->id2slot.as<u32>() : nullptr;
        const u32 egrid = grid_for((u64)h->n_tiles * 32, 256);
        { KScope ks(h, "k_edge_line_len"); k_edge_line_len<<<egrid, 256, 0, h->stream>>>(EP, N, h->el_len.as<u32>()); }
        CK(cudaGetLastError());
        LoadArray<u32> ll{h->el_len.as<u32>()};
        int rc = launch_scan<u64>(h, ll, h->el_off.as<u64>(), nullptr, E, nullptr, nullptr);
        if (rc) return rc;
        CK(cudaMemcpyAsync(&h->h_tail[5], h->el_off.as<u64>() + E, sizeof(u64), cudaMemcpyDeviceToHost, h->stream));
        CK(cudaStreamSynchronize(h->stream));
        h->el_bytes = h->h_tail[5];
        CK(h->el_text.ensure(h->el_bytes + 16));
        { KScope ks(h, "k_edge_line_write"); k_edge_line_write<<<egrid, 256, 0, h->stream>>>(EP, N, h->el_off.as<u64>(), h->el_text.as<uint8_t>()); }
        CK(cudaGetLastError());
    }
    h->el_ready = true;
    return G2N_OK;
}

int g2n_edge_list_bytes(g2n_handle* h, uint64_t* out)
{
    if (!h || !out || !h->built) return G2N_ERR_INVALID;
    CK(cudaSetDevice(h->device));
    int rc = build_edge_list(h);
    if (rc) return rc;
    *out = h->el_bytes;
    return G2N_OK;
}

int g2n_fetch_edge_list(g2n_handle* h, uint8_t* out)
{
    if (!h || !h->built) return G2N_ERR_INVALID;
    CK(cudaSetDevice(h->device));
    int rc = build_edge_list(h);
    if (rc) return rc;
    if (h->el_bytes) CK(cudaMemcpyAsync(out, h->el_text.p, h->el_bytes, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return G2N_OK;
}

// ---- distances on the resident CSR (analysis.py:116-161, 180-272)
static int bfs_check(g2n_handle* h, int32_t slot)
{
    if (!h->built || h->slab_mode || h->result_format == G2N_FMT_COO) { h->err = "distances need the CSR/CSC result of a single-GPU build"; return G2N_ERR_INVALID; }
    if (slot < 0 || slot >= h->bfs_slots || h->bfs_n != h->n_nodes) { h->err = "no such level slot (g2n_bfs first)"; return G2N_ERR_INVALID; }
    return G2N_OK;
}

static int bfs_run(g2n_handle* h, const int32_t* sources, uint64_t n_sources, bool on_device, int32_t slot, int32_t n_slots)
{
    if (!h || (!sources && n_sources) || n_slots < 1 || slot < 0 || slot >= n_slots) return G2N_ERR_INVALID;
    if (!h->built || h->slab_mode || h->result_format == G2N_FMT_COO) { h->err = "distances need the CSR/CSC result of a single-GPU build"; return G2N_ERR_INVALID; }
    CK(cudaSetDevice(h->device));
    const u64 n = h->n_nodes;
    if (h->bfs_slots != n_slots || h->bfs_n != n) {
        CK(h->bfs_levels.ensure((size_t)n_slots * (n + 1) * sizeof(int32_t)));
        h->bfs_slots = n_slots;
        h->bfs_n = n;
    }
    CK(h->bfs_q0.ensure((n + 1) * sizeof(u32)));
    CK(h->bfs_q1.ensure((n + 1) * sizeof(u32)));
    CK(h->bfs_ctl.ensure(sizeof(BfsCtl)));
    int32_t* level = h->bfs_levels.as<int32_t>() + (size_t)slot * (n + 1);
    CK(cudaMemsetAsync(level, 0xFF, (n + 1) * sizeof(int32_t), h->stream));
    CK(cudaMemsetAsync(h->bfs_ctl.p, 0, sizeof(BfsCtl), h->stream));
    if (n == 0 || n_sources == 0) return G2N_OK;
    const int32_t* dsrc = sources;
    if (!on_device) {
        CK(h->bfs_nodes.ensure((n_sources + 1) * sizeof(int32_t)));
        CK(cudaMemcpyAsync(h->bfs_nodes.p, sources, n_sources * sizeof(int32_t), cudaMemcpyHostToDevice, h->stream));
        dsrc = h->bfs_nodes.as<int32_t>();
    }
    { KScope ks(h, "k_bfs_seed"); k_bfs_seed<<<grid_for(n_sources, 256), 256, 0, h->stream>>>(dsrc, n_sources, (u32)n, level, h->bfs_q0.as<u32>(), h->bfs_ctl.as<BfsCtl>()); }
    CK(cudaGetLastError());
    {
        KScope ks(h, "k_bfs_gang");
        int per_sm = 0;
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_bfs_gang, 256, 0));
        if (per_sm < 1) { h->err = "k_bfs_gang cannot be made co-resident"; return G2N_ERR_INTERNAL; }
        const dim3 grid((unsigned)(G2N_SM_COUNT * (per_sm > 4 ? 4 : per_sm)));
        const int32_t* indptr = h->indptr.as<int32_t>();
        const int32_t* indices = h->indices.as<int32_t>();
        u32 *q0 = h->bfs_q0.as<u32>(), *q1 = h->bfs_q1.as<u32>();
        BfsCtl* ctl = h->bfs_ctl.as<BfsCtl>();
        void* args[] = {(void*)&indptr, (void*)&indices, (void*)&level, (void*)&q0, (void*)&q1, (void*)&ctl};
        CK(cudaLaunchCooperativeKernel((const void*)k_bfs_gang, grid, dim3(256), args, 0, h->stream));
    }
    return G2N_OK;
}

int g2n_bfs(g2n_handle* h, const int32_t* sources, uint64_t n_sources, int32_t slot, int32_t n_slots)
{
    return bfs_run(h, sources, n_sources, false, slot, n_slots);
}

static int levels_reduce_run(g2n_handle* h, int32_t slot, const int32_t* nodes, uint64_t n_nodes, bool on_device, int64_t* out3)
{
    if (!h || !out3 || (!nodes && n_nodes)) return G2N_ERR_INVALID;
    CK(cudaSetDevice(h->device));
    int rc = bfs_check(h, slot);
    if (rc) return rc;
    const u64 n = h->n_nodes;
    CK(h->bfs_out.ensure(4 * sizeof(long long)));
    const long long init[3] = {0x7fffffffffffffffLL, 0, 0};
    CK(cudaMemcpyAsync(h->bfs_out.p, init, sizeof init, cudaMemcpyHostToDevice, h->stream));
    if (n_nodes) {
        const int32_t* dn = nodes;
        if (!on_device) {
            CK(h->bfs_nodes.ensure((n_nodes + 1) * sizeof(int32_t)));
            CK(cudaMemcpyAsync(h->bfs_nodes.p, nodes, n_nodes * sizeof(int32_t), cudaMemcpyHostToDevice, h->stream));
            dn = h->bfs_nodes.as<int32_t>();
        }
        KScope ks(h, "k_levels_reduce");
        k_levels_reduce<<<grid_for(n_nodes, 256), 256, 0, h->stream>>>(h->bfs_levels.as<int32_t>() + (size_t)slot * (n + 1), dn, n_nodes, (u32)n, h->bfs_out.as<long long>());
    }
    long long res[3];
    CK(cudaMemcpyAsync(res, h->bfs_out.p, sizeof res, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    out3[0] = res[2] ? res[0] : -1;
    out3[1] = res[1];
    out3[2] = res[2];
    return G2N_OK;
}

int g2n_levels_reduce(g2n_handle* h, int32_t slot, const int32_t* nodes, uint64_t n_nodes, int64_t* out3)
{
    return levels_reduce_run(h, slot, nodes, n_nodes, false, out3);
}

// ---- P / O records resolved on the device (paths.cuh)
int g2n_paths_load(g2n_handle* h, uint64_t* n_paths)
{
    if (!h || !n_paths) return G2N_ERR_INVALID;
    if (!h->built || h->slab_mode || !h->have_edges) { h->err = "paths need a single-GPU build"; return G2N_ERR_INVALID; }
    CK(cudaSetDevice(h->device));
    if (h->paths_ready) { *n_paths = h->h_paths.size(); return G2N_OK; }
    const u64 N = h->nbytes;
    h->h_paths.clear();
    h->h_path_entry0.assign(1, 0);
    CK(h->path_misc.ensure(256));
    u32 cap = 1u << 16, found = 0;
    for (int attempt = 0; attempt < 2 && N; attempt++) {
        CK(h->path_starts.ensure((size_t)cap * sizeof(u64)));
        CK(cudaMemsetAsync(h->path_misc.p, 0, 256, h->stream));
        { KScope ks(h, "k_paths_find"); k_paths_find<<<grid_for((N + 15) / 16, 256), 256, 0, h->stream>>>(h->d_text, N, h->path_starts.as<u64>(), cap, h->path_misc.as<u32>()); }
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(&found, h->path_misc.p, sizeof(u32), cudaMemcpyDeviceToHost, h->stream));
        CK(cudaStreamSynchronize(h->stream));
        if (found <= cap) break;
        cap = found;
    }
    const u32 R = found;
    if (R) {
        std::vector<u64> starts(R);
        CK(cudaMemcpy(starts.data(), h->path_starts.p, (size_t)R * sizeof(u64), cudaMemcpyDeviceToHost));
        std::sort(starts.begin(), starts.end());
        h->h_paths.resize(R);
        for (u32 r = 0; r < R; r++) { memset(&h->h_paths[r], 0, sizeof(PathRec)); h->h_paths[r].line = starts[r]; }
        CK(h->path_recs.ensure((size_t)R * sizeof(PathRec)));
        CK(cudaMemcpyAsync(h->path_recs.p, h->h_paths.data(), (size_t)R * sizeof(PathRec), cudaMemcpyHostToDevice, h->stream));
        { KScope ks(h, "k_path_fields"); k_path_fields<<<R < 4u * G2N_SM_COUNT ? R : 4u * G2N_SM_COUNT, 256, 0, h->stream>>>(h->d_text, N, h->path_recs.as<PathRec>(), R); }
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(h->h_paths.data(), h->path_recs.p, (size_t)R * sizeof(PathRec), cudaMemcpyDeviceToHost, h->stream));
        CK(cudaStreamSynchronize(h->stream));
        u64 n_blocks = 0;
        for (u32 r = 0; r < R; r++) {
            h->h_paths[r].first_blk = n_blocks;
            n_blocks += (h->h_paths[r].list_end - h->h_paths[r].list_off + 1 + PB_BYTES - 1) / PB_BYTES;
        }
        CK(cudaMemcpyAsync(h->path_recs.p, h->h_paths.data(), (size_t)R * sizeof(PathRec), cudaMemcpyHostToDevice, h->stream));
        CK(h->path_cnt.ensure((n_blocks + 1) * sizeof(u32)));
        CK(h->path_off.ensure((n_blocks + 2) * sizeof(u64)));
        { KScope ks(h, "k_path_count"); k_path_count<<<grid_for(n_blocks, 1, 16), 256, 0, h->stream>>>(h->d_text, h->path_recs.as<PathRec>(), R, n_blocks, h->path_cnt.as<u32>()); }
        CK(cudaGetLastError());
        LoadArray<u32> lc{h->path_cnt.as<u32>()};
        int rc = launch_scan<u64>(h, lc, h->path_off.as<u64>(), nullptr, n_blocks, nullptr, nullptr);
        if (rc) return rc;
        h->h_path_blkoff.resize(n_blocks + 1);
        CK(cudaMemcpyAsync(h->h_path_blkoff.data(), h->path_off.p, (n_blocks + 1) * sizeof(u64), cudaMemcpyDeviceToHost, h->stream));
        CK(cudaStreamSynchronize(h->stream));
        const u64 total = h->h_path_blkoff[n_blocks];
        if (total >= 0x7FFFFFFF00ull) { h->err = "too many path entries"; return G2N_ERR_UNSUPPORTED; }
        CK(h->path_ids.ensure((total + 1) * sizeof(int32_t)));
        PathLookup T;
        T.slots = h->d_slots; T.longs = h->longs.as<LongDesc>(); T.slot_id = h->slot_id.as<u32>();
        T.mask = h->table_cap - 1; T.seed = h->seed_used; T.bidirected = h->params.bidirected ? 1 : 0;
        { KScope ks(h, "k_path_lookup"); k_path_lookup<<<grid_for(n_blocks, 1, 16), 256, 0, h->stream>>>(h->d_text, h->path_recs.as<PathRec>(), R, n_blocks, h->path_off.as<u64>(), T, h->path_ids.as<int32_t>()); }
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(h->h_paths.data(), h->path_recs.p, (size_t)R * sizeof(PathRec), cudaMemcpyDeviceToHost, h->stream));
        CK(cudaStreamSynchronize(h->stream));
        h->h_path_entry0.resize(R + 1);
        for (u32 r = 0; r < R; r++) h->h_path_entry0[r] = h->h_path_blkoff[h->h_paths[r].first_blk];
        h->h_path_entry0[R] = total;
    }
    h->paths_ready = true;
    *n_paths = R;
    return G2N_OK;
}

int g2n_path_info(g2n_handle* h, uint64_t i, g2n_path_info_t* out)
{
    if (!h || !out || !h->paths_ready || i >= h->h_paths.size()) return G2N_ERR_INVALID;
    CK(cudaSetDevice(h->device));
    const PathRec& R = h->h_paths[i];
    memset(out, 0, sizeof(*out));
    out->line_offset = R.line;
    out->name_offset = R.name_off;
    out->name_len = R.name_len;
    out->n_entries = h->h_path_entry0[i + 1] - h->h_path_entry0[i];
    out->missing_entry = R.missing == ~0ull ? -1 : (int64_t)R.missing;
    if (R.missing != ~0ull) {
        // where that entry's text is: the block that holds the entry index, then at most 4 KiB of commas
        const u64 target = h->h_path_entry0[i] + R.missing;
        const u64 nb = h->h_path_blkoff.size() - 1;
        u64 a = R.first_blk, z = (i + 1 < h->h_paths.size() ? h->h_paths[i + 1].first_blk : nb);
        while (z - a > 1) { const u64 m = (a + z) >> 1; if (h->h_path_blkoff[m] <= target) a = m; else z = m; }
        const u64 lo = R.list_off + (a - R.first_blk) * PB_BYTES;
        const u64 hi = std::min<u64>(lo + PB_BYTES, R.list_end + 1);
        const u64 from = lo ? lo - 1 : 0, to = std::min<u64>(R.list_end, hi + (1u << 16));
        std::vector<uint8_t> buf(to - from + 1);
        if (to > from) CK(cudaMemcpy(buf.data(), h->d_text + from, to - from, cudaMemcpyDeviceToHost));
        u64 k = h->h_path_blkoff[a];
        for (u64 p = lo; p < hi; p++) {
            const bool start = p == R.list_off || buf[p - 1 - from] == ',';
            if (!start) continue;
            if (k == target) {
                u64 e = p;
                while (e < R.list_end && e - from < buf.size() - 1 && buf[e - from] != ',') e++;
                u64 len = e - p;
                if (len && (buf[e - 1 - from] == '+' || buf[e - 1 - from] == '-')) len--;
                out->missing_offset = p;
                out->missing_len = (uint32_t)len;
                break;
            }
            k++;
        }
    }
    return G2N_OK;
}

int g2n_fetch_text(g2n_handle* h, uint64_t offset, uint64_t len, uint8_t* out)
{
    if (!h || (!out && len) || !h->built || offset + len > h->nbytes) return G2N_ERR_INVALID;
    CK(cudaSetDevice(h->device));
    if (len) CK(cudaMemcpyAsync(out, h->d_text + offset, len, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return G2N_OK;
}

int g2n_path_bfs(g2n_handle* h, uint64_t i, int32_t slot, int32_t n_slots)
{
    if (!h || !h->paths_ready || i >= h->h_paths.size()) return G2N_ERR_INVALID;
    CK(cudaSetDevice(h->device));
    const u64 e0 = h->h_path_entry0[i], n = h->h_path_entry0[i + 1] - e0;
    return bfs_run(h, h->path_ids.as<int32_t>() + e0, n, true, slot, n_slots);
}

int g2n_path_reduce(g2n_handle* h, int32_t slot, uint64_t i, int64_t* out3)
{
    if (!h || !h->paths_ready || i >= h->h_paths.size()) return G2N_ERR_INVALID;
    const u64 e0 = h->h_path_entry0[i], n = h->h_path_entry0[i + 1] - e0;
    return levels_reduce_run(h, slot, h->path_ids.as<int32_t>() + e0, n, true, out3);
}

int g2n_fetch_path_nodes(g2n_handle* h, uint64_t i, int32_t* out)
{
    if (!h || !out || !h->paths_ready || i >= h->h_paths.size()) return G2N_ERR_INVALID;
    CK(cudaSetDevice(h->device));
    const u64 e0 = h->h_path_entry0[i], n = h->h_path_entry0[i + 1] - e0;
    if (n) CK(cudaMemcpyAsync(out, h->path_ids.as<int32_t>() + e0, n * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return G2N_OK;
}

int g2n_fetch_levels(g2n_handle* h, int32_t slot, int32_t* out)
{
    if (!h || !out) return G2N_ERR_INVALID;
    CK(cudaSetDevice(h->device));
    int rc = bfs_check(h, slot);
    if (rc) return rc;
    const u64 n = h->n_nodes;
    if (n) CK(cudaMemcpyAsync(out, h->bfs_levels.as<int32_t>() + (size_t)slot * (n + 1), n * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return G2N_OK;
}

int g2n_coo_to_compressed(g2n_handle* h, const int32_t* row, const int32_t* col, const void* data, uint64_t nnz_in, uint64_t n,
                          int32_t dtype, int32_t want_format, int32_t* indptr, int32_t* indices, void* data_out, uint64_t* nnz_out)
{
    if (!h || !indptr || !nnz_out) return G2N_ERR_INVALID;
    if (want_format != G2N_FMT_CSR && want_format != G2N_FMT_CSC) { h->err = "target must be CSR or CSC"; return G2N_ERR_INVALID; }
    if (dtype < G2N_DTYPE_F64 || dtype > G2N_DTYPE_BOOL) { h->err = "unknown dtype"; return G2N_ERR_INVALID; }
    if (n > 0x7FFFFFFFull || nnz_in >= 0xFFFFFFF0ull) { h->err = "int64 indices are out of scope"; return G2N_ERR_UNSUPPORTED; }
    CK(cudaSetDevice(h->device));
    h->built = false;
    h->have_edges = false;
    const size_t ds = dtype_size(dtype);
    if (nnz_in == 0 || n == 0) {
        memset(indptr, 0, (n + 1) * sizeof(int32_t));
        *nnz_out = 0;
        return G2N_OK;
    }
    CK(h->up_row.ensure(nnz_in * sizeof(int32_t)));
    CK(h->up_col.ensure(nnz_in * sizeof(int32_t)));
    CK(h->up_data.ensure(nnz_in * ds));
    CK(cudaMemcpyAsync(h->up_row.p, row, nnz_in * sizeof(int32_t), cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(h->up_col.p, col, nnz_in * sizeof(int32_t), cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(h->up_data.p, data, nnz_in * ds, cudaMemcpyHostToDevice, h->stream));
    CK(h->entries.ensure((nnz_in + 1) * sizeof(u64)));
    {
        size_t zb = 0;
        int rc0 = layout_zearly(h, 1024, 0, false, &zb);  // only the control block is used here
        if (rc0) return rc0;
        rc0 = layout_zrows(h, n);
        if (rc0) return rc0;
    }
    h->spec = false;
    CK(cudaMemsetAsync(h->zrows.p, 0, h->zrows_bytes, h->stream));
    k_set_sizes<<<1, 32, 0, h->stream>>>(h->d_ds, (u32)n, (u32)nnz_in);
    const int csc = want_format == G2N_FMT_CSC ? 1 : 0;
    { KScope ks(h, "k_coo_count"); k_coo_count<<<grid_for(nnz_in, 256), 256, 0, h->stream>>>(h->up_row.as<int32_t>(), h->up_col.as<int32_t>(), nnz_in, csc, h->d_rowcnt); }
    CK(cudaGetLastError());
    int rc = rows_scan(h, n, nullptr);
    if (rc) return rc;
    { KScope ks(h, "k_coo_scatter"); k_coo_scatter<<<grid_for(nnz_in, 256), 256, 0, h->stream>>>(h->up_row.as<int32_t>(), h->up_col.as<int32_t>(), nnz_in, csc, h->cursor.as<u32>(), h->entries.as<u64>()); }
    CK(cudaGetLastError());
    rc = rows_finalize(h, dtype, true, nnz_in, n, 0, WEmit{nullptr, 0u}, h->up_data.p);
    if (rc) return rc;
    CK(cudaMemcpyAsync(&h->h_ctl->s, h->d_ds, sizeof(DevSizes), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    const u64 nnz = h->h_ctl->s.nnz;
    *nnz_out = nnz;
    CK(cudaMemcpyAsync(indptr, h->indptr.p, (n + 1) * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
    if (nnz) {
        CK(cudaMemcpyAsync(indices, h->indices.p, nnz * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
        CK(cudaMemcpyAsync(data_out, h->data.p, nnz * ds, cudaMemcpyDeviceToHost, h->stream));
    }
    CK(cudaStreamSynchronize(h->stream));
    return G2N_OK;
}

// =====================================================================================
// multi-GPU build (device side and protocol: dist.cuh).  One handle per rank.  The caller
// (gfa2network_b200/dist.py) only bootstraps: it agrees on capacities, exchanges the CUDA IPC handles of
// the exchange arenas and queues the six stages; data moves between GPUs inside the kernels.

static size_t up256(size_t x) { return (x + 255) & ~(size_t)255; }

static void dx_make_layout(DxLayout& L, int world, u64 kcap, u64 pcap, u64 pair_bytes)
{
    L.kcap = kcap;
    L.pcap = pcap;
    L.pair_bytes = pair_bytes;
    size_t o = 0;
    L.off_key = o;   o += up256((size_t)world * kcap * sizeof(TKey));
    L.off_ord = o;   o += up256((size_t)world * kcap * sizeof(u32));
    L.off_first = o; o += up256((size_t)world * kcap);
    L.off_rank = o;  o += up256((size_t)world * kcap * sizeof(u32));
    L.off_id = o;    o += up256((size_t)world * kcap * sizeof(u32));
    L.off_pair = o;  o += up256((size_t)world * pcap * pair_bytes);
    L.bytes = o;
}

int g2n_dist_init(g2n_handle* h, int rank, int world)
{
    if (!h || world < 1 || world > DX_MAXW || rank < 0 || rank >= world) return G2N_ERR_INVALID;
    CK(cudaSetDevice(h->device));
    if (!h->dx_ctl.p) {
        CK(h->dx_ctl.ensure(sizeof(DxCtl)));
        CK(cudaMemset(h->dx_ctl.p, 0, sizeof(DxCtl)));
        CK(h->dx_loc.ensure(sizeof(DxLocal)));
        CK(cudaHostAlloc((void**)&h->h_loc, sizeof(DxLocal), cudaHostAllocDefault));
    }
    memset(&h->dxp, 0, sizeof(h->dxp));
    memset(&h->dxl, 0, sizeof(h->dxl));
    h->dxp.rank = rank;
    h->dxp.world = world;
    h->dxp.ctl[rank] = h->dx_ctl.as<DxCtl>();
    h->dx_inited = true;
    h->dx_probed = false;
    return G2N_OK;
}

// Ranks that share a process but sit on different GPUs reach each other's arenas through plain device pointers:
// peer access from this handle's device to `peer_device` has to be on (a no-op for the same device).
int g2n_dist_enable_peer(g2n_handle* h, int peer_device)
{
    if (!h) return G2N_ERR_INVALID;
    CK(cudaSetDevice(h->device));
    if (peer_device == h->device) return G2N_OK;
    int can = 0;
    CK(cudaDeviceCanAccessPeer(&can, h->device, peer_device));
    if (!can) { h->err = "no peer access between the two devices"; return G2N_ERR_UNSUPPORTED; }
    const cudaError_t e = cudaDeviceEnablePeerAccess(peer_device, 0);
    if (e == cudaErrorPeerAccessAlreadyEnabled) { (void)cudaGetLastError(); return G2N_OK; }
    CK(e);
    return G2N_OK;
}

// This rank's byte range of a file -> the handle's text buffer on the device (pread into two pinned staging buffers,
// the copy of one piece overlaps the read of the next).  *dev_text is 16-byte aligned and stays valid until the handle
// loads another text; pass it to g2n_dist_probe / g2n_dist_stage with text_on_device = 1.
int g2n_load_file_range(g2n_handle* h, const char* path, uint64_t offset, uint64_t nbytes, void** dev_text)
{
    if (!h || !path || !dev_text) return G2N_ERR_INVALID;
    CK(cudaSetDevice(h->device));
    const int fd = open(path, O_RDONLY);
    if (fd < 0) { h->err = std::string("cannot open ") + path; return G2N_ERR_INVALID; }
    int rc = G2N_OK;
    do {
        if (cudaStreamSynchronize(h->stream) != cudaSuccess) { rc = G2N_ERR_CUDA; break; }
        if (h->text.ensure(nbytes + 64) != cudaSuccess) { h->err = "out of device memory for the text"; rc = G2N_ERR_CUDA; break; }
        const u64 piece = G2N_H2D_PIECE;
        while (h->stage.size() < 2) {
            void* b = nullptr;
            if (cudaHostAlloc(&b, piece, cudaHostAllocDefault) != cudaSuccess) { rc = G2N_ERR_CUDA; break; }
            h->stage.push_back(b);
            cudaEvent_t e;
            if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) { rc = G2N_ERR_CUDA; break; }
            h->stage_ev.push_back(e);
        }
        if (rc) break;
        u32 k = 0;
        for (u64 a = 0; a < nbytes && !rc; a += piece, k++) {
            const u32 sb = k & 1;
            cudaEventSynchronize(h->stage_ev[sb]);  // the previous copy out of this staging buffer is done
            const u64 len = a + piece < nbytes ? piece : nbytes - a;
            u64 got = 0;
            while (got < len) {
                const ssize_t n = pread(fd, (uint8_t*)h->stage[sb] + got, len - got, (off_t)(offset + a + got));
                if (n <= 0) { h->err = "reading the input file failed"; rc = G2N_ERR_INVALID; break; }
                got += (u64)n;
            }
            if (rc) break;
            if (cudaMemcpyAsync(h->text.as<uint8_t>() + a, h->stage[sb], len, cudaMemcpyHostToDevice, h->copy_stream) != cudaSuccess ||
                cudaEventRecord(h->stage_ev[sb], h->copy_stream) != cudaSuccess) { h->err = "queueing a copy of the input file failed"; rc = G2N_ERR_CUDA; }
        }
        if (!rc && cudaStreamSynchronize(h->copy_stream) != cudaSuccess) rc = G2N_ERR_CUDA;
    } while (0);
    close(fd);
    if (rc) return rc;
    h->built = false;
    *dev_text = h->text.p;
    return G2N_OK;
}

int g2n_dist_plan(g2n_handle* h, uint64_t key_cap, uint64_t pair_cap, uint64_t rows_cap, uint64_t recv_cap, int dry_run, int* will_realloc)
{
    if (!h || !h->dx_inited) return G2N_ERR_INVALID;
    CK(cudaSetDevice(h->device));
    const u64 kcap = (key_cap + 255) & ~255ull, pcap = (pair_cap + 255) & ~255ull;
    if (kcap >= (1ull << 29) || pcap >= 0xFFFFFF00ull) { h->err = "multi-GPU exchange segment too large"; return G2N_ERR_UNSUPPORTED; }
    DxLayout L;
    // the mode of the last g2n_dist_probe / stage 0 decides the entry format: 16 bytes with a weight, 8 without
    dx_make_layout(L, h->dxp.world, kcap, pcap, h->params.weight_tag_len > 0 ? sizeof(DistPairW) : sizeof(DistPair));
    const bool re = L.bytes > h->dx_arena.cap;
    if (will_realloc) *will_realloc = re ? 1 : 0;
    if (dry_run) return G2N_OK;
    if (re) {
        for (int s = 0; s < h->dxp.world; s++)
            if (h->dx_peer_open[s]) { h->err = "close the peer mappings before the exchange arena grows"; return G2N_ERR_INVALID; }
        CK(cudaStreamSynchronize(h->stream));
        CK(h->dx_arena.ensure(L.bytes));
    }
    h->dxl = L;
    h->dxp.arena[h->dxp.rank] = h->dx_arena.as<uint8_t>();
    h->dx_rows_cap = rows_cap;
    h->dx_recv_cap = recv_cap;
    const u64 want = (u64)h->dxp.world * kcap;
    h->dx_gcap = next_pow2(want + want / 2 < 1024 ? 1024 : want + want / 2);
    CK(h->dx_zg.ensure((size_t)h->dx_gcap * sizeof(Slot)));
    CK(h->dx_gslot.ensure((size_t)want * sizeof(u32) + 256));
    CK(h->dx_gpos.ensure((size_t)h->dx_gcap * sizeof(u32)));
    return G2N_OK;
}

int g2n_dist_local_mem(g2n_handle* h, void** arena, void** ctl, uint8_t* ipc128)
{
    if (!h || !h->dx_inited || !h->dx_arena.p) return G2N_ERR_INVALID;
    CK(cudaSetDevice(h->device));
    if (arena) *arena = h->dx_arena.p;
    if (ctl) *ctl = h->dx_ctl.p;
    if (ipc128) {
        cudaIpcMemHandle_t a, c;
        CK(cudaIpcGetMemHandle(&a, h->dx_arena.p));
        CK(cudaIpcGetMemHandle(&c, h->dx_ctl.p));
        static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle size");
        memcpy(ipc128, &a, 64);
        memcpy(ipc128 + 64, &c, 64);
    }
    return G2N_OK;
}

int g2n_dist_set_peers(g2n_handle* h, void* const* arenas, void* const* ctls)
{
    if (!h || !h->dx_inited || !arenas || !ctls) return G2N_ERR_INVALID;
    for (int s = 0; s < h->dxp.world; s++) {
        if (s == h->dxp.rank) continue;
        h->dxp.arena[s] = (uint8_t*)arenas[s];
        h->dxp.ctl[s] = (DxCtl*)ctls[s];
    }
    return G2N_OK;
}

int g2n_dist_close_peers(g2n_handle* h)
{
    if (!h || !h->dx_inited) return G2N_ERR_INVALID;
    CK(cudaSetDevice(h->device));
    CK(cudaStreamSynchronize(h->stream));
    for (int s = 0; s < h->dxp.world; s++) {
        if (!h->dx_peer_open[s]) continue;
        cudaIpcCloseMemHandle(h->dxp.arena[s]);
        cudaIpcCloseMemHandle(h->dxp.ctl[s]);
        h->dxp.arena[s] = nullptr;
        h->dxp.ctl[s] = nullptr;
        h->dx_peer_open[s] = false;
    }
    return G2N_OK;
}

int g2n_dist_open_peers(g2n_handle* h, const uint8_t* ipc_all)
{
    if (!h || !h->dx_inited || !ipc_all) return G2N_ERR_INVALID;
    int rc = g2n_dist_close_peers(h);
    if (rc) return rc;
    for (int s = 0; s < h->dxp.world; s++) {
        if (s == h->dxp.rank) continue;
        cudaIpcMemHandle_t a, c;
        memcpy(&a, ipc_all + (size_t)s * 128, 64);
        memcpy(&c, ipc_all + (size_t)s * 128 + 64, 64);
        void *pa = nullptr, *pc = nullptr;
        CK(cudaIpcOpenMemHandle(&pa, a, cudaIpcMemLazyEnablePeerAccess));
        CK(cudaIpcOpenMemHandle(&pc, c, cudaIpcMemLazyEnablePeerAccess));
        h->dxp.arena[s] = (uint8_t*)pa;
        h->dxp.ctl[s] = (DxCtl*)pc;
        h->dx_peer_open[s] = true;
    }
    return G2N_OK;
}

// Host-planned pass: tokenize this rank's byte range with the usual host round trip and report the shard's
// counts (or its first parse error -- g2n_status -- so that every rank can raise the same exception).
int g2n_dist_probe(g2n_handle* h, const uint8_t* text, uint64_t nbytes, const g2n_params* p, g2n_dist_info* out)
{
    if (!h || !out || !h->dx_inited) return G2N_ERR_INVALID;
    memset(out, 0, sizeof(*out));
    h->dx_probed = false;
    int rc = tokenize_phase(h, text, nbytes, p, false, false);
    if (rc) return rc;
    // names beyond the 15-byte inline key travel as their 128-bit tagged hash (table.cuh: make_key); every rank must
    // hash with the same seed, so a shard that had to re-seed (a collision among ITS long names) cannot take part
    if (h->reseeded) { h->err = "multi-GPU build: hash collision among long node names (re-seeded table)"; return G2N_ERR_UNSUPPORTED; }
    out->n_keys = h->n_nodes;
    out->n_tiles = h->n_tiles;
    out->n_records = h->n_records;
    out->n_edge_records = h->n_edges;
    out->n_entries = h->n_edges * (u64)h->tpe * (h->symmax ? 2 : 1);
    h->hint_sig = shape_signature(p, nbytes, h->params.weight_tag_len);
    h->hint_valid = true;
    h->dx_probed = true;
    return G2N_OK;
}

int g2n_dist_stage(g2n_handle* h, int stage, const uint8_t* text, uint64_t nbytes, const g2n_params* p, int speculative)
{
    if (!h || !h->dx_inited || stage < 0 || stage > 6) return G2N_ERR_INVALID;
    CK(cudaSetDevice(h->device));
    const int W = h->dxp.world;
    for (int s = 0; s < W; s++)
        if (!h->dxp.arena[s] || !h->dxp.ctl[s]) { h->err = "multi-GPU peers are not connected"; return G2N_ERR_INVALID; }
    if (stage == 0) {
        if (!p) return G2N_ERR_INVALID;
        h->built = false;
        h->dx_spec = speculative != 0;
        if (h->dx_spec) {
            if (!h->hint_valid) { h->err = "speculative multi-GPU build without a previous build on this handle"; return G2N_ERR_INVALID; }
            int rc = tokenize_phase(h, text, nbytes, p, true, false);
            if (rc) return rc;
        } else if (!h->dx_probed) {
            h->err = "g2n_dist_probe must precede a host-planned multi-GPU build";
            return G2N_ERR_INVALID;
        }
        h->dx_probed = false;
        if ((h->params.weight_tag_len > 0) != (h->dxl.pair_bytes == sizeof(DistPairW))) {
            h->err = "the multi-GPU plan was made for a different mode (weight tag)";
            return G2N_ERR_INVALID;
        }
        h->dxp.epoch++;
        if (!h->dx_spec) {  // a repeated host-planned pass over the same probe: the bitmap must start clean
            CK(cudaMemsetAsync(h->zids.p, 0, h->zids_bytes, h->stream));
        }
        CK(cudaMemsetAsync(h->dx_loc.p, 0, sizeof(DxLocal), h->stream));
        CK(cudaMemsetAsync(h->dx_zg.p, 0, (size_t)h->dx_gcap * sizeof(Slot), h->stream));
        const u64 words = (4 * h->cap_R + 31) / 32 + 1;
        CK(h->wprefix.ensure((words + 2) * sizeof(u32)));
        CK(h->slot_id.ensure((size_t)h->table_cap * sizeof(u32)));
        CK(h->dx_sent.ensure((size_t)W * h->dxl.kcap * sizeof(u64) + 256));  // klist: [owner][position] -> order bit | slot
        CK(h->id2slot.ensure((h->cap_n + 1) * sizeof(u32)));
        CK(h->name_off.ensure((h->cap_n + 2) * sizeof(u64)));
        h->slab_mode = true;
        h->names_sized = false;
        h->names_ready = false;
        h->tsv_ready = false;
    }
    const DxPeers X = h->dxp;
    const DxLayout L = h->dxl;
    DxLocal* loc = h->dx_loc.as<DxLocal>();
    const DxCtl* my = h->dx_ctl.as<DxCtl>();
    const u32 cap = h->table_cap;
    DxOwner G;
    G.gslots = h->dx_zg.as<Slot>();
    G.gmask = h->dx_gcap - 1;
    G.gslot = h->dx_gslot.as<u32>();
    G.gpos = h->dx_gpos.as<u32>();
    const u32 kgrid = grid_for((u64)W * L.kcap, 256, 8);  // the kernels walk all per-peer segments as one index space (DxFlat)
    const int sym = h->symmax ? 1 : 0;
    const int csc = (!sym && h->params.want_format == G2N_FMT_CSC) ? 1 : 0;
    switch (stage) {
    case 0: {
        { KScope ks(h, "k_dx_export"); k_dx_export<<<grid_for(cap, DXK_SLOTS), 256, 0, h->stream>>>(h->d_slots, cap, h->tile_base.as<u64>(), h->d_ds, X, L, loc, h->dx_sent.as<u64>()); }
        break;
    }
    case 1: {
        { KScope ks(h, "k_dx_insert"); k_dx_insert<<<kgrid, 256, 0, h->stream>>>(X, L, my, loc, G, h->d_cnt); }
        { KScope ks(h, "k_dx_reply_first"); k_dx_reply_first<<<grid_for((u64)W * L.kcap / 4 + 1, 256, 8), 256, 0, h->stream>>>(X, L, my, loc, G); }
        break;
    }
    case 2: {
        { KScope ks(h, "k_dx_mark"); k_dx_mark<<<kgrid, 256, 0, h->stream>>>(h->d_ds, X, L, my, loc, h->dx_sent.as<u64>(), h->d_bitmap); }
        const u64 words = (4 * h->cap_R + 31) / 32 + 1;
        LoadPopc8 lp{h->d_bitmap};
        int rc = launch_scan<u32>(h, lp, h->wprefix.as<u32>(), nullptr, words / BM_GROUP + 1, &h->d_ds->wgroups, h->d_scan_words);
        if (rc) return rc;
        { KScope ks(h, "k_dx_send_rank"); k_dx_send_rank<<<kgrid, 256, 0, h->stream>>>(h->d_slots, h->d_ds, X, L, h->dx_sent.as<u64>(), h->d_bitmap, h->wprefix.as<u32>(), h->id2slot.as<u32>(), loc); }
        break;
    }
    case 3: {
        { KScope ks(h, "k_dx_reply_ids"); k_dx_reply_ids<<<kgrid, 256, 0, h->stream>>>(X, L, my, loc, G); }
        break;
    }
    case 4: {
        const u32 rows_cap = h->dx_spec ? (u32)h->dx_rows_cap : 0xFFFFFFFFu;
        { KScope ks(h, "k_dx_localmap"); k_dx_localmap<<<kgrid, 256, 0, h->stream>>>(h->d_ds, X, L, my, loc, h->dx_sent.as<u64>(), h->slot_id.as<u32>(), rows_cap); }
        if (h->params.weight_tag_len > 0) {
            // weighted: positions follow the emission order (count per tile and owner, scan, ordered scatter)
            const u64 cells = (u64)W * h->n_tiles;
            CK(h->dx_tcnt.ensure((cells + 2) * sizeof(u32)));
            CK(h->dx_toff.ensure((cells + 2) * sizeof(u32)));
            EmitParams EP = emit_params(h);
            const u32 tgrid = grid_for((u64)h->n_tiles * 32 + 1, 256);
            { KScope ks(h, "k_dxw_count"); k_dxw_count<<<tgrid, 256, 0, h->stream>>>(EP, sym, csc, X, L, loc, h->dx_tcnt.as<u32>()); }
            CK(cudaGetLastError());
            LoadArray<u32> lt{h->dx_tcnt.as<u32>()};
            int rc = launch_scan<u32>(h, lt, h->dx_toff.as<u32>(), nullptr, cells, nullptr, nullptr);
            if (rc) return rc;
            { KScope ks(h, "k_dxw_scatter"); k_dxw_scatter<<<tgrid, 256, 0, h->stream>>>(EP, sym, csc, X, L, loc, h->dx_toff.as<u32>()); }
            break;
        }
        const u32 egrid = grid_for((h->cap_E + 256 * DXE_BATCH - 1) / (256 * DXE_BATCH), 1, 8);
        u32* es = h->edge_slots.as<u32>();
        {
            KScope ks(h, "k_dx_entries");
            switch (h->tpe) {
                case 1: k_dx_entries<1><<<egrid, 256, 0, h->stream>>>(es, h->slot_id.as<u32>(), h->d_ds, sym, csc, X, L, loc); break;
                case 2: k_dx_entries<2><<<egrid, 256, 0, h->stream>>>(es, h->slot_id.as<u32>(), h->d_ds, sym, csc, X, L, loc); break;
                default: k_dx_entries<4><<<egrid, 256, 0, h->stream>>>(es, h->slot_id.as<u32>(), h->d_ds, sym, csc, X, L, loc); break;
            }
        }
        h->edges_are_ids = true;
        break;
    }
    case 5: {
        u64 rows_cap = h->dx_rows_cap, recv_cap = h->dx_recv_cap;
        { KScope ks(h, "k_dx_slab_sizes"); k_dx_slab_sizes<<<1, 32, 0, h->stream>>>(X, L, my, loc, h->d_ds, h->dx_spec ? (u32)recv_cap : 0xFFFFFFF0u, h->dx_spec ? (u32)rows_cap : 0xFFFFFFFFu); }
        if (!h->dx_spec) {
            // host-planned pass: size the slab buffers exactly (one extra round trip)
            CK(cudaMemcpyAsync(h->h_loc, loc, sizeof(DxLocal), cudaMemcpyDeviceToHost, h->stream));
            CK(cudaStreamSynchronize(h->stream));
            rows_cap = h->h_loc->n_rows;
            recv_cap = h->h_loc->n_recv;
            h->dx_rows_cap = rows_cap;
            h->dx_recv_cap = recv_cap;
        }
        if (recv_cap >= 0xFFFFFFF0ull) { h->err = "more than 2^32 entries in one slab"; return G2N_ERR_UNSUPPORTED; }
        const bool weighted = h->params.weight_tag_len > 0;
        CK(h->entries.ensure((recv_cap + 1) * (weighted ? sizeof(u64) : sizeof(u32))));
        if (weighted) CK(h->w_emit.ensure((recv_cap + 1) * sizeof(double)));
        {
            int rc0 = layout_zrows(h, rows_cap);
            if (rc0) return rc0;
        }
        h->spec = false;
        h->rows_presorted = false;
        CK(cudaMemsetAsync(h->zrows.p, 0, h->zrows_bytes, h->stream));
        h->result_format = csc ? G2N_FMT_CSC : G2N_FMT_CSR;
        const u32 pgrid = grid_for(recv_cap + 1, 256, 8);
        const RowPasses rp = row_passes(recv_cap, weighted ? sizeof(u64) : sizeof(u32), rows_cap);
        if (rp.bucketed) {
            // slab far larger than L2: partition the received entries by row bucket, then stream (dist.cuh / rowsort.cuh)
            const BucketPlan BP = plan_buckets(rp, recv_cap, rows_cap, weighted ? sizeof(u64) : sizeof(u32));
            const RowBuckets rb = BP.rb;
            CK(h->pair_major.ensure((recv_cap + 1) * sizeof(u32)));
            CK(h->pair_ent.ensure((recv_cap + 1) * (weighted ? sizeof(u64) : sizeof(u32))));
            CK(h->bucket_ctl.ensure(sizeof(BucketCtl)));
            CK(cudaMemsetAsync(h->bucket_ctl.p, 0, sizeof(BucketCtl), h->stream));
            BucketCtl* ctl = h->bucket_ctl.as<BucketCtl>();
            const u32 sgrid = grid_for((recv_cap + RB_ROUND) / RB_ROUND, 1, 8);
            if (weighted) {
                { KScope ks(h, "k_pairs_bucket_count"); k_pairs_bucket_count<DistPairW><<<pgrid, 256, 0, h->stream>>>(X, L, loc, h->d_ds, rb, ctl, &loc->bad); }
                { KScope ks(h, "k_pairs_bucket_scatter"); k_pairs_bucket_scatter<DistPairW, Ent64><<<sgrid, 256, 0, h->stream>>>(X, L, loc, h->d_ds, rb, ctl, h->pair_major.as<u32>(), h->pair_ent.as<u64>(), h->w_emit.as<double>()); }
            } else {
                { KScope ks(h, "k_pairs_bucket_count"); k_pairs_bucket_count<DistPair><<<pgrid, 256, 0, h->stream>>>(X, L, loc, h->d_ds, rb, ctl, &loc->bad); }
                { KScope ks(h, "k_pairs_bucket_scatter"); k_pairs_bucket_scatter<DistPair, Ent32><<<sgrid, 256, 0, h->stream>>>(X, L, loc, h->d_ds, rb, ctl, h->pair_major.as<u32>(), h->pair_ent.as<u32>(), nullptr); }
            }
            CK(cudaGetLastError());
            int rc = weighted ? bucketed_tail<Ent64>(h, BP, recv_cap, rows_cap, false, sym) : bucketed_tail<Ent32>(h, BP, recv_cap, rows_cap, false, sym);
            if (rc) return rc;
            rc = rows_finalize(h, h->params.dtype, weighted, recv_cap, rows_cap, sym, WEmit{weighted ? h->w_emit.as<double>() : nullptr, 0u}, nullptr);
            if (rc) return rc;
            break;
        }
        {
            const RowRange rr = ROW_RANGE_ALL;
            if (weighted) { KScope ks(h, "k_pairsw_count"); k_pairsw_count<<<pgrid, 256, 0, h->stream>>>(X, L, loc, h->d_ds, h->d_rowcnt, &loc->bad, rr); }
            else { KScope ks(h, "k_pairs_count"); k_pairs_count<<<pgrid, 256, 0, h->stream>>>(X, L, loc, h->d_ds, h->d_rowcnt, &loc->bad, rr); }
        }
        int rc = rows_scan(h, rows_cap, &h->d_ds->rows);
        if (rc) return rc;
        for (u32 ps = 0; ps < rp.count; ps++) {
            const RowRange rr = rp.at(ps, rows_cap);
            if (weighted) { KScope ks(h, "k_pairsw_scatter"); k_pairsw_scatter<<<pgrid, 256, 0, h->stream>>>(X, L, loc, h->d_ds, h->cursor.as<u32>(), h->entries.as<u64>(), h->w_emit.as<double>(), rr); }
            else { KScope ks(h, "k_pairs_scatter"); k_pairs_scatter<<<pgrid, 256, 0, h->stream>>>(X, L, loc, h->d_ds, h->cursor.as<u32>(), h->entries.as<u32>(), rr); }
        }
        CK(cudaGetLastError());
        rc = rows_finalize(h, h->params.dtype, weighted, recv_cap, rows_cap, sym, WEmit{weighted ? h->w_emit.as<double>() : nullptr, 0u}, nullptr);
        if (rc) return rc;
        break;
    }
    case 6: {
        { KScope ks(h, "k_dx_final"); k_dx_final<<<1, 32, 0, h->stream>>>(X, my, loc); }
        CK(cudaEventRecord(h->ev[EV_REDUCE], h->stream));
        CK(cudaMemcpyAsync(h->h_loc, loc, sizeof(DxLocal), cudaMemcpyDeviceToHost, h->stream));
        CK(cudaMemcpyAsync(h->h_ctl, h->d_ctl, sizeof(Ctl), cudaMemcpyDeviceToHost, h->stream));
        break;
    }
    }
    CK(cudaGetLastError());
    return G2N_OK;
}

// The one host round trip of a multi-GPU build.  G2N_ERR_RETRY: some rank exceeded a capacity (the same
// verdict on every rank): repeat with g2n_dist_probe + a host-planned pass.
int g2n_dist_finish(g2n_handle* h, g2n_dist_result* out)
{
    if (!h || !out || !h->dx_inited) return G2N_ERR_INVALID;
    CK(cudaSetDevice(h->device));
    CK(cudaStreamSynchronize(h->stream));
    const DxLocal& l = *h->h_loc;
    memset(out, 0, sizeof(*out));
    out->bad = l.final_bad;
    out->n_global = l.n_global;
    out->row0 = l.row0;
    out->n_rows = l.n_rows;
    out->n_recv = l.n_recv;
    out->n_first = l.n_first;
    out->id0 = l.idbase[h->dxp.rank];
    out->n_keys = l.n_keys;
    out->n_records = l.n_records;
    out->n_edge_records = l.n_edges;
    for (int d = 0; d < h->dxp.world; d++) { out->keys_to[d] = l.cur_keys[d]; out->pairs_to[d] = l.cur_pairs[d]; }
    h->diag.gpu_launches = h->launches;
    if (l.final_bad) {
        char b[96];
        snprintf(b, sizeof b, "multi-GPU build has to be repeated (status bits 0x%x)", l.final_bad);
        h->err = b;
        return (l.final_bad & DXB_TIMEOUT) ? G2N_ERR_INTERNAL : G2N_ERR_RETRY;
    }
    h->n_global = l.n_global;
    h->n_nodes = l.n_global;
    h->slab_rows = l.n_rows;
    h->names_n = l.n_first;
    h->names_id0 = l.idbase[h->dxp.rank];
    h->n_edges = l.n_edges;
    h->n_records = l.n_records;
    h->nnz = h->h_ctl->s.nnz;
    out->nnz = h->nnz;
    h->diag.n_records = l.n_records;
    h->diag.n_edge_records = l.n_edges;
    h->diag.n_triplets = (u64)l.n_edges * (u64)h->tpe;
    h->diag.speculative = h->dx_spec ? 1 : 0;
    float ms = 0;
    if (cudaEventElapsedTime(&ms, h->ev[EV_START], h->ev[EV_REDUCE]) == cudaSuccess) h->diag.ms_total = ms;
    h->have_edges = false;  // g2n_convert does not apply to a slab
    h->built = true;
    return G2N_OK;
}

}  // extern "C"
