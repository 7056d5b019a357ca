// rowsort.cuh -- K4: COO triplets -> indptr / indices / data by row bucketing.
//
// Adjacency rows of sequence graphs are tiny (a handful of links per segment), so instead of a global
// radix sort of (row, col) keys the build is a counting sort by row followed by an in-row sort:
//   k_rows_count     one atomic per triplet into cnt[major]                       (histogram)
//   exclusive scan   cnt -> rowptr + cursors; rows longer than RS_SMALL are listed on the way (common.cuh)
//   k_rows_scatter_flat / k_edges_scatter_flat   entry (minor, dir[, emission index]) -> atomicAdd(cursor[major]) (any order inside a
//                    row) -- while the row arrays fit L2 (<= 96 MB).  Beyond that the entries are PARTITIONED first: by row
//                    bucket, then by sub-bucket, and the histogram / placement / short-row sort of a sub-bucket happen in
//                    shared memory (k_bucket_* / k_sub_* below); the older scheme, one pass of the flat kernels per row
//                    range (RowRange), is kept behind G2N_DBG_NOBUCKET
//   k_rows_big       rows longer than RS_SMALL are sorted in place by a whole CTA (bitonic; rare)
//   k_rows_sort      one CTA per chunk of RF_ROWS consecutive rows, one lane per row: the chunk's entries
//                    are staged in shared memory (coalesced), every row is sorted -- rows of <= 16
//                    entries by a sorting network in registers -- and the number of entries each row
//                    will store (duplicates summed, zeros of max() dropped) is counted
//   exclusive scan   counts -> indptr
//   k_rows_write     same chunking: walk the sorted rows (duplicates summed left to right, max(S, S^T)),
//                    write indices / data
//   (a single-pass variant with a decoupled look-back over chunk totals was measured and dropped: with
//   ~1.5 K entries per chunk the look-back latency, not the work, set the pace -- 119 us vs 90 us on C2)
// An entry is 64 bits (minor << 33 | dir << 32 | emission index) when weights exist -- the emission
// index finds the weight and makes the order inside a row total, so duplicate weights are summed in
// emission order exactly like SciPy does for rows of <= 16 stored entries -- and 32
// bits (minor << 1 | dir) for unweighted builds, where every weight is 1 and only counts matter.
//
// Replaces the SciPy C++ the reference reaches through
//   builders.py:283   out_mat.maximum(out_mat.T)  -> coo_tocsr, csr_sort_indices, csr_sum_duplicates, csr_maximum_csr
//   utils.py:55       A.asformat("csr" | "csc")   -> coo_tocsr + sum_duplicates
#pragma once
#include "ids.cuh"

namespace g2n {

#define RS_SMALL 64      // rows up to this many entries are sorted by one lane
#define RS_GROUP_CAP 1024  // entries of a 32-row group staged in shared memory per warp
#define RS_WARPS 4

// entry = minor << 33 | dir << 32 | emission index of the triplet
__device__ __forceinline__ u64 rs_entry(u32 minor, u32 dir, u32 t) { return ((u64)minor << 33) | ((u64)dir << 32) | t; }
__device__ __forceinline__ u32 rs_minor(u64 e) { return (u32)(e >> 33); }
__device__ __forceinline__ u32 rs_dir(u64 e) { return (u32)(e >> 32) & 1u; }
__device__ __forceinline__ u32 rs_t(u64 e) { return (u32)e; }

struct Ent64 {  // weighted builds
    typedef u64 type;
    static constexpr bool kWeighted = true;
    static __device__ __forceinline__ u64 make(u32 minor, u32 dir, u32 t) { return rs_entry(minor, dir, t); }
    static __device__ __forceinline__ u32 minor(u64 e) { return rs_minor(e); }
    static __device__ __forceinline__ u32 dir(u64 e) { return rs_dir(e); }
    static __device__ __forceinline__ u32 t(u64 e) { return rs_t(e); }
};
struct Ent32 {  // unweighted builds: every weight is 1.0, node IDs are below 2^31
    typedef u32 type;
    static constexpr bool kWeighted = false;
    static __device__ __forceinline__ u32 make(u32 minor, u32 dir, u32) { return (minor << 1) | dir; }
    static __device__ __forceinline__ u32 minor(u32 e) { return e >> 1; }
    static __device__ __forceinline__ u32 dir(u32 e) { return e & 1u; }
    static __device__ __forceinline__ u32 t(u32) { return 0; }
};

// Row range of one bucketing pass.  When the row histogram, the cursors and the entries together are far
// larger than L2, the count / scatter kernels run once per range of majors (every pass re-reads the edge
// records, a sequential stream, and touches only its slice of the random-access arrays, which then stays
// L2-resident) instead of spraying partial sectors over hundreds of megabytes of DRAM.
struct RowRange {
    u32 lo, n;  // majors [lo, lo + n)
    __device__ __forceinline__ bool has(u32 major) const { return major - lo < n; }
};

// Entries of one edge record:  g(major, minor, dir, t)
//   sym == 0: one entry per triplet, major = row (CSR) or col (CSC)
//   sym == 1: two entries per triplet: (row, col, dir 0) and (col, row, dir 1)   [max(S, S^T)]
template <class G>
__device__ __forceinline__ void record_entries(const u32 (&id)[4], int tpe, u32 t0, int sym, int csc, G g)
{
    for (int k = 0; k < tpe; k++) {
        u32 r, c;
        record_triplet(id, k, r, c);
        if (sym) { g(r, c, 0u, t0 + k); g(c, r, 1u, t0 + k); }
        else if (csc) g(c, r, 0u, t0 + k);
        else g(r, c, 0u, t0 + k);
    }
}

// compile-time triplet count: the record's IDs stay in registers (no dynamic indexing)
template <int TPE, class G>
__device__ __forceinline__ void record_entries_t(const u32 (&id)[4], u32 t0, int sym, int csc, G g)
{
#pragma unroll
    for (int k = 0; k < TPE; k++) {
        const u32 a = id[k & 2], b = id[(k & 2) + 1];
        const u32 r = (k & 1) ? b : a, c = (k & 1) ? a : b;
        if (sym) { g(r, c, 0u, t0 + k); g(c, r, 1u, t0 + k); }
        else if (csc) g(c, r, 0u, t0 + k);
        else g(r, c, 0u, t0 + k);
    }
}

// ---------------------------------------------------------------- unweighted builds: flat passes
// Without weights nothing depends on the emission index, so the two bucketing passes run over the
// stored edge records [0, E) directly -- coalesced, no per-tile bookkeeping.  TPE = triplets per edge
// record (1 | 2 | 4); records hold 4 slots iff TPE == 4.
#define EF_BATCH 4
template <int TPE>
__global__ void __launch_bounds__(256) k_edges_count_flat(u32* __restrict__ edge_slots, const u32* __restrict__ slot_id,
                                                           const DevSizes* __restrict__ ds, int sym, int csc, u32* __restrict__ cnt,
                                                           const RowRange rr, int translate)
{
    constexpr int SPE = TPE == 4 ? 4 : 2;
    if (!ds->ok) return;
    const u32 E = ds->E;
    const u32 stride = gridDim.x * blockDim.x;
    for (u32 e0 = blockIdx.x * blockDim.x + threadIdx.x; e0 < E; e0 += stride * EF_BATCH) {
        u32 id[EF_BATCH][4];
#pragma unroll
        for (int u = 0; u < EF_BATCH; u++) {
            const u32 e = e0 + u * stride;
            id[u][0] = id[u][1] = id[u][2] = id[u][3] = 0;
            if (e < E) {
                if (SPE == 4) {
                    const uint4 q = reinterpret_cast<const uint4*>(edge_slots)[e];
                    id[u][0] = q.x; id[u][1] = q.y; id[u][2] = q.z; id[u][3] = q.w;
                } else {
                    const uint2 q = reinterpret_cast<const uint2*>(edge_slots)[e];
                    id[u][0] = q.x; id[u][1] = q.y;
                }
            }
        }
        if (translate) {  // first pass: table slots -> node IDs, left in place for the later passes
#pragma unroll
            for (int u = 0; u < EF_BATCH; u++) {
                if (e0 + u * stride < E) {
#pragma unroll
                    for (int k = 0; k < SPE; k++) id[u][k] = slot_id[id[u][k]];
                }
            }
        }
#pragma unroll
        for (int u = 0; u < EF_BATCH; u++) {
            const u32 e = e0 + u * stride;
            if (e < E) {
                if (translate) {
                    if (SPE == 4) reinterpret_cast<uint4*>(edge_slots)[e] = make_uint4(id[u][0], id[u][1], id[u][2], id[u][3]);
                    else reinterpret_cast<uint2*>(edge_slots)[e] = make_uint2(id[u][0], id[u][1]);
                }
                record_entries_t<TPE>(id[u], 0u, sym, csc, [&](u32 major, u32, u32, u32) { if (rr.has(major)) atomicAdd(&cnt[major], 1u); });
            }
        }
    }
}

// `slot_id` non-NULL: the records still hold table slots (the tokenizer counted the rows, so no count pass has
// translated them): translate here and leave the node IDs in place for later passes / converts.
template <int TPE>
__global__ void __launch_bounds__(256) k_edges_scatter_flat(u32* __restrict__ edge_ids, const u32* __restrict__ slot_id, const DevSizes* __restrict__ ds, int sym, int csc,
                                                             u32* __restrict__ cursor, u32* __restrict__ entries, const RowRange rr)
{
    constexpr int SPE = TPE == 4 ? 4 : 2;
    if (!ds->ok) return;
    const u32 E = ds->E;
    const u32 stride = gridDim.x * blockDim.x;
    for (u32 e0 = blockIdx.x * blockDim.x + threadIdx.x; e0 < E; e0 += stride * EF_BATCH) {
        u32 id[EF_BATCH][4];
#pragma unroll
        for (int u = 0; u < EF_BATCH; u++) {
            const u32 e = e0 + u * stride;
            id[u][0] = id[u][1] = id[u][2] = id[u][3] = 0;
            if (e < E) {
                if (SPE == 4) {
                    const uint4 q = reinterpret_cast<const uint4*>(edge_ids)[e];
                    id[u][0] = q.x; id[u][1] = q.y; id[u][2] = q.z; id[u][3] = q.w;
                } else {
                    const uint2 q = reinterpret_cast<const uint2*>(edge_ids)[e];
                    id[u][0] = q.x; id[u][1] = q.y;
                }
            }
        }
        if (slot_id) {
#pragma unroll
            for (int u = 0; u < EF_BATCH; u++) {
                if (e0 + u * stride < E) {
#pragma unroll
                    for (int k = 0; k < SPE; k++) id[u][k] = slot_id[id[u][k]];
                }
            }
#pragma unroll
            for (int u = 0; u < EF_BATCH; u++) {
                const u32 e = e0 + u * stride;
                if (e < E) {
                    if (SPE == 4) reinterpret_cast<uint4*>(edge_ids)[e] = make_uint4(id[u][0], id[u][1], id[u][2], id[u][3]);
                    else reinterpret_cast<uint2*>(edge_ids)[e] = make_uint2(id[u][0], id[u][1]);
                }
            }
        }
#pragma unroll
        for (int u = 0; u < EF_BATCH; u++) {
            if (e0 + u * stride < E)
                record_entries_t<TPE>(id[u], 0u, sym, csc, [&](u32 major, u32 minor, u32 dir, u32) {
                    if (rr.has(major)) {
                        G2N_CHECK(major < ds->rows && minor < ds->n);
                        const u32 pos = atomicAdd(&cursor[major], 1u);
                        G2N_CHECK(pos < ds->M);
                        entries[pos] = Ent32::make(minor, dir, 0u);
                    }
                });
        }
    }
}

// histogram of majors (cnt == NULL: the tokenizer counted the rows already); translates edge_slots to node IDs in
// place and lays the weights out by emission ordinal of the record (WEmit) when there are any
__global__ void __launch_bounds__(256) k_rows_count(const EmitParams E, int sym, int csc, u32* __restrict__ cnt, double* __restrict__ w_rec,
                                                     const RowRange rr, u32* __restrict__ emit_t0)
{
    for_each_edge(E, [&](u32 stored, u32 t0, const u32 (&id)[4]) {
        if (emit_t0) emit_t0[stored] = t0;  // emission index of the record's first triplet, for the flat scatter passes
        if (cnt) record_entries(id, E.tpe, t0, sym, csc, [&](u32 major, u32, u32, u32) { if (rr.has(major)) atomicAdd(&cnt[major], 1u); });
        if (w_rec) w_rec[t0 / (u32)E.tpe] = E.edge_w[stored];  // one weight per record, at its emission ordinal
    });
}

// cursor[major] starts at rowptr[major]; one atomicAdd hands out the entry's position inside the row's
// range (order inside the range is arbitrary; the in-row sort restores a total order)
// Weighted scatter as a flat pass over the stored edge records (node IDs in place, emission index of every
// record from emit_t0): no per-tile bookkeeping, so a pass per row range is cheap (see RowRange).
template <int TPE>
__global__ void __launch_bounds__(256) k_rows_scatter_flat(const u32* __restrict__ edge_ids, const u32* __restrict__ emit_t0, const DevSizes* __restrict__ ds,
                                                            int sym, int csc, u32* __restrict__ cursor, u64* __restrict__ entries, const RowRange rr)
{
    constexpr int SPE = TPE == 4 ? 4 : 2;
    if (!ds->ok) return;
    const u32 E = ds->E;
    const u32 stride = gridDim.x * blockDim.x;
    for (u32 e0 = blockIdx.x * blockDim.x + threadIdx.x; e0 < E; e0 += stride * EF_BATCH) {
        u32 id[EF_BATCH][4], t0[EF_BATCH];
#pragma unroll
        for (int u = 0; u < EF_BATCH; u++) {
            const u32 e = e0 + u * stride;
            id[u][0] = id[u][1] = id[u][2] = id[u][3] = 0;
            t0[u] = 0;
            if (e < E) {
                if (SPE == 4) {
                    const uint4 q = reinterpret_cast<const uint4*>(edge_ids)[e];
                    id[u][0] = q.x; id[u][1] = q.y; id[u][2] = q.z; id[u][3] = q.w;
                } else {
                    const uint2 q = reinterpret_cast<const uint2*>(edge_ids)[e];
                    id[u][0] = q.x; id[u][1] = q.y;
                }
                t0[u] = emit_t0[e];
            }
        }
#pragma unroll
        for (int u = 0; u < EF_BATCH; u++) {
            if (e0 + u * stride < E)
                record_entries_t<TPE>(id[u], t0[u], sym, csc, [&](u32 major, u32 minor, u32 dir, u32 t) {
                    if (rr.has(major)) {
                        G2N_CHECK(major < ds->rows && minor < ds->n && t < ds->T);
                        const u32 pos = atomicAdd(&cursor[major], 1u);
                        G2N_CHECK(pos < ds->M);
                        entries[pos] = Ent64::make(minor, dir, t);
                    }
                });
        }
    }
}

// ---------------------------------------------------------------- row arrays far larger than L2: bucketed build
// The histogram and the scatter above touch a random row per entry.  Once rowcnt + cursors + entries are hundreds of
// megabytes every such access is a DRAM sector of its own (tools/ubench/randmem.cu: atomics on a working set >> L2 run
// at 20 - 30 G/s against 190 G/s inside L2); a pass per row range keeps them in L2 but re-reads and re-filters all edge
// records once per range.  Instead the entries are first PARTITIONED by row bucket (rows >> shift, <= RB_MAX buckets
// sized to stay L2-resident) into a bucket-major pair list -- sequential reads, writes that fill each bucket's frontier
// in pieces reserved per CTA round -- and the histogram / scatter passes then stream over that list: at any moment
// the grid works inside one or two buckets, so their random accesses hit L2.
//   k_bucket_count        flat over the edge records: slots -> IDs in place, entries per bucket (shared-memory histogram)
//   k_bucket_scatter      flat again: (major, entry) -> its bucket's range, one global atomic per bucket and CTA round
//   k_bucket_rows_count   flat over the pair list: cnt[major]++        (skipped when the tokenizer counted the rows)
//   k_bucket_rows_scatter flat over the pair list: entries[cursor[major]++] = entry
#define RB_MAX 128
struct RowBuckets {
    u32 shift, count;  // bucket of a row = min(row >> shift, count - 1)
    __device__ __forceinline__ u32 of(u32 major) const
    {
        const u32 b = major >> shift;
        return b < count ? b : count - 1;
    }
};
struct BucketCtl {  // zeroed before k_bucket_count
    u32 cnt[RB_MAX];  // entries per bucket
    u32 cur[RB_MAX];  // entries placed so far (k_bucket_scatter)
};

// the entries of one record with a compile-time index each (max(S, S^T) only exists for one-triplet records: symmax
// implies graph_directed, g2n.cu tokenize_phase), so that per-entry state stays in registers
template <int TPE, class G>
__device__ __forceinline__ void rb_entries(const u32 (&id)[4], u32 t0, int sym, int csc, G g)
{
#pragma unroll
    for (int k = 0; k < TPE; k++) {
        const u32 a = id[k & 2], b = id[(k & 2) + 1];
        const u32 r = (k & 1) ? b : a, c = (k & 1) ? a : b;
        if (TPE == 1 && sym) { g(0, r, c, 0u, t0); g(1, c, r, 1u, t0); }
        else if (csc) g(k, c, r, 0u, t0 + k);
        else g(k, r, c, 0u, t0 + k);
    }
}

// exclusive prefix of n <= 256 shared-memory counters by a 256-thread CTA: lo[j] = sum of cnt[0 .. j), lo[n] = total.
// `warp_sums`: 8 words of shared memory.  Starts and ends with a CTA barrier.
__device__ __forceinline__ void rb_prefix256(const u32* cnt, u32 n, u32* lo, u32* warp_sums)
{
    __syncthreads();
    const u32 lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const u32 v = threadIdx.x < n ? cnt[threadIdx.x] : 0u;
    const u32 inc = warp_incl_scan(v);
    if (lane == 31) warp_sums[wid] = inc;
    __syncthreads();
    u32 base = 0;
#pragma unroll
    for (u32 w = 0; w < 8; w++) base += w < wid ? warp_sums[w] : 0u;
    if (threadIdx.x < n) lo[threadIdx.x] = base + inc - v;
    if (threadIdx.x == n) lo[n] = base + inc - v;  // everything below n (v == 0 from n on); n == 256: see below
    if (n == 256 && threadIdx.x == 255) lo[256] = base + inc;
    __syncthreads();
}

template <int SPE>
__device__ __forceinline__ void rb_load(const u32* __restrict__ recs, u32 e, u32 (&id)[4])
{
    if (SPE == 4) {
        const uint4 q = reinterpret_cast<const uint4*>(recs)[e];
        id[0] = q.x; id[1] = q.y; id[2] = q.z; id[3] = q.w;
    } else {
        const uint2 q = reinterpret_cast<const uint2*>(recs)[e];
        id[0] = q.x; id[1] = q.y; id[2] = 0; id[3] = 0;
    }
}

template <int TPE>
__global__ void __launch_bounds__(256) k_bucket_count(u32* __restrict__ edge_slots, const u32* __restrict__ slot_id, const DevSizes* __restrict__ ds,
                                                       int sym, int csc, const RowBuckets rb, BucketCtl* __restrict__ ctl)
{
    constexpr int SPE = TPE == 4 ? 4 : 2;
    __shared__ u32 s_cnt[8][RB_MAX];  // one histogram per warp: less contention on the shared-memory atomics
    for (u32 i = threadIdx.x; i < 8 * RB_MAX; i += 256) (&s_cnt[0][0])[i] = 0;
    __syncthreads();
    if (!ds->ok) return;
    const u32 E = ds->E, wid = threadIdx.x >> 5;
    const u32 stride = gridDim.x * blockDim.x;
    for (u32 e0 = blockIdx.x * blockDim.x + threadIdx.x; e0 < E; e0 += stride * EF_BATCH) {
        u32 id[EF_BATCH][4];
#pragma unroll
        for (int u = 0; u < EF_BATCH; u++) {
            id[u][0] = id[u][1] = id[u][2] = id[u][3] = 0;
            if (e0 + u * stride < E) rb_load<SPE>(edge_slots, e0 + u * stride, id[u]);
        }
        if (slot_id) {  // table slots -> node IDs, left in place for the later passes
#pragma unroll
            for (int u = 0; u < EF_BATCH; u++) {
                if (e0 + u * stride < E) {
#pragma unroll
                    for (int k = 0; k < SPE; k++) id[u][k] = slot_id[id[u][k]];
                }
            }
#pragma unroll
            for (int u = 0; u < EF_BATCH; u++) {
                const u32 e = e0 + u * stride;
                if (e < E) {
                    if (SPE == 4) reinterpret_cast<uint4*>(edge_slots)[e] = make_uint4(id[u][0], id[u][1], id[u][2], id[u][3]);
                    else reinterpret_cast<uint2*>(edge_slots)[e] = make_uint2(id[u][0], id[u][1]);
                }
            }
        }
#pragma unroll
        for (int u = 0; u < EF_BATCH; u++) {
            if (e0 + u * stride < E)
                rb_entries<TPE>(id[u], 0u, sym, csc, [&](int, u32 major, u32, u32, u32) { atomicAdd(&s_cnt[wid][rb.of(major)], 1u); });
        }
    }
    __syncthreads();
    if (threadIdx.x < rb.count) {
        u32 c = 0;
#pragma unroll
        for (int w = 0; w < 8; w++) c += s_cnt[w][threadIdx.x];
        if (c) atomicAdd(&ctl->cnt[threadIdx.x], c);
    }
}

// ENT = Ent32 (unweighted: emit_t0 == NULL) | Ent64 (weighted: the record's first emission index comes from emit_t0)
// A CTA round takes 256 x RB_RECS<TPE> records = at most RB_ROUND entries: they are grouped by bucket in shared memory
// (rank inside the round from a shared-memory atomic, group offsets from a 64-element prefix), the round's range in
// every bucket is reserved with one global atomic, and consecutive threads then copy consecutive staged elements -- the
// bucket frontiers receive whole runs instead of one 4-byte store per lane (uncoalesced stores ran the partition at
// 0.5 TB/s with 32 buckets, profiles/r2b_buckets.md).
#define RB_ROUND 2048
template <int TPE> struct RbRecs { static constexpr int value = TPE == 4 ? 2 : 4; };  // records per thread and round
template <int TPE, class ENT>
__global__ void __launch_bounds__(256) k_bucket_scatter(const u32* __restrict__ edge_ids, const u32* __restrict__ emit_t0, const DevSizes* __restrict__ ds,
                                                         int sym, int csc, const RowBuckets rb, BucketCtl* __restrict__ ctl,
                                                         u32* __restrict__ pair_major, typename ENT::type* __restrict__ pair_ent)
{
    typedef typename ENT::type EV;
    constexpr int SPE = TPE == 4 ? 4 : 2;
    constexpr int EPR = TPE == 1 ? 2 : TPE;  // entries of one record at most (TPE == 1: two when sym)
    constexpr int RECS = RbRecs<TPE>::value;
    static_assert(256 * RECS * EPR <= RB_ROUND, "round size");
    __shared__ u32 s_off[RB_MAX], s_cnt[RB_MAX], s_lo[RB_MAX + 1], s_base[RB_MAX], s_ws[8];
    __shared__ u32 s_major[RB_ROUND];
    __shared__ EV s_ent[RB_ROUND];
    if (!ds->ok) return;
    if (threadIdx.x == 0) {
        u32 run = 0;
        for (u32 b = 0; b < rb.count; b++) { s_off[b] = run; run += ctl->cnt[b]; }
    }
    const u32 E = ds->E, M = ds->M;
    const u32 per_round = 256 * RECS;
    for (u64 r0 = (u64)blockIdx.x * per_round; r0 < E; r0 += (u64)gridDim.x * per_round) {  // uniform per CTA
        if (threadIdx.x < RB_MAX) s_cnt[threadIdx.x] = 0;
        __syncthreads();
        u32 id[RECS][4], t0[RECS];
        unsigned short rk[RECS][EPR];  // rank of my entries among the round's entries of the same bucket
#pragma unroll
        for (int u = 0; u < RECS; u++) {
            const u64 e = r0 + u * 256 + threadIdx.x;
            id[u][0] = id[u][1] = id[u][2] = id[u][3] = 0;
            t0[u] = 0;
            if (e < E) {
                rb_load<SPE>(edge_ids, (u32)e, id[u]);
                if (ENT::kWeighted) t0[u] = emit_t0[e];
            }
        }
#pragma unroll
        for (int u = 0; u < RECS; u++) {
            if (r0 + u * 256 + threadIdx.x < E)
                rb_entries<TPE>(id[u], 0u, sym, csc, [&](int q, u32 major, u32, u32, u32) { rk[u][q] = (unsigned short)atomicAdd(&s_cnt[rb.of(major)], 1u); });
        }
        rb_prefix256(s_cnt, rb.count, s_lo, s_ws);
        if (threadIdx.x < rb.count) {
            const u32 c = s_cnt[threadIdx.x];
            s_base[threadIdx.x] = s_off[threadIdx.x] + (c ? atomicAdd(&ctl->cur[threadIdx.x], c) : 0u);
        }
        __syncthreads();
#pragma unroll
        for (int u = 0; u < RECS; u++) {
            if (r0 + u * 256 + threadIdx.x < E)
                rb_entries<TPE>(id[u], t0[u], sym, csc, [&](int q, u32 major, u32 minor, u32 dir, u32 t) {
                    const u32 p = s_lo[rb.of(major)] + rk[u][q];
                    s_major[p] = major;
                    s_ent[p] = ENT::make(minor, dir, t);
                });
        }
        __syncthreads();
        const u32 total = s_lo[rb.count];
        for (u32 p = threadIdx.x; p < total; p += 256) {
            const u32 major = s_major[p];
            const u32 b = rb.of(major);
            const u32 pos = s_base[b] + (p - s_lo[b]);
            G2N_CHECK(pos < M);
            if (pos < M) {
                pair_major[pos] = major;
                pair_ent[pos] = s_ent[p];
            }
        }
        __syncthreads();
    }
}

// entries in the pair list = sum of the bucket counters (= DevSizes.M on one GPU; a multi-GPU slab may have dropped
// entries of a failed exchange)
__device__ __forceinline__ u64 rb_total(const BucketCtl* ctl, u32 nb)
{
    __shared__ u64 s_total;
    if (threadIdx.x == 0) {
        u64 t = 0;
        for (u32 b = 0; b < nb; b++) t += ctl->cnt[b];
        s_total = t;
    }
    __syncthreads();
    return s_total;
}

__global__ void __launch_bounds__(256) k_bucket_rows_count(const u32* __restrict__ pair_major, const DevSizes* __restrict__ ds, const BucketCtl* __restrict__ ctl, u32 nb,
                                                            u32* __restrict__ cnt)
{
    if (!ds->ok) return;
    const u64 M = rb_total(ctl, nb), stride = (u64)gridDim.x * blockDim.x;
    for (u64 i0 = (u64)blockIdx.x * blockDim.x + threadIdx.x; i0 < M; i0 += stride * 4) {
        u32 m[4];
#pragma unroll
        for (int u = 0; u < 4; u++) m[u] = i0 + u * stride < M ? pair_major[i0 + u * stride] : 0xFFFFFFFFu;
#pragma unroll
        for (int u = 0; u < 4; u++)
            if (m[u] != 0xFFFFFFFFu) { G2N_CHECK(m[u] < ds->rows); atomicAdd(&cnt[m[u]], 1u); }
    }
}

template <class ENT>
__global__ void __launch_bounds__(256) k_bucket_rows_scatter(const u32* __restrict__ pair_major, const typename ENT::type* __restrict__ pair_ent,
                                                              const DevSizes* __restrict__ ds, const BucketCtl* __restrict__ ctl, u32 nb,
                                                              u32* __restrict__ cursor, typename ENT::type* __restrict__ entries)
{
    if (!ds->ok) return;
    const u64 M = rb_total(ctl, nb), stride = (u64)gridDim.x * blockDim.x;
    for (u64 i0 = (u64)blockIdx.x * blockDim.x + threadIdx.x; i0 < M; i0 += stride * 4) {
        u32 m[4];
        typename ENT::type v[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const bool in = i0 + u * stride < M;
            m[u] = in ? pair_major[i0 + u * stride] : 0xFFFFFFFFu;
            v[u] = in ? pair_ent[i0 + u * stride] : (typename ENT::type)0;
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
            if (m[u] != 0xFFFFFFFFu) {
                const u32 pos = atomicAdd(&cursor[m[u]], 1u);
                G2N_CHECK(m[u] < ds->rows && pos < M);
                entries[pos] = v[u];
            }
        }
    }
}

// ---------------------------------------------------------------- second level: sub-buckets that fit shared memory
// Inside an L2-sized bucket every entry still costs an atomic with a return value and a store of its own (about 70 G
// entries/s whatever the locality: the L2 transaction rate, not its hit rate).  So each bucket is partitioned once more,
// into <= SB_FAN sub-buckets of a few thousand entries (a power-of-two number of consecutive rows), and the last two
// passes work on one sub-bucket per CTA entirely in shared memory: the row histogram is written out coalesced, and the
// entries are grouped by row in shared memory and copied to their final place, which is one contiguous range.
//   k_sub_count        bucket-major pair list -> entries per sub-bucket
//   k_sub_scatter      pair list -> sub-bucket-major pair list (same staging as k_bucket_scatter)
//   k_sub_rows_count   one CTA per sub-bucket: cnt[row] for its rows     (skipped when the tokenizer counted the rows)
//   k_sub_rows_place   one CTA per sub-bucket: entries of its rows, grouped by row (any order inside a row; a sub-bucket
//                      that does not fit -- a few very long rows -- goes through the global cursors instead)
#define SB_FAN 256
#define SB_UNSORTED 0xFFFFFFFFu  // ucnt of a row the placement did not sort (longer than RS_SMALL, or an overflowing sub-bucket)
#define SB_ROWS_MAX 4096  // rows of a sub-bucket at most (shared-memory histogram / cursors)
struct SubBuckets {
    u32 shift2;     // sub-bucket of a row = row >> shift2 (numbered across buckets: bucket = sub-bucket >> fan_shift)
    u32 fan_shift;  // bucket shift - shift2, <= 8
    u32 n_sub;      // buckets << fan_shift
    u32 cap;        // entries a sub-bucket may hold for the shared-memory placement
};

// prefix of the bucket counters in shared memory: s_off[0 .. count]
__device__ __forceinline__ void rb_offsets(const BucketCtl* ctl, u32 count, u32* s_off)
{
    if (threadIdx.x == 0) {
        u32 run = 0;
        for (u32 b = 0; b < count; b++) { s_off[b] = run; run += ctl->cnt[b]; }
        s_off[count] = run;
    }
    __syncthreads();
}
// every bucket that intersects the pair-list range [c0, c1): f(bucket, lo, hi), uniform across the CTA
template <class F>
__device__ __forceinline__ void rb_segments(const u32* s_off, u32 count, u64 c0, u64 c1, F f)
{
    u32 b = 0, hi_b = count;  // first bucket that ends after c0 (binary search: this runs once per round in every thread)
    while (b + 1 < hi_b) {
        const u32 mid = (b + hi_b) >> 1;
        if (s_off[mid] <= c0) b = mid; else hi_b = mid;
    }
    for (; b < count && s_off[b] < c1; b++) {
        const u64 lo = c0 > s_off[b] ? c0 : s_off[b], hi = c1 < s_off[b + 1] ? c1 : s_off[b + 1];
        if (lo < hi) f(b, lo, hi);
    }
}

#define SB_CHUNK (RB_ROUND * 8)
__global__ void __launch_bounds__(256) k_sub_count(const u32* __restrict__ pair_major, const DevSizes* __restrict__ ds, const BucketCtl* __restrict__ ctl,
                                                    const RowBuckets rb, const SubBuckets sb, u32* __restrict__ sub_cnt)
{
    __shared__ u32 s_off[RB_MAX + 1];
    __shared__ u32 s_hist[SB_FAN];
    if (!ds->ok) return;
    rb_offsets(ctl, rb.count, s_off);
    const u64 total = s_off[rb.count];
    const u32 fan_mask = (1u << sb.fan_shift) - 1u;
    for (u64 c0 = (u64)blockIdx.x * SB_CHUNK; c0 < total; c0 += (u64)gridDim.x * SB_CHUNK) {
        rb_segments(s_off, rb.count, c0, c0 + SB_CHUNK, [&](u32 b, u64 lo, u64 hi) {
            s_hist[threadIdx.x] = 0;  // SB_FAN == blockDim.x
            __syncthreads();
            for (u64 i = lo + threadIdx.x; i < hi; i += 256) atomicAdd(&s_hist[(pair_major[i] >> sb.shift2) & fan_mask], 1u);
            __syncthreads();
            const u32 c = s_hist[threadIdx.x];
            if (c) atomicAdd(&sub_cnt[(b << sb.fan_shift) + threadIdx.x], c);
            __syncthreads();
        });
    }
}

template <class ENT>
__global__ void __launch_bounds__(256) k_sub_scatter(const u32* __restrict__ pair_major, const typename ENT::type* __restrict__ pair_ent,
                                                      const DevSizes* __restrict__ ds, const BucketCtl* __restrict__ ctl, const RowBuckets rb, const SubBuckets sb,
                                                      const u32* __restrict__ sub_off, u32* __restrict__ sub_cur,
                                                      u32* __restrict__ out_major, typename ENT::type* __restrict__ out_ent)
{
    typedef typename ENT::type EV;
    constexpr int PER = RB_ROUND / 256;
    __shared__ u32 s_off[RB_MAX + 1];
    __shared__ u32 s_cnt[SB_FAN], s_lo[SB_FAN + 1], s_base[SB_FAN], s_ws[8];
    __shared__ u32 s_major[RB_ROUND];
    __shared__ EV s_ent[RB_ROUND];
    if (!ds->ok) return;
    rb_offsets(ctl, rb.count, s_off);
    const u64 total = s_off[rb.count];
    const u32 fan = 1u << sb.fan_shift, fan_mask = fan - 1u;
    for (u64 c0 = (u64)blockIdx.x * RB_ROUND; c0 < total; c0 += (u64)gridDim.x * RB_ROUND) {
        rb_segments(s_off, rb.count, c0, c0 + RB_ROUND, [&](u32 b, u64 lo, u64 hi) {
            s_cnt[threadIdx.x] = 0;
            __syncthreads();
            u32 mj[PER];
            EV en[PER];
            unsigned short rk[PER];
#pragma unroll
            for (int u = 0; u < PER; u++) {
                const u64 i = lo + u * 256 + threadIdx.x;
                mj[u] = 0xFFFFFFFFu;
                en[u] = (EV)0;
                if (i < hi) { mj[u] = pair_major[i]; en[u] = pair_ent[i]; }
            }
#pragma unroll
            for (int u = 0; u < PER; u++)
                if (mj[u] != 0xFFFFFFFFu) rk[u] = (unsigned short)atomicAdd(&s_cnt[(mj[u] >> sb.shift2) & fan_mask], 1u);
            rb_prefix256(s_cnt, fan, s_lo, s_ws);
            if (threadIdx.x < fan) {
                const u32 c = s_cnt[threadIdx.x], g = (b << sb.fan_shift) + threadIdx.x;
                s_base[threadIdx.x] = sub_off[g] + (c ? atomicAdd(&sub_cur[g], c) : 0u);
            }
            __syncthreads();
#pragma unroll
            for (int u = 0; u < PER; u++) {
                if (mj[u] != 0xFFFFFFFFu) {
                    const u32 p = s_lo[(mj[u] >> sb.shift2) & fan_mask] + rk[u];
                    s_major[p] = mj[u];
                    s_ent[p] = en[u];
                }
            }
            __syncthreads();
            const u32 n = s_lo[fan];
            for (u32 p = threadIdx.x; p < n; p += 256) {
                const u32 major = s_major[p];
                const u32 j = (major >> sb.shift2) & fan_mask;
                const u64 pos = (u64)s_base[j] + (p - s_lo[j]);
                out_major[pos] = major;
                out_ent[pos] = s_ent[p];
            }
            __syncthreads();
        });
    }
}

__global__ void __launch_bounds__(256) k_sub_rows_count(const u32* __restrict__ pair_major, const DevSizes* __restrict__ ds, const SubBuckets sb,
                                                         const u32* __restrict__ sub_off, u32* __restrict__ cnt)
{
    __shared__ u32 s_hist[SB_ROWS_MAX];
    if (!ds->ok) return;
    const u32 rows = ds->rows, per = 1u << sb.shift2;
    for (u32 g = blockIdx.x; g < sb.n_sub; g += gridDim.x) {  // uniform per CTA
        const u32 row0 = g << sb.shift2;
        if (row0 >= rows) break;
        const u32 nr = min(per, rows - row0);
        const u32 lo = sub_off[g], hi = sub_off[g + 1];
        for (u32 i = threadIdx.x; i < nr; i += 256) s_hist[i] = 0;
        __syncthreads();
        for (u32 i = lo + threadIdx.x; i < hi; i += 256) {
            const u32 r = pair_major[i] - row0;
            G2N_CHECK(r < nr);
            atomicAdd(&s_hist[r], 1u);
        }
        __syncthreads();
        for (u32 i = threadIdx.x; i < nr; i += 256) cnt[row0 + i] = s_hist[i];
        __syncthreads();
    }
}

// (k_sub_rows_place: at the end of this file, after the row-sorting helpers it uses)

// same two steps for caller-provided COO arrays (g2n_coo_to_compressed)
__global__ void __launch_bounds__(256) k_coo_count(const int32_t* __restrict__ row, const int32_t* __restrict__ col, u64 nnz, int csc, u32* __restrict__ cnt)
{
    for (u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x; t < nnz; t += (u64)gridDim.x * blockDim.x)
        atomicAdd(&cnt[(u32)(csc ? col[t] : row[t])], 1u);
}
__global__ void __launch_bounds__(256) k_coo_scatter(const int32_t* __restrict__ row, const int32_t* __restrict__ col, u64 nnz, int csc,
                                                      u32* __restrict__ cursor, u64* __restrict__ entries)
{
    for (u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x; t < nnz; t += (u64)gridDim.x * blockDim.x) {
        const u32 major = (u32)(csc ? col[t] : row[t]), minor = (u32)(csc ? row[t] : col[t]);
        entries[atomicAdd(&cursor[major], 1u)] = rs_entry(minor, 0u, (u32)t);
    }
}

// ---------------------------------------------------------------- long rows (rare)
// Load functor of the rowptr scan: lists the rows longer than RS_SMALL while the counts stream by.
struct LoadRowCounts {
    const u32* cnt;
    u32* biglist;
    u32* bigcount;
    __device__ __forceinline__ u64 operator()(u64 i) const
    {
        const u32 c = cnt[i];
        if (c > RS_SMALL) biglist[atomicAdd(bigcount, 1u)] = (u32)i;
        return (u64)c;
    }
    __device__ __forceinline__ u64 peek(u64 i) const { return (u64)cnt[i]; }
};

#define RS_BIG_SMEM 4096
// One CTA sorts one long row with a normalized bitonic network (every comparison ascending) over the
// next power of two; indices past the end act as +infinity and are never touched.  Rows that fit are
// sorted in shared memory, longer ones in place in global memory.
template <class ENT>
__global__ void __launch_bounds__(256) k_rows_big(const u32* __restrict__ rowptr, const u32* __restrict__ biglist,
                                                   const u32* __restrict__ bigcount, typename ENT::type* __restrict__ entries)
{
    typedef typename ENT::type E;
    __shared__ E s_big[RS_BIG_SMEM];
    const u32 nbig = *bigcount;
    for (u32 b = blockIdx.x; b < nbig; b += gridDim.x) {
        const u32 r = biglist[b];
        const u32 lo = rowptr[r], len = rowptr[r + 1] - lo;
        E* a = entries + lo;
        const bool in_smem = len <= RS_BIG_SMEM;
        if (in_smem) {
            for (u32 i = threadIdx.x; i < len; i += blockDim.x) s_big[i] = a[i];
            a = s_big;
        }
        __syncthreads();
        u32 p2 = 1;
        while (p2 < len) p2 <<= 1;
        for (u32 k = 2; k <= p2; k <<= 1) {
            for (u32 j = k >> 1; j > 0; j >>= 1) {
                const bool mirror = (j == (k >> 1));
                for (u32 i = threadIdx.x; i < p2; i += blockDim.x) {
                    const u32 l = mirror ? (i ^ (k - 1)) : (i ^ j);
                    if (l > i && l < len) {
                        const E x = a[i], y = a[l];
                        if (x > y) { a[i] = y; a[l] = x; }
                    }
                }
                __syncthreads();
            }
        }
        if (in_smem)
            for (u32 i = threadIdx.x; i < len; i += blockDim.x) entries[lo + i] = s_big[i];
        __syncthreads();
    }
}

// ---------------------------------------------------------------- finalize
template <typename T>
struct RowAcc {
    // accumulates one (major, minor) group: a = sum of dir-0 weights, b = sum of dir-1 weights
    T a, b;
    bool has_a, has_b;
    __device__ __forceinline__ void reset() { a = zero_t<T>(); b = zero_t<T>(); has_a = has_b = false; }
    __device__ __forceinline__ void add(u32 dir, T x)
    {
        if (dir) { b = has_b ? add_t<T>(b, x) : x; has_b = true; }
        else { a = has_a ? add_t<T>(a, x) : x; has_a = true; }
    }
    // sym == 0: csr_sum_duplicates keeps explicit zeros; sym == 1: csr_maximum_csr drops results == 0
    __device__ __forceinline__ bool result(int sym, T& v) const
    {
        if (sym) { v = lt_t<T>(a, b) ? b : a; return nz_t<T>(v); }
        v = a;
        return true;
    }
};

// Weights of a build, one per edge RECORD in emission order: the 1 / 2 / 4 triplets of a record share it, so the
// triplet's emission index t finds it at t >> shift (shift 0: one weight per entry -- multi-GPU slabs)
struct WEmit {
    const double* p;
    u32 shift;
};

template <typename T, class ENT>
__device__ __forceinline__ T entry_weight(typename ENT::type e, const WEmit w_emit, const T* __restrict__ w_typed)
{
    if (ENT::kWeighted) {
        if (w_typed) return w_typed[ENT::t(e)];
        if (w_emit.p) return cast_weight<T>(w_emit.p[ENT::t(e) >> w_emit.shift]);
    }
    return cast_weight<T>(1.0);
}

// Walks one sorted row; emit(k, minor, value) is called for every stored result.  Returns their count.
template <typename T, class ENT, class Emit>
__device__ __forceinline__ u32 walk_row(const typename ENT::type* a, u32 len, int sym, const WEmit w_emit, const T* w_typed, Emit emit)
{
    u32 out = 0;
    u32 i = 0;
    while (i < len) {
        const u32 minor = ENT::minor(a[i]);
        RowAcc<T> acc;
        acc.reset();
        while (i < len && ENT::minor(a[i]) == minor) {
            acc.add(ENT::dir(a[i]), entry_weight<T, ENT>(a[i], w_emit, w_typed));
            i++;
        }
        T v;
        if (acc.result(sym, v)) { emit(out, minor, v); out++; }
    }
    return out;
}

template <typename E>
__device__ __forceinline__ void insertion_sort(E* a, u32 len)
{
    for (u32 i = 1; i < len; i++) {
        const E x = a[i];
        u32 j = i;
        while (j > 0 && a[j - 1] > x) { a[j] = a[j - 1]; j--; }
        a[j] = x;
    }
}

// ---- rows of <= 16 entries live in registers: Batcher's odd-even merge sort network (19 / 63
// comparators for 8 / 16 inputs, every lane runs the same instructions) and an unrolled walk
template <typename E>
__device__ __forceinline__ void cmp_swap(E& x, E& y)
{
    const E lo = x < y ? x : y, hi = x < y ? y : x;
    x = lo;
    y = hi;
}

// comparator lists generated from Batcher's construction (checked with the 0-1 principle); written out
// so that every register index is a literal
#define CS(a, b) cmp_swap(v[a], v[b]);
template <typename E>
__device__ __forceinline__ void sort_network(E (&v)[8])
{
    CS(0, 1) CS(2, 3) CS(4, 5) CS(6, 7) CS(0, 2) CS(1, 3) CS(4, 6) CS(5, 7) CS(1, 2) CS(5, 6) CS(0, 4) CS(1, 5)
    CS(2, 6) CS(3, 7) CS(2, 4) CS(3, 5) CS(1, 2) CS(3, 4) CS(5, 6)
}
template <typename E>
__device__ __forceinline__ void sort_network(E (&v)[16])
{
    CS(0, 1) CS(2, 3) CS(4, 5) CS(6, 7) CS(8, 9) CS(10, 11) CS(12, 13) CS(14, 15) CS(0, 2) CS(1, 3) CS(4, 6) CS(5, 7)
    CS(8, 10) CS(9, 11) CS(12, 14) CS(13, 15) CS(1, 2) CS(5, 6) CS(9, 10) CS(13, 14) CS(0, 4) CS(1, 5) CS(2, 6)
    CS(3, 7) CS(8, 12) CS(9, 13) CS(10, 14) CS(11, 15) CS(2, 4) CS(3, 5) CS(10, 12) CS(11, 13) CS(1, 2) CS(3, 4)
    CS(5, 6) CS(9, 10) CS(11, 12) CS(13, 14) CS(0, 8) CS(1, 9) CS(2, 10) CS(3, 11) CS(4, 12) CS(5, 13) CS(6, 14)
    CS(7, 15) CS(4, 8) CS(5, 9) CS(6, 10) CS(7, 11) CS(2, 4) CS(3, 5) CS(6, 8) CS(7, 9) CS(10, 12) CS(11, 13)
    CS(1, 2) CS(3, 4) CS(5, 6) CS(7, 8) CS(9, 10) CS(11, 12) CS(13, 14)
}
#undef CS

// the sorted row in registers (entries past `len` are all-ones); same contract as walk_row
template <int N, typename T, class ENT, class Emit>
__device__ __forceinline__ u32 walk_regs(const typename ENT::type (&v)[N], u32 len, int sym, const WEmit w_emit, const T* w_typed, Emit emit)
{
    u32 out = 0;
    RowAcc<T> acc;
    acc.reset();
#pragma unroll
    for (int k = 0; k < N; k++) {
        if ((u32)k < len) {
            const u32 minor = ENT::minor(v[k]);
            acc.add(ENT::dir(v[k]), entry_weight<T, ENT>(v[k], w_emit, w_typed));
            bool last = (u32)(k + 1) == len;
            if (k + 1 < N) last = last || ENT::minor(v[k + 1 < N ? k + 1 : k]) != minor;
            if (last) {
                T val;
                if (acc.result(sym, val)) { emit(out, minor, val); out++; }
                acc.reset();
            }
        }
    }
    return out;
}

template <int N, class ENT>
__device__ __forceinline__ void load_row(typename ENT::type (&v)[N], const typename ENT::type* a, u32 len)
{
#pragma unroll
    for (int k = 0; k < N; k++) v[k] = (u32)k < len ? a[k] : ~(typename ENT::type)0;
}

// phase 1: sort the row (written back in place) and count the entries it will store
template <int N, typename T, class ENT>
__device__ __forceinline__ u32 row_sort_count(typename ENT::type* a, u32 len, int sym, const WEmit w_emit, const T* w_typed)
{
    typename ENT::type v[N];
    load_row<N, ENT>(v, a, len);
    sort_network(v);
#pragma unroll
    for (int k = 0; k < N; k++)
        if ((u32)k < len) a[k] = v[k];
    if (ENT::kWeighted && sym) return walk_regs<N, T, ENT>(v, len, sym, w_emit, w_typed, [](u32, u32, T) {});  // zeros of max() are dropped
    u32 heads = 0;  // every distinct minor is stored (explicit zeros are kept, counts are never zero)
#pragma unroll
    for (int k = 0; k < N; k++) heads += ((u32)k < len && (k == 0 || ENT::minor(v[k]) != ENT::minor(v[k > 0 ? k - 1 : 0]))) ? 1u : 0u;
    return heads;
}

// phase 2: walk the sorted row and write its results
template <int N, typename T, class ENT>
__device__ __forceinline__ void row_emit(const typename ENT::type* a, u32 len, int sym, const WEmit w_emit, const T* w_typed, u32 out0,
                                         int32_t* __restrict__ indices, T* __restrict__ data)
{
    typename ENT::type v[N];
    load_row<N, ENT>(v, a, len);
    walk_regs<N, T, ENT>(v, len, sym, w_emit, w_typed, [&](u32 k, u32 minor, T val) {
        indices[out0 + k] = (int32_t)minor;
        data[out0 + k] = val;
    });
}

#define RF_ROWS 256       // rows per chunk = threads per CTA
#define RF_SMEM_ENT 4096  // entries of a chunk staged in shared memory (else: in place in global memory)

// one coalesced copy of the chunk's entries into shared memory, eight loads in flight per thread
template <typename E>
__device__ __forceinline__ void stage_chunk(E* s_ent, const E* __restrict__ src, u32 c_len)
{
    for (u32 i0 = threadIdx.x; i0 < c_len; i0 += 8 * RF_ROWS) {
        E tmp[8];
#pragma unroll
        for (int u = 0; u < 8; u++) {
            const u32 i = i0 + u * RF_ROWS;
            if (i < c_len) tmp[u] = src[i];
        }
#pragma unroll
        for (int u = 0; u < 8; u++) {
            const u32 i = i0 + u * RF_ROWS;
            if (i < c_len) s_ent[i] = tmp[u];
        }
    }
}

// Pass A: one CTA per chunk of RF_ROWS consecutive rows, one lane per row.  The chunk's entries are
// staged in shared memory, every row is sorted (per warp the longest of its 32 rows picks the path: 8- or
// 16-input sorting network in registers, else insertion sort) and written back, and the number of
// entries each row will store goes to ucnt.  No CTA depends on another one.
template <typename T, class ENT>
__global__ void __launch_bounds__(RF_ROWS) k_rows_sort(const u32* __restrict__ rowptr, typename ENT::type* __restrict__ entries,
                                                        const u32* __restrict__ n_dev, int sym, const WEmit w_emit,
                                                        const T* __restrict__ w_typed, u32* __restrict__ ucnt)
{
    typedef typename ENT::type E;
    __shared__ E s_ent[RF_SMEM_ENT];
    const u32 n = *n_dev;
    const u32 n_chunks = (n + RF_ROWS - 1) / RF_ROWS;
    for (u32 chunk = blockIdx.x; chunk < n_chunks; chunk += gridDim.x) {
        const u32 r0 = chunk * RF_ROWS, r = r0 + threadIdx.x;
        const u32 r_end = min(r0 + RF_ROWS, n);
        const bool live = r < n;
        const u32 lo = rowptr[live ? r : r_end], hi = rowptr[live ? r + 1 : r_end];
        const u32 len = hi - lo;
        const u32 c_lo = rowptr[r0], c_len = rowptr[r_end] - c_lo;
        const bool staged = c_len <= RF_SMEM_ENT;
        E* a;
        if (staged) {
            stage_chunk<E>(s_ent, entries + c_lo, c_len);
            __syncthreads();
            a = s_ent + (lo - c_lo);
        } else {
            a = entries + lo;  // in place in global memory (rows > RS_SMALL were sorted by k_rows_big)
        }
        const u32 wmax = __reduce_max_sync(0xffffffffu, len);  // warp-uniform path
        u32 mine;
        if (wmax <= 8) {
            mine = row_sort_count<8, T, ENT>(a, len, sym, w_emit, w_typed);
        } else if (wmax <= 16) {
            mine = row_sort_count<16, T, ENT>(a, len, sym, w_emit, w_typed);
        } else {
            if (len > 1 && len <= RS_SMALL) insertion_sort<E>(a, len);
            mine = walk_row<T, ENT>(a, len, sym, w_emit, w_typed, [](u32, u32, T) {});
        }
        if (live) ucnt[r] = mine;
        if (staged) {
            __syncthreads();
            for (u32 i = threadIdx.x; i < c_len; i += RF_ROWS) entries[c_lo + i] = s_ent[i];
            __syncthreads();  // s_ent is reused by the next chunk
        }
    }
}

// Pass A after a placement that sorted the short rows itself (k_sub_rows_place): only the rows it marked SB_UNSORTED --
// rows longer than RS_SMALL (sorted by k_rows_big meanwhile) and the rows of overflowing sub-buckets -- in place in
// global memory, one lane per row.
template <typename T, class ENT>
__global__ void __launch_bounds__(256) k_rows_sort_rest(const u32* __restrict__ rowptr, typename ENT::type* __restrict__ entries,
                                                         const u32* __restrict__ n_dev, int sym, const WEmit w_emit,
                                                         const T* __restrict__ w_typed, u32* __restrict__ ucnt)
{
    typedef typename ENT::type E;
    const u32 n = *n_dev;
    for (u32 r = blockIdx.x * blockDim.x + threadIdx.x; r < n; r += gridDim.x * blockDim.x) {
        if (ucnt[r] != SB_UNSORTED) continue;
        const u32 lo = rowptr[r], len = rowptr[r + 1] - lo;
        E* a = entries + lo;
        if (len > 1 && len <= RS_SMALL) insertion_sort<E>(a, len);
        ucnt[r] = walk_row<T, ENT>(a, len, sym, w_emit, w_typed, [](u32, u32, T) {});
    }
}

// Pass B: indptr (exclusive scan of ucnt) is known.  Same chunking: stage the sorted entries, walk every
// row, write indices / data.
template <typename T, class ENT>
__global__ void __launch_bounds__(RF_ROWS) k_rows_write(const u32* __restrict__ rowptr, const typename ENT::type* __restrict__ entries,
                                                         const u32* __restrict__ n_dev, int sym, const WEmit w_emit,
                                                         const T* __restrict__ w_typed, const int32_t* __restrict__ indptr,
                                                         int32_t* __restrict__ indices, T* __restrict__ data, u32* __restrict__ nnz_out)
{
    typedef typename ENT::type E;
    __shared__ E s_ent[RF_SMEM_ENT];
    const u32 n = *n_dev;
    const u32 n_chunks = (n + RF_ROWS - 1) / RF_ROWS;
    if (blockIdx.x == 0 && threadIdx.x == 0) *nnz_out = (u32)indptr[n];
    for (u32 chunk = blockIdx.x; chunk < n_chunks; chunk += gridDim.x) {
        const u32 r0 = chunk * RF_ROWS, r = r0 + threadIdx.x;
        const u32 r_end = min(r0 + RF_ROWS, n);
        const bool live = r < n;
        const u32 lo = rowptr[live ? r : r_end], hi = rowptr[live ? r + 1 : r_end];
        const u32 len = hi - lo;
        const u32 c_lo = rowptr[r0], c_len = rowptr[r_end] - c_lo;
        const u32 out0 = (u32)indptr[live ? r : r_end];
        const bool staged = c_len <= RF_SMEM_ENT;
        const E* a;
        if (staged) {
            stage_chunk<E>(s_ent, entries + c_lo, c_len);
            __syncthreads();
            a = s_ent + (lo - c_lo);
        } else {
            a = entries + lo;
        }
        const u32 wmax = __reduce_max_sync(0xffffffffu, len);
        if (wmax <= 8) {
            row_emit<8, T, ENT>(a, len, sym, w_emit, w_typed, out0, indices, data);
        } else if (wmax <= 16) {
            row_emit<16, T, ENT>(a, len, sym, w_emit, w_typed, out0, indices, data);
        } else {
            walk_row<T, ENT>(a, len, sym, w_emit, w_typed, [&](u32 k, u32 minor, T v) {
                indices[out0 + k] = (int32_t)minor;
                data[out0 + k] = v;
            });
        }
        __syncthreads();  // s_ent is reused by the next chunk
    }
}

// dynamic shared memory: cursors[SB_ROWS_MAX] (u32) | staged entries[sb.cap].  SB_PT threads per CTA and SB_PB pairs per
// thread in flight: a sub-bucket is a few thousand entries, so its three phases (row pointers, group, copy out) are one
// or two memory round trips each -- with one pair per thread and iteration the phases were a chain of 24 dependent round
// trips (1.8 ms on C4d, profiles/r2b_buckets.md).
#define SB_PT 512
#define SB_PB 4
// the row sorted in place by a network in registers; returns the number of distinct minors
template <int N, class ENT>
__device__ __forceinline__ u32 row_sort_heads(typename ENT::type* a, u32 len)
{
    typename ENT::type v[N];
    load_row<N, ENT>(v, a, len);
    sort_network(v);
    u32 heads = 0;
#pragma unroll
    for (int k = 0; k < N; k++) {
        if ((u32)k < len) {
            a[k] = v[k];
            heads += (k == 0 || ENT::minor(v[k]) != ENT::minor(v[k > 0 ? k - 1 : 0])) ? 1u : 0u;
        }
    }
    return heads;
}
// `ucnt` non-NULL: every row of <= RS_SMALL entries is also SORTED while its sub-bucket sits in shared memory and the number
// of entries it will store (distinct minors) is written to ucnt -- k_rows_sort's whole pass over the entries (read, sort,
// write back) disappears; the rows left over are marked SB_UNSORTED for k_rows_sort_rest.  (Not for weighted max(S, S^T)
// builds, whose count depends on the summed weights.)
template <class ENT>
__global__ void __launch_bounds__(SB_PT) k_sub_rows_place(const u32* __restrict__ pair_major, const typename ENT::type* __restrict__ pair_ent,
                                                           const DevSizes* __restrict__ ds, const SubBuckets sb, const u32* __restrict__ sub_off,
                                                           const u32* __restrict__ rowptr, u32* __restrict__ cursor, typename ENT::type* __restrict__ entries,
                                                           u32* __restrict__ ucnt)
{
    typedef typename ENT::type EV;
    extern __shared__ __align__(16) uint8_t s_dyn[];
    u32* s_cur = reinterpret_cast<u32*>(s_dyn);
    EV* s_out = reinterpret_cast<EV*>(s_dyn + SB_ROWS_MAX * sizeof(u32));
    if (!ds->ok) return;
    const u32 rows = ds->rows, per = 1u << sb.shift2;
    for (u32 g = blockIdx.x; g < sb.n_sub; g += gridDim.x) {  // uniform per CTA
        const u32 row0 = g << sb.shift2;
        if (row0 >= rows) break;
        const u32 nr = min(per, rows - row0);
        const u32 lo = sub_off[g], hi = sub_off[g + 1], n = hi - lo;
        if (n == 0 || n > sb.cap) {
            // does not fit: global cursors (cursor[] starts as a copy of rowptr[]); nothing is sorted here
            for (u32 i = lo + threadIdx.x; i < hi; i += SB_PT) entries[atomicAdd(&cursor[pair_major[i]], 1u)] = pair_ent[i];
            if (ucnt)
                for (u32 i = threadIdx.x; i < nr; i += SB_PT) ucnt[row0 + i] = n ? SB_UNSORTED : 0u;
            continue;
        }
        const u32 out0 = rowptr[row0];
        for (u32 i = threadIdx.x; i < nr; i += SB_PT) s_cur[i] = rowptr[row0 + i] - out0;
        __syncthreads();
        for (u32 i0 = lo + threadIdx.x; i0 < hi; i0 += SB_PT * SB_PB) {
            u32 r[SB_PB];
            EV v[SB_PB];
#pragma unroll
            for (int u = 0; u < SB_PB; u++) {
                const u32 i = i0 + u * SB_PT;
                r[u] = 0xFFFFFFFFu;
                v[u] = (EV)0;
                if (i < hi) { r[u] = pair_major[i] - row0; v[u] = pair_ent[i]; }
            }
#pragma unroll
            for (int u = 0; u < SB_PB; u++) {
                if (r[u] != 0xFFFFFFFFu) {
                    G2N_CHECK(r[u] < nr);
                    const u32 p = atomicAdd(&s_cur[r[u]], 1u);
                    G2N_CHECK(p < n);
                    if (p < n) s_out[p] = v[u];
                }
            }
        }
        __syncthreads();
        if (ucnt) {  // s_cur[r] is now the END of row r's range; per warp the longest of its 32 rows picks the path
            for (u32 r0 = 0; r0 < nr; r0 += SB_PT) {  // uniform trip count: the warp reduction needs every lane
                const u32 r = r0 + threadIdx.x;
                const bool live = r < nr;
                const u32 a0 = (live && r) ? s_cur[r - 1] : 0u, len = live ? s_cur[r] - a0 : 0u;
                EV* a = s_out + a0;
                const u32 wmax = __reduce_max_sync(0xffffffffu, len);
                u32 heads;
                if (wmax <= 8) heads = row_sort_heads<8, ENT>(a, len);
                else if (wmax <= 16) heads = row_sort_heads<16, ENT>(a, len);
                else if (len <= RS_SMALL) {
                    if (len > 1) insertion_sort<EV>(a, len);
                    heads = 0;
                    for (u32 k = 0; k < len; k++) heads += (k == 0 || ENT::minor(a[k]) != ENT::minor(a[k - 1])) ? 1u : 0u;
                } else heads = SB_UNSORTED;
                if (live) ucnt[row0 + r] = heads;
            }
            __syncthreads();
        }
        for (u32 i = threadIdx.x; i < n; i += SB_PT) entries[(u64)out0 + i] = s_out[i];
        __syncthreads();
    }
}

}  // namespace g2n
