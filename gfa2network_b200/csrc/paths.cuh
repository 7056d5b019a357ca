// paths.cuh -- SURVEY 8(f) row 4: the node lists of the P / O records (analysis.py:164-177 over
// parser.py:229-247, 343-361), resolved to node IDs on the device.
//
// The records are a handful of lines, but each can be hundreds of megabytes (one haplotype walking every
// segment), so nothing per entry happens on the host:
//   k_paths_find     record lines whose first field is P or O               -> line offsets (few)
//   k_path_fields    one CTA per record: end of the name field and of the segment-list field
//   k_path_count     commas per 4 KiB block of every list                    -> exclusive scan = entry index
//   k_path_lookup    every entry "<segment>[+-]": strip one trailing sign, look the name up in the build's
//                    table (read-only probe), node ID (or -1) to ids[entry index]; the first unknown name
//                    of a record is remembered (the reference raises NodeNotFound for it)
#pragma once
#include "table.cuh"

namespace g2n {

struct PathRec {
    u64 line;       // offset of the record's first byte
    u64 name_off;   // fields[1]
    u64 list_off;   // fields[2]
    u64 list_end;   // one past its last byte
    u64 first_blk;  // index of the record's first 4 KiB block in the concatenated block list
    u64 missing;    // smallest entry index (within the record) whose name is not a node; ~0 if none
    u32 name_len;
    u32 n_fields_ok;  // 1: the record has >= 3 fields
};

#define PB_BYTES 4096  // bytes of a list handled by one CTA iteration

__global__ void __launch_bounds__(256) k_paths_find(const uint8_t* __restrict__ text, u64 n, u64* __restrict__ starts, u32 cap, u32* __restrict__ count)
{
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i * 16 < n; i += (u64)gridDim.x * blockDim.x) {
        const u64 base = i * 16;
        __align__(16) uint8_t bb[32];
        uint8_t* b = bb + 15;  // b[0] = the byte before my 16, b[1..16] = mine
        if (base + 16 <= n && base > 0) {
            *reinterpret_cast<uint4*>(bb + 16) = ld_nc_v4(text + base);  // the text is 16-byte aligned
            b[0] = text[base - 1];
        } else {
            b[0] = base ? text[base - 1] : (uint8_t)'\n';
            for (int k = 0; k < 16; k++) b[1 + k] = base + k < n ? text[base + k] : (uint8_t)'\n';
        }
        for (int k = 0; k < 16; k++) {
            if (b[k] != '\n' || (b[k + 1] != 'P' && b[k + 1] != 'O')) continue;
            const u64 p = base + k;
            if (p >= n) continue;
            const uint8_t nx = p + 1 < n ? text[p + 1] : (uint8_t)'\n';
            if (nx != '\t') continue;  // the first FIELD must be exactly "P" / "O" (parser.py:133-134)
            const u32 j = atomicAdd(count, 1u);
            if (j < cap) starts[j] = p;
        }
    }
}

// first byte equal to '\t' or '\n' at or after `from` (n if none): all threads of the CTA, 4 KiB per round
__device__ __forceinline__ u64 cta_find_sep(const uint8_t* text, u64 n, u64 from, u64* s_min)
{
    u64 pos = from;
    while (true) {
        u64 mine = ~0ull;
        const u64 a = pos + (u64)threadIdx.x * 16;
        for (int k = 0; k < 16; k++) {
            const u64 p = a + k;
            if (p >= n) { mine = mine < n ? mine : n; break; }
            const uint8_t c = text[p];
            if (c == '\t' || c == '\n') { mine = p; break; }
        }
        if (threadIdx.x == 0) *s_min = ~0ull;
        __syncthreads();
        if (mine != ~0ull) atomicMin(s_min, mine);
        __syncthreads();
        const u64 r = *s_min;
        __syncthreads();
        if (r != ~0ull) return r;
        pos += 256 * 16;
    }
}

__global__ void __launch_bounds__(256) k_path_fields(const uint8_t* __restrict__ text, u64 n, PathRec* __restrict__ recs, u32 n_recs)
{
    __shared__ u64 s_min;
    for (u32 r = blockIdx.x; r < n_recs; r += gridDim.x) {
        const u64 line = recs[r].line;
        const u64 name_off = line + 2;
        const u64 e1 = cta_find_sep(text, n, name_off, &s_min);  // end of fields[1]
        u64 list_off = e1, list_end = e1;
        u32 ok = 0;
        if (e1 < n && text[e1] == '\t') {
            ok = 1;
            list_off = e1 + 1;
            list_end = cta_find_sep(text, n, list_off, &s_min);
        }
        if (threadIdx.x == 0) {
            PathRec R = recs[r];
            R.name_off = name_off; R.name_len = (u32)(e1 - name_off); R.list_off = list_off; R.list_end = list_end;
            R.n_fields_ok = ok; R.missing = ~0ull;
            recs[r] = R;
        }
        __syncthreads();
    }
}

// block b of the concatenated lists -> (record, byte range)
__device__ __forceinline__ bool path_block(const PathRec* recs, u32 n_recs, u64 b, u32& r, u64& lo, u64& hi)
{
    u32 a = 0, z = n_recs;  // last record whose first_blk <= b
    while (z - a > 1) {
        const u32 m = (a + z) >> 1;
        if (recs[m].first_blk <= b) a = m; else z = m;
    }
    r = a;
    // entry starts live in [list_off, list_end] INCLUSIVE: "a+,b+," ends with an empty entry, an empty field is one
    lo = recs[a].list_off + (b - recs[a].first_blk) * PB_BYTES;
    hi = lo + PB_BYTES < recs[a].list_end + 1 ? lo + PB_BYTES : recs[a].list_end + 1;
    return true;
}

// entries that START in the block: one at list_off itself, one after every ','
__global__ void __launch_bounds__(256) k_path_count(const uint8_t* __restrict__ text, const PathRec* __restrict__ recs, u32 n_recs, u64 n_blocks,
                                                     u32* __restrict__ blk_cnt)
{
    __shared__ u32 s_cnt;
    for (u64 b = blockIdx.x; b < n_blocks; b += gridDim.x) {
        u32 r;
        u64 lo, hi;
        path_block(recs, n_recs, b, r, lo, hi);
        if (threadIdx.x == 0) s_cnt = 0;
        __syncthreads();
        u32 c = 0;
        const u64 a = lo + (u64)threadIdx.x * 16;
        for (int k = 0; k < 16; k++) {
            const u64 p = a + k;
            if (p >= hi) break;
            c += (p == recs[r].list_off || text[p - 1] == ',') ? 1u : 0u;
        }
        if (c) atomicAdd(&s_cnt, c);
        __syncthreads();
        if (threadIdx.x == 0) blk_cnt[b] = s_cnt;
        __syncthreads();
    }
}

struct PathLookup {
    const Slot* slots;
    const LongDesc* longs;
    const u32* slot_id;
    u32 mask;
    u64 seed;
    int bidirected;
};

// node ID of the name text[a, a + len) or -1
__device__ __forceinline__ int32_t path_find_node(const PathLookup& T, const uint8_t* text, u64 a, u32 len)
{
    u64 k0 = 0, k1 = 0;
    const bool is_long = len > 15;
    if (!is_long) {
        for (u32 i = 0; i < len; i++) {
            const u64 c = text[a + i];
            if (i < 8) k0 |= c << (8 * i); else k1 |= c << (8 * (i - 8));
        }
        k1 |= (u64)(len + 1) << 56;
    } else {  // table.cuh: make_key
        u64 h1 = T.seed ^ 0x9e3779b97f4a7c15ULL, h2 = ~T.seed * 0xd6e8feb86659fd93ULL;
        for (u32 i = 0; i < len; i++) {
            const u64 c = text[a + i];
            h1 = (h1 ^ c) * 0x100000001b3ULL;
            h2 = (h2 + c + 1) * 0xc2b2ae3d27d4eb4fULL;
            h2 ^= h2 >> 29;
        }
        k0 = mix64(h1 ^ (h2 << 1));
        k1 = (0xFFull << 56) | ((u64)(len & 0xFFFFFF) << 32) | (mix64(h2 + h1) & 0xFFFFFFFFull);
    }
    ProbeSeq q = probe_seq(k0, k1, T.mask, T.bidirected);
    u32 visited = 0;
    while (true) {
        const u32 j = q.slot();
        const Slot x = T.slots[j];
        if (x.k0 == k0 && x.k1 == k1) {
            const u32 found = j;
            if (is_long) {  // same hash: compare the bytes kept for the slot
                const LongDesc d = T.longs[x.rep - 1];
                if (d.base_len + (d.has_ori ? 1 + d.ori_len : 0) != len) return -1;
                for (u32 i = 0; i < len; i++)
                    if (long_byte(text, d, i) != text[a + i]) return -1;
            }
            return (int32_t)T.slot_id[found];
        }
        if (x.k0 == 0 && x.k1 == 0) return -1;
        if (!q.next(T.mask, visited)) return -1;
    }
}

__global__ void __launch_bounds__(256) k_path_lookup(const uint8_t* __restrict__ text, PathRec* __restrict__ recs, u32 n_recs, u64 n_blocks,
                                                      const u64* __restrict__ blk_off, const PathLookup T, int32_t* __restrict__ ids)
{
    __shared__ u32 s_warp[8];
    for (u64 b = blockIdx.x; b < n_blocks; b += gridDim.x) {
        u32 r;
        u64 lo, hi;
        path_block(recs, n_recs, b, r, lo, hi);
        const u64 list_off = recs[r].list_off, list_end = recs[r].list_end;
        // my 16 bytes: entry starts and their rank inside the block
        const u64 a = lo + (u64)threadIdx.x * 16;
        u32 c = 0;
        for (int k = 0; k < 16; k++) {
            const u64 p = a + k;
            if (p >= hi) break;
            c += (p == list_off || text[p - 1] == ',') ? 1u : 0u;
        }
        const u32 lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
        const u32 inc = warp_incl_scan(c);
        if (lane == 31) s_warp[wid] = inc;
        __syncthreads();
        u32 base = 0;
        for (u32 w = 0; w < wid; w++) base += s_warp[w];
        u64 idx = blk_off[b] + base + inc - c;  // index (over all records) of my first entry
        for (int k = 0; k < 16; k++) {
            const u64 p = a + k;
            if (p >= hi) break;
            if (!(p == list_off || text[p - 1] == ',')) continue;
            u64 e = p;  // end of the entry: next ',' or the end of the field
            while (e < list_end && text[e] != ',') e++;
            u64 len = e - p;
            if (len && (text[e - 1] == '+' || text[e - 1] == '-')) len--;  // parser.py:237-244: ONE trailing sign
            const int32_t id = path_find_node(T, text, p, (u32)len);
            ids[idx] = id;
            if (id < 0) atomicMin((unsigned long long*)&recs[r].missing, (unsigned long long)(idx - blk_off[recs[r].first_blk]));
            idx++;
        }
        __syncthreads();
    }
}

}  // namespace g2n
