// wal.cuh -- row N4 of the scope table: the step before the flush.  LSMTree::read_memtable_from_wal_file
// (lsm_tree.rs:552-574) replays a write-ahead log into a memtable; open_or_create_ex then flushes it (:478-513).
//
// The log is a chain: a record starts on a 4096-byte boundary, and the NEXT record starts on the first boundary strictly
// after this one's last byte (:568-571), so where record i+1 starts is known only after record i has been decoded -- a
// serial dependency through the whole file.  Here it is cut the list-ranking way:
//   k_wal_parse   every page decodes "the record that would start here" on its own: where the chain would go next
//                 (jump[0][page]) and whether that record would be inserted (cnt[0][page]);
//   k_wal_double  pointer doubling: jump[k+1][p] = jump[k][jump[k][p]], cnt[k+1][p] = cnt[k][p] + cnt[k][jump[k][p]];
//                 after log2(pages) rounds cnt[K][0] is the number of replayed records;
//   k_wal_select  thread i walks the doubling tables from page 0 to the i-th replayed record and writes an index record
//                 {offset into the log, key_size, full_size} -- pages inside a multi-page record, padding pages and
//                 whatever follows a torn record are simply never reached.
// The result is an arrival batch whose .data IS the log (index offsets are sparse): the flush front end sorts it by
// (key, arrival), the last arrival of a key wins, and the SSTable comes out as for any other flush.
#pragma once

#include "kernels.cuh"

namespace dbeel {

constexpr uint32_t kWalPage = 4096;
constexpr uint32_t kWalTooLarge = 1u; // flags: a replayed record does not fit EntryOffset's u32 sizes (entry_writer.rs:72-74)

struct WalParams {
    const uint8_t *wal;
    uint64_t len;
    uint32_t n_pages; // ceil(len / 4096); node n_pages is the end of the chain
    uint32_t levels;  // K: 2^K >= n_pages + 1
    uint32_t *jump;   // [levels + 1][n_pages + 1]
    uint32_t *cnt;    // [levels + 1][n_pages + 1]
    uint2 *sizes;     // [n_pages] {key_size, full_size} of the record that would start at each page
    uint4 *index;     // out: one record per replayed entry, arrival order
    unsigned long long *totals; // [0] = replayed records, [1] = sum of their full_size, [2] = flags
};

__device__ __forceinline__ uint64_t wal_ld64(const uint8_t *p) { return ld_u64_unaligned(p); }

__global__ void __launch_bounds__(256) k_wal_parse(WalParams w) {
    pdl_trigger();
    pdl_wait();
    const uint32_t pg = blockIdx.x * 256u + threadIdx.x;
    const uint32_t stride = w.n_pages + 1;
    if (pg > w.n_pages) return;
    if (pg == w.n_pages) { // the end node absorbs
        w.jump[pg] = pg;
        w.cnt[pg] = 0;
        return;
    }
    (void)stride;
    const uint64_t len = w.len, pos = (uint64_t)pg * kWalPage;
    uint32_t next = w.n_pages, take = 0;
    uint2 sz = make_uint2(0, 0);
    // the decode of bincode's Entry from a Cursor: every read that runs out of bytes ends the replay (lsm_tree.rs:561-572)
    do {
        const uint64_t left = len - pos;
        if (left < 8) break;
        const uint64_t klen = wal_ld64(w.wal + pos);
        if (klen > left - 8) break;
        const uint64_t p2 = pos + 8 + klen;
        if (len - p2 < 8) break;
        const uint64_t dlen = wal_ld64(w.wal + p2);
        if (dlen > len - p2 - 8) break;
        const uint64_t p3 = p2 + 8 + dlen;
        if (len - p3 < 16) break;
        const uint64_t end = p3 + 16;
        const uint64_t nx = end / kWalPage + 1; // pos + PAGE_SIZE - pos % PAGE_SIZE
        next = nx * kWalPage < len ? (uint32_t)nx : w.n_pages;
        // an Entry whose timestamp does not deserialize is skipped, but its bytes were consumed (the chain goes on)
        take = ts_decodes(wal_ld64(w.wal + p3), wal_ld64(w.wal + p3 + 8)) ? 1u : 0u;
        const uint64_t ks = klen + 8, fs = klen + dlen + 32;
        if (fs > 0xFFFFFFFFull) sz = make_uint2(0xFFFFFFFFu, 0xFFFFFFFFu); // reported only if the chain reaches it
        else sz = make_uint2((uint32_t)ks, (uint32_t)fs);
    } while (false);
    w.jump[pg] = next;
    w.cnt[pg] = take;
    w.sizes[pg] = sz;
}

__global__ void __launch_bounds__(256) k_wal_double(WalParams w, uint32_t k) {
    pdl_trigger();
    pdl_wait();
    const uint32_t pg = blockIdx.x * 256u + threadIdx.x;
    const uint32_t stride = w.n_pages + 1;
    if (pg >= stride) return;
    const uint32_t *j0 = w.jump + (uint64_t)k * stride, *c0 = w.cnt + (uint64_t)k * stride;
    const uint32_t mid = j0[pg];
    w.jump[(uint64_t)(k + 1) * stride + pg] = j0[mid];
    w.cnt[(uint64_t)(k + 1) * stride + pg] = c0[pg] + c0[mid];
}

__global__ void __launch_bounds__(256) k_wal_select(WalParams w) {
    pdl_trigger();
    pdl_wait();
    const uint32_t stride = w.n_pages + 1;
    const uint32_t total = w.cnt[(uint64_t)w.levels * stride]; // records replayed from page 0 on
    const uint32_t i = blockIdx.x * 256u + threadIdx.x;
    unsigned long long bytes = 0;
    uint32_t bad = 0;
    if (i < total) {
        uint32_t node = 0, rem = i;
        for (int k = (int)w.levels; k >= 0; k--) { // skip whole 2^k-node stretches that hold <= rem replayed records
            const uint32_t c = w.cnt[(uint64_t)k * stride + node];
            if (c <= rem) {
                rem -= c;
                node = w.jump[(uint64_t)k * stride + node];
            }
        }
        const uint2 sz = w.sizes[node];
        const unsigned long long off = (unsigned long long)node * kWalPage;
        if (sz.y == 0xFFFFFFFFu && sz.x == 0xFFFFFFFFu) bad = kWalTooLarge;
        w.index[i] = make_uint4((uint32_t)off, (uint32_t)(off >> 32), sz.x, sz.y);
        bytes = bad ? 0ull : sz.y;
    }
    // job totals: one atomic per warp
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        bytes += __shfl_xor_sync(0xFFFFFFFFu, bytes, o);
        bad |= __shfl_xor_sync(0xFFFFFFFFu, bad, o);
    }
    if ((threadIdx.x & 31) == 0) {
        if (bytes) atomicAdd(&w.totals[1], bytes);
        if (bad) atomicOr(&w.totals[2], (unsigned long long)bad);
    }
    if (i == 0) w.totals[0] = total;
}

} // namespace dbeel
