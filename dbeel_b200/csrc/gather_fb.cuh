// gather_fb.cuh -- K5 with the entry-boundary blocks handled INSIDE the copy loop (round 2).
//
// ncu of k_gather32 (profiles/r02_v2_ncu_full.csv): its busiest unit is the L1 data pipe (70 % of peak in LSU wavefronts), and
// the copy proper is only about a third of those wavefronts.  The rest comes from the passes that run one lane per ENTRY: a
// warp instruction whose 27 lanes touch 27 different 128-byte lines costs 27 wavefronts, where the copy loop's instructions
// (32 lanes on consecutive addresses) cost 8-9.  The dense boundary pass of k_gather32 -- for every entry, the 32-byte block
// that holds its last byte: ~6 scattered loads + 2 scattered 16-byte stores -- is the largest of them.
//
// Here the lane that owns such a block in the copy loop builds it itself.  Entries are >= 32 bytes, so a 32-byte block holds
// at most one boundary: t bytes of entry e, then 32 - t bytes of entry e + 1.  The lane loads TWO windows:
//   * the block in e's coordinates   -- addresses that continue its left neighbour's, same lines, coalesced with them;
//   * the block in e+1's coordinates -- a second set of three loads that only the ~3 boundary lanes of a chunk execute;
// pieces that hold no byte of their entry are not loaded (the window may reach past the end of e / before the start of e+1,
// i.e. outside the run's buffer for its last / first entry).  Both windows are realigned, blended at byte t and leave in the
// same coalesced 256-bit store as the pure blocks.  What remains of the dense pass is the last, partial block of the whole
// stream (byte stores, once per job).  256 threads per CTA and one 1 KB chunk per warp keep the register budget of
// k_gather32 (two windows of three pieces instead of two chunks of one window).
#pragma once
#include "kernels.cuh"

namespace dbeel {

constexpr int kFbThreads = 256;
#ifndef DBEEL_GATHER_FB_MINB
#define DBEEL_GATHER_FB_MINB 6
#endif
static_assert(kGatherTileBytes == 32ull * kFbThreads, "one 32-byte block per thread and tile");

__device__ __forceinline__ uint4 ldg_if(bool need, const uint4 *q) { return need ? __ldg(q) : make_uint4(0, 0, 0, 0); }

__global__ void __launch_bounds__(kFbThreads, DBEEL_GATHER_FB_MINB) k_gather_fb(Params p) {
    pdl_trigger();
    pdl_wait();
    constexpr int NT = kFbThreads;
    __shared__ unsigned long long s_adj[kGatherMaxEntries]; // entry address minus its tile-relative start
    __shared__ int s_r0[kGatherMaxEntries], s_r1[kGatherMaxEntries];
    __shared__ uint32_t s_ks[kGatherMaxEntries];
    const Ctl *c = p.ctl;
    const unsigned long long out_len = c->out_data_len;
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t tile_id = blockIdx.x;
    const unsigned long long T0 = (unsigned long long)tile_id * kGatherTileBytes;
    if (T0 >= out_len) return;
    const uint32_t tile_len = out_len - T0 < kGatherTileBytes ? (uint32_t)(out_len - T0) : (uint32_t)kGatherTileBytes;
    const uint32_t e_lo = p.tile_first[tile_id];
    const uint32_t e_hi = T0 + kGatherTileBytes < out_len ? p.tile_first[tile_id + 1] : c->out_items - 1;
    const uint32_t ne = e_hi - e_lo + 1; // <= kGatherMaxEntries: every entry is >= 32 bytes
    const bool hash_here = p.bloom.words != nullptr && p.hash_rec == nullptr && !p.bloom_elsewhere;
    for (uint32_t j = tid; j < ne; j += NT) {
        const uint4 rec = p.out_index[e_lo + j];
        const unsigned long long d0 = ((unsigned long long)rec.x | ((unsigned long long)rec.y << 32)) - p.out_offset_base;
        const long long r0 = (long long)d0 - (long long)T0; // < 0 only for the tile's first entry
        const long long r1 = r0 + (long long)rec.w;
        s_adj[j] = p.src_ptr[e_lo + j] - (unsigned long long)r0;
        s_r0[j] = r0 < -0x7FFFFFFFll ? -0x7FFFFFFF : (int)r0;
        s_r1[j] = r1 > 0x7FFFFFFFll ? 0x7FFFFFFF : (int)r1;
        if (hash_here) s_ks[j] = rec.z;
    }
    __syncthreads();

    // ---- copy: warp w owns bytes [w * 1 KB, (w + 1) * 1 KB) of the tile, one 32-byte block per lane
    uint8_t *dst_tile = p.out_data + T0;
    const int cb = (int)(warp * 1024u);
    if ((uint32_t)cb < tile_len) {
        uint32_t j = 0; // the entry that holds byte cb = number of entries ending at or before it (ends ascend)
        for (uint32_t base = 0; base + 1 < ne; base += 32) {
            const uint32_t i = base + lane;
            j += __popc(__ballot_sync(0xFFFFFFFFu, i + 1 < ne && s_r1[i] <= cb));
        }
        const uint32_t lanes_le = 0xFFFFFFFFu >> (31 - lane); // bits 0..lane
        const int b0 = cb + (int)lane * 32;
        // Entries that end inside the chunk, i.e. in (cb, cb + 1024]: at most 32, lane l looks at entry j + l.  An end at r1
        // precedes the blocks t = ceil((r1 - cb) / 32) .. 31; distinct entries have distinct t.
        const uint32_t i = j + lane;
        const int r1 = i + 1 < ne ? s_r1[i] : 0x7FFFFFFF;
        const bool ends_here = r1 <= cb + 1024;
        const uint32_t tq = (uint32_t)((ends_here ? r1 : cb + 32) - cb + 31) >> 5; // 1..32 when ends_here
        const uint32_t ends = __reduce_or_sync(0xFFFFFFFFu, (ends_here && tq < 32) ? (1u << tq) : 0u);
        const uint32_t e = j + __popc(ends & lanes_le); // entry that holds byte b0
        const int r1e = s_r1[e];
        const bool whole = (uint32_t)b0 + 32 <= tile_len; // the block lies inside the stream
        const bool pure = whole && b0 + 32 <= r1e;
        const bool edge = whole && !pure && e + 1 < ne;   // e ends inside the block and e + 1 fills the rest of it
        const uint32_t t = edge ? (uint32_t)(r1e - b0) : 32u; // bytes of e in the block: 1..31 on an edge
        // window 1: the block in e's coordinates (pieces that hold no byte of e stay unloaded)
        const uintptr_t sa = (uintptr_t)(s_adj[e] + (unsigned long long)(long long)b0);
        const uint32_t sh = (uint32_t)(sa & 15);
        const uint4 *sv = reinterpret_cast<const uint4 *>(sa - sh);
        const bool any = pure || edge;
        const uint4 A = ldg_if(any, sv);
        const uint4 B = ldg_if(any && 16u - sh < t, sv + 1);
        const uint4 C = ldg_if(any && sh != 0 && 32u - sh < t, sv + 2);
        // window 2 (edge blocks only): the block in e+1's coordinates, pieces that end at or before byte t unloaded
        uint4 D = make_uint4(0, 0, 0, 0), E = D, F = D;
        uint32_t sh2 = 0;
        if (edge) {
            const uintptr_t sa2 = (uintptr_t)(s_adj[e + 1] + (unsigned long long)(long long)b0);
            sh2 = (uint32_t)(sa2 & 15);
            const uint4 *sv2 = reinterpret_cast<const uint4 *>(sa2 - sh2);
            D = ldg_if(16u - sh2 > t, sv2);
            E = ldg_if(32u - sh2 > t, sv2 + 1);
            F = ldg_if(sh2 != 0, sv2 + 2); // bytes [32 - sh2, 32) of the block: always part of e + 1 on an edge
        }
        if (any) {
            uint32_t o[8];
            realign32(A, B, C, sh, o);
            if (edge) {
                uint32_t o2[8];
                realign32(D, E, F, sh2, o2);
#pragma unroll
                for (uint32_t w = 0; w < 8; w++) { // bytes [0, t) from e, [t, 32) from e + 1
                    const uint32_t lo = 4 * w;
                    const uint32_t keep = t >= lo + 4 ? 0xFFFFFFFFu : (t <= lo ? 0u : (0xFFFFFFFFu >> (32 - 8 * (t - lo))));
                    o[w] = (o[w] & keep) | (o2[w] & ~keep);
                }
            }
            stg256(dst_tile + b0, o);
        }
    }

    // ---- the ragged end of the whole stream: the last entry's bytes in the last, partial 32-byte block (once per job)
    if ((tile_len & 31u) != 0 && tid == 0) {
        const uint32_t j = ne - 1;
        const uint32_t b_start = tile_len & ~31u;
        const int r0 = s_r0[j];
        // the partial block belongs to the last entry alone (an entry is >= 32 bytes and ends at tile_len) unless that
        // entry starts inside it, which would make it shorter than the block
        const uint32_t from = r0 > (int)b_start ? (uint32_t)r0 : b_start;
        const uint8_t *src = reinterpret_cast<const uint8_t *>((uintptr_t)(s_adj[j] + from));
        for (uint32_t b = from; b < tile_len; b++) dst_tile[b] = __ldg(src + (b - from));
        if (from > b_start) { // bytes of the entry before it (cannot happen with entries >= 32 bytes; kept for safety)
            const uint8_t *src2 = reinterpret_cast<const uint8_t *>((uintptr_t)(s_adj[j - 1] + b_start));
            for (uint32_t b = b_start; b < from; b++) dst_tile[b] = __ldg(src2 + (b - b_start));
        }
    }

    // ---- bloom (fused epilogue): entries whose first byte lies in this tile
    if (hash_here) {
        for (uint32_t j = tid; j < ne; j += NT) {
            const int r0 = s_r0[j];
            if (r0 < 0 || r0 >= (int)kGatherTileBytes) continue;
            const uint8_t *key = reinterpret_cast<const uint8_t *>((uintptr_t)(s_adj[j] + (unsigned long long)r0)) + 8;
            const uint64_t klen = s_ks[j] - 8;
            uint64_t h0, h1;
            sip13_pair_vec_u8(p.bloom.sip, klen, [key](uint64_t q) { return ld_u64_unaligned(key + 8 * q); }, &h0, &h1);
            uint32_t *words = p.bloom.words;
            bloom_probe_all(h0, h1, p.bloom.k_num, p.bloom.bits, p.bloom.bits_magic,
                            [words](uint64_t bit) { atomicOr(&words[bit >> 5], 1u << (bit & 31)); });
        }
    }
}

} // namespace dbeel
