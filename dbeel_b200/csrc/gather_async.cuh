// gather_async.cuh -- K5 with the payload landing in SHARED MEMORY instead of registers (round 2).
//
// k_gather32 keeps a CTA's payload in flight in registers: three 16-byte loads per 32-byte block, two blocks per lane, 24
// registers that stay allocated from the moment the loads are issued until the realigned block is stored -- which caps the SM
// at 12 CTAs of 128 threads (40 registers) and makes every warp sit on its loads before it can do anything else.  Here the
// same three aligned 16-byte pieces of every block are fetched with cp.async (LDGSTS, 16 bytes, global -> shared, no
// register destination) into a slot that belongs to the lane: 48 bytes per block, lanes 48 bytes apart (a quarter-warp's
// 16-byte accesses at a 48-byte stride cover all 32 banks exactly once).  Consequences:
//   * no payload registers: the kernel fits 32 registers, 16 CTAs of 128 threads per SM instead of 12 (+33 % bytes in flight);
//   * the loads are fire-and-forget, so between issuing them and waiting for them the warp does the work that used to queue
//     up BEHIND the copy: the boundary blocks of its entries and the filter (key loads, 2 x SipHash-1-3, 7 atomicOr);
//   * each lane reads back only its own slot (three LDS.128), so no barrier is needed after the wait.
// Everything else -- tiles, staging, the vector -> entry map, the boundary pass, the bytes written -- is k_gather32's.
#pragma once
#include "kernels.cuh"

namespace dbeel {

#ifndef DBEEL_GATHER_ASYNC_MINB
#define DBEEL_GATHER_ASYNC_MINB 16
#endif

__device__ __forceinline__ void cp_async16_ca(void *smem_dst, const void *gmem_src) { // through L1: neighbouring lanes share pieces
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src) : "memory");
}

__global__ void __launch_bounds__(kGatherThreads, DBEEL_GATHER_ASYNC_MINB) k_gather_async(Params p) {
    pdl_trigger();
    pdl_wait();
    constexpr int NT = kGatherThreads;
    __shared__ unsigned long long s_adj[kGatherMaxEntries]; // entry address minus its tile-relative start
    __shared__ int s_r1[kGatherMaxEntries]; // tile-relative end of every entry (= start of the next one)
    __shared__ int s_r0_first;
    __shared__ __align__(16) uint4 s_land[kG32Vpt][NT][3]; // landing slots: [chunk][thread][16-byte piece]
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
        if (j == 0) s_r0_first = r0 < -0x7FFFFFFFll ? -0x7FFFFFFF : (int)r0;
        s_r1[j] = r1 > 0x7FFFFFFFll ? 0x7FFFFFFF : (int)r1;
    }
    __syncthreads();

    // ---- issue: warp w owns bytes [w * 2 KB, (w + 1) * 2 KB) of the tile, 1 KB (32 lanes x 32 bytes) at a time
    uint8_t *dst_tile = p.out_data + T0;
    const int sub0 = (int)(warp * (uint32_t)(kG32Vpt * 1024));
    const bool copying = (uint32_t)sub0 < tile_len;
    uint32_t shp = 0; // per chunk: bits 0-3 = shift of the block inside its 48-byte window, bit 4 = the block is stored here
    if (copying) {
        uint32_t j = 0; // the entry that holds byte sub0 = number of entries ending at or before it (ends ascend)
        for (uint32_t base = 0; base + 1 < ne; base += 32) {
            const uint32_t i = base + lane;
            j += __popc(__ballot_sync(0xFFFFFFFFu, i + 1 < ne && s_r1[i] <= sub0));
        }
        const uint32_t lanes_le = 0xFFFFFFFFu >> (31 - lane); // bits 0..lane
#pragma unroll
        for (int k = 0; k < kG32Vpt; k++) {
            const int cb = sub0 + k * 1024;
            const int b0 = cb + (int)lane * 32;
            // Entries that end inside the chunk, i.e. in (cb, cb + 1024]: at most 32, lane l looks at entry j + l.  An end at
            // r1 precedes the blocks t = ceil((r1 - cb) / 32) .. 31; distinct entries have distinct t.
            const uint32_t i = j + lane;
            const int r1 = i + 1 < ne ? s_r1[i] : 0x7FFFFFFF;
            const bool ends_here = r1 <= cb + 1024;
            const uint32_t t = (uint32_t)((ends_here ? r1 : cb + 32) - cb + 31) >> 5; // 1..32 when ends_here
            const uint32_t ends = __reduce_or_sync(0xFFFFFFFFu, (ends_here && t < 32) ? (1u << t) : 0u);
            const uint32_t cnt = __popc(ends & lanes_le); // entries ending at or before my block's first byte
            const uint32_t adv = __popc(__ballot_sync(0xFFFFFFFFu, ends_here));
            const uint32_t e = j + cnt; // entry that holds byte b0
            j += adv;                   // entry that holds the next chunk's first byte
            // A block that is not wholly inside entry e (it holds e's end, or lies past the end of the stream) fetches the
            // last 32 bytes of e instead (always valid: entries are >= 32 bytes) and is not stored by this pass.
            const int r1e = s_r1[e];
            const bool pure = (uint32_t)b0 + 32 <= tile_len && b0 + 32 <= r1e;
            const int bl = b0 + 32 <= r1e ? b0 : r1e - 32;
            const uintptr_t sa = (uintptr_t)(s_adj[e] + (unsigned long long)(long long)bl);
            const uint32_t sh = (uint32_t)(sa & 15);
            const uint4 *sv = reinterpret_cast<const uint4 *>(sa - sh);
            cp_async16_ca(&s_land[k][tid][0], sv);
            cp_async16_ca(&s_land[k][tid][1], sv + 1);
            cp_async16_ca(&s_land[k][tid][2], sh ? sv + 2 : sv + 1);
            shp |= (sh | (pure ? 16u : 0u)) << (8 * k);
        }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");

    // ---- while the payload travels: the 32-byte block that holds the last byte of entry j (unless j ends on a block
    // boundary), its two halves
    for (uint32_t j = tid; j < ne; j += NT) {
        const int r1 = s_r1[j];
        if (r1 <= 0 || (r1 & 31) == 0 || r1 > (int)tile_len) continue;
        const bool has_next = j + 1 < ne;
#pragma unroll
        for (int half = 0; half < 2; half++) {
            const uint32_t b0 = ((uint32_t)r1 & ~31u) + 16u * half;
            if (b0 >= tile_len) continue;     // past the end of the stream
            const int t = r1 - (int)b0;       // bytes of entry j in this vector: <= 0 none, >= 16 all
            uint4 o;
            if (t >= 16) {
                o = ld16_any((uintptr_t)(s_adj[j] + b0), 16);
            } else if (t <= 0) {
                if (!has_next) continue;
                o = ld16_any((uintptr_t)(s_adj[j + 1] + b0), 16);
            } else {
                o = ld16_any((uintptr_t)(s_adj[j] + b0), (uint32_t)t);
                if (b0 + 16 <= tile_len) { // blend with the head of entry j + 1
                    const uint4 H = ld16_any((uintptr_t)(s_adj[j + 1] + (unsigned long long)(long long)r1), 16);
                    const uint4 HU = realign16_sel(make_uint4(0, 0, 0, 0), H, 16 - (uint32_t)t);
                    const uint32_t wfull = (uint32_t)t >> 2, bits = ((uint32_t)t & 3) * 8;
                    const uint32_t mmix = bits ? (0xFFFFFFFFu >> (32 - bits)) : 0u;
                    uint32_t ow[4] = {o.x, o.y, o.z, o.w}, hw[4] = {HU.x, HU.y, HU.z, HU.w};
#pragma unroll
                    for (uint32_t q = 0; q < 4; q++) {
                        const uint32_t mk = q < wfull ? 0xFFFFFFFFu : (q == wfull ? mmix : 0u);
                        ow[q] = (ow[q] & mk) | (hw[q] & ~mk);
                    }
                    o = make_uint4(ow[0], ow[1], ow[2], ow[3]);
                } else { // ragged end of the whole stream: never write past out_data_len
                    const uint32_t ow[4] = {o.x, o.y, o.z, o.w};
                    for (int b = 0; b < t; b++) dst_tile[b0 + b] = (uint8_t)(ow[b >> 2] >> ((b & 3) * 8));
                    continue;
                }
            }
            reinterpret_cast<uint4 *>(dst_tile)[b0 >> 4] = o;
        }
    }

    // ---- ... and the filter: entries whose first byte lies in this tile
    if (hash_here) {
        for (uint32_t j = tid; j < ne; j += NT) {
            const int r0 = j ? s_r1[j - 1] : s_r0_first; // entries abut
            if (r0 < 0 || r0 >= (int)kGatherTileBytes) continue;
            const uint8_t *key = reinterpret_cast<const uint8_t *>((uintptr_t)(s_adj[j] + (unsigned long long)r0)) + 8;
            const uint64_t klen = ld_u64_unaligned(key - 8); // the entry's own length prefix (validated against key_size by k_extract)
            uint64_t h0, h1;
            sip13_pair_vec_u8(p.bloom.sip, klen, [key](uint64_t q) { return ld_u64_unaligned(key + 8 * q); }, &h0, &h1);
            uint32_t *words = p.bloom.words;
            bloom_probe_all(h0, h1, p.bloom.k_num, p.bloom.bits, p.bloom.bits_magic,
                            [words](uint64_t bit) { atomicOr(&words[bit >> 5], 1u << (bit & 31)); });
        }
    }

    // ---- the payload has landed: realign from the lane's own slots, 256-bit stores
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    if (copying) {
#pragma unroll
        for (int k = 0; k < kG32Vpt; k++) {
            const uint32_t m = (shp >> (8 * k)) & 0xFF;
            if (m & 16u) {
                const uint4 A = s_land[k][tid][0], B = s_land[k][tid][1], C = s_land[k][tid][2];
                uint32_t o[8];
                realign32(A, B, C, m & 15u, o);
                stg256(dst_tile + sub0 + k * 1024 + (int)lane * 32, o);
            }
        }
    }
}

} // namespace dbeel
