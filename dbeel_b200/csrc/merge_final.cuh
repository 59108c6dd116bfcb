// merge_final.cuh -- the LAST merge level, resolve, the offsets scan and the .index writes of a single compaction in ONE
// persistent kernel (round 2).
//
// Round 1 ran them as five kernels: k_merge_tma (last level) -> k_resolve -> k_scan_tiles -> k_scan_chunks -> k_emit, with the
// merged records and the per-position results (`res`) making a round trip through HBM between them (4 x 16 bytes per input
// entry) and two latency-bound launches of one thread per record.  Here a tile of the last level stays in shared memory
// after it has been merged:
//
//   TMA (2 bulk copies, mbarrier)   A / B ranges of the tile, plus the record on either side of each range (the neighbours
//                                   that decide whether a group of equal keys crosses the tile's edges)
//   merge-path, 7 records / thread  sorted tile in shared memory (as k_merge_tma)
//   resolve                         7 independent chains per thread: index record by gid (64-byte granule), timestamp only
//                                   for members of a group; the head of a group picks max (timestamp, run position) over
//                                   shared memory (lsm_tree.rs:1036-1066, mod.rs:75-81) and applies the tombstone rule
//   scan                            (bytes, entries) of the tile's survivors; chained scan over the tiles with a 256-wide
//                                   look-back (tiles are taken in order by co-resident persistent CTAs, so a predecessor is
//                                   always running or done)
//   emit                            output .index records, src_ptr, tile_first (entry_writer.rs:76-86) straight from registers
//
// What no longer exists: the last level's 16 B/entry write, k_resolve's 16 B read + 16 B `res` write, k_emit's 16 B read,
// three kernel boundaries and the two scan kernels.
#pragma once
#include "kernels.cuh"

namespace dbeel {

constexpr int kFinThreads = kMergeThreads;
constexpr int kFinVT = kMergeVT;
constexpr int kFinTile = kMergeTile;           // capacity of a tile: 7 records per thread
constexpr int kFinNominal = kFinTile - 64;     // nominal tile (k_merge_partition's split points); + up to 31 + 31 records of a straddling group
constexpr int kFinBufRecs = kFinTile + 8;      // A_prev | A | A_next | B_prev | B | B_next, + slack
constexpr int kFinMaxRunsSmem = 64;            // run tables (base / index / data) cached in shared memory up to this many runs
#ifndef DBEEL_FIN_CTAS
#define DBEEL_FIN_CTAS 2
#endif
constexpr uint32_t kFinSmem = 2u * kFinBufRecs * 16u + 2u * kFinTile * 8u + (uint32_t)kFinTile + 16u;

struct FinDesc {
    uint32_t a_src, n_a, b_src, n_b; // record offsets into src, counts
    uint32_t a_end, b_end;           // one past the last record of segment A / B (absolute)
    uint32_t diag0;                  // merged position of the tile's first record, relative to the pair's output
    uint32_t nb;                     // bit 0: A[a0-1] exists, 1: A[a1] exists, 2: B[b0-1] exists, 3: B[b1] exists
};

template <bool kNarrow>
__global__ void __launch_bounds__(kFinThreads, DBEEL_FIN_CTAS) k_merge_final(Params p, uint32_t level, const Rec *src) {
    constexpr int NT = kFinThreads, VT = kFinVT;
    extern __shared__ __align__(128) uint8_t f_raw[];
    Rec *bufs[2] = {reinterpret_cast<Rec *>(f_raw), reinterpret_cast<Rec *>(f_raw) + kFinBufRecs};
    unsigned long long *s_tlo = reinterpret_cast<unsigned long long *>(f_raw + 2u * kFinBufRecs * 16u);
    unsigned long long *s_thi = s_tlo + kFinTile;
    uint8_t *s_flag = reinterpret_cast<uint8_t *>(s_thi + kFinTile); // bit 0: same key as the next merged record
    __shared__ __align__(8) uint64_t s_bar[2];
    __shared__ Rec s_bnd[5]; // A[a0-1], A[a1], B[b0-1], B[b1], the tile's last record
    __shared__ unsigned long long s_wb[NT / 32];
    __shared__ uint32_t s_wc[NT / 32];
    __shared__ unsigned long long s_pref_b;
    __shared__ uint32_t s_pref_c;
    __shared__ uint32_t s_first[NT / 32];
    __shared__ uint32_t s_rbase[kFinMaxRunsSmem + 1];
    __shared__ uint8_t s_rlut[256]; // run that holds gid (b << lut_shift): a record's run is that one or a close successor
    __shared__ const uint4 *s_rindex[kFinMaxRunsSmem];
    __shared__ const uint8_t *s_rdata[kFinMaxRunsSmem];

    pdl_trigger();
    pdl_wait();
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t skip = p.ctl->prefix_len + kWindowBytes;
    const uint32_t n_tiles = p.tile_base[level][p.nseg[level + 1]];
    const uint32_t G = gridDim.x;
    uint32_t tile = blockIdx.x;
    if (tile >= n_tiles) return;
    const bool runs_cached = p.n_runs <= (uint32_t)kFinMaxRunsSmem;
    uint32_t lut_shift = 0;
    while (((uint64_t)p.n_total >> lut_shift) > 256) lut_shift++; // gid >> lut_shift < 256 for every gid < n_total
    if (runs_cached) {
        for (uint32_t r = tid; r < p.n_runs; r += NT) {
            s_rbase[r] = p.runs[r].base;
            s_rindex[r] = p.runs[r].index;
            s_rdata[r] = p.runs[r].data;
        }
        if (tid == 0) s_rbase[p.n_runs] = 0xFFFFFFFFu; // sentinel: no run starts above any gid
        s_rlut[tid] = (uint8_t)find_run(p, (uint32_t)(((uint64_t)tid << lut_shift) < p.n_total ? (uint64_t)tid << lut_shift : p.n_total - 1));
    }
    if (tid == 0) {
        mbar_init(&s_bar[0], 1);
        mbar_init(&s_bar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    // entry address / key_size / full_size of the entry behind a gid: its run's index record (64-byte granule)
    auto run_of = [&](uint32_t gid) -> uint32_t {
        if (!runs_cached) return find_run(p, gid);
        uint32_t r = s_rlut[gid >> lut_shift]; // owner of the bucket's first gid; ours is the last run whose base <= gid
        while (s_rbase[r + 1] <= gid) r++;
        return r;
    };
    auto index_ptr = [&](uint32_t gid, uint32_t r) -> const uint4 * {
        return runs_cached ? s_rindex[r] + (gid - s_rbase[r]) : p.runs[r].index + (gid - p.runs[r].base);
    };
    auto data_ptr = [&](uint32_t r) -> const uint8_t * { return runs_cached ? s_rdata[r] : p.runs[r].data; };

    auto issue = [&](const FinDesc &d, Rec *buf, uint64_t *bar) { // thread 0 only
        const uint32_t ap = d.nb & 1u, an = (d.nb >> 1) & 1u, bp = (d.nb >> 2) & 1u, bn = (d.nb >> 3) & 1u;
        const uint32_t ca = d.n_a + ap + an, cb = d.n_b + bp + bn;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); // the buffer was last written through the generic proxy
        mbar_expect_tx(bar, (ca + cb) * 16u);
        if (ca) tma_load_1d(buf + 1 - ap, &src[d.a_src - ap], ca * 16u, bar);
        if (cb) tma_load_1d(buf + d.n_a + 3 - bp, &src[d.b_src - bp], cb * 16u, bar);
    };

    // Tile descriptors from k_merge_partition's boundary records (already moved past straddling groups): the boundary index
    // three tiles ahead, the two records two tiles ahead -- no dependent chain inside an iteration.  The last level of a
    // single compaction is one pair: segments 0 (A) and 1 (B).
    const Seg segA = p.seg[level][0];
    Seg segB;
    segB.start = 0; segB.len = 0;
    if (p.nseg[level] > 1) segB = p.seg[level][1];
    auto ld_bidx = [&](uint32_t t) -> uint32_t { return t < n_tiles ? __ldg(&p.tile_bnd[t]) : 0xFFFFFFFFu; };
    auto mk_desc = [&](uint32_t bidx) -> FinDesc {
        FinDesc d;
        d.a_src = d.n_a = d.b_src = d.n_b = d.a_end = d.b_end = d.diag0 = d.nb = 0;
        if (bidx == 0xFFFFFFFFu) return d;
        const uint4 b0 = __ldg(&p.bnd[bidx]), b1 = __ldg(&p.bnd[bidx + 1]);
        d.a_src = b0.x; d.n_a = b1.x - b0.x;
        d.b_src = b0.y; d.n_b = b1.y - b0.y;
        d.a_end = segA.start + segA.len;
        d.b_end = segB.start + segB.len;
        d.diag0 = b0.z;
        d.nb = (b0.x > segA.start ? 1u : 0u) | (b1.x < d.a_end ? 2u : 0u) | (b0.y > segB.start ? 4u : 0u) | (b1.y < d.b_end ? 8u : 0u);
        return d;
    };
    FinDesc cur = mk_desc(ld_bidx(tile));
    FinDesc nxt = mk_desc(ld_bidx(tile + G));
    uint32_t bidx2 = ld_bidx(tile + 2 * G);
    if (tid == 0) issue(cur, bufs[0], &s_bar[0]);
    const int keep_tombstones = p.keep_tombstones;

    for (uint32_t q = 0;; q++) {
        Rec *s = bufs[q & 1];
        uint4 *s4 = reinterpret_cast<uint4 *>(s);
        const bool has_next = tile + G < n_tiles;
        if (tid == 0 && has_next) issue(nxt, bufs[(q + 1) & 1], &s_bar[(q + 1) & 1]);
        const FinDesc nn = mk_desc(bidx2);            // consumed one iteration from now
        const uint32_t bidx3 = ld_bidx(tile + 3 * G); // ... and two iterations from now
        while (!mbar_try_wait(&s_bar[q & 1], (q >> 1) & 1)) {}

        // ---- merge-path: thread t produces merged records [7t, 7t + 7) of the tile
        const uint32_t nA = cur.n_a, nB = cur.n_b, n = nA + nB;
        const Rec *A = s + 1, *B = s + nA + 3;
        uint32_t d = tid * VT;
        if (d > n) d = n;
        Rec me[VT + 1]; // this thread's merged records, then the one after them
        {
            uint32_t lo = d > nB ? d - nB : 0;
            uint32_t hi = d < nA ? d : nA;
            while (lo < hi) {
                const uint32_t mid = (lo + hi) >> 1;
                if (!key_less_smem(p, skip, &B[d - 1 - mid], &A[mid])) lo = mid + 1; else hi = mid;
            }
            uint32_t ai = lo, bi = d - lo;
            Rec ak = A[ai], bk = B[bi]; // may read one slot past a range: the neighbour slots
#pragma unroll
            for (int i = 0; i < VT; i++) {
                const bool has_a = ai < nA, has_b = bi < nB;
                const bool take_b = has_b && (!has_a || key_less(p, skip, bk, ak));
                me[i] = take_b ? bk : ak;
                if (take_b) { bi++; bk = B[bi]; } else { ai++; ak = A[ai]; }
            }
            if (tid == 0) {
                s_bnd[0] = s[0];
                s_bnd[1] = s[nA + 1];
                s_bnd[2] = s[nA + 2];
                s_bnd[3] = s[nA + 3 + nB];
            }
            __syncthreads(); // every thread is done reading the unmerged ranges
            // only a thread's first and last record are ever read by its neighbours (group flags across thread borders)
            if (d < n) s[d] = me[0];
            if (d + VT - 1 < n) s[d + VT - 1] = me[VT - 1];
#pragma unroll
            for (int i = 0; i < VT; i++)
                if (d + i == n - 1) s_bnd[4] = me[i];
        }
        __syncthreads();

        // ---- resolve, part 1: group flags from the neighbours' keys, index record of every entry
        uint32_t m_act = 0, m_eqn = 0, m_eqp = 0;
        uint32_t gidv[VT];
        uint4 irec[VT];
        {
#pragma unroll
            for (int i = 0; i < VT; i++)
                if (d + i < n) m_act |= 1u << i;
            me[VT].x = me[VT].y = me[VT].z = me[VT].w = 0;
            if (d + VT < n) me[VT] = s[d + VT];
#pragma unroll
            for (int i = 0; i < VT; i++) {
                if (!(m_act & (1u << i))) continue;
                bool eqn;
                if (d + i + 1 < n) {
                    eqn = key_equal(p, skip, me[i], me[i + 1]);
                } else { // the tile's last record: the next merged record is A[a1] or B[b1]
                    eqn = ((cur.nb & 2u) && key_equal(p, skip, me[i], s_bnd[1])) || ((cur.nb & 8u) && key_equal(p, skip, me[i], s_bnd[3]));
                }
                if (eqn) m_eqn |= 1u << i;
            }
            m_eqp = m_eqn << 1;
            if (m_act & 1u) {
                bool eqp;
                if (d > 0) eqp = key_equal(p, skip, s[d - 1], me[0]);
                else eqp = ((cur.nb & 1u) && key_equal(p, skip, s_bnd[0], me[0])) || ((cur.nb & 4u) && key_equal(p, skip, s_bnd[2], me[0]));
                if (eqp) m_eqp |= 1u;
            }
#pragma unroll
            for (int i = 0; i < VT; i++) {
                gidv[i] = me[i].w;
                irec[i] = make_uint4(0, 0, 0, 0);
                if (m_act & (1u << i)) {
                    const uint32_t r = run_of(gidv[i]);
                    const uint4 *q4 = index_ptr(gidv[i], r);
                    irec[i] = kNarrow ? ldg128_narrow(q4) : __ldg(q4);
                }
            }
        }
        __syncthreads(); // every neighbour key has been read: the tile's slots may now hold {entry address, key_size, full_size}
        uint4 we[VT]; // what is written for position d + i: {entry address, key_size, full_size or 0}; first the entry itself
        {
            const uint32_t m_grp = m_act & (m_eqn | m_eqp); // members of a group of two or more
            unsigned long long tlo[VT], thi[VT];
#pragma unroll
            for (int i = 0; i < VT; i++) {
                tlo[i] = thi[i] = 0;
                we[i] = make_uint4(0, 0, 0, 0);
                if (!(m_act & (1u << i))) continue;
                const uint32_t r = run_of(gidv[i]);
                const uint8_t *entry = data_ptr(r) + ((uint64_t)irec[i].x | ((uint64_t)irec[i].y << 32));
                const unsigned long long ea = (unsigned long long)(uintptr_t)entry;
                we[i] = make_uint4((uint32_t)ea, (uint32_t)(ea >> 32), irec[i].z, irec[i].w);
                if ((m_grp >> i) & 1u) { // its timestamp decides (mod.rs:80); only group members go through shared memory
                    const uint8_t *t = entry + irec[i].w - 16;
                    if (kNarrow) { tlo[i] = ld_u64_unaligned_narrow(t); thi[i] = ld_u64_unaligned_narrow(t + 8); }
                    else { tlo[i] = ld_u64_unaligned(t); thi[i] = ld_u64_unaligned(t + 8); }
                    s4[d + i] = we[i];
                    s_flag[d + i] = (uint8_t)((m_eqn >> i) & 1u);
                }
            }
#pragma unroll
            for (int i = 0; i < VT; i++) {
                if ((m_grp >> i) & 1u) { s_tlo[d + i] = tlo[i]; s_thi[d + i] = thi[i]; }
            }
        }
        __syncthreads();

        // ---- resolve, part 2: heads pick their group's winner; the tombstone rule (lsm_tree.rs:1045-1046)
        unsigned long long vb = 0;
        uint32_t vc = 0;
#pragma unroll
        for (int i = 0; i < VT; i++) {
            uint4 info = we[i];
            we[i].w = 0;
            if (!((m_act >> i) & 1u) || ((m_eqp >> i) & 1u)) continue; // not a head
            const uint32_t k = d + i;
            if ((m_eqn >> i) & 1u) {
                unsigned long long wlo = s_tlo[k], whi = s_thi[k];
                uint32_t j = k;
                bool more = true;
                while (more && j + 1 < n) { // members inside the tile, in run-position order: a later member wins ties
                    j++;
                    const unsigned long long clo = s_tlo[j], chi = s_thi[j];
                    if (!ts_greater(wlo, whi, clo, chi)) { info = s4[j]; wlo = clo; whi = chi; }
                    more = s_flag[j] != 0;
                }
                if (more) {
                    // The group runs past the tile: its remaining members are the records of A from a1 on and then of B from
                    // b1 on that carry the same key -- still in run-position order (a group that continues in A has no member
                    // from B inside the tile: ties take A first).
                    const Rec last = s_bnd[4];
                    for (int side = 0; side < 2; side++) {
                        uint32_t g = side ? cur.b_src + nB : cur.a_src + nA;
                        const uint32_t end = side ? cur.b_end : cur.a_end;
                        for (; g < end; g++) {
                            const Rec nx = ld_rec(&src[g]);
                            if (!key_equal(p, skip, last, nx)) break;
                            const KeyRef ck = key_of_gid(p, nx.w);
                            uint64_t clo, chi;
                            ld_ts(ck.entry, ck.full_size, &clo, &chi);
                            if (!ts_greater(wlo, whi, clo, chi)) {
                                const unsigned long long ea = (unsigned long long)(uintptr_t)ck.entry;
                                info = make_uint4((uint32_t)ea, (uint32_t)(ea >> 32), ck.klen + 8, ck.full_size);
                                wlo = clo; whi = chi;
                            }
                        }
                    }
                }
            }
            const bool tomb = info.w == info.z + 24;
            if (keep_tombstones || !tomb) {
                we[i] = info;
                vb += info.w;
                vc += 1;
            }
        }

        // ---- scan: this thread's survivors -> tile -> all tiles before this one
        unsigned long long ib = vb;
        uint32_t ic = vc;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long xb = __shfl_up_sync(0xFFFFFFFFu, ib, o);
            const uint32_t xc = __shfl_up_sync(0xFFFFFFFFu, ic, o);
            if (lane >= (uint32_t)o) { ib += xb; ic += xc; }
        }
        if (lane == 31) { s_wb[warp] = ib; s_wc[warp] = ic; }
        __syncthreads();
        unsigned long long tb = 0, wbefore = 0;
        uint32_t tc = 0, wcbefore = 0;
#pragma unroll
        for (int w = 0; w < NT / 32; w++) {
            if ((uint32_t)w < warp) { wbefore += s_wb[w]; wcbefore += s_wc[w]; }
            tb += s_wb[w];
            tc += s_wc[w];
        }
        unsigned long long *mine = p.scan_state + 2ull * tile;
        if (tile == 0) {
            if (tid == 0) {
                st_volatile_u64(mine, (kScanPrefix << 62) | tb);
                st_volatile_u64(mine + 1, (kScanPrefix << 32) | tc);
                s_pref_b = 0;
                s_pref_c = 0;
            }
            __syncthreads();
        } else {
            // Each word says what it holds (aggregate or inclusive prefix): a reader needs both words at the same stage and simply
            // reads again when it caught the pair mid-update -- no fence on either side.
            if (tid == 0) {
                st_volatile_u64(mine, (kScanAgg << 62) | tb);
                st_volatile_u64(mine + 1, (kScanAgg << 32) | tc);
            }
            unsigned long long eb = 0;
            uint32_t ec = 0;
            int base = (int)tile - 1;
            while (true) { // 256 predecessors per step: thread t looks at tile base - t
                const int idx = base - (int)tid;
                unsigned long long vb2 = 0, vc2 = 0, stat = kScanPrefix; // tiles before the first one: a prefix of nothing
                if (idx >= 0) {
                    const unsigned long long *qs = p.scan_state + 2ull * (uint32_t)idx;
                    while (true) {
                        vc2 = ld_volatile_u64(qs + 1);
                        vb2 = ld_volatile_u64(qs);
                        if ((vc2 >> 32) != 0 && (vc2 >> 32) == (vb2 >> 62)) break;
                    }
                    stat = vc2 >> 32;
                    vb2 &= (1ull << 62) - 1;
                    vc2 &= 0xFFFFFFFFull;
                }
                // the nearest predecessor that already holds an inclusive prefix ends the walk
                const uint32_t pm = __ballot_sync(0xFFFFFFFFu, stat == kScanPrefix);
                __syncthreads(); // s_first / s_wb / s_wc of the previous step (or of the tile scan) have been read
                if (lane == 0) s_first[warp] = pm ? (warp * 32u + (uint32_t)__ffs((int)pm) - 1u) : 0xFFFFFFFFu;
                __syncthreads();
                uint32_t first = 0xFFFFFFFFu;
#pragma unroll
                for (int w = 0; w < NT / 32; w++) first = s_first[w] < first ? s_first[w] : first;
                unsigned long long cb = tid <= first ? vb2 : 0ull;
                uint32_t cc = tid <= first ? (uint32_t)vc2 : 0u;
#pragma unroll
                for (int o = 16; o; o >>= 1) {
                    cb += __shfl_xor_sync(0xFFFFFFFFu, cb, o);
                    cc += __shfl_xor_sync(0xFFFFFFFFu, cc, o);
                }
                if (lane == 0) { s_wb[warp] = cb; s_wc[warp] = cc; }
                __syncthreads();
#pragma unroll
                for (int w = 0; w < NT / 32; w++) { eb += s_wb[w]; ec += s_wc[w]; }
                if (first != 0xFFFFFFFFu) break;
                base -= NT;
            }
            if (tid == 0) {
                st_volatile_u64(mine, (kScanPrefix << 62) | (eb + tb));
                st_volatile_u64(mine + 1, (kScanPrefix << 32) | (unsigned long long)(ec + tc));
                s_pref_b = eb;
                s_pref_c = ec;
            }
            __syncthreads();
        }
        const unsigned long long pref_b = s_pref_b;
        const uint32_t pref_c = s_pref_c;
        if (tile + 1 == n_tiles && tid == 0) { // the last tile holds the totals
            Ctl *cw = p.ctl;
            cw->out_data_len = pref_b + tb;
            cw->out_items = pref_c + tc;
        }

        // ---- emit: .index records (entry_writer.rs:76-86), source addresses, gather-tile markers
        {
            unsigned long long off = pref_b + wbefore + ib - vb; // within this job's .data
            uint32_t pos = pref_c + wcbefore + ic - vc;
            constexpr unsigned long long gt = kGatherTileBytes;
#pragma unroll
            for (int i = 0; i < VT; i++) {
                const uint32_t fs = we[i].w;
                if (!fs) continue;
                const unsigned long long file_off = off + p.out_offset_base;
                p.out_index[pos] = make_uint4((uint32_t)file_off, (uint32_t)(file_off >> 32), we[i].z, fs);
                p.src_ptr[pos] = (unsigned long long)we[i].x | ((unsigned long long)we[i].y << 32);
                for (unsigned long long bq = (off + gt - 1) / gt; bq * gt < off + fs && bq < p.tile_first_n; bq++) p.tile_first[bq] = pos;
                off += fs;
                pos++;
            }
        }
        if (!has_next) break;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); // generic-proxy accesses of this buffer before the next bulk copy into it
        __syncthreads(); // s_pref / s_bnd / the tile buffer are free for the next tile
        tile += G;
        cur = nxt;
        nxt = nn;
        bidx2 = bidx3;
    }
}

} // namespace dbeel
