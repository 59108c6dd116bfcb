"""BASELINE.json configs[4]: "memtable flush + L0 compaction pipeline, 32M zipf-skewed keys, 512B docs, 8 GPUs".

What the reference does with such a write stream (per shard, on one core each): the client routes every write to the shard
that owns murmur3_32(key) on the hash ring (src/shards.rs:95-109,586-598), the shard inserts it into its red-black-tree
memtable (lsm_tree.rs:731-771), a full tree of 8192 keys is flushed to an SSTable (:844-946), and after every flush
compact_tree (src/tasks/compaction.rs:35-102) merges tiers of >= --compaction-factor tables.

The same pipeline here, device-resident from the raw arrivals on:

    arrivals (HBM) --dbeel_route_device--> one stream per shard (index records only; payload stays put)
                   --dbeel_memtable_cuts_device--> where every memtable of every stream ends
      per wave of W memtables per shard:
                   --dbeel_flush_many_sparse_device--> W SSTables per shard in ONE launch sequence
                   --dbeel_plan_compactions (host picker, per shard) + dbeel_compact_many_device--> every shard's groups of a
                     picker round in ONE launch sequence, repeated to quiescence
    tables stay in HBM between rounds; nothing returns to the host but table sizes and the picker's decisions.

N GPUs: every rank holds the (synthetic, deterministic) stream, routes it, and keeps its own shards -- the client side of the
reference, which also hashes every key.  No data-path collective.  Which shards share a GPU when N < 8 is ours to choose, and it
matters: the reference's ring has no virtual nodes, so the arcs of `dbeel-0..7` are very unequal (24.9 %, 22.7 %, 19.8 %, 14.4 %,
12.1 %, 5.2 %, 0.5 %, 0.4 % of the hash space).  own_positions() deals the ring positions out by arc length, longest first, each to
the least loaded rank (contiguous blocks put 79 % of the writes on one of two ranks: 193 GB/s at N = 2 against 198 at N = 1).
At N = 8 the largest arc alone bounds the speed-up at 4x: that is the reference's ring, not the engine.

The checker lives outside the package (bench_cfg5.py): every shard's stream is replayed through the oracle's red-black-tree
memtable and the recorded compaction plan; every table that is left must be byte-identical (.data, .index, .bloom).
"""
from __future__ import annotations

import time
from typing import Dict, List

import numpy as np

from . import capi
from . import storage_engine as se
from . import workloads as W

N_SHARDS = 8          # one shard per core / GPU (src/main.rs:39-64)
CAPACITY = 8192       # mod.rs:18 DEFAULT_TREE_CAPACITY
FACTOR = 8            # --compaction-factor of the plan
DOC_BYTES = 512
KEY_BYTES = 17        # msgpack fixstr of 16 chars (workloads.format_keys)
FULL = 32 + KEY_BYTES + DOC_BYTES
TOMB = 32 + KEY_BYTES
SEED = 5
ZIPF_S = 0.99
TOMBSTONE_FRAC = 0.01
CFG5_WRITES = 32_000_000
CFG5_IDS = 8_000_000


# ------------------------------------------------------------------------------------ the synthetic stream

def stream_ids(n_writes: int, n_ids: int, seed: int = SEED):
    """Key id and tombstone flag of every arrival (host, deterministic): ids ~ Zipf(0.99) over n_ids, hotness decorrelated
    from key order by a fixed permutation -- the id stream of workloads.make_arrival_batch."""
    rng = np.random.default_rng(seed)
    ranks = np.arange(1, n_ids + 1, dtype=np.float64)
    p = ranks ** (-ZIPF_S)
    p /= p.sum()
    cdf = np.cumsum(p)
    hot = np.searchsorted(cdf, rng.random(n_writes), side="right").clip(0, n_ids - 1)
    perm = np.random.default_rng(seed ^ 0x5EED).permutation(n_ids)
    ids = perm[hot].astype(np.int64)
    tomb = rng.random(n_writes) < TOMBSTONE_FRAC
    return ids, tomb


def _splitmix64(torch, x):
    """splitmix64 finaliser on an int64 tensor (logical shifts emulated; multiplications wrap)."""
    def lsr(v, k):
        return (v >> k) & ((1 << (64 - k)) - 1)
    x = x + (-7046029254386353131)            # 0x9E3779B97F4A7C15
    x = (x ^ lsr(x, 30)) * (-4658895280553007687)   # 0xBF58476D1CE4E5B9
    x = (x ^ lsr(x, 27)) * (-7723592293110705685)   # 0x94D049BB133111EB
    return x ^ lsr(x, 31)


def build_stream_device(torch, dev, ids: np.ndarray, tomb: np.ndarray, chunk: int = 2_000_000):
    """The arrival batch in HBM (run layout, arrival order): (.data uint8 tensor, .index uint8 tensor).  Entry i:
    key = msgpack fixstr "k%015d" % id, document = bin16 blob of DOC_BYTES whose payload is splitmix64(i * 64 + word),
    timestamp = BASE_TS + 1000 i, tombstones carry no document.  Built with torch ops: test tooling, not the product."""
    n = ids.shape[0]
    sizes = np.where(tomb, TOMB, FULL).astype(np.int64)
    total = int(sizes.sum())
    data = torch.empty(total + 256, dtype=torch.uint8, device=dev)
    index = torch.empty(16 * n + 256, dtype=torch.uint8, device=dev)
    base = 0
    words = DOC_BYTES // 8
    col = torch.arange(words, dtype=torch.int64, device=dev)
    for c0 in range(0, n, chunk):
        c1 = min(n, c0 + chunk)
        m = c1 - c0
        idt = torch.from_numpy(ids[c0:c1]).to(dev)
        tb = torch.from_numpy(tomb[c0:c1]).to(dev)
        seq = torch.arange(c0, c1, dtype=torch.int64, device=dev)
        rows = torch.zeros((m, FULL), dtype=torch.uint8, device=dev)
        rows[:, 0] = KEY_BYTES
        rows[:, 8] = 0xB0
        rows[:, 9] = ord("k")
        rem = idt.clone()
        for k in range(16, 1, -1):
            rows[:, 8 + k] = (rem % 10 + 48).to(torch.uint8)
            rem = rem // 10
        c = 8 + KEY_BYTES
        rows[:, c] = DOC_BYTES & 255
        rows[:, c + 1] = DOC_BYTES >> 8
        doc = _splitmix64(torch, seq[:, None] * words + col[None, :]).contiguous().view(torch.uint8).reshape(m, DOC_BYTES)
        rows[:, c + 8:c + 8 + DOC_BYTES] = doc
        rows[:, c + 8] = 0xC5
        rows[:, c + 9] = (DOC_BYTES - 3) >> 8
        rows[:, c + 10] = (DOC_BYTES - 3) & 255
        ts = (W.BASE_TS + seq * 1000).contiguous().view(torch.uint8).reshape(m, 8)
        rows[:, FULL - 16:FULL - 8] = ts
        rows[tb, c] = 0
        rows[tb, c + 1] = 0
        keep = torch.ones((m, FULL), dtype=torch.bool, device=dev)
        keep[tb, c + 8:c + 8 + DOC_BYTES] = False
        flat = rows[keep]
        data[base:base + flat.numel()] = flat
        sz = torch.from_numpy(sizes[c0:c1]).to(dev)
        offs = torch.cumsum(sz, 0) - sz + base
        rec = torch.empty((m, 2), dtype=torch.int64, device=dev)
        rec[:, 0] = offs
        rec[:, 1] = (8 + KEY_BYTES) | (sz << 32)
        index[16 * c0:16 * c1] = rec.view(torch.uint8).reshape(-1)
        base += int(flat.numel())
        del rows, keep, flat, doc, rec
    assert base == total
    return data[:total], index[:16 * n], total


# ------------------------------------------------------------------------------------ the pipeline (product calls only)

def ring_arcs(ring: np.ndarray) -> np.ndarray:
    """Fraction of the 32-bit hash space each ring position owns: position i takes the hashes in [ring[i-1], ring[i]) -- the
    first shard whose hash is GREATER than the key's (shards.rs:586-598) -- and position 0 also the wrap-around."""
    h = np.asarray(ring, dtype=np.int64)
    return np.diff(np.concatenate([[h[-1] - (1 << 32)], h])) / float(1 << 32)


def own_positions(rank: int, world: int, ring=None) -> List[int]:
    """Ring positions of the shards rank `rank` of `world` runs (ascending).  With the ring: balanced by arc length (longest arc
    first, each to the least loaded rank; ties to the lower rank -- every rank computes the same table).  Without: contiguous
    blocks."""
    n = N_SHARDS if ring is None else len(ring)
    if ring is None or world == 1:
        return list(range(rank * n // world, (rank + 1) * n // world))
    arcs = ring_arcs(ring)
    load = [0.0] * world
    mine: List[int] = []
    for pos in sorted(range(n), key=lambda i: (-arcs[i], i)):
        r = min(range(world), key=lambda k: (load[k], k))
        load[r] += float(arcs[pos])
        if r == rank:
            mine.append(pos)
    return sorted(mine)


def _seed(pos: int, out_index: int) -> bytes:
    return bytes([(pos * 37 + out_index * 11 + k) & 255 for k in range(32)])


def pipeline(eng, torch, dev, data, index, ring: np.ndarray, positions, wave: int = FACTOR, capacity: int = CAPACITY,
             factor: int = FACTOR):
    """Route -> cut -> flush waves -> picker rounds, for the shards at ring `positions`.  Returns a dict with the tables left
    per shard (device views), the recorded plan, the memtable boundaries and the timings."""
    t0 = time.perf_counter()
    n = index.numel() // 16
    routed = torch.empty(index.numel() + 16, dtype=torch.uint8, device=dev)
    h64 = torch.empty(n + 2, dtype=torch.int64, device=dev)
    ms = {"route": 0.0, "cut": 0.0, "flush": 0.0, "compact": 0.0}
    flush_stages = {"ms_extract": 0.0, "ms_merge": 0.0, "ms_resolve": 0.0, "ms_gather": 0.0}
    launches = calls = flush_calls = 0
    counts, nbytes = eng.route_device((data.data_ptr(), data.numel(), index.data_ptr(), index.numel()), ring, routed.data_ptr(),
                                      index.numel(), 0, h64.data_ptr())
    st = eng.stats()
    ms["route"] += st["ms_total"]
    launches += st["kernel_launches"]
    calls += 1
    starts = np.concatenate([[0], np.cumsum(counts)]).astype(np.uint64)
    positions = sorted(int(x) for x in positions)
    # ONE call over the block of ring positions that covers the own ones: the kernel walks every stream with a CTA of its own, so
    # the call takes as long as its longest stream whatever else runs next to it -- foreign streams in between cost nothing,
    # while one call per run of neighbouring positions would walk those runs one after the other (measured: 51 ms instead of 22).
    p0, p1 = positions[0], positions[-1] + 1
    cuts = eng.memtable_cuts_device(h64.data_ptr(), starts[p0:p1 + 1], capacity)
    st = eng.stats()
    ms["cut"] += st["ms_total"]
    launches += st["kernel_launches"]
    calls += 1
    cuts_of = {pos: cuts[k] for k, pos in enumerate(range(p0, p1))}
    # memtables of every own shard: (first arrival, arrivals) inside the shard's stream; the tail that never filled is
    # flushed too (what a shutdown / the recovery flush of open_or_create_ex does, lsm_tree.rs:478-513)
    mem: Dict[int, List] = {}
    for pos in positions:
        ends = [int(x) for x in cuts_of[pos]]
        cnt = int(counts[pos])
        if not ends or ends[-1] < cnt:
            ends.append(cnt)
        lo = 0
        mem[pos] = []
        for e in ends:
            if e > lo:
                mem[pos].append((lo, e - lo))
            lo = e
    tables: Dict[int, Dict[int, tuple]] = {pos: {} for pos in mem}   # pos -> {sstable index: (data view, index view, items, bloom view|None)}
    plan_log: Dict[int, List] = {pos: [] for pos in mem}
    short = []  # full memtables whose flush did not yield `capacity` entries (a 64-bit identity collision in the cut)
    hold = []   # output buffers the table views live in
    n_waves = max((len(v) + wave - 1) // wave for v in mem.values()) if mem else 0
    rounds = 0
    for w in range(n_waves):
        batches, owner = [], []
        arrivals = 0
        for pos, lst in mem.items():
            for m_i in range(w * wave, min(len(lst), (w + 1) * wave)):
                lo, cnt = lst[m_i]
                batches.append((data.data_ptr(), data.numel(), routed.data_ptr() + 16 * (int(starts[pos]) + lo), 16 * cnt))
                owner.append((pos, m_i, cnt))
                arrivals += cnt
        if not batches:
            continue
        bound = arrivals * FULL
        fd = torch.empty(bound + 64, dtype=torch.uint8, device=dev)
        fi = torch.empty(16 * arrivals + 64, dtype=torch.uint8, device=dev)
        hold.append((fd, fi))
        _, _, _, rows = eng.flush_many_sparse_device(batches, bound, (fd.data_ptr(), bound, fi.data_ptr(), 16 * arrivals))
        st = eng.stats()
        ms["flush"] += st["ms_total"]
        for k in flush_stages:
            flush_stages[k] += st[k]
        launches += st["kernel_launches"]
        calls += 1
        flush_calls += 1
        for (pos, m_i, cnt), r in zip(owner, rows):
            tables[pos][2 * m_i] = (fd[r["data_off"]:r["data_off"] + r["data_len"]], fi[r["index_off"]:r["index_off"] + r["index_len"]],
                                    r["items"], None)
            if m_i + 1 < len(mem[pos]) and r["items"] != capacity:
                short.append((pos, m_i, r["items"]))
        # compact_tree after the flush events: every shard's picker, all groups of a round in one launch sequence
        while True:
            jobs, meta = [], []
            for pos in tables:
                plan = se.plan_compactions(sorted((i, t[2]) for i, t in tables[pos].items()), factor)
                for indices, out_index, keep in plan:
                    tb = tables[pos]
                    jobs.append(([(tb[i][0].data_ptr(), tb[i][0].numel(), tb[i][1].data_ptr(), tb[i][1].numel()) for i in indices], keep))
                    meta.append((pos, indices, out_index, keep))
            if not jobs:
                break
            rounds += 1
            dc = sum(tables[pos][i][0].numel() for pos, indices, _, _ in meta for i in indices)
            ic = sum(tables[pos][i][1].numel() for pos, indices, _, _ in meta for i in indices)
            bc = sum(16 + capi.lib().dbeel_bloom_file_size(sum(tables[pos][i][2] for i in indices), 0.01) for pos, indices, _, _ in meta)
            od = torch.empty(dc + 64, dtype=torch.uint8, device=dev)
            oi = torch.empty(ic + 64, dtype=torch.uint8, device=dev)
            ob = torch.empty(bc + 64, dtype=torch.uint8, device=dev)
            hold.append((od, oi, ob))
            seeds = [_seed(pos, out_index) for pos, _, out_index, _ in meta]
            res = eng.compact_many_device(jobs, (od.data_ptr(), dc, oi.data_ptr(), ic, ob.data_ptr(), bc), seeds=seeds)
            st = eng.stats()
            ms["compact"] += st["ms_total"]
            launches += st["kernel_launches"]
            calls += 1
            for (pos, indices, out_index, keep), r, seed in zip(meta, res, seeds):
                for i in indices:
                    del tables[pos][i]
                tables[pos][out_index] = (od[r["data_off"]:r["data_off"] + r["data_len"]], oi[r["index_off"]:r["index_off"] + r["index_len"]],
                                          r["items_written"], ob[r["bloom_off"]:r["bloom_off"] + r["bloom_len"]] if r["bloom_len"] else None)
                plan_log[pos].append((indices, out_index, keep, seed))
    torch.cuda.synchronize()
    wall_ms = (time.perf_counter() - t0) * 1e3
    own_bytes = int(sum(int(nbytes[pos]) + 16 * int(counts[pos]) for pos in mem))
    return {"tables": tables, "plan": plan_log, "mem": mem, "starts": starts, "routed": routed, "short": short, "hold": hold,
            "wall_ms": wall_ms, "ms": ms, "kernel_ms": sum(ms.values()), "launches": launches, "calls": calls, "rounds": rounds,
            "flush_calls": flush_calls, "flush_stage_ms": {k: round(v, 3) for k, v in flush_stages.items()},
            "own_arrival_bytes": own_bytes, "memtables": sum(len(v) for v in mem.values()),
            "compactions": sum(len(v) for v in plan_log.values())}
