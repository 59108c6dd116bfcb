"""Shard routing (cfg5): murmur3_32 + the consistent-hash ring.  CPU part: the oracle's restatement and the product's host
arithmetic against golden vectors from an independent implementation (scikit-learn, tests/golden/gen_murmur3_sklearn.py)
and the published SMHasher answers.  GPU part: dbeel_route_device against oracle.route on the same arrival batch."""
import json
import os

import numpy as np
import pytest

import oracle
from dbeel_b200 import capi, sstable
from dbeel_b200 import workloads as W

from helpers import nasty_keys

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "murmur3_32.json")))


def test_oracle_murmur3_matches_the_goldens():
    assert sum(1 for v in GOLD["vectors"] if v.get("published")) >= 16
    for v in GOLD["vectors"]:
        assert oracle.murmur3_32(bytes.fromhex(v["hex"]), v["seed"]) == v["hash"], v


def test_product_host_murmur3_matches_the_goldens():
    for v in GOLD["vectors"]:
        assert capi.murmur3_32(bytes.fromhex(v["hex"]), v["seed"]) == v["hash"], v


def test_ring_of_the_default_node():
    """hash_string("dbeel-<id>") (shards.rs:95-97,213-214), ascending (shards.rs:657-670 up to rotation)."""
    ring, ids = oracle.shard_ring(8)
    assert {r["name"]: r["hash"] for r in GOLD["ring_dbeel_8"]} == {f"dbeel-{i}": int(h) for h, i in zip(ring, ids)}
    assert list(ring) == sorted(ring)
    ph, pi = capi.shard_ring(8)
    assert np.array_equal(ph, ring) and np.array_equal(pi, ids)


def test_owns_key_is_between_semantics():
    """is_between(item, start, end) of shards.rs:103-109 for every shard, every boundary, with wrap-around."""
    ring, _ = oracle.shard_ring(8)

    def is_between(item, start, end):
        return (item >= end or item < start) if end < start else (start <= item < end)

    # NB: shards.rs:103 names its parameters (item, start, end) but owns_key passes (hash, previous shard, this shard)
    # and the wrapped branch reads `item >= end || item < start` with end < start -- restated literally below
    def owner_literal(h):
        n = len(ring)
        for s in range(n):
            start, end = int(ring[(s - 1) % n]), int(ring[s])
            if end < start:
                ok = h >= start or h < end
            else:
                ok = start <= h < end
            if ok:
                return s
        raise AssertionError("no owner")

    for h in [0, 1, 2**32 - 1] + [int(x) + d for x in ring for d in (-1, 0, 1)]:
        h %= 2**32
        assert oracle.ring_owner(ring, h) == owner_literal(h) == capi.ring_owner(ring, h), h
    # a key whose hash equals a shard's hash belongs to the NEXT shard (half-open interval)
    assert oracle.ring_owner(ring, int(ring[3])) == 4 and oracle.ring_owner(ring, int(ring[7])) == 0
    assert oracle.ring_owner(ring[:1], 12345) == 0  # a single shard owns everything (shards.rs:588-590)


def _arrivals(n, seed=5, nasty=False):
    if not nasty:
        return W.make_arrival_batch(n_writes=n, n_ids=max(64, n // 3), doc_bytes=64, seed=seed)
    rng = np.random.default_rng(seed)
    pool = nasty_keys(rng, max(50, n // 4), max_len=70)
    ents = [(pool[int(rng.integers(len(pool)))], bytes(rng.integers(0, 256, int(rng.integers(0, 90)), dtype=np.uint8)), 1_700_000_000 + i)
            for i in range(n)]
    return sstable.build_run(ents)


def test_cfg5_shard_placement_balances_the_ring():
    """dbeel_b200/cfg5.own_positions: every ring position goes to exactly one rank, the same table on every rank, and the ranks'
    shares of the hash space are as even as the (very uneven, no virtual nodes) arcs of dbeel-0..7 allow."""
    from dbeel_b200 import cfg5
    ring, _ = capi.shard_ring(8)
    arcs = cfg5.ring_arcs(ring)
    assert abs(float(arcs.sum()) - 1.0) < 1e-12 and (arcs > 0).all()
    # the arc of a position is the share of uniformly drawn hashes the ring routes to it
    rng = np.random.default_rng(5)
    hs = rng.integers(0, 2**32, 200_000, dtype=np.uint64)
    owners = np.array([capi.ring_owner(ring, int(h)) for h in hs[:20_000]])
    got = np.bincount(owners, minlength=8) / 20_000
    assert np.abs(got - arcs).max() < 0.01
    for world in (1, 2, 4, 8):
        parts = [cfg5.own_positions(r, world, ring) for r in range(world)]
        assert sorted(sum(parts, [])) == list(range(8))
        loads = [float(sum(arcs[p] for p in ps)) for ps in parts]
        assert max(loads) <= max(float(arcs.max()), 1.0 / world + 0.02), (world, loads)
    assert cfg5.own_positions(1, 2) == [4, 5, 6, 7]  # without a ring: contiguous blocks


@pytest.mark.gpu
@pytest.mark.parametrize("n,n_shards,nasty", [(1, 8, False), (255, 8, False), (257, 3, True), (5000, 8, True), (70_000, 8, False),
                                               (33_333, 1, False), (20_000, 64, True), (9_000, 256, False)])
def test_route_device_matches_the_oracle(engine, n, n_shards, nasty):
    import torch
    batch = _arrivals(n, seed=n, nasty=nasty)
    ring, _ = oracle.shard_ring(n_shards)
    exp_shard, _ = oracle.route(batch, ring)
    dev = torch.device("cuda:0")
    d = torch.from_numpy(np.ascontiguousarray(batch[0])).to(dev)
    ix = torch.from_numpy(np.ascontiguousarray(batch[1])).to(dev)
    out = torch.empty(ix.numel() + 16, dtype=torch.uint8, device=dev)
    owner = torch.empty(n, dtype=torch.int32, device=dev)
    counts, nbytes = engine.route_device((d.data_ptr(), d.numel(), ix.data_ptr(), ix.numel()), ring, out.data_ptr(), ix.numel(),
                                         owner.data_ptr())
    assert np.array_equal(owner.cpu().numpy().view(np.uint32), exp_shard)
    recs = batch[1].reshape(-1, 16)
    exp_idx = np.concatenate([recs[exp_shard == s] for s in range(n_shards)])  # stable split, ring order
    assert np.array_equal(out[:ix.numel()].cpu().numpy().reshape(-1, 16), exp_idx)
    assert list(counts) == [int((exp_shard == s).sum()) for s in range(n_shards)]
    fs = recs[:, 12:16].copy().view("<u4").ravel()
    assert list(nbytes) == [int(fs[exp_shard == s].sum()) for s in range(n_shards)]


@pytest.mark.gpu
def test_route_then_sparse_flush_equals_flushing_each_shards_stream(engine):
    """cfg5's front end: route 60k Zipf arrivals to 8 shards, cut every shard's stream into memtables of 512 distinct keys,
    flush all memtables of all shards with ONE dbeel_flush_many_sparse_device, compare every SSTable with the oracle's
    red-black-tree flush of the same arrivals."""
    import torch
    from dbeel_b200 import storage_engine as se
    n, cap = 60_000, 512
    batch = W.make_arrival_batch(n_writes=n, n_ids=9000, doc_bytes=96, seed=15)
    ring, _ = oracle.shard_ring(8)
    shard, _ = oracle.route(batch, ring)
    dev = torch.device("cuda:0")
    d = torch.from_numpy(np.ascontiguousarray(batch[0])).to(dev)
    ix = torch.from_numpy(np.ascontiguousarray(batch[1])).to(dev)
    routed = torch.empty(ix.numel() + 16, dtype=torch.uint8, device=dev)
    counts, nbytes = engine.route_device((d.data_ptr(), d.numel(), ix.data_ptr(), ix.numel()), ring, routed.data_ptr(), ix.numel())
    recs = batch[1].reshape(-1, 16)
    batches, expect = [], []
    start = 0
    for s in range(8):
        stream_idx = np.ascontiguousarray(recs[shard == s]).ravel()
        stream = (batch[0], stream_idx)  # sparse host view of the shard's stream, for the host-side cut and the oracle
        pos, cnt = 0, int(counts[s])
        while pos < cnt:
            m = se.memtable_cut(stream, pos, cap)
            batches.append((d.data_ptr(), d.numel(), routed.data_ptr() + 16 * (start + pos), 16 * m))
            dense = sstable.build_run(sstable.parse_run(batch[0], stream_idx[16 * pos:16 * (pos + m)]))
            (od, oi, on), = oracle.memtable_flushes(dense, capacity=1 << 30)
            expect.append((od, oi, on))
            pos += m
        start += cnt
    bound = int(nbytes.sum())
    od = torch.empty(bound + 64, dtype=torch.uint8, device=dev)
    oi = torch.empty(ix.numel() + 64, dtype=torch.uint8, device=dev)
    _, _, _, rows = engine.flush_many_sparse_device(batches, bound, (od.data_ptr(), bound, oi.data_ptr(), ix.numel()))
    assert len(rows) == len(expect) > 8
    for r, (ed, ei, en) in zip(rows, expect):
        assert r["items"] == en
        assert np.array_equal(od[r["data_off"]:r["data_off"] + r["data_len"]].cpu().numpy(), ed)
        assert np.array_equal(oi[r["index_off"]:r["index_off"] + r["index_len"]].cpu().numpy(), ei)
    written = sum(r["data_len"] for r in rows)  # < bound: arrivals of one key inside a memtable collapse
    with pytest.raises(capi.DbeelError) as ei_:
        engine.flush_many_sparse_device(batches, written - 1, (od.data_ptr(), bound, oi.data_ptr(), ix.numel()))
    assert ei_.value.code == 2  # DBEEL_ERR_CAPACITY: the bound was too low, nothing past it was written


@pytest.mark.gpu
@pytest.mark.parametrize("n,cap,n_ids", [(40_000, 512, 3000), (100_000, 8192, 60_000), (30_000, 1, 50), (5_000, 9000, 100_000),
                                         (70_000, 1000, 1001)])
def test_device_memtable_cuts_match_the_host_trigger(engine, n, cap, n_ids):
    """dbeel_memtable_cuts_device (64-bit key identities, one CTA per stream) against dbeel_memtable_cut (exact, key bytes)
    on every shard's stream: lsm_tree.rs:747-765 -- the flush starts right after the insert that fills the tree."""
    import torch
    from dbeel_b200 import storage_engine as se
    batch = W.make_arrival_batch(n_writes=n, n_ids=n_ids, doc_bytes=40, seed=cap)
    ring, _ = oracle.shard_ring(8)
    shard, _ = oracle.route(batch, ring)
    dev = torch.device("cuda:0")
    d = torch.from_numpy(np.ascontiguousarray(batch[0])).to(dev)
    ix = torch.from_numpy(np.ascontiguousarray(batch[1])).to(dev)
    routed = torch.empty(ix.numel() + 16, dtype=torch.uint8, device=dev)
    h64 = torch.empty(n, dtype=torch.int64, device=dev)
    counts, _ = engine.route_device((d.data_ptr(), d.numel(), ix.data_ptr(), ix.numel()), ring, routed.data_ptr(), ix.numel(), 0,
                                    h64.data_ptr())
    starts = np.concatenate([[0], np.cumsum(counts)]).astype(np.uint64)
    got = engine.memtable_cuts_device(h64.data_ptr(), starts, cap)
    recs = batch[1].reshape(-1, 16)
    for s in range(8):
        stream = (batch[0], np.ascontiguousarray(recs[shard == s]).ravel())
        exp, pos, cnt = [], 0, int(counts[s])
        while pos < cnt:
            m = se.memtable_cut(stream, pos, cap)
            pos += m
            exp.append(pos)
        # the host loop also reports the memtable that is still filling when the stream ends; the device reports full ones
        full = exp if (exp and _distinct(stream, exp[-2] if len(exp) > 1 else 0, exp[-1]) == cap) else exp[:-1]
        assert list(got[s]) == full, (s, list(got[s])[:5], full[:5])


def _distinct(stream, lo, hi):
    return len({k for k, _, _ in sstable.parse_run(stream[0], stream[1][16 * lo:16 * hi])})
