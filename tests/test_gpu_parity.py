"""GPU parity tests: the CUDA path, called through the C ABI (dbeel_b200.capi -> libdbeel_compact.so),
must produce byte-identical .data / .index / .bloom to the CPU oracle on the same inputs.

Restates the reference's own compaction test (lsm_tree.rs:1328-1451) against the GPU engine and
adds the byte-level, tie-break, tombstone, truncation and bloom cases the reference never tests.
"""
import dataclasses

import numpy as np
import pytest

import oracle
from dbeel_b200 import capi, sstable
from dbeel_b200 import workloads as W

from helpers import BASE_TS, assert_run_equal, model_compact, model_flush, nasty_keys, random_runs

pytestmark = pytest.mark.gpu

SEED = bytes(range(32))


def u16key(n):
    return int(n).to_bytes(2, "little")


def check_against_oracle(engine, runs, keep, bloom_min_size=capi.DEFAULT_BLOOM_MIN_SIZE, what=""):
    gd, gi, gb, gn = engine.compact(runs, keep_tombstones=keep, bloom_min_size=bloom_min_size, seed=SEED)
    od, oi, ob, on = oracle.compact(runs, keep_tombstones=keep, bloom_min_size=bloom_min_size, seed=SEED)
    assert gn == on, f"{what}: items_written {gn} != {on}"
    assert_run_equal((gd, gi), (od, oi), what)
    assert (gb is None) == (ob is None), f"{what}: bloom presence"
    if ob is not None:
        assert gb.size == ob.size
        if not np.array_equal(gb, ob):
            bad = int(np.flatnonzero(gb != ob)[0])
            raise AssertionError(f"{what}: .bloom differs at byte {bad} of {ob.size}")
    return gd, gi, gb, gn


def test_cfg1_two_way_1k_keys(engine):
    """BASELINE.json configs[0]: the reference's own CPU-runnable case."""
    runs = W.make_merge_runs(W.CFG1)
    _, _, gb, n = check_against_oracle(engine, runs, keep=False, what="cfg1")
    assert gb is None and 1800 < n < 1900  # 226 KB of .data <= 1 MiB: no bloom (lsm_tree.rs:1027)
    st = engine.stats()
    assert st["entries_in"] == 2000 and st["entries_out"] == n and st["merge_passes"] == 1
    assert st["key_prefix_len"] == 13  # 0xb0 'k' + 11 leading zero digits are common to ids < 1900


def test_get_after_compaction(engine):
    """lsm_tree.rs:1400-1446 through the GPU engine: flushes (0,32),(2,32),(4,32) -> (5, 92)."""
    writes = [(u16key(n), u16key(n), BASE_TS + n) for n in range(94)]
    writes += [(u16key(1), b"", BASE_TS + 1000), (u16key(4), b"", BASE_TS + 1001)]
    runs = [(d, i) for d, i, _ in oracle.memtable_flushes(sstable.build_run(writes), capacity=32)]
    gd, gi, gb, n = check_against_oracle(engine, runs, keep=False, what="get_after_compaction")
    assert n == 92 and gb is None
    got = {k: v for k, v, _ in sstable.parse_run(gd, gi)}
    assert got[u16key(0)] == u16key(0) and got[u16key(2)] == u16key(2) and got[u16key(10)] == u16key(10)
    assert u16key(1) not in got and u16key(4) not in got
    assert [v for k, v, _ in sstable.parse_run(gd, gi) if u16key(1) <= k < u16key(5)] == [u16key(2), u16key(3)]
    _, _, _, n2 = check_against_oracle(engine, runs, keep=True, what="get_after_compaction keep")
    assert n2 == 94


@pytest.mark.parametrize("seed", range(8))
@pytest.mark.parametrize("keep", [False, True])
def test_adversarial_keys_match_oracle_and_model(engine, seed, keep):
    rng = np.random.default_rng(100 + seed)
    pool = nasty_keys(rng, 3000, max_len=48)
    k = int(rng.integers(1, 12))
    runs = random_runs(rng, k, [int(rng.integers(0, 2500)) for _ in range(k)], pool)
    gd, gi, _, n = check_against_oracle(engine, runs, keep, what=f"adversarial {seed}")
    exp, en = model_compact(runs, keep)
    assert n == en
    assert_run_equal((gd, gi), exp, "vs model")


def test_tie_breaks(engine):
    k = b"same"
    runs = [sstable.build_run([(k, b"r0", 5)]), sstable.build_run([(k, b"r1", 5)]), sstable.build_run([(k, b"r2", -7)])]
    gd, gi, _, n = check_against_oracle(engine, runs, False, what="ties")
    assert sstable.parse_run(gd, gi) == [(k, b"r1", 5)]
    # signed 128-bit order: a negative timestamp loses to zero even though its low word is huge
    runs = [sstable.build_run([(k, b"neg", -1)]), sstable.build_run([(k, b"zero", 0)]),
            sstable.build_run([(k, b"big", (1 << 64) + 5)]), sstable.build_run([(k, b"lowword", (1 << 64) - 1)])]
    gd, gi, _, _ = check_against_oracle(engine, runs, False, what="i128")
    assert sstable.parse_run(gd, gi) == [(k, b"big", (1 << 64) + 5)]
    # all timestamps equal over 9 runs: the last run position wins every key
    ents = [(bytes([65 + j]), b"v", 7) for j in range(20)]
    runs = [sstable.build_run([(kk, b"run%d" % r, ts) for kk, _, ts in ents]) for r in range(9)]
    gd, gi, _, n = check_against_oracle(engine, runs, False, what="equal ts")
    assert n == 20 and all(v == b"run8" for _, v, _ in sstable.parse_run(gd, gi))


def test_tombstones(engine):
    runs = [sstable.build_run([(b"a", b"v", 1), (b"b", b"v", 1), (b"c", b"", 1)]),
            sstable.build_run([(b"a", b"", 2), (b"c", b"back", 2)])]
    gd, gi, _, n = check_against_oracle(engine, runs, False, what="tomb drop")
    assert sstable.parse_run(gd, gi) == [(b"b", b"v", 1), (b"c", b"back", 2)]
    gd, gi, _, n = check_against_oracle(engine, runs, True, what="tomb keep")
    assert sstable.parse_run(gd, gi) == [(b"a", b"", 2), (b"b", b"v", 1), (b"c", b"back", 2)]
    only = [sstable.build_run([(bytes([j]), b"", 1) for j in range(50)])]
    _, _, _, n = check_against_oracle(engine, only, False, what="all tombstones")
    assert n == 0


def test_empty_and_ragged_inputs(engine):
    e = (np.zeros(0, np.uint8), np.zeros(0, np.uint8))
    assert engine.compact([], False)[3] == 0
    assert engine.compact([e, e], False)[3] == 0
    a = sstable.build_run([(b"", b"empty key is a key", 1), (b"\x00", b"x", 1)])
    check_against_oracle(engine, [e, a, e], False, what="empty runs around")
    b = sstable.build_run([(bytes([n]), b"B" * n, 2) for n in range(5, 15)])
    ragged = (b[0], np.concatenate([b[1], np.zeros(7, np.uint8)]))
    check_against_oracle(engine, [a, ragged], False, what="ragged index tail")
    # 37 runs of wildly different sizes (odd segment counts at every merge level)
    rng = np.random.default_rng(9)
    pool = nasty_keys(rng, 2000)
    runs = random_runs(rng, 37, [int(rng.integers(0, 200)) if r % 5 else 1500 for r in range(37)], pool)
    check_against_oracle(engine, runs, False, what="37 runs")


def test_truncated_and_corrupt_runs_end_silently(engine):
    """lsm_tree.rs:1014,1063: a run whose next record cannot be read/decoded simply ends."""
    a = sstable.build_run([(bytes([n]), b"A" * 40, 1) for n in range(100)])
    b = sstable.build_run([(bytes([n]), b"B" * 33, 2) for n in range(50, 150)])
    cut = int.from_bytes(bytes(b[1][16 * 30:16 * 30 + 8]), "little") + 5
    check_against_oracle(engine, [a, (b[0][:cut].copy(), b[1])], False, what="short data")
    assert engine.stats()["runs_truncated"] == 1 and engine.stats()["entries_valid"] == 130
    bad_idx = b[1].copy()
    bad_idx[16 * 20 + 12] += 1  # full_size off by one: decode sees trailing bytes
    check_against_oracle(engine, [a, (b[0], bad_idx)], False, what="bad full_size")
    bad_data = b[0].copy()
    off = int.from_bytes(bytes(b[1][16 * 10:16 * 10 + 8]), "little")
    bad_data[off] += 1  # klen prefix disagrees with the index
    check_against_oracle(engine, [(bad_data, b[1]), a], True, what="bad klen")
    first = b[1].copy()
    first[12] += 2  # very first record undecodable (trailing bytes): whole run contributes nothing
    check_against_oracle(engine, [a, (b[0], first)], False, what="first record bad")
    assert engine.stats()["entries_valid"] == 100


def test_index_fields_are_verified_not_ignored(engine):
    """Deliberate strictness (DESIGN.md "divergences"): the reference's compaction reader only
    consults full_size and takes klen from the .data bytes (lsm_tree.rs:1164-1168), so a wrong
    key_size / offset in .index goes unnoticed there -- while its own point reads
    (lsm_tree.rs:628-632) trust key_size.  The engine treats such a record as undecodable and
    ends the run at it, exactly like any other corrupt record."""
    a = sstable.build_run([(bytes([n]), b"A" * 40, 1) for n in range(100)])
    b = sstable.build_run([(bytes([n]), b"B" * 33, 2) for n in range(50, 150)])
    bad = b[1].copy()
    bad[16 * 40 + 8] += 1  # key_size of record 40
    gd, gi, _, n = engine.compact([a, (b[0], bad)], False)
    exp, en = model_compact([a, sstable.build_run([(bytes([n]), b"B" * 33, 2) for n in range(50, 90)])], False)
    assert n == en and engine.stats()["runs_truncated"] == 1
    assert_run_equal((gd, gi), exp, "key_size mismatch ends the run")


def test_long_shared_prefixes_and_window_ties(engine):
    """Keys that agree on the common prefix AND on the whole 11-byte window (full compare path)."""
    rng = np.random.default_rng(21)
    stem = b"tenant/000042/collection/users/"  # 31 bytes shared by every key
    mids = [b"AAAAAAAAAAAAAAAA", b"AAAAAAAAAAAAAAAB", b"AAAAAAAAAAA", b"AAAAAAAAAAAA", b"AAAAAAAAAAAB"]
    pool = sorted({stem + m + bytes(rng.integers(97, 100, int(rng.integers(0, 6)), dtype=np.uint8))
                   for m in mids for _ in range(200)})
    runs = random_runs(rng, 5, 150, pool)
    gd, gi, _, n = check_against_oracle(engine, runs, False, what="window ties")
    assert engine.stats()["key_prefix_len"] >= len(stem)
    exp, en = model_compact(runs, False)
    assert_run_equal((gd, gi), exp, "window ties vs model")


def test_bloom_small(engine):
    rng = np.random.default_rng(5)
    ents = [(b"\xb0k%015d" % n, bytes(rng.integers(0, 256, 90, dtype=np.uint8)), BASE_TS + n) for n in range(0, 6000, 2)]
    run_a, run_b = sstable.build_run(ents[:2000]), sstable.build_run(ents[1000:])
    gd, gi, gb, n = check_against_oracle(engine, [run_a, run_b], False, bloom_min_size=100_000, what="bloom")
    assert gb is not None and n == 3000
    assert all(oracle.bloom_check(gb, k) for k, _, _ in ents)
    total = run_a[0].size + run_b[0].size
    assert engine.compact([run_a, run_b], False, bloom_min_size=total, seed=SEED)[2] is None  # strict '>'
    assert engine.compact([run_a, run_b], False, bloom_min_size=total - 1, seed=SEED)[2] is not None
    # random seed (Bloom::new -> getrandom): still a valid filter for its own seed
    _, _, rb, _ = engine.compact([run_a, run_b], False, bloom_min_size=100_000, seed=None)
    assert all(oracle.bloom_check(rb, k) for k, _, _ in ents[:200])


def test_variable_key_lengths_in_bloom(engine):
    rng = np.random.default_rng(6)
    pool = nasty_keys(rng, 4000, max_len=70)
    runs = random_runs(rng, 4, 2500, pool, max_doc=200)
    check_against_oracle(engine, runs, True, bloom_min_size=1000, what="bloom nasty keys")


@pytest.mark.parametrize("cfg,keys", [(W.CFG2, 100_000), (W.CFG3, 20_000)])
def test_scaled_benchmark_shapes(engine, cfg, keys):
    c = W.scaled(cfg, keys)
    runs = W.make_merge_runs(c)
    _, _, gb, n = check_against_oracle(engine, runs, c.keep_tombstones, what=c.name)
    assert gb is not None
    eq = W.make_merge_runs(dataclasses.replace(c, seed=c.seed + 100), equal_ts=True)
    check_against_oracle(engine, eq, True, what=c.name + " equal-ts keep")


def test_unsorted_run_is_refused(engine):
    ents = [(b"\xb0k%015d" % n, b"v" * 20, BASE_TS) for n in range(200)]
    good = sstable.build_run(ents)
    swapped = list(ents)
    swapped[50], swapped[51] = swapped[51], swapped[50]
    with pytest.raises(capi.DbeelError) as ei:
        engine.compact([good, sstable.build_run(swapped)], False, flags=capi.FLAG_VERIFY_SORTED)
    assert ei.value.code == capi.ERR_UNSORTED_RUN
    engine.compact([good, good], False, flags=capi.FLAG_VERIFY_SORTED)  # sorted input passes the check


def test_capacity_and_argument_errors(engine):
    import ctypes as C
    run = sstable.build_run([(b"k", b"v", 1)])
    arr = (capi.Run * 1)(capi.Run(run[0].ctypes.data, run[0].size, run[1].ctypes.data, run[1].size))
    small = np.empty(8, np.uint8)
    out = capi.Out(small.ctypes.data, 8, 0, small.ctypes.data, 8, 0, None, 0, 0, 0)
    opts = capi.make_opts()
    assert capi.lib().dbeel_compact(engine._h, arr, 1, C.byref(opts), C.byref(out)) == capi.ERR_CAPACITY
    assert capi.lib().dbeel_compact(engine._h, None, 1, C.byref(opts), C.byref(out)) == capi.ERR_INVALID_ARG
    assert b"too small" in capi.lib().dbeel_strerror(capi.ERR_CAPACITY)


def test_device_resident_entry_point(engine):
    import torch
    c = W.scaled(W.CFG2, 50_000)
    runs = W.make_merge_runs(c)
    dev = torch.device("cuda:0")
    t_runs = [(torch.from_numpy(d).to(dev), torch.from_numpy(i).to(dev)) for d, i in runs]
    opts = capi.make_opts(False, seed=SEED)
    dc, ic, bc = capi.compact_bound([(d.size, i.size) for d, i in runs], opts)
    od = torch.empty(dc + 16, dtype=torch.uint8, device=dev)
    oi = torch.empty(ic + 16, dtype=torch.uint8, device=dev)
    ob = torch.empty(bc + 16, dtype=torch.uint8, device=dev)
    torch.cuda.synchronize()
    dl, il, bl, n = engine.compact_device([(d.data_ptr(), d.numel(), i.data_ptr(), i.numel()) for d, i in t_runs],
                                          (od.data_ptr(), dc, oi.data_ptr(), ic, ob.data_ptr(), bc), opts)
    ed, ei, eb, en = oracle.compact(runs, False, seed=SEED)
    assert n == en and bl == eb.size
    assert_run_equal((od[:dl].cpu().numpy(), oi[:il].cpu().numpy()), (ed, ei), "device path")
    assert np.array_equal(ob[:bl].cpu().numpy(), eb)
    st = engine.stats()
    assert st["ms_total"] > 0 and st["kernel_launches"] >= 10


def test_flush_matches_memtable(engine):
    """dbeel_flush == RedBlackTree inserts + flush_memtable_to_disk for one memtable's writes."""
    rng = np.random.default_rng(31)
    pool = nasty_keys(rng, 500)
    writes = []
    for s in range(3000):
        k = pool[int(rng.integers(len(pool)))]
        v = b"" if rng.random() < 0.1 else bytes(rng.integers(0, 256, int(rng.integers(1, 50)), dtype=np.uint8))
        writes.append((k, v, BASE_TS - s))  # decreasing timestamps: arrival order must win, not time
    batch = sstable.build_run(writes)
    gd, gi, n = engine.flush(batch)
    (od, oi, on), = oracle.memtable_flushes(batch, capacity=1 << 20)
    assert n == on
    assert_run_equal((gd, gi), (od, oi), "flush")
    assert_run_equal((gd, gi), model_flush(batch, 1 << 20)[0], "flush vs model")


@pytest.fixture()
def tiny_partition_engine(monkeypatch):
    """An engine whose host entry point cuts even small jobs into many key-range partitions."""
    monkeypatch.setenv("DBEEL_PIPELINE_MIN_KB", "1")
    monkeypatch.setenv("DBEEL_PARTITION_KB", "24")
    eng = capi.Engine(0)
    yield eng
    eng.close()


@pytest.mark.parametrize("seed", range(4))
def test_pipelined_host_path_is_byte_identical(tiny_partition_engine, seed):
    """dbeel_compact with host buffers runs as a pipeline of key-range partitions (H2D / kernels / D2H
    overlapped).  Same bytes as the oracle: keys never straddle partitions, offsets continue across them,
    one shared bloom filter."""
    eng = tiny_partition_engine
    rng = np.random.default_rng(300 + seed)
    pool = nasty_keys(rng, 5000, max_len=40)
    k = int(rng.integers(2, 10))
    runs = random_runs(rng, k, [int(rng.integers(0, 3000)) for _ in range(k)], pool, max_doc=120)
    keep = bool(seed & 1)
    check_against_oracle(eng, runs, keep, bloom_min_size=10_000, what=f"pipelined {seed}")
    st = eng.stats()
    assert st["kernel_launches"] > 60  # several partitions actually ran
    # equal timestamps everywhere: the run-position tie-break must survive partitioning
    c = W.scaled(W.CFG2, 3000)
    check_against_oracle(eng, W.make_merge_runs(c, equal_ts=True), False, bloom_min_size=10_000, what="pipelined equal ts")


def test_pipelined_path_falls_back_on_corrupt_input(tiny_partition_engine):
    eng = tiny_partition_engine
    a = sstable.build_run([(b"k%06d" % n, b"A" * 40, 1) for n in range(3000)])
    b = sstable.build_run([(b"k%06d" % n, b"B" * 33, 2) for n in range(1500, 4500)])
    bad = b[1].copy()
    bad[16 * 2000 + 12] += 1  # undecodable record in the middle of run b: everything after it is dropped
    check_against_oracle(eng, [a, (b[0], bad)], False, what="pipelined corrupt")
    assert eng.stats()["runs_truncated"] == 1


def test_flush_large_zipf_batch(engine):
    """cfg5-shaped arrivals: Zipf-skewed key ids, 512-byte docs, enough writes for a dozen sort tiles and
    four merge levels; cut into memtables of 3000 distinct keys exactly like set_ex would (lsm_tree.rs:757-765)."""
    from dbeel_b200 import storage_engine as se
    batch = W.make_arrival_batch(n_writes=30_000, n_ids=20_000, doc_bytes=512, seed=5)
    ents = sstable.parse_run(*batch)
    expected = oracle.memtable_flushes(batch, capacity=3000)
    pos = 0
    for od, oi, on in expected:
        n = se.memtable_cut(batch, pos, 3000)
        sub = sstable.build_run(ents[pos:pos + n])
        gd, gi, gn = engine.flush(sub)
        assert gn == on
        assert_run_equal((gd, gi), (od, oi), f"memtable at arrival {pos}")
        pos += n
    assert pos == len(ents) and len(expected) >= 2
    st = engine.stats()
    assert st["merge_passes"] >= 1


def test_flush_device_entry_point(engine):
    import torch
    rng = np.random.default_rng(77)
    writes = [(b"\xb0k%015d" % int(rng.integers(0, 3000)), bytes(rng.integers(0, 256, 64, dtype=np.uint8)), BASE_TS + s)
              for s in range(9000)]
    batch = sstable.build_run(writes)
    dev = torch.device("cuda:0")
    d, i = torch.from_numpy(batch[0]).to(dev), torch.from_numpy(batch[1]).to(dev)
    od = torch.empty(d.numel() + 16, dtype=torch.uint8, device=dev)
    oi = torch.empty(i.numel() + 16, dtype=torch.uint8, device=dev)
    torch.cuda.synchronize()
    dl, il, n = engine.flush_device((d.data_ptr(), d.numel(), i.data_ptr(), i.numel()), (od.data_ptr(), d.numel(), oi.data_ptr(), i.numel()))
    (ed, ei, en), = oracle.memtable_flushes(batch, capacity=1 << 20)
    assert n == en
    assert_run_equal((od[:dl].cpu().numpy(), oi[:il].cpu().numpy()), (ed, ei), "flush_device")
    assert engine.stats()["key_prefix_len"] >= 12  # the prefix reduction over ALL arrivals found '\xb0k' + zero digits


def test_async_submit_poll_wait(engine):
    """dbeel_compact_submit / dbeel_poll / dbeel_wait: the non-blocking form a reactor-driven caller uses."""
    import time
    runs = W.make_merge_runs(W.scaled(W.CFG2, 20_000))
    reap = engine.compact_async(runs, False, seed=SEED)
    with pytest.raises(capi.DbeelError) as ei:  # one job per engine
        engine.compact(runs, False, seed=SEED)
    assert ei.value.code == 10
    res = None
    deadline = time.time() + 60
    while res is None and time.time() < deadline:
        res = reap(False)
        time.sleep(0.001)
    assert res is not None
    od, oi, ob, on = oracle.compact(runs, False, seed=SEED)
    assert res[3] == on
    assert_run_equal((res[0], res[1]), (od, oi), "async")
    assert np.array_equal(res[2], ob)
    reap2 = engine.compact_async(runs, True, seed=SEED)
    r2 = reap2(True)  # blocking wait
    assert r2[3] == oracle.compact(runs, True, seed=SEED)[3]


def test_large_and_tiny_entries_mixed(engine):
    """Entry sizes from 32 bytes (empty key, tombstone) to ~64 KB (the reference's practical request cap,
    db_server.rs:399-403): gather tiles that hold hundreds of entries next to entries that span several tiles."""
    rng = np.random.default_rng(41)
    sizes = [0, 1, 15, 16, 17, 31, 32, 33, 255, 4095, 4096, 4097, 16383, 16384, 16385, 40000, 65000]
    runs = []
    for r in range(3):
        ents = []
        for n in range(400):
            k = b"key-%05d" % (3 * n + r % 2)
            dl = sizes[int(rng.integers(len(sizes)))]
            ents.append((k, bytes(rng.integers(0, 256, dl, dtype=np.uint8)), BASE_TS + r))
        if r == 0:
            ents.append((b"", b"", BASE_TS))  # 32-byte entry: empty key, tombstone
        runs.append(sstable.build_run(sorted(ents)))
    check_against_oracle(engine, runs, True, bloom_min_size=1000, what="mixed sizes keep")
    check_against_oracle(engine, runs, False, bloom_min_size=1000, what="mixed sizes drop")


def test_three_hundred_small_runs(engine):
    rng = np.random.default_rng(43)
    pool = [b"\xb0k%015d" % n for n in range(5000)]
    runs = random_runs(rng, 300, [int(rng.integers(0, 40)) for _ in range(300)], pool, max_doc=30)
    gd, gi, _, n = check_against_oracle(engine, runs, False, what="300 runs")
    assert engine.stats()["merge_passes"] == 9
    exp, en = model_compact(runs, False)
    assert_run_equal((gd, gi), exp, "300 runs vs model")


def test_common_prefix_longer_than_the_cap(engine):
    """Keys sharing 400 bytes: the skipped prefix is capped at 255, the 11-byte window sits inside the shared
    part, every comparison ties on it and goes through the full byte compare."""
    rng = np.random.default_rng(44)
    stem = bytes(rng.integers(1, 255, 400, dtype=np.uint8))
    pool = sorted({stem + bytes(rng.integers(0, 256, int(rng.integers(0, 12)), dtype=np.uint8)) for _ in range(600)})
    runs = random_runs(rng, 4, 250, pool, max_doc=40)
    gd, gi, _, _ = check_against_oracle(engine, runs, True, bloom_min_size=1000, what="long prefix")
    assert engine.stats()["key_prefix_len"] == 255
    exp, _ = model_compact(runs, True)
    assert_run_equal((gd, gi), exp, "long prefix vs model")


def test_flush_many_equals_separate_flushes(engine):
    """dbeel_flush_many: several memtables in one launch sequence -- each SSTable must equal what the red-black-tree
    replay writes for that memtable alone, offsets restarting at 0 per file."""
    from dbeel_b200 import storage_engine as se
    rng = np.random.default_rng(91)
    batches = []
    # (a) a Zipf stream cut at 3000 distinct keys (multi-tile memtables of different sizes)
    stream = W.make_arrival_batch(n_writes=30_000, n_ids=20_000, doc_bytes=200, seed=9)
    ents = sstable.parse_run(*stream)
    pos = 0
    while pos < len(ents):
        n = se.memtable_cut(stream, pos, 3000)
        batches.append(sstable.build_run(ents[pos:pos + n]))
        pos += n
    # (b) odd ones: an empty memtable, a single write, adversarial keys with many overwrites, all tombstones
    batches.append((np.zeros(0, np.uint8), np.zeros(0, np.uint8)))
    batches.append(sstable.build_run([(b"only", b"one", 5)]))
    pool = nasty_keys(rng, 200)
    batches.append(sstable.build_run([(pool[int(rng.integers(200))], bytes(rng.integers(0, 256, int(rng.integers(0, 40)), dtype=np.uint8)),
                                       BASE_TS - s) for s in range(2500)]))
    batches.append(sstable.build_run([(b"k%03d" % (s % 50), b"", s) for s in range(300)]))
    got = engine.flush_many(batches)
    assert len(got) == len(batches)
    for k, (b, (gd, gi, gn)) in enumerate(zip(batches, got)):
        if b[1].size == 0:
            assert gn == 0 and gd.size == 0 and gi.size == 0
            continue
        (od, oi, on), = oracle.memtable_flushes(b, capacity=1 << 20)
        assert gn == on, f"memtable {k}"
        assert_run_equal((gd, gi), (od, oi), f"memtable {k}")
    # and the same through separate calls
    for b, (gd, gi, gn) in zip(batches[:3], got[:3]):
        sd, si, sn = engine.flush(b)
        assert sn == gn
        assert_run_equal((sd, si), (gd, gi), "flush vs flush_many")


def _compact_many_check(engine, jobs, bloom_min_size, what):
    seeds = [bytes([(7 * j + k) % 256 for k in range(32)]) for j in range(len(jobs))]
    got = engine.compact_many(jobs, bloom_min_size=bloom_min_size, seeds=seeds)
    assert len(got) == len(jobs)
    for j, ((runs, keep), (gd, gi, gb, gn)) in enumerate(zip(jobs, got)):
        od, oi, ob, on = oracle.compact(runs, keep_tombstones=keep, bloom_min_size=bloom_min_size, seed=seeds[j])
        assert gn == on, f"{what} job {j}: items {gn} != {on}"
        assert_run_equal((gd, gi), (od, oi), f"{what} job {j}")
        assert (gb is None) == (ob is None), f"{what} job {j}: bloom presence"
        if ob is not None:
            assert np.array_equal(gb, ob), f"{what} job {j}: .bloom differs"
    return got


@pytest.mark.parametrize("seed", range(4))
def test_compact_many_equals_separate_compactions(engine, seed):
    """N1: many compactions in one launch sequence -- every job's three files equal a separate dbeel_compact
    (own run count, own keep_tombstones, own filter with its own seed, enabled from its own input size)."""
    rng = np.random.default_rng(500 + seed)
    jobs = []
    for j in range(int(rng.integers(2, 12))):
        pool = nasty_keys(rng, int(rng.integers(20, 900)), max_len=50)
        n_runs = int(rng.integers(1, 10))
        sizes = [int(rng.integers(0, len(pool) + 1)) for _ in range(n_runs)]
        runs = random_runs(rng, n_runs, sizes, pool, max_doc=int(rng.choice([20, 200, 1500])))
        jobs.append((runs, bool(rng.integers(2))))
    _compact_many_check(engine, jobs, bloom_min_size=int(rng.choice([1000, 30_000, 1 << 20])), what=f"many {seed}")


def test_compact_many_edge_shapes(engine):
    rng = np.random.default_rng(77)
    pool = nasty_keys(rng, 300, max_len=30)
    empty = sstable.build_run([])
    one = random_runs(rng, 1, 200, pool)
    jobs = [([], False), ([empty, empty], True), (one, False), (random_runs(rng, 7, 250, pool), False), ([empty] + one, True),
            (random_runs(rng, 16, 100, pool, tombstone_frac=0.9), False)]
    got = _compact_many_check(engine, jobs, bloom_min_size=2000, what="edges")
    assert got[0][3] == 0 and got[1][3] == 0 and got[0][0].size == 0
    # a single job through the batch entry point is the ordinary compaction
    _compact_many_check(engine, [jobs[3]], bloom_min_size=2000, what="single")
    assert engine.compact_many([]) == []
    # a corrupt run ends early inside its own job only (lsm_tree.rs:1014,1063)
    bad_d, bad_i = one[0][0].copy(), one[0][1].copy()
    bad_i[16 * 50 + 12:16 * 50 + 16] = np.frombuffer((7).to_bytes(4, "little"), np.uint8)
    _compact_many_check(engine, [(random_runs(rng, 3, 100, pool), False), ([(bad_d, bad_i)] + one, True)], 1 << 20, "truncated job")


def test_multi_megabyte_entries_and_keys(engine):
    """Far beyond the server's request cap: documents of several MB (one entry spans hundreds of gather tiles) and keys of
    100 KB that differ only in their last byte (the 12-byte comparison window never decides: every comparison walks the
    full keys), through compact, flush, flush_many and the WAL replay."""
    rng = np.random.default_rng(91)
    big = lambda n: bytes(rng.integers(0, 256, n, dtype=np.uint8))
    stem = big(100_000)
    keys = [stem + bytes([j]) for j in range(6)] + [b"a", b"b" * 70_000, b"b" * 70_000 + b"\x00"]
    runs = []
    for r in range(3):
        ents = [(k, big(int(rng.choice([0, 10, 3_000_000, 5_500_001]))), BASE_TS + r) for k in keys if rng.random() < 0.8]
        runs.append(sstable.build_run(sorted(ents)))
    check_against_oracle(engine, runs, True, bloom_min_size=1000, what="huge entries keep")
    check_against_oracle(engine, runs, False, bloom_min_size=1000, what="huge entries drop")
    arrivals = [(keys[int(rng.integers(len(keys)))], big(int(rng.choice([0, 7, 2_100_000]))), BASE_TS + j) for j in range(40)]
    batch = sstable.build_run(arrivals)
    od, oi, on = oracle.memtable_flushes(batch, capacity=8192)[0]
    gd, gi, gn = engine.flush(batch)
    assert gn == on
    assert_run_equal((gd, gi), (od, oi), "huge flush")
    (md, mi, mn), (sd, si, sn) = engine.flush_many([batch, sstable.build_run(arrivals[:5])])
    assert mn == on and np.array_equal(md, od) and np.array_equal(mi, oi) and sn <= 5
    wal = sstable.build_wal(arrivals, pad_byte=0x11)
    wd, wi, wn = engine.wal_flush(wal)
    assert wn == on and np.array_equal(wd, od) and np.array_equal(wi, oi)
    res = engine.get_many([(od, oi, None)], keys + [stem, stem + b"\xff"], capi.LOOKUP_EXACT)
    present = {k for k, _, _ in arrivals}
    assert [int(t) for t in res["table"]] == [0 if k in present else -1 for k in keys] + [-1, -1]



# ---- row N3: dbeel_compact_stream -- the same compaction with read / write callbacks instead of buffers

@pytest.mark.parametrize("seed", range(3))
def test_compact_stream_equals_compact_and_the_oracle(tiny_partition_engine, seed):
    """In-memory "files" behind ctypes callbacks: many partitions through the pinned rings, same bytes as the oracle."""
    eng = tiny_partition_engine
    rng = np.random.default_rng(900 + seed)
    pool = nasty_keys(rng, 5000, max_len=40)
    k = int(rng.integers(2, 9))
    runs = random_runs(rng, k, [int(rng.integers(0, 3000)) for _ in range(k)], pool, max_doc=120)
    keep = bool(seed & 1)
    exp = oracle.compact(runs, keep, bloom_min_size=10_000, seed=SEED)
    got = eng.compact_stream(runs, keep, bloom_min_size=10_000, seed=SEED)
    assert got[3] == exp[3]
    assert_run_equal(got[:2], exp[:2], f"stream {seed}")
    assert (got[2] is None) == (exp[2] is None) and (got[2] is None or np.array_equal(got[2], exp[2]))
    assert eng.stats()["partitions"] > 3
    # the reference's reader semantics (DBEEL_FLAG_REFERENCE_READER) through the same callbacks
    got = eng.compact_stream(runs, keep, bloom_min_size=10_000, seed=SEED, flags=capi.FLAG_REFERENCE_READER)
    assert_run_equal(got[:2], exp[:2], f"stream, reference reader {seed}")


def test_compact_stream_returns_the_callbacks_error_code(tiny_partition_engine):
    eng = tiny_partition_engine
    runs = W.make_merge_runs(W.scaled(W.CFG2, 3000))
    for kw in ({"fail_read_at": 0}, {"fail_read_at": 200}, {"fail_write_at": 0}, {"fail_write_at": 7}):
        with pytest.raises(capi.DbeelError) as ei:
            eng.compact_stream(runs, False, bloom_min_size=10_000, seed=SEED, **kw)
        assert ei.value.code == 4242, kw
    # the engine is usable afterwards
    exp = oracle.compact(runs, False, bloom_min_size=10_000, seed=SEED)
    got = eng.compact_stream(runs, False, bloom_min_size=10_000, seed=SEED)
    assert_run_equal(got[:2], exp[:2], "stream after failures")
