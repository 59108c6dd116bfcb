"""DBEEL_FLAG_REFERENCE_READER: the input runs are decoded exactly like the reference's sequential reader
(read_next_entry, lsm_tree.rs:1158-1170): `offset` and `key_size` of an index record are ignored, the key length comes from
the .data bytes, and a timestamp outside `time`'s range ends the run (utils/timestamp_nanos.rs:15-24).  Every case here is
compared with the C oracle, which restates that reader line by line (oracle/dbeel_oracle.c: read_next_entry, entry_decode).
"""
import numpy as np
import pytest

import oracle
from dbeel_b200 import capi, sstable
from dbeel_b200 import workloads as W

from helpers import BASE_TS, assert_run_equal, nasty_keys, random_runs

pytestmark = pytest.mark.gpu

SEED = bytes(range(32))
REF = capi.FLAG_REFERENCE_READER

# time 0.3: seconds must lie in -9999-01-01T00:00:00Z ..= 9999-12-31T23:59:59Z
TS_MAX = 253402300799 * 10**9 + 999_999_999
TS_MIN = -377705116800 * 10**9


def check(engine, runs, keep=False, bloom_min_size=capi.DEFAULT_BLOOM_MIN_SIZE, what=""):
    gd, gi, gb, gn = engine.compact(runs, keep_tombstones=keep, bloom_min_size=bloom_min_size, seed=SEED, flags=REF)
    od, oi, ob, on = oracle.compact(runs, keep_tombstones=keep, bloom_min_size=bloom_min_size, seed=SEED)
    assert gn == on, f"{what}: items_written {gn} != {on}"
    assert_run_equal((gd, gi), (od, oi), what)
    assert (gb is None) == (ob is None)
    if ob is not None:
        assert np.array_equal(gb, ob), f"{what}: .bloom differs"
    return gd, gi, gn


def two_runs():
    a = sstable.build_run([(bytes([n]), b"A" * 40, 1) for n in range(100)])
    b = sstable.build_run([(bytes([n]), b"B" * (20 + n % 7), 2) for n in range(50, 150)])
    return a, b


def test_wrong_key_size_is_ignored(engine):
    a, b = two_runs()
    bad = b[1].copy()
    bad[16 * 40 + 8] += 1  # key_size of record 40: the reference never reads it
    gd, gi, n = check(engine, [a, (b[0], bad)], what="key_size ignored")
    assert n == 150 and engine.stats()["index_repaired"] == 1 and engine.stats()["runs_truncated"] == 0
    # the output .index carries the key_size EntryWriter computes from the entry itself (entry_writer.rs:76-86)
    assert sstable.parse_run(gd, gi)[90][0] == bytes([90])
    # default mode is the strict one: the run ends at record 40
    _, _, _, n_strict = engine.compact([a, (b[0], bad)], False)
    assert n_strict == 100  # b contributes keys 50..89 only, all of them also in a


def test_wrong_offsets_are_ignored(engine):
    a, b = two_runs()
    rng = np.random.default_rng(3)
    bad = b[1].copy()
    for rec in (0, 7, 8, 55, 99):  # garbage in the offset field, including the very first and the very last record
        bad[16 * rec:16 * rec + 8] = rng.integers(0, 256, 8, dtype=np.uint8)
    _, _, n = check(engine, [(b[0], bad), a], keep=True, what="offsets ignored")
    assert n == 150 and engine.stats()["index_repaired"] == 1
    zeros = b[1].copy()
    zeros.reshape(-1, 16)[:, :8] = 0  # every offset zero: the stream cursor alone positions the reads
    check(engine, [a, (b[0], zeros)], what="all offsets zero")


def test_full_size_is_what_positions_the_stream(engine):
    """A wrong full_size derails the sequential reader from that record on (trailing bytes / garbage lengths): the run ends
    there in both implementations, no matter what the other index fields say."""
    a, b = two_runs()
    bad = b[1].copy()
    bad[16 * 20 + 12] += 3   # full_size of record 20
    bad[16 * 60 + 8] ^= 0x55  # and a wrong key_size further on (never reached)
    _, _, n = check(engine, [a, (b[0], bad)], what="bad full_size")
    assert engine.stats()["runs_truncated"] == 1
    short = (b[0][:b[0].size - 9].copy(), b[1])  # last record cut short: read_exact fails
    check(engine, [short, a], what="short data")
    klen = b[0].copy()
    off = int.from_bytes(bytes(b[1][16 * 33:16 * 33 + 8]), "little")
    klen[off] += 1  # the length prefix in .data itself is wrong: bincode sees trailing / missing bytes
    check(engine, [(klen, b[1]), a], what="bad klen in data")
    huge = b[0].copy()
    huge[off:off + 8] = 0xFF  # klen = 2^64 - 1
    check(engine, [(huge, b[1]), a], what="huge klen in data")


@pytest.mark.parametrize("ts,valid", [(TS_MAX, True), (TS_MAX + 1, False), (TS_MIN, True), (TS_MIN - 1, False),
                                       (-1, True), (1 << 64, True), ((1 << 100), False), (-(1 << 100), False),
                                       ((1 << 127) - 1, False), (-(1 << 127), False)])
def test_timestamp_range_ends_the_run(engine, ts, valid):
    """utils/timestamp_nanos.rs:15-24: OffsetDateTime::from_unix_timestamp_nanos fails outside +-9999 years, the
    deserialize error ends the run (lsm_tree.rs:1014,1063)."""
    a = sstable.build_run([(bytes([n]), b"A" * 40, 1) for n in range(100)])
    ents = [(bytes([n]), b"B" * 33, 2) for n in range(50, 150)]
    ents[30] = (ents[30][0], b"odd one", ts)
    b = sstable.build_run(ents)
    _, _, n = check(engine, [a, b], what=f"ts {ts}")
    st = engine.stats()
    assert st["entries_valid"] == (200 if valid else 130)
    assert n == (150 if valid else 100)  # cut at record 30, b contributes keys 50..79 only


def test_timestamp_of_unique_and_duplicate_keys(engine):
    """The range check applies to every entry the reader decodes, whether or not its key occurs twice."""
    rng = np.random.default_rng(11)
    pool = nasty_keys(rng, 800, max_len=30)
    runs = random_runs(rng, 5, 400, pool)
    # poison one entry per run at different depths
    poisoned = []
    for r, (d, i) in enumerate(runs):
        ents = sstable.parse_run(d, i)
        j = 50 + 60 * r
        ents[j] = (ents[j][0], ents[j][1], TS_MAX + 1 + r)
        poisoned.append(sstable.build_run(ents))
    check(engine, poisoned, keep=True, what="poisoned timestamps")
    assert engine.stats()["runs_truncated"] == 5


def test_clean_inputs_take_the_fast_path_and_match_the_default_mode(engine):
    c = W.scaled(W.CFG2, 30_000)
    runs = W.make_merge_runs(c)
    gd, gi, n = check(engine, runs, what="clean cfg2")
    assert engine.stats()["index_repaired"] == 0
    dd, di, _, dn = engine.compact(runs, False, seed=SEED)
    assert dn == n and np.array_equal(dd, gd) and np.array_equal(di, gi)


def test_random_index_damage_matches_the_oracle(engine):
    rng = np.random.default_rng(77)
    pool = nasty_keys(rng, 1500, max_len=40)
    for trial in range(12):
        k = int(rng.integers(1, 7))
        runs = random_runs(rng, k, [int(rng.integers(1, 900)) for _ in range(k)], pool)
        damaged = []
        for d, i in runs:
            i = i.copy()
            nrec = i.size // 16
            for _ in range(int(rng.integers(0, 4))):
                rec = int(rng.integers(nrec))
                field = int(rng.integers(3))  # 0 offset, 1 key_size, 2 full_size
                lo, hi = ((0, 8), (8, 12), (12, 16))[field]
                i[16 * rec + int(rng.integers(lo, hi))] ^= int(rng.integers(1, 256))
            damaged.append((d, i))
        check(engine, damaged, keep=bool(trial & 1), what=f"damage trial {trial}")


def test_pipelined_host_path_falls_back(monkeypatch):
    """Key-range partitions are cut by index offsets on the host; when the device finds the index inconsistent the job is
    redone in one piece with the reference semantics."""
    monkeypatch.setenv("DBEEL_PIPELINE_MIN_KB", "1")
    monkeypatch.setenv("DBEEL_PARTITION_KB", "24")
    eng = capi.Engine(0)
    try:
        rng = np.random.default_rng(5)
        pool = nasty_keys(rng, 3000, max_len=40)
        runs = random_runs(rng, 4, 2000, pool)
        check(eng, runs, what="clean, partitioned")
        assert eng.stats()["partitions"] > 1
        d, i = runs[2]
        bad = i.copy()
        bad[16 * 1500 + 8] += 2  # key_size deep inside some partition
        check(eng, runs[:2] + [(d, bad)] + runs[3:], what="damaged, partitioned")
        assert eng.stats()["partitions"] == 1 and eng.stats()["index_repaired"] == 1
    finally:
        eng.close()
