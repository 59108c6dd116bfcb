"""Row N4 (the step before the flush): dbeel_wal_flush vs the oracle's restatement of read_memtable_from_wal_file
(lsm_tree.rs:552-574) + the recovery flush (:478-513).  The SSTable must be byte-identical."""
import numpy as np
import pytest

import oracle
from dbeel_b200 import capi, sstable
from dbeel_b200 import workloads as W

from helpers import BASE_TS, assert_run_equal, nasty_keys

pytestmark = pytest.mark.gpu


def check(engine, wal, capacity=8192, what=""):
    gd, gi, gn = engine.wal_flush(wal, capacity)
    od, oi, on, _ = oracle.wal_flush(wal, capacity)
    assert gn == on, f"{what}: items {gn} != {on}"
    assert_run_equal((gd, gi), (od, oi), what)
    return gd, gi, gn


def test_replay_of_a_write_heavy_log(engine):
    """BASELINE configs[4] shape in small: Zipf-distributed writes, 512-byte documents, one memtable's worth."""
    ents = sstable.parse_run(*W.make_arrival_batch(30_000, 200_000, 512, seed=5))
    distinct = len({k for k, _, _ in ents})
    wal = sstable.build_wal(ents, pad_byte=0xCD)
    _, _, n = check(engine, wal, capacity=max(8192, distinct), what="zipf log")
    assert n == distinct
    st = engine.stats()
    assert st["kernel_launches"] > 20 and st["input_bytes"] == wal.size


@pytest.mark.parametrize("seed", range(4))
def test_adversarial_keys_sizes_and_padding(engine, seed):
    rng = np.random.default_rng(100 + seed)
    pool = nasty_keys(rng, 400, max_len=60)
    ents = []
    for j in range(3000):
        k = pool[int(rng.integers(len(pool)))]
        size = int(rng.choice([0, 1, 30, 500, 4000, 4096 - 32 - len(k), 9000, 20_000], p=[.1, .2, .3, .2, .05, .05, .05, .05]))
        ents.append((k, bytes(rng.integers(0, 256, max(0, size), dtype=np.uint8)), BASE_TS + int(rng.integers(-5, 5))))
    check(engine, sstable.build_wal(ents, pad_byte=int(rng.integers(0, 256))), what=f"nasty {seed}")


def test_records_that_hide_records(engine):
    """A multi-page record whose payload holds decodable records at its inner page boundaries: the chain must step over
    them (they are reachable only from a page the replay never starts at)."""
    fake = sstable.encode_entry(b"ghost", b"boo", 5)
    page = fake.ljust(4096, b"\x00")
    ents = [(b"aaa", b"x" * (4096 - 8 - 3 - 8) + page * 5 + b"tail", 10), (b"bbb", b"v", 11),
            (b"p" * (4096 - 32 - 100), b"q" * 100, 12), (b"ccc", page * 2, 13), (b"aaa", b"", 14), (b"ddd", b"z" * 8159, 15)]
    wal = sstable.build_wal(ents)
    gd, gi, n = check(engine, wal, what="hidden records")
    keys = [k for k, _, _ in sstable.parse_run(gd, gi)]
    assert b"ghost" not in keys and n == 5 and keys[0] == b"aaa"
    assert sstable.parse_run(gd, gi)[0][1] == b""  # the later tombstone of "aaa" won


def test_zero_pages_torn_tails_corrupt_lengths_bad_timestamps(engine):
    ents = [(b"a", b"1", 100), (b"b", b"2", 101), (b"c", b"3", 102), (b"d", b"4" * 5000, 103)]
    wal = sstable.build_wal(ents)
    gd, gi, n = check(engine, np.concatenate([wal, np.zeros(3 * 4096, np.uint8)]), what="zero pages")
    assert n == 5 and sstable.parse_run(gd, gi)[0] == (b"", b"", 0)
    for cut in (1, 7, 8, 20, 4096 - 1, 4096 + 8 + 1 + 7, 3 * 4096 + 100, wal.size - 1, wal.size - 4096):
        check(engine, wal[:cut], what=f"torn at {cut}")
    bad = wal.copy()
    bad[4096:4104] = np.frombuffer((1 << 40).to_bytes(8, "little"), np.uint8)
    _, _, n = check(engine, bad, what="corrupt klen")
    assert n == 1
    bad = wal.copy()
    bad[2 * 4096 + 8 + 1:2 * 4096 + 8 + 1 + 8] = np.frombuffer((wal.size).to_bytes(8, "little"), np.uint8)  # dlen of "c"
    _, _, n = check(engine, bad, what="corrupt dlen")
    assert n == 2
    ents2 = [(b"a", b"1", 100), (b"b", b"2", 1 << 100), (b"c", b"3", 102), (b"b", b"5", -(1 << 120)), (b"e", b"6", -5)]
    _, _, n = check(engine, sstable.build_wal(ents2), what="bad timestamps")
    assert n == 3
    assert engine.wal_flush(np.zeros(0, np.uint8))[2] == 0
    only_bad = sstable.build_wal([(b"x", b"y", 1 << 100)])
    assert engine.wal_flush(only_bad)[2] == 0 and oracle.wal_flush(only_bad)[2] == 0


def test_capacity_is_enforced_like_the_tree(engine):
    ents = [(b"k%03d" % j, b"v", j) for j in range(300)]
    wal = sstable.build_wal(ents + ents)
    assert engine.wal_flush(wal, capacity=300)[2] == 300
    with pytest.raises(capi.DbeelError) as ei:
        engine.wal_flush(wal, capacity=299)
    assert ei.value.code == capi.ERR_TREE_FULL
    with pytest.raises(oracle.OracleError, match="ReachedCapacity"):
        oracle.wal_flush(wal, capacity=299)


def test_device_resident_log(engine):
    import torch
    rng = np.random.default_rng(33)
    ents = [(b"\xb0k%015d" % int(rng.integers(0, 6000)), bytes(rng.integers(0, 256, 512, dtype=np.uint8)), BASE_TS + j)
            for j in range(20_000)]
    wal = sstable.build_wal(ents)
    dev = torch.device("cuda:0")
    d_wal = torch.from_numpy(wal).to(dev)
    od = torch.empty(wal.size // 4 + 64, dtype=torch.uint8, device=dev)
    oi = torch.empty(20_000 * 16 + 64, dtype=torch.uint8, device=dev)
    torch.cuda.synchronize()
    dl, il, n = engine.wal_flush_device(d_wal.data_ptr(), wal.size, (od.data_ptr(), od.numel() - 32, oi.data_ptr(), oi.numel() - 32))
    ed, ei, en, _ = oracle.wal_flush(wal)
    assert n == en
    assert_run_equal((od[:dl].cpu().numpy(), oi[:il].cpu().numpy()), (ed, ei), "device log")
    # and the same bytes the ordinary flush of the same arrivals produces
    fd, fi, fn = engine.flush(sstable.build_run(ents))
    assert fn == n and np.array_equal(fd, ed) and np.array_equal(fi, ei)


def test_engines_give_their_device_memory_back():
    """Every grow-only scratch buffer of an engine (compaction workspace, WAL replay tables, routing tables, staging) is freed
    by dbeel_engine_destroy: create / use / destroy in a loop must not move cudaMemGetInfo's free figure."""
    import torch
    from dbeel_b200 import workloads as W
    rng = np.random.default_rng(2)
    ents = [(b"k%06d" % int(rng.integers(0, 3000)), bytes(rng.integers(0, 256, 200, dtype=np.uint8)), BASE_TS + j) for j in range(6000)]
    wal = sstable.build_wal(ents)
    runs = W.make_merge_runs(W.scaled(W.CFG2, 5000))

    def cycle():
        eng = capi.Engine(0)
        eng.wal_flush(wal, capacity=8192)
        eng.compact(runs, False, seed=bytes(32))
        eng.close()

    cycle()  # first use may load modules / grow the context's own pools
    torch.cuda.synchronize()
    free0, _ = torch.cuda.mem_get_info(0)
    for _ in range(5):
        cycle()
    torch.cuda.synchronize()
    free1, _ = torch.cuda.mem_get_info(0)
    assert free0 - free1 < (8 << 20), f"device memory leaked: {free0 - free1} bytes over 5 engine lifetimes"
