"""Host-side mirror of the file protocol around the GPU path (include/dbeel_tree.h).

CPU tests: the picker (tasks/compaction.rs:35-102), the memtable cut (rbtree_arena set semantics),
journal replay on open (lsm_tree.rs:424-438,576-590).  GPU tests: LSMTree.compact / flush on real
files, restating lsm_tree.rs:1328-1451 (get_after_compaction) at the file level."""
import ctypes as C
import os

import numpy as np
import pytest

import oracle
from dbeel_b200 import capi, sstable, storage_engine as se

from helpers import BASE_TS, assert_run_equal, model_flush, nasty_keys

SEED = bytes(range(32))


def u16key(n):
    return int(n).to_bytes(2, "little")


def test_tree_symbols_exported():
    lib = capi.lib()
    assert all(hasattr(lib, n) for n in se.TREE_EXPORTS)


def test_memtable_cut_counts_distinct_keys():
    rng = np.random.default_rng(2)
    pool = nasty_keys(rng, 200)
    writes = [(pool[int(rng.integers(len(pool)))], b"v", BASE_TS + s) for s in range(3000)]
    batch = sstable.build_run(writes)
    pos, cuts = 0, []
    while pos < len(writes):
        n = se.memtable_cut(batch, pos, 64)
        assert n > 0
        cuts.append(n)
        pos += n
    # same boundaries as the red-black-tree replay
    exp, pos, seen = [], 0, set()
    for s, (k, _, _) in enumerate(writes):
        seen.add(k)
        if len(seen) == 64:
            exp.append(s + 1 - pos)
            pos, seen = s + 1, set()
    if pos < len(writes):
        exp.append(len(writes) - pos)
    assert cuts == exp
    flushed = oracle.memtable_flushes(batch, capacity=64)
    assert len(flushed) == len(cuts) and all(n <= 64 for _, _, n in flushed)


def test_picker_matches_compaction_rs():
    # three flushed memtables of 32 entries, factor 2: one group [0,2,4] -> output 1, final level drops tombstones
    assert se.plan_compactions([(0, 32), (2, 32), (4, 32)], 2) == [([0, 2, 4], 1, False)]
    # factor above the group size: nothing to do (compaction.rs:83-87); factor < 2 disables (:105-108)
    assert se.plan_compactions([(0, 32), (2, 32), (4, 32)], 4) == []
    assert se.plan_compactions([(0, 32), (2, 32)], 1) == []
    # output index = max odd + 2 (compaction.rs:38-43)
    assert se.plan_compactions([(5, 92), (6, 32), (8, 32)], 2)[0][1] == 7
    # two tiers: 8 tables of 8192 (order lz=50) and 2 of 65536 (lz=47); the small tier sums to 65536 -> promoted
    tables = [(2 * i, 8192) for i in range(8)] + [(101, 65536), (103, 65536)]
    plan = se.plan_compactions(tables, 2)
    assert len(plan) == 1 and sorted(plan[0][0]) == sorted(i for i, _ in tables) and plan[0][2] is False
    # no promotion: 3 tables of 8192 (sum 24576, lz 49 < 50 -> promoted to order 49 alone) + 2 big ones
    plan = se.plan_compactions([(0, 8192), (2, 8192), (4, 8192), (101, 1 << 20), (103, 1 << 20)], 2)
    assert [sorted(g[0]) for g in plan] == [[101, 103], [0, 2, 4]]
    assert [g[2] for g in plan] == [False, True]  # only the largest tier drops tombstones


def _fake_engine():
    return C.c_void_p(1)  # dbeel_tree_open never dereferences the engine


def test_open_discovers_sstables_and_replays_journal(tmp_path):
    d = str(tmp_path)
    run = sstable.build_run([(u16key(n), b"v", 1) for n in range(10)])
    sstable.write_run_files(d, 0, run)
    sstable.write_run_files(d, 2, run)
    # a crash after the journal was written but before the renames: compact_* files + journal on disk
    out = sstable.build_run([(u16key(n), b"w", 2) for n in range(10)])
    out[0].tofile(os.path.join(d, sstable.file_name(1, sstable.COMPACT_DATA_FILE_EXT)))
    out[1].tofile(os.path.join(d, sstable.file_name(1, sstable.COMPACT_INDEX_FILE_EXT)))

    def path(i, ext):
        return os.path.join(d, sstable.file_name(i, ext)).encode()

    def s(b):
        return len(b).to_bytes(8, "little") + b

    renames = [(path(1, "compact_data"), path(1, "data")), (path(1, "compact_index"), path(1, "index")),
               (path(1, "compact_bloom"), path(1, "bloom"))]
    deletes = [path(i, e) for i in (0, 2) for e in ("data", "index", "bloom")]
    journal = (len(renames).to_bytes(8, "little") + b"".join(s(a) + s(b) for a, b in renames)
               + len(deletes).to_bytes(8, "little") + b"".join(s(x) for x in deletes))
    open(os.path.join(d, sstable.file_name(1, sstable.COMPACT_ACTION_FILE_EXT)), "wb").write(journal)

    t = se.LSMTree(d, _fake_engine())
    assert t.sstable_indices_and_sizes() == [(1, 10)]  # inputs deleted, output renamed into place
    assert t.write_sstable_index == 2  # lsm_tree.rs:461-465: odd max index i -> i + 1
    assert sorted(os.listdir(d)) == [sstable.file_name(1, "data"), sstable.file_name(1, "index")]
    assert_run_equal(sstable.read_run_files(d, 1), out)
    t.close()
    t = se.LSMTree(d, _fake_engine())  # idempotent
    assert t.sstable_indices_and_sizes() == [(1, 10)]


def test_missing_sstable_is_an_error(tmp_path):
    t = se.LSMTree(str(tmp_path), _fake_engine())
    assert t.sstable_indices_and_sizes() == [] and t.write_sstable_index == 0
    with pytest.raises(capi.DbeelError) as ei:
        t.compact([0, 2], 1, False)
    assert ei.value.code == se.ERR_NO_SSTABLE


@pytest.mark.gpu
def test_get_after_compaction_on_files(engine, tmp_path):
    """lsm_tree.rs:1400-1446 with real files: three flushes (0,32),(2,32),(4,32) -> compact(&[0,2,4], 5, false)."""
    d = str(tmp_path)
    tree = se.LSMTree.open_or_create(d, engine)
    assert tree.write_sstable_index == 0 and tree.sstable_indices_and_sizes() == []
    writes = [(u16key(n), u16key(n), BASE_TS + n) for n in range(94)]
    writes += [(u16key(1), b"", BASE_TS + 1000), (u16key(4), b"", BASE_TS + 1001)]
    batch = sstable.build_run(writes)
    pos = 0
    while pos < len(writes):
        n = se.memtable_cut(batch, pos, 32)
        lo = int.from_bytes(bytes(batch[1][16 * pos:16 * pos + 8]), "little")
        hi = batch[0].size if pos + n == len(writes) else int.from_bytes(bytes(batch[1][16 * (pos + n):16 * (pos + n) + 8]), "little")
        sub = sstable.build_run(sstable.parse_run(batch[0], batch[1])[pos:pos + n])
        assert sub[0].size == hi - lo
        tree.flush(sub)
        pos += n
    assert tree.sstable_indices_and_sizes() == [(0, 32), (2, 32), (4, 32)]
    for (gd, gi), (od, oi, _) in zip([sstable.read_run_files(d, i) for i in (0, 2, 4)],
                                     oracle.memtable_flushes(batch, capacity=32)):
        assert_run_equal((gd, gi), (od, oi), "flushed sstable")
    inputs = [sstable.read_run_files(d, i) for i in (0, 2, 4)]
    tree.compact([0, 2, 4], 5, False)
    assert tree.sstable_indices_and_sizes() == [(5, 32 * 3 - 4)] and tree.write_sstable_index == 6
    assert sorted(os.listdir(d)) == [sstable.file_name(5, "data"), sstable.file_name(5, "index")]  # no bloom: <= 1 MiB
    od, oi, _, on = oracle.compact(inputs, False)
    assert_run_equal(sstable.read_run_files(d, 5), (od, oi), "compacted files")
    # validate_tree_after_compaction's reads (:1386-1390), batched through the GPU read path, both search modes
    for mode in (capi.LOOKUP_REFERENCE, capi.LOOKUP_EXACT):
        assert tree.get_many([u16key(0), u16key(2), u16key(10), u16key(1), u16key(4)], mode) == \
            [u16key(0), u16key(2), u16key(10), None, None]
    assert tree.get_many([u16key(n) for n in range(94)], capi.LOOKUP_EXACT) == \
        [None if n in (1, 4) else u16key(n) for n in range(94)]
    tree.close()
    tree = se.LSMTree.open_or_create(d, engine)  # reopening the tree (:1439-1443)
    assert tree.sstable_indices_and_sizes() == [(5, 92)] and tree.write_sstable_index == 6
    assert tree.get_many([u16key(10), u16key(4)]) == [u16key(10), None]


@pytest.mark.gpu
def test_compact_tree_writes_bloom_file(engine, tmp_path):
    d = str(tmp_path)
    rng = np.random.default_rng(8)
    tree = se.LSMTree(d, engine, sstable_bloom_min_size=50_000)
    runs = []
    for r in range(4):
        ents = [(b"\xb0k%015d" % n, bytes(rng.integers(0, 256, 120, dtype=np.uint8)), BASE_TS + r)
                for n in sorted(rng.choice(5000, 800, replace=False).tolist())]
        run = sstable.build_run(ents)
        runs.append(run)
        sstable.write_run_files(d, 2 * r, run)
    tree.close()
    tree = se.LSMTree(d, engine, sstable_bloom_min_size=50_000)
    plan = tree.compact_tree(compaction_factor=2, bloom_seed=SEED)
    assert plan == [([0, 2, 4, 6], 1, False)]
    od, oi, ob, on = oracle.compact(runs, False, bloom_min_size=50_000, seed=SEED)
    assert tree.sstable_indices_and_sizes() == [(1, on)]
    assert_run_equal(sstable.read_run_files(d, 1), (od, oi), "compact_tree output")
    bloom = np.fromfile(os.path.join(d, sstable.file_name(1, "bloom")), dtype=np.uint8)
    assert np.array_equal(bloom, ob)
    assert sorted(os.listdir(d)) == [sstable.file_name(1, e) for e in ("bloom", "data", "index")]
    # reads through the filter the compaction just wrote (SSTable::new_with_bloom_read, lsm_tree.rs:94-101)
    want = {k: v for k, v, _ in sstable.parse_run(od, oi)}
    probe = [b"\xb0k%015d" % n for n in range(0, 5000, 7)]
    assert tree.get_many(probe, capi.LOOKUP_EXACT) == [want.get(k) for k in probe]


@pytest.mark.gpu
def test_cfg5_pipeline_flush_then_tiered_compactions(engine, tmp_path):
    """BASELINE.json configs[4] in miniature: a Zipf write stream is cut into memtables (distinct-key capacity),
    every memtable is flushed to the next even index, and after every flush the size-tiered picker runs
    (tasks/compaction.rs:104-137).  The recorded plan is replayed on the CPU oracle; every file that is left
    must match byte for byte."""
    d = str(tmp_path)
    cap, factor = 2048, 4
    batch = W_arrivals = __import__("dbeel_b200.workloads", fromlist=["x"]).make_arrival_batch(
        n_writes=40_000, n_ids=24_000, doc_bytes=100, seed=55)
    ents = sstable.parse_run(*batch)
    flushed = oracle.memtable_flushes(batch, capacity=cap)
    tree = se.LSMTree(d, engine, sstable_bloom_min_size=200_000)
    model = {}  # index -> (data, index, bloom|None) as the oracle would have them
    pos = 0
    n_compactions = 0
    for od, oi, on in flushed:
        n = se.memtable_cut(batch, pos, cap)
        idx, items = tree.flush(sstable.build_run(ents[pos:pos + n]))
        assert items == on
        model[idx] = (od, oi, None)
        pos += n
        for indices, out_idx, keep in tree.compact_tree(compaction_factor=factor, bloom_seed=SEED):
            runs = [(model[i][0], model[i][1]) for i in indices]
            cd, ci, cb, cn = oracle.compact(runs, keep, bloom_min_size=200_000, seed=SEED)
            for i in indices:
                del model[i]
            model[out_idx] = (cd, ci, cb)
            n_compactions += 1
    assert pos == len(ents) and n_compactions >= 2
    assert [i for i, _ in tree.sstable_indices_and_sizes()] == sorted(model)
    for i, (md, mi, mb) in model.items():
        assert_run_equal(sstable.read_run_files(d, i), (md, mi), f"sstable {i}")
        bloom_path = os.path.join(d, sstable.file_name(i, "bloom"))
        assert os.path.exists(bloom_path) == (mb is not None)
        if mb is not None:
            assert np.array_equal(np.fromfile(bloom_path, dtype=np.uint8), mb)
    # nothing but live SSTable files is left behind (journals and compact_* temporaries are gone)
    assert all(f.split(".")[1] in ("data", "index", "bloom") for f in os.listdir(d))


def test_wal_recovery_without_an_unflushed_log_needs_no_gpu(tmp_path):
    """0 or 1 `.memtable` files: nothing to replay (lsm_tree.rs:478-480); 3 is the reference's panic (:513)."""
    t = se.LSMTree(str(tmp_path), _fake_engine())
    assert t.recover_wal() == (0, 0)
    open(os.path.join(str(tmp_path), sstable.file_name(6, "memtable")), "wb").close()
    assert t.recover_wal() == (6, 0)
    for i in (8, 10):
        open(os.path.join(str(tmp_path), sstable.file_name(i, "memtable")), "wb").close()
    with pytest.raises(capi.DbeelError):
        t.recover_wal()


@pytest.mark.gpu
def test_unflushed_log_is_replayed_and_flushed_on_open(engine, tmp_path):
    """lsm_tree.rs:481-511: two logs on disk -> the older one becomes an SSTable under the NEWER log's index (the
    reference's own choice, :491-492) and is removed; the files equal the oracle's replay + flush."""
    d = str(tmp_path)
    rng = np.random.default_rng(8)
    ents = [(b"user%04d" % int(rng.integers(0, 700)), bytes(rng.integers(0, 256, int(rng.integers(0, 900)), dtype=np.uint8)),
             BASE_TS + j) for j in range(2500)]
    wal = sstable.build_wal(ents, pad_byte=0x5A)
    with open(os.path.join(d, sstable.file_name(4, "memtable")), "wb") as f:
        f.write(wal.tobytes())
    with open(os.path.join(d, sstable.file_name(6, "memtable")), "wb") as f:
        f.write(sstable.build_wal(ents[:3]).tobytes())
    tree = se.LSMTree.open_or_create(d, engine)
    current, items = tree.recover_wal()
    od, oi, on, _ = oracle.wal_flush(wal)
    assert (current, items) == (6, on)
    assert_run_equal(sstable.read_run_files(d, 6), (od, oi), "recovered sstable")
    assert not os.path.exists(os.path.join(d, sstable.file_name(4, "memtable")))
    assert os.path.exists(os.path.join(d, sstable.file_name(6, "memtable")))
    assert tree.sstable_indices_and_sizes() == []  # the list was built before the recovery (:440-459): next open sees it
    tree.close()
    tree = se.LSMTree.open_or_create(d, engine)
    assert tree.sstable_indices_and_sizes() == [(6, on)] and tree.recover_wal() == (6, 0)
    # a log with more distinct keys than the memtable holds cannot be replayed (memtable.set(..)? -> ReachedCapacity)
    with open(os.path.join(d, sstable.file_name(2, "memtable")), "wb") as f:  # older than the active log 6
        f.write(wal.tobytes())
    with pytest.raises(capi.DbeelError) as ei:
        tree.recover_wal(tree_capacity=100)
    assert ei.value.code == capi.ERR_TREE_FULL
    assert os.path.exists(os.path.join(d, sstable.file_name(2, "memtable")))  # nothing was removed


@pytest.mark.gpu
def test_compact_tree_batched_equals_group_by_group(engine, tmp_path):
    """compact_tree's groups (tasks/compaction.rs:82-101) through one dbeel_compact_many + per-group commit: the
    directory ends up byte-identical to running LSMTree::compact once per group."""
    rng = np.random.default_rng(31)
    sizes = {0: 100, 2: 100, 4: 100, 101: 5000, 103: 5000}
    tables = {}
    for idx, n in sizes.items():
        ents = [(b"\xb0k%015d" % k, b"" if rng.random() < 0.1 else bytes(rng.integers(0, 256, 150, dtype=np.uint8)), BASE_TS + idx)
                for k in sorted(rng.choice(20_000, n, replace=False).tolist())]
        tables[idx] = sstable.build_run(ents)
    dirs = [str(tmp_path / "seq"), str(tmp_path / "batched")]
    plans = []
    for d, batched in zip(dirs, (False, True)):
        os.makedirs(d)
        for idx, run in tables.items():
            sstable.write_run_files(d, idx, run)
        tree = se.LSMTree(d, engine, sstable_bloom_min_size=100_000)
        plans.append(tree.compact_tree(compaction_factor=2, bloom_seed=SEED, batched=batched))
        assert tree.sstable_indices_and_sizes() == sorted(tree.sstable_indices_and_sizes())
        tree.close()
    assert plans[0] == plans[1] and [sorted(g[0]) for g in plans[0]] == [[101, 103], [0, 2, 4]]
    assert [g[2] for g in plans[0]] == [False, True]
    names = sorted(os.listdir(dirs[0]))
    assert names == sorted(os.listdir(dirs[1])) and any(n.endswith(".bloom") for n in names)
    for n in names:
        a, b = (np.fromfile(os.path.join(d, n), dtype=np.uint8) for d in dirs)
        assert np.array_equal(a, b), n
    for indices, out, keep in plans[1]:  # and both equal the oracle's compaction of each group
        od, oi, ob, on = oracle.compact([tables[i] for i in indices], keep, bloom_min_size=100_000, seed=SEED)
        assert_run_equal(sstable.read_run_files(dirs[1], out), (od, oi), f"group -> {out}")
        bp = os.path.join(dirs[1], sstable.file_name(out, "bloom"))
        assert os.path.exists(bp) == (ob is not None)
        if ob is not None:
            assert np.array_equal(np.fromfile(bp, dtype=np.uint8), ob)


def _entry_writer_model(entries, files_index):
    """entry_writer.rs:71-156 restated in Python: write() per entry, write_to_cache() per stream, close()."""
    PAGE = 4096
    seq = []
    state = {se.FILE_DATA: [bytearray(PAGE), 0], se.FILE_INDEX: [bytearray(PAGE), 0]}

    def write_to_cache(raw, kind):
        st = state[kind]
        for c0 in range(0, len(raw), PAGE):
            chunk = raw[c0:c0 + PAGE]
            off = st[1] % PAGE
            end = min(off + len(chunk), PAGE)
            st[0][off:end] = chunk[:end - off]
            first = end - off
            st[1] += first
            if st[1] % PAGE == 0:
                seq.append((kind, files_index, st[1] - PAGE, bytes(st[0])))
                st[0] = bytearray(PAGE)
                left = len(chunk) - first
                st[0][:left] = chunk[first:]
                st[1] += left

    for k, v, ts in entries:
        rec = sstable.encode_entry(k, v, ts)
        idx = sstable.encode_index_record(state[se.FILE_DATA][1], len(k), len(rec))
        write_to_cache(rec, se.FILE_DATA)
        write_to_cache(idx, se.FILE_INDEX)
    for kind in (se.FILE_DATA, se.FILE_INDEX):  # close()
        left = state[kind][1] % PAGE
        if left:
            seq.append((kind, files_index, state[kind][1] - left, bytes(state[kind][0])))
    return seq


def test_out_pages_replays_entry_writers_write_through():
    """dbeel_out_pages vs a literal restatement of EntryWriter::{write, write_to_cache, close}: same pages, same keys, same
    order -- including entries larger than a page, entries that end exactly on a page boundary and the zero-padded tails
    (the reference's own entry_writer_cache_equals_disk test, lsm_tree.rs:1489-1556, checks pages against stream bytes)."""
    rng = np.random.default_rng(3)
    cases = [
        [(b"k%04d" % i, bytes(rng.integers(0, 256, int(rng.integers(0, 300)), dtype=np.uint8)), 7 + i) for i in range(700)],
        [(b"big%d" % i, bytes(rng.integers(0, 256, 9000 + 517 * i, dtype=np.uint8)), i) for i in range(5)],
        [(b"a", b"x" * (4096 - 32 - 1), 1), (b"b", b"y" * (4096 - 32 - 1), 2), (b"c", b"", 3)],  # entries of exactly one page
        [(b"only", b"", 1)],
        [(b"i%05d" % i, b"", i) for i in range(300)],  # 300 index records: the index stream crosses a page too
    ]
    for n, ents in enumerate(cases):
        d, i = sstable.build_run(ents)
        got = se.out_pages(d, i, 40 + n)
        assert got == _entry_writer_model(ents, 40 + n), f"case {n}"
        # and the pages are the files, zero-padded
        for kind, buf in ((se.FILE_DATA, d), (se.FILE_INDEX, i)):
            pages = [p for k, _, _, p in got if k == kind]
            assert b"".join(pages)[:buf.size] == bytes(buf) and not any(b"".join(pages)[buf.size:])
    assert se.out_pages(np.zeros(0, np.uint8), np.zeros(0, np.uint8), 1) == []


@pytest.mark.gpu
def test_tree_write_through_warms_a_page_cache(engine, tmp_path):
    """entry_writer.rs:94-156 as a side effect of the tree's own writes: with a page sink installed, every SSTable a flush
    or a compaction writes also lands in the cache, page by page, under ((FileTypeKind, files_index), address) -- and what
    the cache holds is what the files hold (the reference's entry_writer_cache_equals_disk, lsm_tree.rs:1489-1556)."""
    d = str(tmp_path)
    rng = np.random.default_rng(12)
    tree = se.LSMTree(d, engine, sstable_bloom_min_size=1 << 40)
    cache = {}
    tree.set_page_cache(cache)
    for r in range(3):
        ents = [(b"key%05d" % int(k), bytes(rng.integers(0, 256, int(rng.integers(0, 700)), dtype=np.uint8)), BASE_TS + r)
                for k in rng.choice(4000, 900, replace=False)]
        tree.flush(sstable.build_run(ents))
    tree.compact([0, 2, 4], 5, False)
    for files_index in (0, 2, 4, 5):  # the flushed tables were cached too (their files are gone after the compaction)
        assert any(k[0] == (se.FILE_DATA, files_index) for k in cache)
    for kind, ext in ((se.FILE_DATA, "data"), (se.FILE_INDEX, "index")):
        raw = open(os.path.join(d, sstable.file_name(5, ext)), "rb").read()
        pages = sorted((addr, pg) for (fid, addr), pg in cache.items() if fid == (kind, 5))
        assert [a for a, _ in pages] == list(range(0, len(raw), 4096))
        joined = b"".join(pg for _, pg in pages)
        assert joined[:len(raw)] == raw and not any(joined[len(raw):])
    tree.close()


# ---- row N3: compactions streamed file -> pinned ring -> GPU -> pinned ring -> file (dbeel_compact_stream)

@pytest.fixture()
def tiny_partition_stream_engine(monkeypatch):
    """An engine that cuts even small jobs into many key-range partitions, so a few hundred KB of files exercise the
    whole streaming pipeline (reader / writer threads, ring reuse)."""
    monkeypatch.setenv("DBEEL_PIPELINE_MIN_KB", "1")
    monkeypatch.setenv("DBEEL_PARTITION_KB", "24")
    monkeypatch.setenv("DBEEL_IO_THREADS", "3")
    eng = capi.Engine(0)
    yield eng
    eng.close()


def _random_tree_runs(rng, n_runs, n_keys, id_space, doc=120):
    runs = []
    for r in range(n_runs):
        ids = sorted(rng.choice(id_space, n_keys, replace=False).tolist())
        ents = [(b"\xb0k%015d" % n, b"" if rng.random() < 0.02 else bytes(rng.integers(0, 256, int(rng.integers(1, doc)), dtype=np.uint8)),
                 BASE_TS + r * 10_000 + j) for j, n in enumerate(ids)]
        runs.append(sstable.build_run(ents))
    return runs


@pytest.mark.gpu
@pytest.mark.parametrize("ring", [2, 3])
def test_streamed_compaction_is_byte_identical(monkeypatch, tmp_path, ring):
    """dbeel_tree_compact streams its files through the engine's pinned rings (the default).  Same files as the oracle's
    compaction and as the whole-buffer path, bloom included; many partitions actually ran."""
    monkeypatch.setenv("DBEEL_PIPELINE_MIN_KB", "1")
    monkeypatch.setenv("DBEEL_PARTITION_KB", "24")
    monkeypatch.setenv("DBEEL_IO_THREADS", "3")
    monkeypatch.setenv("DBEEL_STREAM_RING", str(ring))
    eng = capi.Engine(0)
    try:
        rng = np.random.default_rng(70 + ring)
        runs = _random_tree_runs(rng, 5, 1500, 6000)
        od, oi, ob, on = oracle.compact(runs, False, bloom_min_size=50_000, seed=SEED)
        outs = {}
        for mode in ("1", "0"):  # streamed, then whole buffers
            d = str(tmp_path / f"stream{mode}")
            os.makedirs(d)
            for r, run in enumerate(runs):
                sstable.write_run_files(d, 2 * r, run)
            monkeypatch.setenv("DBEEL_TREE_STREAM", mode)
            tree = se.LSMTree(d, eng, sstable_bloom_min_size=50_000)
            tree.compact([0, 2, 4, 6, 8], 9, False, bloom_seed=SEED)
            assert tree.sstable_indices_and_sizes() == [(9, on)]
            assert sorted(os.listdir(d)) == [sstable.file_name(9, e) for e in ("bloom", "data", "index")]
            gd, gi = sstable.read_run_files(d, 9)
            assert_run_equal((gd, gi), (od, oi), f"streamed={mode}")
            assert np.array_equal(np.fromfile(os.path.join(d, sstable.file_name(9, "bloom")), dtype=np.uint8), ob)
            outs[mode] = eng.stats()
            tree.close()
        assert outs["1"]["partitions"] > 4 and outs["1"]["kernel_launches"] > 60
    finally:
        eng.close()


@pytest.mark.gpu
def test_streamed_compaction_small_job_and_damaged_run(tiny_partition_stream_engine, monkeypatch, tmp_path):
    """The two ways out of the pipeline: a job too small to partition (one piece, files whole through the callbacks), and a
    run that ends early -- its .data file cut short -- which is redone in one piece; the files are truncated to the final
    lengths.  Both equal the oracle."""
    eng = tiny_partition_stream_engine
    rng = np.random.default_rng(91)
    # (a) below the partition threshold: no bloom, a single piece
    monkeypatch.setenv("DBEEL_PIPELINE_MIN_KB", "100000")
    small = capi.Engine(0)
    try:
        d = str(tmp_path / "small")
        os.makedirs(d)
        runs = _random_tree_runs(rng, 3, 200, 500)
        for r, run in enumerate(runs):
            sstable.write_run_files(d, 2 * r, run)
        tree = se.LSMTree(d, small)
        tree.compact([0, 2, 4], 5, True)
        od, oi, _, on = oracle.compact(runs, True)
        assert tree.sstable_indices_and_sizes() == [(5, on)]
        assert sorted(os.listdir(d)) == [sstable.file_name(5, "data"), sstable.file_name(5, "index")]
        assert_run_equal(sstable.read_run_files(d, 5), (od, oi), "small streamed job")
        tree.close()
    finally:
        small.close()
    # (b) a run whose .data stops in the middle of an entry (lsm_tree.rs:1014,1063: the run simply ends there)
    d = str(tmp_path / "cut")
    os.makedirs(d)
    runs = _random_tree_runs(rng, 4, 1200, 4000)
    cut = (runs[2][0][: runs[2][0].size * 2 // 3 + 5].copy(), runs[2][1])
    runs[2] = cut
    for r, run in enumerate(runs):
        sstable.write_run_files(d, 2 * r, run)
    tree = se.LSMTree(d, eng, sstable_bloom_min_size=50_000)
    tree.compact([0, 2, 4, 6], 7, False, bloom_seed=SEED)
    od, oi, ob, on = oracle.compact(runs, False, bloom_min_size=50_000, seed=SEED)
    assert tree.sstable_indices_and_sizes() == [(7, on)]
    assert_run_equal(sstable.read_run_files(d, 7), (od, oi), "streamed job with a run that ends early")
    assert np.array_equal(np.fromfile(os.path.join(d, sstable.file_name(7, "bloom")), dtype=np.uint8), ob)
    tree.close()
    # (c) an entry header in the middle of a run disagrees with its index record: the planner does not see it, the partition
    # that holds it reports a run that ended early, the job is redone in one piece over what the pipeline had already written
    d = str(tmp_path / "mid")
    os.makedirs(d)
    runs = _random_tree_runs(rng, 4, 1200, 4000)
    data1 = runs[1][0].copy()
    off = int.from_bytes(bytes(runs[1][1][16 * 700:16 * 700 + 8]), "little")
    data1[off] ^= 0x40  # the key-length prefix of entry 700
    runs[1] = (data1, runs[1][1])
    for r, run in enumerate(runs):
        sstable.write_run_files(d, 2 * r, run)
    tree = se.LSMTree(d, eng, sstable_bloom_min_size=50_000)
    tree.compact([0, 2, 4, 6], 7, False, bloom_seed=SEED)
    od, oi, ob, on = oracle.compact(runs, False, bloom_min_size=50_000, seed=SEED)
    assert tree.sstable_indices_and_sizes() == [(7, on)]
    assert eng.stats()["runs_truncated"] == 1
    assert_run_equal(sstable.read_run_files(d, 7), (od, oi), "streamed job with a damaged entry")
    assert np.array_equal(np.fromfile(os.path.join(d, sstable.file_name(7, "bloom")), dtype=np.uint8), ob)
    tree.close()


@pytest.mark.gpu
def test_streamed_compaction_reports_a_missing_input(tiny_partition_stream_engine, tmp_path):
    d = str(tmp_path)
    run = sstable.build_run([(b"k%04d" % n, b"v", BASE_TS + n) for n in range(50)])
    sstable.write_run_files(d, 0, run)
    tree = se.LSMTree(d, tiny_partition_stream_engine)
    with pytest.raises(capi.DbeelError):
        tree.compact([0, 2], 3, False)
    assert sorted(os.listdir(d)) == [sstable.file_name(0, "data"), sstable.file_name(0, "index")]  # nothing left behind
    tree.close()
