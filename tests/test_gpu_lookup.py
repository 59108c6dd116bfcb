"""Row N2 (read side of the produced files): dbeel_get_many vs the oracle's restatement of get_entry's SSTable loop
(lsm_tree.rs:686-719: Bloom::check, then binary_search :605-670), through the C ABI, on files the GPU engine wrote."""
import numpy as np
import pytest

import oracle
from dbeel_b200 import capi, sstable
from dbeel_b200 import workloads as W

from helpers import BASE_TS, nasty_keys, random_runs

pytestmark = pytest.mark.gpu

SEED = bytes(range(32))


def check(engine, tables, keys, mode=capi.LOOKUP_REFERENCE):
    """GPU rows == the oracle's restatement of get_entry's loop (newest table first, bloom.check, binary_search)."""
    res = engine.get_many(tables, keys, mode)
    assert len(res) == len(keys)
    blob, off = capi.pack_keys(keys)
    et, er, ej = oracle.get_many(tables, blob, off)
    for name, got, exp in (("table", res["table"], et), ("bloom_rejects", res["bloom_rejects"], ej),
                           ("record", np.where(res["table"] >= 0, res["record"], 0), er)):
        if not np.array_equal(got, exp):
            q = int(np.flatnonzero(got != exp)[0])
            raise AssertionError(f"key {keys[q]!r}: {name} GPU {got[q]} != oracle {exp[q]}")
    return res


def keys_of(table):
    return [k for k, _, _ in sstable.parse_run(table[0], table[1])]


def test_reference_loop_on_a_compacted_table_with_bloom(engine):
    c = W.scaled(W.CFG2, 20_000)
    runs = W.make_merge_runs(c)
    gd, gi, gb, n = engine.compact(runs, keep_tombstones=False, seed=SEED)
    assert gb is not None
    table = (gd, gi, gb)
    present = keys_of(table)
    rng = np.random.default_rng(1)
    sample = [present[j] for j in rng.choice(len(present), 3000, replace=False)] + present[:40] + present[-40:]
    absent = [b"\xb0k%015d" % int(x) for x in rng.integers(0, 1 << 40, 1500)] + [b"", b"\xb0", b"\xb0k", present[7] + b"\x00",
                                                                                   present[9][:-1], b"\xff" * 40]
    res = check(engine, [table], sample + absent)
    found = int((res["table"][:len(sample)] >= 0).sum())
    assert found >= len(sample) - 64  # the reference's loop gives up on a few present keys (tests/test_oracle_goldens.py)
    assert engine.stats()["kernel_launches"] == 1


def test_exact_mode_finds_every_present_key(engine):
    rng = np.random.default_rng(2)
    pool = nasty_keys(rng, 3000, max_len=70)
    runs = random_runs(rng, 3, 1800, pool, max_doc=80, tombstone_frac=0.0)
    gd, gi, gb, n = engine.compact(runs, keep_tombstones=True, bloom_min_size=1000, seed=SEED)
    present = keys_of((gd, gi))
    assert len(present) == n and gb is not None
    res = engine.get_many([(gd, gi, gb)], present, capi.LOOKUP_EXACT)
    assert np.array_equal(res["table"], np.zeros(n, np.int32))
    assert np.array_equal(res["record"], np.arange(n, dtype=np.uint64))
    others = [k for k in pool if k not in set(present)] + [p + b"\x01" for p in present[:200]]
    miss = engine.get_many([(gd, gi, gb)], others, capi.LOOKUP_EXACT)
    assert (miss["table"] == -1).all()
    # the reference-faithful mode agrees with the oracle on the very same keys, hits and misses alike
    check(engine, [(gd, gi, gb)], present[::3] + others[:300])


def test_newest_table_answers_first_and_filters_are_counted(engine):
    rng = np.random.default_rng(3)
    ents = [(b"\xb0k%015d" % n, bytes(rng.integers(0, 256, 40, dtype=np.uint8)), BASE_TS + n) for n in range(0, 9000)]
    tables = []
    for t, (lo, hi, step) in enumerate([(0, 6000, 1), (3000, 9000, 2), (100, 8000, 3), (5000, 5400, 1)]):
        run = sstable.build_run(ents[lo:hi:step])
        d, i, b, _ = engine.compact([run], keep_tombstones=True, bloom_min_size=1000 if t != 2 else 1 << 40, seed=SEED)
        assert (b is None) == (t == 2)  # the third table has no filter (lsm_tree.rs:94-101: .bloom file absent)
        tables.append((d, i, b))
    keys = [ents[int(j)][0] for j in rng.integers(0, 9000, 2500)] + [b"\xb0k%015d" % n for n in range(9000, 9300)]
    res = check(engine, tables, keys)
    assert set(np.unique(res["table"])) >= {-1, 0, 1, 2, 3}
    assert int(res["bloom_rejects"].max()) >= 2
    check(engine, tables[:1], keys[:200])
    assert (engine.get_many([], keys[:10])["table"] == -1).all()
    assert len(engine.get_many(tables, [])) == 0


def test_tiny_tables_and_the_reference_early_exit(engine):
    """The reference's loop leaves right after probing index 0 (`if half == 0 ... break`, lsm_tree.rs:660), whatever the
    comparison said: on tiny tables that is a large share of the keys.  REFERENCE mode restates it, EXACT repairs it."""
    for n in (1, 2, 3, 4, 5, 8, 33):
        ents = [(bytes([10 + 2 * j]), b"v", BASE_TS + j) for j in range(n)]
        d, i = sstable.build_run(ents)
        keys = [bytes([x]) for x in range(8, 12 + 2 * n)]
        check(engine, [(d, i, None)], keys)
        ex = engine.get_many([(d, i, None)], [k for k, _, _ in ents], capi.LOOKUP_EXACT)
        assert np.array_equal(ex["record"], np.arange(n, dtype=np.uint64)) and (ex["table"] == 0).all()
    empty = (np.zeros(0, np.uint8), np.zeros(0, np.uint8), None)
    assert (engine.get_many([empty], [b"a", b""])["table"] == -1).all()


def test_device_resident_lookup_on_compaction_output(engine):
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
    dl, il, bl, n = engine.compact_device([(d.data_ptr(), d.numel(), i.data_ptr(), i.numel()) for d, i in t_runs],
                                          (od.data_ptr(), dc, oi.data_ptr(), ic, ob.data_ptr(), bc), opts)
    host_table = (od[:dl].cpu().numpy(), oi[:il].cpu().numpy(), ob[:bl].cpu().numpy())
    present = keys_of(host_table)
    rng = np.random.default_rng(4)
    keys = [present[j] for j in rng.choice(n, 20_000, replace=False)] + [b"\xb0k%015d" % int(x) for x in rng.integers(0, 1 << 40, 5000)]
    blob, off = capi.pack_keys(keys)
    d_keys, d_off = torch.from_numpy(blob.copy()).to(dev), torch.from_numpy(off.view(np.int64)).to(dev)
    d_res = torch.zeros(len(keys) * 2, dtype=torch.int64, device=dev)
    torch.cuda.synchronize()
    for mode in (capi.LOOKUP_REFERENCE, capi.LOOKUP_EXACT):
        engine.get_many_device([(od.data_ptr(), dl, oi.data_ptr(), il, ob.data_ptr(), bl)], d_keys.data_ptr(), d_off.data_ptr(),
                               len(keys), d_res.data_ptr(), mode)
        got = d_res.cpu().numpy().view(capi.LOOKUP_DTYPE)
        want = engine.get_many([host_table], keys, mode)
        assert np.array_equal(got, want)
        if mode == capi.LOOKUP_EXACT:
            assert (got["table"][:20_000] == 0).all()
            pos = {k: j for j, k in enumerate(present)}
            assert all(int(r) == pos[k] for k, r in zip(keys[:20_000:97], got["record"][:20_000:97]))
    assert engine.stats()["ms_total"] > 0


def test_malformed_inputs_are_reported(engine):
    ents = [(b"key%04d" % j, b"value", BASE_TS + j) for j in range(500)]
    run = sstable.build_run(ents)
    d, i, b, _ = engine.compact([run], keep_tombstones=True, bloom_min_size=1000, seed=SEED)
    with pytest.raises(capi.DbeelError) as ei:
        engine.get_many([(d, i, b[:-4])], [b"key0001"])
    assert ei.value.code == capi.ERR_BAD_BLOOM
    bad = b.copy()
    bad[:8] = np.frombuffer((12345).to_bytes(8, "little"), np.uint8)
    with pytest.raises(capi.DbeelError) as ei:
        engine.get_many([(d, i, bad)], [b"key0001"])
    assert ei.value.code == capi.ERR_BAD_BLOOM
    with pytest.raises(capi.DbeelError) as ei:
        engine.get_many([(d, i[:-3], None)], [b"key0001"])
    assert ei.value.code == capi.ERR_INVALID_ARG
    # an index record that points past the end of .data: the reference's read_at fails, the row carries the flag
    broken = i.copy()
    broken[16 * 250:16 * 250 + 8] = np.frombuffer((1 << 40).to_bytes(8, "little"), np.uint8)
    res = engine.get_many([(d, broken, None)], [b"key0100", b"key0400"])
    assert (res["table"] == -1).all() and (res["bloom_rejects"] & capi.LOOKUP_CORRUPT).all()
