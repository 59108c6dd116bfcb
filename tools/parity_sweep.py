#!/usr/bin/env python
"""Randomised byte-parity sweep against the oracle, through the C ABI, on one GPU: many small jobs of random shape
(run counts, sizes, key pools with shared prefixes / empty keys / 0x00 / 0xFF, tombstones, colliding timestamps) over
every entry point: compact (single shot and pipelined in tiny key-range partitions), compact_many, flush, flush_many,
wal_flush, get_many.  A mismatch prints the seed and exits 1.  Usage: tools/parity_sweep.py [iterations (default 300)] [first seed]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402

import oracle  # noqa: E402  (the checker)
from dbeel_b200 import capi, sstable  # noqa: E402
from helpers import BASE_TS, nasty_keys, random_runs  # noqa: E402

SEED32 = bytes(range(32))


def same(a, b):
    return a.size == b.size and np.array_equal(a, b)


def arrivals(rng, pool, n, max_doc):
    out = []
    for j in range(n):
        k = pool[int(rng.integers(len(pool)))]
        v = b"" if rng.random() < 0.1 else bytes(rng.integers(0, 256, int(rng.integers(0, max_doc + 1)), dtype=np.uint8))
        out.append((k, v, BASE_TS + int(rng.integers(-3, 3)) + (j if rng.random() < 0.5 else 0)))
    return out


def main():
    iters = int(sys.argv[1]) if len(sys.argv) > 1 else 300
    seed0 = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
    eng = capi.Engine(0)
    os.environ["DBEEL_PARTITION_KB"] = "24"
    os.environ["DBEEL_PIPELINE_MIN_KB"] = "1"
    eng_p = capi.Engine(0)  # pipelined host path with tiny partitions
    for k in ("DBEEL_PARTITION_KB", "DBEEL_PIPELINE_MIN_KB"):
        del os.environ[k]
    counts = {}
    t0 = time.perf_counter()
    for it in range(iters):
        seed = seed0 + it
        rng = np.random.default_rng(seed)
        kind = ["compact", "pipelined", "flush", "flush_many", "wal", "get", "compact_many"][it % 7]
        pool = nasty_keys(rng, int(rng.integers(5, 1500)), max_len=int(rng.integers(4, 90)))
        what = f"seed {seed} ({kind})"
        try:
            if kind in ("compact", "pipelined"):
                n_runs = int(rng.integers(1, 21))
                sizes = [int(rng.integers(0, min(len(pool), 3000) + 1)) for _ in range(n_runs)]
                runs = random_runs(rng, n_runs, sizes, pool, max_doc=int(rng.choice([8, 60, 300, 2000])),
                                   tombstone_frac=float(rng.choice([0.0, 0.15, 0.6])), equal_ts_frac=float(rng.choice([0.0, 0.3, 1.0])))
                keep = bool(rng.integers(2))
                bms = int(rng.choice([1, 1000, 1 << 20]))
                e = eng_p if kind == "pipelined" else eng
                gd, gi, gb, gn = e.compact(runs, keep_tombstones=keep, bloom_min_size=bms, seed=SEED32)
                od, oi, ob, on = oracle.compact(runs, keep_tombstones=keep, bloom_min_size=bms, seed=SEED32)
                ok = gn == on and same(gd, od) and same(gi, oi) and (gb is None) == (ob is None) and (gb is None or same(gb, ob))
                if kind == "pipelined":
                    counts["partitions>1"] = counts.get("partitions>1", 0) + (e.stats()["partitions"] > 1)
            elif kind == "compact_many":
                jobs, seeds = [], []
                for j in range(int(rng.integers(1, 9))):
                    n_runs = int(rng.integers(0, 10))
                    sizes = [int(rng.integers(0, min(len(pool), 1200) + 1)) for _ in range(n_runs)]
                    jobs.append((random_runs(rng, n_runs, sizes, pool, max_doc=int(rng.choice([8, 60, 600])),
                                             tombstone_frac=float(rng.choice([0.0, 0.15, 0.6])),
                                             equal_ts_frac=float(rng.choice([0.0, 0.3, 1.0]))) if n_runs else [], bool(rng.integers(2))))
                    seeds.append(bytes(rng.integers(0, 256, 32, dtype=np.uint8)))
                bms = int(rng.choice([1, 5000, 1 << 20]))
                got = eng.compact_many(jobs, bloom_min_size=bms, seeds=seeds)
                ok = len(got) == len(jobs)
                for (runs, keep), sd, (gd, gi, gb, gn) in zip(jobs, seeds, got):
                    od, oi, ob, on = oracle.compact(runs, keep_tombstones=keep, bloom_min_size=bms, seed=sd)
                    ok = ok and gn == on and same(gd, od) and same(gi, oi) and (gb is None) == (ob is None) and (gb is None or same(gb, ob))
            elif kind == "flush":
                ents = arrivals(rng, pool, int(rng.integers(1, 6000)), int(rng.choice([4, 80, 700])))
                gd, gi, gn = eng.flush(sstable.build_run(ents))
                od, oi, on = oracle.memtable_flushes(sstable.build_run(ents), capacity=1 << 20)[0]
                ok = gn == on and same(gd, od) and same(gi, oi)
            elif kind == "flush_many":
                batches = [sstable.build_run(arrivals(rng, pool, int(rng.integers(0, 4000)), int(rng.choice([4, 80, 400]))))
                           for _ in range(int(rng.integers(1, 9)))]
                got = eng.flush_many(batches)
                ok = True
                for b, (gd, gi, gn) in zip(batches, got):
                    if b[1].size == 0:
                        ok = ok and gn == 0 and gd.size == 0
                        continue
                    od, oi, on = oracle.memtable_flushes(b, capacity=1 << 20)[0]
                    ok = ok and gn == on and same(gd, od) and same(gi, oi)
            elif kind == "wal":
                ents = arrivals(rng, pool, int(rng.integers(1, 1500)), int(rng.choice([4, 500, 5000, 9000])))
                if rng.random() < 0.3:
                    j = int(rng.integers(len(ents)))
                    ents[j] = (ents[j][0], ents[j][1], 1 << 100)  # undecodable timestamp: skipped
                wal = sstable.build_wal(ents, pad_byte=int(rng.integers(0, 256)))
                if rng.random() < 0.3:
                    wal = wal[:int(rng.integers(1, wal.size + 1))]  # torn tail
                if rng.random() < 0.2:
                    wal = np.concatenate([wal, np.zeros(4096 * int(rng.integers(1, 3)), np.uint8)])
                gd, gi, gn = eng.wal_flush(wal, capacity=1 << 20)
                od, oi, on, _ = oracle.wal_flush(wal, capacity=1 << 20)
                ok = gn == on and same(gd, od) and same(gi, oi)
            else:
                tables = []
                for t in range(int(rng.integers(1, 5))):
                    runs = random_runs(rng, int(rng.integers(1, 4)), int(rng.integers(1, min(len(pool), 2000) + 1)), pool, max_doc=40)
                    d, i, b, _ = oracle.compact(runs, True, bloom_min_size=int(rng.choice([1, 1 << 40])), seed=SEED32)
                    tables.append((d, i, b))
                keys = [pool[int(j)] for j in rng.integers(0, len(pool), 800)] + [bytes(rng.integers(0, 256, int(rng.integers(0, 30)), dtype=np.uint8)) for _ in range(200)]
                blob, off = capi.pack_keys(keys)
                et, er, ej = oracle.get_many(tables, blob, off)
                res = eng.get_many(tables, keys, capi.LOOKUP_REFERENCE)
                ok = np.array_equal(res["table"], et) and np.array_equal(res["bloom_rejects"], ej) and \
                    np.array_equal(np.where(res["table"] >= 0, res["record"], 0), er)
                ex = eng.get_many(tables, keys, capi.LOOKUP_EXACT)
                present = [{k for k, _, _ in sstable.parse_run(d, i)} for d, i, _ in tables]
                for k, row in zip(keys, ex):  # exact mode: the newest table that holds the key answers
                    want = max((t for t in range(len(tables)) if k in present[t]), default=-1)
                    ok = ok and int(row["table"]) == want
        except Exception as ex_:  # noqa: BLE001
            print(f"{what}: EXCEPTION {ex_!r}", flush=True)
            sys.exit(1)
        if not ok:
            print(f"{what}: MISMATCH", flush=True)
            sys.exit(1)
        counts[kind] = counts.get(kind, 0) + 1
    print(f"parity sweep: {iters} jobs byte-identical to the oracle in {time.perf_counter() - t0:.0f} s  {counts}", flush=True)


if __name__ == "__main__":
    main()
