"""Shared test helpers: adversarial run generators and a pure-Python model of the merge
semantics that is independent of both the C oracle and the GPU path."""
from __future__ import annotations

import numpy as np

from dbeel_b200 import sstable

BASE_TS = 1_700_000_000_000_000_000


def nasty_keys(rng: np.random.Generator, n: int, max_len: int = 40) -> list:
    """Distinct keys engineered to stress byte-wise ordering: shared prefixes, keys that are
    prefixes of other keys, 0x00 / 0xFF bytes, the empty key, keys longer than any
    fixed-width comparison window."""
    pool = set()
    stems = [b"", b"\x00", b"\xff", b"\xb0k0000", b"ab", b"ab\x00", b"ab\x00\x00", b"\xff\xff\xff",
             b"common-prefix-that-is-quite-long/", b"common-prefix-that-is-quite-long/\x00"]
    while len(pool) < n:
        stem = stems[int(rng.integers(len(stems)))]
        extra = int(rng.integers(0, max(1, max_len - len(stem)) + 1))
        style = int(rng.integers(4))
        if style == 0:
            tail = bytes(rng.integers(0, 256, extra, dtype=np.uint8))
        elif style == 1:
            tail = bytes(rng.integers(0, 2, extra, dtype=np.uint8) * 255)
        elif style == 2:
            tail = b"\x00" * extra
        else:
            tail = bytes(rng.integers(48, 58, extra, dtype=np.uint8))
        pool.add(stem + tail)
    keys = sorted(pool)
    rng.shuffle(keys)
    return keys[:n]


def random_runs(rng: np.random.Generator, n_runs: int, keys_per_run, key_pool: list, max_doc: int = 60,
                tombstone_frac: float = 0.15, equal_ts_frac: float = 0.3):
    """n_runs sorted runs over a shared key pool (so keys collide across runs).  Some
    timestamps collide exactly to exercise the run-position tie-break."""
    runs = []
    fixed_ts = BASE_TS + int(rng.integers(0, 5))
    for r in range(n_runs):
        n = keys_per_run[r] if isinstance(keys_per_run, (list, tuple)) else keys_per_run
        n = min(n, len(key_pool))
        idx = rng.choice(len(key_pool), size=n, replace=False)
        keys = sorted(key_pool[i] for i in idx)
        ents = []
        for k in keys:
            if rng.random() < tombstone_frac:
                v = b""
            else:
                v = bytes(rng.integers(0, 256, int(rng.integers(1, max_doc + 1)), dtype=np.uint8))
            if rng.random() < equal_ts_frac:
                ts = fixed_ts
            else:
                ts = BASE_TS + int(rng.integers(-50, 50)) + r * 3
            ents.append((k, v, ts))
        runs.append(sstable.build_run(ents))
    return runs


def model_compact(runs, keep_tombstones: bool):
    """Appendix A.3 in ten lines: per key the max (timestamp, run position) entry survives;
    a surviving tombstone is dropped unless keep_tombstones; output ascending by key."""
    best = {}
    for r, (d, i) in enumerate(runs):
        for k, v, ts in sstable.parse_run(d, i):
            if k not in best or (ts, r) > (best[k][1], best[k][2]):
                best[k] = (v, ts, r)
    out = [(k, v, ts) for k, (v, ts, _) in sorted(best.items()) if keep_tombstones or v != b""]
    return sstable.build_run(out), len(out)


def model_flush(batch, capacity: int):
    """Appendix A.5: cut the arrival stream every time `capacity` distinct keys are live
    (set_ex flushes right after the insert that fills the tree), last arrival wins."""
    ents = sstable.parse_run(*batch)
    outs = []
    cur = {}
    for k, v, ts in ents:
        cur[k] = (v, ts)
        if len(cur) == capacity:
            outs.append(sstable.build_run([(kk, vv, tt) for kk, (vv, tt) in sorted(cur.items())]))
            cur = {}
    if cur:
        outs.append(sstable.build_run([(kk, vv, tt) for kk, (vv, tt) in sorted(cur.items())]))
    return outs


def assert_run_equal(got, exp, what=""):
    gd, gi = np.asarray(got[0]), np.asarray(got[1])
    ed, ei = np.asarray(exp[0]), np.asarray(exp[1])
    assert gi.size == ei.size, f"{what}: index length {gi.size} != {ei.size}"
    if not np.array_equal(gi, ei):
        bad = int(np.flatnonzero(gi != ei)[0])
        raise AssertionError(f"{what}: .index differs at byte {bad} (record {bad // 16})")
    assert gd.size == ed.size, f"{what}: data length {gd.size} != {ed.size}"
    if not np.array_equal(gd, ed):
        bad = int(np.flatnonzero(gd != ed)[0])
        raise AssertionError(f"{what}: .data differs at byte {bad}")
