"""Synthetic SSTable runs of the shapes BASELINE.json names.

Keys are what db_server.rs:98-105 stores: the msgpack encoding of the user's key -- here a
``fixstr`` of 16 ASCII chars ``"k%015d" % id`` (17 bytes).  Documents are msgpack ``bin``
blobs of V bytes including their header.  Timestamps are unix nanos, strictly increasing
with write order so that a newer run wins a key; ``equal_ts=True`` makes every timestamp
identical, which exercises the run-position tie-break of lsm_tree.rs:58-65.

Everything is seeded; the same (config, seed) always yields the same bytes.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List

import numpy as np

from .sstable import Run, build_run_dense

BASE_TS = 1_700_000_000_000_000_000  # 2023-11-14T22:13:20Z in nanos


@dataclass(frozen=True)
class MergeConfig:
    name: str
    n_runs: int
    keys_per_run: int
    doc_bytes: int  # V: msgpack-encoded document size
    id_space: int  # key ids are drawn from [0, id_space)
    tombstone_frac: float
    keep_tombstones: bool
    seed: int
    contiguous_overlap: float = -1.0  # >=0: run r covers a sliding id window (cfg1/cfg3 style)


# BASELINE.json configs[0..3] (cfg4 = 8 x cfg2 with seeds 40..47)
CFG1 = MergeConfig("cfg1-2way-1k-64B", 2, 1_000, 64, 1_900, 0.01, False, 1, contiguous_overlap=0.10)
CFG2 = MergeConfig("cfg2-8way-1M-256B", 8, 1_000_000, 256, 16_000_000, 0.01, False, 2)
CFG3 = MergeConfig("cfg3-16way-4M-1KB", 16, 250_000, 1024, 2_125_000, 0.01, False, 3, contiguous_overlap=0.5)


def cfg4_shard(shard: int) -> MergeConfig:
    return MergeConfig(f"cfg4-shard{shard}-8way-1M-256B", 8, 1_000_000, 256, 16_000_000, 0.01, False, 40 + shard)


def scaled(cfg: MergeConfig, keys_per_run: int) -> MergeConfig:
    """Same shape at a smaller key count (parity tests, CPU-baseline samples)."""
    f = keys_per_run / cfg.keys_per_run
    return MergeConfig(f"{cfg.name}@{keys_per_run}", cfg.n_runs, keys_per_run, cfg.doc_bytes,
                       max(keys_per_run, int(cfg.id_space * f)), cfg.tombstone_frac, cfg.keep_tombstones,
                       cfg.seed, cfg.contiguous_overlap)


def format_keys(ids: np.ndarray) -> np.ndarray:
    """ids -> [n, 17] uint8: 0xb0 (fixstr len 16), 'k', 15 zero-padded decimal digits."""
    n = ids.shape[0]
    out = np.empty((n, 17), dtype=np.uint8)
    out[:, 0] = 0xB0
    out[:, 1] = ord("k")
    rem = ids.astype(np.uint64).copy()
    for col in range(16, 1, -1):
        out[:, col] = (rem % 10).astype(np.uint8) + ord("0")
        rem //= 10
    return out


def _dense_rows(seed_seq, n: int, klen: int, doc_bytes: int) -> np.ndarray:
    """[n, 32+klen+V] uint8 row buffer whose document columns hold a msgpack bin8/bin16 blob
    with pseudo-random payload (SFC64 raw words: payload bytes only need to be
    incompressible-looking and reproducible)."""
    full = 32 + klen + doc_bytes
    words = (n * full + 7) // 8
    raw = np.random.SFC64(seed_seq).random_raw(words).view(np.uint8)[:n * full]
    dense = raw.reshape(n, full)
    c = 16 + klen
    if doc_bytes - 2 <= 255:
        dense[:, c] = 0xC4
        dense[:, c + 1] = doc_bytes - 2
    else:
        dense[:, c] = 0xC5
        dense[:, c + 1] = (doc_bytes - 3) >> 8
        dense[:, c + 2] = (doc_bytes - 3) & 0xFF
    return dense


def _run_ids(cfg: MergeConfig, rng: np.random.Generator, r: int) -> np.ndarray:
    n = cfg.keys_per_run
    if cfg.contiguous_overlap >= 0:
        # run r covers ids [r*stride, r*stride + n): neighbouring runs share `overlap` of their keys
        stride = int(round(n * (1.0 - cfg.contiguous_overlap)))
        return np.arange(r * stride, r * stride + n, dtype=np.int64)
    ids = rng.choice(cfg.id_space, size=n, replace=False)
    ids.sort()
    return ids.astype(np.int64)


def make_merge_runs(cfg: MergeConfig, equal_ts: bool = False) -> List[Run]:
    runs = []
    for r in range(cfg.n_runs):
        rng = np.random.default_rng([cfg.seed, r])
        ids = _run_ids(cfg, rng, r)
        n = ids.shape[0]
        keys = format_keys(ids)
        dense = _dense_rows([cfg.seed, r, 7], n, keys.shape[1], cfg.doc_bytes)
        tomb = rng.random(n) < cfg.tombstone_frac
        if equal_ts:
            ts = np.full(n, BASE_TS, dtype=np.int64)
        else:
            # write order inside a memtable is unrelated to key order
            ts = BASE_TS + (np.int64(r) * n + rng.permutation(n).astype(np.int64)) * 1000
        runs.append(build_run_dense(dense, keys, ts, tomb))
    return runs


def make_arrival_batch(n_writes: int, n_ids: int, doc_bytes: int, seed: int, zipf_s: float = 0.99,
                       tombstone_frac: float = 0.01) -> Run:
    """cfg5-style write stream: key ids ~ Zipf(s) over n_ids, arrival order, strictly
    increasing timestamps.  Output layout = run layout (unsorted, keys repeat)."""
    rng = np.random.default_rng(seed)
    ranks = np.arange(1, n_ids + 1, dtype=np.float64)
    p = ranks ** (-zipf_s)
    p /= p.sum()
    cdf = np.cumsum(p)
    hot = np.searchsorted(cdf, rng.random(n_writes), side="right").clip(0, n_ids - 1)
    # decorrelate hotness from key order with a fixed permutation of the id space
    perm = np.random.default_rng(seed ^ 0x5EED).permutation(n_ids)
    ids = perm[hot].astype(np.int64)
    keys = format_keys(ids)
    dense = _dense_rows([seed, 7], n_writes, keys.shape[1], doc_bytes)
    tomb = rng.random(n_writes) < tombstone_frac
    ts = BASE_TS + np.arange(n_writes, dtype=np.int64) * 1000
    return build_run_dense(dense, keys, ts, tomb)
