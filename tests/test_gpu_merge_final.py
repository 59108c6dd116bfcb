"""k_merge_final (DBEEL_FUSED_FINAL=1): the last merge level, resolve, the chained offsets scan and the .index writes in one
persistent kernel.  Off by default (measured slower than the five kernels it replaces), but a selectable
path of the product library: same bytes as the oracle on everything the default path is tested on, plus the cases only this
kernel has -- groups of equal keys that straddle its tile borders, tiles whose neighbours decide a group's head."""
import numpy as np
import pytest

from dbeel_b200 import capi, sstable
from dbeel_b200 import workloads as W

from helpers import BASE_TS, assert_run_equal, model_compact, nasty_keys, random_runs
from test_gpu_parity import check_against_oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def fused_engine():
    import os
    old = os.environ.get("DBEEL_FUSED_FINAL")
    os.environ["DBEEL_FUSED_FINAL"] = "1"
    eng = capi.Engine(0)
    if old is None:
        del os.environ["DBEEL_FUSED_FINAL"]
    else:
        os.environ["DBEEL_FUSED_FINAL"] = old
    yield eng
    eng.close()


@pytest.mark.parametrize("seed", range(6))
@pytest.mark.parametrize("keep", [False, True])
def test_fused_adversarial_keys(fused_engine, seed, keep):
    rng = np.random.default_rng(1000 + seed)
    pool = nasty_keys(rng, 3000, max_len=40)
    k = int(rng.integers(2, 12))
    runs = random_runs(rng, k, [int(rng.integers(0, 2500)) for _ in range(k)], pool, max_doc=90)
    gd, gi, gb, n = check_against_oracle(fused_engine, runs, keep, what=f"fused adversarial {seed} keep={keep}")
    exp, en = model_compact(runs, keep)
    assert n == en
    assert_run_equal((gd, gi), exp, "fused adversarial vs model")
    st = fused_engine.stats()
    # header + prefix/extract (x2) + plan, two launches per level, gather, publish (+ the filter's frame): the separate
    # resolve / scan / scan / emit launches of the default path are gone
    assert st["kernel_launches"] == 8 + 2 * st["merge_passes"] + (0 if gb is None else 1)


@pytest.mark.parametrize("cfg,keys", [(W.CFG2, 100_000), (W.CFG3, 20_000)])
def test_fused_scaled_benchmark_shapes(fused_engine, cfg, keys):
    c = W.scaled(cfg, keys)
    check_against_oracle(fused_engine, W.make_merge_runs(c), c.keep_tombstones, what=c.name)
    check_against_oracle(fused_engine, W.make_merge_runs(c, equal_ts=True), c.keep_tombstones, what=c.name + " equal ts")


def test_fused_edge_shapes(fused_engine):
    one = sstable.build_run([(b"k%05d" % n, b"v" * 9, BASE_TS + n) for n in range(5000)])
    empty = sstable.build_run([])
    other = sstable.build_run([(b"k%05d" % n, b"", BASE_TS + 10_000 + n) for n in range(0, 5000, 3)])
    for runs, keep in [([one, empty], False), ([empty, one], True), ([empty, empty], False), ([one, other], False),
                       ([one, other], True), ([other, one, empty, one], False), ([one], False)]:
        check_against_oracle(fused_engine, runs, keep, what=f"fused edge {len(runs)} runs keep={keep}")
    # truncated run: the reader stops at the first undecodable entry, the re-extraction pass runs before the merge levels
    d, i = one
    cut = (d[: d.size - 7], i)
    check_against_oracle(fused_engine, [cut, other], False, what="fused truncated")


@pytest.mark.parametrize("seed", range(3))
def test_groups_straddling_merge_tiles(fused_engine, seed):
    """80 runs share 70 hot keys, so the last merge level meets groups of up to 80 equal keys (40 from either side).  Tile
    borders that fall inside such a group are moved past at most 31 + 31 of its records; the rest is finished by the head's
    walk through global memory.  Equal timestamps exercise the run-position tie-break across the border, tombstones the
    drop rule on a winner found on the far side."""
    rng = np.random.default_rng(900 + seed)
    hot = [b"\xb0k%015d" % (n * 997) for n in range(70)]
    runs = []
    for r in range(80):
        fill = [b"\xb0k%015d" % int(x) for x in rng.choice(70_000, size=int(rng.integers(300, 700)), replace=False)]
        keys = sorted(set(fill) | {k for k in hot if rng.random() < 0.97})
        ents = []
        for k in keys:
            v = b"" if rng.random() < 0.2 else bytes(rng.integers(0, 256, int(rng.integers(1, 40)), dtype=np.uint8))
            ts = BASE_TS + (5 if rng.random() < 0.5 else int(rng.integers(0, 9)))
            ents.append((k, v, ts))
        runs.append(sstable.build_run(ents))
    for keep in (False, True):
        gd, gi, _, n = check_against_oracle(fused_engine, runs, keep, what=f"straddling groups keep={keep}")
        exp, en = model_compact(runs, keep)
        assert n == en
        assert_run_equal((gd, gi), exp, "straddling groups vs model")
    assert fused_engine.stats()["merge_passes"] == 7


def test_fused_reference_reader_and_host_pipeline(fused_engine):
    """The fused tail under the other job flavours that reach it: reference-reader mode and the key-range partitions of the
    host entry point (every partition is a single compaction with an output offset base and a shared filter)."""
    rng = np.random.default_rng(77)
    pool = nasty_keys(rng, 4000, max_len=30)
    runs = random_runs(rng, 5, [2500, 1800, 2900, 700, 2200], pool, max_doc=100)
    gd, gi, gb, gn = fused_engine.compact(runs, keep_tombstones=False, bloom_min_size=10_000, seed=bytes(range(32)),
                                          flags=capi.FLAG_REFERENCE_READER)
    import oracle
    od, oi, ob, on = oracle.compact(runs, keep_tombstones=False, bloom_min_size=10_000, seed=bytes(range(32)))
    assert gn == on and np.array_equal(gd, od) and np.array_equal(gi, oi) and np.array_equal(gb, ob)
