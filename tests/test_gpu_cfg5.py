"""BASELINE.json configs[4] at test scale: Zipf writes -> murmur3 ring -> per-shard memtables -> flush waves -> size-tiered
compactions, all through the product's device entry points (dbeel_b200/cfg5.py); every table that is left is compared with
the oracle's replay of the same shard (red-black-tree memtables + the recorded plan)."""
import numpy as np
import pytest

import bench_cfg5
from dbeel_b200 import capi, cfg5

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n_writes,capacity,factor,wave,world", [(120_000, 512, 4, 4, 1), (90_000, 300, 2, 3, 1), (200_000, 1024, 8, 8, 2),
                                                                 (40_000, 8192, 8, 8, 1)])
def test_cfg5_pipeline_matches_the_oracle(engine, n_writes, capacity, factor, wave, world):
    import torch
    dev = torch.device("cuda:0")
    ids, tomb = cfg5.stream_ids(n_writes, max(2000, n_writes // 5))
    data, index, total = cfg5.build_stream_device(torch, dev, ids, tomb, chunk=50_000)
    ring, _ = capi.shard_ring(cfg5.N_SHARDS)
    data_host = data.cpu().numpy()
    seen = 0
    for rank in range(world):  # the ranks of an N-GPU run, one after the other on this GPU
        res = cfg5.pipeline(engine, torch, dev, data, index, ring, cfg5.own_positions(rank, world, ring), wave=wave, capacity=capacity, factor=factor)
        assert not res["short"], res["short"]
        ok, _, _, why = bench_cfg5.check_against_oracle(res, data_host, res["routed"][:index.numel()].cpu().numpy(), capacity=capacity)
        assert ok, why
        seen += res["own_arrival_bytes"]
        if capacity < 2000:
            assert res["compactions"] > 0 and res["rounds"] > 0
    assert seen == total + index.numel()  # every arrival belongs to exactly one rank


def test_stream_generator_matches_the_host_generator_layout():
    """The device generator's bytes parse as the run layout the rest of the repo uses (a handful of entries, host side)."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    from dbeel_b200 import sstable
    ids, tomb = cfg5.stream_ids(300, 50)
    tomb[:5] = [True, False, True, False, False]
    data, index, total = cfg5.build_stream_device(torch, torch.device("cuda:0"), ids, tomb, chunk=64)
    ents = sstable.parse_run(data.cpu().numpy(), index.cpu().numpy())
    assert len(ents) == 300
    for i, (k, v, ts) in enumerate(ents):
        assert k == b"\xb0" + (b"k%015d" % ids[i]) and ts == 1_700_000_000_000_000_000 + 1000 * i
        assert (v == b"") == bool(tomb[i]) and (len(v) in (0, cfg5.DOC_BYTES))
        if v:
            assert v[:3] == bytes([0xC5, (cfg5.DOC_BYTES - 3) >> 8, (cfg5.DOC_BYTES - 3) & 255])
