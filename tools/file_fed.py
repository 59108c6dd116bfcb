#!/usr/bin/env python
"""Row N3: one headline job (8 runs x 1M keys x 256 B docs) file to file on tmpfs through dbeel_tree_compact --
streamed through the engine's pinned rings (dbeel_compact_stream) vs whole files through fresh pinned buffers.
    python tools/file_fed.py [--parity]      (env: DBEEL_IO_THREADS, DBEEL_STREAM_RING, DBEEL_PARTITION_MB, DBEEL_TRACE)
Prints one JSON line (bench.py's other_configs.file_fed carries the same record)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402
from dbeel_b200 import capi  # noqa: E402


def main():
    node, cpus = capi.bind_to_gpu(0) if hasattr(capi, "bind_to_gpu") else (None, None)
    expected = None
    if "--parity" in sys.argv:
        import oracle
        from dbeel_b200 import workloads as W
        cfg = W.cfg4_shard(0)
        expected = oracle.compact(bench.make_runs_parallel(cfg), cfg.keep_tombstones, seed=bench.SEED32, emulate_page_cache=True)
    eng = capi.Engine(0)
    res = bench.file_fed_job(eng, expected)
    res["env"] = {k: os.environ[k] for k in ("DBEEL_IO_THREADS", "DBEEL_STREAM_RING", "DBEEL_PARTITION_MB") if k in os.environ}
    res["host_placement"] = {"numa_node": node, "cpus": cpus}
    eng.close()
    print(json.dumps(res))


if __name__ == "__main__":
    main()
