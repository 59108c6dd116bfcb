#!/usr/bin/env python
"""End-to-end timing of dbeel_compact on the cfg2 workload with ONE engine (jobs one after the other) and with TWO engines
driven by two host threads (two shards' compactions in flight on one GPU: the fill of one overlaps the drain of the other).
Usage: tools/e2e_two_engines.py [jobs]"""
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from bench import SEED32, make_runs_parallel  # noqa: E402
from dbeel_b200 import capi  # noqa: E402
from dbeel_b200 import workloads as W  # noqa: E402


def main():
    n_jobs = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    capi.bind_to_gpu(0)
    cfg = W.CFG2
    runs = make_runs_parallel(cfg)
    torch.cuda.init()
    opts = capi.make_opts(cfg.keep_tombstones, seed=SEED32)
    dc, ic, bc = capi.compact_bound([(d.size, i.size) for d, i in runs], opts)
    pins, h_runs = [], []
    for d, i in runs:
        pd, pi = capi.PinnedBuffer(d.size), capi.PinnedBuffer(i.size)
        pd.array[:] = d
        pi.array[:] = i
        pins += [pd, pi]
        h_runs.append((pd.array, pi.array))
    in_bytes = sum(d.size + i.size for d, i in runs)
    outs = [(capi.PinnedBuffer(dc), capi.PinnedBuffer(ic), capi.PinnedBuffer(max(1, bc))) for _ in range(2)]
    for n_eng in (1, 2, 1, 2):
        engs = [capi.Engine(0) for _ in range(n_eng)]
        res = [None] * n_eng

        def worker(k, reps):
            for _ in range(reps):
                o = outs[k]
                res[k] = engs[k].compact(h_runs, cfg.keep_tombstones, seed=SEED32, out_buffers=(o[0].array, o[1].array, o[2].array))

        for k in range(n_eng):
            worker(k, 1)
        t0 = time.perf_counter()
        ths = [threading.Thread(target=worker, args=(k, n_jobs // n_eng)) for k in range(n_eng)]
        for t in ths:
            t.start()
        for t in ths:
            t.join()
        ms = (time.perf_counter() - t0) * 1e3 / n_jobs
        chk = [(r[3], int(r[0][: r[0].size // 8 * 8].view("<u8").sum())) for r in res]
        print(f"{n_eng} engine(s): {ms:7.2f} ms/job  {in_bytes / ms / 1e6:7.2f} GB/s in   {chk}", flush=True)
        for e in engs:
            e.close()


if __name__ == "__main__":
    main()
