#!/usr/bin/env python
"""End-to-end (host pinned buffers in, host pinned buffers out) timing of dbeel_compact on the cfg2 workload for a
list of environment configurations, optionally with the per-partition timeline (DBEEL_TRACE=1).
Usage: tools/e2e_probe.py "DBEEL_PARTITION_MB=256" "DBEEL_PARTITION_MB=128,DBEEL_TRACE=1" ..."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from bench import SEED32, make_runs_parallel  # noqa: E402
from dbeel_b200 import capi  # noqa: E402
from dbeel_b200 import workloads as W  # noqa: E402


def main():
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
    pod, poi, pob = capi.PinnedBuffer(dc), capi.PinnedBuffer(ic), capi.PinnedBuffer(max(1, bc))
    ref = None
    for spec in sys.argv[1:] or [""]:
        env = dict(kv.split("=") for kv in spec.split(",") if kv)
        for k in [k for k in os.environ if k.startswith("DBEEL_")]:
            del os.environ[k]
        trace = env.pop("DBEEL_TRACE", None)
        os.environ.update(env)
        eng = capi.Engine(0)
        for _ in range(2):
            res = eng.compact(h_runs, cfg.keep_tombstones, seed=SEED32, out_buffers=(pod.array, poi.array, pob.array))
        steps = 5
        t0 = time.perf_counter()
        for _ in range(steps):
            res = eng.compact(h_runs, cfg.keep_tombstones, seed=SEED32, out_buffers=(pod.array, poi.array, pob.array))
        ms = (time.perf_counter() - t0) * 1e3 / steps
        st = eng.stats()
        n8 = res[0].size // 8 * 8
        chk = (res[3], int(res[0][:n8].view("<u8").sum()), int(res[1].view("<u8").sum()))
        if ref is None:
            ref = chk
        print(f"{spec or 'default':50s} {ms:7.2f} ms/step  {in_bytes / ms / 1e6:7.2f} GB/s in  partitions={st['partitions']} "
              f"kernels={st['ms_total']:.2f} ms  {'same-output' if chk == ref else 'OUTPUT-DIFFERS'}", flush=True)
        if trace:
            os.environ["DBEEL_TRACE"] = "1"
            eng.compact(h_runs, cfg.keep_tombstones, seed=SEED32, out_buffers=(pod.array, poi.array, pob.array))
            del os.environ["DBEEL_TRACE"]
        eng.close()


if __name__ == "__main__":
    main()
