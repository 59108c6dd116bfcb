#!/usr/bin/env python
"""Many small compactions (row N1): J level-0 merges of the cfg5 shape (8 runs x 8192 entries, 512-byte documents, one
memtable's worth per run) -- one dbeel_compact_device call per job vs one dbeel_compact_many_device launch sequence for
all of them, next to the CPU oracle.  Usage: tools/many_bench.py [jobs (default 32)]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

import oracle  # noqa: E402  (CPU baseline + spot check)
from dbeel_b200 import capi  # noqa: E402
from dbeel_b200 import workloads as W  # noqa: E402


def main():
    n_jobs = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    dev = torch.device("cuda:0")
    jobs, t_jobs, in_bytes = [], [], 0
    for j in range(n_jobs):
        cfg = W.MergeConfig(f"l0-merge-{j}", 8, 8192, 512, 40_000, 0.01, j % 2 == 0, 900 + j)
        runs = W.make_merge_runs(cfg)
        jobs.append((runs, cfg.keep_tombstones))
        t_jobs.append([(torch.from_numpy(d).to(dev), torch.from_numpy(i).to(dev)) for d, i in runs])
        in_bytes += sum(d.size + i.size for d, i in runs)
    seeds = [bytes([(j + k) % 256 for k in range(32)]) for j in range(n_jobs)]
    eng = capi.Engine(0)
    ptrs = [([(d.data_ptr(), d.numel(), i.data_ptr(), i.numel()) for d, i in tj], keep) for tj, (_, keep) in zip(t_jobs, jobs)]
    dc = sum(sum(d.numel() for d, _ in tj) for tj in t_jobs)
    ic = sum(sum(i.numel() for _, i in tj) for tj in t_jobs)
    bc = 16 * n_jobs + sum(capi.lib().dbeel_bloom_file_size(sum(i.numel() // 16 for _, i in tj), 0.01) for tj in t_jobs)
    od = torch.empty(dc + 64, dtype=torch.uint8, device=dev)
    oi = torch.empty(ic + 64, dtype=torch.uint8, device=dev)
    ob = torch.empty(bc + 64, dtype=torch.uint8, device=dev)
    torch.cuda.synchronize()

    def one_by_one():
        ms, items = 0.0, 0
        for (runs, keep), seed in zip(ptrs, seeds):
            opts = capi.make_opts(keep, seed=seed)
            res = eng.compact_device(runs, (od.data_ptr(), dc, oi.data_ptr(), ic, ob.data_ptr(), bc), opts)
            ms += eng.stats()["ms_total"]
            items += res[3]
        return ms, items

    one_by_one()
    t0 = time.perf_counter()
    ms_each, items_each = one_by_one()
    torch.cuda.synchronize()
    wall_each = time.perf_counter() - t0

    out_ptrs = (od.data_ptr(), dc, oi.data_ptr(), ic, ob.data_ptr(), bc)
    rows = eng.compact_many_device(ptrs, out_ptrs, seeds=seeds)
    t1 = time.perf_counter()
    rows = eng.compact_many_device(ptrs, out_ptrs, seeds=seeds)
    torch.cuda.synchronize()
    wall_many = time.perf_counter() - t1
    ms_many = eng.stats()["ms_total"]
    assert sum(r["items_written"] for r in rows) == items_each

    # spot check against the oracle + CPU time of one job
    r = rows[n_jobs // 2]
    t2 = time.perf_counter()
    xd, xi, xb, xn = oracle.compact(jobs[n_jobs // 2][0], jobs[n_jobs // 2][1], seed=seeds[n_jobs // 2], emulate_page_cache=True)
    cpu_one = time.perf_counter() - t2
    same = (np.array_equal(od[r["data_off"]:r["data_off"] + r["data_len"]].cpu().numpy(), xd)
            and np.array_equal(oi[r["index_off"]:r["index_off"] + r["index_len"]].cpu().numpy(), xi)
            and (xb is None) == (r["bloom_len"] == 0)
            and (xb is None or np.array_equal(ob[r["bloom_off"]:r["bloom_off"] + r["bloom_len"]].cpu().numpy(), xb)))
    print(f"{n_jobs} compactions of 8 runs x 8192 entries x 512 B documents: {in_bytes / 1e6:.0f} MB in, {items_each} entries out")
    print(f"GPU device-resident, one dbeel_compact_device per job : {ms_each:.2f} ms in kernels ({in_bytes / 1e6 / ms_each * 1e3:.0f} MB/s), "
          f"{wall_each * 1e3:.1f} ms wall ({in_bytes / 1e6 / wall_each:.0f} MB/s)")
    print(f"GPU device-resident, dbeel_compact_many_device (one launch sequence): {ms_many:.2f} ms in kernels "
          f"({in_bytes / 1e6 / ms_many * 1e3:.0f} MB/s), {wall_many * 1e3:.1f} ms wall ({in_bytes / 1e6 / wall_many:.0f} MB/s)")
    print(f"CPU oracle, 1 core: {cpu_one * 1e3:.0f} ms per job -> {cpu_one * n_jobs * 1e3:.0f} ms for all "
          f"({in_bytes / 1e6 / (cpu_one * n_jobs):.0f} MB/s); job {n_jobs // 2} of the batch byte-identical to it: {same}")
    eng.close()


if __name__ == "__main__":
    main()
