#!/usr/bin/env python
"""BASELINE.json configs[4] for one shard, device-resident: a Zipf(0.99) write stream (512-byte documents) is cut into
memtables of 8192 distinct keys; the memtables arrive in waves of 8, every wave is flushed by one
dbeel_flush_many_device, and after every wave the size-tiered picker (compaction factor 8) runs to quiescence -- every
round's groups merged by ONE dbeel_compact_many_device.  The same
recorded plan is replayed on the CPU oracle (red-black-tree memtables, one core) and every table that is left is
compared byte for byte.  Usage: tools/cfg5_bench.py [n_writes (default 1500000)]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

import oracle  # noqa: E402  (CPU arm + checker)
from dbeel_b200 import capi, storage_engine as se  # noqa: E402
from dbeel_b200 import workloads as W  # noqa: E402

FACTOR = 8


def main():
    n_writes = int(sys.argv[1]) if len(sys.argv) > 1 else 1_500_000
    batch = W.make_arrival_batch(n_writes=n_writes, n_ids=max(8192, n_writes // 4), doc_bytes=512, seed=5)
    nbytes = batch[0].size + batch[1].size
    cuts, pos = [], 0
    while pos < n_writes:
        n = se.memtable_cut(batch, pos, 8192)
        cuts.append((pos, n))
        pos += n
    idx = batch[1].view("<u8").reshape(-1, 2)[:, 0]
    dev = torch.device("cuda:0")
    eng = capi.Engine(0)
    d_data = torch.from_numpy(batch[0]).to(dev)
    subs = []
    for p0, n in cuts:  # each memtable's arrivals: a slice of the stream + an index whose offsets restart at 0
        lo = int(idx[p0])
        hi = int(idx[p0 + n]) if p0 + n < n_writes else batch[0].size
        ix = batch[1][16 * p0:16 * (p0 + n)].copy()
        ix.view("<u8").reshape(-1, 2)[:, 0] -= lo
        # flush batches need 16-byte aligned .data: copy the slice into its own buffer
        subs.append((d_data[lo:hi].clone(), torch.from_numpy(ix).to(dev)))
    del d_data
    torch.cuda.synchronize()

    # ---- GPU: memtables arrive in waves of FACTOR; each wave is flushed by one dbeel_flush_many_device, then the
    # picker runs to quiescence (as compact_tree does after flush events), every round's groups in one compact_many
    t0 = time.perf_counter()
    ms_kernels, launches = 0.0, 0
    tables = {}
    plan_log, rounds = [], 0
    for w0 in range(0, len(subs), FACTOR):
        wave = subs[w0:w0 + FACTOR]
        tot_d = sum(d.numel() for d, _ in wave)
        tot_i = sum(i.numel() for _, i in wave)
        fd = torch.empty(tot_d + 64, dtype=torch.uint8, device=dev)
        fi = torch.empty(tot_i + 64, dtype=torch.uint8, device=dev)
        _, _, _, rows = eng.flush_many_device([(d.data_ptr(), d.numel(), i.data_ptr(), i.numel()) for d, i in wave],
                                              (fd.data_ptr(), tot_d, fi.data_ptr(), tot_i))
        ms_kernels += eng.stats()["ms_total"]
        launches += eng.stats()["kernel_launches"]
        # a table = (data, index, entries[, bloom]); slices of fd / fi are 16-byte aligned only by luck, so every table is
        # cloned into its own allocation before it becomes a compaction input
        for m, r in enumerate(rows):
            tables[2 * (w0 + m)] = (fd[r["data_off"]:r["data_off"] + r["data_len"]].clone(),
                                    fi[r["index_off"]:r["index_off"] + r["index_len"]].clone(), r["items"])
        while True:
            plan = se.plan_compactions(sorted((i, t[2]) for i, t in tables.items()), FACTOR)
            if not plan:
                break
            rounds += 1
            jobs = [([(tables[i][0].data_ptr(), tables[i][0].numel(), tables[i][1].data_ptr(), tables[i][1].numel()) for i in indices], keep)
                    for indices, _, keep in plan]
            dc = sum(tables[i][0].numel() for indices, _, _ in plan for i in indices)
            ic = sum(tables[i][1].numel() for indices, _, _ in plan for i in indices)
            bc = sum(16 + capi.lib().dbeel_bloom_file_size(sum(tables[i][2] for i in indices), 0.01) for indices, _, _ in plan)
            od = torch.empty(dc + 64, dtype=torch.uint8, device=dev)
            oi = torch.empty(ic + 64, dtype=torch.uint8, device=dev)
            ob = torch.empty(bc + 64, dtype=torch.uint8, device=dev)
            seeds = [bytes([(g + rounds) % 256] * 32) for g in range(len(plan))]
            res = eng.compact_many_device(jobs, (od.data_ptr(), dc, oi.data_ptr(), ic, ob.data_ptr(), bc), seeds=seeds)
            ms_kernels += eng.stats()["ms_total"]
            launches += eng.stats()["kernel_launches"]
            for (indices, out_index, keep), r, seed in zip(plan, res, seeds):
                for i in indices:
                    del tables[i]
                tables[out_index] = (od[r["data_off"]:r["data_off"] + r["data_len"]].clone(),
                                     oi[r["index_off"]:r["index_off"] + r["index_len"]].clone(), r["items_written"],
                                     ob[r["bloom_off"]:r["bloom_off"] + r["bloom_len"]].clone() if r["bloom_len"] else None)
                plan_log.append((indices, out_index, keep, seed))
    torch.cuda.synchronize()
    gpu_wall = time.perf_counter() - t0

    # ---- CPU oracle: the same flushes and the same recorded plan, one core
    t1 = time.perf_counter()
    flushed = oracle.memtable_flushes(batch, capacity=8192, emulate_page_cache=True)
    otab = {2 * m: (d, i, n, None) for m, (d, i, n) in enumerate(flushed)}
    for indices, out_index, keep, seed in plan_log:
        d, i, b, n = oracle.compact([(otab[k][0], otab[k][1]) for k in indices], keep, seed=seed, emulate_page_cache=True)
        for k in indices:
            del otab[k]
        otab[out_index] = (d, i, n, b)
    cpu_wall = time.perf_counter() - t1

    same = sorted(tables) == sorted(otab)
    for k in sorted(tables):
        if not same:
            break
        g, o = tables[k], otab[k]
        same = same and g[2] == o[2] and np.array_equal(g[0].cpu().numpy(), o[0]) and np.array_equal(g[1].cpu().numpy(), o[1])
        gb = g[3] if len(g) > 3 else None
        same = same and (gb is None) == (o[3] is None) and (gb is None or np.array_equal(gb.cpu().numpy(), o[3]))
    print(f"cfg5, one shard: {n_writes} writes, {nbytes / 1e6:.0f} MB of arrivals -> {len(cuts)} memtables -> {len(plan_log)} compactions in "
          f"{rounds} picker rounds (factor {FACTOR}) -> {len(tables)} tables of {sorted(t[2] for t in tables.values())} entries")
    print(f"GPU, device-resident: {ms_kernels:.2f} ms in kernels over {launches} launches ({nbytes / 1e6 / ms_kernels * 1e3:.0f} MB/s of arrivals), "
          f"{gpu_wall * 1e3:.1f} ms wall incl. picker, allocations and table copies ({nbytes / 1e6 / gpu_wall:.0f} MB/s)")
    print(f"CPU oracle, 1 core  : {cpu_wall * 1e3:.0f} ms ({nbytes / 1e6 / cpu_wall:.0f} MB/s of arrivals)")
    print(f"every table left byte-identical to the oracle's (data, index, bloom): {same}")
    eng.close()
    if not same:
        sys.exit(1)


if __name__ == "__main__":
    main()
